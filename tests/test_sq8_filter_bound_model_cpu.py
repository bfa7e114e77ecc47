"""CPU model of the SQ8 decode-GEMM filter score (vecgo_b200/csrc/vg_quant_tc.cu) against its certificate bound E.

x^_d = mid_d + w_d b_d with b = code - 128 (exact in fp16), a_d = fp16(q_d w_d 2^e), f_q = -2 / 2^e, c_q = -2 q.mid:
    s' = f_q * sum_d a_d b_d + ||x^||^2        and        ||q - x^||^2 = s' + c_q + ||q||^2  up to E.
The model multiplies the rounded operands out in float64 (the tensor core's fp16 x fp16 products are exact) and checks
|s'_model + c_q - (||x^||^2 - 2 q.x^)| <= E for the constants the kernel uses, on several data scales.
"""
import numpy as np

F = np.float32


def bound(q, xhat, mid, dim, G=128):
    qn = np.linalg.norm(q.astype(np.float64))
    bn = np.linalg.norm((xhat - mid).astype(np.float64), axis=1).max()
    xx = (xhat.astype(np.float64) ** 2).sum(1).max()
    c1 = 1.125 / 1024.0
    c2 = 1.0 / 16384.0 + dim / 8388608.0
    smax = xx + 2.0 * qn * bn
    return (c1 * qn * bn + c2 * (qn * qn + max(xx, bn * bn)) + smax * (1.0 / 4194304.0 + G / 8388608.0)
            + qn * (np.linalg.norm(mid.astype(np.float64)) + bn) / 2097152.0)


def test_sq8_filter_score_stays_inside_the_certificate_bound():
    rng = np.random.default_rng(21)
    dim, rows = 768, 3000
    for scale, offset in ((1.0, 0.0), (1e-3, 0.0), (250.0, 1000.0), (1.0, -7.0)):
        mins = (rng.standard_normal(dim) * 0.1 - 4.0).astype(F) * F(scale) + F(offset)
        inv = ((8.0 + rng.random(dim)) / 255.0).astype(F) * F(scale)
        codes = rng.integers(0, 256, (rows, dim)).astype(np.uint8)
        codes[0] = 0
        codes[1] = 255
        for _ in range(4):
            q = (rng.standard_normal(dim) * scale + offset).astype(F)
            w = inv
            mid = (mins.astype(np.float64) + 128.0 * inv.astype(np.float64)).astype(F)
            # reference decode: rec = fma(code, inv, min) rounded once to float32
            xhat = (codes.astype(np.float64) * inv.astype(np.float64) + mins.astype(np.float64)).astype(F)
            qw = (q * w).astype(F)
            mx = np.abs(qw).max()
            e = 12 - np.frexp(mx)[1]
            a = (qw * F(2.0) ** e).astype(F).astype(np.float16)
            assert np.isfinite(a.astype(F)).all() and 2048 <= np.abs(a.astype(F)).max() <= 4096
            b = codes.astype(np.float64) - 128.0
            acc = b @ a.astype(np.float64)
            fq = -2.0 / 2.0 ** e
            xn = np.zeros(rows, F)
            v = xhat
            xn = (v.astype(np.float64) ** 2).sum(1).astype(F)          # any float32 evaluation of ||x^||^2 (error inside c2)
            s_model = fq * acc + xn.astype(np.float64)
            cq = -2.0 * float(q.astype(np.float64) @ mid.astype(np.float64))
            s_true = (xhat.astype(np.float64) ** 2).sum(1) - 2.0 * (xhat.astype(np.float64) @ q.astype(np.float64))
            err = np.abs(s_model + cq - s_true).max()
            E = bound(q, xhat, mid, dim)
            assert err <= E, (scale, offset, err, E)
            assert err <= 0.6 * E    # head-room for the fp32 accumulation inside the tensor core, not modelled here
