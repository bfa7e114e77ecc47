"""CPU model of the Flat fp16 filter score (flat2_kernel in vecgo_b200/csrc/vg_flat_tc.cu) against the certificate
bound of tc_exact_kernel: x16 = fp16(x 2^sx) (sx from the largest row norm), a16 = fp16(q 2^e) (e from max|q_i|),
s = f_q * sum a16 x16 + ||x||^2 (L2, f_q = -2 / 2^(e+sx)) or f_q * sum (dot, f_q = -1 / 2^(e+sx)).
The rounded operands are multiplied out in float64 (fp16 x fp16 products are exact on the tensor core)."""
import numpy as np

F = np.float32


def model_and_bound(x, q, is_dot, G=128):
    d = x.shape[1]
    xn64 = (x.astype(np.float64) ** 2).sum(1)
    xx = xn64.max()
    qq = float((q.astype(np.float64) ** 2).sum())
    sx = 12 - np.frexp(np.sqrt(xx))[1]
    e = 12 - np.frexp(np.abs(q).max())[1]
    x16 = (x * F(2.0) ** sx).astype(F).astype(np.float16)
    a16 = (q * F(2.0) ** e).astype(F).astype(np.float16)
    assert np.isfinite(x16.astype(F)).all() and np.isfinite(a16.astype(F)).all()
    acc = x16.astype(np.float64) @ a16.astype(np.float64)
    fq = -(1.0 if is_dot else 2.0) / 2.0 ** (e + sx)
    if is_dot:
        s_model = fq * acc
        s_true = -(x.astype(np.float64) @ q.astype(np.float64))
    else:
        s_model = fq * acc + xn64.astype(F).astype(np.float64)
        s_true = xn64 - 2.0 * (x.astype(np.float64) @ q.astype(np.float64))
    c1 = (1.0 / 512.0 if is_dot else 1.0 / 256.0) * 1.125 * 0.5
    c2 = 1.0 / 16384.0 + d / 8388608.0
    smax = np.sqrt(qq * xx) if is_dot else xx + 2.0 * np.sqrt(qq * xx)
    E = c1 * np.sqrt(qq * xx) + c2 * (qq + xx) + smax * G / 8388608.0
    return np.abs(s_model - s_true).max(), E


def test_flat_fp16_filter_score_stays_inside_the_certificate_bound():
    rng = np.random.default_rng(33)
    for dim in (128, 768):
        for gen in ("uniform", "gauss", "scaled", "sparse"):
            n = 2000
            if gen == "uniform":
                x, qs = rng.random((n, dim)), rng.random((4, dim))
            elif gen == "gauss":
                x, qs = rng.standard_normal((n, dim)), rng.standard_normal((4, dim))
            elif gen == "scaled":
                x, qs = rng.standard_normal((n, dim)) * 3e4 + 1e5, rng.standard_normal((4, dim)) * 1e-3
            else:   # a few large components, the rest tiny: fp16 subnormals in the shadow
                x = rng.standard_normal((n, dim)) * 1e-7
                x[:, :4] = rng.standard_normal((n, 4)) * 10.0
                qs = rng.standard_normal((4, dim))
            x, qs = x.astype(F), qs.astype(F)
            for q in qs:
                for is_dot in (False, True):
                    err, E = model_and_bound(x, q, is_dot)
                    assert err <= 0.6 * E, (dim, gen, is_dot, err, E)   # head-room for the in-MMA fp32 accumulation
