"""CPU model of the kind::i8 filter scores (vecgo_b200/csrc/vg_quant_tc.cu, Q_SQ8I / Q_INT4I / Q_PQI) against the
certificate term that replaces c1 ||q|| max||x^ - mid|| when the query tile is quantised to signed 8-bit.

x^_p = mid_p + w_p b_p with the integer b_p (SQ8: code - 128, INT4: nibble - 8, PQ: the int8 codebook entry).  The
preparation kernel (prep_queries_i8_kernel) forms a_p = q_p w_p, Delta = max|a| / 127, ah_p = rint(a_p / Delta),
e_p = a_p - Delta ah_p, and the GEMM multiplies ah with the UNSIGNED stored value c_p = b_p + offset exactly (int32):
    s' = ||x^||^2 - 2 Delta acc,   acc = sum ah_p c_p,   c_q = -2 q.mid + 2 offset Delta sum ah_p
so that  s' + c_q  differs from the true  ||x^||^2 - 2 q.x^  by  2 sum e_p b_p  (+ float32 roundings), and
    |2 sum e_p b_p| <= 2 ||e / w||_2 ||x^ - mid||_2  =  ea * max||x^ - mid||        (Cauchy-Schwarz).
The model evaluates everything in float64 / exact integers and checks the inequality — the whole bound E the kernel uses,
and the new term alone — on several data scales and query shapes (Gaussian, one dominant component, constant dimensions).
"""
import numpy as np

F = np.float32


def prep_i8(q, w):
    a = (q * w).astype(F)
    mx = np.abs(a).max()
    delta = F(mx) / F(127.0) if mx > 0 else F(0.0)
    inv = F(127.0) / F(mx) if mx > 0 else F(0.0)
    ah = np.clip(np.rint((a * inv).astype(F)), -127, 127).astype(np.int64)
    e = a.astype(np.float64) - float(delta) * ah
    ew = np.where(w != 0, e / np.where(w != 0, w, 1).astype(np.float64), 0.0)
    ea = 2.0 * np.sqrt((ew * ew).sum()) * 1.000001
    return ah, float(delta), ea


def full_bound(q, xhat, mid, dim, ea, G=128):
    qn = np.linalg.norm(q.astype(np.float64))
    bn = np.linalg.norm((xhat - mid).astype(np.float64), axis=1).max()
    xx = (xhat.astype(np.float64) ** 2).sum(1).max()
    c2 = 1.0 / 16384.0 + dim / 8388608.0
    smax = xx + 2.0 * qn * bn
    return (ea * bn + c2 * (qn * qn + max(xx, bn * bn)) + smax * (1.0 / 4194304.0 + G / 8388608.0)
            + qn * (np.linalg.norm(mid.astype(np.float64)) + bn) / 2097152.0), ea * bn


def queries(rng, dim, scale, offset):
    yield (rng.standard_normal(dim) * scale + offset).astype(F)
    q = (rng.standard_normal(dim) * scale * 0.01 + offset).astype(F)
    q[rng.integers(dim)] += F(50.0 * scale)          # one dominant component: Delta is coarse for all the others
    yield q
    yield (rng.random(dim) * scale + offset).astype(F)  # all of one sign
    yield np.zeros(dim, F)


def check(codes_b, offset_c, w, mid, xhat, q, dim):
    ah, delta, ea = prep_i8(q, w)
    c = codes_b + offset_c                                   # what the kernel multiplies (unsigned for SQ8 / INT4)
    acc = c @ ah                                             # exact integers
    assert np.abs(acc).max() < 2 ** 31
    xn = (xhat.astype(np.float64) ** 2).sum(1).astype(F)
    s_model = xn.astype(np.float64) - 2.0 * delta * acc
    cq = -2.0 * float(q.astype(np.float64) @ mid.astype(np.float64)) + 2.0 * offset_c * delta * float(ah.sum())
    s_true = (xhat.astype(np.float64) ** 2).sum(1) - 2.0 * (xhat.astype(np.float64) @ q.astype(np.float64))
    err = np.abs(s_model + cq - s_true).max()
    E, term = full_bound(q, xhat, mid, dim, ea)
    assert err <= E, (err, E)
    # the quantisation part alone: 2 sum e_p b_p against ea * max||x^ - mid||
    e = (q * w).astype(F).astype(np.float64) - delta * ah
    qerr = np.abs(2.0 * (codes_b @ e)).max()
    assert qerr <= term * (1 + 1e-9) + 1e-300, (qerr, term)
    return err, E


def test_sq8_i8_filter_score_stays_inside_the_certificate_bound():
    rng = np.random.default_rng(31)
    dim, rows = 768, 2000
    for scale, offset in ((1.0, 0.0), (1e-3, 0.0), (250.0, 1000.0), (1.0, -7.0)):
        mins = (rng.standard_normal(dim) * 0.1 - 4.0).astype(F) * F(scale) + F(offset)
        inv = ((8.0 + rng.random(dim)) / 255.0).astype(F) * F(scale)
        inv[::97] = 0.0                                       # constant dimensions (SetBounds: diff < 1e-9 -> 0)
        codes = rng.integers(0, 256, (rows, dim)).astype(np.int64)
        codes[0], codes[1] = 0, 255
        mid = (mins.astype(np.float64) + 128.0 * inv.astype(np.float64)).astype(F)
        xhat = (codes * inv.astype(np.float64) + mins.astype(np.float64)).astype(F)
        for q in queries(rng, dim, scale, offset):
            check(codes - 128, 128, inv, mid, xhat, q, dim)


def test_int4_i8_filter_score_stays_inside_the_certificate_bound():
    rng = np.random.default_rng(32)
    dim, rows = 256, 2000
    for scale, offset in ((1.0, 0.0), (30.0, -100.0)):
        mn = (rng.standard_normal(dim) * 0.1 - 4.0).astype(F) * F(scale) + F(offset)
        diff = ((8.0 + rng.random(dim))).astype(F) * F(scale)
        w = (diff * F(1.0 / 15.0)).astype(F)
        nib = rng.integers(0, 16, (rows, dim)).astype(np.int64)
        nib[0], nib[1] = 0, 15
        mid = (mn.astype(np.float64) + 8.0 * w.astype(np.float64)).astype(F)
        xhat = (nib * w.astype(np.float64) + mn.astype(np.float64)).astype(F)
        for q in queries(rng, dim, scale, offset):
            check(nib - 8, 8, w, mid, xhat, q, dim)


def test_pq_i8_filter_score_stays_inside_the_certificate_bound():
    rng = np.random.default_rng(33)
    dim, m, rows = 256, 32, 1500
    ds = dim // m
    cb = rng.integers(-128, 128, (m, 256, ds)).astype(np.int64)
    sc = (0.01 + 0.002 * rng.random(m)).astype(F)
    of = (0.05 * rng.standard_normal(m)).astype(F)
    codes = rng.integers(0, 256, (rows, m))
    b = np.concatenate([cb[j, codes[:, j]] for j in range(m)], axis=1)        # [rows, dim] int8 entries
    w = np.repeat(sc, ds)
    mid = np.repeat(of, ds)
    xhat = (b * w.astype(np.float64) + mid.astype(np.float64)).astype(F)
    for q in queries(rng, dim, 1.0, 0.0):
        check(b, 0, w, mid, xhat, q, dim)
