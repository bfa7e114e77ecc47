"""CPU model of the operand encoding of the tensor-core PQ-training assignment (vecgo_b200/csrc/vg_pq_assign_tc.cu).

The kernel's exactness argument rests on two numeric claims that do not need a GPU to check:
  1. the hi/lo fp16 split of samples and centroids plus the three-way split of -|c|^2/2, multiplied out exactly (as the
     tensor core does for fp16 x fp16) reproduces x.c - |c|^2/2 to well below 2^-20 * B, B = max|x|^2 + max|c|^2;
  2. the epilogue's indicator sat((v - lim) * 2^60) is exactly 0 or 1 for float32 scores, and exactly 1 for the row
     maximum, so "sum of indicators == 1.0" means exactly one column within the margin.
"""
import numpy as np

F = np.float32
H = np.float16


def split16(v):
    hi = v.astype(H)
    lo = (v.astype(F) - hi.astype(F)).astype(H)
    return hi, lo


def model_scores(x, c):
    """x [n,8], c [k,8] already scaled (max|x| in [8,16)).  Returns (tensor-core model score, true score) in float64."""
    xh, xl = split16(x)
    ch, cl = split16(c)
    cn = np.zeros(len(c), F)
    for i in range(8):  # cn = fma(v, v, cn) in float32, as centroid_kernel does
        cn = (cn.astype(np.float64) + c[:, i].astype(np.float64) ** 2).astype(F)
    v = (F(-0.5) * cn).astype(F)
    h = v.astype(H)
    r1 = (v - h.astype(F)).astype(F)
    m = r1.astype(H)
    l = (r1 - m.astype(F)).astype(F).astype(H)
    d = np.float64
    acc = xh.astype(d) @ ch.astype(d).T + xl.astype(d) @ ch.astype(d).T + xh.astype(d) @ cl.astype(d).T
    acc = acc + (h.astype(d) + m.astype(d) + l.astype(d))[None, :]
    true = x.astype(d) @ c.astype(d).T - 0.5 * (c.astype(d) ** 2).sum(1)[None, :]
    return acc, true


def scaled(rng, n, kind):
    if kind == "gauss":
        x = rng.standard_normal((n, 8))
    elif kind == "wide":      # six decades between dimensions and a far offset
        x = rng.standard_normal((n, 8)) * np.logspace(-3, 3, 8)[None, :] + 100.0
    else:                     # tiny values next to large ones
        x = rng.standard_normal((n, 8)) * rng.choice([1e-6, 1.0], size=(n, 8))
    x = x.astype(F)
    mx = np.abs(x).max()
    e = 4 - (np.frexp(mx)[1])          # max|x| * 2^e in [8, 16)
    return (x * F(2.0) ** e).astype(F)


def test_split_score_error_is_far_below_the_certificate_margin():
    rng = np.random.default_rng(3)
    for kind in ("gauss", "wide", "tiny"):
        x = scaled(rng, 4000, kind)
        assert 8.0 <= np.abs(x).max() < 16.0
        c = x[rng.choice(len(x), 256, replace=False)] * F(0.75) + x[rng.choice(len(x), 256, replace=False)] * F(0.25)  # convex mixes
        acc, true = model_scores(x, c.astype(F))
        B = float((x.astype(np.float64) ** 2).sum(1).max() + (c.astype(np.float64) ** 2).sum(1).max())
        err = np.abs(acc - true).max()
        assert err <= B * 2.0 ** -21, (kind, err / B)          # kernel budget: 2^-20 B in total (fp32 accumulation included)
        assert B * 2.0 ** -16 >= 16 * 2 * err                  # the margin used by the certificate is >= 16x twice the error


def test_indicator_is_zero_or_one_and_one_for_the_maximum():
    rng = np.random.default_rng(4)
    big = F(2.0 ** 60)
    for _ in range(200):
        v = (rng.standard_normal(256) * rng.choice([1e-3, 1.0, 500.0])).astype(F)
        v[rng.integers(0, 256, 5)] = v[0]                        # exact duplicates of one score
        thr = F(abs(rng.standard_normal()) * 1e-2 + 1e-3)
        mx = v.max()
        nlim = ((thr - mx).astype(F) * big).astype(F)
        ind = np.clip((v.astype(np.float64) * float(big) + float(nlim)), 0.0, 1.0)   # FFMA.SAT: one rounding after an exact fma
        ind = ind.astype(F)
        assert set(np.unique(ind)) <= {F(0.0), F(1.0)}
        assert np.all(ind[v == mx] == 1.0)
        lim = float(mx) - float(thr)
        slack = abs(lim) * 2.0 ** -22 + 1e-30
        assert np.all(ind[v.astype(np.float64) > lim + slack] == 1.0) and np.all(ind[v.astype(np.float64) < lim - slack] == 0.0)
