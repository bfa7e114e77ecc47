"""CPU model of the BQ (Hamming) path through the group-minima filter (vecgo_b200/csrc/vg_quant_tc.cu, qtc2_kernel<Q_BQ>
+ tc_select_kernel + qtc_exact_kernel): soundness of the strict integer certificate.

Scores are integers with massive ties and ties are decided by row id, so the filter may only be trusted when the k-th
exact distance is STRICTLY below the Hamming part of tau.  Property: whenever the certificate holds, the top-k of the
scored candidates equals the top-k of the full scan under (distance, row); when it does not hold the library re-runs the
query on the exact scan, so nothing is claimed.
"""
import numpy as np


def full_topk(h, k):
    order = np.lexsort((np.arange(len(h)), h))[:k]
    return order


def filter_topk(h, k, kc, G):
    n = len(h)
    groups = (n + G - 1) // G
    pad = np.full(groups * G, 1 << 30, np.int64)
    pad[:n] = h
    blk = pad.reshape(groups, G)
    arg = blk.argmin(1)                                   # first arg-min: (value, index bits) minimum of the epilogue
    m1 = blk[np.arange(groups), arg]
    part = np.partition(blk, 1, axis=1)
    m2 = part[:, 1]                                       # second smallest value of the group (duplicates count)
    sel = np.lexsort((np.arange(groups), arg, m1))[:kc]   # key order of the selection: (value with index bits, group id)
    tau_h = m1[sel[-1]] if len(sel) >= kc else (1 << 30)
    rows = []
    for g in sel:
        if m2[g] < tau_h:                                 # crowded (the kernel's test m2 <= tau also takes some equal ones)
            rows.extend(range(g * G, min(n, (g + 1) * G)))
        else:
            r = g * G + arg[g]
            if r < n:
                rows.append(r)
    rows = np.array(sorted(set(rows)), np.int64)
    order = np.lexsort((rows, h[rows]))[:k]
    got = rows[order]
    certified = len(got) >= k and h[got[k - 1]] < tau_h
    return got, certified


def test_strict_integer_certificate_is_sound():
    rng = np.random.default_rng(12)
    certified_cases = failed_cases = 0
    for trial in range(300):
        D = int(rng.choice([64, 128, 256, 1536]))
        n = int(rng.integers(9000, 30000))
        k = int(rng.choice([1, 10, 100]))
        kc = 32 if k <= 16 else 2 * k
        G = int(rng.choice([32, 64, 128]))
        # Hamming distances of random codes ~ Binomial(D, 1/2); some trials plant many duplicates of the best rows
        h = rng.binomial(D, 0.5, n).astype(np.int64)
        if trial % 3 == 0:
            h[rng.integers(0, n, 400)] = h.min()
        if trial % 5 == 0:
            h[rng.integers(0, n, 50)] = 0
        got, certified = filter_topk(h, k, kc, G)
        if certified:
            certified_cases += 1
            assert np.array_equal(got, full_topk(h, k)), (trial, D, n, k, G)
        else:
            failed_cases += 1
    assert certified_cases > 50 and failed_cases > 10   # both branches were exercised
