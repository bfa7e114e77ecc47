"""CPU model of the streaming group selection of the tensor-core filters (tc_select_kernel + select_compact_approx in
vecgo_b200/csrc/vg_flat_tc.cu): a bounded buffer with a threshold, compacted WITHOUT sorting by a sampled pivot.

Invariant checked: whatever the input order, the buffer always holds a superset of the kc smallest keys seen so far and
tau never drops below the kc-th smallest, so the final exact sort returns exactly the kc smallest keys (= the groups the
exact stage must score) — the same set the bitonic-sort compaction returned.
"""
import numpy as np

EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)


def orderable(f):
    u = (np.asarray(f, np.float32) + np.float32(0.0)).view(np.uint32).astype(np.uint64)
    return np.where(u & np.uint64(0x80000000), ~u & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))


def capacity(k, burst):
    need, c = k + 2 * burst, 64
    while c < need:
        c <<= 1
    return c


def compact_approx(buf, kc, C):
    """select_compact_approx: returns (kept keys, tau) or None when the kernel would fall back to the exact sort."""
    n = len(buf)
    if n <= kc:
        return buf, None
    sample = np.sort(np.array([buf[(lane * n) >> 5] for lane in range(32)], np.uint64))
    lo, hi, best, kept = 0, 31, -1, 0
    while lo <= hi:
        mid = (lo + hi) >> 1
        c = int(np.count_nonzero(buf < sample[mid]))
        if c >= kc:
            best, kept, hi = mid, c, mid - 1
        else:
            lo = mid + 1
    if best < 0 or kept > kc + (C - kc) // 2:
        return None
    pivot = sample[best]
    return buf[buf < pivot], pivot


def stream_select(values, kc):
    C = capacity(kc, 32)
    trigger = C - 32
    keys_all = (orderable(values) << np.uint64(32)) | np.arange(len(values), dtype=np.uint64)
    buf = np.zeros(0, np.uint64)
    tau = EMPTY
    exact_sorts = approx = 0
    for g0 in range(0, len(keys_all), 32):
        chunk = keys_all[g0:g0 + 32]
        buf = np.concatenate([buf, chunk[chunk < tau]])
        assert len(buf) <= C
        if len(buf) > trigger:
            r = compact_approx(buf, kc, C)
            if r is None:                                  # topk_compact_warp: exact sort, keep kc
                buf = np.sort(buf)[:kc]
                tau = buf[kc - 1] if len(buf) >= kc else EMPTY
                exact_sorts += 1
            else:
                buf, t = r
                if t is not None:
                    tau = t
                approx += 1
            # invariant: nothing that belongs to the kc smallest so far was dropped
            want = np.sort(keys_all[:g0 + 32])[:kc]
            assert np.all(np.isin(want, buf)), "a key of the current top-kc was dropped"
            assert len(buf) <= trigger, "compaction made no room"
    out = np.sort(buf)[:kc]
    return out, np.sort(keys_all)[:kc], approx, exact_sorts


def test_streaming_selection_with_sampled_pivots_returns_the_kc_smallest():
    rng = np.random.default_rng(8)
    cases = []
    for kc in (32, 200, 2000):
        n = 40 * kc + 17
        cases.append((rng.standard_normal(n).astype(np.float32), kc))                 # random
        cases.append((np.sort(rng.standard_normal(n)).astype(np.float32), kc))        # ascending: tau tight at once
        cases.append((np.sort(rng.standard_normal(n))[::-1].astype(np.float32).copy(), kc))  # descending: everything passes
        cases.append((np.zeros(n, np.float32), kc))                                   # all equal: order by group id only
        cases.append((np.round(rng.standard_normal(n) * 2).astype(np.float32), kc))   # few distinct values, massive ties
        cases.append((np.concatenate([np.full(n // 2, 3.0e38), rng.standard_normal(n - n // 2)]).astype(np.float32), kc))  # masked rows first
    total_approx = 0
    for values, kc in cases:
        got, want, approx, exact = stream_select(values, kc)
        assert np.array_equal(got, want)
        total_approx += approx
    assert total_approx > 0   # the sort-free path is what ran
