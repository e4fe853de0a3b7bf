"""Tie-aware comparison of a device top-k with the oracle.

The oracle consumes the SAME stored (rounded) rows the device scored, so the only
difference left is fp32 accumulation order (~1e-7).  Rules checked:
  1. every device score is within `score_tol` of the exact (fp64) dot of the stored rows;
  2. the device list is best-first under its own scores, ties by lower row;
  3. the device row set equals the oracle's, except rows whose exact score lies within
     `tie_eps` of the oracle's k-th score (a boundary near-tie an fp32 reorder can flip);
  4. exact duplicates (bit-identical stored rows) appear lower row first.
Returns the number of boundary near-ties it had to excuse.
"""
import numpy as np


def exact_scores(d_stored: np.ndarray, q_stored: np.ndarray) -> np.ndarray:
    return np.atleast_2d(q_stored).astype(np.float64) @ d_stored.astype(np.float64).T


def assert_topk_matches(gpu_s, gpu_i, d_stored, q_stored, k, score_tol=2e-6, tie_eps=1e-6,
                        idx_offset=0, s64=None):
    gpu_s = np.asarray(gpu_s)
    gpu_i = np.asarray(gpu_i)
    if s64 is None:
        s64 = exact_scores(d_stored, q_stored)
    b, n = s64.shape
    kk = min(k, n)
    assert gpu_s.shape == (b, k) and gpu_i.shape == (b, k)
    excused = 0
    for r in range(b):
        row = s64[r]
        order = np.lexsort((np.arange(n), -row))[:kk]
        gi = gpu_i[r, :kk] - idx_offset
        gs = gpu_s[r, :kk]
        assert (gpu_i[r, kk:] == -1).all() and np.isneginf(gpu_s[r, kk:]).all(), "empty slots"
        assert ((gi >= 0) & (gi < n)).all(), f"query {r}: row out of range {gi}"
        assert len(set(gi.tolist())) == kk, f"query {r}: duplicate rows {gi}"
        # 1. scores
        err = np.abs(gs.astype(np.float64) - row[gi]).max() if kk else 0.0
        assert err <= score_tol, f"query {r}: score error {err}"
        # 2. own order
        for a in range(kk - 1):
            assert gs[a] > gs[a + 1] or (gs[a] == gs[a + 1] and gi[a] < gi[a + 1]), \
                f"query {r}: not best-first at {a}: {gs[a]},{gi[a]} then {gs[a+1]},{gi[a+1]}"
        # 3. set
        want = set(order.tolist())
        got = set(gi.tolist())
        if want != got:
            kth = row[order[-1]]
            for x in (want ^ got):
                assert abs(row[x] - kth) <= tie_eps, \
                    f"query {r}: row {x} (score {row[x]}) differs from oracle set, k-th {kth}"
            excused += len(want ^ got) // 2
    return excused
