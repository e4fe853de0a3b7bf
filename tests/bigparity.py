"""Oracle comparison for shards too large for a dense fp64 score matrix (BASELINE config sizes).

The oracle side is `oracle.topk_cosine`'s arithmetic -- fp32 `Q @ D.T` of the STORED rows upcast
to fp32 -- evaluated block by block over the WHOLE shard on the CPU (numpy /
BLAS), keeping per query every row whose fp32 score is within `margin` of the running k-th best;
those candidates (a superset of the exact top-k: fp32 BLAS error is ~1e-7, margin 1e-5) and every
row the device returned are then rescored in fp64, and the device lists are held to the same four
rules as `parity.assert_topk_matches` (scores, own order, row set up to boundary near-ties, ties
by lower row).  torch is used only to move stored rows off the GPU and to upcast them (exact).
"""
import numpy as np
import torch


def stored_block_f32(shard: torch.Tensor, dtype: str, lo: int, hi: int) -> np.ndarray:
    """Stored rows [lo, hi) upcast to fp32 on the host (exact for every storage class)."""
    blk = shard[lo:hi]
    if dtype == "bf16x2":
        d = blk.shape[1] // 2
        blk = blk[:, :d].float() + blk[:, d:].float()              # hi + lo, exact in fp32
    else:
        blk = blk.float()
    return blk.cpu().numpy()


def stored_rows_f64(shard: torch.Tensor, dtype: str, rows: np.ndarray) -> np.ndarray:
    idx = torch.from_numpy(np.asarray(rows, dtype=np.int64)).to(shard.device)
    blk = shard[idx]
    if dtype == "bf16x2":
        d = blk.shape[1] // 2
        blk = blk[:, :d].double() + blk[:, d:].double()
    else:
        blk = blk.double()
    return blk.cpu().numpy()


def oracle_candidates(shard: torch.Tensor, dtype: str, n: int, q_st: np.ndarray, k: int,
                      block: int = 250_000, margin: float = 1e-5):
    """Per query: rows whose fp32 oracle score is within `margin` of the oracle's k-th best
    (every row of the shard is scored).  Returns a list of int64 arrays."""
    q = np.ascontiguousarray(q_st, dtype=np.float32)
    b = q.shape[0]
    tau = np.full((b,), -np.inf, dtype=np.float32)
    pool_i = [[] for _ in range(b)]
    pool_s = [[] for _ in range(b)]
    for lo in range(0, n, block):
        hi = min(n, lo + block)
        s = q @ stored_block_f32(shard, dtype, lo, hi).T            # [b, hi-lo] fp32: the oracle's scores
        m = hi - lo
        if m >= k:                                                  # the block's own k-th best bounds the global one
            kth = np.partition(s, m - k, axis=1)[:, m - k]
            tau = np.maximum(tau, kth)
        rr, cc = np.nonzero(s >= (tau - np.float32(margin))[:, None])
        for r in np.unique(rr):
            sel = cc[rr == r]
            pool_i[r].append(sel.astype(np.int64) + lo)
            pool_s[r].append(s[r, sel])
    out = []
    for r in range(b):
        i = np.concatenate(pool_i[r]) if pool_i[r] else np.empty((0,), dtype=np.int64)
        sc = np.concatenate(pool_s[r]) if pool_s[r] else np.empty((0,), dtype=np.float32)
        if len(i) > k:                                              # tighten with the final k-th best
            kth = np.partition(sc, len(sc) - k)[len(sc) - k]
            i = i[sc >= kth - np.float32(margin)]
        out.append(i)
    return out


def assert_topk_matches_at_size(gpu_s, gpu_i, shard: torch.Tensor, dtype: str, n: int,
                                q_st: np.ndarray, k: int, score_tol: float, tie_eps: float,
                                idx_offset: int = 0, cands=None):
    """Returns (excused boundary near-ties, worst |device score - exact fp64 score|)."""
    gpu_s = np.asarray(gpu_s)
    gpu_i = np.asarray(gpu_i) - idx_offset
    b = q_st.shape[0]
    assert gpu_s.shape == (b, k) and gpu_i.shape == (b, k)
    if cands is None:
        cands = oracle_candidates(shard, dtype, n, q_st, k)
    q64 = q_st.astype(np.float64)
    kk = min(k, n)
    excused, worst = 0, 0.0
    for r in range(b):
        gi, gs = gpu_i[r, :kk], gpu_s[r, :kk]
        assert (gpu_i[r, kk:] + idx_offset == -1).all() and np.isneginf(gpu_s[r, kk:]).all(), "empty slots"
        assert ((gi >= 0) & (gi < n)).all(), f"query {r}: row out of range {gi}"
        assert len(set(gi.tolist())) == kk, f"query {r}: duplicate rows {gi}"
        rows = np.unique(np.concatenate([cands[r], gi]))
        s64 = stored_rows_f64(shard, dtype, rows) @ q64[r]
        exact = dict(zip(rows.tolist(), s64.tolist()))
        err = max(abs(float(gs[a]) - exact[int(gi[a])]) for a in range(kk)) if kk else 0.0
        worst = max(worst, err)
        assert err <= score_tol, f"query {r}: score error {err}"
        for a in range(kk - 1):
            assert gs[a] > gs[a + 1] or (gs[a] == gs[a + 1] and gi[a] < gi[a + 1]), \
                f"query {r}: not best-first at {a}: {gs[a]},{gi[a]} then {gs[a + 1]},{gi[a + 1]}"
        order = np.lexsort((rows, -s64))[:kk]
        want = set(rows[order].tolist())
        got = set(gi.tolist())
        if want != got:
            kth = s64[order[-1]]
            for x in (want ^ got):
                assert abs(exact[x] - kth) <= tie_eps, \
                    f"query {r}: row {x} (score {exact[x]}) differs from the oracle set, k-th {kth}"
            excused += len(want ^ got) // 2
    return excused, worst
