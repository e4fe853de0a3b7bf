"""numpy restatement of the reference's similarity path.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py for who may import this and how it is pinned).

Every function cites the reference lines it follows; paths are relative to
/root/reference.
"""
from __future__ import annotations

import json
from typing import List, Optional, Sequence, Tuple

import numpy as np

# app/main.py:38,42,44
EMBED_DIM = 1024
REDIS_MAX_ITEMS = 1000
CACHE_SIM_THRESHOLD = 0.96


# ----------------------------------------------------------------------------
# a1  cosine_similarity                                     app/main.py:59-64
# ----------------------------------------------------------------------------
def cosine_similarity(a: np.ndarray, b: np.ndarray) -> float:
    """dot(a,b)/(|a|*|b|) with the zero-norm guard -> 0.0 (main.py:59-64)."""
    norm_a = np.linalg.norm(a)
    norm_b = np.linalg.norm(b)
    if norm_a == 0.0 or norm_b == 0.0:
        return 0.0
    return float(np.dot(a, b) / (norm_a * norm_b))


# ----------------------------------------------------------------------------
# a4/a5/a6  row L2 normalise      app/main.py:315-316, :353-354; embedding_gen.py:215-216
# ----------------------------------------------------------------------------
def normalize_rows(emb: np.ndarray) -> np.ndarray:
    """`E / (||E||_row + 1e-9)` exactly as the reference writes it.

    fp32 in, fp32 out.  Zero rows stay exactly zero (0 / 1e-9)."""
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    norms = np.linalg.norm(emb, axis=1, keepdims=True)
    return emb / (norms + 1e-9)


def pairwise_sumsq_f32(emb: np.ndarray) -> np.ndarray:
    """Sum of squares per row in the exact fp32 order numpy uses.

    `np.linalg.norm(E, axis=1)` is `sqrt(add.reduce(E*E, axis=1))`; for a
    contiguous fp32 row numpy's add.reduce is its pairwise summation: blocks of
    128 elements, each block summed with 8 strided accumulators
    r[j] += x[8i+j] combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), blocks
    combined by recursive halving.  The CUDA ingest kernel reproduces this order
    so its fp32 output is bit-identical to `normalize_rows`; this function is the
    explicit statement of that order (checked against numpy in the tests).
    Only dims that are a power-of-two multiple of 128 are supported here."""
    x = np.ascontiguousarray(emb, dtype=np.float32)
    n, d = x.shape
    assert d % 128 == 0 and (d // 128) & (d // 128 - 1) == 0
    s = (x * x).reshape(n, d // 128, 16, 8)
    r = s[:, :, 0, :].copy()
    for i in range(1, 16):
        r = r + s[:, :, i, :]
    blk = ((r[..., 0] + r[..., 1]) + (r[..., 2] + r[..., 3])) + (
        (r[..., 4] + r[..., 5]) + (r[..., 6] + r[..., 7])
    )
    while blk.shape[1] > 1:
        blk = blk[:, 0::2] + blk[:, 1::2]
    return blk[:, 0]


# ----------------------------------------------------------------------------
# storage rounding (new capability; the reference stores fp32 JSON lists)
# ----------------------------------------------------------------------------
def to_storage(x: np.ndarray, dtype: str) -> np.ndarray:
    """Round fp32 values to the shard storage type, round-to-nearest-even.

    Returns the raw stored array: float32 for "fp32", float16 for "fp16",
    uint16 bit patterns for "bf16" (numpy has no bfloat16)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if dtype == "fp32":
        return x.copy()
    if dtype == "fp16":
        return x.astype(np.float16)
    if dtype == "bf16":
        u = x.view(np.uint32)
        nan = (u & 0x7FFFFFFF) > 0x7F800000
        rounded = (u + (0x7FFF + ((u >> 16) & 1))) >> 16
        rounded = np.where(nan, (u >> 16) | 0x0040, rounded)
        return rounded.astype(np.uint16)
    if dtype == "bf16x2":
        # split bf16 (new storage class): hi = bf16(x), lo = bf16(x - hi), row = [hi | lo]
        hi = to_storage(x, "bf16")
        rest = x - from_storage(hi, "bf16")
        lo = to_storage(rest.astype(np.float32), "bf16")
        return np.concatenate([hi, lo], axis=-1)
    raise ValueError(dtype)


def from_storage(s: np.ndarray, dtype: str) -> np.ndarray:
    """Upcast stored values to fp32 (exact)."""
    if dtype == "fp32":
        return np.asarray(s, dtype=np.float32)
    if dtype == "fp16":
        return np.asarray(s, dtype=np.float16).astype(np.float32)
    if dtype == "bf16":
        return (np.asarray(s, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)
    if dtype == "bf16x2":
        s = np.asarray(s, dtype=np.uint16)
        d = s.shape[-1] // 2
        return from_storage(s[..., :d], "bf16") + from_storage(s[..., d:], "bf16")   # exact in fp32
    raise ValueError(dtype)


# ----------------------------------------------------------------------------
# a6  exact scoring + top-k (replaces the external HNSW leg, main.py:356-367)
# ----------------------------------------------------------------------------
def topk_from_scores(scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Best-first top-k of each row of `scores`; ties -> lower index.

    Identical to `np.argsort(-s, kind="stable")[:k]` (the stable sort is what
    makes "first maximum wins", main.py:84, hold for k>1) but uses a partition
    so it stays fast at N = 1M."""
    s = np.atleast_2d(np.asarray(scores))
    b, n = s.shape
    kk = min(k, n)
    out_s = np.full((b, k), -np.inf, dtype=s.dtype)
    out_i = np.full((b, k), -1, dtype=np.int64)
    if kk == 0:
        return out_s, out_i
    for r in range(b):
        row = s[r]
        if n > 4 * kk + 64:
            kth = np.partition(row, n - kk)[n - kk]
            cand = np.nonzero(row >= kth)[0]
        else:
            cand = np.arange(n)
        order = np.argsort(-row[cand], kind="stable")[:kk]
        idx = cand[order]
        out_i[r, :kk] = idx
        out_s[r, :kk] = row[idx]
    return out_s, out_i


def topk_cosine(d_stored: np.ndarray, q_stored: np.ndarray, k: int,
                chunk: int = 262144) -> Tuple[np.ndarray, np.ndarray]:
    """Exact cosine top-k of stored (already normalised, already rounded) rows.

    `d_stored` [N,dim] and `q_stored` [B,dim] are fp32 upcasts of what sits in
    HBM, so the only difference left against the device is fp32 accumulation
    order.  Scores are fp32 `Q @ D.T` -- the vectorised form of
    main.py:59-64 on unit rows -- followed by the stable top-k above."""
    d = np.ascontiguousarray(d_stored, dtype=np.float32)
    q = np.atleast_2d(np.ascontiguousarray(q_stored, dtype=np.float32))
    n = d.shape[0]
    b = q.shape[0]
    if n <= chunk:
        return topk_from_scores(q @ d.T, k)
    best_s = np.empty((b, 0), dtype=np.float32)
    best_i = np.empty((b, 0), dtype=np.int64)
    for lo in range(0, n, chunk):
        s, i = topk_from_scores(q @ d[lo:lo + chunk].T, k)
        i = np.where(i >= 0, i + lo, -1)
        cs = np.concatenate([best_s, s], axis=1)
        ci = np.concatenate([best_i, i], axis=1)
        best_s, best_i = merge_topk(cs, ci, k, presplit=False)
    return best_s, best_i


def merge_topk(scores: np.ndarray, idx: np.ndarray, k: int, presplit: bool = True
               ) -> Tuple[np.ndarray, np.ndarray]:
    """Merge candidate lists into one best-first top-k per query.

    presplit=True : scores/idx are [lists, B, k'] (one list per shard).
    presplit=False: scores/idx are [B, M] flat candidate rows.
    Order is (score desc, index asc); idx < 0 marks an empty slot."""
    s = np.asarray(scores)
    i = np.asarray(idx)
    if presplit:
        s = np.concatenate(list(s), axis=1)
        i = np.concatenate(list(i), axis=1)
    b = s.shape[0]
    out_s = np.full((b, k), -np.inf, dtype=s.dtype)
    out_i = np.full((b, k), -1, dtype=np.int64)
    for r in range(b):
        valid = np.nonzero(i[r] >= 0)[0]
        order = np.lexsort((i[r][valid], -s[r][valid]))[:k]
        sel = valid[order]
        out_s[r, :len(sel)] = s[r][sel]
        out_i[r, :len(sel)] = i[r][sel]
    return out_s, out_i


def opensearch_score(cos: np.ndarray) -> np.ndarray:
    """[external] OpenSearch k-NN `cosinesimil` _score = 1/(1+d), d = 1-cos.
    Monotone in cosine; offered so results can look like main.py:364-367's."""
    return 1.0 / (2.0 - np.asarray(cos, dtype=np.float64))


# ----------------------------------------------------------------------------
# a2  cache lookup: top-1 + threshold                       app/main.py:73-90
# ----------------------------------------------------------------------------
def cache_lookup(query_vec: np.ndarray, cache_embs: Sequence[np.ndarray],
                 threshold: float = CACHE_SIM_THRESHOLD) -> Tuple[int, float, bool]:
    """Row-by-row restatement of the scan in lfu_cache_get (main.py:73-90).

    Returns (best_index, best_sim, hit).  Strict `>` keeps the first maximum
    (lowest list index); the running maximum starts at -1.0 with index -1, so a
    cache whose every similarity is <= -1.0 never hits; miss iff
    best_sim < threshold."""
    best_sim = -1.0
    best_index = -1
    for i, emb in enumerate(cache_embs):
        sim = cosine_similarity(query_vec, np.asarray(emb, dtype=np.float32))
        if sim > best_sim:
            best_sim = sim
            best_index = i
    hit = not (best_sim < threshold) and best_index >= 0
    return best_index, best_sim, hit


def cache_lookup_batched(q_stored: np.ndarray, c_stored: np.ndarray, threshold: float
                         ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Vectorised form of `cache_lookup` for stored unit rows (config 5 sizes).
    Returns (idx int32 [B], score f32 [B], hit u8 [B])."""
    s, i = topk_cosine(c_stored, q_stored, 1)
    s = s[:, 0]
    i = i[:, 0]
    valid = (i >= 0) & (s > np.float32(-1.0))
    hit = valid & ~(s.astype(np.float64) < float(threshold))     # Python-float compare, main.py:89
    return np.where(valid, i, -1).astype(np.int32), s.astype(np.float32), hit.astype(np.uint8)


# ----------------------------------------------------------------------------
# a2/a3  the Redis list + LFU eviction, as a plain Python model   main.py:67-128
# ----------------------------------------------------------------------------
class LfuCacheModel:
    """In-memory model of the reference's Redis-list cache.

    `items` mirrors the Redis list `query_cache_lfu`: index 0 is the newest entry
    (LPUSH, main.py:128); each item is the JSON string the reference would store
    (main.py:123)."""

    def __init__(self, max_items: int = REDIS_MAX_ITEMS,
                 threshold: float = CACHE_SIM_THRESHOLD):
        self.max_items = max_items
        self.threshold = threshold
        self.items: List[str] = []

    def get(self, query_emb: np.ndarray) -> Optional[str]:
        if not self.items:                                    # main.py:70-71
            return None
        entries = [json.loads(s) for s in self.items]
        embs = [np.array(e["embedding"], dtype=np.float32) for e in entries]
        idx, _sim, hit = cache_lookup(query_emb[0], embs, self.threshold)
        if not hit:                                           # main.py:89-90
            return None
        e = entries[idx]
        e["freq"] = e.get("freq", 1) + 1                      # main.py:94
        self.items[idx] = json.dumps(e)                       # main.py:95
        return e["response"]

    def _remove_least_frequent_item(self) -> None:            # main.py:101-118
        if not self.items:
            return
        min_freq = float("inf")
        min_index = -1
        for i, s in enumerate(self.items):
            f = json.loads(s).get("freq", 1)
            if f < min_freq:
                min_freq = f
                min_index = i
        if min_index >= 0:
            # LREM count=1 removes the first element EQUAL to that string, which
            # may sit before min_index if an identical JSON string exists.
            self.items.remove(self.items[min_index])

    def put(self, query_emb: np.ndarray, response: str) -> None:   # main.py:121-128
        entry = {"embedding": query_emb.tolist()[0], "response": response, "freq": 1}
        if len(self.items) >= self.max_items:
            self._remove_least_frequent_item()
        self.items.insert(0, json.dumps(entry))

    def freqs(self) -> List[int]:
        return [json.loads(s).get("freq", 1) for s in self.items]

    def responses(self) -> List[str]:
        return [json.loads(s)["response"] for s in self.items]


# ----------------------------------------------------------------------------
# K3p  int8 prefilter: the bound that makes it exact (no reference line: the reference scores
# every row, main.py:59-64; this is the checker of the device-side error bound)
# ----------------------------------------------------------------------------
PREFILTER_SLACK = np.float32(4e-6)
PREFILTER_INFLATE = np.float32(1.001)


def quantize_rows_int8(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Per-row max-abs int8 quantisation as `sqe_quantize_rows` does it (fp32 arithmetic):
    d8 = clip(rint(x * (127 / max|x|))), meta = [sd, eps, nd, 0] with sd = max|x| / 127,
    eps >= |x - sd d8|_2, nd >= |sd d8|_2.  Non-finite rows: d8 = 0, eps = +inf."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = x.shape[0]
    with np.errstate(all="ignore"):
        finite = np.isfinite((x.astype(np.float64) ** 2).sum(axis=1).astype(np.float32))
        mx = np.nanmax(np.abs(np.where(np.isnan(x), np.float32(0), x)), axis=1).astype(np.float32)
        live = finite & (mx > 0)
        sd = np.where(live, mx / np.float32(127), np.float32(0)).astype(np.float32)
        inv = np.where(live, np.float32(127) / np.where(live, mx, np.float32(1)), np.float32(0)).astype(np.float32)
        xs = np.where(live[:, None], x, np.float32(0))
        r = np.clip(np.rint(xs * inv[:, None]), -127, 127).astype(np.float32)
        back = (sd[:, None] * r).astype(np.float64)
        err = xs.astype(np.float64) - back
        eps = (np.sqrt((err ** 2).sum(axis=1)).astype(np.float32) * PREFILTER_INFLATE + np.float32(1e-12))
        nd = np.sqrt((back ** 2).sum(axis=1)).astype(np.float32) * PREFILTER_INFLATE
    meta = np.zeros((n, 4), dtype=np.float32)
    meta[:, 0] = sd
    meta[:, 1] = np.where(finite, eps, np.float32(np.inf))
    meta[:, 2] = np.where(finite, nd, np.float32(0))
    return r.astype(np.int8), meta


def prefilter_bounds(d_stored: np.ndarray, q_stored: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(L, U) [B,N]: the lower / upper bounds the int8 prefilter puts on every exact score
    `q_stored . d_stored`:  s8 = sd sq (q8 . d8),  m = |eq| nd + |q| eps + slack,  L = s8 - m,
    U = s8 + m (Cauchy-Schwarz on q = sq q8 + eq, d = sd d8 + ed)."""
    d8, dm = quantize_rows_int8(d_stored)
    q8, qm = quantize_rows_int8(q_stored)
    q = np.ascontiguousarray(q_stored, dtype=np.float32)
    qn = np.sqrt((q.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32) * PREFILTER_INFLATE
    qe = qm[:, 1]
    acc = q8.astype(np.int32) @ d8.astype(np.int32).T                      # exact integers
    s8 = (dm[None, :, 0] * qm[:, None, 0]).astype(np.float32) * acc.astype(np.float32)
    with np.errstate(all="ignore"):
        m = ((qe[:, None] * dm[None, :, 2] + qn[:, None] * dm[None, :, 1]
              + PREFILTER_SLACK * qn[:, None] * (dm[None, :, 2] + dm[None, :, 1])) * PREFILTER_INFLATE
             + np.float32(1e-30)).astype(np.float32)
        return (s8 - m).astype(np.float32), (s8 + m).astype(np.float32)


def k3_score_emulation(d_row: np.ndarray, q_row: np.ndarray, per: int = 8) -> np.float32:
    """The fp32 operation order of the exact GEMV scan (topk_gemv.cu, also the rescoring pass of
    K3p) for ONE row: lane l of a warp holds elements g*(32*per) + l*per + e (per = 8 for 16-bit
    storage, 4 for fp32), accumulates the even e into one FMA chain and the odd e into another
    (g outer, e inner), adds the two chains, then a 5-step xor butterfly over the 32 lanes.
    Each FMA rounds once (the product is formed exactly in float64).  Used to check the rounding
    allowance of the prefilter bound: |this - exact| <= 22 * 2^-24 * |q| |d|."""
    d = np.asarray(d_row, dtype=np.float32).astype(np.float64)
    q = np.asarray(q_row, dtype=np.float32).astype(np.float64)
    groups = EMBED_DIM // (32 * per)
    lanes = np.zeros(32, dtype=np.float32)
    for l in range(32):
        a0 = np.float32(0.0)
        a1 = np.float32(0.0)
        for g in range(groups):
            base = g * 32 * per + l * per
            for e in range(0, per, 2):
                a0 = np.float32(d[base + e] * q[base + e] + np.float64(a0))
                a1 = np.float32(d[base + e + 1] * q[base + e + 1] + np.float64(a1))
        lanes[l] = np.float32(a0 + a1)
    idx = np.arange(32)
    for step in (16, 8, 4, 2, 1):
        lanes = (lanes + lanes[idx ^ step]).astype(np.float32)
    return lanes[0]
