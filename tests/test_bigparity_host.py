"""The at-size comparator (tests/bigparity.py) is itself pinned on the CPU: fed with the oracle's
own answers it must accept them with nothing excused, and it must reject a wrong row, a wrong
score, a wrong order and a wrong tie order."""
import numpy as np
import pytest
import torch

import oracle
from bigparity import assert_topk_matches_at_size, oracle_candidates, stored_block_f32

DIM = 1024


def _shard(dtype, n, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, DIM)).astype(np.float32)
    x[n - 2] = x[5]                                     # exact duplicates (tie -> lower row)
    st = oracle.to_storage(oracle.normalize_rows(x), dtype)
    q = oracle.to_storage(oracle.normalize_rows(np.concatenate([x[5:6] * 2, rng.standard_normal((6, DIM)).astype(np.float32)])), dtype)
    if dtype == "fp32":
        t, tq = torch.from_numpy(st), torch.from_numpy(q)
    elif dtype == "fp16":
        t, tq = torch.from_numpy(st), torch.from_numpy(q)
    else:
        t = torch.from_numpy(st.view(np.int16)).view(torch.bfloat16)
        tq = torch.from_numpy(q.view(np.int16)).view(torch.bfloat16)
    return t, tq, oracle.from_storage(st, dtype), oracle.from_storage(q, dtype)


@pytest.mark.parametrize("dtype", ["fp32", "bf16", "fp16", "bf16x2"])
def test_comparator_accepts_the_oracle_and_rejects_damage(dtype):
    n, k = 3001, 10
    t, tq, d_st, q_st = _shard(dtype, n, 3)
    np.testing.assert_array_equal(stored_block_f32(t, dtype, 0, n), d_st)
    np.testing.assert_array_equal(stored_block_f32(tq, dtype, 0, 7), q_st)
    s, i = oracle.topk_cosine(d_st, q_st, k)
    cands = oracle_candidates(t, dtype, n, q_st, k, block=700)
    for r in range(7):
        assert set(i[r].tolist()) <= set(cands[r].tolist())
    exc, worst = assert_topk_matches_at_size(s, i, t, dtype, n, q_st, k, 2e-6, 1e-6, cands=cands)
    assert exc == 0 and worst < 1e-6
    assert i[0, :2].tolist() == [5, n - 2]
    exc, _ = assert_topk_matches_at_size(s, i + 10**10, t, dtype, n, q_st, k, 2e-6, 1e-6, idx_offset=10**10)
    assert exc == 0
    bad = i.copy(); bad[3, k - 1] = (set(range(n)) - set(i[3].tolist())).pop()
    with pytest.raises(AssertionError):
        assert_topk_matches_at_size(s, bad, t, dtype, n, q_st, k, 2e-6, 1e-6)
    bad_s = s.copy(); bad_s[2, 4] += 1e-4
    with pytest.raises(AssertionError):
        assert_topk_matches_at_size(bad_s, i, t, dtype, n, q_st, k, 2e-6, 1e-6)
    sw_i, sw_s = i.copy(), s.copy()
    sw_i[0, [0, 1]] = sw_i[0, [1, 0]]                   # the tie the wrong way round
    with pytest.raises(AssertionError):
        assert_topk_matches_at_size(sw_s, sw_i, t, dtype, n, q_st, k, 2e-6, 1e-6)


def test_comparator_handles_short_shards():
    t, tq, d_st, q_st = _shard("fp32", 6, 4)
    s, i = oracle.topk_cosine(d_st, q_st, 10)
    exc, _ = assert_topk_matches_at_size(s, i, t, "fp32", 6, q_st, 10, 2e-6, 1e-6)
    assert exc == 0
