"""CPU checks of the error bound that makes the int8-prefiltered scan (K3p) exact: for every
row, L <= exact score <= U, whatever the data -- so a row with U below the k-th best L can be
skipped without changing the result.  The device kernels are checked against the exact scan bit
for bit in tests/test_gpu_parity.py; this file checks the mathematics (oracle restatement)."""
import numpy as np
import pytest

import oracle
from oracle import numpy_oracle as no

DIM = 1024


def _stored(x, dtype):
    return oracle.from_storage(oracle.to_storage(oracle.normalize_rows(x), dtype), dtype)


def _check(d, q, k=10):
    L, U = oracle.prefilter_bounds(d, q)
    s64 = q.astype(np.float64) @ d.astype(np.float64).T
    s32 = q @ d.T                                        # an fp32 summation order (BLAS)
    for s in (s64, s32.astype(np.float64)):
        ok = np.isfinite(s)
        with np.errstate(invalid="ignore"):
            assert not (s[ok] < L.astype(np.float64)[ok]).any()          # NaN bounds never exclude
            assert not (U.astype(np.float64)[ok] < s[ok]).any()
    # the selection argument: every row of the exact top-k survives the filter U >= tau
    kk = min(k, d.shape[0])
    with np.errstate(invalid="ignore"):
        tau = -np.sort(-np.where(np.isnan(L), -np.inf, L), axis=1)[:, kk - 1]
        keep = ~(U < tau[:, None])
    _, idx = oracle.topk_from_scores(s32, kk)
    for b in range(q.shape[0]):
        assert keep[b, idx[b][idx[b] >= 0]].all()
    return keep.sum(axis=1)


@pytest.mark.parametrize("dtype", ["fp32", "bf16", "fp16"])
def test_bounds_hold_on_random_unit_rows(dtype):
    rng = np.random.default_rng(5)
    d = _stored(rng.standard_normal((3000, DIM)).astype(np.float32), dtype)
    q = _stored(rng.standard_normal((5, DIM)).astype(np.float32), dtype)
    kept = _check(d, q)
    assert kept.max() < 600                              # the filter does filter (3000 rows, k = 10)


def test_bounds_hold_in_the_cauchy_schwarz_worst_case():
    """Queries aligned with a row's own quantisation error (the direction in which the bound is
    tight), rows with outliers, near-duplicates of the query, scaled (non-unit) rows."""
    rng = np.random.default_rng(6)
    x = rng.standard_normal((400, DIM)).astype(np.float32)
    x[5, 17] = 40.0                                      # outlier -> coarse scale
    x[6] = 0.0
    d = _stored(x, "bf16")
    d[7] *= np.float32(3.5)                              # not unit norm
    d8, meta = oracle.quantize_rows_int8(d)
    err = d - meta[:, :1] * d8.astype(np.float32)
    q = np.stack([err[3], -err[4], d[9] + 0.01 * err[9], d[5], d[7], np.zeros(DIM, np.float32)]).astype(np.float32)
    q[:5] = oracle.normalize_rows(q[:5])
    _check(d, q, k=3)
    near = oracle.normalize_rows(d[9][None, :] + 1e-3 * rng.standard_normal((200, DIM)).astype(np.float32))
    _check(np.concatenate([d, near]), q, k=10)


def test_non_finite_rows_and_queries_are_never_filtered_out():
    rng = np.random.default_rng(7)
    d = _stored(rng.standard_normal((64, DIM)).astype(np.float32), "fp32")
    d[3, 5] = np.nan
    d[4, 6] = np.inf
    q = _stored(rng.standard_normal((3, DIM)).astype(np.float32), "fp32")
    q[2, 0] = np.nan
    L, U = oracle.prefilter_bounds(d, q)
    with np.errstate(invalid="ignore"):
        assert not (U[:, 3] < np.inf).any() and not (U[:, 4] < np.inf).any()      # +inf or NaN
        assert not (U[2] < np.inf).any()                                          # NaN query: all rows
    d8, meta = oracle.quantize_rows_int8(d)
    assert not d8[3].any() and not d8[4].any() and np.isinf(meta[3, 1]) and np.isinf(meta[4, 1])


def test_quantiser_is_max_abs_round_to_nearest():
    rng = np.random.default_rng(8)
    x = rng.standard_normal((16, DIM)).astype(np.float32)
    d8, meta = oracle.quantize_rows_int8(x)
    assert np.abs(d8).max(axis=1).tolist() == [127] * 16
    back = meta[:, :1] * d8.astype(np.float32)
    assert np.abs(x - back).max() <= meta[:, 0].max() * 0.5001
    assert np.all(np.linalg.norm((x - back).astype(np.float64), axis=1) <= meta[:, 1])
    assert np.all(np.linalg.norm(back.astype(np.float64), axis=1) <= meta[:, 2])


def test_bounds_hold_for_arbitrary_rows_hypothesis():
    """Property test: any mixture of scales, sparsity, outliers and sign patterns, unit or not."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(seed=st.integers(0, 2 ** 31 - 1), scale_exp=st.integers(-30, 12), sparsity=st.floats(0.0, 0.99),
           outliers=st.integers(0, 5), dtype=st.sampled_from(["fp32", "bf16", "fp16"]))
    def prop(seed, scale_exp, sparsity, outliers, dtype):
        rng = np.random.default_rng(seed)
        x = rng.standard_normal((48, DIM)).astype(np.float32)
        x[rng.random(x.shape) < sparsity] = 0.0
        for _ in range(outliers):
            x[rng.integers(48), rng.integers(DIM)] = np.float32(rng.choice([-1, 1]) * 10.0 ** rng.integers(1, 4))
        d = oracle.from_storage(oracle.to_storage(oracle.normalize_rows(x), dtype), dtype)
        d[:8] *= np.float32(2.0 ** scale_exp)            # some rows far from unit norm
        q = rng.standard_normal((3, DIM)).astype(np.float32)
        q[1] = d[9] + 1e-3 * q[1]
        q = oracle.from_storage(oracle.to_storage(q if dtype == "fp32" else oracle.normalize_rows(q), dtype), dtype)
        _check(d, q, k=5)
    prop()


def test_plain_c_twin_agrees_with_the_numpy_restatement():
    """oracle/c_oracle.c states the quantiser and the bound a second time, independently: same
    int8 rows and scales bit for bit, same bounds up to the rounding of the norms, and its bounds
    hold as well."""
    from oracle import c_oracle as co
    if not co.available():
        pytest.skip("no gcc and no prebuilt C oracle")
    rng = np.random.default_rng(11)
    x = rng.standard_normal((300, DIM)).astype(np.float32)
    x[3] = 0.0
    x[4, 9] = 30.0
    d = _stored(x, "bf16")
    d[5, 1] = np.nan
    q = _stored(rng.standard_normal((4, DIM)).astype(np.float32), "bf16")
    w8, wm = oracle.quantize_rows_int8(d)
    c8, cm = co.quantize_rows_int8(d)
    np.testing.assert_array_equal(c8, w8)
    np.testing.assert_array_equal(cm[:, 0], wm[:, 0])
    np.testing.assert_allclose(cm[:, 1:3], wm[:, 1:3], rtol=1e-6)
    Ln, Un = oracle.prefilter_bounds(d, q)
    Lc, Uc = co.prefilter_bounds(d, q)
    ok = np.isfinite(Un)
    np.testing.assert_allclose(Lc[ok], Ln[ok], rtol=0, atol=1e-6)
    np.testing.assert_allclose(Uc[ok], Un[ok], rtol=0, atol=1e-6)
    assert not np.isfinite(Uc[:, 5]).any()                 # the NaN row can never be ruled out
    s = q.astype(np.float64) @ np.where(np.isnan(d), 0, d).astype(np.float64).T
    fin = np.isfinite(Uc)
    assert not (s[fin] < Lc[fin]).any() and not (Uc[fin] < s[fin]).any()


def test_rounding_allowance_covers_the_exact_scan_s_own_summation_order():
    """The bound is on the real-number dot product; what the device compares against is the fp32
    score of the exact scan.  Its summation tree (two FMA chains of 16 per lane, one add, a 5-step
    butterfly) has at most 22 roundings on any path, i.e. an error <= 22 * 2^-24 * sum|q_i d_i|
    <= 1.3e-6 |q||d|; the bound reserves 4e-6 |q|(|d8| + eps) for it.  Checked here on an emulation
    of that order, including all-positive vectors where the rounding errors do not cancel."""
    rng = np.random.default_rng(21)
    worst = 0.0
    for trial in range(24):
        per = 8 if trial % 2 == 0 else 4
        x = rng.standard_normal((2, DIM)).astype(np.float32)
        if trial % 3 == 0:
            x = np.abs(x)                                   # no cancellation
        if trial % 4 == 1:
            x[1] = x[0] + 1e-3 * x[1]                       # score near 1
        d, q = _stored(x[:1], "bf16" if per == 8 else "fp32")[0], _stored(x[1:], "bf16" if per == 8 else "fp32")[0]
        s = float(no.k3_score_emulation(d, q, per))
        exact = float(d.astype(np.float64) @ q.astype(np.float64))
        scale = float(np.linalg.norm(d.astype(np.float64)) * np.linalg.norm(q.astype(np.float64)))
        worst = max(worst, abs(s - exact) / scale)
        L, U = oracle.prefilter_bounds(d[None, :], q[None, :])
        assert L[0, 0] <= s <= U[0, 0]
    assert worst <= 22 * 2.0 ** -24, worst                  # the analytical allowance ...
    assert worst < 4e-6 / 3                                 # ... and the reserved slack is 3x that
