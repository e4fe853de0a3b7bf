/*
 * Plain-C restatement of the reference's retrieval hot path.  TEST INFRASTRUCTURE ONLY
 * (oracle/__init__.py says who may use it); an independent second statement of the arithmetic
 * next to oracle/numpy_oracle.py, compiled by __graft_entry__.build() with gcc.  Nothing under
 * semantic-query-engine_b200/ links or loads it.
 *
 * Reference lines (relative to /root/reference):
 *   app/main.py:59-64    cosine_similarity: dot/(|a||b|), zero norm -> 0.0
 *   app/main.py:73-90    lfu_cache_get scan: running max from (-1.0, -1), strict '>', miss iff
 *                        best < CACHE_SIM_THRESHOLD (a Python float, i.e. a double)
 *   app/main.py:315-316, :353-354, app/embedding_gen.py:215-216
 *                        E / (np.linalg.norm(E, axis=1, keepdims=True) + 1e-9)
 *   app/main.py:356-367  k-NN leg (external index), restated as exact scoring + stable order
 *
 * Built with -O2 -ffp-contract=off: every fp32 operation below rounds once, as numpy's does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SQE_ORACLE_DIM 1024

/* numpy's fp32 add.reduce over a contiguous row of n (multiple of 8, n <= 128) elements:
 * eight strided accumulators, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
 * (numpy/core/src/umath/loops_utils.h.src, pairwise sum, the n <= PW_BLOCKSIZE branch). */
static float block_sumsq(const float *x, int n) {
    float r[8];
    for (int j = 0; j < 8; ++j) r[j] = x[j] * x[j];
    for (int i = 8; i < n; i += 8)
        for (int j = 0; j < 8; ++j) r[j] = r[j] + x[i + j] * x[i + j];
    return ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
}

/* ... and its recursion for n > 128: split at n/2 rounded down to a multiple of 8 */
static float pairwise_sumsq(const float *x, int n) {
    if (n <= 128) return block_sumsq(x, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sumsq(x, n2) + pairwise_sumsq(x + n2, n - n2);
}

/* sum of squares of one 1024-d row in numpy's order (np.linalg.norm = sqrt(add.reduce(x*x))) */
float sqe_oracle_row_sumsq(const float *row) { return pairwise_sumsq(row, SQE_ORACLE_DIM); }

/* out[i,:] = in[i,:] / (|in[i,:]| + 1e-9)          main.py:315-316.
 * numpy evaluates `norms + 1e-9` in fp32 (a Python float does not promote a float32 array). */
void sqe_oracle_normalize_rows(const float *in, float *out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        const float *x = in + i * SQE_ORACLE_DIM;
        float *y = out + i * SQE_ORACLE_DIM;
        const float den = sqrtf(sqe_oracle_row_sumsq(x)) + 1e-9f;
        for (int j = 0; j < SQE_ORACLE_DIM; ++j) y[j] = x[j] / den;
    }
}

/* main.py:59-64.  For 1-D inputs np.linalg.norm is sqrt(dot(x, x)) and np.dot is a BLAS call
 * whose summation order is not specified, so this function is compared with a tolerance (the
 * zero guard is exact): norms in numpy's row order, the dot product accumulated in double. */
double sqe_oracle_cosine(const float *a, const float *b) {
    const float na = sqrtf(sqe_oracle_row_sumsq(a));
    const float nb = sqrtf(sqe_oracle_row_sumsq(b));
    if (na == 0.0f || nb == 0.0f) return 0.0;
    double dot = 0.0;
    for (int j = 0; j < SQE_ORACLE_DIM; ++j) dot += (double)a[j] * (double)b[j];
    return dot / ((double)na * (double)nb);
}

/* scores[q, r] = dot(Q[q], D[r]) accumulated in double, rounded to fp32 once */
void sqe_oracle_scores(const float *D, int64_t n, const float *Q, int b, float *scores) {
    for (int q = 0; q < b; ++q)
        for (int64_t r = 0; r < n; ++r) {
            const float *x = D + r * SQE_ORACLE_DIM, *y = Q + (int64_t)q * SQE_ORACLE_DIM;
            double acc = 0.0;
            for (int j = 0; j < SQE_ORACLE_DIM; ++j) acc += (double)x[j] * (double)y[j];
            scores[(int64_t)q * n + r] = (float)acc;
        }
}

/* Best-first top-k of each row of `scores` [b, n]: score descending, ties -> lower index
 * (the stable order that generalises "first maximum wins", main.py:84); NaN ranks last like
 * numpy's sort.  Empty slots: (-inf, -1).  Simple insertion into a k-list: O(n k). */
void sqe_oracle_topk(const float *scores, int b, int64_t n, int k, float *out_score, int64_t *out_idx) {
    for (int q = 0; q < b; ++q) {
        float *os = out_score + (int64_t)q * k;
        int64_t *oi = out_idx + (int64_t)q * k;
        int filled = 0;
        for (int i = 0; i < k; ++i) { os[i] = -INFINITY; oi[i] = -1; }
        for (int64_t r = 0; r < n; ++r) {
            float s = scores[(int64_t)q * n + r];
            if (s != s) s = -INFINITY;                       /* NaN -> last */
            if (s == 0.0f) s = 0.0f;                         /* -0.0 == +0.0 */
            /* rows arrive in increasing order: an equal score never displaces an earlier row */
            int pos = filled;
            while (pos > 0 && s > os[pos - 1]) --pos;
            if (pos >= k) continue;
            const int last = filled < k ? filled : k - 1;
            for (int i = last; i > pos; --i) { os[i] = os[i - 1]; oi[i] = oi[i - 1]; }
            os[pos] = s;
            oi[pos] = r;
            if (filled < k) ++filled;
        }
    }
}

/* lfu_cache_get's scan over stored unit rows (main.py:73-90): returns the list index of the
 * first maximum (or -1), its similarity, and whether it is a hit (!(best < threshold), compared
 * in double like Python does). */
void sqe_oracle_cache_lookup(const float *C, int64_t n, const float *q, double threshold,
                             int32_t *out_idx, float *out_sim, uint8_t *out_hit) {
    double best = -1.0;                                      /* main.py:74 */
    int64_t best_i = -1;                                     /* main.py:75 */
    for (int64_t r = 0; r < n; ++r) {
        const float *x = C + r * SQE_ORACLE_DIM;
        double acc = 0.0;
        for (int j = 0; j < SQE_ORACLE_DIM; ++j) acc += (double)x[j] * (double)q[j];
        const double sim = (double)(float)acc;
        if (sim > best) { best = sim; best_i = r; }          /* main.py:84: strict '>' */
    }
    *out_idx = (int32_t)best_i;
    *out_sim = (float)best;
    *out_hit = (best_i >= 0 && !(best < threshold)) ? 1 : 0; /* main.py:89 */
}

/* ------------------------------------------------------------------------------------------
 * The int8 prefilter (K3p).  No reference line: the reference scores every row (main.py:59-64);
 * this restates the quantiser and the Cauchy-Schwarz bound that lets the device skip rows
 * without changing the answer:  L <= exact score <= U  for every row.
 *   d8 = clip(rint(x * (127 / max|x|))), sd = max|x| / 127, eps >= |x - sd d8|_2, nd >= |sd d8|_2
 *   s8 = sd sq (q8 . d8)  (exact integer dot product),  m = |eq| nd + |q| eps + slack
 * Norms are accumulated in double and inflated by 0.1 % like the device does in fp32.
 * ------------------------------------------------------------------------------------------ */
#define SQE_PF_INFLATE 1.001f
#define SQE_PF_SLACK 4e-6f

void sqe_oracle_quantize_row(const float *x, int8_t *d8, float *meta /* [4] */) {
    float mx = 0.0f;
    double ss = 0.0;
    for (int j = 0; j < SQE_ORACLE_DIM; ++j) {
        const float a = fabsf(x[j]);
        if (a > mx) mx = a;                                  /* NaN never wins, like fmaxf */
        ss += (double)x[j] * (double)x[j];
    }
    const int finite = isfinite((float)ss);
    const int live = finite && mx > 0.0f;
    const float sd = live ? mx / 127.0f : 0.0f;
    const float inv = live ? 127.0f / mx : 0.0f;
    double e2 = 0.0, n2 = 0.0;
    for (int j = 0; j < SQE_ORACLE_DIM; ++j) {
        const float v = live ? x[j] : 0.0f;
        float r = rintf(v * inv);
        if (r > 127.0f) r = 127.0f;
        if (r < -127.0f) r = -127.0f;
        const double back = (double)(sd * r), err = (double)v - back;
        e2 += err * err;
        n2 += back * back;
        d8[j] = (int8_t)r;
    }
    meta[0] = sd;
    meta[1] = finite ? (float)sqrt(e2) * SQE_PF_INFLATE + 1e-12f : INFINITY;
    meta[2] = finite ? (float)sqrt(n2) * SQE_PF_INFLATE : 0.0f;
    meta[3] = 0.0f;
}

/* L[q, r], U[q, r] for stored rows D [n, 1024] and stored queries Q [b, 1024] */
void sqe_oracle_prefilter_bounds(const float *D, int64_t n, const float *Q, int b, float *L, float *U) {
    int8_t *d8 = (int8_t *)malloc((size_t)n * SQE_ORACLE_DIM);
    float *dm = (float *)malloc((size_t)n * 4 * sizeof(float));
    for (int64_t r = 0; r < n; ++r) sqe_oracle_quantize_row(D + r * SQE_ORACLE_DIM, d8 + r * SQE_ORACLE_DIM, dm + r * 4);
    for (int q = 0; q < b; ++q) {
        const float *y = Q + (int64_t)q * SQE_ORACLE_DIM;
        int8_t q8[SQE_ORACLE_DIM];
        float qm[4];
        sqe_oracle_quantize_row(y, q8, qm);
        double ss = 0.0;
        for (int j = 0; j < SQE_ORACLE_DIM; ++j) ss += (double)y[j] * (double)y[j];
        const float qn = (float)sqrt(ss) * SQE_PF_INFLATE, qe = qm[1], sq = qm[0];
        for (int64_t r = 0; r < n; ++r) {
            int32_t acc = 0;
            const int8_t *x = d8 + r * SQE_ORACLE_DIM;
            for (int j = 0; j < SQE_ORACLE_DIM; ++j) acc += (int32_t)x[j] * (int32_t)q8[j];
            const float sd = dm[r * 4], eps = dm[r * 4 + 1], nd = dm[r * 4 + 2];
            const float s8 = (sd * sq) * (float)acc;
            const float m = ((qe * nd + qn * eps) + SQE_PF_SLACK * qn * (nd + eps)) * SQE_PF_INFLATE + 1e-30f;
            L[(int64_t)q * n + r] = s8 - m;
            U[(int64_t)q * n + r] = s8 + m;
        }
    }
    free(d8);
    free(dm);
}
