/* CPU oracle for the MMR re-rank.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates rerank_with_mmr (main.py:133-169): the best-scored candidate is taken first; then, until
 * min(top_k, C) items are chosen, the remaining candidate (visited in ranked order, unmapped items skipped)
 * with the largest
 *     mmr = lambda * score - (1 - lambda) * max_{s in selected, mapped} cos(v_c, v_s)        (main.py:161)
 * is appended; `>` at main.py:162 keeps the FIRST maximum.  With no mapped item selected yet the
 * similarity term is 0 (main.py:154-155).  cos() is scikit-learn's cosine_similarity (main.py:159,
 * unpinned, 1.9.0 here): normalize both rows, dot product -- summation order left to BLAS, so THE
 * CONTRACT fixes it like oracle/knn_oracle.c: sequential-fma norms and dots in fp32.  The mmr
 * expression is evaluated in fp32 (NumPy >= 2 promotion of `python_float * np.float32`).
 * Parity status: pinned against the reference's own rerank_with_mmr on tie-free random data
 * (tests/golden/mmr_*.npz from tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* emb: [n_items, d]; scores / emb_idx: [C] in ranked order (emb_idx < 0: unmapped); order_out: [top_k]
 * positions into the candidate list; returns the number selected. */
int mmr_rerank(const float *emb, int64_t n_items, int d, const float *scores, const int64_t *emb_idx, int C,
               float lambda, int top_k, int32_t *order_out) {
    (void)n_items;
    if (C <= 0) return 0;
    float *vhat = (float *)malloc((size_t)C * d * sizeof(float));
    float *maxsim = (float *)malloc((size_t)C * sizeof(float));
    char *taken = (char *)calloc((size_t)C, 1);
    for (int c = 0; c < C; ++c) {
        maxsim[c] = -INFINITY;
        if (emb_idx[c] < 0) continue;
        const float *v = emb + emb_idx[c] * (int64_t)d;
        float ss = 0.0f;
        for (int j = 0; j < d; ++j) ss = fmaf(v[j], v[j], ss);
        float nrm = sqrtf(ss);
        if (nrm == 0.0f) nrm = 1.0f;
        for (int j = 0; j < d; ++j) vhat[(size_t)c * d + j] = v[j] / nrm;
    }
    const float one_minus = (float)(1.0 - (double)lambda);      /* np.float32(1 - lambda_param) */
    int n = 0, have_sel = 0, last = 0;
    const int want = top_k < C ? top_k : C;
    order_out[n++] = 0;
    taken[0] = 1;
    while (n < want) {
        if (emb_idx[last] >= 0) {
            have_sel = 1;
            for (int c = 0; c < C; ++c) {
                if (taken[c] || emb_idx[c] < 0) continue;
                float s = 0.0f;
                for (int j = 0; j < d; ++j) s = fmaf(vhat[(size_t)c * d + j], vhat[(size_t)last * d + j], s);
                if (s > maxsim[c]) maxsim[c] = s;
            }
        }
        int best = -1;
        float best_v = -INFINITY;
        for (int c = 0; c < C; ++c) {
            if (taken[c] || emb_idx[c] < 0) continue;
            const float ms = have_sel ? maxsim[c] : 0.0f;
            const float v = lambda * scores[c] - one_minus * ms;        /* two roundings, no fma */
            if (v > best_v) { best_v = v; best = c; }
        }
        if (best < 0) break;
        order_out[n++] = best;
        taken[best] = 1;
        last = best;
    }
    free(vhat); free(maxsim); free(taken);
    return n;
}
