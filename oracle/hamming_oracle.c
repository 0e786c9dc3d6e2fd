/* CPU oracle (TEST INFRASTRUCTURE ONLY) - weighted Hamming distance matrix.
 *
 * Restates SciPy 1.18.1's `cdist(..., metric="hamming", w=w)`
 * ($SP/scipy/spatial/distance.py:1718-1723 -> _distance_pybind.cdist_hamming), which is
 * what sknnr's RFNN path runs through sklearn's brute fallback
 * ($SP/sklearn/neighbors/_base.py:879-908) with w = hamming_weights_
 * (ref:src/sknnr/_weighted_trees.py:65-98,139-140):
 *
 *     d(q, r) = (sum_{t : q_t != r_t} w_t) / (sum_t w_t)
 *
 * with both sums accumulated strictly left to right in float64 (verified bit for bit
 * against SciPy in tests/test_oracle.py).  Compiled with -ffp-contract=off so no FMA
 * contraction changes the rounding.
 */
#include <stdint.h>

void hamming_cdist_w(const int64_t *Q, const int64_t *R, const double *w, int64_t n_q,
                     int64_t n_r, int64_t T, double *out)
{
    double den = 0.0;
    for (int64_t t = 0; t < T; ++t) den += w[t];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_q; ++i) {
        const int64_t *q = Q + i * T;
        for (int64_t j = 0; j < n_r; ++j) {
            const int64_t *r = R + j * T;
            double num = 0.0;
            for (int64_t t = 0; t < T; ++t) {
                if (q[t] != r[t]) num += w[t];
            }
            out[i * n_r + j] = num / den;
        }
    }
}
