/* ORACLE — TEST INFRASTRUCTURE ONLY.  Compiled C restatement of the row loops
 * of the reference's smoother, used by oracle/openmg_oracle.py so parity cases
 * of 64^3..128^3 finish in seconds and so the CPU baseline is compiled code.
 * Single thread, plain C, no SIMD intrinsics.
 *
 * oracle_gs_sweeps   : openmg/solvers.py:56-68 (lexicographic forward GS over CSR rows,
 *                      x[i] += (b[i] - A[i,:]·x) / A[i,i], in place)
 * oracle_jacobi_sweeps / oracle_rbgs_sweeps : same row update (:68) applied to all rows
 *                      at once (weight omega) / colour by colour with lagged same-colour
 *                      values — definitions in oracle/openmg_oracle.py (jacobi, rbgs).
 */
#include <stdint.h>
#include <string.h>

static int diag_of(const int32_t *indptr, const int32_t *indices, const double *data,
                   int64_t i, double *out) {
    /* A[i,i]: scipy sums duplicates on scalar __getitem__; do the same */
    double d = 0.0; int found = 0;
    for (int32_t p = indptr[i]; p < indptr[i + 1]; ++p)
        if (indices[p] == i) { d += data[p]; found = 1; }
    *out = d;
    return found && d != 0.0;
}

int oracle_gs_sweeps(int64_t n, const int32_t *indptr, const int32_t *indices,
                     const double *data, const double *b, double *x, int sweeps) {
    for (int s = 0; s < sweeps; ++s) {
        for (int64_t i = 0; i < n; ++i) {
            double aix = 0.0, d;
            for (int32_t p = indptr[i]; p < indptr[i + 1]; ++p)
                aix += data[p] * x[indices[p]];
            if (!diag_of(indptr, indices, data, i, &d)) return 1;
            x[i] = x[i] + (b[i] - aix) / d;
        }
    }
    return 0;
}

int oracle_jacobi_sweeps(int64_t n, const int32_t *indptr, const int32_t *indices,
                         const double *data, const double *b, double *x, double *tmp,
                         double omega, int sweeps) {
    for (int s = 0; s < sweeps; ++s) {
        for (int64_t i = 0; i < n; ++i) {
            double aix = 0.0, d;
            for (int32_t p = indptr[i]; p < indptr[i + 1]; ++p)
                aix += data[p] * x[indices[p]];
            if (!diag_of(indptr, indices, data, i, &d)) return 1;
            tmp[i] = x[i] + omega * (b[i] - aix) / d;
        }
        memcpy(x, tmp, (size_t)n * sizeof(double));
    }
    return 0;
}

int oracle_rbgs_sweeps(int64_t n, const int32_t *indptr, const int32_t *indices,
                       const double *data, const double *b, double *x, double *tmp,
                       const uint8_t *colour, int sweeps) {
    for (int s = 0; s < sweeps; ++s) {
        for (int c = 0; c < 2; ++c) {
            for (int64_t i = 0; i < n; ++i) {
                if (colour[i] != c) { tmp[i] = x[i]; continue; }
                double aix = 0.0, d;
                for (int32_t p = indptr[i]; p < indptr[i + 1]; ++p)
                    aix += data[p] * x[indices[p]];
                if (!diag_of(indptr, indices, data, i, &d)) return 1;
                tmp[i] = x[i] + (b[i] - aix) / d;
            }
            memcpy(x, tmp, (size_t)n * sizeof(double));
        }
    }
    return 0;
}
