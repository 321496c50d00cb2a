// omg_setup.cu — hierarchy setup on the device:
//   * depth rule + closed-form restriction descriptors   (openmg/operators.py:15-141)
//   * Galerkin coarse operators A_{l+1} = R_l A_l R_l^T   (openmg/operators.py:144-188)
//   * constant-band ("stencil") detection with exception rows
//   * dense inverse of the coarsest operator (coarse "factor") for the direct coarse solve
//     (openmg/solvers.py:16-26 uses SuperLU every cycle; we invert once)
#include <algorithm>
#include <map>
#include <stdarg.h>
#include <string.h>

#include "omg_hier.cuh"
#include "omg_kernels.cuh"

// ------------------------------------------------------------------ allocation

int h_alloc(omg_hierarchy *h, void **p, size_t bytes, bool zero) {
    *p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return omg_set_error(OMG_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    h->allocs.push_back(*p);
    if (zero) CUDA_TRY(cudaMemsetAsync(*p, 0, bytes, g.stream));
    return OMG_OK;
}

void h_free(omg_hierarchy *h, void *p) {
    if (!p) return;
    for (size_t i = 0; i < h->allocs.size(); ++i)
        if (h->allocs[i] == p) {
            h->allocs[i] = h->allocs.back();
            h->allocs.pop_back();
            break;
        }
    cudaFree(p);
}

// ------------------------------------------------------------------ exclusive scan (int32)

#define SCAN_TPB 512
#define SCAN_IPT 8
#define SCAN_TILE (SCAN_TPB * SCAN_IPT)

__global__ void __launch_bounds__(SCAN_TPB) k_scan_tile(const int *__restrict__ in, int *__restrict__ out, int n,
                                                        int *__restrict__ tile_sums) {
    __shared__ int sh[SCAN_TPB];
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
    int v[SCAN_IPT];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        int idx = base + k;
        v[k] = idx < n ? in[idx] : 0;
        s += v[k];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < SCAN_TPB; o <<= 1) {   // Hillis-Steele inclusive scan of thread sums
        int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    int excl = sh[threadIdx.x] - s;
    if (threadIdx.x == SCAN_TPB - 1 && tile_sums) tile_sums[blockIdx.x] = sh[threadIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        int idx = base + k;
        if (idx < n) out[idx] = excl;
        excl += v[k];
    }
}

__global__ void __launch_bounds__(SCAN_TPB) k_scan_add(int *__restrict__ out, int n,
                                                       const int *__restrict__ tile_offs) {
    int add = tile_offs[blockIdx.x];
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k)
        if (base + k < n) out[base + k] += add;
}

// out[i] = sum(in[0..i)), i < n; *total_host = sum(in[0..n)) (synchronises). in may alias out.
int exclusive_scan_i32(const int *in, int *out, int n, int *total_host, cudaStream_t st) {
    if (n <= 0) {
        if (total_host) *total_host = 0;
        return OMG_OK;
    }
    int tiles = cdiv(n, SCAN_TILE);
    int *sums = nullptr;
    CUDA_TRY(cudaMalloc(&sums, sizeof(int) * (size_t)(tiles + 1)));
    k_scan_tile<<<tiles, SCAN_TPB, 0, st>>>(in, out, n, sums);
    int total = 0;
    int rc = OMG_OK;
    if (tiles > 1) {
        rc = exclusive_scan_i32(sums, sums, tiles, &total, st);
        if (rc == OMG_OK) k_scan_add<<<tiles, SCAN_TPB, 0, st>>>(out, n, sums);
        if (rc == OMG_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = omg_set_error(OMG_ECUDA, "scan failed");
    } else {
        cudaError_t e = cudaMemcpyAsync(&total, sums, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = omg_set_error(OMG_ECUDA, "scan failed: %s", cudaGetErrorString(e));
    }
    cudaFree(sums);
    if (total_host) *total_host = total;
    cudaError_t e = cudaGetLastError();
    if (rc == OMG_OK && e != cudaSuccess) rc = omg_set_error(OMG_ECUDA, "scan: %s", cudaGetErrorString(e));
    return rc;
}

// ------------------------------------------------------------------ level geometry (depth rule)

struct RDesc {
    int alpha;
    int shape[3];
    int64_t N, n;
    int k;
    int offs[8];
    bool regular;
};

// operators.restriction(shape) validity + descriptor, openmg/operators.py:45-84
static int make_rdesc(int alpha, const int64_t *shape, RDesc *d) {
    int64_t N = 1;
    for (int i = 0; i < alpha; ++i) N *= shape[i];
    int64_t n = N / (1ll << alpha);                                   // :52
    if (n == 0 || n == 1)                                              // :53-56
        return omg_set_error(OMG_ESHAPE,
                             "New restriction matrix would have shape (%lld, %lld). Coarse set would have %lld "
                             "point(s)! Try a larger problem or fewer gridLevels.",
                             (long long)n, (long long)N, (long long)n);
    if (alpha > 3 || alpha < 1)                                        // :69-71
        return omg_set_error(OMG_EDIM, "restriction(): Greater than 3 dimensions is not implemented.");
    if (N >= (1ll << 31)) return omg_set_error(OMG_EUNSUPPORTED, "level with %lld rows exceeds int32 indexing", (long long)N);
    d->alpha = alpha;
    d->N = N;
    d->n = n;
    for (int i = 0; i < 3; ++i) d->shape[i] = i < alpha ? (int)shape[i] : 1;
    int64_t NX = shape[0], NY = alpha >= 2 ? shape[1] : 1;
    int64_t o[8];
    int k = 0;
    o[k++] = 0;
    o[k++] = 1;
    if (alpha >= 2) {
        o[k++] = NX;
        o[k++] = NX + 1;
        if (alpha == 3) {
            o[k++] = NX * NY;
            o[k++] = NX * NY + 1;
            o[k++] = NX * NY + NX;
            o[k++] = NX * NY + NX + 1;
        }
    }
    std::sort(o, o + k);
    k = (int)(std::unique(o, o + k) - o);      // duplicates collapse by overwrite in the lil_matrix (:75-84)
    d->k = k;
    for (int i = 0; i < k; ++i) d->offs[i] = (int)o[i];
    // largest column written: cc[n-1] + max offset must be < N (else lil_matrix raises IndexError)
    int64_t g[3], st[3];
    for (int i = 0; i < alpha; ++i) g[i] = (shape[i] + 1) / 2;
    st[alpha - 1] = 1;
    for (int i = alpha - 2; i >= 0; --i) st[i] = st[i + 1] * shape[i + 1];
    int64_t r = n - 1, cc = 0;
    for (int i = alpha - 1; i >= 0; --i) {
        int64_t c = (i == 0) ? r : r % g[i];
        r = (i == 0) ? 0 : r / g[i];
        cc += 2 * c * st[i];
    }
    if (cc + o[k - 1] >= N)
        return omg_set_error(OMG_EINDEX, "restriction(): column index %lld out of range for %lld columns",
                             (long long)(cc + o[k - 1]), (long long)N);
    bool even = true;
    for (int i = 0; i < alpha; ++i) even = even && (shape[i] % 2 == 0);
    bool strides = (alpha == 1) || (alpha == 2 && NX == shape[1]) ||
                   (alpha == 3 && NX == shape[2] && NX * NY == shape[1] * shape[2]);
    d->regular = even && strides && k == (1 << alpha);
    return OMG_OK;
}

static void fill_reg(const RDesc &d, RegR *R) {
    R->alpha = d.alpha;
    R->k = d.k;
    if (d.alpha == 1) {
        R->fs1 = 1;
        R->fs2 = d.shape[0];
    } else if (d.alpha == 2) {
        R->fs1 = d.shape[0];
        R->fs2 = d.shape[1];
    } else {
        R->fs1 = d.shape[1];
        R->fs2 = d.shape[2];
    }
    R->cs1 = R->fs1 > 1 ? R->fs1 / 2 : 1;
    R->cs2 = R->fs2 / 2;
    for (int i = 0; i < 8; ++i) R->o[i] = i < d.k ? d.offs[i] : 0;
    R->nc = (int)d.n;
    R->nf = (int)d.N;
    R->w = 1.0 / (double)(1 << d.alpha);
}

// restrictionList depth rule, openmg/operators.py:128-140
int setup_levels(omg_hierarchy *h, int ndim, const int64_t *shape, int coarsestLevel, int minSize) {
    if (ndim < 1) return omg_set_error(OMG_EINVAL, "problemShape must have at least one dimension");
    std::vector<RDesc> R;
    RDesc d;
    int64_t sh[8];
    if (ndim > 8) return omg_set_error(OMG_EDIM, "restriction(): Greater than 3 dimensions is not implemented.");
    auto level_shape = [&](int level) {
        for (int i = 0; i < ndim; ++i) sh[i] = shape[i] / (1ll << level);
    };
    level_shape(0);
    OMG_TRY(make_rdesc(ndim, sh, &d));
    R.push_back(d);
    int level = 0;
    while (level < coarsestLevel) {
        ++level;
        level_shape(level);
        OMG_TRY(make_rdesc(ndim, sh, &d));
        if (d.n <= minSize) break;
        R.push_back(d);
    }
    int nlev = (int)R.size() + 1;
    // coeffecientList needs R[l-1].rows == R[l].cols == rows(A_l) (scipy raises ValueError otherwise)
    if (R[0].N != h->lv[0].n)
        return omg_set_error(OMG_EINVAL, "dimension mismatch: A is %d x %d but problemShape has %lld points",
                             h->lv[0].n, h->lv[0].n, (long long)R[0].N);
    for (size_t l = 1; l < R.size(); ++l)
        if (R[l].N != R[l - 1].n)
            return omg_set_error(OMG_EINVAL, "dimension mismatch between restriction levels %zu and %zu (%lld vs %lld)",
                                 l - 1, l, (long long)R[l - 1].n, (long long)R[l].N);
    h->nlev = nlev;
    h->lv.resize(nlev);
    for (int l = 0; l < nlev; ++l) {
        Level &L = h->lv[l];
        L.ndim = ndim;
        for (int i = 0; i < 3; ++i) L.shape[i] = i < ndim ? (int)(shape[i] / (1ll << l)) : 1;
        if (l > 0) L.n = (int)R[l - 1].n;
        L.row0 = 0;
        L.nloc = L.n;
        L.colour.flat = (ndim == 1) || (ndim == 2 && l == 0);
        L.colour.alpha = ndim;
        L.colour.s1 = ndim == 3 ? std::max(L.shape[1], 1) : 1;
        L.colour.s2 = std::max(L.shape[ndim - 1], 1);
        if (l < nlev - 1) {
            const RDesc &r = R[l];
            L.hasR = true;
            L.regular = r.regular;
            L.Rk = r.k;
            for (int i = 0; i < 8; ++i) L.Roffs[i] = i < r.k ? r.offs[i] : 0;
            L.Rw = 1.0 / (double)(1 << r.alpha);
            L.nc = (int)r.n;
            fill_reg(r, &L.reg);
        }
    }
    return OMG_OK;
}

// ------------------------------------------------------------------ explicit restriction pattern (non-regular shapes)

// Rcc[r] = first fine column of coarse row r (C-order unravel over the [::2] grid, openmg/operators.py:63-68)
__global__ void k_build_rcc(int n, int alpha, int g1, int g2, int st0, int st1, int *__restrict__ Rcc) {
    int r = blockIdx.x * OMG_TPB + threadIdx.x;
    if (r >= n) return;
    int c2 = r % g2, t = r / g2;
    int c1 = t % g1, c0 = t / g1;
    Rcc[r] = 2 * c0 * st0 + 2 * c1 * st1 + 2 * c2;
}

__global__ void k_rt_count(const int *__restrict__ Rcc, int n, int k, RegR offs, int *__restrict__ cnt) {
    int r = blockIdx.x * OMG_TPB + threadIdx.x;
    if (r >= n) return;
    for (int q = 0; q < k; ++q) atomicAdd(cnt + Rcc[r] + offs.o[q], 1);
}

__global__ void k_rt_fill(const int *__restrict__ Rcc, int n, int k, RegR offs, const int *__restrict__ RTptr,
                          int *__restrict__ cursor, int *__restrict__ RTcol) {
    int r = blockIdx.x * OMG_TPB + threadIdx.x;
    if (r >= n) return;
    for (int q = 0; q < k; ++q) {
        int j = Rcc[r] + offs.o[q];
        int p = atomicAdd(cursor + j, 1);
        RTcol[RTptr[j] + p] = r;
    }
}

__global__ void k_rt_sort(int N, const int *__restrict__ RTptr, int *__restrict__ RTcol) {
    int j = blockIdx.x * OMG_TPB + threadIdx.x;
    if (j >= N) return;
    int p0 = RTptr[j], p1 = RTptr[j + 1];
    for (int a = p0 + 1; a < p1; ++a) {
        int key = RTcol[a], b = a - 1;
        while (b >= p0 && RTcol[b] > key) {
            RTcol[b + 1] = RTcol[b];
            --b;
        }
        RTcol[b + 1] = key;
    }
}

static void rcc_params(int alpha, const int *shape, int *g1, int *g2, int *st0, int *st1);

static int build_explicit_R(omg_hierarchy *h, Level &L) {
    int n = L.nc, N = L.n, k = L.Rk, alpha = L.ndim;
    int g1, g2, st0, st1;
    rcc_params(alpha, L.shape, &g1, &g2, &st0, &st1);
    OMG_TRY(h_alloc_t(h, &L.Rcc, (size_t)n));
    k_build_rcc<<<cdiv(n, OMG_TPB), OMG_TPB, 0, g.stream>>>(n, alpha, g1, g2, st0, st1, L.Rcc);
    OMG_TRY(h_alloc_t(h, &L.RTptr, (size_t)N + 1, true));
    int *cursor = nullptr;
    OMG_TRY(h_alloc_t(h, &cursor, (size_t)N + 1, true));
    k_rt_count<<<cdiv(n, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.Rcc, n, k, L.reg, L.RTptr);
    int total = 0;
    OMG_TRY(exclusive_scan_i32(L.RTptr, L.RTptr, N + 1, &total, g.stream));
    OMG_TRY(h_alloc_t(h, &L.RTcol, (size_t)std::max(total, 1)));
    k_rt_fill<<<cdiv(n, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.Rcc, n, k, L.reg, L.RTptr, cursor, L.RTcol);
    k_rt_sort<<<cdiv(N, OMG_TPB), OMG_TPB, 0, g.stream>>>(N, L.RTptr, L.RTcol);
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    h_free(h, cursor);
    return OMG_OK;
}

__global__ void k_r_fill(const int *__restrict__ Rcc, int n, RegR R, int *__restrict__ ptr, int *__restrict__ col,
                         double *__restrict__ val) {
    int r = blockIdx.x * OMG_TPB + threadIdx.x;
    if (r > n) return;
    ptr[r] = r * R.k;
    if (r == n) return;
    for (int q = 0; q < R.k; ++q) {
        col[r * R.k + q] = Rcc[r] + R.o[q];
        val[r * R.k + q] = R.w;
    }
}

static void rcc_params(int alpha, const int *shape, int *g1, int *g2, int *st0, int *st1) {
    int s3[3] = {1, 1, 1};
    for (int i = 0; i < alpha; ++i) s3[3 - alpha + i] = shape[i];
    *g1 = (s3[1] + 1) / 2;
    *g2 = (s3[2] + 1) / 2;
    *st1 = s3[2];
    *st0 = s3[1] * s3[2];
}

// operators.restriction(shape) built on the device as CSR (openmg/operators.py:15-89)
int device_restriction_csr(int ndim, const int64_t *shape, int64_t *n_rows, int64_t *nnz, int **dptr, int **dcol,
                           double **dval) {
    RDesc d;
    OMG_TRY(make_rdesc(ndim, shape, &d));
    *n_rows = d.n;
    *nnz = d.n * d.k;
    if (!dptr) return OMG_OK;
    if (d.n * d.k >= (1ll << 31)) return omg_set_error(OMG_EUNSUPPORTED, "restriction too large to export");
    RegR R;
    fill_reg(d, &R);
    int n = (int)d.n;
    int *Rcc = nullptr;
    CUDA_TRY(cudaMalloc(&Rcc, sizeof(int) * (size_t)n));
    CUDA_TRY(cudaMalloc(dptr, sizeof(int) * ((size_t)n + 1)));
    CUDA_TRY(cudaMalloc(dcol, sizeof(int) * (size_t)n * d.k));
    CUDA_TRY(cudaMalloc(dval, sizeof(double) * (size_t)n * d.k));
    int g1, g2, st0, st1;
    rcc_params(d.alpha, d.shape, &g1, &g2, &st0, &st1);
    k_build_rcc<<<cdiv(n, OMG_TPB), OMG_TPB, 0, g.stream>>>(n, d.alpha, g1, g2, st0, st1, Rcc);
    k_r_fill<<<cdiv(n + 1, OMG_TPB), OMG_TPB, 0, g.stream>>>(Rcc, n, R, *dptr, *dcol, *dval);
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    CUDA_TRY(cudaGetLastError());
    cudaFree(Rcc);
    return OMG_OK;
}

// ------------------------------------------------------------------ Galerkin R A R^T

struct BandRows {
    BandOp b;
    int n;
    template <class F>
    __device__ __forceinline__ void for_each(int i, F f) const {
        bool dd = false;
        for (int k = 0; k < b.nb; ++k) {
            if (!dd && b.off[k] > 0) {
                f(i, b.diag);
                dd = true;
            }
            int j = i + b.off[k];
            if (j >= 0 && j < n) f(j, b.coef[k]);
        }
        if (!dd) f(i, b.diag);
    }
};

struct CsrRows {
    const int *ptr;
    const int *col;
    const double *val;
    template <class F>
    __device__ __forceinline__ void for_each(int i, F f) const {
        for (int p = ptr[i]; p < ptr[i + 1]; ++p) f(col[p], val[p]);
    }
};

struct RegMap {
    RegR R;
    __device__ __forceinline__ int row_len(int) const { return R.k; }
    __device__ __forceinline__ int row_col(int I, int k) const { return reg_cc(R, I) + R.o[k]; }
    template <class F>
    __device__ __forceinline__ void for_each_t(int j, F f) const {
        f(reg_agg(R, j));
    }
};

struct ExpMap {
    RegR R;   // only k and o[] are used
    const int *Rcc, *RTptr, *RTcol;
    __device__ __forceinline__ int row_len(int) const { return R.k; }
    __device__ __forceinline__ int row_col(int I, int k) const { return Rcc[I] + R.o[k]; }
    template <class F>
    __device__ __forceinline__ void for_each_t(int j, F f) const {
        for (int p = RTptr[j]; p < RTptr[j + 1]; ++p) f(RTcol[p]);
    }
};

// One thread per coarse row I in [clo,chi).  Scratch is interleaved ([t*B + r]) so that
// neighbouring threads touch neighbouring words.  Products are accumulated in
// (fine row ascending, A-entry order, R^T order); exact zeros are dropped and the row is
// emitted with ascending columns (canonical form of scipy's (R*A)*R.T, which drops zero
// sums in csr_matmat — openmg/operators.py:184-186).
template <class ARows, class RMap, bool WRITE>
__global__ void __launch_bounds__(OMG_TPB) k_galerkin(ARows A, RMap M, int clo, int chi, int B, int *__restrict__ sJ,
                                                      double *__restrict__ sV, int cap, int *__restrict__ ocnt,
                                                      const int *__restrict__ optr, int *__restrict__ ocol,
                                                      double *__restrict__ oval, double w, int *__restrict__ overflow) {
    int r = blockIdx.x * OMG_TPB + threadIdx.x;
    int I = clo + r;
    if (I >= chi) return;
    int m = 0;
    int len = M.row_len(I);
    for (int k = 0; k < len; ++k) {
        int i = M.row_col(I, k);
        A.for_each(i, [&](int j, double a) {
            double v = (w * a) * w;
            M.for_each_t(j, [&](int J) {
                int t = 0;
                for (; t < m; ++t)
                    if (sJ[(size_t)t * B + r] == J) break;
                if (t < m) {
                    sV[(size_t)t * B + r] += v;
                } else if (m < cap) {
                    sJ[(size_t)m * B + r] = J;
                    sV[(size_t)m * B + r] = v;
                    ++m;
                } else {
                    *overflow = 1;
                }
            });
        });
    }
    int m2 = 0;
    for (int t = 0; t < m; ++t) {
        double v = sV[(size_t)t * B + r];
        if (v != 0.0) {
            int J = sJ[(size_t)t * B + r];
            sJ[(size_t)m2 * B + r] = J;
            sV[(size_t)m2 * B + r] = v;
            ++m2;
        }
    }
    if constexpr (!WRITE) {
        ocnt[I] = m2;
    } else {
    for (int a = 1; a < m2; ++a) {   // insertion sort by column
        int kJ = sJ[(size_t)a * B + r];
        double kV = sV[(size_t)a * B + r];
        int b2 = a - 1;
        while (b2 >= 0 && sJ[(size_t)b2 * B + r] > kJ) {
            sJ[(size_t)(b2 + 1) * B + r] = sJ[(size_t)b2 * B + r];
            sV[(size_t)(b2 + 1) * B + r] = sV[(size_t)b2 * B + r];
            --b2;
        }
        sJ[(size_t)(b2 + 1) * B + r] = kJ;
        sV[(size_t)(b2 + 1) * B + r] = kV;
    }
    int o = optr[I];
    for (int t = 0; t < m2; ++t) {
        ocol[o + t] = sJ[(size_t)t * B + r];
        oval[o + t] = sV[(size_t)t * B + r];
    }
    }
}

__global__ void k_max_rowlen(const int *__restrict__ ptr, int n, int *__restrict__ out) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    int v = i < n ? ptr[i + 1] - ptr[i] : 0;
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

static int device_max_rowlen(const int *ptr, int n, int *out_host) {
    int *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(d, 0, sizeof(int), g.stream));
    k_max_rowlen<<<cdiv(n, OMG_TPB), OMG_TPB, 0, g.stream>>>(ptr, n, d);
    CUDA_TRY(cudaMemcpyAsync(out_host, d, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    cudaFree(d);
    return OMG_OK;
}

template <class ARows, class RMap>
static int galerkin_typed(omg_hierarchy *h, const ARows &A, const RMap &M, int nc, int cap_per_row, double w,
                          Level &C) {
    const int64_t SCRATCH_ENTRIES = 48ll << 20;   // <= 576 MB of scratch per batch
    int B = (int)std::min<int64_t>(nc, std::max<int64_t>(SCRATCH_ENTRIES / cap_per_row, 1024));
    B = (B + OMG_TPB - 1) / OMG_TPB * OMG_TPB;
    int *sJ = nullptr, *ocnt = nullptr, *overflow = nullptr;
    double *sV = nullptr;
    OMG_TRY(h_alloc_t(h, &sJ, (size_t)B * cap_per_row));
    OMG_TRY(h_alloc_t(h, &sV, (size_t)B * cap_per_row));
    OMG_TRY(h_alloc_t(h, &C.ptr, (size_t)nc + 1, true));
    OMG_TRY(h_alloc_t(h, &overflow, 1, true));
    ocnt = C.ptr;
    for (int clo = 0; clo < nc; clo += B) {
        int chi = std::min(nc, clo + B);
        k_galerkin<ARows, RMap, false><<<cdiv(chi - clo, OMG_TPB), OMG_TPB, 0, g.stream>>>(
            A, M, clo, chi, B, sJ, sV, cap_per_row, ocnt, nullptr, nullptr, nullptr, w, overflow);
    }
    int total = 0, ovf = 0;
    OMG_TRY(exclusive_scan_i32(C.ptr, C.ptr, nc + 1, &total, g.stream));
    CUDA_TRY(cudaMemcpy(&ovf, overflow, sizeof(int), cudaMemcpyDeviceToHost));
    if (ovf) return omg_set_error(OMG_ECUDA, "internal: Galerkin scratch overflow");
    C.nnz = total;
    OMG_TRY(h_alloc_t(h, &C.col, (size_t)std::max(total, 1)));
    OMG_TRY(h_alloc_t(h, &C.val, (size_t)std::max(total, 1)));
    for (int clo = 0; clo < nc; clo += B) {
        int chi = std::min(nc, clo + B);
        k_galerkin<ARows, RMap, true><<<cdiv(chi - clo, OMG_TPB), OMG_TPB, 0, g.stream>>>(
            A, M, clo, chi, B, sJ, sV, cap_per_row, nullptr, C.ptr, C.col, C.val, w, overflow);
    }
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    CUDA_TRY(cudaGetLastError());
    h_free(h, sJ);
    h_free(h, sV);
    h_free(h, overflow);
    return OMG_OK;
}

static int galerkin_level(omg_hierarchy *h, int l) {
    Level &F = h->lv[l];
    Level &C = h->lv[l + 1];
    int maxrow = 0, maxmult = 1;
    bool band_rows = (F.ptr == nullptr);
    if (band_rows)
        maxrow = F.band.nb + 1;
    else
        OMG_TRY(device_max_rowlen(F.ptr, F.n, &maxrow));
    if (!F.regular) {
        OMG_TRY(build_explicit_R(h, F));
        OMG_TRY(device_max_rowlen(F.RTptr, F.n, &maxmult));
        maxmult = std::max(maxmult, 1);
    }
    int64_t cap64 = (int64_t)F.Rk * maxrow * maxmult;
    if (cap64 > (1 << 20)) return omg_set_error(OMG_EUNSUPPORTED, "Galerkin product: rows too dense (%lld products per coarse row)", (long long)cap64);
    int cap = (int)std::max<int64_t>(cap64, 1);
    if (F.regular) {
        RegMap M{F.reg};
        if (band_rows) {
            BandRows A{F.band, F.n};
            OMG_TRY(galerkin_typed(h, A, M, F.nc, cap, F.Rw, C));
        } else {
            CsrRows A{F.ptr, F.col, F.val};
            OMG_TRY(galerkin_typed(h, A, M, F.nc, cap, F.Rw, C));
        }
    } else {
        ExpMap M{F.reg, F.Rcc, F.RTptr, F.RTcol};
        if (band_rows) {
            BandRows A{F.band, F.n};
            OMG_TRY(galerkin_typed(h, A, M, F.nc, cap, F.Rw, C));
        } else {
            CsrRows A{F.ptr, F.col, F.val};
            OMG_TRY(galerkin_typed(h, A, M, F.nc, cap, F.Rw, C));
        }
    }
    return OMG_OK;
}

// ------------------------------------------------------------------ diagonal

__global__ void k_extract_diag(const int *__restrict__ ptr, const int *__restrict__ col,
                               const double *__restrict__ val, int n, int row0, double *__restrict__ diag) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= n) return;
    double d = 0.0;
    for (int p = ptr[i]; p < ptr[i + 1]; ++p)
        if (col[p] == i + row0) d += val[p];     // duplicates sum, like scipy's A[i,i]
    diag[i] = d;
}

// ------------------------------------------------------------------ band detection

struct BandSig {        // stencil incl. the diagonal, ascending offsets
    int m;
    int off[OMG_MAXBAND + 1];
    double coef[OMG_MAXBAND + 1];
};

// bit i of mask = row i does NOT equal the truncated stencil
__global__ void __launch_bounds__(OMG_TPB) k_flag_rows(const int *__restrict__ ptr, const int *__restrict__ col,
                                                       const double *__restrict__ val, int n, BandSig s,
                                                       unsigned *__restrict__ mask, int *__restrict__ wcnt) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    bool bad = false;
    if (i < n) {
        int p = ptr[i], pe = ptr[i + 1];
        for (int k = 0; k < s.m; ++k) {
            int j = i + s.off[k];
            if (j < 0 || j >= n) continue;
            if (p >= pe || col[p] != j || val[p] != s.coef[k]) {
                bad = true;
                break;
            }
            ++p;
        }
        if (p != pe) bad = true;
    }
    unsigned m = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && (i >> 5) < (n + 31) / 32) {
        mask[i >> 5] = m;
        wcnt[i >> 5] = __popc(m);
    }
}

__global__ void __launch_bounds__(OMG_TPB) k_exc_rows(const unsigned *__restrict__ mask, const int *__restrict__ wpre,
                                                      const int *__restrict__ ptr, int n, int *__restrict__ rows,
                                                      int *__restrict__ cnt) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= n) return;
    unsigned w = mask[i >> 5];
    if ((w >> (i & 31)) & 1u) {
        int s = wpre[i >> 5] + __popc(w & ((1u << (i & 31)) - 1u));
        rows[s] = i;
        cnt[s] = ptr[i + 1] - ptr[i];
    }
}

__global__ void __launch_bounds__(OMG_TPB) k_exc_fill(const int *__restrict__ rows, int nexc,
                                                      const int *__restrict__ ptr, const int *__restrict__ col,
                                                      const double *__restrict__ val, const int *__restrict__ eptr,
                                                      int row0, int *__restrict__ ecol, double *__restrict__ eval,
                                                      double *__restrict__ ediag) {
    int s = blockIdx.x * OMG_TPB + threadIdx.x;
    if (s >= nexc) return;
    int i = rows[s];
    int o = eptr[s];
    double d = 0.0;
    for (int p = ptr[i]; p < ptr[i + 1]; ++p, ++o) {
        ecol[o] = col[p] - row0;
        eval[o] = val[p];
        if (col[p] == i + row0) d += val[p];
    }
    ediag[s] = d;
}

// max |column - row| over the exception rows: a slab level exchanges halos sized from the BAND reach, so exception
// rows reaching further (possible when the level comes from a general input matrix) rule out sharding it
__global__ void k_exc_reach(const int *__restrict__ rows, int nexc, const int *__restrict__ eptr,
                            const int *__restrict__ ecol, int *__restrict__ reach) {
    int s = blockIdx.x * OMG_TPB + threadIdx.x;
    if (s >= nexc) return;
    int i = rows[s], m = 0;
    for (int p = eptr[s]; p < eptr[s + 1]; ++p) m = max(m, abs(ecol[p] - i));
    atomicMax(reach, m);
}

// global coarse row of every exception row -> bit mask over coarse rows (global index - 0 on one GPU)
__global__ void k_mark_crows(const int *__restrict__ rows, int nexc, RegR R, int frow0, unsigned *__restrict__ cmask) {
    int s = blockIdx.x * OMG_TPB + threadIdx.x;
    if (s >= nexc) return;
    int I = reg_agg(R, rows[s] + frow0);
    atomicOr(cmask + (I >> 5), 1u << (I & 31));
}
__global__ void k_popc_words(const unsigned *__restrict__ mask, int nw, int *__restrict__ cnt) {
    int w = blockIdx.x * OMG_TPB + threadIdx.x;
    if (w < nw) cnt[w] = __popc(mask[w]);
}
__global__ void k_compact_bits(const unsigned *__restrict__ mask, const int *__restrict__ wpre, int n,
                               int *__restrict__ out) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= n) return;
    unsigned w = mask[i >> 5];
    if ((w >> (i & 31)) & 1u) out[wpre[i >> 5] + __popc(w & ((1u << (i & 31)) - 1u))] = i;
}

__global__ void k_diag_differs(const double *__restrict__ ediag, int nexc, double d, int *__restrict__ flag) {
    int s = blockIdx.x * OMG_TPB + threadIdx.x;
    if (s < nexc && ediag[s] != d) *flag = 1;
}

// ---- positional stencil classes (3-D Galerkin levels)

struct ClsFlat {     // host/device helper: flat offsets + deltas per class
    int ntap[9];
    int off[9][OMG_CLS_TAPS];
    double coef[9][OMG_CLS_TAPS];
};

// Every row must equal band + class correction (truncated to [0,n), exact zeros absent); interior-class rows must
// not be exception rows.  One thread per row.
__global__ void __launch_bounds__(OMG_TPB) k_check_classes(const int *__restrict__ ptr, const int *__restrict__ col,
                                                           const double *__restrict__ val, const unsigned *__restrict__ mask,
                                                           int n, BandOp b, ClsFlat c, int S1, int NY,
                                                           int *__restrict__ bad) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= n) return;
    int x = i % S1, y = (i / S1) % NY;
    int cls = 3 * (y == 0 ? 0 : (y == NY - 1 ? 2 : 1)) + (x == 0 ? 0 : (x == S1 - 1 ? 2 : 1));
    bool exc = (mask[i >> 5] >> (i & 31)) & 1u;
    if (cls == 4) {
        if (exc) *bad = 1;
        return;
    }
    if (!exc && c.ntap[cls] == 0) return;
    // expected value at flat offset o
    auto expected = [&](int o, bool &known) {
        double v = 0.0;
        known = false;
        if (o == 0) {
            v = b.diag;
            known = true;
        }
        for (int k = 0; k < b.nb; ++k)
            if (b.off[k] == o) {
                v += b.coef[k];
                known = true;
            }
        for (int t = 0; t < c.ntap[cls]; ++t)
            if (c.off[cls][t] == o) {
                v += c.coef[cls][t];
                known = true;
            }
        return v;
    };
    int seen = 0;
    for (int p = ptr[i]; p < ptr[i + 1]; ++p) {
        bool known;
        double v = expected(col[p] - i, known);
        if (!known || v != val[p]) {
            *bad = 1;
            return;
        }
        ++seen;
    }
    // no expected nonzero in-range entry may be missing
    int want = 0;
    bool known;
    if (expected(0, known) != 0.0) ++want;
    for (int k = 0; k < b.nb; ++k) {
        int j = i + b.off[k];
        if (j >= 0 && j < n && expected(b.off[k], known) != 0.0) ++want;
    }
    for (int t = 0; t < c.ntap[cls]; ++t) {
        int o = c.off[cls][t];
        bool in_band = (o == 0);
        for (int k = 0; k < b.nb; ++k) in_band = in_band || b.off[k] == o;
        int j = i + o;
        if (!in_band && j >= 0 && j < n && expected(o, known) != 0.0) ++want;
    }
    if (want != seen) *bad = 1;
}

static int detect_classes(omg_hierarchy *h, Level &L) {
    L.classed = false;
    if (L.kind != OMG_KIND_BAND_EXC || L.band.nb != 6 || !L.ptr || getenv("OMG_NO_CLASSES")) return OMG_OK;
    const BandOp &B = L.band;
    if (B.off[3] != 1 || B.off[2] != -1 || B.off[4] != -B.off[1] || B.off[5] != -B.off[0]) return OMG_OK;
    int S1 = B.off[4], S2 = B.off[5];
    if (S1 < 8 || S2 % S1 != 0 || L.n % S2 != 0) return OMG_OK;
    int NY = S2 / S1, NZ = L.n / S2;
    if (NY < 4 || NZ < 3) return OMG_OK;
    ClsFlat cf{};
    std::vector<int> hp(2);
    for (int cy = 0; cy < 3; ++cy)
        for (int cx = 0; cx < 3; ++cx) {
            int cls = 3 * cy + cx;
            cf.ntap[cls] = 0;
            if (cls == 4) continue;
            int x = cx == 0 ? 0 : (cx == 2 ? S1 - 1 : S1 / 2);
            int y = cy == 0 ? 0 : (cy == 2 ? NY - 1 : NY / 2);
            int r = (NZ / 2) * S2 + y * S1 + x;
            CUDA_TRY(cudaMemcpy(hp.data(), L.ptr + r, 2 * sizeof(int), cudaMemcpyDeviceToHost));
            int len = hp[1] - hp[0];
            if (len <= 0 || len > 32) return OMG_OK;
            std::vector<int> hc(len);
            std::vector<double> hv(len);
            CUDA_TRY(cudaMemcpy(hc.data(), L.col + hp[0], len * sizeof(int), cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(hv.data(), L.val + hp[0], len * sizeof(double), cudaMemcpyDeviceToHost));
            // delta = row - band over the union of offsets
            std::map<int, double> delta;
            for (int t = 0; t < len; ++t) delta[hc[t] - r] += hv[t];
            delta[0] -= B.diag;
            for (int k = 0; k < B.nb; ++k) delta[B.off[k]] -= B.coef[k];
            for (auto &kv : delta) {
                if (kv.second == 0.0) continue;
                if (kv.first == 0) return OMG_OK;                 // classes need a uniform diagonal
                int o = kv.first;
                int dz = o > S2 / 2 ? 1 : (o < -S2 / 2 ? -1 : 0);
                int so = o - dz * S2;
                if (so < -S1 || so > S1) return OMG_OK;          // outside the staged +-1 rows
                if (cf.ntap[cls] >= OMG_CLS_TAPS) return OMG_OK;
                cf.off[cls][cf.ntap[cls]] = o;
                cf.coef[cls][cf.ntap[cls]] = kv.second;
                cf.ntap[cls]++;
            }
        }
    int *bad = nullptr, hb = 0;
    OMG_TRY(h_alloc_t(h, &bad, 1, true));
    k_check_classes<<<cdiv(L.n, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.ptr, L.col, L.val, L.exc_mask, L.n, B, cf, S1, NY, bad);
    CUDA_TRY(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    CUDA_TRY(cudaGetLastError());
    h_free(h, bad);
    if (hb) return OMG_OK;
    for (int c = 0; c < 9; ++c) {
        L.cls.ntap[c] = cf.ntap[c];
        for (int t = 0; t < cf.ntap[c]; ++t) {
            int o = cf.off[c][t];
            int dz = o > S2 / 2 ? 1 : (o < -S2 / 2 ? -1 : 0);
            L.cls.dz[c][t] = dz;
            L.cls.soff[c][t] = o - dz * S2;
            L.cls.coef[c][t] = cf.coef[c][t];
        }
    }
    L.classed = true;
    return OMG_OK;
}

// 2-D Galerkin levels: the rows of the first and of the last grid column deviate from the band by three taps each
// (offsets -1, -(N+1), +(N-1) resp. +1, +(N+1), -(N-1)); everything else is the truncated band.
static int detect_classes2(omg_hierarchy *h, Level &L) {
    L.classed2 = false;
    if (L.kind != OMG_KIND_BAND_EXC || L.band.nb != 6 || !L.ptr || getenv("OMG_NO_CLASSES")) return OMG_OK;
    const BandOp &B = L.band;
    int N = B.off[4];
    if (B.off[2] != -1 || B.off[3] != 1 || B.off[5] != N + 1 || B.off[1] != -N || B.off[0] != -(N + 1)) return OMG_OK;
    if (N < 8 || L.n % N != 0 || L.n / N < 4) return OMG_OK;
    int NY = L.n / N;
    const int allowed[2][3] = {{-1, -(N + 1), N - 1}, {1, N + 1, -(N - 1)}};
    double dl[2][3] = {{0, 0, 0}, {0, 0, 0}};
    for (int side = 0; side < 2; ++side) {
        int r = (NY / 2) * N + (side ? N - 1 : 0);
        int hp[2];
        CUDA_TRY(cudaMemcpy(hp, L.ptr + r, 2 * sizeof(int), cudaMemcpyDeviceToHost));
        int len = hp[1] - hp[0];
        if (len <= 0 || len > 32) return OMG_OK;
        std::vector<int> hc(len);
        std::vector<double> hv(len);
        CUDA_TRY(cudaMemcpy(hc.data(), L.col + hp[0], len * sizeof(int), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(hv.data(), L.val + hp[0], len * sizeof(double), cudaMemcpyDeviceToHost));
        std::map<int, double> delta;
        for (int t = 0; t < len; ++t) delta[hc[t] - r] += hv[t];
        delta[0] -= B.diag;
        for (int k = 0; k < B.nb; ++k) delta[B.off[k]] -= B.coef[k];
        for (auto &kv : delta) {
            if (kv.second == 0.0) continue;
            int t = 0;
            while (t < 3 && allowed[side][t] != kv.first) ++t;
            if (t == 3) return OMG_OK;
            dl[side][t] = kv.second;
        }
    }
    ClsFlat cf{};
    for (int cy = 0; cy < 3; ++cy)
        for (int side = 0; side < 2; ++side) {
            int cls = 3 * cy + (side ? 2 : 0);
            for (int t = 0; t < 3; ++t)
                if (dl[side][t] != 0.0) {
                    cf.off[cls][cf.ntap[cls]] = allowed[side][t];
                    cf.coef[cls][cf.ntap[cls]] = dl[side][t];
                    cf.ntap[cls]++;
                }
        }
    int *bad = nullptr, hb = 0;
    OMG_TRY(h_alloc_t(h, &bad, 1, true));
    k_check_classes<<<cdiv(L.n, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.ptr, L.col, L.val, L.exc_mask, L.n, B, cf, N, NY, bad);
    CUDA_TRY(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    CUDA_TRY(cudaGetLastError());
    h_free(h, bad);
    if (hb) return OMG_OK;
    for (int t = 0; t < 3; ++t) {
        L.c2l[t] = dl[0][t];
        L.c2r[t] = dl[1][t];
    }
    L.classed2 = true;
    return OMG_OK;
}

static int detect_band(omg_hierarchy *h, Level &L) {
    L.kind = OMG_KIND_CSR;
    if (h->flags & OMG_FLAG_FORCE_CSR) return OMG_OK;
    int n = L.n;
    if (n < 32 || L.ptr == nullptr) return OMG_OK;
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    // --- candidate stencil: majority signature of rows sampled from the middle half
    const int S = 41;
    std::map<std::vector<std::pair<int, double>>, int> votes;
    std::vector<int> hp(2);
    for (int s = 0; s < S; ++s) {
        int r = n / 4 + (int)(((int64_t)(n / 2) * s) / S) + (s * 7919) % std::max(1, n / (2 * S));
        r = std::min(std::max(r, 0), n - 1);
        CUDA_TRY(cudaMemcpy(hp.data(), L.ptr + r, 2 * sizeof(int), cudaMemcpyDeviceToHost));
        int len = hp[1] - hp[0];
        if (len <= 0 || len > OMG_MAXBAND + 1) continue;
        std::vector<int> hc(len);
        std::vector<double> hv(len);
        CUDA_TRY(cudaMemcpy(hc.data(), L.col + hp[0], len * sizeof(int), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(hv.data(), L.val + hp[0], len * sizeof(double), cudaMemcpyDeviceToHost));
        std::vector<std::pair<int, double>> sig;
        bool has_diag = false;
        for (int t = 0; t < len; ++t) {
            sig.push_back({hc[t] - r, hv[t]});
            has_diag = has_diag || hc[t] == r;
        }
        if (!has_diag) continue;
        // rows near the global ends are truncated; only full-reach rows vote
        if (r + sig.front().first < 0 || r + sig.back().first >= n) continue;
        votes[sig]++;
    }
    if (votes.empty()) return OMG_OK;
    auto best = votes.begin();
    for (auto it = votes.begin(); it != votes.end(); ++it)
        if (it->second > best->second) best = it;
    if (best->second * 3 < S) return OMG_OK;
    BandSig sig{};
    sig.m = (int)best->first.size();
    BandOp band{};
    band.nb = 0;
    for (int t = 0; t < sig.m; ++t) {
        sig.off[t] = best->first[t].first;
        sig.coef[t] = best->first[t].second;
        if (sig.off[t] == 0)
            band.diag = sig.coef[t];
        else {
            band.off[band.nb] = sig.off[t];
            band.coef[band.nb] = sig.coef[t];
            band.nb++;
        }
    }
    if (band.diag == 0.0) return OMG_OK;
    // --- flag deviating rows
    int nw = (n + 31) / 32;
    unsigned *mask = nullptr;
    int *wpre = nullptr;
    OMG_TRY(h_alloc_t(h, &mask, (size_t)nw + 1, true));
    OMG_TRY(h_alloc_t(h, &wpre, (size_t)nw + 1, true));
    k_flag_rows<<<cdiv(n, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.ptr, L.col, L.val, n, sig, mask, wpre);
    int nexc = 0;
    OMG_TRY(exclusive_scan_i32(wpre, wpre, nw + 1, &nexc, g.stream));
    if ((double)nexc > 0.30 * n) {
        h_free(h, mask);
        h_free(h, wpre);
        return OMG_OK;
    }
    L.band = band;
    L.nexc = nexc;
    if (nexc == 0) {
        h_free(h, mask);
        h_free(h, wpre);
        L.kind = OMG_KIND_BAND;
        return OMG_OK;
    }
    int *rows = nullptr, *eptr = nullptr;
    OMG_TRY(h_alloc_t(h, &rows, (size_t)nexc));
    OMG_TRY(h_alloc_t(h, &eptr, (size_t)nexc + 1, true));
    k_exc_rows<<<cdiv(n, OMG_TPB), OMG_TPB, 0, g.stream>>>(mask, wpre, L.ptr, n, rows, eptr);
    int ennz = 0;
    OMG_TRY(exclusive_scan_i32(eptr, eptr, nexc + 1, &ennz, g.stream));
    OMG_TRY(h_alloc_t(h, &L.exc_col, (size_t)std::max(ennz, 1)));
    OMG_TRY(h_alloc_t(h, &L.exc_val, (size_t)std::max(ennz, 1)));
    OMG_TRY(h_alloc_t(h, &L.exc_diag, (size_t)nexc));
    k_exc_fill<<<cdiv(nexc, OMG_TPB), OMG_TPB, 0, g.stream>>>(rows, nexc, L.ptr, L.col, L.val, eptr, L.row0,
                                                              L.exc_col, L.exc_val, L.exc_diag);
    {   // are all exception diagonals equal to the stencil diagonal? (true for the Poisson hierarchies)
        // and how far do the exception rows reach?
        int *flag = nullptr;
        OMG_TRY(h_alloc_t(h, &flag, 2, true));
        k_diag_differs<<<cdiv(nexc, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.exc_diag, nexc, band.diag, flag);
        k_exc_reach<<<cdiv(nexc, OMG_TPB), OMG_TPB, 0, g.stream>>>(rows, nexc, eptr, L.exc_col, flag + 1);
        int f[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(f, flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
        CUDA_TRY(cudaStreamSynchronize(g.stream));
        L.exc_diag_uniform = (f[0] == 0);
        L.exc_reach = f[1];
        h_free(h, flag);
    }
    L.exc_rows = rows;
    L.exc_mask = mask;
    L.exc_wpre = wpre;
    L.exc_ptr = eptr;
    L.kind = OMG_KIND_BAND_EXC;
    // coarse rows touched by exception rows (fix-up list of the fused residual+restriction)
    if (L.hasR && L.regular) {
        int ncw = (L.nc + 31) / 32;
        unsigned *cmask = nullptr;
        int *cpre = nullptr;
        OMG_TRY(h_alloc_t(h, &cmask, (size_t)ncw + 1, true));
        OMG_TRY(h_alloc_t(h, &cpre, (size_t)ncw + 1, true));
        k_mark_crows<<<cdiv(nexc, OMG_TPB), OMG_TPB, 0, g.stream>>>(rows, nexc, L.reg, L.row0, cmask);
        k_popc_words<<<cdiv(ncw, OMG_TPB), OMG_TPB, 0, g.stream>>>(cmask, ncw, cpre);
        int ncr = 0;
        OMG_TRY(exclusive_scan_i32(cpre, cpre, ncw + 1, &ncr, g.stream));
        OMG_TRY(h_alloc_t(h, &L.exc_crows, (size_t)std::max(ncr, 1)));
        k_compact_bits<<<cdiv(L.nc, OMG_TPB), OMG_TPB, 0, g.stream>>>(cmask, cpre, L.nc, L.exc_crows);
        CUDA_TRY(cudaStreamSynchronize(g.stream));
        L.nexc_crows = ncr;
        h_free(h, cmask);
        h_free(h, cpre);
    }
    if (L.band.nb == 6 && L.band.off[5] == L.band.off[4] + 1) return detect_classes2(h, L);
    return detect_classes(h, L);
}

// ------------------------------------------------------------------ band -> CSR (export of a band-only level 0)

__global__ void k_band_count(BandOp b, int n, int *__restrict__ cnt) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= n) return;
    int c = 1;
    for (int k = 0; k < b.nb; ++k) {
        int j = i + b.off[k];
        c += (j >= 0 && j < n);
    }
    cnt[i] = c;
}

__global__ void k_band_fill(BandOp b, int n, const int *__restrict__ ptr, int *__restrict__ col,
                            double *__restrict__ val) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= n) return;
    int o = ptr[i];
    BandRows rows{b, n};
    rows.for_each(i, [&](int j, double a) {
        col[o] = j;
        val[o] = a;
        ++o;
    });
}

// temp CSR of a level that has none (caller frees with cudaFree)
int materialize_level_csr(omg_hierarchy *h, const Level &L, int **ptr, int **col, double **val, int64_t *nnz) {
    (void)h;
    int n = L.n;
    CUDA_TRY(cudaMalloc(ptr, sizeof(int) * ((size_t)n + 1)));
    CUDA_TRY(cudaMemsetAsync(*ptr, 0, sizeof(int) * ((size_t)n + 1), g.stream));
    k_band_count<<<cdiv(n, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.band, n, *ptr);
    int total = 0;
    OMG_TRY(exclusive_scan_i32(*ptr, *ptr, n + 1, &total, g.stream));
    *nnz = total;
    if (!col) return OMG_OK;
    CUDA_TRY(cudaMalloc(col, sizeof(int) * (size_t)std::max(total, 1)));
    CUDA_TRY(cudaMalloc(val, sizeof(double) * (size_t)std::max(total, 1)));
    k_band_fill<<<cdiv(n, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.band, n, *ptr, *col, *val);
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return OMG_OK;
}

// ------------------------------------------------------------------ coarse factor: dense inverse by Gauss-Jordan

__global__ void k_dense_from_csr(const int *__restrict__ ptr, const int *__restrict__ col,
                                 const double *__restrict__ val, int n, double *__restrict__ M) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= n) return;
    for (int p = ptr[i]; p < ptr[i + 1]; ++p) M[(size_t)i * n + col[p]] += val[p];
}

// partial pivoting: piv[k] = argmax_{i>=k} |M[i][k]|
__global__ void __launch_bounds__(1024) k_gj_pivot(const double *__restrict__ M, int n, int k, int *__restrict__ piv,
                                                   int *__restrict__ singular) {
    __shared__ double sv[1024];
    __shared__ int si[1024];
    double best = -1.0;
    int bi = k;
    for (int i = k + threadIdx.x; i < n; i += blockDim.x) {
        double v = fabs(M[(size_t)i * n + k]);
        if (v > best) {
            best = v;
            bi = i;
        }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            double v2 = sv[threadIdx.x + o];
            int i2 = si[threadIdx.x + o];
            if (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x])) {
                sv[threadIdx.x] = v2;
                si[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        piv[k] = si[0];
        if (!(sv[0] > 0.0)) *singular = 1;
    }
}

__global__ void k_gj_swap(double *__restrict__ M, int n, int k, const int *__restrict__ piv) {
    int p = piv[k];
    if (p == k) return;
    int j = blockIdx.x * OMG_TPB + threadIdx.x;
    if (j >= n) return;
    double a = M[(size_t)k * n + j], b = M[(size_t)p * n + j];
    M[(size_t)k * n + j] = b;
    M[(size_t)p * n + j] = a;
}

__global__ void k_gj_col(double *__restrict__ M, int n, int k, double *__restrict__ colk) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= n) return;
    colk[i] = M[(size_t)i * n + k];
    M[(size_t)i * n + k] = (i == k) ? 1.0 : 0.0;
}

__global__ void k_gj_scale(double *__restrict__ M, int n, int k, const double *__restrict__ colk) {
    int j = blockIdx.x * OMG_TPB + threadIdx.x;
    if (j >= n) return;
    M[(size_t)k * n + j] = M[(size_t)k * n + j] / colk[k];
}

__global__ void __launch_bounds__(OMG_TPB) k_gj_elim(double *__restrict__ M, int n, int k,
                                                     const double *__restrict__ colk) {
    int j = blockIdx.x * OMG_TPB + threadIdx.x;
    if (j >= n) return;
    double pk = M[(size_t)k * n + j];
    int i0 = blockIdx.y * 16;
#pragma unroll 4
    for (int i = i0; i < min(i0 + 16, n); ++i) {
        if (i == k) continue;
        double f = colk[i];
        if (f != 0.0) M[(size_t)i * n + j] -= f * pk;
    }
}

__global__ void k_gj_unperm(double *__restrict__ M, int n, const int *__restrict__ piv) {
    int r = blockIdx.x * OMG_TPB + threadIdx.x;
    if (r >= n) return;
    double *row = M + (size_t)r * n;
    for (int k = n - 1; k >= 0; --k) {
        int p = piv[k];
        if (p != k) {
            double a = row[k];
            row[k] = row[p];
            row[p] = a;
        }
    }
}

// ---- blocked Gauss-Jordan inversion without pivoting (block 32): 4 launches per 32 columns instead
// of 5 per column.  Used first; if a diagonal block meets a (near-)zero pivot the pivoted unblocked
// elimination above is the fallback.  (The Galerkin operators of definite problems need no pivoting.)
#define BGJ 32

__global__ void k_bgj_pad(double *__restrict__ W, int n, int np) {
    int i = n + blockIdx.x * OMG_TPB + threadIdx.x;
    if (i < np) W[(size_t)i * np + i] = 1.0;
}

// D = inverse of the diagonal block k (in shared memory), written to Dbuf; C = column panel copy
__global__ void __launch_bounds__(BGJ *BGJ) k_bgj_diag(const double *__restrict__ W, int np, int k,
                                                        double *__restrict__ Dbuf, int *__restrict__ bad) {
    __shared__ double S[BGJ][BGJ + 1], V[BGJ][BGJ + 1];
    int r = threadIdx.y, c = threadIdx.x;
    S[r][c] = W[(size_t)(k * BGJ + r) * np + k * BGJ + c];
    V[r][c] = (r == c) ? 1.0 : 0.0;
    __syncthreads();
    for (int p = 0; p < BGJ; ++p) {
        double pv = S[p][p];
        if (r == 0 && c == 0 && !(fabs(pv) > 1e-280)) *bad = 1;
        __syncthreads();
        if (r == p) {
            S[p][c] = S[p][c] / pv;
            V[p][c] = V[p][c] / pv;
        }
        __syncthreads();
        double f = S[r][p];
        __syncthreads();
        if (r != p) {
            S[r][c] -= f * S[p][c];
            V[r][c] -= f * V[p][c];
        }
        __syncthreads();
    }
    Dbuf[r * BGJ + c] = V[r][c];
}

__global__ void k_bgj_savecol(const double *__restrict__ W, int np, int k, double *__restrict__ Cbuf) {
    int i = blockIdx.x * 8 + threadIdx.y, c = threadIdx.x;      // block (32, 8)
    if (i < np) Cbuf[(size_t)i * BGJ + c] = W[(size_t)i * np + k * BGJ + c];
}

// row panel: W[K, J] = D * W[K, J] (J != K), W[K, K] = D.  One CTA (32x32) per 32-column tile.
__global__ void __launch_bounds__(BGJ *BGJ) k_bgj_rowpanel(double *__restrict__ W, int np, int k,
                                                            const double *__restrict__ Dbuf) {
    __shared__ double D[BGJ][BGJ + 1], T[BGJ][BGJ + 1];
    int r = threadIdx.y, c = threadIdx.x, j = blockIdx.x;
    D[r][c] = Dbuf[r * BGJ + c];
    double *t = W + (size_t)(k * BGJ + r) * np + j * BGJ + c;
    T[r][c] = *t;
    __syncthreads();
    if (j == k) {
        *t = D[r][c];
        return;
    }
    double acc = 0.0;
#pragma unroll 8
    for (int q = 0; q < BGJ; ++q) acc += D[r][q] * T[q][c];
    *t = acc;
}

// trailing update: W[I, J] -= C[I] * W[K, J]   (rows outside block K; for J == K the old value counts as 0,
// giving W[I, K] = -C[I] * D).  64x64 tile per CTA (16x16 threads, 4x4 outputs each).
__global__ void __launch_bounds__(256) k_bgj_trailing(double *__restrict__ W, int np, int k,
                                                       const double *__restrict__ Cbuf) {
    __shared__ double Cs[64][BGJ + 1], Ps[BGJ][64 + 1];
    int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    for (int t = threadIdx.x; t < 64 * BGJ; t += 256) {
        int rr = t / BGJ, cc = t % BGJ;
        Cs[rr][cc] = Cbuf[(size_t)(i0 + rr) * BGJ + cc];
        int pr = t / 64, pc = t % 64;
        Ps[pr][pc] = W[(size_t)(k * BGJ + pr) * np + j0 + pc];
    }
    __syncthreads();
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 4
    for (int q = 0; q < BGJ; ++q) {
        double cv[4], pv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) cv[a] = Cs[ty * 4 + a][q];
#pragma unroll
        for (int b = 0; b < 4; ++b) pv[b] = Ps[q][tx * 4 + b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] += cv[a] * pv[b];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        int i = i0 + ty * 4 + a;
        if (i / BGJ == k) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int j = j0 + tx * 4 + b;
            double *w = W + (size_t)i * np + j;
            double base = (j / BGJ == k) ? 0.0 : *w;
            *w = base - acc[a][b];
        }
    }
}

__global__ void k_bgj_copyout(const double *__restrict__ W, int np, int n, double *__restrict__ out) {
    int j = blockIdx.x * OMG_TPB + threadIdx.x, i = blockIdx.y;
    if (j < n) out[(size_t)i * n + j] = W[(size_t)i * np + j];
}

__global__ void k_dense_from_csr_ld(const int *__restrict__ ptr, const int *__restrict__ col,
                                    const double *__restrict__ val, int n, int ld, double *__restrict__ M) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= n) return;
    for (int p = ptr[i]; p < ptr[i + 1]; ++p) M[(size_t)i * ld + col[p]] += val[p];
}

// returns OMG_OK with *ok = false when a pivot was unusable (caller falls back)
static int coarse_factor_blocked(omg_hierarchy *h, Level &L, int n, bool *ok) {
    int np = (n + 63) / 64 * 64;
    double *W = nullptr, *Dbuf = nullptr, *Cbuf = nullptr;
    int *bad = nullptr;
    OMG_TRY(h_alloc_t(h, &W, (size_t)np * np, true));
    OMG_TRY(h_alloc_t(h, &Dbuf, (size_t)BGJ * BGJ));
    OMG_TRY(h_alloc_t(h, &Cbuf, (size_t)np * BGJ));
    OMG_TRY(h_alloc_t(h, &bad, 1, true));
    k_dense_from_csr_ld<<<cdiv(n, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.ptr, L.col, L.val, n, np, W);
    if (np > n) k_bgj_pad<<<cdiv(np - n, OMG_TPB), OMG_TPB, 0, g.stream>>>(W, n, np);
    int nblk = np / BGJ;
    for (int k = 0; k < nblk; ++k) {
        k_bgj_diag<<<1, dim3(BGJ, BGJ), 0, g.stream>>>(W, np, k, Dbuf, bad);
        k_bgj_savecol<<<cdiv(np, 8), dim3(BGJ, 8), 0, g.stream>>>(W, np, k, Cbuf);
        k_bgj_rowpanel<<<nblk, dim3(BGJ, BGJ), 0, g.stream>>>(W, np, k, Dbuf);
        k_bgj_trailing<<<dim3(np / 64, np / 64), 256, 0, g.stream>>>(W, np, k, Cbuf);
    }
    k_bgj_copyout<<<dim3(cdiv(n, OMG_TPB), n), OMG_TPB, 0, g.stream>>>(W, np, n, h->Ainv);
    int b = 0;
    CUDA_TRY(cudaMemcpyAsync(&b, bad, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    CUDA_TRY(cudaGetLastError());
    h_free(h, W);
    h_free(h, Dbuf);
    h_free(h, Cbuf);
    h_free(h, bad);
    *ok = (b == 0);
    return OMG_OK;
}

// max |A*Ainv - I| over all entries (non-finite entries count as +inf): the unpivoted blocked elimination is only
// kept when this is at rounding level, so a small-but-not-vanishing pivot of an indefinite / nonsymmetric coarse
// operator cannot slip an inaccurate inverse into every cycle (the reference solves with pivoted SuperLU,
// openmg/solvers.py:16-26).  One thread per (row i, column j); bit pattern of a non-negative double orders like
// an unsigned integer.
__global__ void k_inverse_defect(const int *__restrict__ ptr, const int *__restrict__ col,
                                 const double *__restrict__ val, int n, const double *__restrict__ Ainv,
                                 unsigned long long *__restrict__ worst) {
    int j = blockIdx.x * OMG_TPB + threadIdx.x, i = blockIdx.y;
    double r = 0.0;
    if (j < n) {
        double acc = (i == j) ? -1.0 : 0.0;
        for (int p = ptr[i]; p < ptr[i + 1]; ++p) acc += val[p] * Ainv[(size_t)col[p] * n + j];
        r = isfinite(acc) ? fabs(acc) : __longlong_as_double(0x7ff0000000000000ll);
    }
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    if ((threadIdx.x & 31) == 0 && r > 0.0) atomicMax(worst, (unsigned long long)__double_as_longlong(r));
}

static int inverse_defect(omg_hierarchy *h, Level &L, int n, double *defect) {
    unsigned long long *worst = nullptr, w = 0;
    OMG_TRY(h_alloc_t(h, &worst, 1, true));
    k_inverse_defect<<<dim3(cdiv(n, OMG_TPB), n), OMG_TPB, 0, g.stream>>>(L.ptr, L.col, L.val, n, h->Ainv, worst);
    CUDA_TRY(cudaMemcpyAsync(&w, worst, sizeof(w), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    CUDA_TRY(cudaGetLastError());
    h_free(h, worst);
    long long bits = (long long)w;
    memcpy(defect, &bits, sizeof(double));
    return OMG_OK;
}

static int coarse_factor(omg_hierarchy *h) {
    Level &L = h->lv[h->nlev - 1];
    int n = L.n;
    if (n > 8192)
        return omg_set_error(OMG_EUNSUPPORTED,
                             "coarsest level has %d rows; the dense device factor supports <= 8192 "
                             "(use more gridLevels / a smaller minSize)", n);
    h->ncoarse = n;
    OMG_TRY(h_alloc_t(h, &h->Ainv, (size_t)n * n, true));
    if (!getenv("OMG_PIVOTED_FACTOR")) {
        bool ok = false;
        OMG_TRY(coarse_factor_blocked(h, L, n, &ok));
        if (ok) {   // no pivoting: keep the result only if A * Ainv = I to rounding (also catches inf/nan)
            double defect = 0.0;
            OMG_TRY(inverse_defect(h, L, n, &defect));
            h->coarse_defect = defect;
            if (defect <= 1e-9) return OMG_OK;
        }
        CUDA_TRY(cudaMemsetAsync(h->Ainv, 0, sizeof(double) * (size_t)n * n, g.stream));
    }
    double *colk = nullptr;
    int *piv = nullptr, *sing = nullptr;
    OMG_TRY(h_alloc_t(h, &colk, (size_t)n));
    OMG_TRY(h_alloc_t(h, &piv, (size_t)n));
    OMG_TRY(h_alloc_t(h, &sing, 1, true));
    int gb = cdiv(n, OMG_TPB);
    k_dense_from_csr<<<gb, OMG_TPB, 0, g.stream>>>(L.ptr, L.col, L.val, n, h->Ainv);
    dim3 ge(gb, cdiv(n, 16));
    for (int k = 0; k < n; ++k) {
        k_gj_pivot<<<1, 1024, 0, g.stream>>>(h->Ainv, n, k, piv, sing);
        k_gj_swap<<<gb, OMG_TPB, 0, g.stream>>>(h->Ainv, n, k, piv);
        k_gj_col<<<gb, OMG_TPB, 0, g.stream>>>(h->Ainv, n, k, colk);
        k_gj_scale<<<gb, OMG_TPB, 0, g.stream>>>(h->Ainv, n, k, colk);
        k_gj_elim<<<ge, OMG_TPB, 0, g.stream>>>(h->Ainv, n, k, colk);
    }
    k_gj_unperm<<<gb, OMG_TPB, 0, g.stream>>>(h->Ainv, n, piv);
    int s = 0;
    CUDA_TRY(cudaMemcpyAsync(&s, sing, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    CUDA_TRY(cudaGetLastError());
    h_free(h, colk);
    h_free(h, piv);
    h_free(h, sing);
    if (s) return omg_set_error(OMG_ESINGULAR, "coarsest-level operator is singular");
    OMG_TRY(inverse_defect(h, L, n, &h->coarse_defect));
    return OMG_OK;
}

// ------------------------------------------------------------------ vectors

// first index t in sorted a[0..n) with a[t] >= v, for two values (device array)
__global__ void k_lower_bounds(const int *__restrict__ a, int n, int v0, int v1, int *__restrict__ out) {
    if (threadIdx.x > 1 || blockIdx.x) return;
    int v = threadIdx.x ? v1 : v0;
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < v)
            lo = mid + 1;
        else
            hi = mid;
    }
    out[threadIdx.x] = lo;
}

static int lower_bounds(const int *a, int n, int v0, int v1, int *o0, int *o1) {
    int *d = nullptr, hres[2] = {0, 0};
    if (n > 0 && a) {
        CUDA_TRY(cudaMalloc(&d, 2 * sizeof(int)));
        k_lower_bounds<<<1, 2, 0, g.stream>>>(a, n, v0, v1, d);
        CUDA_TRY(cudaMemcpyAsync(hres, d, 2 * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
        CUDA_TRY(cudaStreamSynchronize(g.stream));
        cudaFree(d);
    }
    *o0 = hres[0];
    *o1 = hres[1];
    return OMG_OK;
}

static int alloc_vectors(omg_hierarchy *h) {
    int nlev = h->nlev;
    // ---- pads / halo widths from the band reach
    std::vector<int> reach(nlev, 0), reach2(nlev, 0);
    for (int l = 0; l < nlev; ++l) {
        Level &L = h->lv[l];
        if (L.kind != OMG_KIND_CSR)
            for (int k = 0; k < L.band.nb; ++k) {
                int a = std::abs(L.band.off[k]);
                if (a > reach[l]) {
                    reach2[l] = reach[l];
                    reach[l] = a;
                } else if (a > reach2[l] && a < reach[l])
                    reach2[l] = a;
            }
        L.pad = (2 * reach[l] + 15) / 16 * 16;    // the 2.5-D stencil path stages plane -1 with its halo rows
        if (L.kind != OMG_KIND_CSR && L.band.nb == 2 && L.n >= 512) L.pad = std::max(L.pad, 2 * 2048 + 16);   // 1-D rows view (two rows: fused two-colour sweep)
    }
    // ---- slab partition (multi-GPU): row0 / nloc per level
    std::vector<int64_t> lead(nlev), rows(nlev), row0(nlev), nloc(nlev);
    std::vector<int32_t> regular(nlev);
    for (int l = 0; l < nlev; ++l) {
        Level &L = h->lv[l];
        lead[l] = L.shape[0];
        rows[l] = L.n;
        // a slab level needs the closed-form restriction and a band operator (halo = band reach) whose exception
        // rows, if any, stay inside that reach
        regular[l] = (L.hasR && L.regular && L.kind != OMG_KIND_CSR) ? 1 : 0;
        if (L.kind == OMG_KIND_BAND_EXC && L.exc_reach > ((reach[l] + reach2[l] + 1) & ~1)) regular[l] = 0;
    }
    const char *env = getenv("OMG_AGGLOMERATE_BELOW");
    int64_t agg_below = env ? atoll(env) : (1ll << 19);
    int32_t ld = 0;
    for (;;) {
        OMG_TRY(omg_partition(nlev, lead.data(), rows.data(), regular.data(), g.nranks, g.rank, agg_below, &ld,
                              row0.data(), nloc.data()));
        bool ok = true;
        for (int l = 0; l < ld; ++l)
            if (nloc[l] < 2 * (int64_t)h->lv[l].pad) {    // slab thinner than its halos: replicate from here on
                regular[l] = 0;
                ok = false;
                break;
            }
        if (ok) break;
    }
    h->first_replicated = ld;
    for (int l = 0; l < nlev; ++l) {
        Level &L = h->lv[l];
        L.slab = l < ld;
        L.row0 = (int)row0[l];
        L.nloc = (int)nloc[l];
        L.halo = L.slab ? std::min(L.pad, (reach[l] + reach2[l] + 1) & ~1) : 0;
        if (L.hasR) {
            Level &C = h->lv[l + 1];
            int k = L.Rk > 0 ? L.Rk : 1;
            L.piece_row0 = L.slab ? (int)(row0[l] / k) : 0;
            L.piece_n = L.slab ? (int)(nloc[l] / k) : C.n;
        }
        L.exc_s0 = 0;
        L.exc_s1 = (int)L.nexc;
        L.crow_t0 = 0;
        L.crow_t1 = L.nexc_crows;
        if (L.slab && L.kind == OMG_KIND_BAND_EXC) {
            OMG_TRY(lower_bounds(L.exc_rows, (int)L.nexc, L.row0, L.row0 + L.nloc, &L.exc_s0, &L.exc_s1));
            OMG_TRY(lower_bounds(L.exc_crows, L.nexc_crows, L.piece_row0, L.piece_row0 + L.piece_n, &L.crow_t0,
                                 &L.crow_t1));
        }
    }
    for (int l = 0; l < nlev; ++l) {
        Level &L = h->lv[l];
        size_t len = (size_t)L.pad * 2 + (size_t)L.nloc + 16;
        OMG_TRY(h_alloc_t(h, &L.xa_base, len, true));
        OMG_TRY(h_alloc_t(h, &L.xb_base, len, true));
        OMG_TRY(h_alloc_t(h, &L.b_base, len, true));
        L.xa = L.xa_base + L.pad;
        L.xb = L.xb_base + L.pad;
        L.b = L.b_base + L.pad;
        if (L.hasR && !L.regular) {   // unfused residual -> explicit restriction needs r in HBM
            OMG_TRY(h_alloc_t(h, &L.r_base, len, true));
            L.r = L.r_base + L.pad;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(g.stream));      // buffers are zeroed before any neighbour may write into them
    OMG_TRY(dist_peer_setup(h));
    h->npartial = std::max(g.sm_count, 1) * 8;
    OMG_TRY(h_alloc_t(h, &h->partial, (size_t)h->npartial, true));
    OMG_TRY(h_alloc_t(h, &h->norm2_dev, 4, true));
    CUDA_TRY(cudaMallocHost((void **)&h->norm2_host, 4 * sizeof(double)));
    return OMG_OK;
}

// ------------------------------------------------------------------ driver

static double now_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// Level 0 operator is already in h->lv[0] (band and/or CSR).  Builds everything else.
// ---- sliced ELLPACK (SellOp) from the sorted CSR of a level

__global__ void k_sell_width(const int *__restrict__ ptr, int n, int *__restrict__ rlen, int *__restrict__ slots) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    int len = (i < n) ? ptr[i + 1] - ptr[i] : 0;
    if (i < n) rlen[i] = len;
    int w = len;
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if ((threadIdx.x & 31) == 0 && i < n) slots[i >> 5] = ((w + 1) & ~1) * 32;
}

__global__ void k_sell_fill(const int *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ val,
                            int n, const int *__restrict__ off, int *__restrict__ scol, double *__restrict__ sval) {
    int i = blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= n) return;
    int p0 = ptr[i], len = ptr[i + 1] - p0;
    int base = off[i >> 5] + (i & 31) * 2;
    for (int k = 0; k < len; ++k) {
        int t = base + (k >> 1) * 64 + (k & 1);
        scol[t] = col[p0 + k];
        sval[t] = val[p0 + k];
    }
}

static int build_sell(omg_hierarchy *h, Level &L) {
    if (L.kind != OMG_KIND_CSR || !L.ptr || L.n < 64 || L.row0 != 0 || L.nloc != L.n || getenv("OMG_NO_SELL"))
        return OMG_OK;
    int nsl = cdiv(L.n, 32);
    int *slots = nullptr;
    OMG_TRY(h_alloc_t(h, &slots, (size_t)nsl + 1, true));
    OMG_TRY(h_alloc_t(h, &L.sell_rlen, (size_t)L.n));
    OMG_TRY(h_alloc_t(h, &L.sell_off, (size_t)nsl + 1, true));
    k_sell_width<<<cdiv(L.n, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.ptr, L.n, L.sell_rlen, slots);
    int total = 0;
    OMG_TRY(exclusive_scan_i32(slots, L.sell_off, nsl, &total, g.stream));
    h_free(h, slots);
    if (total <= 0 || (double)total > 2.0 * (double)L.nnz + 64.0 * nsl) {
        // very uneven row lengths: the padded slices would cost more than they save; keep the scalar CSR walk
        h_free(h, L.sell_rlen);
        h_free(h, L.sell_off);
        L.sell_rlen = L.sell_off = nullptr;
        return OMG_OK;
    }
    OMG_TRY(h_alloc_t(h, &L.sell_col, (size_t)total, true));
    OMG_TRY(h_alloc_t(h, &L.sell_val, (size_t)total, true));
    k_sell_fill<<<cdiv(L.n, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.ptr, L.col, L.val, L.n, L.sell_off, L.sell_col, L.sell_val);
    L.sell_entries = total;
    CUDA_TRY(cudaGetLastError());
    return OMG_OK;
}

int build_hierarchy(omg_hierarchy *h) {
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&e2);
    cudaEventRecord(e0, g.stream);
    int rc = OMG_OK;
    for (int l = 0; l + 1 < h->nlev && rc == OMG_OK; ++l) {
        rc = galerkin_level(h, l);
        if (rc != OMG_OK) break;
        Level &C = h->lv[l + 1];
        rc = h_alloc_t(h, &C.diag, (size_t)C.n);
        if (rc != OMG_OK) break;
        k_extract_diag<<<cdiv(C.n, OMG_TPB), OMG_TPB, 0, g.stream>>>(C.ptr, C.col, C.val, C.n, C.row0, C.diag);
        rc = detect_band(h, C);
    }
    for (int l = 0; l < h->nlev && rc == OMG_OK; ++l) rc = build_sell(h, h->lv[l]);
    cudaEventRecord(e1, g.stream);
    if (rc == OMG_OK && (h->nlev > 1 || (h->flags & OMG_FLAG_FACTOR))) rc = coarse_factor(h);
    cudaEventRecord(e2, g.stream);
    if (rc == OMG_OK) rc = alloc_vectors(h);
    if (rc == OMG_OK) {
        cudaError_t e = cudaStreamSynchronize(g.stream);
        if (e != cudaSuccess) rc = omg_set_error(OMG_ECUDA, "setup failed: %s", cudaGetErrorString(e));
    }
    if (rc == OMG_OK) {
        h->t_galerkin_ms = now_ms(e0, e1);
        h->t_coarse_ms = now_ms(e1, e2);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    return rc;
}

// used by omg_api.cu for the level-0 CSR
int level0_from_csr(omg_hierarchy *h) {
    Level &L = h->lv[0];
    OMG_TRY(h_alloc_t(h, &L.diag, (size_t)L.n));
    k_extract_diag<<<cdiv(L.n, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.ptr, L.col, L.val, L.n, L.row0, L.diag);
    return detect_band(h, L);
}
