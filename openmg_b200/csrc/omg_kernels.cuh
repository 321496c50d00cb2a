// omg_kernels.cuh — generic (any dimension, any band set, CSR) sm_100a kernels of the
// V-cycle.  The structured 3-D/2-D/1-D fast paths live in omg_stencil.cuh; these are
// the always-available versions and the ones the fast paths are parity-checked against.
//
// All kernels are HBM-bound fp64 streaming kernels (no tensor cores): one thread per
// row (or per coarse row for the fused residual+restriction), coalesced along the flat
// index, read-only data through ld.global.nc, zero-padded vectors instead of bounds
// branches.
#pragma once
#include "omg_common.cuh"

#define OMG_TPB 256

// ---------------------------------------------------------------- operator policies

template <int HAS_EXC>
struct BandA {
    BandOp b;
    ExcOp e;
    // (A x)_i and a_ii.  x points at owned row 0 and is zero-padded by >= max|off|.
    __device__ __forceinline__ double ax(int i, const double *x, double &d) const {
        if (HAS_EXC) {
            unsigned w = __ldg(e.mask + (i >> 5));
            if ((w >> (i & 31)) & 1u) {
                int s = __ldg(e.wpre + (i >> 5)) + __popc(w & ((1u << (i & 31)) - 1u));
                int p0 = __ldg(e.ptr + s), p1 = __ldg(e.ptr + s + 1);
                double acc = 0.0;
                for (int p = p0; p < p1; ++p) acc += __ldg(e.val + p) * x[__ldg(e.col + p)];
                d = __ldg(e.diag + s);
                return acc;
            }
        }
        double acc = b.diag * x[i];
#pragma unroll
        for (int k = 0; k < OMG_MAXBAND; ++k)
            if (k < b.nb) acc += b.coef[k] * x[i + b.off[k]];
        d = b.diag;
        return acc;
    }
    __device__ __forceinline__ double dg(int i) const {
        if (HAS_EXC) {
            unsigned w = __ldg(e.mask + (i >> 5));
            if ((w >> (i & 31)) & 1u)
                return __ldg(e.diag + __ldg(e.wpre + (i >> 5)) + __popc(w & ((1u << (i & 31)) - 1u)));
        }
        return b.diag;
    }
};

struct CsrA {
    CsrOp c;
    __device__ __forceinline__ double ax(int i, const double *x, double &d) const {
        int p0 = __ldg(c.ptr + i), p1 = __ldg(c.ptr + i + 1);
        double acc = 0.0;
        for (int p = p0; p < p1; ++p) acc += __ldg(c.val + p) * x[__ldg(c.col + p)];
        d = __ldg(c.diag + i);
        return acc;
    }
    __device__ __forceinline__ double dg(int i) const { return __ldg(c.diag + i); }
};

// the vectorised path of CSR levels / general input matrices (openmg/operators.py:172-186): see SellOp
struct SellA {
    SellOp c;
    __device__ __forceinline__ double ax(int i, const double *x, double &d) const {
        const int len = __ldg(c.rlen + i);
        const int base = (__ldg(c.off + (i >> 5)) >> 1) + (i & 31);      // in pairs
        const double2 *v2 = reinterpret_cast<const double2 *>(c.val) + base;
        const int2 *c2 = reinterpret_cast<const int2 *>(c.col) + base;
        double acc = 0.0;
        for (int k = 0; k < len; k += 2) {
            const double2 v = __ldg(v2 + (k >> 1) * 32);
            const int2 j = __ldg(c2 + (k >> 1) * 32);
            acc += v.x * x[j.x];
            if (k + 1 < len) acc += v.y * x[j.y];
        }
        d = __ldg(c.diag + i);
        return acc;
    }
    __device__ __forceinline__ double dg(int i) const { return __ldg(c.diag + i); }
};

__device__ __forceinline__ int colour_of(const ColourRule &c, int ig) {
    if (c.flat) return ig & 1;
    int s = ig % c.s2;
    int t = ig / c.s2;
    if (c.alpha == 3) {
        s += t % c.s1;
        t /= c.s1;
    }
    return (s + t) & 1;
}

// coarse row I (global) -> first fine column (global), openmg/operators.py:63-68
__device__ __forceinline__ int reg_cc(const RegR &R, int I) {
    int I2 = I % R.cs2;
    int t = I / R.cs2;
    int I1 = t % R.cs1;
    int I0 = t / R.cs1;
    return ((2 * I0) * R.fs1 + 2 * I1) * R.fs2 + 2 * I2;
}
// fine column j (global) -> the coarse row whose aggregate contains it
__device__ __forceinline__ int reg_agg(const RegR &R, int j) {
    int j2 = j % R.fs2;
    int t = j / R.fs2;
    int j1 = t % R.fs1;
    int j0 = t / R.fs1;
    return ((j0 >> 1) * R.cs1 + (j1 >> 1)) * R.cs2 + (j2 >> 1);
}

// ---------------------------------------------------------------- smoothers

// weighted Jacobi sweep: xo = xi + omega (b - A xi)/diag      rows [lo,hi)
template <class AOp>
__global__ void __launch_bounds__(OMG_TPB) k_jacobi(AOp A, int lo, int hi, const double *__restrict__ xi,
                                                    const double *__restrict__ b, double *__restrict__ xo,
                                                    double omega) {
    int i = lo + blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= hi) return;
    double d;
    double ax = A.ax(i, xi, d);
    xo[i] = xi[i] + omega * (__ldg(b + i) - ax) / d;
}

// first Jacobi sweep from a zero iterate (coarse levels start from zeros,
// openmg/__init__.py:191-192): x = omega b / diag
template <class AOp>
__global__ void __launch_bounds__(OMG_TPB) k_jacobi_zero(AOp A, int lo, int hi, const double *__restrict__ b,
                                                         double *__restrict__ xo, double omega) {
    int i = lo + blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= hi) return;
    xo[i] = omega * __ldg(b + i) / A.dg(i);
}

// one colour half-sweep of the two-colour Gauss-Seidel, out of place:
//   xo_i = xi_i + [colour(i)==c] (b_i - (A xi)_i)/a_ii
// (same-colour couplings use xi, i.e. lagged values — oracle.rbgs)
template <class AOp>
__global__ void __launch_bounds__(OMG_TPB) k_colour_relax(AOp A, ColourRule cr, int colour, int row0, int lo, int hi,
                                                          const double *__restrict__ xi,
                                                          const double *__restrict__ b, double *__restrict__ xo) {
    int i = lo + blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= hi) return;
    double v = xi[i];
    if (colour_of(cr, i + row0) == colour) {
        double d;
        double ax = A.ax(i, xi, d);
        v = v + (__ldg(b + i) - ax) / d;
    }
    xo[i] = v;
}

// the reference's lexicographic Gauss-Seidel (openmg/solvers.py:56-68), one thread,
// strictly sequential: exact semantics for small parity cases, not a performance path.
template <class AOp>
__global__ void k_lexgs(AOp A, int n, double *x, const double *b, int sweeps) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int s = 0; s < sweeps; ++s)
        for (int i = 0; i < n; ++i) {
            double d;
            double ax = A.ax(i, (const double *)x, d);
            x[i] = x[i] + (b[i] - ax) / d;
            __threadfence_block();
        }
}

// ---------------------------------------------------------------- residual (+ restriction)

template <class AOp>
__global__ void __launch_bounds__(OMG_TPB) k_residual(AOp A, int lo, int hi, const double *__restrict__ x,
                                                      const double *__restrict__ b, double *__restrict__ r) {
    int i = lo + blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= hi) return;
    double d;
    double ax = A.ax(i, x, d);
    r[i] = __ldg(b + i) - ax;
}

template <class AOp>
__global__ void __launch_bounds__(OMG_TPB) k_matvec(AOp A, int lo, int hi, const double *__restrict__ x,
                                                    double *__restrict__ y) {
    int i = lo + blockIdx.x * OMG_TPB + threadIdx.x;
    if (i >= hi) return;
    double d;
    y[i] = A.ax(i, x, d);
}

// fused  rc = R (b - A x)  for the closed-form restriction: one thread per coarse row,
// the fine residual is never written to HBM (openmg/__init__.py:209-210).
// crow0 / frow0: global index of local coarse / fine row 0.
template <class AOp>
__global__ void __launch_bounds__(OMG_TPB) k_residual_restrict(AOp A, RegR R, int crow0, int frow0, int clo, int chi,
                                                               const double *__restrict__ x,
                                                               const double *__restrict__ b,
                                                               double *__restrict__ rc) {
    int I = clo + blockIdx.x * OMG_TPB + threadIdx.x;
    if (I >= chi) return;
    int cc = reg_cc(R, I + crow0) - frow0;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < R.k) {
            int i = cc + R.o[k];
            double d;
            double ax = A.ax(i, x, d);
            acc += __ldg(b + i) - ax;      // ascending-column order, like scipy's csr_matvec of R
        }
    }
    rc[I] = R.w * acc;
}

// rc = R r for an explicit R (non-regular shapes): row I has columns Rcc[I] + o[k], weight w
static __global__ void __launch_bounds__(OMG_TPB) k_restrict_explicit(const int *__restrict__ Rcc, RegR R, int clo, int chi,
                                                               const double *__restrict__ r,
                                                               double *__restrict__ rc) {
    int I = clo + blockIdx.x * OMG_TPB + threadIdx.x;
    if (I >= chi) return;
    int cc = Rcc[I];
    double acc = 0.0;
    for (int k = 0; k < R.k; ++k) acc += R.w * r[cc + R.o[k]];
    rc[I] = acc;
}

// ---------------------------------------------------------------- prolongation + correction

// x += R^T e (openmg/__init__.py:214,224), closed-form R: thread per fine row
static __global__ void __launch_bounds__(OMG_TPB) k_prolong_correct(RegR R, int crow0, int frow0, int lo, int hi,
                                                             const double *__restrict__ e,
                                                             const double *__restrict__ xi, double *__restrict__ xo) {
    int j = lo + blockIdx.x * OMG_TPB + threadIdx.x;
    if (j >= hi) return;
    int I = reg_agg(R, j + frow0) - crow0;
    xo[j] = xi[j] + R.w * __ldg(e + I);
}

// x += R^T e for explicit R^T (pattern-only CSR over fine rows)
static __global__ void __launch_bounds__(OMG_TPB) k_prolong_correct_csr(const int *__restrict__ RTptr,
                                                                 const int *__restrict__ RTcol, double w, int lo,
                                                                 int hi, const double *__restrict__ e,
                                                                 const double *__restrict__ xi,
                                                                 double *__restrict__ xo) {
    int j = lo + blockIdx.x * OMG_TPB + threadIdx.x;
    if (j >= hi) return;
    double acc = 0.0;
    for (int p = RTptr[j]; p < RTptr[j + 1]; ++p) acc += w * e[RTcol[p]];
    xo[j] = xi[j] + acc;
}

// ---------------------------------------------------------------- norms (deterministic)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (wid == 0) v = warp_sum(v);
    return v;   // valid in thread 0
}

// partial[blockIdx] = sum over this block's grid-stride rows of (b - A x)^2
template <class AOp>
__global__ void __launch_bounds__(OMG_TPB) k_resnorm_partial(AOp A, int lo, int hi, const double *__restrict__ x,
                                                             const double *__restrict__ b,
                                                             double *__restrict__ partial) {
    double acc = 0.0;
    for (int i = lo + blockIdx.x * OMG_TPB + threadIdx.x; i < hi; i += gridDim.x * OMG_TPB) {
        double d;
        double r = __ldg(b + i) - A.ax(i, x, d);
        acc += r * r;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// out[0] = sum(partial[0..m)) in fixed order (single block)
static __global__ void k_final_sum(const double *__restrict__ partial, int m, double *__restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) acc += partial[i];
    acc = block_sum(acc);
    if (threadIdx.x == 0) out[0] = acc;
}

// ---------------------------------------------------------------- coarse solve

// x = Ainv b, dense row-major n x n inverse computed at setup: one warp per row, 128-bit loads,
// four independent accumulators (2 KB in flight per warp).  One CTA when n <= 512, else all SMs.
static __global__ void __launch_bounds__(OMG_TPB) k_coarse_gemv(const double *__restrict__ Ainv, int n,
                                                         const double *__restrict__ b, double *__restrict__ x) {
    int warp = (blockIdx.x * OMG_TPB + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * OMG_TPB) >> 5;
    for (int r = warp; r < n; r += nwarps) {
        const double *row = Ainv + (size_t)r * n;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int c = 0;
        if ((n & 1) == 0) {
            const double2 *row2 = reinterpret_cast<const double2 *>(row);
            const double2 *b2 = reinterpret_cast<const double2 *>(b);
            int n2 = n >> 1;
            int c2 = lane;
            for (; c2 + 96 < n2; c2 += 128) {
                double2 m0 = __ldg(row2 + c2), m1 = __ldg(row2 + c2 + 32), m2 = __ldg(row2 + c2 + 64),
                        m3 = __ldg(row2 + c2 + 96);
                double2 v0 = __ldg(b2 + c2), v1 = __ldg(b2 + c2 + 32), v2 = __ldg(b2 + c2 + 64),
                        v3 = __ldg(b2 + c2 + 96);
                a0 += m0.x * v0.x;
                a0 += m0.y * v0.y;
                a1 += m1.x * v1.x;
                a1 += m1.y * v1.y;
                a2 += m2.x * v2.x;
                a2 += m2.y * v2.y;
                a3 += m3.x * v3.x;
                a3 += m3.y * v3.y;
            }
            for (; c2 < n2; c2 += 32) {
                double2 m0 = __ldg(row2 + c2), v0 = __ldg(b2 + c2);
                a0 += m0.x * v0.x;
                a0 += m0.y * v0.y;
            }
            c = n;
        }
        for (c += lane; c < n; c += 32) a0 += __ldg(row + c) * __ldg(b + c);
        double acc = warp_sum((a0 + a1) + (a2 + a3));
        if (lane == 0) x[r] = acc;
    }
}
