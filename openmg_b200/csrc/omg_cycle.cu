// omg_cycle.cu — one V(pre,post) cycle (openmg/__init__.py:151-236) as a fixed sequence of
// kernel launches on the library stream; omg_api.cu captures it into a CUDA graph.
//
// Per level l < coarsest:
//   pre-smooth (first sweep from zero specialised on l >= 1)      :201
//   b_{l+1} = R_l (b_l - A_l x_l)            one fused kernel     :209-210
//   recurse with zero initial iterate                              :213, :191-192
//   x_l += R_l^T x_{l+1}, post-smooth        fused when possible   :214-224
// coarsest: x = A_L^{-1} b (dense inverse applied as a GEMV)      :234
// The per-level residual norm of :227 is only consumed at level 0 (:113,136); it is
// computed there, on request.
#include "omg_hier.cuh"
#include "omg_kernels.cuh"
#include "omg_stencil.cuh"

#define GRID(n) cdiv((n), OMG_TPB), OMG_TPB, 0, g.stream

#define DISPATCH_A(L, ...)                                   \
    do {                                                     \
        if ((L).kind == OMG_KIND_BAND) {                     \
            BandA<0> A{(L).band, ExcOp{}};                   \
            __VA_ARGS__;                                     \
        } else if ((L).kind == OMG_KIND_BAND_EXC) {          \
            BandA<1> A{(L).band, (L).exc_op()};              \
            __VA_ARGS__;                                     \
        } else {                                             \
            CsrA A{(L).csr_op()};                            \
            __VA_ARGS__;                                     \
        }                                                    \
    } while (0)

static inline double *other(Level &L, double *cur) { return cur == L.xa ? L.xb : L.xa; }
static inline int lvl(omg_hierarchy *h, const Level &L) { return (int)(&L - h->lv.data()); }

int launch_matvec(omg_hierarchy *h, Level &L, const double *x, double *y) {
    ProfScope ps(h, "matvec", lvl(h, L), 16.0 * L.nloc);
    DISPATCH_A(L, (k_matvec<decltype(A)><<<GRID(L.nloc)>>>(A, 0, L.nloc, x, y)));
    h->launches++;
    return OMG_OK;
}

int launch_residual(omg_hierarchy *h, Level &L, const double *x, const double *b, double *r) {
    ProfScope ps(h, "residual", lvl(h, L), 24.0 * L.nloc);
    DISPATCH_A(L, (k_residual<decltype(A)><<<GRID(L.nloc)>>>(A, 0, L.nloc, x, b, r)));
    h->launches++;
    return OMG_OK;
}

// sum of squares of b - A x into h->norm2_dev[slot]
int launch_resnorm2(omg_hierarchy *h, Level &L, const double *x, const double *b, int slot) {
    int blocks = std::min(h->npartial, cdiv(L.nloc, OMG_TPB));
    blocks = std::max(blocks, 1);
    ProfScope ps(h, "residual_norm", lvl(h, L), 16.0 * L.nloc);
    DISPATCH_A(L, (k_resnorm_partial<decltype(A)><<<blocks, OMG_TPB, 0, g.stream>>>(A, 0, L.nloc, x, b, h->partial)));
    k_final_sum<<<1, 1024, 0, g.stream>>>(h->partial, blocks, h->norm2_dev + slot);
    h->launches += 2;
    return OMG_OK;
}

// `sweeps` smoothing iterations on level L starting from `cur` (nullptr = zero iterate).
// Returns the buffer holding the result.
double *launch_smooth(omg_hierarchy *h, Level &L, int smoother, double omega, int sweeps, double *cur,
                      const double *b) {
    int n = L.nloc;
    if (cur == nullptr && (sweeps == 0 || smoother != OMG_SMOOTH_JACOBI)) {
        cudaMemsetAsync(L.xa, 0, sizeof(double) * (size_t)n, g.stream);
        cur = L.xa;
    }
    if (smoother == OMG_SMOOTH_JACOBI) {
        for (int s = 0; s < sweeps; ++s) {
            if (cur == nullptr) {
                ProfScope ps(h, "jacobi_zero", lvl(h, L), 16.0 * n);
                DISPATCH_A(L, (k_jacobi_zero<decltype(A)><<<GRID(n)>>>(A, 0, n, b, L.xa, omega)));
                cur = L.xa;
            } else {
                ProfScope ps(h, "jacobi", lvl(h, L), 24.0 * n);
                double *out = other(L, cur);
                if (!(h->flags & OMG_FLAG_NO_FUSED) && stencil_jacobi(h, L, cur, b, out, omega)) {
                } else {
                    DISPATCH_A(L, (k_jacobi<decltype(A)><<<GRID(n)>>>(A, 0, n, cur, b, out, omega)));
                }
                cur = out;
            }
            h->launches++;
        }
    } else if (smoother == OMG_SMOOTH_RBGS) {
        for (int s = 0; s < sweeps; ++s)
            for (int c = 0; c < 2; ++c) {
                ProfScope ps(h, "rbgs_half", lvl(h, L), 12.0 * n);
                double *out = other(L, cur);
                DISPATCH_A(L, (k_colour_relax<decltype(A)><<<GRID(n)>>>(A, L.colour, c, L.row0, 0, n, cur, b, out)));
                cur = out;
                h->launches++;
            }
    } else {   // lexicographic GS, in place
        if (sweeps > 0) {
            ProfScope ps(h, "lexgs", lvl(h, L), 24.0 * n * sweeps);
            DISPATCH_A(L, (k_lexgs<decltype(A)><<<1, 32, 0, g.stream>>>(A, n, cur, b, sweeps)));
            h->launches++;
        }
    }
    return cur;
}

// b_{l+1} = R_l (b_l - A_l x)
int launch_residual_restrict(omg_hierarchy *h, int l, const double *x, const double *b, double *rc) {
    Level &L = h->lv[l];
    Level &C = h->lv[l + 1];
    if (L.regular) {
        ProfScope ps(h, "residual_restrict", l, 16.0 * L.nloc + 8.0 * C.nloc);
        if (!(h->flags & OMG_FLAG_NO_FUSED) && stencil_residual_restrict(h, L, C, x, b, rc)) {
        } else {
            DISPATCH_A(L, (k_residual_restrict<decltype(A)><<<GRID(C.nloc)>>>(A, L.reg, C.row0, L.row0, 0, C.nloc, x, b, rc)));
        }
        h->launches++;
    } else {
        if (!L.r) {
            if (h_alloc_t(h, &L.r_base, (size_t)L.nloc + 16, true) != OMG_OK) return OMG_ENOMEM;
            L.r = L.r_base;
        }
        launch_residual(h, L, x, b, L.r);
        ProfScope ps(h, "restrict", l, 8.0 * L.nloc + 8.0 * C.nloc);
        k_restrict_explicit<<<GRID(C.nloc)>>>(L.Rcc, L.reg, 0, C.nloc, L.r, rc);
        h->launches++;
    }
    return OMG_OK;
}

// xo = xi + R_l^T e   (xo may alias xi)
int launch_prolong_correct(omg_hierarchy *h, int l, const double *e, const double *xi, double *xo) {
    Level &L = h->lv[l];
    Level &C = h->lv[l + 1];
    ProfScope ps(h, "prolong_correct", l, 16.0 * L.nloc + 8.0 * C.nloc);
    if (L.regular)
        k_prolong_correct<<<GRID(L.nloc)>>>(L.reg, C.row0, L.row0, 0, L.nloc, e, xi, xo);
    else
        k_prolong_correct_csr<<<GRID(L.nloc)>>>(L.RTptr, L.RTcol, L.Rw, 0, L.nloc, e, xi, xo);
    h->launches++;
    return OMG_OK;
}

// correction + post-smoothing (openmg/__init__.py:214-224).  Returns the result buffer.
double *launch_prolong_correct_smooth(omg_hierarchy *h, int l, int smoother, double omega, int sweeps, double *cur,
                                      const double *e, const double *b) {
    Level &L = h->lv[l];
    if (sweeps > 0 && smoother == OMG_SMOOTH_JACOBI && L.regular && !(h->flags & OMG_FLAG_NO_FUSED)) {
        double *out = other(L, cur);
        bool ok;
        {
            ProfScope ps(h, "prolong_jacobi", l, 24.0 * L.nloc + 8.0 * h->lv[l + 1].nloc);
            ok = stencil_prolong_jacobi(h, L, h->lv[l + 1], cur, e, b, out, omega);
            if (!ok && ps.idx >= 0) {   // not applicable: drop the record
                cudaEventDestroy(h->prof.back().e0);
                cudaEventDestroy(h->prof.back().e1);
                h->prof.pop_back();
                ps.idx = -1;
            }
        }
        if (ok) {
            h->launches++;
            return launch_smooth(h, L, smoother, omega, sweeps - 1, out, b);
        }
    }
    // in place: each thread reads and writes only its own x_j
    launch_prolong_correct(h, l, e, cur, cur);
    return launch_smooth(h, L, smoother, omega, sweeps, cur, b);
}

int launch_coarse_solve(omg_hierarchy *h, const double *b, double *x) {
    int n = h->ncoarse;
    ProfScope ps(h, "coarse_solve", h->nlev - 1, 16.0 * n);
    int blocks = n <= 512 ? 1 : std::min(cdiv((int64_t)n * 32, OMG_TPB), std::max(g.sm_count, 1) * 8);
    k_coarse_gemv<<<blocks, OMG_TPB, 0, g.stream>>>(h->Ainv, n, b, x);
    h->launches++;
    return OMG_OK;
}

static double *cycle_level(omg_hierarchy *h, int l, const CycleCfg &cfg, double *cur) {
    Level &L = h->lv[l];
    if (l == h->nlev - 1) {
        launch_coarse_solve(h, L.b, L.xa);
        return L.xa;
    }
    Level &C = h->lv[l + 1];
    bool fused0 = false;
    if (cur == nullptr && cfg.pre == 1 && cfg.smoother == OMG_SMOOTH_JACOBI && L.regular &&
        !(h->flags & OMG_FLAG_NO_FUSED)) {
        // zero initial iterate (openmg/__init__.py:191-192): sweep + residual + restriction in one pass over b
        ProfScope ps(h, "jacobi0_residual_restrict", l, 16.0 * L.nloc + 8.0 * C.nloc);
        fused0 = stencil_jacobi0_residual_restrict(h, L, C, L.b, L.xa, C.b, cfg.omega);
        if (fused0) {
            cur = L.xa;
            h->launches++;
        } else if (ps.idx >= 0) {
            cudaEventDestroy(h->prof.back().e0);
            cudaEventDestroy(h->prof.back().e1);
            h->prof.pop_back();
            ps.idx = -1;
        }
    }
    if (!fused0) {
        cur = launch_smooth(h, L, cfg.smoother, cfg.omega, cfg.pre, cur, L.b);
        launch_residual_restrict(h, l, cur, L.b, C.b);
    }
    double *e = cycle_level(h, l + 1, cfg, nullptr);
    return launch_prolong_correct_smooth(h, l, cfg.smoother, cfg.omega, cfg.post, cur, e, L.b);
}

double *cycle_from_level(omg_hierarchy *h, int l, const CycleCfg &cfg, double *cur) {
    return cycle_level(h, l, cfg, cur);
}

// Issues one cycle on g.stream; level-0 iterate buffer tracked in h->cur0.
int run_cycle(omg_hierarchy *h, const CycleCfg &cfg) {
    Level &L0 = h->lv[0];
    double *cur = h->cur0 ? L0.xb : L0.xa;
    cur = cycle_level(h, 0, cfg, cur);
    h->cur0 = (cur == L0.xb) ? 1 : 0;
    if (cfg.with_norm) launch_resnorm2(h, L0, cur, L0.b, 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return omg_set_error(OMG_ECUDA, "cycle launch failed: %s", cudaGetErrorString(e));
    return OMG_OK;
}
