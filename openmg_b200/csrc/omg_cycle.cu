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
//
// Indexing convention: level vectors are stored [pad | owned rows | pad]; `L.xa`, `L.xb`,
// `L.b` point at the first OWNED row.  The generic kernels index by GLOBAL row, so they get
// "virtual" pointers V(p) = p - row0 and the row range [row0, row0 + nloc); on one GPU
// row0 == 0 and this is the identity.  Before an operator is applied to a vector of a slab
// level its halos are filled from the neighbouring ranks (dist_halo_exchange).
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
        } else if ((L).sell_val) {                           \
            SellA A{(L).sell_op()};                          \
            __VA_ARGS__;                                     \
        } else {                                             \
            CsrA A{(L).csr_op()};                            \
            __VA_ARGS__;                                     \
        }                                                    \
    } while (0)

static inline double *other(Level &L, double *cur) { return cur == L.xa ? L.xb : L.xa; }
static inline int lvl(omg_hierarchy *h, const Level &L) { return (int)(&L - h->lv.data()); }
template <class T>
static inline T *V(const Level &L, T *p) { return p - L.row0; }

int launch_matvec(omg_hierarchy *h, Level &L, double *x, double *y) {
    dist_halo_exchange(h, L, x);
    dist_halo_wait(h);
    ProfScope ps(h, "matvec", lvl(h, L), 16.0 * L.nloc + L.matrix_bytes());
    int lo = L.row0, hi = L.row0 + L.nloc;
    DISPATCH_A(L, (k_matvec<decltype(A)><<<GRID(L.nloc)>>>(A, lo, hi, V(L, x), V(L, y))));
    h->launches++;
    return OMG_OK;
}

int launch_residual(omg_hierarchy *h, Level &L, double *x, const double *b, double *r) {
    dist_halo_exchange(h, L, x);
    dist_halo_wait(h);
    ProfScope ps(h, "residual", lvl(h, L), 24.0 * L.nloc + L.matrix_bytes());
    int lo = L.row0, hi = L.row0 + L.nloc;
    DISPATCH_A(L, (k_residual<decltype(A)><<<GRID(L.nloc)>>>(A, lo, hi, V(L, x), V(L, b), V(L, r))));
    h->launches++;
    return OMG_OK;
}

// sum of squares of b - A x (all ranks) into h->norm2_dev[slot]
int launch_resnorm2(omg_hierarchy *h, Level &L, double *x, const double *b, int slot) {
    dist_halo_exchange(h, L, x);
    dist_halo_wait(h);
    int blocks = std::min(h->npartial, cdiv(L.nloc, OMG_TPB));
    blocks = std::max(blocks, 1);
    ProfScope ps(h, "residual_norm", lvl(h, L), 16.0 * L.nloc + L.matrix_bytes());
    int lo = L.row0, hi = L.row0 + L.nloc;
    DISPATCH_A(L, (k_resnorm_partial<decltype(A)><<<blocks, OMG_TPB, 0, g.stream>>>(A, lo, hi, V(L, x), V(L, b), h->partial)));
    k_final_sum<<<1, 1024, 0, g.stream>>>(h->partial, blocks, h->norm2_dev + slot);
    h->launches += 2;
    if (L.slab) dist_allreduce_sum(h, h->norm2_dev + slot, 1);
    return OMG_OK;
}

// `sweeps` smoothing iterations on level L starting from `cur` (nullptr = zero iterate).
// Returns the buffer holding the result.
double *launch_smooth(omg_hierarchy *h, Level &L, int smoother, double omega, int sweeps, double *cur,
                      const double *b) {
    int n = L.nloc, lo = L.row0, hi = L.row0 + L.nloc;
    if (cur == nullptr && sweeps > 0 && smoother == OMG_SMOOTH_RBGS && !(h->flags & OMG_FLAG_NO_FUSED) &&
        stencil_rb_sweep0(h, L, nullptr, nullptr)) {
        // zero initial iterate (openmg/__init__.py:191-192): the first two-colour sweep reads only b
        ProfScope ps(h, "rbgs_sweep0", lvl(h, L), 16.0 * n);
        if (stencil_rb_sweep0(h, L, b, L.xa)) {
            cur = L.xa;
            h->launches++;
            --sweeps;
        }
    }
    if (cur == nullptr && (sweeps == 0 || smoother != OMG_SMOOTH_JACOBI)) {
        cudaMemsetAsync(L.xa, 0, sizeof(double) * (size_t)n, g.stream);
        cur = L.xa;
    }
    if (smoother == OMG_SMOOTH_JACOBI) {
        for (int s = 0; s < sweeps; ++s) {
            if (cur == nullptr) {
                ProfScope ps(h, "jacobi_zero", lvl(h, L), 16.0 * n);
                DISPATCH_A(L, (k_jacobi_zero<decltype(A)><<<GRID(n)>>>(A, lo, hi, V(L, b), V(L, L.xa), omega)));
                cur = L.xa;
            } else {
                dist_halo_exchange(h, L, cur);
                ProfScope ps(h, "jacobi", lvl(h, L), 24.0 * n + L.matrix_bytes());
                double *out = other(L, cur);
                if (!(h->flags & OMG_FLAG_NO_FUSED) && stencil_jacobi(h, L, cur, b, out, omega)) {
                } else {
                    dist_halo_wait(h);
                    DISPATCH_A(L, (k_jacobi<decltype(A)><<<GRID(n)>>>(A, lo, hi, V(L, cur), V(L, b), V(L, out), omega)));
                }
                cur = out;
            }
            h->launches++;
        }
    } else if (smoother == OMG_SMOOTH_RBGS) {
        for (int s = 0; s < sweeps; ++s) {
            if (!(h->flags & OMG_FLAG_NO_FUSED) && stencil_rb_sweep(h, L, nullptr, nullptr, nullptr)) {
                // both colours in one pass (unsharded pure-band levels)
                ProfScope ps(h, "rbgs_sweep", lvl(h, L), 24.0 * n);
                double *out = other(L, cur);
                if (stencil_rb_sweep(h, L, cur, b, out)) {
                    cur = out;
                    h->launches++;
                    continue;
                }
            }
            for (int c = 0; c < 2; ++c) {
                dist_halo_exchange(h, L, cur);
                ProfScope ps(h, "rbgs_half", lvl(h, L), 12.0 * n + 0.5 * L.matrix_bytes());
                double *out = other(L, cur);
                if (!(h->flags & OMG_FLAG_NO_FUSED) && stencil_colour_relax(h, L, c, cur, b, out)) {
                } else {
                    dist_halo_wait(h);
                    DISPATCH_A(L, (k_colour_relax<decltype(A)><<<GRID(n)>>>(A, L.colour, c, 0, lo, hi, V(L, cur), V(L, b), V(L, out))));
                }
                cur = out;
                h->launches++;
            }
        }
    } else {   // lexicographic GS, in place (sequential: single GPU / replicated levels only)
        dist_halo_wait(h);
        if (sweeps > 0) {
            ProfScope ps(h, "lexgs", lvl(h, L), 24.0 * n * sweeps);
            DISPATCH_A(L, (k_lexgs<decltype(A)><<<1, 32, 0, g.stream>>>(A, n, cur, b, sweeps)));
            h->launches++;
        }
    }
    return cur;
}

// b_{l+1} = R_l (b_l - A_l x); at the slab -> replicated transition the pieces are all-gathered
int launch_residual_restrict(omg_hierarchy *h, int l, double *x, const double *b, double *rc) {
    Level &L = h->lv[l];
    Level &C = h->lv[l + 1];
    dist_halo_exchange(h, L, x);
    double *rcv = V(C, rc);                    // indexable by global coarse row
    if (L.regular) {
        ProfScope ps(h, "residual_restrict", l, 16.0 * L.nloc + 8.0 * L.piece_n + L.matrix_bytes());
        if (!(h->flags & OMG_FLAG_NO_FUSED) && stencil_residual_restrict(h, L, C, x, b, rcv)) {
        } else {
            dist_halo_wait(h);
            int clo = L.piece_row0, chi = L.piece_row0 + L.piece_n;
            DISPATCH_A(L, (k_residual_restrict<decltype(A)><<<GRID(L.piece_n)>>>(A, L.reg, 0, 0, clo, chi, V(L, x), V(L, b), rcv)));
        }
        h->launches++;
    } else {
        launch_residual(h, L, x, b, L.r);
        ProfScope ps(h, "restrict", l, 8.0 * L.nloc + 8.0 * C.nloc);
        k_restrict_explicit<<<GRID(C.nloc)>>>(L.Rcc, L.reg, 0, C.nloc, L.r, rc);
        h->launches++;
    }
    if (L.slab && !C.slab) return dist_allgather(h, rcv + L.piece_row0, rcv, (size_t)L.piece_n);
    return OMG_OK;
}

// pre-smoothing + restricted residual (openmg/__init__.py:201, :209-210): `sweeps` sweeps from *cur (nullptr = zero
// iterate), then b_{l+1} = R_l (b_l - A_l x).  A last Jacobi sweep from a non-zero iterate and the restricted residual
// run as ONE pass over x where the level allows it (k_jr3 / k_jr2: unsharded levels only, so no halo is involved).
// *cur: the buffer holding the smoothed iterate.
int launch_smooth_residual_restrict(omg_hierarchy *h, int l, int smoother, double omega, int sweeps, double **cur,
                                    const double *b, double *rc) {
    Level &L = h->lv[l];
    Level &C = h->lv[l + 1];
    if (smoother == OMG_SMOOTH_JACOBI && sweeps >= 1 && (*cur != nullptr || sweeps >= 2) && L.regular &&
        !(h->flags & OMG_FLAG_NO_FUSED) &&
        stencil_jacobi_residual_restrict(h, L, C, nullptr, nullptr, nullptr, nullptr, omega)) {     // applicability probe
        double *c = launch_smooth(h, L, smoother, omega, sweeps - 1, *cur, b);
        double *out = other(L, c);
        bool ok;
        {
            ProfScope ps(h, "jacobi_residual_restrict", l, 24.0 * L.nloc + 8.0 * L.piece_n);
            ok = stencil_jacobi_residual_restrict(h, L, C, c, b, out, V(C, rc), omega);
        }
        if (ok) {
            h->launches++;
            *cur = out;
            return OMG_OK;
        }
        *cur = launch_smooth(h, L, smoother, omega, 1, c, b);      // launch failed after a successful probe
        return launch_residual_restrict(h, l, *cur, b, rc);
    }
    *cur = launch_smooth(h, L, smoother, omega, sweeps, *cur, b);
    return launch_residual_restrict(h, l, *cur, b, rc);
}

// xo = xi + R_l^T e   (xo may alias xi)
int launch_prolong_correct(omg_hierarchy *h, int l, const double *e, const double *xi, double *xo) {
    Level &L = h->lv[l];
    Level &C = h->lv[l + 1];
    ProfScope ps(h, "prolong_correct", l, 16.0 * L.nloc + 8.0 * L.piece_n);
    if (L.regular)
        k_prolong_correct<<<GRID(L.nloc)>>>(L.reg, 0, 0, L.row0, L.row0 + L.nloc, V(C, e), V(L, xi), V(L, xo));
    else
        k_prolong_correct_csr<<<GRID(L.nloc)>>>(L.RTptr, L.RTcol, L.Rw, 0, L.nloc, e, xi, xo);
    h->launches++;
    return OMG_OK;
}

// correction + post-smoothing (openmg/__init__.py:214-224).  Returns the result buffer.
// cur_halo_valid: cur's halos were already filled (it was the input of the restriction).
double *launch_prolong_correct_smooth(omg_hierarchy *h, int l, int smoother, double omega, int sweeps, double *cur,
                                      double *e, const double *b, bool cur_halo_valid) {
    Level &L = h->lv[l];
    Level &C = h->lv[l + 1];
    if (sweeps > 0 && smoother == OMG_SMOOTH_JACOBI && L.regular && !(h->flags & OMG_FLAG_NO_FUSED) &&
        stencil_prolong_jacobi(h, L, C, nullptr, nullptr, nullptr, nullptr, omega)) {     // applicability probe
        // (a fused pull reads the neighbours' rows inside the kernel and leaves nothing in the local halo rows)
        if (!cur_halo_valid || h->peer.pull_ok) dist_halo_exchange(h, L, cur);
        dist_halo_exchange(h, C, e);
        double *out = other(L, cur);
        bool ok;
        {
            ProfScope ps(h, "prolong_jacobi", l, 24.0 * L.nloc + 8.0 * L.piece_n);
            ok = stencil_prolong_jacobi(h, L, C, cur, e, b, out, omega);
        }
        if (ok) {
            h->launches++;
            return launch_smooth(h, L, smoother, omega, sweeps - 1, out, b);
        }
        dist_halo_wait(h);          // launch failed after a successful probe: unfused path below
        launch_prolong_correct(h, l, e, cur, cur);
        return launch_smooth(h, L, smoother, omega, sweeps, cur, b);
    }
    if (sweeps > 0 && smoother == OMG_SMOOTH_RBGS && L.regular && !(h->flags & OMG_FLAG_NO_FUSED) &&
        stencil_prolong_rb_sweep(h, L, C, nullptr, nullptr, nullptr, nullptr)) {          // applicability probe
        // correction fused with the whole first post-smoothing sweep
        double *out = other(L, cur);
        bool ok;
        {
            ProfScope ps(h, "prolong_rbgs_sweep", l, 24.0 * L.nloc + 8.0 * L.piece_n);
            ok = stencil_prolong_rb_sweep(h, L, C, cur, e, b, out);
        }
        if (ok) {
            h->launches++;
            return launch_smooth(h, L, smoother, omega, sweeps - 1, out, b);
        }
        launch_prolong_correct(h, l, e, cur, cur);
        return launch_smooth(h, L, smoother, omega, sweeps, cur, b);
    }
    if (sweeps > 0 && smoother == OMG_SMOOTH_RBGS && L.regular && !(h->flags & OMG_FLAG_NO_FUSED) &&
        stencil_prolong_colour_relax(h, L, C, 0, nullptr, nullptr, nullptr, nullptr)) {   // applicability probe
        // correction fused with the colour-0 half of the first post-smoothing sweep
        if (!cur_halo_valid || h->peer.pull_ok) dist_halo_exchange(h, L, cur);
        dist_halo_exchange(h, C, e);
        double *out = other(L, cur);
        bool ok;
        {
            ProfScope ps(h, "prolong_rbgs_half", l, 12.0 * L.nloc + 8.0 * L.piece_n);
            ok = stencil_prolong_colour_relax(h, L, C, 0, cur, e, b, out);
        }
        if (!ok) {
            dist_halo_wait(h);
            launch_prolong_correct(h, l, e, cur, cur);
            return launch_smooth(h, L, smoother, omega, sweeps, cur, b);
        }
        h->launches++;
        cur = out;
        {   // colour-1 half of that sweep
            dist_halo_exchange(h, L, cur);
            ProfScope ps(h, "rbgs_half", l, 12.0 * L.nloc + 0.5 * L.matrix_bytes());
            out = other(L, cur);
            int lo = L.row0, hi = L.row0 + L.nloc;
            if (!stencil_colour_relax(h, L, 1, cur, b, out)) {
                dist_halo_wait(h);
                DISPATCH_A(L, (k_colour_relax<decltype(A)><<<GRID(L.nloc)>>>(A, L.colour, 1, 0, lo, hi, V(L, cur), V(L, b), V(L, out))));
            }
            h->launches++;
            cur = out;
        }
        return launch_smooth(h, L, smoother, omega, sweeps - 1, cur, b);
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
        !(h->flags & OMG_FLAG_NO_FUSED) &&
        stencil_jacobi0_residual_restrict(h, L, C, nullptr, nullptr, nullptr, cfg.omega)) {   // applicability probe
        // zero initial iterate (openmg/__init__.py:191-192): sweep + residual + restriction in one pass over b
        dist_halo_exchange(h, L, L.b);
        {
            ProfScope ps(h, "jacobi0_residual_restrict", l, 16.0 * L.nloc + 8.0 * L.piece_n);
            fused0 = stencil_jacobi0_residual_restrict(h, L, C, L.b, L.xa, V(C, C.b), cfg.omega);
        }
        if (fused0) {
            cur = L.xa;
            h->launches++;
            if (L.slab && !C.slab) dist_allgather(h, V(C, C.b) + L.piece_row0, V(C, C.b), (size_t)L.piece_n);
        } else {
            dist_halo_wait(h);      // the real launch failed after a successful probe: generic path from the zero iterate
        }
    }
    if (!fused0) launch_smooth_residual_restrict(h, l, cfg.smoother, cfg.omega, cfg.pre, &cur, L.b, C.b);
    double *e = cycle_level(h, l + 1, cfg, nullptr);
    // cur's halos were filled for the restriction (unfused path) and cur has not changed since (the single-pass descent
    // kernels only run on unsharded levels, which have no halos)
    return launch_prolong_correct_smooth(h, l, cfg.smoother, cfg.omega, cfg.post, cur, e, L.b, !fused0);
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
    dist_halo_wait(h);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return omg_set_error(OMG_ECUDA, "cycle launch failed: %s", cudaGetErrorString(e));
    return OMG_OK;
}
