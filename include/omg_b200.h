/* omg_b200.h — C-ABI of libomg_b200.so: the B200 (sm_100a) implementation of
 * openmg's multigrid V-cycle hot path.
 *
 * This is the drop-in boundary.  The reference (tsbertalan/openmg) is pure
 * Python with no FFI of its own; each entry point below states which
 * reference function (file:line, relative to the reference root) it replaces.
 * The Python package `openmg_b200` binds these with ctypes and re-exposes the
 * reference's own names (mgSolve, mgCycle, smooth, coarseSolve, operators.*,
 * tools.*).  INTEGRATION.md shows the stub a maintainer of the reference
 * would add.
 *
 * Conventions
 *  - plain C types only; all index arrays are int32 (scipy's default), all
 *    values fp64; "host" pointers are caller-owned host memory, never freed
 *    or retained by the library after the call returns.
 *  - every function returns 0 on success or an OMG_E* code; the message of
 *    the last failure on the calling thread is omg_last_error().
 *  - there is NO CPU fallback: without a CUDA device omg_init() fails.
 *  - one omg_hierarchy is used by one host thread at a time.
 */
#ifndef OMG_B200_H
#define OMG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct omg_hierarchy omg_hierarchy;

enum {
    OMG_OK = 0,
    OMG_EINVAL = 1,       /* bad argument                       -> ValueError   */
    OMG_ECUDA = 2,        /* CUDA runtime failure               -> RuntimeError */
    OMG_ENODEV = 3,       /* no usable CUDA device              -> RuntimeError */
    OMG_ESHAPE = 4,       /* restriction() would have 0/1 rows  -> ValueError (openmg/operators.py:53-56) */
    OMG_EDIM = 5,         /* more than 3 dimensions             -> ValueError (openmg/operators.py:69-71,273) */
    OMG_EINDEX = 6,       /* restriction column out of range    -> IndexError (lil_matrix, openmg/operators.py:75-84) */
    OMG_ESINGULAR = 7,    /* zero diagonal / singular coarse op -> ZeroDivisionError / LinAlgError */
    OMG_ENOMEM = 8,       /* device memory                      -> MemoryError  */
    OMG_ENCCL = 9,        /* NCCL failure                       -> RuntimeError */
    OMG_EUNSUPPORTED = 10 /* valid in the reference, not implemented here -> NotImplementedError */
};

/* smoother ids (parameters['smoother']) */
enum {
    OMG_SMOOTH_JACOBI = 0, /* weighted Jacobi, x += omega*(b-Ax)/diag                       (new; row update = openmg/solvers.py:68) */
    OMG_SMOOTH_RBGS = 1,   /* two-colour Gauss-Seidel, colour rule of oracle.colouring()     (replaces openmg/solvers.py:34-75)      */
    OMG_SMOOTH_LEXGS = 2   /* the reference's lexicographic GS, sequential single-CTA kernel (openmg/solvers.py:56-68); parity only  */
};

/* level operator storage kinds reported by omg_level_info */
enum { OMG_KIND_BAND = 0, OMG_KIND_BAND_EXC = 1, OMG_KIND_CSR = 2 };

/* hierarchy creation flags */
enum {
    OMG_FLAG_FORCE_CSR = 1,   /* never use the constant-band fast path (testing)                 */
    OMG_FLAG_NO_GRAPH = 2,    /* launch kernels directly instead of replaying a CUDA graph       */
    OMG_FLAG_NO_FUSED = 4,    /* use the unfused generic kernels only (testing / A-B comparison)  */
    OMG_FLAG_KEEP_CSR = 8,    /* keep full CSR copies of band levels on device after setup        */
    OMG_FLAG_FACTOR = 16      /* omg_operator_create_csr: also build the dense direct-solve factor */
};

/* ---- process / device --------------------------------------------------- */

/* Select the CUDA device of this process (one process per GPU) and create the
 * library's streams.  device < 0 -> LOCAL_RANK env or 0.  Idempotent. */
int omg_init(int device);
void omg_finalize(void);
const char *omg_last_error(void);
/* "sm_100a B200 148SM ..." style description of the active device. */
int omg_device_info(char *buf, int buflen, int *sm_count, int64_t *mem_bytes);

/* Multi-GPU: row-slab sharding of every fine level across `nranks` processes.
 * The caller (Python, via torch.distributed) broadcasts the 128-byte NCCL
 * unique id produced by rank 0. */
int omg_nccl_unique_id(unsigned char id[128]);
int omg_dist_init(int rank, int nranks, const unsigned char id[128]);
int omg_dist_rank(int *rank, int *nranks);
/* Pure host logic: the row-slab partition of a hierarchy.  level_lead[l] = leading grid extent
 * (problemShape[0] >> l), level_rows[l] = rows of A_l, level_regular[l] = level may be a slab
 * (closed-form restriction, band operator).  Levels [0, *first_replicated) are slabs cut so that no
 * restriction aggregate straddles two ranks; the others are replicated on every rank. */
int omg_partition(int nlevels, const int64_t *level_lead, const int64_t *level_rows,
                  const int32_t *level_regular, int nranks, int rank, int64_t agglomerate_below,
                  int32_t *first_replicated, int64_t *row0, int64_t *nloc);
/* rows [row0, row0+nloc) of level `level` live on this rank; slab = 0 for replicated levels */
int omg_level_partition(const omg_hierarchy *h, int level, int64_t *row0, int64_t *nloc, int *slab);

/* pinned host staging buffers for the host<->device legs of omg_solve */
int omg_host_alloc(void **ptr, int64_t bytes);
int omg_host_free(void *ptr);

/* ---- setup: level hierarchy (A_l, R_l, P_l = R_l^T) ---------------------- */

/* Replaces operators.restrictionList (openmg/operators.py:92-141) +
 * operators.coeffecientList (openmg/operators.py:144-188) as called by
 * mgSolve (openmg/__init__.py:103-109): uploads A_0 once as device CSR,
 * detects the constant-band (stencil) structure, builds R_l in closed form,
 * and the Galerkin operators A_{l+1} = R_l A_l R_l^T on the device.
 * shape[ndim] = parameters['problemShape']; coarsestLevel, minSize as in the
 * reference (depth rule openmg/operators.py:128-140).  indptr has n+1 entries. */
int omg_hierarchy_create_csr(omg_hierarchy **out, int ndim, const int64_t *shape,
                             int coarsestLevel, int minSize, int64_t n,
                             const int32_t *indptr, const int32_t *indices,
                             const double *data, int flags);

/* Same, for A_0 = diag*I + sum_k coeffs[k]*(S^{+offsets[k]} + S^{-offsets[k]})
 * truncated at the two global ends — the matrices of operators.poisson
 * (openmg/operators.py:191-256) — without ever materialising a host CSR
 * (needed from 512^3 up).  offsets[k] > 0. */
int omg_hierarchy_create_band(omg_hierarchy **out, int ndim, const int64_t *shape,
                              int coarsestLevel, int minSize, int64_t n, double diag,
                              int nband, const int64_t *offsets, const double *coeffs,
                              int flags);

/* A single operator (1 level, no restriction) for the standalone reference entry points
 * solvers.smooth / gaussSeidel / smoothToThreshold (openmg/solvers.py:28-75),
 * tools.getresidual / flexibleMmult (openmg/tools.py:12-26) and, with OMG_FLAG_FACTOR,
 * solvers.coarseSolve (openmg/solvers.py:16-26).  Usable with the level-0 unit entry points. */
int omg_operator_create_csr(omg_hierarchy **out, int64_t n, const int32_t *indptr, const int32_t *indices,
                            const double *data, int flags);

void omg_hierarchy_destroy(omg_hierarchy *h);

/* number of grids = len(R)+1 (openmg/operators.py:169-170) */
int omg_level_count(const omg_hierarchy *h, int *nlevels);
/* n = rows of A_l; nnzA = stored entries of A_l; nnzR = entries of R_l (0 on the
 * coarsest); kind = OMG_KIND_*; nexc = rows deviating from the band stencil. */
int omg_level_info(const omg_hierarchy *h, int level, int64_t *n, int64_t *nnzA,
                   int64_t *nnzR, int *kind, int64_t *nexc);
/* band description of a level (nband <= 16 signed offsets); returns nband=0 for CSR levels */
int omg_level_band(const omg_hierarchy *h, int level, double *diag, int *nband,
                   int64_t *offsets, double *coeffs);
/* Canonical (sorted, zero-free) CSR of A_l / R_l for infoDict['A'], infoDict['R']
 * (openmg/__init__.py:142-143).  Buffers sized from omg_level_info. */
int omg_level_export_A(const omg_hierarchy *h, int level, int32_t *indptr, int32_t *indices, double *data);
int omg_level_export_R(const omg_hierarchy *h, int level, int32_t *indptr, int32_t *indices, double *data);
/* setup timings of the last create call, milliseconds */
int omg_setup_times(const omg_hierarchy *h, double *upload_ms, double *galerkin_ms, double *coarse_factor_ms);

/* ---- standalone operators (device-built, exported to host CSR) ----------- */

/* operators.restriction(shape) (openmg/operators.py:15-89).  Call with
 * indptr==NULL to query n (rows) and nnz. */
int omg_restriction(int ndim, const int64_t *shape, int64_t *n_rows, int64_t *nnz,
                    int32_t *indptr, int32_t *indices, double *data);

/* ---- solve --------------------------------------------------------------- */

/* The cycle loop of mgSolve (openmg/__init__.py:112-138) around mgCycle
 * (openmg/__init__.py:151-236): copies b (and x if has_initial) host->device,
 * runs V(pre,post) cycles until `cycles` (>0) are done or the level-0 residual
 * 2-norm falls below `threshold` (>0), copies x back.  At least one cycle always
 * runs; cycles<=0 && threshold<=0 is OMG_EINVAL *after* that cycle, as in the
 * reference (:112 then :118-119).  norm_hist (nullable, capacity hist_cap)
 * receives the per-cycle norms when requested or needed for the stop rule. */
int omg_solve(omg_hierarchy *h, const double *b_host, double *x_host, int has_initial,
              int pre, int post, int smoother, double omega, int cycles, double threshold,
              int *cycles_done, double *final_norm, double *norm_hist, int hist_cap);

/* Diagnostics of the last omg_solve / of the setup: host_syncs = stream synchronisations the cycle loop issued
 * (a thresholded solve tests its stop rule on the device, openmg/__init__.py:118-138, and synchronises once per
 * batch of cycles, not once per cycle); coarse_defect = max |A_L * Ainv - I| of the coarse factor
 * (openmg/solvers.py:16-26 solves with pivoted SuperLU; the device inverse is only kept unpivoted when this is
 * at rounding level). */
int omg_solve_stats(const omg_hierarchy *h, int64_t *host_syncs, double *coarse_defect);

/* One mgCycle(A, b, level, R, parameters, initial) (openmg/__init__.py:151-236) entered at
 * `level`: b_host / x_host have n_level entries; x_host is the initial iterate when
 * has_initial, and receives uOut.  norm = ||b - A_level uOut||_2 (:227; 0 on the coarsest, :232). */
int omg_cycle(omg_hierarchy *h, int level, const double *b_host, double *x_host, int has_initial,
              int pre, int post, int smoother, double omega, double *norm);

/* Device-resident timing: b already in HBM (set by omg_set_rhs), zero initial
 * guess, `ncycles` back-to-back V-cycles timed with CUDA events on the
 * library's stream.  ms = total milliseconds; launches = kernels launched. */
int omg_set_rhs(omg_hierarchy *h, const double *b_host);
int omg_bench_cycles(omg_hierarchy *h, int pre, int post, int smoother, double omega,
                     int ncycles, int with_norm, float *ms, int64_t *launches);
/* Per-kernel CUDA-event timing of `reps` V-cycles launched directly (no graph) on the rhs set
 * by omg_set_rhs: JSON array of {"name","level","launches","ms" (avg per launch),"bytes"
 * (algorithmic bytes per launch)} written to json[cap]. */
int omg_profile_cycle(omg_hierarchy *h, int pre, int post, int smoother, double omega, int reps,
                      char *json, int cap);
int omg_get_solution(omg_hierarchy *h, double *x_host);
int omg_current_norm(omg_hierarchy *h, double *norm);

/* ---- unit entry points: one reference operation each (parity tests) ------ */

/* smooth / gaussSeidel (openmg/solvers.py:28-75) on level `level`: x in/out, `sweeps` iterations. */
int omg_smooth(omg_hierarchy *h, int level, const double *b_host, double *x_host, int sweeps,
               int smoother, double omega);
/* smoothToThreshold (openmg/solvers.py:31-32,43-50): sweep until ||b - A x||_2 < threshold
 * (checked before the first sweep and after every sweep) or max_sweeps; *sweeps_done out. */
int omg_smooth_to_threshold(omg_hierarchy *h, int level, const double *b_host, double *x_host, double threshold,
                            int max_sweeps, int smoother, double omega, int *sweeps_done, double *norm);
/* R_l (b - A_l x)  — getresidual + restriction (openmg/__init__.py:209-210), fused. rc has n_{l+1} entries. */
int omg_residual_restrict(omg_hierarchy *h, int level, const double *b_host, const double *x_host, double *rc_host);
/* The descent step of mgCycle on one level (openmg/__init__.py:201 pre-smoothing, :209-210 residual + restriction):
 * x := smooth(A_l, b, x, sweeps) in/out, rc := R_l (b - A_l x) with n_{l+1} entries.  The last Jacobi sweep and the
 * restricted residual run as one pass over x where the level allows it (k_jr3 / k_jr2). */
int omg_smooth_residual_restrict(omg_hierarchy *h, int level, const double *b_host, double *x_host, int sweeps,
                                 int smoother, double omega, double *rc_host);
/* x += R_l^T e   (openmg/__init__.py:214,224) */
int omg_prolong_correct(omg_hierarchy *h, int level, const double *ec_host, double *x_host);
/* smooth(A, b, x + R^T e, sweeps) (openmg/__init__.py:216-222), fused */
int omg_prolong_correct_smooth(omg_hierarchy *h, int level, const double *b_host, const double *ec_host,
                               double *x_host, int sweeps, int smoother, double omega);
/* coarseSolve (openmg/solvers.py:16-26) on the coarsest level */
int omg_coarse_solve(omg_hierarchy *h, const double *b_host, double *x_host);
/* ||b - A_l x||_2 (openmg/__init__.py:227) and the residual vector itself (openmg/tools.py:12-15) */
int omg_residual_norm(omg_hierarchy *h, int level, const double *b_host, const double *x_host, double *norm);
int omg_residual(omg_hierarchy *h, int level, const double *b_host, const double *x_host, double *r_host);
/* y = A_l x (tools.flexibleMmult(A, x), openmg/tools.py:18-26) */
int omg_matvec(omg_hierarchy *h, int level, const double *x_host, double *y_host);

#ifdef __cplusplus
}
#endif
#endif /* OMG_B200_H */
