// omg_stencil.cuh — structured fast paths (constant 1-D/2-D/3-D band stencils with the
// closed-form restriction).  Each function returns false when the level does not match
// its preconditions; the caller then launches the generic kernel of omg_kernels.cuh.
#pragma once
#include "omg_hier.cuh"

// xo = xi + omega (b - A xi)/diag
bool stencil_jacobi(omg_hierarchy *h, Level &L, const double *xi, const double *b, double *xo, double omega);
// rc = R (b - A x)
bool stencil_residual_restrict(omg_hierarchy *h, Level &L, Level &C, const double *x, const double *b, double *rc);
// y = xi + R^T e ; xo = y + omega (b - A y)/diag      (prolong + correct + first post-smoothing sweep)
bool stencil_prolong_jacobi(omg_hierarchy *h, Level &L, Level &C, const double *xi, const double *e,
                            const double *b, double *xo, double omega);
// xo = omega b/diag (first Jacobi sweep from the zero iterate) ; rc = R (b - A xo)
bool stencil_jacobi0_residual_restrict(omg_hierarchy *h, Level &L, Level &C, const double *b, double *xo,
                                       double *rc, double omega);
// two-colour Gauss-Seidel half sweeps (grid-parity colouring of 3-D levels)
bool stencil_colour_relax(omg_hierarchy *h, Level &L, int colour, const double *xi, const double *b, double *xo);
bool stencil_prolong_colour_relax(omg_hierarchy *h, Level &L, Level &C, int colour, const double *xi,
                                  const double *e, const double *b, double *xo);
// one full two-colour sweep (colour 0 then 1) in a single pass over x; with R^T e added first.  xi == nullptr: probe
bool stencil_rb_sweep(omg_hierarchy *h, Level &L, const double *xi, const double *b, double *xo);
bool stencil_prolong_rb_sweep(omg_hierarchy *h, Level &L, Level &C, const double *xi, const double *e,
                              const double *b, double *xo);
// one full two-colour sweep from the zero iterate, x never read (coarse levels of a cycle).  b == nullptr: probe
bool stencil_rb_sweep0(omg_hierarchy *h, Level &L, const double *b, double *xo);
// xo = xi + omega (b - A xi)/diag ; rc = R (b - A xo): last pre-smoothing sweep + restricted residual in one pass over x
// (unsharded pure-band levels: k_jr3 in 3-D, k_jr2 in 2-D / 1-D).  xi == nullptr: probe
bool stencil_jacobi_residual_restrict(omg_hierarchy *h, Level &L, Level &C, const double *xi, const double *b,
                                      double *xo, double *rc, double omega);
