// omg_stencil.cu — structured fast paths; see omg_stencil.cuh.
#include "omg_stencil.cuh"
#include "omg_kernels.cuh"

bool stencil_jacobi(omg_hierarchy *, Level &, const double *, const double *, double *, double) { return false; }
bool stencil_residual_restrict(omg_hierarchy *, Level &, Level &, const double *, const double *, double *) {
    return false;
}
bool stencil_prolong_jacobi(omg_hierarchy *, Level &, Level &, const double *, const double *, const double *,
                            double *, double) {
    return false;
}
