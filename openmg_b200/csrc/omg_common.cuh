// omg_common.cuh — shared types of libomg_b200 (host + device).
//
// Data layout in HBM:
//  * every level vector (x, x', b) is one allocation [pad | owned rows | pad];
//    the pads are zero on the two global ends (they implement the reference's
//    "out-of-range band terms are dropped" semantics, SURVEY §0.2 item 4,
//    branch-free) and are neighbour halos on interior slab cuts (multi-GPU).
//    Kernels receive the pointer to owned row 0 and index relative to it.
//  * a level operator is either BAND (constant-coefficient truncated-Toeplitz
//    stencil in the flat index, optionally + a compact CSR of "exception rows"
//    selected by a bit mask) or full CSR.
//  * restriction R_l is matrix-free (closed form of openmg/operators.py:15-89)
//    for "regular" shapes; prolongation is R_l^T with the same weights
//    (openmg/__init__.py:214).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/omg_b200.h"

#define OMG_MAXBAND 16

struct BandOp {
    int nb;                  // number of off-diagonal taps (both signs listed)
    int off[OMG_MAXBAND];    // signed flat-index offsets, ascending
    double coef[OMG_MAXBAND];
    double diag;
};

// exception rows of a BAND level: rows whose entries differ from the stencil
struct ExcOp {
    const unsigned *mask;    // bit i%32 of word i/32: row i is an exception
    const int *wpre;         // exclusive prefix of popc(mask[w])
    const int *ptr;          // CSR over exception slots
    const int *col;          // local-relative column (may reach into the pads)
    const double *val;
    const double *diag;      // a_ii per slot
};

struct CsrOp {
    const int *ptr;
    const int *col;          // local-relative
    const double *val;
    const double *diag;      // a_ii per row
};

// Sliced ELLPACK view of a CSR level (slices of 32 rows, entries stored column-major in PAIRS): entry k of row i
// sits at off[i >> 5] + (k >> 1) * 64 + (i & 31) * 2 + (k & 1), so a thread reads its row with one 128-bit value
// load and one 64-bit column load per two entries and a warp's loads are contiguous 512 B / 256 B segments.
// Rows keep their CSR entry order (bit-identical sums); slots beyond a row's length are never read.
struct SellOp {
    const int *off;          // per slice: first entry (even)
    const int *rlen;         // per row: entries
    const int *col;
    const double *val;
    const double *diag;
};

// closed-form restriction for regular shapes (all dims even, C-order strides
// equal to the reference's NX / NX*NY offsets).  Dims padded to 3 with leading 1s.
struct RegR {
    int alpha;               // 1..3
    int k;                   // entries per row = 2^alpha
    int fs1, fs2;            // fine trailing dims   (fine  flat = (i0*fs1 + i1)*fs2 + i2)
    int cs1, cs2;            // coarse trailing dims (coarse flat = (I0*cs1 + I1)*cs2 + I2)
    int o[8];                // fine column offsets of one coarse row, ascending
    int nc, nf;              // global rows coarse / fine
    double w;                // 1/2^alpha
};

// Positional stencil classes of a 3-D Galerkin level.  The reference's operators have no boundary breaks, so their
// Galerkin images deviate from the interior stencil exactly on the x- and y-boundaries of the grid, and all rows of
// one position class (cx, cy in {first, interior, last}) share one row pattern.  A class is stored as the
// correction taps (row entry minus band entry); class index = 3*cy + cx, class 4 is the interior (no taps).
#define OMG_CLS_TAPS 6
struct ClsTab {
    int ntap[9];
    int soff[9][OMG_CLS_TAPS];     // offset inside the plane (flat, may reach the +-1 halo rows)
    int dz[9][OMG_CLS_TAPS];       // plane offset -1, 0, +1
    double coef[9][OMG_CLS_TAPS];
};

// two-colouring rule of a level (oracle.colouring)
struct ColourRule {
    int flat;                // 1: colour = i & 1
    int alpha;
    int s1, s2;              // trailing dims of the level shape (C order)
};

struct OmgError {
    int code;
    std::string msg;
};

int omg_set_error(int code, const char *fmt, ...);

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return omg_set_error(_e == cudaErrorMemoryAllocation ? OMG_ENOMEM : OMG_ECUDA,   \
                                 "%s:%d %s -> %s", __FILE__, __LINE__, #expr,                \
                                 cudaGetErrorString(_e));                                    \
    } while (0)

#define OMG_TRY(expr)                 \
    do {                              \
        int _rc = (expr);             \
        if (_rc != OMG_OK) return _rc; \
    } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
