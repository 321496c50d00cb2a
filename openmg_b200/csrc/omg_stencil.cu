// omg_stencil.cu — structured fast path for 3-D constant band stencils
//   A = d I + c1 (S^1 + S^-1) + cS (S^S1 + S^-S1) + cP (S^S2 + S^-S2)      (flat index, no
//   boundary breaks — the matrices of openmg/operators.py:244-256 and their Galerkin images)
// with the closed-form 2x2x2 restriction (openmg/operators.py:73-84).
//
// Design (sm_100a):
//  * 2.5-D blocking.  A CTA owns TY full rows of every plane ("chunk", W = TY*S1 points) of a
//    z-segment and marches along z.  Because rows are full, the chunk plus its +-S1 halo rows is
//    ONE contiguous span of the flat vector per plane, so each plane is staged into shared
//    memory by a single TMA bulk copy (cp.async.bulk.shared.global, SASS UBLKCP) completing on
//    an mbarrier; a ring of NS spans keeps planes z-1, z, z+1 resident and the rest in flight.
//    The zero pads of the vectors make the global ends branch-free.
//  * Each thread owns 2x2 (x,y) patches: all 7 taps are shared-memory reads (LDS.128 for the
//    aligned pairs), b streams through ld.global.nc, results leave as 128-bit stores.
//  * MODE 1 (residual+restriction) accumulates the 2x2 patch over two planes in registers and
//    writes one coarse value: the fine residual never exists in HBM.
//  * MODE 2 (prolong+correct+Jacobi) and MODE 3 (first Jacobi sweep from zero + residual +
//    restriction) transform each staged plane in place in shared memory when it lands
//    (y = x + w e[agg], resp. x = omega b / a_ii), so the 7-point pass itself is unchanged.
//  * Rows deviating from the stencil ("exception rows" of Galerkin levels) are recomputed from
//    their compact CSR by a small fix-up kernel right after (out-of-place, so still exact).
#include <stdlib.h>

#include "omg_stencil.cuh"
#include "omg_kernels.cuh"

#define ST_PPT 2           // 2x2 patches per thread per plane (max)
#define ST_TPT 6           // span pairs per thread in the in-smem transform (max)
#define ST_MAXNS 6

struct St3 {
    const double *xi;      // staged vector: owned row 0 (zero/halo padded).  MODE 3: this is b.
    const double *b;
    double *xo;            // fine output (MODE 0,2: new iterate; MODE 3: x = omega b/diag)
    double *rc;            // coarse output (MODE 1,3)
    const double *e;       // coarse correction (MODE 2)
    ExcOp exc;             // exception rows (MODE 3 needs their a_ii)
    int has_exc;
    int nloc;
    int S1, S2, NZ;        // row length, plane size, local planes
    int TY, NP, SPAN;      // rows per chunk, patches per plane-chunk, staged elements per plane (TY+2)*PITCH
    int XW, XC, PITCH, XH; // chunk width, chunks per row, staged row pitch (S1, or XW+4), x-halo columns (0 or 2)
    int NS;                // ring stages
    int ZL;                // planes per z-segment (even)
    int cs1, cs2;          // coarse rows per plane, coarse row length
    int NYg;               // rows per plane
    int zg0, NZg;          // global index of local plane 0, global planes
    int cz0;               // global index of local coarse plane 0
    int colour;            // -1: all rows (Jacobi); 0/1: only rows of that grid-parity colour are relaxed
    int zlo, zhi, boundary;   // planes [zlo, zhi) in segments of ZL; `boundary`: CTA row 0 -> [0, zlo), row 1 -> [zhi, NZ)
    // Fused halo pull (slab levels over NVLink peer memory, see st_pull_*): one grid of `nseg` interior segment rows
    // plus two boundary rows ([0, zlo) and [zhi, NZ)) whose CTAs stage the planes -1 / NZ straight from the
    // neighbours' copies of the vector once those are final.  pull == nullptr: not a fused launch.
    unsigned long long *pull, *pull_dn, *pull_up;   // this slot's 8 flag words here / on the neighbours
    const double *xi_dn, *xi_up;                    // the neighbours' copies of the staged vector (their owned row 0)
    const double *e_dn, *e_up;                      // MODE 2: the neighbours' coarse correction (owned row 0)
    int e_nzc;                                      // MODE 2: coarse planes of e this rank owns
    int *pull_timeout;
    int nseg;
    int bpf;               // planes by which thread 0 prefetches b into L2 ahead of its loads (0: off)
    int epf;               // the same for the coarse correction e (MODE 2), in fine planes
    int needs_fix;         // exception rows are corrected by fix-up kernels after this one (they read local halos)
    int use_cls;           // rows on the x/y grid boundaries get their class correction taps in-kernel (no fix-up)
    ClsTab cls;
    double d, c1, cS, cP, wod, w, omega;   // wod = omega/d
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// TMA bulk prefetch into L2 (no destination, no registers): thread 0 pulls the b rows a later step will load with
// ld.global.nc out of HBM ahead of time, so those loads are L2 hits instead of exposed DRAM latency.
__device__ __forceinline__ void bulk_prefetch_l2(const double *src, long long lo, long long hi) {
    if (hi > lo)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + lo), "r"((uint32_t)((hi - lo) * 8))
                     : "memory");
}

__device__ __forceinline__ double2 lds2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void sts2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }
__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }



// ---------------------------------------------------------------- fused halo pull (compute + neighbour exchange in one kernel)
//
// Slab levels: the planes next to a cut need one plane (+ one row) of the neighbour's vector.  Instead of a separate
// exchange (flag kernels + copy-engine pulls into local halo rows, then a second "boundary" launch behind a stream
// dependency) the stencil kernel does it itself over NVLink peer memory.  Per (level, vector) slot, 8 words in
// memory the neighbours have mapped: {kernels done, ready<-dn, ready<-up, arrivals, pulled<-dn, pulled<-up}.
//   G = kernels done + 1 identifies this launch on every rank (all ranks run the same kernel sequence).
//   CTA (0,0), first wave:   my vector is final (stream order) -> ready := G in both neighbours' words
//   boundary CTAs, last rows: wait for ready<-neighbour >= G, then the TMA ring stages plane -1 / NZ from the
//                            neighbour's copy (cp.async.bulk from the peer-mapped address); everything else is local
//   last CTA to arrive:      pulled := G in both neighbours' words, wait until both have pulled from me (so nothing
//                            that follows this kernel — not even a host copy — can overwrite rows still being
//                            read), kernels done := G
// The boundary planes of the OUTPUT are written by the boundary CTAs only, i.e. after the neighbours reported the
// vector they are reading final, which orders them behind the neighbours' previous kernel: the ping-pong partner
// of the staged vector is never overwritten while a neighbour still reads it.
// Every wait is bounded (~4 s): on expiry a flag is raised (checked by the host after the solve) and the kernel
// carries on, so a protocol error cannot hang the GPU.
__device__ __forceinline__ void st_sys_release(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long st_sys_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_wait_ge(const unsigned long long *p, unsigned long long G, int *timeout_flag) {
    const long long t0 = clock64();
    while (st_sys_acquire(p) < G) {
        if (clock64() - t0 > 8000000000ll) {
            *timeout_flag = 1;
            break;
        }
        __nanosleep(100);
    }
}
// thread 0 of CTA (0,0): publish; thread 0 of a boundary CTA: wait for the neighbour this piece reads from
__device__ __forceinline__ void st_pull_begin(const St3 &P, int brow, unsigned long long G) {
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0 && blockIdx.y == 0) {
            __threadfence_system();
            if (P.pull_dn) st_sys_release(P.pull_dn + 2, G);       // I am the up neighbour of dn
            if (P.pull_up) st_sys_release(P.pull_up + 1, G);       // and the dn neighbour of up
        }
        if (brow == 0 && P.pull_dn) st_wait_ge(P.pull + 1, G, P.pull_timeout);
        if (brow == 1 && P.pull_up) st_wait_ge(P.pull + 2, G, P.pull_timeout);
        if (brow >= 0) asm volatile("fence.proxy.async;" ::: "memory");
    }
    if (brow >= 0) __syncthreads();
}
// thread 0 of CTA (0,0) (after publishing) and of every boundary CTA (after its last plane)
__device__ __forceinline__ void st_pull_arrive(const St3 &P, unsigned long long G, unsigned int nboundary) {
    __threadfence_system();
    if (atomicAdd(P.pull + 3, 1ull) == (unsigned long long)nboundary) {       // nboundary + 1 arrivals in all
        P.pull[3] = 0ull;
        if (P.pull_dn) st_sys_release(P.pull_dn + 5, G);       // "your up neighbour has pulled"
        if (P.pull_up) st_sys_release(P.pull_up + 4, G);       // "your dn neighbour has pulled"
        if (P.pull_dn) st_wait_ge(P.pull + 4, G, P.pull_timeout);
        if (P.pull_up) st_wait_ge(P.pull + 5, G, P.pull_timeout);
        __threadfence();
        *(volatile unsigned long long *)P.pull = G;
    }
}

// MODE 0: xo = xi + omega (b - A xi)/d
// MODE 1: rc = R (b - A xi)
// MODE 2: y = xi + R^T e ; xo = y + omega (b - A y)/d
// MODE 3: xo = omega b/diag (first Jacobi sweep from x = 0) ; rc = R (b - A xo)      [xi == b]
// MODE 4: the same for a uniform diagonal: x = (omega/d) b needs no transform pass — the 7-point pass runs on the staged
//         b planes and its result is scaled (no second barrier per plane, b is not loaded a second time): level 1 of the
//         512^3 hierarchy 0.105 -> 0.070 ms
// PULL: the fused-halo-pull instantiation (slab levels); the single-GPU instantiation carries none of its code — the
// prolong+Jacobi kernel loses 15 % when merely compiled with it (0.57 -> 0.65 ms at 512^3).
template <int MODE, int NT, bool CLS, bool PULL>
__global__ void __launch_bounds__(NT, (NT <= 256 ? 2 : 1)) k_st3(const St3 P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)P.NS * P.SPAN * sizeof(double));
    constexpr bool XF = (MODE == 2 || MODE == 3);
    const int NS = P.NS;
    const int tid = threadIdx.x;
    // blockIdx.y: interior segment, or (fused-pull / boundary launches) one of the two boundary pieces
    const int brow = PULL ? (int)blockIdx.y - P.nseg : -1;
    const int z0 = PULL ? (brow >= 0 ? (brow == 0 ? 0 : P.zhi) : P.zlo + (int)blockIdx.y * P.ZL)
                        : (P.boundary ? (blockIdx.y == 0 ? 0 : P.zhi) : P.zlo + (int)blockIdx.y * P.ZL);
    const int z1 = PULL ? (brow >= 0 ? (brow == 0 ? P.zlo : P.NZ) : min(z0 + P.ZL, P.zhi))
                        : (P.boundary ? (blockIdx.y == 0 ? P.zlo : P.NZ) : min(z0 + P.ZL, P.zhi));
    const int y0 = blockIdx.x * P.TY;
    const long long span0 = (long long)y0 * P.S1 - P.S1;      // in-plane start of the span (row y0-1)
    const uint32_t span_bytes = (uint32_t)P.SPAN * 8u;
    unsigned long long pullG = 0;
    if constexpr (PULL) {
        pullG = *(volatile unsigned long long *)P.pull + 1ull;     // stable until this launch's last arrival
        st_pull_begin(P, brow, pullG);
        if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) st_pull_arrive(P, pullG, 2u * gridDim.x);
    }
    // Stage the span of plane p into ring slot k.  Fused pull: the part of the span below row 0 of the slab comes from
    // the lower neighbour's copy of the vector, the part beyond the last row from the upper neighbour's (a span is
    // whole rows, the slab a whole number of rows; the +-1-row halo of the first / last chunk makes the spans of the
    // planes 0 and NZ-1 straddle a cut).  Without neighbour (or not fused): the local pad / halo rows.
    auto issue_span = [&](int k, int p) {
        const long long f0 = (long long)p * P.S2 + span0;
        mbar_expect_tx(full + k, span_bytes);
        double *dst = stage + (size_t)k * P.SPAN;
        if constexpr (!PULL) {
            bulk_g2s(dst, P.xi + f0, span_bytes, full + k);
            return;
        }
        const long long f1 = f0 + P.SPAN, nl = (long long)P.NZ * P.S2;
        long long a = f0, b = f1 < 0 ? f1 : 0;
        if (b > a) bulk_g2s(dst, P.xi_dn ? P.xi_dn + (nl + a) : P.xi + a, (uint32_t)((b - a) * 8), full + k);
        a = f0 > 0 ? f0 : 0;
        b = f1 < nl ? f1 : nl;
        if (b > a) bulk_g2s(dst + (a - f0), P.xi + a, (uint32_t)((b - a) * 8), full + k);
        a = f0 > nl ? f0 : nl;
        b = f1;
        if (b > a) bulk_g2s(dst + (a - f0), P.xi_up ? P.xi_up + (a - nl) : P.xi + a, (uint32_t)((b - a) * 8), full + k);
    };

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < NS; ++k) {
            int p = z0 - 1 + k;
            if (p > z1) break;
            if constexpr (PULL) {
                issue_span(k, p);
            } else {
                mbar_expect_tx(full + k, span_bytes);
                bulk_g2s(stage + (size_t)k * P.SPAN, P.xi + (long long)p * P.S2 + span0, span_bytes, full + k);
            }
        }
    }

    // fixed 2x2 patch assignment
    const int HX = P.S1 >> 1;
    int px[ST_PPT], py[ST_PPT];
    bool act[ST_PPT];
#pragma unroll
    for (int k = 0; k < ST_PPT; ++k) {
        int p = tid + k * NT;
        act[k] = p < P.NP;
        p = act[k] ? p : 0;
        py[k] = p / HX;
        px[k] = p - py[k] * HX;
        if (CLS && k == 1 && (HX & 63) == 0 && (NT % HX) == 0) {
            // the class corrections run on the lanes px == 0 and px == HX-1 only: rotate the second patch of every
            // thread by half a row so that each warp owns one of them instead of half the warps owning two.
            // (Only when a thread's two patches sit in different rows, NT a multiple of HX: with 128 threads on
            // 512-wide rows — thin slabs of a sharded level — both are in the same row and the rotation would map
            // the second half of the row onto the first.)
            px[k] += HX >> 1;
            if (px[k] >= HX) px[k] -= HX;
        }
    }
    double acc[ST_PPT];
#pragma unroll
    for (int k = 0; k < ST_PPT; ++k) acc[k] = 0.0;

    // fixed span-pair assignment of the in-smem transform
    int tq[ST_TPT], tx[ST_TPT];
    bool tact[ST_TPT];
    double aux[ST_TPT];        // MODE 2: e of the pair's coarse cell
    unsigned auxm[ST_TPT];     // MODE 3: exception mask word of the pair
    if (XF) {
#pragma unroll
        for (int k = 0; k < ST_TPT; ++k) {
            int t = tid + k * NT;
            tact[k] = 2 * t < P.SPAN;
            t = tact[k] ? t : 0;
            tq[k] = (2 * t) / P.S1;
            tx[k] = 2 * t - tq[k] * P.S1;
            aux[k] = 0.0;
            auxm[k] = 0u;
        }
    }
    // aux of local plane p for pair k
    auto load_aux = [&](int p) {
#pragma unroll
        for (int k = 0; k < ST_TPT; ++k) {
            if (!tact[k]) continue;
            if (MODE == 2) {
                int yy = y0 - 1 + tq[k];
                int zz = p + P.zg0;
                if (yy < 0) {              // flat-index wrap into the neighbouring plane
                    yy = P.NYg - 1;
                    zz -= 1;
                } else if (yy >= P.NYg) {
                    yy = 0;
                    zz += 1;
                }
                double v = 0.0;
                if constexpr (!PULL) {
                    if (zz >= 0 && zz < P.NZg)
                        v = __ldg(P.e + ((long long)((zz >> 1) - P.cz0) * P.cs1 + (yy >> 1)) * P.cs2 + (tx[k] >> 1));
                } else if (zz >= 0 && zz < P.NZg) {
                    int cz = (zz >> 1) - P.cz0;
                    const long long eo = (long long)(yy >> 1) * P.cs2 + (tx[k] >> 1);
                    if (PULL && cz < 0 && P.e_dn)                   // the neighbours' coarse planes: NVLink loads
                        v = __ldcv(P.e_dn + (long long)(cz + P.e_nzc) * P.cs1 * P.cs2 + eo);
                    else if (PULL && cz >= P.e_nzc && P.e_up)
                        v = __ldcv(P.e_up + (long long)(cz - P.e_nzc) * P.cs1 * P.cs2 + eo);
                    else
                        v = __ldg(P.e + (long long)cz * P.cs1 * P.cs2 + eo);
                }
                aux[k] = v;
            } else if (MODE == 3) {
                unsigned m = 0u;
                if (P.has_exc) {
                    long long gl = (long long)p * P.S2 + span0 + (long long)tq[k] * P.S1 + tx[k];
                    if (gl >= 0 && gl < P.nloc) m = __ldg(P.exc.mask + (gl >> 5));
                }
                auxm[k] = m;
            }
        }
    };
    auto transform = [&](int slot, int p) {
        double *sp_ = stage + (size_t)slot * P.SPAN;
#pragma unroll
        for (int k = 0; k < ST_TPT; ++k) {
            if (!tact[k]) continue;
            int o = tq[k] * P.S1 + tx[k];
            double2 v = lds2(sp_ + o);
            if (MODE == 2) {
                v.x += P.w * aux[k];
                v.y += P.w * aux[k];
            } else {
                double s0 = P.wod, s1 = P.wod;
                if (P.has_exc) {
                    unsigned m = auxm[k];
                    long long gl = (long long)p * P.S2 + span0 + o;
                    int bit = (int)(gl & 31);
                    if ((m >> bit) & 3u) {
                        int base = __ldg(P.exc.wpre + (gl >> 5));
                        if ((m >> bit) & 1u)
                            s0 = P.omega / __ldg(P.exc.diag + base + __popc(m & ((1u << bit) - 1u)));
                        if ((m >> (bit + 1)) & 1u)
                            s1 = P.omega / __ldg(P.exc.diag + base + __popc(m & ((2u << bit) - 1u)));
                    }
                }
                v.x *= s0;
                v.y *= s1;
            }
            sts2(sp_ + o, v);
        }
    };

    if (XF) {
        // planes z0-1 and z0 are transformed up front, z0+1 inside the loop
        load_aux(z0 - 1);
        mbar_wait(full + 0, 0);
        transform(0, z0 - 1);
        load_aux(z0);
        mbar_wait(full + 1, 0);
        transform(1, z0);
        load_aux(z0 + 1);
    }

    for (int z = z0; z < z1; ++z) {
        const int q = z - (z0 - 1);                // ring position of plane z (plane z0-1 is 0)
        // b for this plane: issue the loads before blocking on the barrier
        double2 ba[ST_PPT], bb[ST_PPT];
#pragma unroll
        for (int k = 0; k < ST_PPT; ++k) {
            if (MODE != 4 && act[k]) {
                int gi = z * P.S2 + (y0 + 2 * py[k]) * P.S1 + 2 * px[k];
                ba[k] = ldg2(P.b + gi);
                bb[k] = ldg2(P.b + gi + P.S1);
            }
        }
        if (!XF && z == z0) {
            mbar_wait(full + 0, 0);
            mbar_wait(full + 1, 0);
        }
        {
            int qq = q + 1;
            mbar_wait(full + (qq % NS), (uint32_t)((qq / NS) & 1));
            if (XF) {
                // (transforming one plane further ahead, so that the closing barrier of the previous step publishes
                // it and this barrier goes away, was measured SLOWER: 0.65 vs 0.57 ms at 512^3 — with four planes
                // resident the ring has no stage left in flight)
                transform(qq % NS, z + 1);
                if (z + 2 <= z1) load_aux(z + 2);
                __syncthreads();
            }
        }
        const double *sm = stage + (size_t)((q - 1) % NS) * P.SPAN;
        const double *sc = stage + (size_t)(q % NS) * P.SPAN;
        const double *sp = stage + (size_t)((q + 1) % NS) * P.SPAN;
#pragma unroll
        for (int k = 0; k < ST_PPT; ++k) {
            if (!act[k]) continue;
            const int oa = (2 * py[k] + 1) * P.S1 + 2 * px[k];
            const int ob = oa + P.S1;
            double2 va = lds2(sc + oa), vb = lds2(sc + ob);
            double2 vn = lds2(sc + oa - P.S1), vs = lds2(sc + ob + P.S1);
            double2 ma = lds2(sm + oa), mb = lds2(sm + ob);
            double2 pa = lds2(sp + oa), pb = lds2(sp + ob);
            double la = sc[oa - 1], ra = sc[oa + 2], lb = sc[ob - 1], rb = sc[ob + 2];
            double ax0 = P.d * va.x + P.c1 * (la + va.y) + P.cS * (vn.x + vb.x) + P.cP * (ma.x + pa.x);
            double ax1 = P.d * va.y + P.c1 * (va.x + ra) + P.cS * (vn.y + vb.y) + P.cP * (ma.y + pa.y);
            double ax2 = P.d * vb.x + P.c1 * (lb + vb.y) + P.cS * (va.x + vs.x) + P.cP * (mb.x + pb.x);
            double ax3 = P.d * vb.y + P.c1 * (vb.x + rb) + P.cS * (va.y + vs.y) + P.cP * (mb.y + pb.y);
            if (CLS) {
                // positional stencil classes: elements on the x/y grid boundaries add their correction taps
                const int ya = y0 + 2 * py[k];
                const int cya = (ya == 0) ? 0 : 3, cyb = (ya + 1 == P.NYg - 1) ? 6 : 3;      // 3*cy
                const int cxx = (px[k] == 0) ? 0 : 1, cxy = (px[k] == HX - 1) ? 2 : 1;       // cx of .x / .y
                auto corr = [&](int cls, int o) {
                    double a = 0.0;
                    for (int t = 0; t < P.cls.ntap[cls]; ++t) {
                        const int dz = P.cls.dz[cls][t];
                        const double *pl = dz == 0 ? sc : (dz > 0 ? sp : sm);
                        a += P.cls.coef[cls][t] * pl[o + P.cls.soff[cls][t]];
                    }
                    return a;
                };
                if (cya + cxx != 4) ax0 += corr(cya + cxx, oa);
                if (cya + cxy != 4) ax1 += corr(cya + cxy, oa + 1);
                if (cyb + cxx != 4) ax2 += corr(cyb + cxx, ob);
                if (cyb + cxy != 4) ax3 += corr(cyb + cxy, ob + 1);
            }
            const int gi = z * P.S2 + (y0 + 2 * py[k]) * P.S1 + 2 * px[k];
            if (MODE == 1 || MODE == 3 || MODE == 4) {
                double a = acc[k];
                if (MODE == 4) {
                    // the staged planes hold b and x = (omega/d) b everywhere: A x = (omega/d) (A b)
                    a += va.x - P.wod * ax0;
                    a += va.y - P.wod * ax1;
                    a += vb.x - P.wod * ax2;
                    a += vb.y - P.wod * ax3;
                } else {
                    a += ba[k].x - ax0;
                    a += ba[k].y - ax1;
                    a += bb[k].x - ax2;
                    a += bb[k].y - ax3;
                }
                if (z & 1) {
                    int Z = z >> 1;
                    P.rc[((long long)Z * P.cs1 + (y0 >> 1) + py[k]) * P.cs2 + px[k]] = P.w * a;
                    a = 0.0;
                }
                acc[k] = a;
                if (MODE == 3) {
                    *reinterpret_cast<double2 *>(P.xo + gi) = va;
                    *reinterpret_cast<double2 *>(P.xo + gi + P.S1) = vb;
                }
                if (MODE == 4) {
                    *reinterpret_cast<double2 *>(P.xo + gi) = make_double2(P.wod * va.x, P.wod * va.y);
                    *reinterpret_cast<double2 *>(P.xo + gi + P.S1) = make_double2(P.wod * vb.x, P.wod * vb.y);
                }
            } else {
                double2 oa2, ob2;
                oa2.x = va.x + P.wod * (ba[k].x - ax0);
                oa2.y = va.y + P.wod * (ba[k].y - ax1);
                ob2.x = vb.x + P.wod * (bb[k].x - ax2);
                ob2.y = vb.y + P.wod * (bb[k].y - ax3);
                if (P.colour >= 0) {
                    // two-colour half sweep: colour = (x + y + z) & 1; rows ya are even, x = 2px is even
                    bool even_match = (((z + P.zg0) & 1) == P.colour);    // colour of (ya, 2px)
                    if (even_match) {
                        oa2.y = va.y;
                        ob2.x = vb.x;
                    } else {
                        oa2.x = va.x;
                        ob2.y = vb.y;
                    }
                }
                *reinterpret_cast<double2 *>(P.xo + gi) = oa2;
                *reinterpret_cast<double2 *>(P.xo + gi + P.S1) = ob2;
            }
        }
        __syncthreads();        // everyone is done with plane z-1's slot
        if (tid == 0) {
            int p = z - 1 + NS;
            if (p <= z1) {
                int k = (q - 1 + NS) % NS;      // == slot of plane z-1
                // generic-proxy reads/writes of this slot are ordered before the async-proxy refill
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if constexpr (PULL) {
                    issue_span(k, p);
                } else {
                    mbar_expect_tx(full + k, span_bytes);
                    bulk_g2s(stage + (size_t)k * P.SPAN, P.xi + (long long)p * P.S2 + span0, span_bytes, full + k);
                }
            }
            if (!PULL && P.bpf > 0 && z + P.bpf < z1) {
                const long long bo = (long long)(z + P.bpf) * P.S2 + (long long)y0 * P.S1;
                bulk_prefetch_l2(P.b, bo, bo + (long long)P.TY * P.S1);
            }
            if (!PULL && MODE == 2 && P.epf > 0 && z + P.epf <= z1 && ((z + P.epf) & 1) == 0) {
                // coarse correction of the plane pair starting at z + epf: coarse rows y0/2 - 1 .. (y0 + TY)/2
                const long long nc = (long long)(P.NZg >> 1) * P.cs1 * P.cs2;
                const long long eo = ((long long)(((z + P.epf + P.zg0) >> 1) - P.cz0) * P.cs1 + (y0 >> 1) - 1) * P.cs2;
                bulk_prefetch_l2(P.e, max(eo, 0ll), min(eo + (long long)((P.TY >> 1) + 2) * P.cs2, nc));
            }
        }
    }
    if (PULL && brow >= 0 && tid == 0) st_pull_arrive(P, pullG, 2u * gridDim.x);
}

// ---------------------------------------------------------------- split-row variant (rows wider than 2*NT)
//
// Chunk geometry: TY rows x XW columns.  XW == S1 (full rows): the staged span of a plane is one
// contiguous piece of the flat vector, one bulk copy.  XW < S1 (rows >= 1024 wide): every staged row
// is its own bulk copy of XW+4 elements starting 2 elements left of the chunk — in the FLAT vector, so
// the left/right halo columns of the first/last chunk of a row are the neighbouring rows' ends,
// exactly the reference's no-boundary-break semantics.
template <int MODE, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256 ? 2 : 1)) k_st3x(const St3 P) {
    constexpr bool SPLIT = true;
    constexpr int TPT = ST_TPT + 1;   // transform pairs per thread (one more than the full-row kernel)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)P.NS * P.SPAN * sizeof(double));
    constexpr bool XF = (MODE == 2 || MODE == 3);
    const int NS = P.NS;
    const int tid = threadIdx.x;
    const int z0 = P.boundary ? (blockIdx.y == 0 ? 0 : P.zhi) : P.zlo + (int)blockIdx.y * P.ZL;
    const int z1 = P.boundary ? (blockIdx.y == 0 ? P.zlo : P.NZ) : min(z0 + P.ZL, P.zhi);
    const int xc = SPLIT ? (int)blockIdx.x % P.XC : 0, yc = SPLIT ? (int)blockIdx.x / P.XC : (int)blockIdx.x;
    const int y0 = yc * P.TY, x0 = SPLIT ? xc * P.XW : 0;
    const int PITCH = SPLIT ? P.PITCH : P.S1;
    constexpr int XH = SPLIT ? 2 : 0;
    const long long span0 = (long long)(y0 - 1) * P.S1 + x0 - XH;      // in-plane flat start of staged row 0
    const uint32_t span_bytes = (uint32_t)P.SPAN * 8u;

    auto issue_plane = [&](int slot, int p) {
        double *dst = stage + (size_t)slot * P.SPAN;
        const double *src = P.xi + (long long)p * P.S2 + span0;
        mbar_expect_tx(full + slot, span_bytes);
        if (!SPLIT) {
            bulk_g2s(dst, src, span_bytes, full + slot);
        } else {
            const uint32_t row_bytes = (uint32_t)PITCH * 8u;
            for (int q = 0; q < P.TY + 2; ++q)
                bulk_g2s(dst + q * PITCH, src + (long long)q * P.S1, row_bytes, full + slot);
        }
    };

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < NS; ++k) {
            int p = z0 - 1 + k;
            if (p > z1) break;
            issue_plane(k, p);
        }
    }

    // fixed 2x2 patch assignment
    const int HX = (SPLIT ? P.XW : P.S1) >> 1;
    int px[ST_PPT], py[ST_PPT];
    bool act[ST_PPT];
#pragma unroll
    for (int k = 0; k < ST_PPT; ++k) {
        int p = tid + k * NT;
        act[k] = p < P.NP;
        p = act[k] ? p : 0;
        py[k] = p / HX;
        px[k] = p - py[k] * HX;
    }
    double acc[ST_PPT];
#pragma unroll
    for (int k = 0; k < ST_PPT; ++k) acc[k] = 0.0;

    // fixed pair assignment of the in-smem transform: staged row and (even) column inside the staged row
    int tq[TPT], tc[TPT];
    bool tact[TPT];
    double aux[TPT];        // MODE 2: e of the pair's coarse cell
    unsigned auxm[TPT];     // MODE 3: exception mask word of the pair
    if (XF) {
        const int HP = PITCH >> 1;
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
            int t = tid + k * NT;
            tact[k] = 2 * t < P.SPAN;
            t = tact[k] ? t : 0;
            tq[k] = t / HP;
            tc[k] = 2 * (t - tq[k] * HP);
            aux[k] = 0.0;
            auxm[k] = 0u;
        }
    }
    // flat row (may leave [0,NY) by one) and column of pair k: halo columns of split rows wrap into the
    // neighbouring flat row
    auto pair_rc = [&](int k, int &row, int &col) {
        row = y0 - 1 + tq[k];
        col = tc[k];
        if (SPLIT) {
            int xg = x0 - XH + tc[k];
            int dq = xg < 0 ? -1 : (xg >= P.S1 ? 1 : 0);
            row += dq;
            col = xg - dq * P.S1;
        }
    };
    // aux of local plane p for pair k
    auto load_aux = [&](int p) {
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
            if (!tact[k]) continue;
            int yy, col;
            pair_rc(k, yy, col);
            if (MODE == 2) {
                int zz = p + P.zg0;
                if (yy < 0) {              // flat-index wrap into the neighbouring plane
                    yy += P.NYg;
                    zz -= 1;
                } else if (yy >= P.NYg) {
                    yy -= P.NYg;
                    zz += 1;
                }
                double v = 0.0;
                if (zz >= 0 && zz < P.NZg)
                    v = __ldg(P.e + ((long long)((zz >> 1) - P.cz0) * P.cs1 + (yy >> 1)) * P.cs2 + (col >> 1));
                aux[k] = v;
            } else if (MODE == 3) {
                unsigned m = 0u;
                if (P.has_exc) {
                    long long gl = (long long)p * P.S2 + (long long)yy * P.S1 + col;
                    if (gl >= 0 && gl < P.nloc) m = __ldg(P.exc.mask + (gl >> 5));
                }
                auxm[k] = m;
            }
        }
    };
    auto transform = [&](int slot, int p) {
        double *sp_ = stage + (size_t)slot * P.SPAN;
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
            if (!tact[k]) continue;
            int o = tq[k] * PITCH + tc[k];
            double2 v = lds2(sp_ + o);
            if (MODE == 2) {
                v.x += P.w * aux[k];
                v.y += P.w * aux[k];
            } else {
                double s0 = P.wod, s1 = P.wod;
                if (P.has_exc) {
                    unsigned m = auxm[k];
                    int yy, col;
                    pair_rc(k, yy, col);
                    long long gl = (long long)p * P.S2 + (long long)yy * P.S1 + col;
                    int bit = (int)(gl & 31);
                    if ((m >> bit) & 3u) {
                        int base = __ldg(P.exc.wpre + (gl >> 5));
                        if ((m >> bit) & 1u)
                            s0 = P.omega / __ldg(P.exc.diag + base + __popc(m & ((1u << bit) - 1u)));
                        if ((m >> (bit + 1)) & 1u)
                            s1 = P.omega / __ldg(P.exc.diag + base + __popc(m & ((2u << bit) - 1u)));
                    }
                }
                v.x *= s0;
                v.y *= s1;
            }
            sts2(sp_ + o, v);
        }
    };

    if (XF) {
        // planes z0-1 and z0 are transformed up front, z0+1 inside the loop
        load_aux(z0 - 1);
        mbar_wait(full + 0, 0);
        transform(0, z0 - 1);
        load_aux(z0);
        mbar_wait(full + 1, 0);
        transform(1, z0);
        load_aux(z0 + 1);
    }

    for (int z = z0; z < z1; ++z) {
        const int q = z - (z0 - 1);                // ring position of plane z (plane z0-1 is 0)
        // b for this plane: issue the loads before blocking on the barrier
        double2 ba[ST_PPT], bb[ST_PPT];
#pragma unroll
        for (int k = 0; k < ST_PPT; ++k) {
            if (act[k]) {
                int gi = z * P.S2 + (y0 + 2 * py[k]) * P.S1 + x0 + 2 * px[k];
                ba[k] = ldg2(P.b + gi);
                bb[k] = ldg2(P.b + gi + P.S1);
            }
        }
        if (!XF && z == z0) {
            mbar_wait(full + 0, 0);
            mbar_wait(full + 1, 0);
        }
        {
            int qq = q + 1;
            mbar_wait(full + (qq % NS), (uint32_t)((qq / NS) & 1));
            if (XF) {
                transform(qq % NS, z + 1);
                if (z + 2 <= z1) load_aux(z + 2);
                __syncthreads();
            }
        }
        const double *sm = stage + (size_t)((q - 1) % NS) * P.SPAN;
        const double *sc = stage + (size_t)(q % NS) * P.SPAN;
        const double *sp = stage + (size_t)((q + 1) % NS) * P.SPAN;
#pragma unroll
        for (int k = 0; k < ST_PPT; ++k) {
            if (!act[k]) continue;
            const int oa = (2 * py[k] + 1) * PITCH + XH + 2 * px[k];
            const int ob = oa + PITCH;
            double2 va = lds2(sc + oa), vb = lds2(sc + ob);
            double2 vn = lds2(sc + oa - PITCH), vs = lds2(sc + ob + PITCH);
            double2 ma = lds2(sm + oa), mb = lds2(sm + ob);
            double2 pa = lds2(sp + oa), pb = lds2(sp + ob);
            double la = sc[oa - 1], ra = sc[oa + 2], lb = sc[ob - 1], rb = sc[ob + 2];
            double ax0 = P.d * va.x + P.c1 * (la + va.y) + P.cS * (vn.x + vb.x) + P.cP * (ma.x + pa.x);
            double ax1 = P.d * va.y + P.c1 * (va.x + ra) + P.cS * (vn.y + vb.y) + P.cP * (ma.y + pa.y);
            double ax2 = P.d * vb.x + P.c1 * (lb + vb.y) + P.cS * (va.x + vs.x) + P.cP * (mb.x + pb.x);
            double ax3 = P.d * vb.y + P.c1 * (vb.x + rb) + P.cS * (va.y + vs.y) + P.cP * (mb.y + pb.y);
            const int gi = z * P.S2 + (y0 + 2 * py[k]) * P.S1 + x0 + 2 * px[k];
            if (MODE == 1 || MODE == 3) {
                double a = acc[k];
                a += ba[k].x - ax0;
                a += ba[k].y - ax1;
                a += bb[k].x - ax2;
                a += bb[k].y - ax3;
                if (z & 1) {
                    int Z = z >> 1;
                    P.rc[((long long)Z * P.cs1 + (y0 >> 1) + py[k]) * P.cs2 + (x0 >> 1) + px[k]] = P.w * a;
                    a = 0.0;
                }
                acc[k] = a;
                if (MODE == 3) {
                    *reinterpret_cast<double2 *>(P.xo + gi) = va;
                    *reinterpret_cast<double2 *>(P.xo + gi + P.S1) = vb;
                }
            } else {
                double2 oa2, ob2;
                oa2.x = va.x + P.wod * (ba[k].x - ax0);
                oa2.y = va.y + P.wod * (ba[k].y - ax1);
                ob2.x = vb.x + P.wod * (bb[k].x - ax2);
                ob2.y = vb.y + P.wod * (bb[k].y - ax3);
                if (P.colour >= 0) {
                    // two-colour half sweep: colour = (x + y + z) & 1; rows ya and columns x0 + 2px are even
                    bool even_match = (((z + P.zg0) & 1) == P.colour);    // colour of (ya, x0 + 2px)
                    if (even_match) {
                        oa2.y = va.y;
                        ob2.x = vb.x;
                    } else {
                        oa2.x = va.x;
                        ob2.y = vb.y;
                    }
                }
                *reinterpret_cast<double2 *>(P.xo + gi) = oa2;
                *reinterpret_cast<double2 *>(P.xo + gi + P.S1) = ob2;
            }
        }
        __syncthreads();        // everyone is done with plane z-1's slot
        if (tid == 0) {
            int p = z - 1 + NS;
            if (p <= z1) {
                // generic-proxy reads/writes of this slot are ordered before the async-proxy refill
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue_plane((q - 1 + NS) % NS, p);      // == slot of plane z-1
            }
        }
    }
}

// ================================================================ 2-D structured path
//
//   A = d I + c1 (S^1 + S^-1) + cN (S^N + S^-N) + cD (S^(N+1) + S^-(N+1))       (flat index, no row breaks:
//   openmg/operators.py:221-241 has cN = 0; its Galerkin images have all three pairs)
// with the closed-form 2x2 restriction.  Same machinery as the 3-D kernel, one dimension down: a CTA owns an
// x-chunk of XW columns of a y-segment and marches over rows; every row chunk (+2 halo columns each side, taken
// from the FLAT vector so that row ends continue into the neighbouring rows) is one TMA bulk copy into an
// mbarrier ring; rows y-1, y, y+1 are resident.  Each thread owns x-pairs; the restriction accumulates a pair
// over two consecutive rows.
#define ST2_PP 4           // x-pairs per thread per row (max)
#define ST2_TPT 5          // staged pairs per thread in the in-smem transform (max)
#define ST2_NT 256

struct St2 {
    const double *xi;      // staged vector (MODE 3: b)
    const double *b;
    double *xo;
    double *rc;
    const double *e;
    int N, NY;             // row length, local rows
    int XW, XC, PITCH;     // chunk width, chunks per row, staged row pitch XW + 4
    int NS, YL;            // ring stages, rows per y-segment (even)
    int cs;                // coarse row length N/2
    int yg0, NYg;          // global index of local row 0, global rows
    int colour, cflat;     // -1: all rows; else the colour relaxed; cflat: colour = x & 1, else (x + y) & 1
    int oned;              // 1-D problem viewed as rows of N: restriction pairs x only, coarse row length N/2 per row
    int use_cls;           // rows of the first / last grid column get their correction taps in-kernel (no fix-up)
    int bpf;               // rows by which thread 0 prefetches b (MODE 2: and e) into L2 ahead of their loads (0: off)
    double c2l[3], c2r[3]; // deltas of the taps -1, -(N+1), +(N-1)  /  +1, +(N+1), -(N-1)
    double d, c1, cN, cD, wod, w;
};

template <int MODE, bool CLS>
__global__ void __launch_bounds__(ST2_NT, 2) k_st2(const St2 P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)P.NS * P.PITCH * sizeof(double));
    constexpr bool XF = (MODE == 2 || MODE == 3);
    constexpr int NT = ST2_NT;
    const int NS = P.NS, PITCH = P.PITCH;
    const int tid = threadIdx.x;
    const int x0 = (int)blockIdx.x * P.XW;
    const int y0 = (int)blockIdx.y * P.YL;
    const int y1 = min(y0 + P.YL, P.NY);
    const uint32_t row_bytes = (uint32_t)PITCH * 8u;

    auto issue_row = [&](int slot, int y) {
        mbar_expect_tx(full + slot, row_bytes);
        bulk_g2s(stage + (size_t)slot * PITCH, P.xi + (long long)y * P.N + x0 - 2, row_bytes, full + slot);
    };
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < NS; ++k) {
            int y = y0 - 1 + k;
            if (y > y1) break;
            issue_row(k, y);
        }
    }
    const int HXW = P.XW >> 1;       // pairs per row chunk
    const int HP = PITCH >> 1;       // staged pairs per row
    double acc[ST2_PP];
#pragma unroll
    for (int k = 0; k < ST2_PP; ++k) acc[k] = 0.0;
    double aux[ST2_TPT];
#pragma unroll
    for (int k = 0; k < ST2_TPT; ++k) aux[k] = 0.0;

    // e of the coarse cell of staged pair k of local row y (MODE 2); halo columns wrap into neighbouring flat rows
    auto load_aux = [&](int y) {
        if constexpr (MODE == 2) {
#pragma unroll
        for (int k = 0; k < ST2_TPT; ++k) {
            int t = tid + k * NT;
            if (t >= HP) continue;
            int xg = x0 - 2 + 2 * t;
            int dq = xg < 0 ? -1 : (xg >= P.N ? 1 : 0);
            int row = y + P.yg0 + dq, col = xg - dq * P.N;
            double v = 0.0;
            if (row >= 0 && row < P.NYg) {
                long long crow = P.oned ? (long long)(row - P.yg0) : (long long)((row >> 1) - (P.yg0 >> 1));
                v = __ldg(P.e + crow * P.cs + (col >> 1));
            }
            aux[k] = v;
        }
        }
    };
    auto transform = [&](int slot) {
        double *sp_ = stage + (size_t)slot * PITCH;
#pragma unroll
        for (int k = 0; k < ST2_TPT; ++k) {
            int t = tid + k * NT;
            if (t >= HP) continue;
            double2 v = lds2(sp_ + 2 * t);
            if (MODE == 2) {
                v.x += P.w * aux[k];
                v.y += P.w * aux[k];
            } else {
                v.x *= P.wod;
                v.y *= P.wod;
            }
            sts2(sp_ + 2 * t, v);
        }
    };
    if (XF) {
        load_aux(y0 - 1);
        mbar_wait(full + 0, 0);
        transform(0);
        load_aux(y0);
        mbar_wait(full + 1, 0);
        transform(1);
        load_aux(y0 + 1);
    }

    for (int y = y0; y < y1; ++y) {
        const int q = y - (y0 - 1);
        double2 bv[ST2_PP];
#pragma unroll
        for (int k = 0; k < ST2_PP; ++k) {
            int pi = tid + k * NT;
            if (pi < HXW) bv[k] = ldg2(P.b + (long long)y * P.N + x0 + 2 * pi);
        }
        if (!XF && y == y0) {
            mbar_wait(full + 0, 0);
            mbar_wait(full + 1, 0);
        }
        {
            int qq = q + 1;
            mbar_wait(full + (qq % NS), (uint32_t)((qq / NS) & 1));
            if (XF) {
                transform(qq % NS);
                if (y + 2 <= y1) load_aux(y + 2);
                __syncthreads();
            }
        }
        const double *sm = stage + (size_t)((q - 1) % NS) * PITCH;
        const double *sc = stage + (size_t)(q % NS) * PITCH;
        const double *sp = stage + (size_t)((q + 1) % NS) * PITCH;
#pragma unroll
        for (int k = 0; k < ST2_PP; ++k) {
            int pi = tid + k * NT;
            if (pi >= HXW) continue;
            const int o = 2 + 2 * pi;
            double2 c = lds2(sc + o), m = lds2(sm + o), p = lds2(sp + o);
            double l = sc[o - 1], r = sc[o + 2], ml = sm[o - 1], pr = sp[o + 2];
            double ax0 = P.d * c.x + P.c1 * (l + c.y) + P.cN * (m.x + p.x) + P.cD * (ml + p.y);
            double ax1 = P.d * c.y + P.c1 * (c.x + r) + P.cN * (m.y + p.y) + P.cD * (m.x + pr);
            if (CLS) {
                if (x0 + 2 * pi == 0) ax0 += P.c2l[0] * l + P.c2l[1] * ml + P.c2l[2] * sp[o - 1];
                if (x0 + 2 * pi == P.N - 2) ax1 += P.c2r[0] * r + P.c2r[1] * pr + P.c2r[2] * sm[o + 2];
            }
            const long long gi = (long long)y * P.N + x0 + 2 * pi;
            if (MODE == 1 || MODE == 3) {
                double a = acc[k];
                a += bv[k].x - ax0;
                a += bv[k].y - ax1;
                if (P.oned) {
                    P.rc[(long long)y * P.cs + (x0 >> 1) + pi] = P.w * a;
                    a = 0.0;
                } else if (y & 1) {
                    P.rc[(long long)(y >> 1) * P.cs + (x0 >> 1) + pi] = P.w * a;
                    a = 0.0;
                }
                acc[k] = a;
                if (MODE == 3) *reinterpret_cast<double2 *>(P.xo + gi) = c;
            } else {
                double2 o2;
                o2.x = c.x + P.wod * (bv[k].x - ax0);
                o2.y = c.y + P.wod * (bv[k].y - ax1);
                if (P.colour >= 0) {
                    // x = x0 + 2 pi is even: its colour is 0 (flat parity) or the parity of the global row
                    bool even_match = P.cflat ? (P.colour == 0) : ((((y + P.yg0) & 1)) == P.colour);
                    if (even_match)
                        o2.y = c.y;
                    else
                        o2.x = c.x;
                }
                *reinterpret_cast<double2 *>(P.xo + gi) = o2;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int yn = y - 1 + NS;
            if (yn <= y1) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue_row((q - 1 + NS) % NS, yn);
            }
            const int yb = y + P.bpf;
            if (P.bpf > 0 && yb < y1) {
                const long long bo = (long long)yb * P.N + x0;
                bulk_prefetch_l2(P.b, bo, bo + P.XW);
                if (MODE == 2 && (P.oned || (yb & 1) == 0)) {
                    const long long eo = (P.oned ? (long long)yb : (long long)(yb >> 1)) * P.cs + (x0 >> 1);
                    bulk_prefetch_l2(P.e, eo, eo + (P.XW >> 1));
                }
            }
        }
    }
}


// ---------------------------------------------------------------- 2-D fused two-colour sweep
//
// Both colour half-sweeps of one two-colour Gauss-Seidel sweep (oracle.rbgs: colour c0 first, then 1 - c0,
// same-colour couplings lagged) in ONE pass over x: marching over rows, pass A relaxes the colour-c0 points of
// row r from the staged raw rows r-1, r, r+1 into a second shared-memory ring ("mid": c0 points new, the others
// old); one row behind, pass B relaxes the other colour of row r-1 from mid rows r-2, r-1, r and writes the row
// out.  Pass B needs mid one column / one row beyond the chunk, so pass A also runs on one halo pair per side and
// one halo row per segment end (recomputed, bit-identical to the owner's values); raw rows carry 4 halo columns.
// Points outside the global vector stay zero (the pads), as in the reference's truncated band.
// MODE 0: x = xi.   MODE 2: x = xi + R^T e (raw rows transformed in shared memory when they land).
// Levels with exception rows qualify only when the kernel corrects them itself (CLS: the first / last grid column
// of 2-D Galerkin levels): a fix-up kernel would have to run between the two passes.
#define ST2RB_PP 5         // pairs per thread per row, halo pairs included
#define ST2RB_TPT 5        // staged raw pairs per thread in the transform

template <int MODE, bool CLS, bool HASN>
__global__ void __launch_bounds__(ST2_NT, 2) k_st2rb(const St2 P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NT = ST2_NT;
    const int NS = P.NS, RP = P.XW + 8, MP = P.XW + 4;
    double *raw = reinterpret_cast<double *>(smem_raw);
    double *mid = raw + (size_t)NS * RP;
    uint64_t *full = reinterpret_cast<uint64_t *>(mid + 3 * MP);
    const int tid = threadIdx.x;
    const int x0 = (int)blockIdx.x * P.XW;
    const int y0 = (int)blockIdx.y * P.YL;
    const int y1 = min(y0 + P.YL, P.NY);
    // Pass A also runs on the row slots -1 and NY: their wrapped halo pairs are the first / last pair of the
    // vector, which the column-correction taps +-(N-1) of the corner rows read.  Everything else there is pad.
    const int rlo = y0 - 2, rhi = y1 + 1;                           // staged raw rows
    const int first = y0 - 1;                                       // first row pass A runs on
    const long long ntot = (long long)P.NY * P.N;
    const uint32_t row_bytes = (uint32_t)RP * 8u;
    const int c0 = P.colour;

    auto issue_row = [&](int r) {
        int slot = (r - rlo) % NS;
        mbar_expect_tx(full + slot, row_bytes);
        bulk_g2s(raw + (size_t)slot * RP, P.xi + (long long)r * P.N + x0 - 4, row_bytes, full + slot);
    };
    auto wait_row = [&](int r) {
        int q = r - rlo;
        mbar_wait(full + (q % NS), (uint32_t)((q / NS) & 1));
    };
    auto raw_row = [&](int r) { return raw + (size_t)((r - rlo) % NS) * RP; };
    auto mid_row = [&](int r) { return mid + (size_t)((r - (y0 - 1)) % 3) * MP; };

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int r = rlo; r < rlo + NS && r <= rhi; ++r) issue_row(r);

    const int HXW = P.XW >> 1;
    const int NPA = HXW + 2;         // pass-A pairs: pair p covers columns x0 - 2 + 2p, +1
    const int HRP = RP >> 1;
    const int MPH = MP >> 1;         // pairs per mid row

    double bold[ST2RB_PP], bnew[ST2RB_PP];      // b of the element pass B relaxes: row it-1 / row it
#pragma unroll
    for (int k = 0; k < ST2RB_PP; ++k) bold[k] = bnew[k] = 0.0;

    // raw row r += w e[agg]: staged pair t covers columns x0 - 4 + 2t, +1 (halo columns wrap into the flat neighbours)
    auto transform = [&](int r) {
        if constexpr (MODE == 2) {
            double *rr_ = raw_row(r);
#pragma unroll
            for (int k = 0; k < ST2RB_TPT; ++k) {
                int t = tid + k * NT;
                if (t >= HRP) continue;
                int xg = x0 - 4 + 2 * t;
                int dq = xg < 0 ? -1 : (xg >= P.N ? 1 : 0);
                int row = r + dq, col = xg - dq * P.N;
                if (row < 0 || row >= P.NY) continue;
                long long crow = P.oned ? (long long)row : (long long)(row >> 1);
                double v = P.w * __ldg(P.e + crow * P.cs + (col >> 1));
                double2 x = lds2(rr_ + 2 * t);
                x.x += v;
                x.y += v;
                sts2(rr_ + 2 * t, x);
            }
        }
    };
    if (MODE == 2) {
        wait_row(first - 1);
        transform(first - 1);
        wait_row(first);
        transform(first);
    }

    // (fetching b / e one row ahead in registers was measured and was slower: 0.38 vs 0.34 ms on 8192^2)
    for (int it = y0 - 1; it <= y1; ++it) {
        double *mw = mid_row(it);
        {
            double2 bv[ST2RB_PP];
#pragma unroll
            for (int k = 0; k < ST2RB_PP; ++k) {
                int p = tid + k * NT;
                if (p < NPA) bv[k] = ldg2(P.b + (long long)it * P.N + x0 - 2 + 2 * p);
            }
            if (MODE == 2) {
                wait_row(it + 1);
                transform(it + 1);
                __syncthreads();
            } else {
                if (it == first) {
                    wait_row(it - 1);
                    wait_row(it);
                }
                wait_row(it + 1);
            }
            const double *sm = raw_row(it - 1), *sc = raw_row(it), *sp = raw_row(it + 1);
#pragma unroll
            for (int k = 0; k < ST2RB_PP; ++k) {
                int p = tid + k * NT;
                if (p >= NPA) continue;
                const int o = 2 + 2 * p;
                const int xg = x0 - 2 + 2 * p;
                const long long gi = (long long)it * P.N + xg;
                double2 c = lds2(sc + o);
                // does the even element of the pair have colour c0?  (wrapped halo columns sit one grid row off)
                const int flip = (xg < 0 || xg >= P.N) ? 1 : 0;
                const bool ex = P.cflat ? (c0 == 0) : (((it + flip) & 1) == c0);
                if (gi >= 0 && gi < ntot) {
                    // only the taps with a non-zero coefficient are read (level 0 has no +-N pair: HASN = false)
                    if (ex) {
                        const double l = sc[o - 1], ml = sm[o - 1];
                        double ax0 = P.d * c.x + P.c1 * (l + c.y);
                        if (HASN) {
                            double2 q = lds2(sp + o);
                            ax0 += P.cN * (sm[o] + q.x) + P.cD * (ml + q.y);
                        } else {
                            ax0 += P.cD * (ml + sp[o + 1]);
                        }
                        if (CLS && (xg == 0 || xg == P.N))         // first grid column (also as the wrapped right halo pair)
                            ax0 += P.c2l[0] * l + P.c2l[1] * ml + P.c2l[2] * sp[o - 1];
                        c.x += P.wod * (bv[k].x - ax0);
                    } else {
                        const double r = sc[o + 2], pr = sp[o + 2];
                        double ax1 = P.d * c.y + P.c1 * (c.x + r);
                        if (HASN) {
                            double2 m = lds2(sm + o);
                            ax1 += P.cN * (m.y + sp[o + 1]) + P.cD * (m.x + pr);
                        } else {
                            ax1 += P.cD * (sm[o] + pr);
                        }
                        if (CLS && (xg == P.N - 2 || xg == -2))    // last grid column (also as the wrapped left halo pair)
                            ax1 += P.c2r[0] * r + P.c2r[1] * pr + P.c2r[2] * sm[o + 2];
                        c.y += P.wod * (bv[k].y - ax1);
                    }
                }
                bnew[k] = ex ? bv[k].y : bv[k].x;
                // mid rows are stored de-interleaved (even elements, then odd elements): pass B reads them conflict-free
                mw[p] = c.x;
                mw[MPH + p] = c.y;
            }
        }
        __syncthreads();        // mid row `it` complete
        const int r = it - 1;
        if (r >= y0) {
            const double *em = mid_row(r - 1), *ec = mid_row(r), *ep = mid_row(r + 1);     // even elements of the pairs
            const double *om = em + MPH, *oc = ec + MPH, *op = ep + MPH;                   // odd elements
            const bool ex = P.cflat ? (c0 == 0) : ((r & 1) == c0);
#pragma unroll
            for (int k = 0; k < ST2RB_PP; ++k) {
                int p = tid + k * NT;
                if (p < 1 || p > HXW) continue;
                double2 c = make_double2(ec[p], oc[p]);
                const int xg = x0 - 2 + 2 * p;
                if (ex) {       // the odd element has the second colour
                    const double rr = ec[p + 1], pr = ep[p + 1];
                    double ax1 = P.d * c.y + P.c1 * (c.x + rr) + P.cD * (em[p] + pr);
                    if (HASN) ax1 += P.cN * (om[p] + op[p]);
                    if (CLS && xg == P.N - 2) ax1 += P.c2r[0] * rr + P.c2r[1] * pr + P.c2r[2] * em[p + 1];
                    c.y += P.wod * (bold[k] - ax1);
                } else {
                    const double l = oc[p - 1], ml = om[p - 1];
                    double ax0 = P.d * c.x + P.c1 * (l + c.y) + P.cD * (ml + op[p]);
                    if (HASN) ax0 += P.cN * (em[p] + ep[p]);
                    if (CLS && xg == 0) ax0 += P.c2l[0] * l + P.c2l[1] * ml + P.c2l[2] * op[p - 1];
                    c.x += P.wod * (bold[k] - ax0);
                }
                *reinterpret_cast<double2 *>(P.xo + (long long)r * P.N + x0 + 2 * (p - 1)) = c;
            }
        }
#pragma unroll
        for (int k = 0; k < ST2RB_PP; ++k) bold[k] = bnew[k];
        __syncthreads();        // raw row it-1 and mid row it-2 are free
        if (tid == 0) {
            int rn = it - 1 + NS;
            if (rn <= rhi) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue_row(rn);
            }
            const int yb = it + P.bpf;
            if (P.bpf > 0 && yb >= 0 && yb <= y1 && yb < P.NY) {
                const long long bo = (long long)yb * P.N + x0;
                bulk_prefetch_l2(P.b, bo, bo + P.XW);
                if (MODE == 2 && yb + 1 < P.NY && (P.oned || ((yb + 1) & 1) == 0)) {      // e of raw row yb + 1
                    const long long eo = (P.oned ? (long long)(yb + 1) : (long long)((yb + 1) >> 1)) * P.cs + (x0 >> 1);
                    bulk_prefetch_l2(P.e, eo, eo + (P.XW >> 1));
                }
            }
        }
    }
}

// ================================================================ 2-D / 1-D Jacobi sweep + residual + restriction in one pass
//
// k_jr3 one dimension down, on the skeleton of k_st2rb (x-chunks of y-segments, marching over rows; one TMA bulk copy
// per raw row with 4 halo columns per side from the FLAT vector, ring of NS stages; 3 de-interleaved mid rows):
//   pass A(it)   Jacobi on every point of row `it` from the raw rows it-1, it, it+1 -> mid row `it` (and, for the rows
//                and columns the CTA owns, the new iterate in global memory); it also runs on one halo pair per side
//                and on the rows y0-1 and y1 (recomputed, never stored)
//   pass B(it-1) residual of row it-1 from the mid rows it-2, it-1, it, summed over the x-pair and (2-D) the row pair
//                of the aggregate -> one coarse value
// x and b are read once: 24 n + 8 n_c bytes instead of 24 n + (16 n + 8 n_c).  Pure-band, unsharded levels (level 0 of
// the 2-D / 1-D Poisson hierarchies: openmg/operators.py:191-241); 1-D vectors are viewed as rows of N.
#define JR2_PP 5           // pairs per thread per row, halo pairs included

template <bool HASN>
__global__ void __launch_bounds__(ST2_NT, 2) k_jr2(const St2 P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NT = ST2_NT;
    const int NS = P.NS, RP = P.XW + 8, MP = P.XW + 4;
    double *raw = reinterpret_cast<double *>(smem_raw);
    double *mid = raw + (size_t)NS * RP;
    uint64_t *full = reinterpret_cast<uint64_t *>(mid + 3 * MP);
    const int tid = threadIdx.x;
    const int x0 = (int)blockIdx.x * P.XW;
    const int y0 = (int)blockIdx.y * P.YL;
    const int y1 = min(y0 + P.YL, P.NY);
    const int rlo = y0 - 2, rhi = y1 + 1;                           // staged raw rows
    const int first = y0 - 1;                                       // first row pass A runs on
    const long long ntot = (long long)P.NY * P.N;
    const uint32_t row_bytes = (uint32_t)RP * 8u;

    auto issue_row = [&](int r) {
        int slot = (r - rlo) % NS;
        mbar_expect_tx(full + slot, row_bytes);
        bulk_g2s(raw + (size_t)slot * RP, P.xi + (long long)r * P.N + x0 - 4, row_bytes, full + slot);
    };
    auto wait_row = [&](int r) {
        int q = r - rlo;
        mbar_wait(full + (q % NS), (uint32_t)((q / NS) & 1));
    };
    auto raw_row = [&](int r) { return raw + (size_t)((r - rlo) % NS) * RP; };
    auto mid_row = [&](int r) { return mid + (size_t)((r - (y0 - 1)) % 3) * MP; };

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int r = rlo; r < rlo + NS && r <= rhi; ++r) issue_row(r);

    const int HXW = P.XW >> 1;
    const int NPA = HXW + 2;         // pass-A pairs: pair p covers columns x0 - 2 + 2p, +1
    const int MPH = MP >> 1;         // pairs per mid row

    double2 bold[JR2_PP];            // b of row it-1
    double acc[JR2_PP];              // residual sum of the open aggregate
#pragma unroll
    for (int k = 0; k < JR2_PP; ++k) {
        bold[k] = make_double2(0.0, 0.0);
        acc[k] = 0.0;
    }

    for (int it = y0 - 1; it <= y1; ++it) {
        double *mw = mid_row(it);
        double2 bv[JR2_PP];
#pragma unroll
        for (int k = 0; k < JR2_PP; ++k) {
            int p = tid + k * NT;
            bv[k] = make_double2(0.0, 0.0);
            if (p < NPA) bv[k] = ldg2(P.b + (long long)it * P.N + x0 - 2 + 2 * p);
        }
        if (it == first) {
            wait_row(it - 1);
            wait_row(it);
        }
        wait_row(it + 1);
        {
            // ---- pass A
            const double *sm = raw_row(it - 1), *sc = raw_row(it), *sp = raw_row(it + 1);
            const bool own_row = it >= y0 && it < y1;
#pragma unroll
            for (int k = 0; k < JR2_PP; ++k) {
                int p = tid + k * NT;
                if (p >= NPA) continue;
                const int o = 2 + 2 * p;
                const int xg = x0 - 2 + 2 * p;
                const long long gi = (long long)it * P.N + xg;
                double2 c = lds2(sc + o);
                if (gi >= 0 && gi < ntot) {         // points outside the vector are never relaxed: they stay zero (pads)
                    const double l = sc[o - 1], r = sc[o + 2];
                    const double2 q = lds2(sp + o);
                    const double ml = sm[o - 1], m0 = sm[o], pr = sp[o + 2];
                    double ax0 = P.d * c.x + P.c1 * (l + c.y) + P.cD * (ml + q.y);
                    double ax1 = P.d * c.y + P.c1 * (c.x + r) + P.cD * (m0 + pr);
                    if (HASN) {
                        ax0 += P.cN * (m0 + q.x);
                        ax1 += P.cN * (sm[o + 1] + q.y);
                    }
                    c.x += P.wod * (bv[k].x - ax0);
                    c.y += P.wod * (bv[k].y - ax1);
                    if (own_row && p >= 1 && p <= HXW) *reinterpret_cast<double2 *>(P.xo + gi) = c;
                }
                // mid rows are stored de-interleaved (even elements, then odd elements): pass B reads them conflict-free
                mw[p] = c.x;
                mw[MPH + p] = c.y;
            }
        }
        __syncthreads();        // mid row `it` complete
        const int r = it - 1;
        if (r >= y0) {
            // ---- pass B
            const double *em = mid_row(r - 1), *ec = mid_row(r), *ep = mid_row(r + 1);     // even elements of the pairs
            const double *om = em + MPH, *oc = ec + MPH, *op = ep + MPH;                   // odd elements
            const bool close_agg = P.oned || (r & 1);
            double *rcrow = P.rc + (P.oned ? (long long)r : (long long)(r >> 1)) * P.cs + (x0 >> 1) - 1;
#pragma unroll
            for (int k = 0; k < JR2_PP; ++k) {
                int p = tid + k * NT;
                if (p < 1 || p > HXW) continue;
                const double cx = ec[p], cy = oc[p];
                double ax0 = P.d * cx + P.c1 * (oc[p - 1] + cy) + P.cD * (om[p - 1] + op[p]);
                double ax1 = P.d * cy + P.c1 * (cx + ec[p + 1]) + P.cD * (em[p] + ep[p + 1]);
                if (HASN) {
                    ax0 += P.cN * (em[p] + ep[p]);
                    ax1 += P.cN * (om[p] + op[p]);
                }
                double a = acc[k] + ((bold[k].x - ax0) + (bold[k].y - ax1));
                if (close_agg) {
                    rcrow[p] = P.w * a;
                    a = 0.0;
                }
                acc[k] = a;
            }
        }
#pragma unroll
        for (int k = 0; k < JR2_PP; ++k) bold[k] = bv[k];
        __syncthreads();        // raw row it-1 and mid row it-2 are free
        if (tid == 0) {
            int rn = it - 1 + NS;
            if (rn <= rhi) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue_row(rn);
            }
            const int yb = it + P.bpf;
            if (P.bpf > 0 && yb >= 0 && yb <= y1 && yb < P.NY) {
                const long long bo = (long long)yb * P.N + x0;
                bulk_prefetch_l2(P.b, bo, bo + P.XW);
            }
        }
    }
}

// ---------------------------------------------------------------- fix-ups for exception rows

// xo_i for exception rows, MODE 0 (jacobi) and MODE 2 (prolong + jacobi, y = x + R^T e on the fly).
// 8 lanes per exception row: entries are loaded and gathered in parallel, fixed-order shuffle sum.
template <int MODE>
__global__ void __launch_bounds__(OMG_TPB) k_fix_rows(const int *__restrict__ rows, int s0, int cnt, ExcOp E, RegR R,
                                                      int nglob, const double *__restrict__ xi,
                                                      const double *__restrict__ e, const double *__restrict__ b,
                                                      double *__restrict__ xo, double omega, ColourRule cr,
                                                      int colour) {
    // all vector pointers are indexable by GLOBAL row / coarse row
    const int crow0 = 0, frow0 = 0;
    int gt = blockIdx.x * OMG_TPB + threadIdx.x;
    int s = s0 + (gt >> 3), k = gt & 7;
    bool valid = (gt >> 3) < cnt;
    double acc = 0.0;
    int i = 0;
    if (valid) {
        i = __ldg(rows + s);
        if (colour >= 0 && colour_of(cr, i) != colour) valid = false;   // other colour: the main kernel copied it
    }
    if (valid) {
        int p0 = __ldg(E.ptr + s), p1 = __ldg(E.ptr + s + 1);
        for (int p = p0 + k; p < p1; p += 8) {
            int j = __ldg(E.col + p);
            double v = xi[j];
            if (MODE == 2) {
                int jg = j + frow0;
                if (jg >= 0 && jg < nglob) v += R.w * __ldg(e + reg_agg(R, jg) - crow0);
            }
            acc += __ldg(E.val + p) * v;
        }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (valid && k == 0) {
        double xc = xi[i];
        if (MODE == 2) xc += R.w * __ldg(e + reg_agg(R, i + frow0) - crow0);
        xo[i] = xc + omega * (b[i] - acc) / __ldg(E.diag + s);
    }
}

// rc_I for coarse rows whose aggregate contains an exception row: one warp per coarse row,
// 4 lanes per fine row (entries of the exception CSR or the 7 band taps spread over the lanes),
// fixed-order shuffle reductions.
__global__ void __launch_bounds__(OMG_TPB) k_fix_crows(const int *__restrict__ crows, int ncrows, BandA<1> A, RegR R,
                                                       const double *__restrict__ x, const double *__restrict__ b,
                                                       double *__restrict__ rc, double xscale) {
    const int crow0 = 0, frow0 = 0;      // global-index pointers; xscale != 0: x aliases b and x_j = xscale*b_j
    int gt = blockIdx.x * OMG_TPB + threadIdx.x;
    int t = gt >> 5, lane = gt & 31;
    int k = lane >> 2, sub = lane & 3;
    bool valid = t < ncrows && k < R.k;
    double acc = 0.0;
    int I = 0, i = 0;
    if (t < ncrows) I = __ldg(crows + t);
    if (valid) {
        i = reg_cc(R, I + crow0) - frow0 + R.o[k];
        unsigned w = __ldg(A.e.mask + (i >> 5));
        if ((w >> (i & 31)) & 1u) {
            int s = __ldg(A.e.wpre + (i >> 5)) + __popc(w & ((1u << (i & 31)) - 1u));
            int p0 = __ldg(A.e.ptr + s), p1 = __ldg(A.e.ptr + s + 1);
            for (int p = p0 + sub; p < p1; p += 4) acc += __ldg(A.e.val + p) * x[__ldg(A.e.col + p)];
        } else {
            if (sub == 0) acc = A.b.diag * x[i];
            for (int q = sub; q < A.b.nb; q += 4) acc += A.b.coef[q] * x[i + A.b.off[q]];
        }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (xscale != 0.0) acc *= xscale;
    double r = (valid && sub == 0) ? (__ldg(b + i) - acc) : 0.0;
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 16);
    if (t < ncrows && lane == 0) rc[I] = R.w * r;
}

// ---------------------------------------------------------------- host side

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// Does level L carry a 3-D 7-point constant band this path can run?  Fills geometry + tiling.
static bool st3_params(Level &L, St3 *P, int *NT_out, bool xf) {
    if (L.kind == OMG_KIND_CSR || L.band.nb != 6) return false;
    if (L.nloc < env_int("OMG_ST_MIN_ROWS", 0)) return false;      // experiment knob: generic kernels on small levels
    const BandOp &B = L.band;
    if (B.off[3] != 1 || B.off[2] != -1 || B.off[4] != -B.off[1] || B.off[5] != -B.off[0]) return false;
    if (B.coef[2] != B.coef[3] || B.coef[1] != B.coef[4] || B.coef[0] != B.coef[5]) return false;
    int S1 = B.off[4], S2 = B.off[5];
    if (S1 < 32 || S1 > 2048 || (S1 & 1) || S2 % S1 != 0 || L.nloc % S2 != 0 || L.row0 % S2 != 0) return false;
    int NY = S2 / S1, NZ = L.nloc / S2;
    if ((NY & 1) || (NZ & 1) || NZ < 2) return false;
    if (L.pad < S2 + S1) return false;
    // rows up to 512 wide: 256 threads, 2 CTAs/SM; 1024 wide: one 512-thread CTA with full rows (measured equal to
    // or better than split rows); wider: split rows
    int NT = env_int("OMG_ST_NT", (S1 > 512 && S1 <= 1024) ? 512 : 256);
    if (NT != 256 && NT != 512 && NT != 128) NT = 256;
    if (L.nloc <= (1 << 19)) NT = 128;        // small levels: more, smaller CTAs
    // chunk width: full rows up to 512 columns per 256 threads, else split rows (largest even divisor)
    int XW = S1;
    int maxXW = env_int("OMG_ST_XW", 2 * NT);
    if (S1 > maxXW && NT == 256) {
        XW = 0;
        for (int w = maxXW; w >= 64; w -= 2)
            if (S1 % w == 0) {
                XW = w;
                break;
            }
        if (XW == 0) XW = S1;
    }
    const int XH = (XW == S1) ? 0 : 2;
    const int PITCH = (XW == S1) ? S1 : XW + 4;
    // rows per chunk: even, divides NY, patches and transform pairs within the per-thread maxima
    int maxTY = (4 * NT * ST_PPT) / XW;
    int envTY = env_int("OMG_ST_TY", 0);
    if (envTY > 0) maxTY = std::min(maxTY, envTY);
    if (L.nloc <= (1 << 19)) maxTY = std::min(maxTY, 8);
    int TY = 0;
    for (int t = std::min(maxTY, NY); t >= 2; --t) {
        if ((t & 1) || NY % t != 0) continue;
        if (xf && (t + 2) * PITCH > 2 * NT * (XH ? ST_TPT + 1 : ST_TPT)) continue;
        TY = t;
        break;
    }
    if (TY < 2) return false;
    *NT_out = NT;
    P->nloc = L.nloc;
    P->S1 = S1;
    P->S2 = S2;
    P->NZ = NZ;
    P->TY = TY;
    P->NP = TY * XW / 4;
    P->SPAN = (TY + 2) * PITCH;
    P->XW = XW;
    P->XC = S1 / XW;
    P->PITCH = PITCH;
    P->XH = XH;
    P->NYg = NY;
    P->cs1 = NY / 2;
    P->cs2 = S1 / 2;
    P->zg0 = L.row0 / S2;
    P->NZg = L.n / S2;
    P->cz0 = P->zg0 / 2;
    P->d = B.diag;
    P->c1 = B.coef[3];
    P->cS = B.coef[4];
    P->cP = B.coef[5];
    // ring depth: as deep as shared memory allows for the targeted CTAs per SM
    size_t budget = (NT <= 256 ? 113 : 226) * 1024;
    int NS = (int)std::min<size_t>(ST_MAXNS, (budget - 64) / ((size_t)P->SPAN * 8));
    int envNS = env_int("OMG_ST_NS", 0);
    if (envNS >= 4) NS = std::min(NS, envNS);
    if (NS < 4) return false;
    P->NS = NS;
    // z-segments (even length).  Two losses to balance: every segment re-stages 2 halo planes and ramps its
    // pipeline (~2.5 planes of dead time), and a grid that is not a whole number of waves of resident CTAs leaves
    // SMs idle in the last wave.  Pick the segment count that maximises ZL/(ZL+2.5) * CTAs/(waves*slots).
    int chunks = (NY / TY) * (S1 / XW);
    int per_sm = NT <= 128 ? 4 : (NT <= 256 ? 2 : 1);
    int slots = per_sm * std::max(g.sm_count, 1);
    int ZL = NZ;
    double best = -1.0;
    for (int nseg = 1; nseg <= NZ / 2; ++nseg) {
        int zl = (NZ + nseg - 1) / nseg;
        zl += zl & 1;
        int ns = (NZ + zl - 1) / zl;
        long long ctas = (long long)chunks * ns;
        long long waves = (ctas + slots - 1) / slots;
        if (waves > 12) break;      // (a cap of 6 left 256- and 512-chunk planes — the slab shapes — at 0.86 waves)
        double eff = (zl / (zl + 2.5)) * ((double)ctas / (double)(waves * slots));
        if (eff > best + 1e-9) {
            best = eff;
            ZL = zl;
        }
    }
    {
        int envZL = env_int("OMG_ST_ZL", 0);
        if (envZL >= 2) ZL = envZL + (envZL & 1);
    }
    ZL = std::min(std::max(ZL, 2), NZ);
    P->ZL = ZL;
    P->has_exc = 0;
    P->colour = -1;
    P->bpf = L.slab ? 0 : env_int("OMG_BPF", 2);      // (slab levels: the sharded kernels are left as measured at 2-8 ranks)
    P->epf = L.slab ? 0 : env_int("OMG_EPF", 4);
    P->use_cls = 0;
    if (L.kind == OMG_KIND_BAND_EXC && L.classed && XH == 0) {
        P->use_cls = 1;
        P->cls = L.cls;
    }
    P->needs_fix = (L.kind == OMG_KIND_BAND_EXC && !P->use_cls) ? 1 : 0;
    return true;
}

static size_t st3_smem(const St3 &P) { return (size_t)P.NS * P.SPAN * sizeof(double) + ST_MAXNS * sizeof(uint64_t); }

template <int MODE, int NT, bool SPLIT>
static bool st3_launch_nt(omg_hierarchy *h, St3 P) {
    static bool attr_set[2] = {false, false};
    size_t smem = st3_smem(P);
    if (smem > 227 * 1024) return false;
    void (*kern)(const St3);
    if constexpr (SPLIT)
        kern = k_st3x<MODE, 256>;
    else
        kern = P.use_cls ? k_st3<MODE, NT, true, false> : k_st3<MODE, NT, false, false>;
    if (!attr_set[P.use_cls ? 1 : 0]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr_set[P.use_cls ? 1 : 0] = true;
    }
    const int chunks = (P.NYg / P.TY) * P.XC;
    const int ZB = 2;       // planes next to a slab cut: the only ones that read the halo planes
    if (!SPLIT && !h->halo_req.empty() && !P.needs_fix && P.NZ >= 4 * ZB + 2) {
        // fused pull: every pending exchange must be one this kernel can do itself (its staged vector, MODE 2: e)
        HaloPull hx{}, he{};
        bool ok = true, have_x = false, have_e = false;
        int e_nzc = 0;
        for (auto &r : h->halo_req) {
            if (r.second == P.xi && !have_x) {
                have_x = dist_pull_params(h, *r.first, r.second, &hx);
                ok = ok && have_x;
            } else if (MODE == 2 && r.second == P.e && !have_e) {
                have_e = dist_pull_params(h, *r.first, r.second, &he);
                ok = ok && have_e;
                e_nzc = r.first->nloc / (P.cs1 * P.cs2);
            } else {
                ok = false;
            }
        }
        if (ok && have_x) {
            h->halo_req.clear();
            P.pull = hx.mine;
            P.pull_dn = hx.peer_dn;
            P.pull_up = hx.peer_up;
            P.xi_dn = hx.v_dn;
            P.xi_up = hx.v_up;
            P.pull_timeout = hx.timeout;
            P.e_dn = have_e ? he.v_dn : nullptr;
            P.e_up = have_e ? he.v_up : nullptr;
            P.e_nzc = e_nzc;
            P.boundary = 0;
            P.zlo = ZB;
            P.zhi = P.NZ - ZB;
            P.nseg = (P.zhi - P.zlo + P.ZL - 1) / P.ZL;
            if constexpr (!SPLIT) {
                static bool attr_p[2] = {false, false};
                void (*kp)(const St3) = P.use_cls ? k_st3<MODE, NT, true, true> : k_st3<MODE, NT, false, true>;
                if (!attr_p[P.use_cls ? 1 : 0]) {
                    if (cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
                        cudaGetLastError();
                        return false;
                    }
                    attr_p[P.use_cls ? 1 : 0] = true;
                }
                kp<<<dim3(chunks, P.nseg + 2), NT, smem, g.stream>>>(P);
            }
            return true;
        }
    }
    if (!h->halo_req.empty()) dist_halo_flush(h);
    if (h->halo_pending && P.NZ >= 4 * ZB + 2) {
        // interior planes do not touch the halos: run them while the exchange is in flight
        P.boundary = 0;
        P.zlo = ZB;
        P.zhi = P.NZ - ZB;
        kern<<<dim3(chunks, (P.zhi - P.zlo + P.ZL - 1) / P.ZL), NT, smem, g.stream>>>(P);
        dist_halo_wait(h);
        P.boundary = 1;
        kern<<<dim3(chunks, 2), NT, smem, g.stream>>>(P);
        h->launches++;
    } else {
        dist_halo_wait(h);
        P.boundary = 0;
        P.zlo = 0;
        P.zhi = P.NZ;
        kern<<<dim3(chunks, (P.NZ + P.ZL - 1) / P.ZL), NT, smem, g.stream>>>(P);
    }
    return true;
}

template <int MODE>
static bool st3_launch(omg_hierarchy *h, const St3 &P, int NT) {
    if constexpr (MODE == 4) {
        if (P.XH) return false;        // (no split-row instantiation of MODE 4)
    } else {
        if (P.XH) return NT == 256 ? st3_launch_nt<MODE, 256, true>(h, P) : false;     // split rows: 256-thread CTAs only
    }
    if (NT == 128) return st3_launch_nt<MODE, 128, false>(h, P);
    if (NT == 256) return st3_launch_nt<MODE, 256, false>(h, P);
    return st3_launch_nt<MODE, 512, false>(h, P);
}

static bool regular_matches(const Level &L, const St3 &P) {
    return L.regular && L.reg.alpha == 3 && L.reg.fs2 == P.S1 && L.reg.fs1 == P.NYg;
}

template <class T>
static inline T *V(const Level &L, T *p) { return p - L.row0; }

// fix-ups run on the owned exception slots / produced coarse rows only, with global-index pointers
static void fix_rows(Level &L, Level *C, int mode, const double *xi, const double *e, const double *b, double *xo,
                     double omega, int colour = -1) {
    int cnt = L.exc_s1 - L.exc_s0;
    if (L.kind != OMG_KIND_BAND_EXC || cnt <= 0) return;
    int grid = cdiv((int64_t)cnt * 8, OMG_TPB);
    if (mode == 0)
        k_fix_rows<0><<<grid, OMG_TPB, 0, g.stream>>>(L.exc_rows, L.exc_s0, cnt, L.exc_op(), L.reg, L.n, V(L, xi),
                                                      nullptr, V(L, b), V(L, xo), omega, L.colour, colour);
    else
        k_fix_rows<2><<<grid, OMG_TPB, 0, g.stream>>>(L.exc_rows, L.exc_s0, cnt, L.exc_op(), L.reg, L.n, V(L, xi),
                                                      V(*C, e), V(L, b), V(L, xo), omega, L.colour, colour);
}

// xscale != 0: x is not read, x_j := xscale * b_j (first sweep from zero with a uniform diagonal)
static void fix_crows(Level &L, const double *x, const double *b, double *rcv, double xscale = 0.0) {
    int cnt = L.crow_t1 - L.crow_t0;
    if (L.kind != OMG_KIND_BAND_EXC || cnt <= 0) return;
    BandA<1> A{L.band, L.exc_op()};
    k_fix_crows<<<cdiv((int64_t)cnt * 32, OMG_TPB), OMG_TPB, 0, g.stream>>>(
        L.exc_crows + L.crow_t0, cnt, A, L.reg, xscale != 0.0 ? V(L, b) : V(L, x), V(L, b), rcv, xscale);
}

// ---------------------------------------------------------------- 2-D host side

static bool st2_params(Level &L, St2 *P, bool need_regular, bool rb = false) {
    if (L.kind == OMG_KIND_CSR) return false;
    if (L.slab && (rb || getenv("OMG_NO_ST2_SLAB"))) return false;     // the single-pass sweep needs two halo rows
    const BandOp &B = L.band;
    int N = 0;
    double c1 = 0, cN = 0, cD = 0;
    bool oned = false;
    if (B.nb == 2) {          // 1-D {-1, 1}: view the vector as rows of N (any even divisor): flat semantics are the same
        if (B.off[0] != -1 || B.off[1] != 1 || B.coef[0] != B.coef[1]) return false;
        for (int w = 2048; w >= 128; w >>= 1)
            if (L.nloc % w == 0 && L.nloc / w >= 4) {
                N = w;
                break;
            }
        if (N == 0) return false;
        c1 = B.coef[1];
        oned = true;
    } else if (B.nb == 4) {          // {-(N+1), -1, 1, N+1}
        if (B.off[1] != -1 || B.off[2] != 1 || B.off[0] != -B.off[3]) return false;
        if (B.coef[1] != B.coef[2] || B.coef[0] != B.coef[3]) return false;
        N = B.off[3] - 1;
        c1 = B.coef[2];
        cD = B.coef[3];
    } else if (B.nb == 6) {   // {-(N+1), -N, -1, 1, N, N+1}
        if (B.off[2] != -1 || B.off[3] != 1 || B.off[5] != B.off[4] + 1 || B.off[1] != -B.off[4] ||
            B.off[0] != -B.off[5])
            return false;
        if (B.coef[2] != B.coef[3] || B.coef[1] != B.coef[4] || B.coef[0] != B.coef[5]) return false;
        N = B.off[4];
        c1 = B.coef[3];
        cN = B.coef[4];
        cD = B.coef[5];
    } else
        return false;
    if (N < 128 || (N & 1) || L.nloc % N != 0) return false;
    int NY = L.nloc / N;
    if ((!oned && (NY & 1)) || NY < 4) return false;
    if (L.pad < N + 4) return false;
    if (L.slab && L.halo < N + 2) return false;       // rows -1 and NY (+2 halo columns) come from the neighbours
    if (need_regular && !oned && !(L.regular && L.reg.alpha == 2 && L.reg.fs2 == N)) return false;
    if (need_regular && oned && !(L.regular && L.reg.alpha == 1)) return false;
    P->oned = oned ? 1 : 0;
    int XW = 0;
    // fused two-colour sweep (k_st2rb): 2 halo pairs per row on top, 4 halo columns per side, a 3-row mid ring
    const int maxw = rb ? std::min(env_int("OMG_RB_MAXW", 2048), 2 * (ST2_NT * ST2RB_PP - 2)) : 2 * ST2_NT * ST2_PP;
    for (int w = std::min(N, maxw); w >= 64; w -= 2)
        if (N % w == 0) {
            XW = w;
            break;
        }
    if (XW == 0) return false;
    if (rb && L.pad < 2 * N + 4) return false;        // the fused sweep stages the row slots -2 and NY+1
    if (!rb && (XW + 4) / 2 > ST2_NT * ST2_TPT) return false;
    if (rb && (XW + 8) / 2 > ST2_NT * ST2RB_TPT) return false;
    P->N = N;
    P->NY = NY;
    P->XW = XW;
    P->XC = N / XW;
    P->PITCH = rb ? XW + 8 : XW + 4;
    size_t budget = 113 * 1024;
    size_t fixed = 64 + (rb ? (size_t)3 * (XW + 4) * 8 : 0);
    if (budget < fixed + 4 * (size_t)P->PITCH * 8) return false;
    int NS = (int)std::min<size_t>(8, (budget - fixed) / ((size_t)P->PITCH * 8));
    if (NS < 4) return false;
    P->NS = NS;
    P->cs = N / 2;
    P->yg0 = L.row0 / N;
    P->NYg = L.n / N;
    P->d = B.diag;
    P->c1 = c1;
    P->cN = cN;
    P->cD = cD;
    P->colour = -1;
    // (16-byte alignment of the prefetched b / e row pieces; slab levels are left as measured at 2-8 ranks)
    // rows ahead 0 / 1 / 2 / 3 / 4 / 8: two-colour V(1,1) on 8192^2 1.52 / 1.47 / 1.33 / 1.35 / 1.43 / 1.63 ms per cycle
    // (Jacobi 1.09 -> 1.01 at 2); 1-D 2^24 (rows of 2048) 0.63 -> 0.59 at 2, 0.66 at 8
    P->bpf = (L.slab || (N & 3) || (XW & 3)) ? 0 : env_int("OMG_BPF2", 2);
    P->cflat = L.colour.flat;
    P->use_cls = 0;
    if (L.kind == OMG_KIND_BAND_EXC && L.classed2 && !oned && B.nb == 6) {
        P->use_cls = 1;
        for (int t = 0; t < 3; ++t) {
            P->c2l[t] = L.c2l[t];
            P->c2r[t] = L.c2r[t];
        }
    }
    {   // y-segments: same trade-off as the z-segments of the 3-D kernel
        int slots = 2 * std::max(g.sm_count, 1);
        int YL = NY;
        double best = -1.0;
        for (int nseg = 1; nseg <= NY / 2; ++nseg) {
            int yl = (NY + nseg - 1) / nseg;
            yl += yl & 1;
            int ns = (NY + yl - 1) / yl;
            long long ctas = (long long)P->XC * ns;
            long long waves = (ctas + slots - 1) / slots;
            if (waves > 6) break;
            double eff = (yl / (yl + (rb ? 4.5 : 2.5))) * ((double)ctas / (double)(waves * slots));   // rows staged beyond the segment
            if (eff > best + 1e-9) {
                best = eff;
                YL = yl;
            }
        }
        P->YL = std::min(std::max(YL, 2), NY);
    }
    return true;
}

template <int MODE>
static bool st2_launch(omg_hierarchy *h, const St2 &P) {
    static bool attr_set[2] = {false, false};
    size_t smem = (size_t)P.NS * P.PITCH * sizeof(double) + 8 * sizeof(uint64_t);
    void (*kern)(const St2) = P.use_cls ? k_st2<MODE, true> : k_st2<MODE, false>;
    if (!attr_set[P.use_cls ? 1 : 0]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr_set[P.use_cls ? 1 : 0] = true;
    }
    dist_halo_wait(h);
    kern<<<dim3(P.XC, (P.NY + P.YL - 1) / P.YL), ST2_NT, smem, g.stream>>>(P);
    return true;
}

// The 2-D kernels index the coarse vector by LOCAL coarse row (fine local row >> 1); the coarse level may be a slab
// (its owned row 0 is that row) or replicated (global rows): rebase accordingly.
static const double *st2_coarse_base(const Level &L, const Level &C, const St2 &Q, const double *e) {
    (void)L;
    long long first = (long long)(Q.oned ? Q.yg0 : (Q.yg0 >> 1)) * Q.cs;    // global coarse row*cs of local coarse row 0
    return e - C.row0 + first;
}

static bool colour2_ok(const Level &L, const St2 &P) {
    return L.colour.flat || (!P.oned && L.colour.alpha == 2 && L.colour.s2 == P.N);
}

// fused two-colour sweep (k_st2rb): pure-band, unsharded levels
static bool st2rb_params(Level &L, St2 *P, bool need_regular) {
    static const bool off = getenv("OMG_NO_RBFUSE") != nullptr;
    if (off || L.kind == OMG_KIND_CSR || L.slab) return false;
    if (!st2_params(L, P, need_regular, true) || !colour2_ok(L, *P)) return false;
    return L.kind == OMG_KIND_BAND || P->use_cls;     // exception rows only if the kernel corrects them itself
}

template <int MODE>
static bool st2rb_launch(omg_hierarchy *h, const St2 &P) {
    static bool attr_set[4] = {false, false, false, false};
    size_t smem = ((size_t)P.NS * (P.XW + 8) + (size_t)3 * (P.XW + 4)) * sizeof(double) + 8 * sizeof(uint64_t);
    const bool hasn = P.cN != 0.0;
    void (*kern)(const St2) = P.use_cls ? (hasn ? k_st2rb<MODE, true, true> : k_st2rb<MODE, true, false>)
                                        : (hasn ? k_st2rb<MODE, false, true> : k_st2rb<MODE, false, false>);
    const int vi = (P.use_cls ? 1 : 0) + (hasn ? 2 : 0);
    if (!attr_set[vi]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr_set[vi] = true;
    }
    dist_halo_wait(h);
    kern<<<dim3(P.XC, (P.NY + P.YL - 1) / P.YL), ST2_NT, smem, g.stream>>>(P);
    return true;
}

bool stencil_jacobi(omg_hierarchy *h, Level &L, const double *xi, const double *b, double *xo, double omega) {
    St3 P{};
    int NT;
    St2 Q{};
    if (st2_params(L, &Q, false)) {
        if (L.kind == OMG_KIND_BAND_EXC && !L.exc_rows) return false;
        Q.xi = xi;
        Q.b = b;
        Q.xo = xo;
        Q.wod = omega / Q.d;
        if (!st2_launch<0>(h, Q)) return false;
        if (!Q.use_cls) fix_rows(L, nullptr, 0, xi, nullptr, b, xo, omega);
        return true;
    }
    if (!st3_params(L, &P, &NT, false)) return false;
    if (L.kind == OMG_KIND_BAND_EXC && !L.exc_rows) return false;
    P.xi = xi;
    P.b = b;
    P.xo = xo;
    P.wod = omega / P.d;
    if (!st3_launch<0>(h, P, NT)) return false;
    if (!P.use_cls) fix_rows(L, nullptr, 0, xi, nullptr, b, xo, omega);
    return true;
}

// rcv: coarse output indexable by GLOBAL coarse row
bool stencil_residual_restrict(omg_hierarchy *h, Level &L, Level &C, const double *x, const double *b, double *rcv) {
    (void)C;
    St3 P{};
    int NT;
    St2 Q{};
    if (st2_params(L, &Q, true)) {
        if (L.kind == OMG_KIND_BAND_EXC && !L.exc_crows) return false;
        Q.xi = x;
        Q.b = b;
        Q.rc = rcv + L.piece_row0;
        Q.w = L.Rw;
        if (!st2_launch<1>(h, Q)) return false;
        if (!Q.use_cls) fix_crows(L, x, b, rcv);
        return true;
    }
    if (!st3_params(L, &P, &NT, false) || !regular_matches(L, P)) return false;
    if (L.kind == OMG_KIND_BAND_EXC && !L.exc_crows) return false;
    P.xi = x;
    P.b = b;
    P.rc = rcv + L.piece_row0;
    P.w = L.Rw;
    if (!st3_launch<1>(h, P, NT)) return false;
    if (!P.use_cls) fix_crows(L, x, b, rcv);
    return true;
}

// xi == nullptr: applicability probe only
bool stencil_prolong_jacobi(omg_hierarchy *h, Level &L, Level &C, const double *xi, const double *e,
                            const double *b, double *xo, double omega) {
    St3 P{};
    int NT;
    St2 Q{};
    if (st2_params(L, &Q, true)) {
        if (L.kind == OMG_KIND_BAND_EXC && !L.exc_rows) return false;
        if (!xi) return true;
        Q.xi = xi;
        Q.b = b;
        Q.xo = xo;
        Q.e = st2_coarse_base(L, C, Q, e);
        Q.w = L.Rw;
        Q.wod = omega / Q.d;
        if (!st2_launch<2>(h, Q)) return false;
        if (!Q.use_cls) fix_rows(L, &C, 2, xi, e, b, xo, omega);
        return true;
    }
    if (!st3_params(L, &P, &NT, true) || !regular_matches(L, P)) return false;
    if (L.kind == OMG_KIND_BAND_EXC && !L.exc_rows) return false;
    if (!xi) return true;
    P.xi = xi;
    P.b = b;
    P.xo = xo;
    P.e = e;
    P.cz0 = C.row0 / (P.cs1 * P.cs2);
    P.w = L.Rw;
    P.wod = omega / P.d;
    if (!st3_launch<2>(h, P, NT)) return false;
    if (!P.use_cls) fix_rows(L, &C, 2, xi, e, b, xo, omega);
    return true;
}

// first Jacobi sweep from the zero iterate + residual + restriction, one pass over b.
// rcv: coarse output indexable by GLOBAL coarse row.  On a slab level b's halos must be valid.
bool stencil_jacobi0_residual_restrict(omg_hierarchy *h, Level &L, Level &C, const double *b, double *xo,
                                       double *rcv, double omega) {
    (void)C;
    St3 P{};
    int NT;
    St2 Q{};
    if (st2_params(L, &Q, true)) {
        if (L.kind == OMG_KIND_BAND_EXC && (!L.exc_crows || !L.exc_diag_uniform)) return false;
        if (!b) return true;
        Q.xi = b;
        Q.b = b;
        Q.xo = xo;
        Q.rc = rcv + L.piece_row0;
        Q.w = L.Rw;
        Q.wod = omega / Q.d;
        if (!st2_launch<3>(h, Q)) return false;
        if (!Q.use_cls) fix_crows(L, xo, b, rcv, Q.wod);
        return true;
    }
    if (!st3_params(L, &P, &NT, true) || !regular_matches(L, P)) return false;
    if (L.kind == OMG_KIND_BAND_EXC && !L.exc_crows) return false;
    if (L.slab && L.kind == OMG_KIND_BAND_EXC && !L.exc_diag_uniform) return false;   // halo rows' a_ii unknown here
    if (!b) return true;       // applicability probe
    P.xi = b;
    P.b = b;
    P.xo = xo;
    P.rc = rcv + L.piece_row0;
    P.w = L.Rw;
    P.omega = omega;
    P.wod = omega / P.d;
    if (L.kind == OMG_KIND_BAND_EXC && !L.exc_diag_uniform) {
        P.has_exc = 1;       // per-row a_ii lookups in the transform (slow path; not hit by Poisson hierarchies)
        P.exc = L.exc_op();
    }
    if (!P.has_exc && !P.XH && !L.slab && !getenv("OMG_NO_MODE4")) {        // (read per call: the parity test runs both)
        if (!st3_launch<4>(h, P, NT)) return false;
    } else if (!st3_launch<3>(h, P, NT)) {
        return false;
    }
    // fix-up: x_j = omega b_j / d is recomputed from b when the diagonal is uniform (x's halos are not filled yet)
    if (!P.use_cls) fix_crows(L, xo, b, rcv, L.exc_diag_uniform ? P.wod : 0.0);
    return true;
}

static bool grid_colour_matches(const Level &L, const St3 &P) {
    return !L.colour.flat && L.colour.alpha == 3 && L.colour.s2 == P.S1 && L.colour.s1 == P.NYg;
}

// one colour half-sweep of the two-colour Gauss-Seidel: xo = xi, rows of `colour` relaxed with omega = 1
bool stencil_colour_relax(omg_hierarchy *h, Level &L, int colour, const double *xi, const double *b, double *xo) {
    St3 P{};
    int NT;
    St2 Q{};
    if (st2_params(L, &Q, false) && colour2_ok(L, Q)) {
        if (L.kind == OMG_KIND_BAND_EXC && !L.exc_rows) return false;
        Q.xi = xi;
        Q.b = b;
        Q.xo = xo;
        Q.wod = 1.0 / Q.d;
        Q.colour = colour;
        if (!st2_launch<0>(h, Q)) return false;
        if (!Q.use_cls) fix_rows(L, nullptr, 0, xi, nullptr, b, xo, 1.0, colour);
        return true;
    }
    if (!st3_params(L, &P, &NT, false) || !grid_colour_matches(L, P)) return false;
    if (L.kind == OMG_KIND_BAND_EXC && !L.exc_rows) return false;
    P.xi = xi;
    P.b = b;
    P.xo = xo;
    P.wod = 1.0 / P.d;
    P.colour = colour;
    if (!st3_launch<0>(h, P, NT)) return false;
    if (!P.use_cls) fix_rows(L, nullptr, 0, xi, nullptr, b, xo, 1.0, colour);
    return true;
}

// y = xi + R^T e ; xo = y with the rows of `colour` relaxed (first half of the post-smoothing sweep)
bool stencil_prolong_colour_relax(omg_hierarchy *h, Level &L, Level &C, int colour, const double *xi,
                                  const double *e, const double *b, double *xo) {
    St3 P{};
    int NT;
    St2 Q{};
    if (st2_params(L, &Q, true) && colour2_ok(L, Q)) {
        if (L.kind == OMG_KIND_BAND_EXC && !L.exc_rows) return false;
        if (!xi) return true;
        Q.xi = xi;
        Q.b = b;
        Q.xo = xo;
        Q.e = st2_coarse_base(L, C, Q, e);
        Q.w = L.Rw;
        Q.wod = 1.0 / Q.d;
        Q.colour = colour;
        if (!st2_launch<2>(h, Q)) return false;
        if (!Q.use_cls) fix_rows(L, &C, 2, xi, e, b, xo, 1.0, colour);
        return true;
    }
    if (!st3_params(L, &P, &NT, true) || !regular_matches(L, P) || !grid_colour_matches(L, P)) return false;
    if (L.kind == OMG_KIND_BAND_EXC && !L.exc_rows) return false;
    if (!xi) return true;
    P.xi = xi;
    P.b = b;
    P.xo = xo;
    P.e = e;
    P.cz0 = C.row0 / (P.cs1 * P.cs2);
    P.w = L.Rw;
    P.wod = 1.0 / P.d;
    P.colour = colour;
    if (!st3_launch<2>(h, P, NT)) return false;
    if (!P.use_cls) fix_rows(L, &C, 2, xi, e, b, xo, 1.0, colour);
    return true;
}

// ================================================================ 3-D single-pass two-colour sweep
//
// Both colour half-sweeps of oracle.rbgs (grid-parity colouring, colour c0 first, same-colour couplings lagged) in
// ONE pass over x: 24 B/row instead of 2 x 24.  Same idea as k_st2rb one dimension up, marching in z over full-row
// chunks of TY rows:
//   pass A(p)   relaxes the colour-c0 points of plane p from the RAW planes p-1, p, p+1  -> "mid" plane p
//   pass B(p-1) relaxes the other colour of plane p-1 from the MID planes p-2, p-1, p    -> output plane p-1
// Pass B needs mid one row beyond the chunk, so pass A also runs on the rows y0-1 and y0+TY (recomputed, bit-identical
// to the owner's values) and the raw planes are staged with two halo rows per side (one TMA bulk copy per plane, ring
// of NS stages, as in k_st3).
//
// Work items are 2x2 (x,y) patches on even rows/columns: on every plane exactly one diagonal of a patch has the
// colour relaxed by pass A and the other diagonal the colour relaxed by pass B one plane earlier — which diagonal is
// the same for ALL patches of a plane (it flips from plane to plane), so the position logic is resolved by one
// uniform branch per patch instead of per-value selects.  A thread keeps its patches in registers from plane to
// plane (raw patch of plane p, mid patch of plane p-1, the two relaxed values of mid plane p-2, the two b values
// pass B needs), so centres, z-neighbours and the in-patch x/y neighbours never touch shared memory: per patch and
// plane pass A reads 2 LDS.128 (its raw patch of plane p+1) + 4 LDS.64 (out-of-patch neighbours) and writes its mid
// patch (2 STS.128), pass B reads 4 LDS.64.  One __syncthreads per plane.
// Flat-index semantics (openmg/operators.py:244-256: no boundary breaks): full rows make the x-wraps contiguous;
// patch rows outside [0,NY) are rows of the neighbouring plane, so their diagonal is the flipped one; points
// outside [0,n) are never relaxed and stay zero (the pads of the vectors).
#define RB3_PPT 2          // full patch slots per thread
#define RB3_NT 512
#define RB3_TC 3           // MODE 2: cells of the staged span per thread (max)

struct Rb3 {
    const double *xi;
    const double *b;
    const double *e;       // coarse correction (MODE 2)
    double *xo;
    int S1, S2, NY, NZ;    // row length, plane size, rows per plane, planes
    int TY, ZL;            // rows per chunk, planes per z-segment
    int RS, MS;            // staged raw plane (TY+4 rows) / mid plane (TY+2 rows) in doubles
    int cs1, cs2;          // coarse rows per plane, coarse row length
    int c0;                // colour relaxed first
    int bpf;               // planes by which thread 0 prefetches b into L2 ahead of its loads (0: off)
    int epf;               // the same for the coarse correction e (MODE 2), in fine planes
    ClsTab cls;            // Galerkin levels (CLS): class correction taps on the x/y grid boundaries
    double d, c1, cS, cP, wod, w;
};

// class-correction taps of one point: planes lo/c/hi are indexed like the staged raw planes, o = the point's offset
__device__ __forceinline__ double rb3_corr(const ClsTab &T, int cls, const double *lo, const double *c,
                                           const double *hi, int o) {
    double a = 0.0;
    for (int t = 0; t < T.ntap[cls]; ++t) {
        const int dz = T.dz[cls][t];
        const double *pl = dz == 0 ? c : (dz > 0 ? hi : lo);
        a += T.coef[cls][t] * pl[o + T.soff[cls][t]];
    }
    return a;
}

// MODE 0: one sweep on xi.  MODE 2: y = xi + R^T e, then one sweep on y.  MODE 3: one sweep from the ZERO iterate
// (openmg/__init__.py:191-192: coarse levels start from zeros): xi is b, x is neither read nor cleared beforehand.
template <int MODE, bool CLS>
__global__ void __launch_bounds__(RB3_NT, 1) k_rb3(const Rb3 P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NT = RB3_NT;
    // class taps reach one plane down: the CLS variant keeps raw plane p-1 and mid plane p-2 resident as well
    constexpr int NS = CLS ? 4 : 3;
    constexpr int NM = CLS ? 3 : 2;
    constexpr int KEEP = CLS ? 1 : 0;
    constexpr int KS = 4 * NT;             // slot k+1 is NT patches = 2 NT / HX patch rows = 4 NT doubles further on
    const int S1 = P.S1, HX = S1 >> 1;
    const int RS = P.RS, MS = P.MS;
    double *raw = reinterpret_cast<double *>(smem_raw);
    double *mid = raw + (size_t)NS * RS;
    uint64_t *full = reinterpret_cast<uint64_t *>(mid + NM * (size_t)MS);
    const int tid = threadIdx.x;
    const int y0 = (int)blockIdx.x * P.TY;
    const int z0 = (int)blockIdx.y * P.ZL;
    const int z1 = min(z0 + P.ZL, P.NZ);
    const uint32_t span_bytes = (uint32_t)RS * 8u;
    const int pfirst = max(z0 - 2, -1);           // raw planes below -1 / above NZ are all zero and never staged
    const int plast = min(z1 + 1, P.NZ);
    const long long gbase = (long long)(y0 - 2) * S1;      // staged offset o <-> in-plane offset gbase + o

    auto slot_of = [&](int p) { return (p - pfirst) % NS; };
    auto issue = [&](int p) {
        int s_ = slot_of(p);
        mbar_expect_tx(full + s_, span_bytes);
        bulk_g2s(raw + (size_t)s_ * RS, P.xi + (long long)p * P.S2 + gbase, span_bytes, full + s_);
    };
    auto wait_plane = [&](int p) {
        int k = p - pfirst;
        mbar_wait(full + (k % NS), (uint32_t)((k / NS) & 1));
    };
    // MODE 2: y = x + R^T e is formed in the staged raw planes: plane pl += w e[cell] on the whole staged span.  A
    // thread owns up to RB3_TC cells (coarse row x coarse column; 2 rows x 1 pair of the plane) of the span; their e
    // values only change every other plane and are fetched one coarse plane ahead.  Rows outside [0,NY) are rows of
    // the neighbouring plane.
    // Cell m of thread tid is the staged row pair pj + m NT/HX, pair column pi_ (NT is a multiple of HX).
    constexpr int TC = RB3_TC;
    const int ncr = (P.TY + 4) >> 1;             // staged row pairs
    const int tpj = tid / HX, tpi = tid - tpj * HX;
    const bool firstc = (y0 == 0), lastc = (y0 + P.TY == P.NY);
    double enxt[TC];
    int eplane_nxt = -(1 << 30);       // the raw plane enxt was fetched for
    auto load_e = [&](int pl, double *ev) {      // w e of this thread's cells for raw plane pl
        if constexpr (MODE == 2) {
            const double *ep = P.e + tpi;
#pragma unroll
            for (int m = 0; m < TC; ++m) {
                const int cr = tpj + m * (NT / HX);
                int zz = pl, crow = (y0 >> 1) - 1 + cr;
                if (firstc && cr == 0) {          // rows -2, -1: the last coarse row of the plane below
                    crow = P.cs1 - 1;
                    zz = pl - 1;
                }
                if (lastc && cr == ncr - 1) {     // rows NY, NY+1: the first coarse row of the plane above
                    crow = 0;
                    zz = pl + 1;
                }
                ev[m] = 0.0;
                if (cr < ncr && zz >= 0 && zz < P.NZ)
                    ev[m] = P.w * __ldg(ep + ((zz >> 1) * P.cs1 + crow) * P.cs2);
            }
        }
    };
    auto apply_e = [&](int pl, const double *ev) {
        if constexpr (MODE == 2) {
            double *sp_ = raw + (size_t)slot_of(pl) * RS + 2 * tpj * S1 + 2 * tpi;
#pragma unroll
            for (int m = 0; m < TC; ++m) {
                if (tpj + m * (NT / HX) >= ncr) continue;
                double *q = sp_ + m * KS;
                double2 x0 = lds2(q), x1 = lds2(q + S1);
                x0.x += ev[m];
                x0.y += ev[m];
                x1.x += ev[m];
                x1.y += ev[m];
                sts2(q, x0);
                sts2(q + S1, x1);
            }
        }
    };
    // wrapped staged rows shift the plane by one, so a thread's cells do not all change coarse plane on the same
    // step: simply reload e for every plane (L1/L2 hits), one plane ahead of its use
    auto transform = [&](int pl) {       // raw plane pl has landed: add its e (fetched during the previous call)
        if constexpr (MODE == 2) {
            if (eplane_nxt != pl) load_e(pl, enxt);
            apply_e(pl, enxt);
            if (pl + 1 <= plast) {
                load_e(pl + 1, enxt);
                eplane_nxt = pl + 1;
            }
        }
    };

    if (tid == 0) {
        for (int s_ = 0; s_ < NS; ++s_) mbar_init(full + s_, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int p = pfirst; p < pfirst + NS && p <= plast; ++p) issue(p);

    // ---- work items
    // full slots: the TY/2 patch rows of the chunk (grid rows y0 .. y0+TY-1), 2x2 patches on even rows / columns;
    //             pass A and pass B, all state in registers.  Slot k of thread tid is patch tid + k NT.
    // halo item:  one x-pair of the grid row y0-1 (threads [0,HX)) or y0+TY (threads [HX,2HX)); pass A only.  These
    //             rows may belong to the neighbouring plane (first / last chunk): their colour parity then flips.
    const int nslots = ((P.TY >> 1) * HX) / NT;          // 1 or 2 (host: TY/2 * HX is a multiple of NT)
    const int pj = tid / HX, pi_ = tid - pj * HX;
    const int ro = (2 * pj + 2) * S1 + 2 * pi_;          // row a of slot 0 inside a staged raw plane; mid: ro - S1
    const bool hact = tid < 2 * HX;
    const bool hwhich = tid >= HX;
    const int ho = (hwhich ? (P.TY + 2) * S1 + 2 * (tid - HX) : S1 + 2 * tid);
    const bool hwrap = hwhich ? (y0 + P.TY == P.NY) : (y0 == 0);
    // in-pair position the halo item relaxes on plane p: (hpar + p) & 1
    const int hpar = (P.c0 + (hwhich ? 0 : 1) + (hwrap ? 1 : 0)) & 1;     // y0, TY even: row y0-1 is odd, y0+TY even

    // carried from plane to plane, per slot: the mid patch of plane p-1, the relaxed values of mid plane p-2 and the
    // b values of plane p-1 at the positions pass B(p-1) relaxes, the b patch of plane p (fetched one step ahead).
    // The raw patch of plane p is re-read from the staged plane (it is resident anyway): carrying it would cost 8
    // more registers per slot, and a single spilled register serialises the whole b prefetch behind it.
    double2 ma[RB3_PPT], mb[RB3_PPT];
    double ua[RB3_PPT], ub[RB3_PPT];
    double ga[RB3_PPT], gb[RB3_PPT];
    double2 ba[RB3_PPT], bb[RB3_PPT];
    double hzm = 0.0, hb = 0.0;           // halo item: raw plane p-1 at the relaxed position, its b value
#pragma unroll
    for (int k = 0; k < RB3_PPT; ++k) {
        ma[k] = mb[k] = ba[k] = bb[k] = make_double2(0.0, 0.0);
        ua[k] = ub[k] = ga[k] = gb[k] = 0.0;
    }

    // One plane step.  DG = 0: pass A relaxes (row a, .x) and (row b, .y) of every patch; DG = 1: (row a, .y) and
    // (row b, .x).  Steps p < z0-1 only fill the register pipeline.
    auto step = [&](auto dgc, const int p) {
        constexpr int DG = decltype(dgc)::value;
        constexpr int EA = DG, EB = 1 - DG;
        auto el = [](const double2 &v, int e_) { return e_ ? v.y : v.x; };
        const bool real = p >= z0 - 1;
        const bool relax = real && p >= 0 && p < P.NZ;
        const bool has_cur = (p >= pfirst && p <= plast);
        const bool has_next = (p + 1 >= pfirst && p + 1 <= plast);
        const int he = (hpar + p) & 1;
        if (MODE != 2 && has_next) wait_plane(p + 1);      // MODE 2: waited for and transformed during the previous step
        const double *rawc = raw + (size_t)slot_of(max(p, pfirst)) * RS;
        const double *rawn = raw + (size_t)slot_of(max(p + 1, pfirst)) * RS;
        const double *rawm = raw + (size_t)slot_of(max(p - 1, pfirst)) * RS;
        const int mq = p - z0 + 1 + NM;                                     // mid plane p lives in slot mq % NM
        double *midw = mid + (size_t)(mq % NM) * MS - S1;                  // mid row = staged row - 1
        const double *midp = mid + (size_t)((mq - 1) % NM) * MS - S1;
        const double *midpp = mid + (size_t)((mq - 2) % NM) * MS - S1;
        double2 ra[RB3_PPT], rb[RB3_PPT];      // raw patch of plane p; pass A turns it into the mid patch in place
#pragma unroll
        for (int k = 0; k < RB3_PPT; ++k) {
            ra[k] = rb[k] = make_double2(0.0, 0.0);
            if (has_cur && k < nslots) {
                ra[k] = lds2(rawc + ro + k * KS);
                rb[k] = lds2(rawc + ro + k * KS + S1);
            }
        }
        double g3a[RB3_PPT], g3b[RB3_PPT];     // MODE 3: b of plane p at the positions pass B(p) relaxes
        if (MODE == 3) {
            // ---- pass A from the zero iterate (the staged planes hold b): every coupling multiplies a zero, so the
            // colour-c0 points become b/a_ii and everything else stays zero — no neighbour is read
#pragma unroll
            for (int k = 0; k < RB3_PPT; ++k) {
                g3a[k] = el(ra[k], 1 - EA);
                g3b[k] = el(rb[k], 1 - EB);
                const double nva = P.wod * el(ra[k], EA), nvb = P.wod * el(rb[k], EB);
                ra[k] = EA ? make_double2(0.0, nva) : make_double2(nva, 0.0);
                rb[k] = EB ? make_double2(0.0, nvb) : make_double2(nvb, 0.0);
            }
        } else if (relax) {
            // ---- pass A
#pragma unroll
            for (int k = 0; k < RB3_PPT; ++k) {
                if (k >= nslots) continue;
                const int o = ro + k * KS;
                const double ca = el(ra[k], EA), cb = el(rb[k], EB);
                const double za = has_next ? rawn[o + EA] : 0.0, zb = has_next ? rawn[o + S1 + EB] : 0.0;
                double nva, nvb;
                {
                    const double xl = EA ? ra[k].x : rawc[o - 1];
                    const double xr = EA ? rawc[o + 2] : ra[k].y;
                    const double yn = rawc[o - S1 + EA], ys = el(rb[k], EA);
                    double ax = P.d * ca + P.c1 * (xl + xr) + P.cS * (yn + ys) + P.cP * (el(ma[k], EA) + za);
                    if (CLS) {
                        const int cls = ((y0 + 2 * (pj + k * (NT / HX)) == 0) ? 0 : 3) +
                                        (EA ? (pi_ == HX - 1 ? 2 : 1) : (pi_ == 0 ? 0 : 1));
                        if (cls != 4) ax += rb3_corr(P.cls, cls, rawm, rawc, rawn, o + EA);
                    }
                    nva = ca + P.wod * (el(ba[k], EA) - ax);
                }
                {
                    const double xl = EB ? rb[k].x : rawc[o + S1 - 1];
                    const double xr = EB ? rawc[o + S1 + 2] : rb[k].y;
                    const double yn = el(ra[k], EB), ys = rawc[o + 2 * S1 + EB];
                    double ax = P.d * cb + P.c1 * (xl + xr) + P.cS * (yn + ys) + P.cP * (el(mb[k], EB) + zb);
                    if (CLS) {
                        const int cls = ((y0 + 2 * (pj + k * (NT / HX)) + 1 == P.NY - 1) ? 6 : 3) +
                                        (EB ? (pi_ == HX - 1 ? 2 : 1) : (pi_ == 0 ? 0 : 1));
                        if (cls != 4) ax += rb3_corr(P.cls, cls, rawm, rawc, rawn, o + S1 + EB);
                    }
                    nvb = cb + P.wod * (el(bb[k], EB) - ax);
                }
                if (EA) ra[k].y = nva; else ra[k].x = nva;
                if (EB) rb[k].y = nvb; else rb[k].x = nvb;
            }
        }
        if (real) {
            // mid plane p (planes -1 and NZ: nothing was relaxed, it is the all-zero raw plane)
#pragma unroll
            for (int k = 0; k < RB3_PPT; ++k) {
                if (k >= nslots) continue;
                sts2(midw + ro + k * KS, ra[k]);
                sts2(midw + ro + k * KS + S1, rb[k]);
            }
        }
        if (hact) {
            // the halo item's row may belong to the neighbouring plane: relaxed iff that point lies inside [0,n)
            const int hshift = hwrap ? (hwhich ? 1 : -1) : 0;
            double2 hr = make_double2(0.0, 0.0);
            if (has_cur) hr = lds2(rawc + ho);
            double2 hm = hr;
            if (MODE == 3) {
                const double nv = P.wod * (he ? hr.y : hr.x);
                hm = he ? make_double2(0.0, nv) : make_double2(nv, 0.0);
            } else if (real && p + hshift >= 0 && p + hshift < P.NZ) {
                const double c = he ? hr.y : hr.x;
                const double xl = he ? hr.x : rawc[ho - 1];
                const double xr = he ? rawc[ho + 2] : hr.y;
                double ax = P.d * c + P.c1 * (xl + xr) + P.cS * (rawc[ho - S1 + he] + rawc[ho + S1 + he]) +
                            P.cP * (hzm + (has_next ? rawn[ho + he] : 0.0));
                if (CLS) {
                    const int hx_ = hwhich ? tid - HX : tid;
                    const int cls = (hwrap ? (hwhich ? 0 : 6) : 3) + (he ? (hx_ == HX - 1 ? 2 : 1) : (hx_ == 0 ? 0 : 1));
                    if (cls != 4) ax += rb3_corr(P.cls, cls, rawm, rawc, rawn, ho + he);
                }
                const double nv = c + P.wod * (hb - ax);
                if (he) hm.y = nv; else hm.x = nv;
            }
            if (real) sts2(midw + ho, hm);
            hzm = he ? hr.x : hr.y;          // raw plane p at the position relaxed on plane p+1
            if (MODE != 3 && p + 1 >= z0 - 1 && p + 1 <= z1 && p + 1 + hshift >= 0 && p + 1 + hshift < P.NZ)
                hb = __ldg(P.b + (long long)(p + 1) * P.S2 + gbase + ho + (1 - he));
        }
        if (CLS && real) __syncthreads();     // the class taps of pass B read mid plane p at other threads' positions
        {
            // ---- pass B: the other colour of plane p-1 sits at the positions pass A just relaxed on plane p;
            // then rotate the slot's registers and fetch b of plane p+1 (in flight during the rest of the step, the
            // barrier and the next step's staging wait)
            const bool doB = (p - 1 >= z0);
            const bool bnext = (p + 1 >= max(z0 - 1, 0)) && (p + 1 <= min(z1, P.NZ - 1));
            double *outp = P.xo + (long long)(p - 1) * P.S2 + gbase;
            const double *bpn = P.b + (long long)(p + 1) * P.S2 + gbase;
#pragma unroll
            for (int k = 0; k < RB3_PPT; ++k) {
                if (k >= nslots) continue;
                const int o = ro + k * KS;
                if (doB) {
                    double2 oa = ma[k], ob = mb[k];
                    {
                        const double c = el(ma[k], EA);
                        const double xl = EA ? ma[k].x : midp[o - 1];
                        const double xr = EA ? midp[o + 2] : ma[k].y;
                        const double yn = midp[o - S1 + EA], ys = el(mb[k], EA);
                        double ax = P.d * c + P.c1 * (xl + xr) + P.cS * (yn + ys) + P.cP * (ua[k] + el(ra[k], EA));
                        if (CLS) {
                            const int cls = ((y0 + 2 * (pj + k * (NT / HX)) == 0) ? 0 : 3) +
                                            (EA ? (pi_ == HX - 1 ? 2 : 1) : (pi_ == 0 ? 0 : 1));
                            if (cls != 4) ax += rb3_corr(P.cls, cls, midpp, midp, midw, o + EA);
                        }
                        const double nv = c + P.wod * (ga[k] - ax);
                        if (EA) oa.y = nv; else oa.x = nv;
                    }
                    {
                        const double c = el(mb[k], EB);
                        const double xl = EB ? mb[k].x : midp[o + S1 - 1];
                        const double xr = EB ? midp[o + S1 + 2] : mb[k].y;
                        const double yn = el(ma[k], EB), ys = midp[o + 2 * S1 + EB];
                        double ax = P.d * c + P.c1 * (xl + xr) + P.cS * (yn + ys) + P.cP * (ub[k] + el(rb[k], EB));
                        if (CLS) {
                            const int cls = ((y0 + 2 * (pj + k * (NT / HX)) + 1 == P.NY - 1) ? 6 : 3) +
                                            (EB ? (pi_ == HX - 1 ? 2 : 1) : (pi_ == 0 ? 0 : 1));
                            if (cls != 4) ax += rb3_corr(P.cls, cls, midpp, midp, midw, o + S1 + EB);
                        }
                        const double nv = c + P.wod * (gb[k] - ax);
                        if (EB) ob.y = nv; else ob.x = nv;
                    }
                    *reinterpret_cast<double2 *>(outp + o) = oa;
                    *reinterpret_cast<double2 *>(outp + o + S1) = ob;
                }
                // pass B(p) relaxes the positions (row a, 1-EA), (row b, 1-EB)
                ua[k] = el(ma[k], 1 - EA);
                ub[k] = el(mb[k], 1 - EB);
                // (opaque moves: if ga/gb merely aliased halves of ba/bb, the b registers could not be reloaded in place
                // and the loop back-edge would have to move b values that are still in flight)
                if (MODE == 3) {
                    ga[k] = g3a[k];
                    gb[k] = g3b[k];
                } else {
                    asm volatile("mov.f64 %0, %1;" : "=d"(ga[k]) : "d"(el(ba[k], 1 - EA)));
                    asm volatile("mov.f64 %0, %1;" : "=d"(gb[k]) : "d"(el(bb[k], 1 - EB)));
                }
                ma[k] = ra[k];
                mb[k] = rb[k];
                if (MODE != 3 && bnext) {
                    ba[k] = ldg2(bpn + o);
                    bb[k] = ldg2(bpn + o + S1);
                }
            }
        }
        // (measured alternatives at 512^3, prolong + sweep 0.79 ms as is: adding e at the start of the step behind a
        // second barrier 0.89 ms; adding it to plane p+1 only here and giving the own-cell z-neighbours their e on the
        // fly 0.93 ms — the e values are then consumed a few hundred ns after their loads are issued)
        if (MODE == 2 && p + 2 >= pfirst && p + 2 <= plast) {
            wait_plane(p + 2);      // one plane ahead, so that the barrier below publishes the transformed plane
            transform(p + 2);
        }
        __syncthreads();        // raw plane p-KEEP and the oldest mid plane are free, mid plane p is complete
        if (tid == 0 && p - KEEP >= pfirst && p - KEEP + NS <= plast) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(p - KEEP + NS);       // into the slot of plane p-KEEP
        }
        if (!CLS && MODE != 3 && tid == 0 && P.bpf > 0) {      // (level 0; the CLS variants have no register to spare)
            // b of plane q (rows y0-1 .. y0+TY) is loaded at the end of step q-1
            const int q = p + 1 + P.bpf;
            if (q >= 0 && q <= min(z1, P.NZ - 1)) {
                const long long lo = (long long)q * P.S2 + gbase + S1, nn = (long long)P.NZ * P.S2;
                bulk_prefetch_l2(P.b, max(lo, 0ll), min(lo + (long long)(P.TY + 2) * S1, nn));
            }
        }
        if (!CLS && MODE == 2 && tid == 0 && P.epf > 0) {
            // e of raw plane q is fetched two steps before the plane is used: coarse rows y0/2 - 1 .. (y0 + TY)/2
            const int q = p + P.epf;
            if ((q & 1) == 0 && q >= 0 && q < P.NZ) {
                const long long nc = (long long)(P.NZ >> 1) * P.cs1 * P.cs2;
                const long long eo = ((long long)(q >> 1) * P.cs1 + (y0 >> 1) - 1) * P.cs2;
                bulk_prefetch_l2(P.e, max(eo, 0ll), min(eo + (long long)((P.TY >> 1) + 2) * P.cs2, nc));
            }
        }
    };

    // start on a DG = 0 plane at or below z0-2: at least one fill step precedes the first relaxed plane z0-1
    int p = z0 - 2;
    if ((p ^ P.c0) & 1) --p;
    if (MODE == 2) {
        for (int q = p; q <= p + 1; ++q)
            if (q >= pfirst && q <= plast) {
                wait_plane(q);
                transform(q);
            }
        __syncthreads();
    } else if (p >= pfirst) {
        wait_plane(p);
    }
    for (; p <= z1; p += 2) {
        step(std::integral_constant<int, 0>{}, p);
        if (p + 1 <= z1) step(std::integral_constant<int, 1>{}, p + 1);
    }
}

static bool rb3_params(Level &L, Rb3 *P, bool *use_cls, bool need_regular) {
    const bool off = getenv("OMG_NO_RB3") != nullptr;       // read per call: the parity test runs both paths
    if (off || L.kind == OMG_KIND_CSR || L.slab || L.band.nb != 6) return false;
    const bool cls = (L.kind == OMG_KIND_BAND_EXC);
    if (cls && !L.classed) return false;             // exception rows only if the kernel corrects them itself
    const BandOp &B = L.band;
    if (B.off[3] != 1 || B.off[2] != -1 || B.off[4] != -B.off[1] || B.off[5] != -B.off[0]) return false;
    if (B.coef[2] != B.coef[3] || B.coef[1] != B.coef[4] || B.coef[0] != B.coef[5]) return false;
    int S1 = B.off[4], S2 = B.off[5];
    if (S1 < 64 || (S1 & 63) || S2 % S1 != 0 || L.n % S2 != 0) return false;     // patch rows must be warp-uniform
    int NY = S2 / S1, NZ = L.n / S2;
    if ((NY & 1) || NY < 4 || NZ < 2) return false;
    if (L.colour.flat || L.colour.alpha != 3 || L.colour.s2 != S1 || L.colour.s1 != NY) return false;
    if (L.pad < S2 + 2 * S1) return false;        // plane -1 is staged from row y0-2
    if (need_regular && !(L.regular && L.reg.alpha == 3 && L.reg.fs2 == S1 && L.reg.fs1 == NY)) return false;
    const int NT = RB3_NT;
    const int NS = cls ? 4 : 3, NM = cls ? 3 : 2;
    int TY = 0;
    for (int t = std::min(env_int("OMG_RB3_TY", 64), NY); t >= 2; --t) {
        if ((t & 1) || NY % t != 0) continue;
        if ((t / 2) * (S1 / 2) > NT * RB3_PPT || ((t / 2) * (S1 / 2)) % NT != 0 || S1 > NT) continue;   // whole slots; one halo item per thread
        size_t smem = ((size_t)NS * (t + 4) + (size_t)NM * (t + 2)) * S1 * 8 + 64;
        if (smem > 227 * 1024) continue;
        TY = t;
        break;
    }
    if (TY < 2) return false;
    P->S1 = S1;
    P->S2 = S2;
    P->NY = NY;
    P->NZ = NZ;
    P->TY = TY;
    P->RS = (TY + 4) * S1;
    P->MS = (TY + 2) * S1;
    P->cs1 = NY / 2;
    P->cs2 = S1 / 2;
    P->c0 = 0;
    P->bpf = env_int("OMG_BPF", 2);
    P->epf = env_int("OMG_EPF", 4);
    P->d = B.diag;
    P->c1 = B.coef[3];
    P->cS = B.coef[4];
    P->cP = B.coef[5];
    P->wod = 1.0 / B.diag;
    *use_cls = cls;
    if (cls) P->cls = L.cls;
    {   // z-segments: 2 extra steps + ~1.5 planes of pipeline ramp per segment against whole waves of one CTA per SM
        int chunks = NY / TY, slots = std::max(g.sm_count, 1);
        int ZL = NZ;
        double best = -1.0;
        for (int nseg = 1; nseg <= NZ; ++nseg) {
            int zl = (NZ + nseg - 1) / nseg;
            int ns = (NZ + zl - 1) / zl;
            long long ctas = (long long)chunks * ns;
            long long waves = (ctas + slots - 1) / slots;
            if (waves > 16) break;
            double eff = (zl / (zl + 3.5)) * ((double)ctas / (double)(waves * slots));
            if (eff > best + 1e-9) {
                best = eff;
                ZL = zl;
            }
        }
        int envZL = env_int("OMG_RB3_ZL", 0);
        if (envZL >= 1) ZL = envZL;
        P->ZL = std::min(std::max(ZL, 1), NZ);
    }
    return true;
}

template <int MODE>
static bool rb3_launch(omg_hierarchy *h, const Rb3 &P, bool cls) {
    static bool attr_set[2] = {false, false};
    const int NS = cls ? 4 : 3, NM = cls ? 3 : 2;
    size_t smem = ((size_t)NS * P.RS + (size_t)NM * P.MS) * sizeof(double) + 64;
    void (*kern)(const Rb3) = cls ? k_rb3<MODE, true> : k_rb3<MODE, false>;
    if (!attr_set[cls]) {        // (static per MODE instantiation)
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr_set[cls] = true;
    }
    dist_halo_wait(h);
    kern<<<dim3(P.NY / P.TY, (P.NZ + P.ZL - 1) / P.ZL), RB3_NT, smem, g.stream>>>(P);
    return true;
}

// one full two-colour Gauss-Seidel sweep (colour 0, then colour 1) in a single pass.  xi == nullptr: probe.
bool stencil_rb_sweep(omg_hierarchy *h, Level &L, const double *xi, const double *b, double *xo) {
    St2 Q{};
    Rb3 T{};
    bool cls3;
    if (rb3_params(L, &T, &cls3, false)) {
        if (!xi) return true;
        T.xi = xi;
        T.b = b;
        T.xo = xo;
        return rb3_launch<0>(h, T, cls3);
    }
    if (!st2rb_params(L, &Q, false)) return false;
    if (!xi) return true;
    Q.xi = xi;
    Q.b = b;
    Q.xo = xo;
    Q.wod = 1.0 / Q.d;
    Q.colour = 0;
    return st2rb_launch<0>(h, Q);
}

// one full two-colour sweep from the zero iterate: xo = rbgs(A, b, 0), x is never read.  b == nullptr: probe.
bool stencil_rb_sweep0(omg_hierarchy *h, Level &L, const double *b, double *xo) {
    Rb3 T{};
    bool cls3;
    if (!rb3_params(L, &T, &cls3, false)) return false;
    if (!b) return true;
    T.xi = b;
    T.b = b;
    T.xo = xo;
    return rb3_launch<3>(h, T, cls3);
}

// y = xi + R^T e, then one full two-colour sweep on y, in a single pass.  xi == nullptr: probe.
bool stencil_prolong_rb_sweep(omg_hierarchy *h, Level &L, Level &C, const double *xi, const double *e,
                              const double *b, double *xo) {
    (void)C;
    St2 Q{};
    Rb3 T{};
    bool cls3;
    if (rb3_params(L, &T, &cls3, true)) {
        if (!xi) return true;
        T.xi = xi;
        T.b = b;
        T.xo = xo;
        T.e = e;
        T.w = L.Rw;
        return rb3_launch<2>(h, T, cls3);
    }
    if (!st2rb_params(L, &Q, true)) return false;
    if (!xi) return true;
    Q.xi = xi;
    Q.b = b;
    Q.xo = xo;
    Q.e = e;
    Q.w = L.Rw;
    Q.wod = 1.0 / Q.d;
    Q.colour = 0;
    return st2rb_launch<2>(h, Q);
}

// ================================================================ 3-D Jacobi sweep + residual + restriction in one pass
//
// The last pre-smoothing sweep of a level and the restricted residual that follows it (openmg/__init__.py:201 and
// :209-210) in ONE pass over x: x and b are read once, the new iterate and the coarse right-hand side written once —
// 24 n + 8 n_c bytes instead of 24 n for the sweep plus 16 n + 8 n_c for the residual.  Temporal blocking in z on
// the skeleton of k_rb3 (full-row chunks of TY rows, one TMA bulk copy per raw plane with two halo rows per side,
// 3-stage mbarrier ring, one __syncthreads per plane):
//   pass A(p)   Jacobi on EVERY point of plane p from the RAW planes p-1, p, p+1  -> "mid" plane p in shared memory,
//               and, for the planes the segment owns, the new iterate in global memory
//   pass B(p-1) residual of plane p-1 from the MID planes p-2, p-1, p, summed over the 2x2 patch and over the plane
//               pair of the aggregate                                             -> one coarse value per patch
// Pass B needs mid one row beyond the chunk and one plane beyond the segment: pass A also runs on the rows y0-1 and
// y0+TY (one x-pair per thread) and on the planes z0-1 and z1, recomputed with the same expression as the owner's and
// never stored.
// The restriction only wants the SUM of the residual over an aggregate, and a 2x2 patch on even rows/columns is one
// plane of an aggregate, so pass B never forms a point residual:
//   sum_patch (A x)(p-1) = (d + c1 + cS) S(p-1) + c1 (left + right columns) + cS (row above + row below)
//                          + cP (S(p-2) + S(p)),          S(q) = sum of the thread's own patch of mid plane q
// — two scalars carried in registers per patch, 2 LDS.128 + 4 LDS.64 from mid plane p-1, a dozen flops.  The result
// differs from the point-wise evaluation by rounding only (different summation order).
// Flat-index semantics (openmg/operators.py:244-256) as in k_rb3: full rows make the x-wraps contiguous, halo rows of
// the first / last chunk are rows of the neighbouring plane and are relaxed iff the point they really are lies in
// [0,n); everything outside [0,n) is the zero pad and stays zero in mid.
// Pure band levels only (level 0 of the Poisson hierarchies — the only level whose pre-smoothing does not start from
// the zero iterate when nu1 = 1), unsharded.
#define JR3_NT 512
#define JR3_PPT 2

struct Jr3 {
    const double *xi;
    const double *b;
    double *xo;
    double *rc;
    int S1, S2, NY, NZ;    // row length, plane size, rows per plane, planes
    int TY, ZL;            // rows per chunk, planes per z-segment (even)
    int RS, MS;            // staged raw plane (TY+4 rows) / mid plane (TY+2 rows) in doubles
    int cs1, cs2;          // coarse rows per plane, coarse row length
    int bpf;               // planes by which thread 0 prefetches b into L2 ahead of its loads (0: off)
    double d, c1, cS, cP, wod, w, dsum;     // wod = omega/d, dsum = d + c1 + cS
};

__device__ __forceinline__ double jr3_relax(const Jr3 &P, double c, double l, double r, double n, double s, double zm,
                                            double zp, double b) {
    const double ax = P.d * c + P.c1 * (l + r) + P.cS * (n + s) + P.cP * (zm + zp);
    return c + P.wod * (b - ax);
}

__global__ void __launch_bounds__(JR3_NT, 1) k_jr3(const Jr3 P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NT = JR3_NT, NS = 3, NM = 2;
    constexpr int KS = 4 * NT;             // slot k+1 is NT patches = 2 NT / HX patch rows = 4 NT doubles further on
    const int S1 = P.S1, HX = S1 >> 1;
    const int RS = P.RS, MS = P.MS;
    double *raw = reinterpret_cast<double *>(smem_raw);
    double *mid = raw + (size_t)NS * RS;
    uint64_t *full = reinterpret_cast<uint64_t *>(mid + NM * (size_t)MS);
    const int tid = threadIdx.x;
    const int y0 = (int)blockIdx.x * P.TY;
    const int z0 = (int)blockIdx.y * P.ZL;
    const int z1 = min(z0 + P.ZL, P.NZ);
    const uint32_t span_bytes = (uint32_t)RS * 8u;
    const int pfirst = max(z0 - 2, -1);           // raw planes below -1 / above NZ are all zero and never staged
    const int plast = min(z1 + 1, P.NZ);
    const long long gbase = (long long)(y0 - 2) * S1;      // staged offset o <-> in-plane offset gbase + o

    auto slot_of = [&](int p) { return (p - pfirst) % NS; };
    auto issue = [&](int p) {
        int s_ = slot_of(p);
        mbar_expect_tx(full + s_, span_bytes);
        bulk_g2s(raw + (size_t)s_ * RS, P.xi + (long long)p * P.S2 + gbase, span_bytes, full + s_);
    };
    auto wait_plane = [&](int p) {
        int k = p - pfirst;
        mbar_wait(full + (k % NS), (uint32_t)((k / NS) & 1));
    };

    if (tid == 0) {
        for (int s_ = 0; s_ < NS; ++s_) mbar_init(full + s_, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int p = pfirst; p < pfirst + NS && p <= plast; ++p) issue(p);

    // work items as in k_rb3: slot k of thread tid is the 2x2 patch tid + k NT of the chunk (passes A and B); the halo
    // item is one x-pair of the grid row y0-1 (threads [0,HX)) or y0+TY (threads [HX,2HX)), pass A only.
    const int nslots = ((P.TY >> 1) * HX) / NT;          // 1 or 2 (host: TY/2 * HX is a multiple of NT)
    const int pj = tid / HX, pi_ = tid - pj * HX;
    const int ro = (2 * pj + 2) * S1 + 2 * pi_;          // row a of slot 0 inside a staged raw plane; mid: ro - S1
    const bool hact = tid < 2 * HX;
    const bool hwhich = tid >= HX;
    const int ho = (hwhich ? (P.TY + 2) * S1 + 2 * (tid - HX) : S1 + 2 * tid);
    const bool hwrap = hwhich ? (y0 + P.TY == P.NY) : (y0 == 0);
    const int hshift = hwrap ? (hwhich ? 1 : -1) : 0;     // the plane the halo row really belongs to: p + hshift

    // carried from plane to plane, per slot: the raw patches of the planes p-1 and p (z- neighbours and centres of pass
    // A: the patch of plane p+1 is read from its staged plane once and moves down), the b patch of plane p (fetched one
    // step ahead), and for pass B the patch sum S(p-1), the part of the residual sum of plane p-1 that was known when
    // its mid patch was formed, and the residual sum of the open aggregate.
    // The shared-memory pipe is what limits this kernel (ncu, first version, which re-read the patch of plane p from its
    // staged plane: l1tex 72 %, short-scoreboard stalls on top): carrying it took the level-0 pass at 512^3 from 0.62 to
    // 0.58 ms.  Fetching the x-neighbours of a patch — the neighbouring lanes' own values — by warp shuffle instead of
    // the 2-way bank-conflicted LDS.64 was measured too and LOST (0.66 ms): the shuffles sit on the dependency chain in
    // front of the FMAs.
    double2 za[JR3_PPT], zb[JR3_PPT], ca[JR3_PPT], cb[JR3_PPT], ba[JR3_PPT], bb[JR3_PPT];
    double s1[JR3_PPT], part[JR3_PPT], acc[JR3_PPT];
    double2 hz = make_double2(0.0, 0.0), hc = hz, hb = hz;
#pragma unroll
    for (int k = 0; k < JR3_PPT; ++k) {
        za[k] = zb[k] = ca[k] = cb[k] = ba[k] = bb[k] = make_double2(0.0, 0.0);
        s1[k] = part[k] = acc[k] = 0.0;
    }

    if (z0 - 2 >= pfirst) {
        wait_plane(z0 - 2);
        const double *rawf = raw + (size_t)slot_of(z0 - 2) * RS;
#pragma unroll
        for (int k = 0; k < JR3_PPT; ++k) {
            if (k >= nslots) continue;
            ca[k] = lds2(rawf + ro + k * KS);
            cb[k] = lds2(rawf + ro + k * KS + S1);
        }
        if (hact) hc = lds2(rawf + ho);
    }
    // step z0-2 only fills the register pipeline (raw patches of the planes z0-2 and z0-1, b of plane z0-1)
    for (int p = z0 - 2; p <= z1; ++p) {
        const bool real = p >= z0 - 1;
        const bool relax = real && p >= 0 && p < P.NZ;
        const bool has_next = (p + 1 >= pfirst && p + 1 <= plast);
        if (has_next) wait_plane(p + 1);
        const double *rawc = raw + (size_t)slot_of(max(p, pfirst)) * RS;
        const double *rawn = raw + (size_t)slot_of(max(p + 1, pfirst)) * RS;
        const int mq = p - z0 + 2 + NM;                                     // mid plane p lives in slot mq % NM
        double *midw = mid + (size_t)(mq % NM) * MS - S1;                  // mid row = staged row - 1
        const double *midp = mid + (size_t)((mq - 1) % NM) * MS - S1;
        const bool owned = p >= z0 && p < z1;
        const bool doB = p - 1 >= z0;
        const bool bnext = (p + 1 >= max(z0 - 1, 0)) && (p + 1 <= min(z1, P.NZ - 1));
        double *outp = P.xo + (long long)p * P.S2 + gbase;
        const double *bpn = P.b + (long long)(p + 1) * P.S2 + gbase;
#pragma unroll
        for (int k = 0; k < JR3_PPT; ++k) {
            if (k >= nslots) continue;
            const int o = ro + k * KS;
            const double2 ra = ca[k], rb = cb[k];
            double2 pa = make_double2(0.0, 0.0), pb = pa;
            if (has_next) {
                pa = lds2(rawn + o);
                pb = lds2(rawn + o + S1);
            }
            double2 na = ra, nb = rb;
            if (relax) {
                // ---- pass A
                const double2 vn = lds2(rawc + o - S1), vs = lds2(rawc + o + 2 * S1);
                const double la = rawc[o - 1], rra = rawc[o + 2], lb = rawc[o + S1 - 1], rrb = rawc[o + S1 + 2];
                na.x = jr3_relax(P, ra.x, la, ra.y, vn.x, rb.x, za[k].x, pa.x, ba[k].x);
                na.y = jr3_relax(P, ra.y, ra.x, rra, vn.y, rb.y, za[k].y, pa.y, ba[k].y);
                nb.x = jr3_relax(P, rb.x, lb, rb.y, ra.x, vs.x, zb[k].x, pb.x, bb[k].x);
                nb.y = jr3_relax(P, rb.y, rb.x, rrb, ra.y, vs.y, zb[k].y, pb.y, bb[k].y);
                if (owned) {
                    *reinterpret_cast<double2 *>(outp + o) = na;
                    *reinterpret_cast<double2 *>(outp + o + S1) = nb;
                }
            }
            if (real) {
                // mid plane p (planes -1 and NZ: nothing was relaxed, it is the all-zero raw plane)
                sts2(midw + o, na);
                sts2(midw + o + S1, nb);
            }
            const double s0 = (na.x + na.y) + (nb.x + nb.y);
            if (doB) {
                // ---- pass B, second half: what the residual sum of plane p-1 still lacks — the rows above and below
                // the patch and the columns left and right of it (other threads' values: mid plane p-1), S(p)
                const double2 mn = lds2(midp + o - S1), ms = lds2(midp + o + 2 * S1);
                const double ml = midp[o - 1] + midp[o + S1 - 1], mr = midp[o + 2] + midp[o + S1 + 2];
                const double rest = P.c1 * (ml + mr) + P.cS * ((mn.x + mn.y) + (ms.x + ms.y)) + P.cP * s0;
                double a = acc[k] + (part[k] - rest);
                if ((p - 1) & 1) {
                    P.rc[((long long)((p - 1) >> 1) * P.cs1 + (y0 >> 1) + pj + k * (NT / HX)) * P.cs2 + pi_] = P.w * a;
                    a = 0.0;
                }
                acc[k] = a;
            }
            {
                // ---- pass B, first half, for plane p: everything known while its mid patch is in registers — the b
                // sum, the own patch, S(p-1)
                const double sbk = (ba[k].x + ba[k].y) + (bb[k].x + bb[k].y);
                part[k] = sbk - (P.dsum * s0 + P.cP * s1[k]);
            }
            s1[k] = s0;
            za[k] = ra;
            zb[k] = rb;
            ca[k] = pa;
            cb[k] = pb;
            if (bnext) {
                ba[k] = ldg2(bpn + o);
                bb[k] = ldg2(bpn + o + S1);
            }
        }
        if (hact) {
            // the halo item's row may belong to the neighbouring plane: relaxed iff that point lies inside [0,n)
            const double2 hr = hc;
            double2 hp = make_double2(0.0, 0.0);
            if (has_next) hp = lds2(rawn + ho);
            double2 hm = hr;
            if (real && p + hshift >= 0 && p + hshift < P.NZ) {
                const double2 hn = lds2(rawc + ho - S1), hs = lds2(rawc + ho + S1);
                const double hl = rawc[ho - 1], hrr = rawc[ho + 2];
                hm.x = jr3_relax(P, hr.x, hl, hr.y, hn.x, hs.x, hz.x, hp.x, hb.x);
                hm.y = jr3_relax(P, hr.y, hr.x, hrr, hn.y, hs.y, hz.y, hp.y, hb.y);
            }
            if (real) sts2(midw + ho, hm);
            hz = hr;
            hc = hp;
            if (p + 1 >= z0 - 1 && p + 1 <= z1 && p + 1 + hshift >= 0 && p + 1 + hshift < P.NZ)
                hb = ldg2(P.b + (long long)(p + 1) * P.S2 + gbase + ho);
        }
        __syncthreads();        // raw plane p and mid plane p-1 are free, mid plane p is complete
        if (tid == 0 && p >= pfirst && p + NS <= plast) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(p + NS);       // into the slot of plane p
        }
        if (tid == 0 && P.bpf > 0) {
            // b of plane q (rows y0-1 .. y0+TY) is loaded at the end of step q-1
            const int q = p + 1 + P.bpf;
            if (q >= 0 && q <= min(z1, P.NZ - 1)) {
                const long long lo = (long long)q * P.S2 + gbase + S1, nn = (long long)P.NZ * P.S2;
                bulk_prefetch_l2(P.b, max(lo, 0ll), min(lo + (long long)(P.TY + 2) * S1, nn));
            }
        }
    }
}

static bool jr3_params(Level &L, Jr3 *P) {
    if (getenv("OMG_NO_JR3") != nullptr) return false;      // read per call: the parity test runs both paths
    if (L.kind != OMG_KIND_BAND || L.slab || L.band.nb != 6) return false;
    const BandOp &B = L.band;
    if (B.off[3] != 1 || B.off[2] != -1 || B.off[4] != -B.off[1] || B.off[5] != -B.off[0]) return false;
    if (B.coef[2] != B.coef[3] || B.coef[1] != B.coef[4] || B.coef[0] != B.coef[5]) return false;
    int S1 = B.off[4], S2 = B.off[5];
    if (S1 < 64 || (S1 & 63) || S2 % S1 != 0 || L.n % S2 != 0) return false;     // patch rows must be warp-uniform
    int NY = S2 / S1, NZ = L.n / S2;
    if ((NY & 1) || NY < 4 || (NZ & 1) || NZ < 2) return false;
    if (L.pad < S2 + 2 * S1) return false;        // plane -1 is staged from row y0-2
    if (!(L.regular && L.reg.alpha == 3 && L.reg.fs2 == S1 && L.reg.fs1 == NY)) return false;
    const int NT = JR3_NT;
    int TY = 0;
    for (int t = std::min(env_int("OMG_JR3_TY", 64), NY); t >= 2; --t) {
        if ((t & 1) || NY % t != 0) continue;
        if ((t / 2) * (S1 / 2) > NT * JR3_PPT || ((t / 2) * (S1 / 2)) % NT != 0 || S1 > NT) continue;   // whole slots; one halo item per thread
        size_t smem = ((size_t)3 * (t + 4) + (size_t)2 * (t + 2)) * S1 * 8 + 64;
        if (smem > 227 * 1024) continue;
        TY = t;
        break;
    }
    if (TY < 2) return false;
    P->S1 = S1;
    P->S2 = S2;
    P->NY = NY;
    P->NZ = NZ;
    P->TY = TY;
    P->RS = (TY + 4) * S1;
    P->MS = (TY + 2) * S1;
    P->cs1 = NY / 2;
    P->cs2 = S1 / 2;
    P->d = B.diag;
    P->c1 = B.coef[3];
    P->cS = B.coef[4];
    P->cP = B.coef[5];
    P->dsum = B.diag + B.coef[3] + B.coef[4];
    P->bpf = env_int("OMG_BPF", 2);
    {   // z-segments of even length: 3 extra steps + ~1.5 planes of pipeline ramp per segment against whole waves of
        // one CTA per SM
        int chunks = NY / TY, slots = std::max(g.sm_count, 1);
        int ZL = NZ;
        double best = -1.0;
        for (int nseg = 1; nseg <= NZ / 2; ++nseg) {
            int zl = (NZ + nseg - 1) / nseg;
            zl += zl & 1;
            int ns = (NZ + zl - 1) / zl;
            long long ctas = (long long)chunks * ns;
            long long waves = (ctas + slots - 1) / slots;
            if (waves > 16) break;
            double eff = (zl / (zl + 4.5)) * ((double)ctas / (double)(waves * slots));
            if (eff > best + 1e-9) {
                best = eff;
                ZL = zl;
            }
        }
        int envZL = env_int("OMG_JR3_ZL", 0);
        if (envZL >= 2) ZL = envZL + (envZL & 1);
        P->ZL = std::min(std::max(ZL, 2), NZ);
    }
    return true;
}

// xo = xi + omega (b - A xi)/diag ; rc = R (b - A xo) — the last pre-smoothing sweep and the restricted residual in a
// single pass.  xi == nullptr: applicability probe.
bool stencil_jacobi_residual_restrict(omg_hierarchy *h, Level &L, Level &C, const double *xi, const double *b,
                                      double *xo, double *rcv, double omega) {
    (void)C;
    Jr3 P{};
    if (!jr3_params(L, &P)) {
        // 2-D / 1-D levels: k_jr2
        St2 Q{};
        if (getenv("OMG_NO_JR3") != nullptr || L.kind != OMG_KIND_BAND || L.slab) return false;
        if (!st2_params(L, &Q, true, true) || (Q.XW + 4) / 2 > ST2_NT * JR2_PP) return false;
        if (!xi) return true;
        Q.xi = xi;
        Q.b = b;
        Q.xo = xo;
        Q.rc = rcv + L.piece_row0;
        Q.w = L.Rw;
        Q.wod = omega / Q.d;
        static bool attr2[2] = {false, false};
        const bool hasn = Q.cN != 0.0;
        size_t smem2 = ((size_t)Q.NS * (Q.XW + 8) + (size_t)3 * (Q.XW + 4)) * sizeof(double) + 8 * sizeof(uint64_t);
        void (*kern)(const St2) = hasn ? k_jr2<true> : k_jr2<false>;
        if (!attr2[hasn]) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
            attr2[hasn] = true;
        }
        dist_halo_wait(h);
        kern<<<dim3(Q.XC, (Q.NY + Q.YL - 1) / Q.YL), ST2_NT, smem2, g.stream>>>(Q);
        return true;
    }
    if (!xi) return true;
    P.xi = xi;
    P.b = b;
    P.xo = xo;
    P.rc = rcv + L.piece_row0;
    P.w = L.Rw;
    P.wod = omega / P.d;
    static bool attr_set = false;
    size_t smem = ((size_t)3 * P.RS + (size_t)2 * P.MS) * sizeof(double) + 64;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_jr3, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr_set = true;
    }
    dist_halo_wait(h);
    k_jr3<<<dim3(P.NY / P.TY, (P.NZ + P.ZL - 1) / P.ZL), JR3_NT, smem, g.stream>>>(P);
    return true;
}
