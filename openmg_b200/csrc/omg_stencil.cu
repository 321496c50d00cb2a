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
//    an mbarrier; a ring of NS spans keeps planes z-1, z, z+1 resident and 2 more in flight.
//    The zero pads of the vectors make the global ends branch-free.
//  * Each thread owns 2x2 (x,y) patches: all 7 taps are shared-memory reads (LDS.128 for the
//    aligned pairs), b streams through ld.global.nc, results leave as 128-bit stores.
//  * The fused residual+restriction accumulates the 2x2 patch over two planes in registers and
//    writes one coarse value: the fine residual never exists in HBM.
//  * Rows deviating from the stencil ("exception rows" of Galerkin levels) are recomputed from
//    their compact CSR by a small fix-up kernel right after (out-of-place, so still exact).
#include "omg_stencil.cuh"
#include "omg_kernels.cuh"

#define ST_NT 512          // threads per CTA
#define ST_PPT 2           // 2x2 patches per thread per plane (max)
#define ST_NS 5            // ring stages

struct St3 {
    const double *xi;      // owned row 0 of the input vector (zero/halo padded)
    const double *b;
    double *xo;            // fine output (jacobi) or coarse output (residual+restrict)
    const double *e;       // coarse correction (prolong+jacobi)
    int S1, S2, NZ;        // row length, plane size, local planes
    int TY, NP, SPAN;      // rows per chunk, patches per plane-chunk, span length (TY+2)*S1
    int ZL;                // planes per z-segment (even)
    int cs1, cs2;          // coarse rows per plane, coarse row length
    int NYg;               // rows per plane
    int zg0, NZg;          // global index of local plane 0, global planes (prolong validity)
    int cz0;               // global index of local coarse plane 0
    double d, c1, cS, cP, wod, w;   // wod = omega/d
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

__device__ __forceinline__ double2 lds2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

// MODE 0: xo = xi + omega (b - A xi)/d
// MODE 1: rc = R (b - A xi)
// MODE 2: y = xi + R^T e ; xo = y + omega (b - A y)/d
template <int MODE>
__global__ void __launch_bounds__(ST_NT, 1) k_st3(const St3 P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)ST_NS * P.SPAN * sizeof(double));
    const int tid = threadIdx.x;
    const int z0 = blockIdx.y * P.ZL;
    const int z1 = min(z0 + P.ZL, P.NZ);
    const int y0 = blockIdx.x * P.TY;
    const long long span0 = (long long)y0 * P.S1 - P.S1;      // in-plane start of the span (row y0-1)
    const uint32_t span_bytes = (uint32_t)P.SPAN * 8u;

    if (tid == 0) {
        for (int s = 0; s < ST_NS; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < ST_NS; ++k) {
            int p = z0 - 1 + k;
            if (p > z1) break;
            mbar_expect_tx(full + k, span_bytes);
            bulk_g2s(stage + (size_t)k * P.SPAN, P.xi + (long long)p * P.S2 + span0, span_bytes, full + k);
        }
    }

    // fixed patch assignment
    const int HX = P.S1 >> 1;
    int px[ST_PPT], py[ST_PPT];
    bool act[ST_PPT];
#pragma unroll
    for (int k = 0; k < ST_PPT; ++k) {
        int p = tid + k * ST_NT;
        act[k] = p < P.NP;
        p = act[k] ? p : 0;
        py[k] = p / HX;
        px[k] = p - py[k] * HX;
    }
    double acc[ST_PPT];
#pragma unroll
    for (int k = 0; k < ST_PPT; ++k) acc[k] = 0.0;

    for (int z = z0; z < z1; ++z) {
        const int q = z - (z0 - 1);                // ring position of plane z (plane z0-1 is 0)
        // b for this plane: issue the loads before blocking on the barrier
        double2 ba[ST_PPT], bb[ST_PPT];
#pragma unroll
        for (int k = 0; k < ST_PPT; ++k) {
            if (act[k]) {
                int gi = z * P.S2 + (y0 + 2 * py[k]) * P.S1 + 2 * px[k];
                ba[k] = ldg2(P.b + gi);
                bb[k] = ldg2(P.b + gi + P.S1);
            }
        }
        if (z == z0) {
            mbar_wait(full + 0, 0);
            mbar_wait(full + 1, 0);
        }
        {
            int qq = q + 1;
            mbar_wait(full + (qq % ST_NS), (uint32_t)((qq / ST_NS) & 1));
        }
        const double *sm = stage + (size_t)((q - 1) % ST_NS) * P.SPAN;
        const double *sc = stage + (size_t)(q % ST_NS) * P.SPAN;
        const double *sp = stage + (size_t)((q + 1) % ST_NS) * P.SPAN;
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
            if (MODE == 2) {
                // y = x + w e[agg]: add the coarse correction to every tap (flat-index neighbours)
                const int Y = (y0 >> 1) + py[k];                         // coarse row in plane
                const int zg = z + P.zg0;                                // global fine plane
                const int Zc = (zg >> 1) - P.cz0;                        // local coarse plane of z
                const double *ec = P.e + ((long long)Zc * P.cs1 + Y) * P.cs2;
                const double w = P.w;
                // rows a and b (same coarse row), centre and in-row neighbours
                const int X = px[k];
                double e0 = __ldg(ec + X);
                // north row ya-1: flat row index rho-1 (may fall into the previous plane)
                const int ya = y0 + 2 * py[k];
                double en = 0.0, es = 0.0, enl = 0.0, esr = 0.0;
                {
                    // row ya-1: same plane, or the last row of plane zg-1 (flat-index wrap)
                    int zn = ya > 0 ? zg : zg - 1;
                    int yn = ya > 0 ? ya - 1 : P.NYg - 1;
                    if (zn >= 0) {
                        const double *er = P.e + ((long long)((zn >> 1) - P.cz0) * P.cs1 + (yn >> 1)) * P.cs2;
                        en = __ldg(er + X);
                        enl = __ldg(er + P.cs2 - 1);
                    }
                    // row yb+1 = ya+2: same plane, or the first row of plane zg+1
                    int zs = ya + 2 < P.NYg ? zg : zg + 1;
                    int ys = ya + 2 < P.NYg ? ya + 2 : 0;
                    if (zs < P.NZg) {
                        const double *er = P.e + ((long long)((zs >> 1) - P.cz0) * P.cs1 + (ys >> 1)) * P.cs2;
                        es = __ldg(er + X);
                        esr = __ldg(er + 0);
                    }
                }
                // left/right neighbours: same coarse row unless at the row ends (flat wrap)
                double ela = (X > 0) ? __ldg(ec + X - 1) : enl;            // left of (ya, x=0) is (ya-1, S1-1)
                double elb = (X > 0) ? ela : __ldg(ec + P.cs2 - 1);        // left of (yb, x=0) is (ya, S1-1)
                double erb = (X < P.cs2 - 1) ? __ldg(ec + X + 1) : esr;    // right of (yb, S1-1) is (yb+1, 0)
                double era = (X < P.cs2 - 1) ? erb : __ldg(ec + 0);        // right of (ya, S1-1) is (yb, 0)
                // z neighbours
                double em = 0.0, ep = 0.0;
                if (zg - 1 >= 0) em = __ldg(P.e + ((long long)(((zg - 1) >> 1) - P.cz0) * P.cs1 + Y) * P.cs2 + X);
                if (zg + 1 < P.NZg) ep = __ldg(P.e + ((long long)(((zg + 1) >> 1) - P.cz0) * P.cs1 + Y) * P.cs2 + X);
                va.x += w * e0; va.y += w * e0; vb.x += w * e0; vb.y += w * e0;
                vn.x += w * en; vn.y += w * en; vs.x += w * es; vs.y += w * es;
                ma.x += w * em; ma.y += w * em; mb.x += w * em; mb.y += w * em;
                pa.x += w * ep; pa.y += w * ep; pb.x += w * ep; pb.y += w * ep;
                la += w * ela; lb += w * elb; ra += w * era; rb += w * erb;
            }
            double ax0 = P.d * va.x + P.c1 * (la + va.y) + P.cS * (vn.x + vb.x) + P.cP * (ma.x + pa.x);
            double ax1 = P.d * va.y + P.c1 * (va.x + ra) + P.cS * (vn.y + vb.y) + P.cP * (ma.y + pa.y);
            double ax2 = P.d * vb.x + P.c1 * (lb + vb.y) + P.cS * (va.x + vs.x) + P.cP * (mb.x + pb.x);
            double ax3 = P.d * vb.y + P.c1 * (vb.x + rb) + P.cS * (va.y + vs.y) + P.cP * (mb.y + pb.y);
            if (MODE == 1) {
                double a = acc[k];
                a += ba[k].x - ax0;
                a += ba[k].y - ax1;
                a += bb[k].x - ax2;
                a += bb[k].y - ax3;
                if (z & 1) {
                    int Z = z >> 1;
                    P.xo[((long long)Z * P.cs1 + (y0 >> 1) + py[k]) * P.cs2 + px[k]] = P.w * a;
                    a = 0.0;
                }
                acc[k] = a;
            } else {
                int gi = z * P.S2 + (y0 + 2 * py[k]) * P.S1 + 2 * px[k];
                double2 oa2, ob2;
                oa2.x = va.x + P.wod * (ba[k].x - ax0);
                oa2.y = va.y + P.wod * (ba[k].y - ax1);
                ob2.x = vb.x + P.wod * (bb[k].x - ax2);
                ob2.y = vb.y + P.wod * (bb[k].y - ax3);
                *reinterpret_cast<double2 *>(P.xo + gi) = oa2;
                *reinterpret_cast<double2 *>(P.xo + gi + P.S1) = ob2;
            }
        }
        __syncthreads();        // everyone is done with plane z-1's slot
        if (tid == 0) {
            int p = z - 1 + ST_NS;
            if (p <= z1) {
                int k = (q - 1 + ST_NS) % ST_NS;      // == slot of plane z-1
                mbar_expect_tx(full + k, span_bytes);
                bulk_g2s(stage + (size_t)k * P.SPAN, P.xi + (long long)p * P.S2 + span0, span_bytes, full + k);
            }
        }
    }
}

// ---------------------------------------------------------------- fix-ups for exception rows

// xo_i for exception rows, MODE 0 (jacobi) and MODE 2 (prolong + jacobi, y = x + R^T e on the fly)
template <int MODE>
__global__ void __launch_bounds__(OMG_TPB) k_fix_rows(const int *__restrict__ rows, int nexc, ExcOp E, RegR R,
                                                      int crow0, int frow0, int nglob,
                                                      const double *__restrict__ xi, const double *__restrict__ e,
                                                      const double *__restrict__ b, double *__restrict__ xo,
                                                      double omega) {
    int s = blockIdx.x * OMG_TPB + threadIdx.x;
    if (s >= nexc) return;
    int i = rows[s];
    int p0 = E.ptr[s], p1 = E.ptr[s + 1];
    double acc = 0.0;
    for (int p = p0; p < p1; ++p) {
        int j = E.col[p];
        double v = xi[j];
        if (MODE == 2) {
            int jg = j + frow0;
            if (jg >= 0 && jg < nglob) v += R.w * __ldg(e + reg_agg(R, jg) - crow0);
        }
        acc += E.val[p] * v;
    }
    double xc = xi[i];
    if (MODE == 2) xc += R.w * __ldg(e + reg_agg(R, i + frow0) - crow0);
    xo[i] = xc + omega * (b[i] - acc) / E.diag[s];
}

// rc_I for coarse rows whose aggregate contains an exception row: the generic fused formula
__global__ void __launch_bounds__(OMG_TPB) k_fix_crows(const int *__restrict__ crows, int ncrows, BandA<1> A, RegR R,
                                                       int crow0, int frow0, const double *__restrict__ x,
                                                       const double *__restrict__ b, double *__restrict__ rc) {
    int t = blockIdx.x * OMG_TPB + threadIdx.x;
    if (t >= ncrows) return;
    int I = crows[t];
    int cc = reg_cc(R, I + crow0) - frow0;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < R.k) {
            int i = cc + R.o[k];
            double d;
            double ax = A.ax(i, x, d);
            acc += __ldg(b + i) - ax;
        }
    }
    rc[I] = R.w * acc;
}

// ---------------------------------------------------------------- host side

// Does level L carry a 3-D 7-point constant band this path can run?
static bool st3_params(Level &L, St3 *P) {
    if (L.kind == OMG_KIND_CSR || L.band.nb != 6) return false;
    const BandOp &B = L.band;
    if (B.off[3] != 1 || B.off[2] != -1 || B.off[4] != -B.off[1] || B.off[5] != -B.off[0]) return false;
    if (B.coef[2] != B.coef[3] || B.coef[1] != B.coef[4] || B.coef[0] != B.coef[5]) return false;
    int S1 = B.off[4], S2 = B.off[5];
    if (S1 < 64 || S1 > 2048 || (S1 & 1) || S2 % S1 != 0 || L.nloc % S2 != 0 || L.row0 % S2 != 0) return false;
    int NY = S2 / S1, NZ = L.nloc / S2;
    if ((NY & 1) || (NZ & 1) || NZ < 2) return false;
    if (L.pad < S2 + S1) return false;
    // rows per chunk: even, divides NY, TY*S1/4 patches <= ST_NT*ST_PPT
    int maxTY = (4 * ST_NT * ST_PPT) / S1;
    int TY = 0;
    for (int t = std::min(maxTY, NY); t >= 2; --t)
        if ((t % 2 == 0) && NY % t == 0) {
            TY = t;
            break;
        }
    if (TY < 2) return false;
    P->S1 = S1;
    P->S2 = S2;
    P->NZ = NZ;
    P->TY = TY;
    P->NP = TY * S1 / 4;
    P->SPAN = (TY + 2) * S1;
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
    // z-segments: even length, aim at >= 4 waves of CTAs
    int chunks = NY / TY;
    int target = std::max(1, (4 * std::max(g.sm_count, 1) + chunks - 1) / chunks);
    int ZL = std::max(2, (NZ + target - 1) / target);
    ZL += ZL & 1;
    ZL = std::min(ZL, NZ);
    P->ZL = ZL;
    return true;
}

static size_t st3_smem(const St3 &P) { return (size_t)ST_NS * P.SPAN * sizeof(double) + ST_NS * sizeof(uint64_t); }

template <int MODE>
static bool st3_launch(const St3 &P) {
    static bool attr_set = false;
    size_t smem = st3_smem(P);
    if (smem > 227 * 1024) return false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_st3<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr_set = true;
    }
    dim3 grid(P.NYg / P.TY, (P.NZ + P.ZL - 1) / P.ZL);
    k_st3<MODE><<<grid, ST_NT, smem, g.stream>>>(P);
    return true;
}

static bool regular_matches(const Level &L, const St3 &P) {
    return L.regular && L.reg.alpha == 3 && L.reg.fs2 == P.S1 && L.reg.fs1 == P.NYg;
}

bool stencil_jacobi(omg_hierarchy *h, Level &L, const double *xi, const double *b, double *xo, double omega) {
    (void)h;
    St3 P{};
    if (!st3_params(L, &P)) return false;
    if (L.kind == OMG_KIND_BAND_EXC && !L.exc_rows) return false;
    P.xi = xi;
    P.b = b;
    P.xo = xo;
    P.wod = omega / P.d;
    if (!st3_launch<0>(P)) return false;
    if (L.kind == OMG_KIND_BAND_EXC)
        k_fix_rows<0><<<cdiv(L.nexc, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.exc_rows, (int)L.nexc, L.exc_op(), L.reg, 0,
                                                                       L.row0, L.n, xi, nullptr, b, xo, omega);
    return true;
}

bool stencil_residual_restrict(omg_hierarchy *h, Level &L, Level &C, const double *x, const double *b, double *rc) {
    (void)h;
    St3 P{};
    if (!st3_params(L, &P) || !regular_matches(L, P)) return false;
    if (L.kind == OMG_KIND_BAND_EXC && !L.exc_crows) return false;
    P.xi = x;
    P.b = b;
    P.xo = rc;
    P.w = L.Rw;
    if (!st3_launch<1>(P)) return false;
    if (L.kind == OMG_KIND_BAND_EXC && L.nexc_crows > 0) {
        BandA<1> A{L.band, L.exc_op()};
        k_fix_crows<<<cdiv(L.nexc_crows, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.exc_crows, L.nexc_crows, A, L.reg, C.row0,
                                                                          L.row0, x, b, rc);
    }
    return true;
}

bool stencil_prolong_jacobi(omg_hierarchy *h, Level &L, Level &C, const double *xi, const double *e,
                            const double *b, double *xo, double omega) {
    (void)h;
    St3 P{};
    if (!st3_params(L, &P) || !regular_matches(L, P)) return false;
    if (L.kind == OMG_KIND_BAND_EXC && !L.exc_rows) return false;
    P.xi = xi;
    P.b = b;
    P.xo = xo;
    P.e = e;
    P.w = L.Rw;
    P.wod = omega / P.d;
    if (!st3_launch<2>(P)) return false;
    if (L.kind == OMG_KIND_BAND_EXC)
        k_fix_rows<2><<<cdiv(L.nexc, OMG_TPB), OMG_TPB, 0, g.stream>>>(L.exc_rows, (int)L.nexc, L.exc_op(), L.reg,
                                                                       C.row0, L.row0, L.n, xi, e, b, xo, omega);
    return true;
}
