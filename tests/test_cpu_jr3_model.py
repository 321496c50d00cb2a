"""CPU model of `k_jr3` (openmg_b200/csrc/omg_stencil.cu): the last pre-smoothing Jacobi sweep and the restricted
residual of a 3-D level (openmg/__init__.py:201,209-210) in one pass over x.

The CUDA kernel marches in z over full-row chunks: pass A(p) relaxes every point of plane p from the raw planes
p-1, p, p+1 into a "mid" plane (and stores the planes the segment owns), pass B(p-1) forms the PATCH SUM of the
residual of plane p-1 from the mid planes p-2, p-1, p —
    sum_patch (A x) = (d + c1 + cS) S(p-1) + c1 (left + right columns) + cS (row above + row below) + cP (S(p-2) + S(p))
— in two halves (what is known while the mid patch of plane p-1 is still in registers; the rest one step later from
shared memory) and accumulates it over the plane pair of the aggregate.  The raw patches of the planes p-1 and p are
carried in registers.  Halo rows y0-1 / y0+TY and halo planes z0-1 / z1 are
recomputed, the 3-stage raw ring and the 2-plane mid ring are reused, rows outside [0,NY) are rows of the neighbouring
plane (flat-index semantics, openmg/operators.py:244-256).  This file restates exactly that index logic thread by
thread (threads run one after another between barriers; ring residency is asserted) and checks it against the oracle's
Jacobi sweep and R (b - A x), so the algorithm is pinned on the CPU; the GPU parity tests pin the kernel itself.
"""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle.openmg_oracle as orc


def jr3_model(NZ, NY, S1, NT, TY, ZL, seed=0):
    shape = (S1, NY, NZ)
    A = sp.csr_matrix(orc.poisson_csr(shape))
    n = A.shape[0]
    S2 = NY * S1
    rs = np.random.RandomState(seed)
    x, b = rs.random_sample(n), rs.random_sample(n)
    omega = 0.8
    i = S2 + S1 + 1
    d, c1, cS, cP = A[i, i], A[i, i + 1], A[i, i + S1], A[i, i + S2]
    wod = omega / d
    dsum = d + c1 + cS
    R = orc.restrictionList(shape, 1, 2)[0]
    w = R.data[0]
    pad = S2 + 2 * S1
    X = np.zeros(n + 2 * pad)
    X[pad:pad + n] = x
    B = np.zeros(n + 2 * pad)
    B[pad:pad + n] = b
    XO = np.zeros(n + 2 * pad)
    cs1, cs2 = NY // 2, S1 // 2
    RC = np.full(n // 8, np.nan)
    HX = S1 // 2
    RS, MS = (TY + 4) * S1, (TY + 2) * S1
    NS, NM, KS = 3, 2, 4 * NT
    nslots = ((TY // 2) * HX) // NT
    assert nslots in (1, 2) and ((TY // 2) * HX) % NT == 0 and S1 <= NT and NT % HX == 0

    def relaxf(c, l, r, nn, s, zm, zp, bv):
        ax = d * c + c1 * (l + r) + cS * (nn + s) + cP * (zm + zp)
        return c + wod * (bv - ax)

    Z2 = lambda: np.zeros(2)
    for by in range((NZ + ZL - 1) // ZL):
        for bx in range(NY // TY):
            y0, z0 = bx * TY, by * ZL
            z1 = min(z0 + ZL, NZ)
            pfirst, plast = max(z0 - 2, -1), min(z1 + 1, NZ)
            gbase = (y0 - 2) * S1
            raw = {}            # resident staged planes (the ring: at most NS, refilled after the barrier)

            def stage(q):
                assert len(raw) < NS
                raw[q] = X[pad + q * S2 + gbase: pad + q * S2 + gbase + RS].copy()
            for q in range(pfirst, min(pfirst + NS, plast + 1)):
                stage(q)
            mid = [np.full(MS, np.nan) for _ in range(NM)]
            st = [dict(za=[Z2(), Z2()], zb=[Z2(), Z2()], ca=[Z2(), Z2()], cb=[Z2(), Z2()], ba=[Z2(), Z2()],
                       bb=[Z2(), Z2()], s1=[0.0] * 2, part=[0.0] * 2, acc=[0.0] * 2, hz=Z2(), hc=Z2(), hb=Z2())
                  for _ in range(NT)]
            geo = []
            for tid in range(NT):
                pj, pi_ = divmod(tid, HX)
                hwhich = tid >= HX
                ho = (TY + 2) * S1 + 2 * (tid - HX) if hwhich else S1 + 2 * tid
                hwrap = (y0 + TY == NY) if hwhich else (y0 == 0)
                geo.append((pj, pi_, (2 * pj + 2) * S1 + 2 * pi_, ho, ((1 if hwhich else -1) if hwrap else 0)))
            if z0 - 2 >= pfirst:
                rawf = raw[z0 - 2]
                for tid in range(NT):
                    ro, ho = geo[tid][2], geo[tid][3]
                    for k in range(nslots):
                        st[tid]['ca'][k] = rawf[ro + k * KS:ro + k * KS + 2].copy()
                        st[tid]['cb'][k] = rawf[ro + k * KS + S1:ro + k * KS + S1 + 2].copy()
                    if tid < 2 * HX:
                        st[tid]['hc'] = rawf[ho:ho + 2].copy()
            for p in range(z0 - 2, z1 + 1):
                real = p >= z0 - 1
                relax = real and 0 <= p < NZ
                has_next = pfirst <= p + 1 <= plast
                if relax:
                    assert p in raw
                if has_next:
                    assert p + 1 in raw
                rawc, rawn = raw.get(p), raw.get(p + 1)
                mq = p - z0 + 2 + NM
                midw, midp = mid[mq % NM], mid[(mq - 1) % NM]          # staged-row offsets: index - S1
                owned, doB = z0 <= p < z1, p - 1 >= z0
                bnext = (p + 1 >= max(z0 - 1, 0)) and (p + 1 <= min(z1, NZ - 1))
                for k in range(nslots):
                    new = []
                    for tid in range(NT):           # pass A
                        T = st[tid]
                        pj, pi_, ro, ho, hshift = geo[tid]
                        o = ro + k * KS
                        ra, rb = T['ca'][k], T['cb'][k]
                        pa = rawn[o:o + 2].copy() if has_next else Z2()
                        pb = rawn[o + S1:o + S1 + 2].copy() if has_next else Z2()
                        na, nb = ra.copy(), rb.copy()
                        if relax:
                            vn, vs = rawc[o - S1:o - S1 + 2], rawc[o + 2 * S1:o + 2 * S1 + 2]
                            la, lb, rra, rrb = rawc[o - 1], rawc[o + S1 - 1], rawc[o + 2], rawc[o + S1 + 2]
                            za, zb, ba, bb = T['za'][k], T['zb'][k], T['ba'][k], T['bb'][k]
                            na[0] = relaxf(ra[0], la, ra[1], vn[0], rb[0], za[0], pa[0], ba[0])
                            na[1] = relaxf(ra[1], ra[0], rra, vn[1], rb[1], za[1], pa[1], ba[1])
                            nb[0] = relaxf(rb[0], lb, rb[1], ra[0], vs[0], zb[0], pb[0], bb[0])
                            nb[1] = relaxf(rb[1], rb[0], rrb, ra[1], vs[1], zb[1], pb[1], bb[1])
                            if owned:
                                gi = pad + p * S2 + gbase + o
                                XO[gi:gi + 2] = na
                                XO[gi + S1:gi + S1 + 2] = nb
                        if real:
                            midw[o - S1:o - S1 + 2] = na
                            midw[o:o + 2] = nb
                        new.append((na, nb, pa, pb))
                    for tid in range(NT):           # pass B
                        T = st[tid]
                        pj, pi_, ro, ho, hshift = geo[tid]
                        o = ro + k * KS
                        na, nb, pa, pb = new[tid]
                        s0 = (na[0] + na[1]) + (nb[0] + nb[1])
                        if doB:
                            m = lambda off: midp[off - S1]
                            mn, ms = m(o - S1) + m(o - S1 + 1), m(o + 2 * S1) + m(o + 2 * S1 + 1)
                            ml, mr = m(o - 1) + m(o + S1 - 1), m(o + 2) + m(o + S1 + 2)
                            rest = c1 * (ml + mr) + cS * (mn + ms) + cP * s0
                            a = T['acc'][k] + (T['part'][k] - rest)
                            if (p - 1) & 1:
                                RC[(((p - 1) >> 1) * cs1 + (y0 >> 1) + pj + k * (NT // HX)) * cs2 + pi_] = w * a
                                a = 0.0
                            T['acc'][k] = a
                        sbk = (T['ba'][k][0] + T['ba'][k][1]) + (T['bb'][k][0] + T['bb'][k][1])
                        T['part'][k] = sbk - (dsum * s0 + cP * T['s1'][k])
                        T['s1'][k] = s0
                    for tid in range(NT):           # register rotation, b of plane p+1
                        T = st[tid]
                        o = geo[tid][2] + k * KS
                        T['za'][k], T['zb'][k] = T['ca'][k], T['cb'][k]
                        T['ca'][k], T['cb'][k] = new[tid][2], new[tid][3]
                        if bnext:
                            gi = pad + (p + 1) * S2 + gbase + o
                            T['ba'][k] = B[gi:gi + 2].copy()
                            T['bb'][k] = B[gi + S1:gi + S1 + 2].copy()
                hnew = {}
                for tid in range(min(NT, 2 * HX)):  # halo items
                    T = st[tid]
                    pj, pi_, ro, ho, hshift = geo[tid]
                    hr = T['hc']
                    hp = rawn[ho:ho + 2].copy() if has_next else Z2()
                    hm = hr.copy()
                    if real and 0 <= p + hshift < NZ:
                        hn, hs = rawc[ho - S1:ho - S1 + 2], rawc[ho + S1:ho + S1 + 2]
                        hl, hrr = rawc[ho - 1], rawc[ho + 2]
                        hm[0] = relaxf(hr[0], hl, hr[1], hn[0], hs[0], T['hz'][0], hp[0], T['hb'][0])
                        hm[1] = relaxf(hr[1], hr[0], hrr, hn[1], hs[1], T['hz'][1], hp[1], T['hb'][1])
                    if real:
                        midw[ho - S1:ho - S1 + 2] = hm
                    hnew[tid] = hp
                for tid in range(min(NT, 2 * HX)):
                    T = st[tid]
                    ho, hshift = geo[tid][3], geo[tid][4]
                    T['hz'], T['hc'] = T['hc'], hnew[tid]
                    if p + 1 >= z0 - 1 and p + 1 <= z1 and 0 <= p + 1 + hshift < NZ:
                        gi = pad + (p + 1) * S2 + gbase + ho
                        T['hb'] = B[gi:gi + 2].copy()
                # barrier; thread 0 refills the slot of plane p
                if p >= pfirst and p + NS <= plast:
                    raw.pop(p)
                    stage(p + NS)
    return A, R, x, b, omega, XO[pad:pad + n], RC


# (NZ, NY, S1, threads, rows per chunk, planes per segment): one and two patch slots per thread, one to several
# chunks and segments, a last segment shorter than the others.  (The reference's restriction numbers the coarse cells
# of shape (a, b, c) consistently with the band strides only when c == a, operators.py:63-84 — the only shapes the
# structured kernels accept.)
@pytest.mark.parametrize("cfg", [(16, 8, 16, 16, 4, 6), (16, 16, 16, 16, 8, 4), (16, 8, 16, 16, 8, 16),
                                 (16, 4, 16, 16, 4, 2), (16, 12, 16, 16, 4, 10), (8, 8, 8, 8, 8, 2)])
def test_jr3_model_matches_oracle(cfg):
    A, R, x, b, omega, xo, rc = jr3_model(*cfg)
    xw = orc.jacobi(A, b, x.copy(), 1, omega)
    rw = R.dot(b - A.dot(xw))
    assert not np.isnan(rc).any(), "every coarse row is written exactly by its owner"
    assert np.abs(xo - xw).max() <= 1e-14
    assert np.abs(rc - rw).max() <= 1e-13


def jr2_model(shape, XW, YL, NS=4, seed=0):
    """`k_jr2`, the row-marching 2-D / 1-D member (x-chunks of XW columns of y-segments of YL rows, one halo pair per
    side and one halo row per segment end recomputed, raw ring of NS rows, 3 mid rows); 1-D vectors are rows of N."""
    oned = len(shape) == 1
    A = sp.csr_matrix(orc.poisson_csr(shape, sparse_1d=oned))
    n = A.shape[0]
    if oned:
        N = XW * 2 if n % (XW * 2) == 0 else XW       # view as rows of N
    else:
        N = shape[0]
    NY = n // N
    rs = np.random.RandomState(seed)
    x, b = rs.random_sample(n), rs.random_sample(n)
    omega = 0.8
    i = 2 * N + 3 if not oned else 5
    d = A[i, i]; c1 = A[i, i + 1]
    cN = A[i, i + N] if not oned else 0.0
    cD = A[i, i + N + 1] if not oned else 0.0
    wod = omega / d
    R = orc.restrictionList(shape, 1, 2)[0]
    w = R.data[0]
    pad = 2 * N + 8
    X = np.zeros(n + 2 * pad); X[pad:pad + n] = x
    B = np.zeros(n + 2 * pad); B[pad:pad + n] = b
    XO = np.full(n, np.nan)
    cs = N // 2
    RC = np.full(R.shape[0], np.nan)
    RP, MP = XW + 8, XW + 4
    HXW, NPA, MPH = XW // 2, XW // 2 + 2, (XW + 4) // 2
    ntot = NY * N
    for by in range((NY + YL - 1) // YL):
        for bx in range(N // XW):
            x0, y0 = bx * XW, by * YL
            y1 = min(y0 + YL, NY)
            rlo, rhi, first = y0 - 2, y1 + 1, y0 - 1
            raw = {}
            def stage(r):
                assert len(raw) < NS
                raw[r] = X[pad + r * N + x0 - 4: pad + r * N + x0 - 4 + RP].copy()
            for r in range(rlo, min(rlo + NS, rhi + 1)): stage(r)
            mid = {}
            bold = {}; acc = {}
            for it in range(y0 - 1, y1 + 1):
                sm, sc, spp = raw[it - 1], raw[it], raw[it + 1]
                mw = np.full(MP, np.nan)
                bvs = {}
                for p in range(NPA):
                    bv = B[pad + it * N + x0 - 2 + 2 * p: pad + it * N + x0 - 2 + 2 * p + 2].copy()
                    bvs[p] = bv
                    o = 2 + 2 * p; xg = x0 - 2 + 2 * p; gi = it * N + xg
                    c = sc[o:o + 2].copy()
                    if 0 <= gi < ntot:
                        l, r_ = sc[o - 1], sc[o + 2]
                        q = spp[o:o + 2]
                        ml, m0, pr = sm[o - 1], sm[o], spp[o + 2]
                        ax0 = d * c[0] + c1 * (l + c[1]) + cD * (ml + q[1])
                        ax1 = d * c[1] + c1 * (c[0] + r_) + cD * (m0 + pr)
                        if cN != 0.0:
                            ax0 += cN * (m0 + q[0]); ax1 += cN * (sm[o + 1] + q[1])
                        c = np.array([c[0] + wod * (bv[0] - ax0), c[1] + wod * (bv[1] - ax1)])
                        if y0 <= it < y1 and 1 <= p <= HXW:
                            XO[gi:gi + 2] = c
                    mw[p] = c[0]; mw[MPH + p] = c[1]
                mid[it] = mw
                for k_ in list(mid):
                    if k_ < it - 2: mid.pop(k_)
                assert len(mid) <= 3
                r = it - 1
                if r >= y0:
                    em, ec, ep = mid[r - 1], mid[r], mid[r + 1]
                    om, oc, op = em[MPH:], ec[MPH:], ep[MPH:]
                    close = oned or (r & 1)
                    base = (r if oned else r >> 1) * cs + (x0 >> 1) - 1
                    for p in range(1, HXW + 1):
                        cx, cy = ec[p], oc[p]
                        ax0 = d * cx + c1 * (oc[p - 1] + cy) + cD * (om[p - 1] + op[p])
                        ax1 = d * cy + c1 * (cx + ec[p + 1]) + cD * (em[p] + ep[p + 1])
                        if cN != 0.0:
                            ax0 += cN * (em[p] + ep[p]); ax1 += cN * (om[p] + op[p])
                        a = acc.get(p, 0.0) + ((bold[p][0] - ax0) + (bold[p][1] - ax1))
                        if close:
                            RC[base + p] = w * a; a = 0.0
                        acc[p] = a
                bold = bvs
                rn = it - 1 + NS
                if rn <= rhi:
                    raw.pop(it - 1); stage(rn)
    return A, R, x, b, omega, XO, RC


@pytest.mark.parametrize("cfg", [((16, 16), 8, 4), ((16, 16), 16, 16), ((16, 16), 8, 6), ((32, 32), 8, 10, 5),
                                 ((256,), 8, 4), ((256,), 16, 6), ((512,), 32, 16, 6)])
def test_jr2_model_matches_oracle(cfg):
    A, R, x, b, omega, xo, rc = jr2_model(*cfg)
    xw = orc.jacobi(A, b, x.copy(), 1, omega)
    rw = R.dot(b - A.dot(xw))
    assert not np.isnan(rc).any() and not np.isnan(xo).any(), "every row is written exactly by its owner"
    assert np.abs(xo - xw).max() <= 1e-14
    assert np.abs(rc - rw).max() <= 1e-13
