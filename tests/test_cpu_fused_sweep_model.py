"""CPU models of the single-pass two-colour sweeps (`k_st2rb` in 2-D / 1-D, `k_rb3` in 3-D),
openmg_b200/csrc/omg_stencil.cu.

The CUDA kernel relaxes colour 0 of row r from raw rows r-1..r+1 into a "mid" ring and, one row behind,
colour 1 of row r-1 from mid rows r-2..r, per x-chunk and y-segment, with halo pairs / halo rows recomputed
redundantly, flat-index wraps at the row ends, the first/last-column correction taps of 2-D Galerkin levels and
pass A also run on the row slots -1 and NY.  This file restates exactly that index logic in numpy and checks it
against the oracle's definition of the sweep (oracle.rbgs: colour 0 then colour 1, same-colour couplings lagged)
on small hierarchies, so the *algorithm* is pinned on the CPU; the GPU parity tests pin the kernel itself.
(The model is how the missing row slots -1 / NY were found: the +-(N-1) taps of the two corner rows read the
first / last pair of the vector through those slots.)
"""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle.openmg_oracle as orc


def band_and_classes(Al, N):
    """Majority band {offset: coefficient} and the two column classes of a 2-D level (openmg_b200 setup logic)."""
    n = Al.shape[0]
    C = Al.tocoo()
    band = {}
    for d in np.unique(C.col - C.row):
        vals = C.data[(C.col - C.row) == d]
        u, c = np.unique(vals, return_counts=True)
        if c.max() > n / 2:
            band[int(d)] = float(u[np.argmax(c)])

    def delta(r):
        row = Al.getrow(r)
        dl = {}
        for c, v in zip(row.indices, row.data):
            dl[int(c - r)] = dl.get(int(c - r), 0.0) + v
        for k, v in band.items():
            dl[k] = dl.get(k, 0.0) - v
        return {k: v for k, v in dl.items() if v != 0.0}

    NY = n // N
    dL, dR = delta((NY // 2) * N), delta((NY // 2) * N + N - 1)
    assert set(dL) <= {-1, -(N + 1), N - 1} and set(dR) <= {1, N + 1, -(N - 1)}
    c2l = [dL.get(-1, 0.0), dL.get(-(N + 1), 0.0), dL.get(N - 1, 0.0)]
    c2r = [dR.get(1, 0.0), dR.get(N + 1, 0.0), dR.get(-(N - 1), 0.0)]
    return band, c2l, c2r


def fused_sweep_model(x, b, N, band, c2l, c2r, cflat, XW, YL, c0=0):
    """One two-colour sweep the way k_st2rb computes it (chunks of XW columns, segments of YL rows)."""
    n = x.size
    NY = n // N
    d = band[0]
    c1, cN, cD = band.get(1, 0.0), band.get(N, 0.0), band.get(N + 1, 0.0)
    wod = 1.0 / d
    PAD = 2 * N + 16
    xp = np.zeros(n + 2 * PAD)
    xp[PAD:PAD + n] = x
    bp = np.zeros(n + 2 * PAD)
    bp[PAD:PAD + n] = b
    out = np.full(n, np.nan)
    HXW, NPA = XW // 2, XW // 2 + 2
    for x0 in range(0, N, XW):
        for y0 in range(0, NY, YL):
            y1 = min(y0 + YL, NY)

            def raw(r):              # staged raw row slot r: columns x0-4 .. x0+XW+3 of the FLAT vector
                s = PAD + r * N + x0 - 4
                return xp[s:s + XW + 8]

            mid = {}
            for it in range(y0 - 1, y1 + 1):
                sm, sc, sq = raw(it - 1), raw(it), raw(it + 1)
                mw = np.zeros(XW + 4)
                for p in range(NPA):             # pass A, halo pairs included
                    o, xg = 2 + 2 * p, x0 - 2 + 2 * p
                    gi = it * N + xg
                    cx, cy = sc[o], sc[o + 1]
                    flip = 1 if (xg < 0 or xg >= N) else 0
                    ex = (c0 == 0) if cflat else (((it + flip) & 1) == c0)
                    if 0 <= gi < n:
                        if ex:
                            ax = d * cx + c1 * (sc[o - 1] + cy) + cN * (sm[o] + sq[o]) + cD * (sm[o - 1] + sq[o + 1])
                            if xg == 0 or xg == N:
                                ax += c2l[0] * sc[o - 1] + c2l[1] * sm[o - 1] + c2l[2] * sq[o - 1]
                            cx += wod * (bp[PAD + gi] - ax)
                        else:
                            ax = d * cy + c1 * (cx + sc[o + 2]) + cN * (sm[o + 1] + sq[o + 1]) + cD * (sm[o] + sq[o + 2])
                            if xg == N - 2 or xg == -2:
                                ax += c2r[0] * sc[o + 2] + c2r[1] * sq[o + 2] + c2r[2] * sm[o + 2]
                            cy += wod * (bp[PAD + gi + 1] - ax)
                    mw[2 * p], mw[2 * p + 1] = cx, cy
                mid[it] = mw
                r = it - 1
                if r < y0:
                    continue
                sm, sc, sq = mid[r - 1], mid[r], mid[r + 1]
                ex = (c0 == 0) if cflat else ((r & 1) == c0)
                for p in range(1, HXW + 1):      # pass B, owned pairs
                    o, xg = 2 * p, x0 - 2 + 2 * p
                    cx, cy = sc[o], sc[o + 1]
                    if ex:
                        ax = d * cy + c1 * (cx + sc[o + 2]) + cN * (sm[o + 1] + sq[o + 1]) + cD * (sm[o] + sq[o + 2])
                        if xg == N - 2:
                            ax += c2r[0] * sc[o + 2] + c2r[1] * sq[o + 2] + c2r[2] * sm[o + 2]
                        cy += wod * (b[r * N + xg + 1] - ax)
                    else:
                        ax = d * cx + c1 * (sc[o - 1] + cy) + cN * (sm[o] + sq[o]) + cD * (sm[o - 1] + sq[o + 1])
                        if xg == 0:
                            ax += c2l[0] * sc[o - 1] + c2l[1] * sm[o - 1] + c2l[2] * sq[o - 1]
                        cx += wod * (b[r * N + xg] - ax)
                    out[r * N + xg], out[r * N + xg + 1] = cx, cy
    assert not np.isnan(out).any()
    return out


@pytest.mark.parametrize("shape,XW_div,YL", [((16, 16), 1, 16), ((32, 32), 2, 6), ((32, 32), 1, 2), ((64, 64), 4, 64)])
def test_single_pass_sweep_equals_two_half_sweeps_2d(shape, XW_div, YL):
    A0 = orc.poisson_csr(shape)
    R = orc.restrictionList(shape, 2, 4)
    A = orc.coeffecientList(A0, R)
    rs = np.random.RandomState(3)
    for l in range(len(A)):
        Al = sp.csr_matrix(A[l])
        n = Al.shape[0]
        N = shape[0] >> l
        if N < 8:
            break
        band, c2l, c2r = band_and_classes(Al, N)
        if l == 0:
            assert c2l == [0, 0, 0] and c2r == [0, 0, 0] and N not in band          # level 0: pure band, no +-N pair
        x, b = rs.random_sample(n), rs.random_sample(n)
        col = orc.colouring(shape, l, n)
        want = orc.rbgs(Al, b, x.copy(), 1, col)
        XW = max(N // XW_div, 4)
        got = fused_sweep_model(x, b, N, band, c2l, c2r, cflat=(l == 0), XW=XW, YL=min(YL, n // N))
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-13 * np.abs(want).max(), err_msg="level %d" % l)


def test_single_pass_sweep_1d_rows_view():
    """1-D levels run on the same kernel with the vector viewed as rows of N (flat colouring, c1 only)."""
    shape = (256,)
    A0 = sp.csr_matrix(orc.poisson_csr(shape, sparse_1d=True))
    n = A0.shape[0]
    rs = np.random.RandomState(4)
    x, b = rs.random_sample(n), rs.random_sample(n)
    want = orc.rbgs(A0, b, x.copy(), 1, orc.colouring(shape, 0, n))
    band = {0: A0[1, 1], 1: A0[1, 2], -1: A0[1, 0]}
    for N, XW, YL in ((32, 32, 8), (64, 16, 2), (16, 8, 3)):
        got = fused_sweep_model(x, b, N, band, [0, 0, 0], [0, 0, 0], cflat=True, XW=XW, YL=YL)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-13 * np.abs(want).max())


def test_corner_rows_couple_through_the_wrapped_pairs():
    """Why pass A also runs on the row slots -1 and NY: on a 2-D Galerkin level row (0, N-1) couples to (0, 0)
    through the tap -(N-1) and row (NY-1, 0) to (NY-1, N-1) through +(N-1); in the staged layout those neighbours
    are the wrapped halo pairs of the slots -1 and NY, so their mid values must exist."""
    shape = (32, 32)
    A = orc.coeffecientList(orc.poisson_csr(shape), orc.restrictionList(shape, 1, 4))
    Al = sp.csr_matrix(A[1])
    n, N = Al.shape[0], 16
    assert Al[N - 1, 0] != 0.0 and Al[n - N, n - 1] != 0.0
    band, c2l, c2r = band_and_classes(Al, N)
    assert c2r[2] == Al[N - 1, 0] and c2l[2] == Al[n - N, n - 1]


# ------------------------------------------------------------------ 3-D: the register-pipelined sweep (k_rb3)

def fused_sweep_model_3d(x, b, S1, NY, NZ, d, c1, cS, cP, TY, ZL, c0=0):
    """One two-colour sweep the way k_rb3 computes it: full-row chunks of TY rows, z-segments of ZL planes; pass A
    relaxes the colour-c0 points of plane p (rows y0-1 .. y0+TY, i.e. including the two recomputed halo rows, which
    in the first / last chunk are rows of the neighbouring plane with flipped colour parity) from the staged raw
    planes into the mid plane p; one plane behind, pass B relaxes the other colour of plane p-1 from the mid planes.
    Centres and z-neighbours are carried from step to step ("registers"), in-plane neighbours come from the staged
    raw plane p (pass A) resp. the stored mid plane p-1 (pass B).  The kernel groups the x-pairs modelled here into
    2x2 patches (interior rows) and single pairs (halo rows); the arithmetic per point is the same."""
    S2 = S1 * NY
    n = S2 * NZ
    wod = 1.0 / d
    PAD = 2 * S2
    xp = np.zeros(n + 2 * PAD)
    xp[PAD:PAD + n] = x
    bp = np.zeros(n + 2 * PAD)
    bp[PAD:PAD + n] = b
    out = np.full(n, np.nan)
    HX = S1 // 2
    RS, MS = (TY + 4) * S1, (TY + 2) * S1
    for y0 in range(0, NY, TY):
        for z0 in range(0, NZ, ZL):
            z1 = min(z0 + ZL, NZ)

            def rawplane(p):                 # staged rows y0-2 .. y0+TY+1; planes < -1 or > NZ are all zero
                if p < -1 or p > NZ:
                    return None
                s = PAD + p * S2 + (y0 - 2) * S1
                return xp[s:s + RS]

            Q, J = np.meshgrid(np.arange(TY + 2), np.arange(HX), indexing='ij')
            Q, J = Q.ravel(), J.ravel()
            idx = np.arange(Q.size)
            yr = y0 - 1 + Q
            wy = ((yr < 0) | (yr >= NY)).astype(int)       # the row belongs to the neighbouring plane: parity flips
            roff, moff = (Q + 1) * S1 + 2 * J, Q * S1 + 2 * J
            own = (Q >= 1) & (Q <= TY)

            def e_of(p):                     # in-pair position of the colour-c0 point on plane p
                return np.where(((yr + p + wy) & 1) == c0, 0, 1)

            first = rawplane(z0 - 1)
            rc = np.stack([first[roff], first[roff + 1]], 1)
            below = rawplane(z0 - 2)
            rme = np.zeros(Q.size) if below is None else below[roff + e_of(z0 - 1)]
            mc, mme, bk = np.zeros((Q.size, 2)), np.zeros(Q.size), np.zeros(Q.size)
            mid = {0: np.zeros(MS), 1: np.zeros(MS)}
            for p in range(z0 - 1, z1 + 1):
                e = e_of(p)
                rawc, nxt = rawplane(p), rawplane(p + 1)
                rp = np.zeros((Q.size, 2)) if nxt is None else np.stack([nxt[roff], nxt[roff + 1]], 1)
                g = p * S2 + yr * S1 + 2 * J
                valid = (g >= 0) & (g < n)
                c = rc[idx, e]
                xl = np.where(e == 0, rawc[roff - 1], rc[:, 0])
                xr = np.where(e == 0, rc[:, 1], rawc[roff + 2])
                ax = d * c + c1 * (xl + xr) + cS * (rawc[roff - S1 + e] + rawc[roff + S1 + e]) + cP * (rme + rp[idx, e])
                b0, b1 = bp[PAD + g], bp[PAD + g + 1]
                mp = rc.copy()
                mp[idx, e] = np.where(valid, c + wod * (np.where(e == 0, b0, b1) - ax), c)
                bknew = np.where(e == 0, b1, b0)
                mid[p & 1][moff], mid[p & 1][moff + 1] = mp[:, 0], mp[:, 1]
                if p - 1 >= z0:              # pass B: the colour-1 point of plane p-1 sits where e points on plane p
                    mprev = mid[(p - 1) & 1]
                    c = mc[idx, e]
                    xl = np.where(e == 0, mprev[np.maximum(moff - 1, 0)], mc[:, 0])
                    xr = np.where(e == 0, mc[:, 1], mprev[np.minimum(moff + 2, MS - 1)])
                    yu, yd = mprev[np.maximum(moff - S1 + e, 0)], mprev[np.minimum(moff + S1 + e, MS - 1)]
                    ax = d * c + c1 * (xl + xr) + cS * (yu + yd) + cP * (mme + mp[idx, e])
                    o = mc.copy()
                    o[idx, e] = c + wod * (bk - ax)
                    go = (p - 1) * S2 + yr * S1 + 2 * J
                    out[go[own]], out[go[own] + 1] = o[own, 0], o[own, 1]
                mme, mc = mc[idx, 1 - e], mp
                rme, rc = rc[idx, 1 - e], rp
                bk = bknew
    assert not np.isnan(out).any()
    return out


@pytest.mark.parametrize("shape,TY,ZL", [((8, 8, 8), 4, 8), ((8, 8, 8), 2, 3), ((16, 16, 16), 4, 5),
                                         ((16, 8, 16), 4, 16), ((32, 16, 32), 8, 7)])
def test_single_pass_sweep_3d_register_pipeline_model(shape, TY, ZL):
    A0 = sp.csr_matrix(orc.poisson_csr(shape))
    n = A0.shape[0]
    NZ, NY, S1 = shape
    rs = np.random.RandomState(0)
    x, b = rs.random_sample(n), rs.random_sample(n)
    want = orc.rbgs(A0, b, x.copy(), 1, orc.colouring(shape, 0, n))
    got = fused_sweep_model_3d(x, b, S1, NY, NZ, -12.0, 1.0, 1.0, 1.0, TY, ZL)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-13 * np.abs(want).max())
