"""Multi-GPU parity check, launched by torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tools/dist_check.py [--shape 64 64 64] [--gl 3]

Every rank builds (a) the row-slab sharded hierarchy and (b) a fully replicated one
(OMG_AGGLOMERATE_BELOW=huge: the single-GPU code path), runs the same V-cycles on both and
compares the gathered slab solution with the replicated one and, for small sizes, with the oracle.
Exit code 0 = parity.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs="+", default=[64, 64, 64])
    ap.add_argument("--gl", type=int, default=3)
    ap.add_argument("--cycles", type=int, default=4)
    ap.add_argument("--agg", type=int, default=4096)
    ap.add_argument("--oracle", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import openmg_b200 as omg
    from openmg_b200 import dist as odist
    from openmg_b200.hierarchy import Hierarchy
    rank, world = odist.init_from_torch(dist)
    shape = tuple(a.shape)
    A = omg.operators.poisson_band(shape, sparse_1d=(len(shape) == 1))
    N = A.n
    u = np.random.RandomState(0).random_sample(N)
    fails = 0

    os.environ["OMG_AGGLOMERATE_BELOW"] = str(1 << 62)
    h_rep = Hierarchy(A, shape, a.gl - 1, 8)
    assert not h_rep.local_range(0)[2]
    b = h_rep.matvec(u, 0)
    os.environ["OMG_AGGLOMERATE_BELOW"] = str(a.agg)
    h = Hierarchy(A, shape, a.gl - 1, 8)
    ranges = [h.local_range(l) for l in range(h.nlevels)]
    if rank == 0:
        print("levels (row0, nloc, slab) on rank 0:", ranges, flush=True)
    assert ranges[0][2], "level 0 should be a slab in this check"

    # operator application across the slab cut
    y = odist.gather_solution(dist, h, h.matvec(u, 0))
    err = np.abs(y - b).max() / np.abs(b).max()
    if rank == 0:
        print("matvec          max rel err %.2e" % err, flush=True)
    fails += err > 1e-14

    # one reference operation at a time, slab vs replicated
    rs = np.random.RandomState(5)
    for l in range(h.nlevels - 1):
        n = h.n(l)
        x, bb = rs.random_sample(n), rs.random_sample(n)
        e = rs.random_sample(h.n(l + 1))
        ops = [
            ("jacobi x2", lambda H: H.smooth(l, bb, x, 2, "jacobi", 0.8), l),
            ("rbgs x1", lambda H: H.smooth(l, bb, x, 1, "rbgs"), l),
            ("residual_restrict", lambda H: H.residual_restrict(l, bb, x), l + 1),
            ("prolong_correct", lambda H: H.prolong_correct(l, e, x), l),
            ("prolong+jacobi", lambda H: H.prolong_correct_smooth(l, bb, e, x, 1, "jacobi", 0.8), l),
            ("prolong+rbgs", lambda H: H.prolong_correct_smooth(l, bb, e, x, 1, "rbgs"), l),
        ]
        for name, fn, out_level in ops:
            want = fn(h_rep)
            got = odist.gather_solution(dist, h, fn(h), out_level)
            err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-300)
            if rank == 0:
                print("L%d %-18s max rel err %.2e%s" % (l, name, err, "  FAIL" if err > 1e-13 else ""), flush=True)
                if err > 1e-13 and os.environ.get("OMG_DIST_DEBUG"):
                    bad = np.nonzero(np.abs(got - want) > 1e-12 * np.abs(want).max())[0]
                    shp = tuple(int(v) >> (l if out_level == l else l + 1) for v in shape)
                    idx = np.array(np.unravel_index(bad, shp)).T
                    print("    %d bad entries; leading-index values %s; first %s" % (
                        len(bad), np.unique(idx[:, 0]).tolist()[:16], idx[:4].tolist()), flush=True)
            fails += err > 1e-13

    for smoother, pre, post in (("jacobi", 1, 1), ("jacobi", 2, 0), ("rbgs", 1, 1)):
        x_rep, c0, n_rep, hist_rep = h_rep.solve(b, None, pre, post, smoother, 0.8, a.cycles, 0.0, want_history=True)
        x_loc, c1, n_loc, hist = h.solve(b, None, pre, post, smoother, 0.8, a.cycles, 0.0, want_history=True)
        x = odist.gather_solution(dist, h, x_loc)
        err = np.abs(x - x_rep).max() / np.abs(x_rep).max()
        nerr = np.abs(hist - hist_rep).max() / hist_rep[0]
        line = "%-6s V(%d,%d): slab vs replicated max rel err %.2e, norm history err %.2e" % (
            smoother, pre, post, err, nerr)
        tol = 1e-12
        bad = err > tol or nerr > 1e-12
        if a.oracle:
            import oracle.openmg_oracle as orc
            A0 = orc.poisson_csr(shape, sparse_1d=(len(shape) == 1))
            params = {'problemShape': shape, 'gridLevels': a.gl, 'preIterations': pre, 'postIterations': post,
                      'verbose': False, 'minSize': 8}
            R = orc.restrictionList(shape, a.gl - 1, 8)
            params['coarsestLevel'] = len(R)
            Al = orc.coeffecientList(A0, R)
            sm = orc.make_smoother(smoother, shape, 0.8)
            xo = None
            for _ in range(a.cycles):
                xo, info = orc.mgCycle(Al, b, 0, R, params, initial=xo, smooth=sm)
            eo = np.abs(x - xo).max() / np.abs(xo).max()
            line += ", vs oracle %.2e" % eo
            bad = bad or eo > (1e-12 if smoother == "jacobi" else 1e-10)
        if rank == 0:
            print(line + ("  FAIL" if bad else ""), flush=True)
        fails += bad
    # the drop-in entry point itself: every rank calls openmg.mgSolve with the same global b and gets the full x
    for smoother, thr, cyc in (("jacobi", 0.0, a.cycles), ("rbgs", 1e-3 * float(np.linalg.norm(b)), 50)):
        params = {'problemShape': shape, 'gridLevels': a.gl, 'cycles': cyc, 'threshold': thr, 'preIterations': 1,
                  'postIterations': 1, 'smoother': smoother, 'giveInfo': True}
        os.environ["OMG_AGGLOMERATE_BELOW"] = str(a.agg)
        x_s, info_s = omg.mgSolve(A, b, dict(params))
        x_r, cyc_r, norm_r, _ = h_rep.solve(b, None, 1, 1, smoother, 0.8, cyc, thr)
        err = np.abs(x_s - x_r).max() / np.abs(x_r).max()
        bad = err > 1e-12 or info_s['cycle'] != cyc_r or abs(info_s['norm'] - norm_r) > 1e-9 * norm_r + 1e-300
        if rank == 0:
            print("mgSolve under torchrun (%s, %d cycles): full x on every rank, vs replicated max rel err %.2e%s"
                  % (smoother, info_s['cycle'], err, "  FAIL" if bad else ""), flush=True)
        fails += bad
    t = torch.tensor([float(fails)], device="cuda")
    dist.all_reduce(t)
    dist.barrier()
    dist.destroy_process_group()
    return 1 if t.item() > 0 else 0


if __name__ == "__main__":
    sys.exit(main())
