"""wOversetGeneric (src/utils/haloExchange.F90:1471-1654) host model: donor-side 8-weight interpolation, same-rank
copies and a world_size-2 gloo exchange; weights as fracToWeights (src/overset/oversetUtilities.F90:2505-2519)."""
import os

import numpy as np

from adflow_b200 import make_params
from adflow_b200 import synthetic as syn
from adflow_b200.halo import build_overset_pattern, comm_vars, exchange_numpy_overset, trilinear_weights


def overset_entries(n0=(9, 8, 7), n1=(6, 7, 8), seed=5):
    """a fabricated overset connectivity between two blocks: a fringe shell of block 1 (its iMin/iMax owned layers,
    iblank = -1 in the reference) interpolates from block 0 and a few cells of block 0 interpolate from block 1.
    As in a valid overset assembly no donor stencil contains a fringe cell (the reference's same-rank loop is
    sequential and in place, so overlapping entries would make the result order dependent)."""
    rng = np.random.default_rng(seed)
    ent = []
    for k in range(2, n1[2] + 2):
        for j in range(2, n1[1] + 2):
            for i in (2, n1[0] + 1):
                dijk = (int(rng.integers(5, n0[0] + 1)), int(rng.integers(1, n0[1] + 1)), int(rng.integers(1, n0[2] + 1)))
                ent.append((0, dijk, tuple(rng.random(3)), 1, (i, j, k)))
    seen = set()
    for _ in range(40):
        fijk = (int(rng.integers(2, 4)), int(rng.integers(2, n0[1] + 2)), int(rng.integers(2, n0[2] + 2)))
        if fijk in seen:      # a cell is fringe at most once
            continue
        seen.add(fijk)
        dijk = (int(rng.integers(3, n1[0])), int(rng.integers(1, n1[1] + 1)), int(rng.integers(1, n1[2] + 1)))
        ent.append((1, dijk, tuple(rng.random(3)), 0, fijk))
    return ent


def two_blocks(prm, n0=(9, 8, 7), n1=(6, 7, 8)):
    return [syn.make_block(*n0, prm, seed=11), syn.make_block(*n1, prm, seed=12)]


VARS = lambda hb: comm_vars(hb, 1, 6, True, True, True, True)  # noqa: E731


def test_weights_partition_unity_and_reproduce_linear_fields():
    w = trilinear_weights((0.3, 0.6, 0.1))
    assert abs(w.sum() - 1.0) < 1e-15 and (w >= 0).all()
    prm = make_params()
    blocks = two_blocks(prm)
    for b in blocks:      # linear field in index space: interpolation is exact
        I, J, K = np.meshgrid(*[np.arange(n) for n in b.d.box], indexing="ij")
        b.p[...] = 2.0 * I - 3.0 * J + 0.5 * K
    ent = overset_entries()
    pat = build_overset_pattern(ent)
    assert len(pat["nbrRank"]) == 0 and len(pat["donorList"]) == len(ent)
    before = [b.p.copy() for b in blocks]
    exchange_numpy_overset(blocks, pat, lambda hb: [hb.p])
    for db, dijk, frac, fb, fijk in ent[:50]:
        want = 2.0 * (dijk[0] + frac[0]) - 3.0 * (dijk[1] + frac[1]) + 0.5 * (dijk[2] + frac[2])
        assert abs(blocks[fb].p[fijk] - want) < 1e-12
    changed = sum(int((a != b.p).sum()) for a, b in zip(before, blocks))
    assert changed > 0


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prm = make_params()
    blocks = two_blocks(prm)
    mine = [blocks[rank]]           # block b lives on rank b
    pat = build_overset_pattern(overset_entries(), rank=rank, owner=[0, 1])

    def sendrecv(peer, sendbuf, rshape):
        recv = torch.empty(rshape, dtype=torch.float64)
        ops = [dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(sendbuf)), peer),
               dist.P2POp(dist.irecv, recv, peer)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        return recv.numpy()

    exchange_numpy_overset(mine, pat, VARS, sendrecv)
    q.put((rank, mine[0].w.copy(), mine[0].p.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_overset_matches_single_rank():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, w, p_ = q.get(timeout=180)
        got[rank] = (w, p_)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prm = make_params()
    blocks = two_blocks(prm)
    # the send side of a rank reads its block BEFORE its own fringes are overwritten; the single-rank model
    # interpolates all donors first too (vals computed before any halo write), so results are identical
    exchange_numpy_overset(blocks, build_overset_pattern(overset_entries()), VARS)
    for b in range(2):
        assert np.array_equal(got[b][0], blocks[b].w) and np.array_equal(got[b][1], blocks[b].p)
