"""world_size-2 (gloo, CPU) run of a two-level multigrid V cycle: every rank owns one fine and one coarse block, the halo
exchange of EACH level goes through that level's own communication pattern (whalo2 on level 1, whalo1 on level 2) with
point-to-point messages between the ranks.  The per-block arithmetic is the oracle's; what is tested is the host-side
multi-level pattern logic the NCCL path shares (tests/test_mg_gpu.py runs the same cycle on two GPUs).  The result has
to equal the single-process run of the same global problem bit for bit."""
import os

import numpy as np

from adflow_b200 import make_params
from adflow_b200 import synthetic as syn
from adflow_b200.halo import BlockGrid, build_cartesian_pattern, comm_vars, exchange_numpy, make_grid_blocks
from oracle.pyoracle import Oracle

OPTS = {"equationType": "laminar NS", "nRKStages": 3, "resAveraging": "never"}
SHAPE = (8, 8, 6)


def v_cycle(prm, fine, coarse, pf, pc, sendrecv=None):
    vars_ = lambda hb: comm_vars(hb, 1, 5, True, True, True, False)  # noqa: E731

    def rk_smoother(blocks, pat):
        for hb in blocks:
            np.copyto(hb.wn, hb.w[..., :5]); np.copyto(hb.pn, hb.p)
        for st in range(1, prm.nRKStages + 1):
            for hb in blocks:
                Oracle(hb, prm).rk_stage(st)
            exchange_numpy(blocks, pat, vars_, sendrecv)
            if st < prm.nRKStages:
                for hb in blocks:
                    Oracle(hb, prm).residual_block(prm.cdisRK[st])

    for hb in fine:
        Oracle(hb, prm).apply_flow_bc(True)
    exchange_numpy(fine, pf, vars_, sendrecv)
    for hb in fine:
        o = Oracle(hb, prm)
        o.time_step(True); hb.fw[...] = 0; o.residual_block(prm.cdisRK[0])
    rk_smoother(fine, pf)
    for f, c in zip(fine, coarse):
        of, oc = Oracle(f, prm), Oracle(c, prm)
        of.time_step(False); of.residual_block(prm.cdisRK[0])
        oc.mg_restrict(of); oc.apply_flow_bc(False)
    exchange_numpy(coarse, pc, vars_, sendrecv)
    for c in coarse:
        oc = Oracle(c, prm)
        oc.time_step(True); oc.mg_store_w1(); oc.residual_block_coarse(prm.cdisRK[0], init=0); oc.mg_forcing()
    rk_smoother(coarse, pc)
    for f, c in zip(fine, coarse):
        of, oc = Oracle(f, prm), Oracle(c, prm)
        of.mg_prolong(oc); of.apply_flow_bc(True)
    exchange_numpy(fine, pf, vars_, sendrecv)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prm = make_params(OPTS)
    gf = BlockGrid((2, 1, 1), SHAPE, nranks=world)
    gc = BlockGrid((2, 1, 1), tuple(n // 2 for n in SHAPE), nranks=world)
    fine = make_grid_blocks(gf, rank, prm)
    coarse = [syn.make_coarse_block(hb, prm) for hb in fine]

    def sendrecv(peer, sendbuf, rshape):
        recv = torch.empty(rshape, dtype=torch.float64)
        ops = [dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(sendbuf)), peer), dist.P2POp(dist.irecv, recv, peer)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        return recv.numpy()

    v_cycle(prm, fine, coarse, build_cartesian_pattern(gf, rank), build_cartesian_pattern(gc, rank), sendrecv)
    q.put((rank, {b: hb.w.copy() for b, hb in zip(gf.local_blocks(rank), fine)}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_multigrid_cycle_matches_single_process():
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        _rank, out = q.get(timeout=300)
        got.update(out)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prm = make_params(OPTS)
    gf = BlockGrid((2, 1, 1), SHAPE, nranks=1)
    gc = BlockGrid((2, 1, 1), tuple(n // 2 for n in SHAPE), nranks=1)
    fine = make_grid_blocks(gf, 0, prm)
    coarse = [syn.make_coarse_block(hb, prm) for hb in fine]
    w0 = [hb.w.copy() for hb in fine]
    v_cycle(prm, fine, coarse, build_cartesian_pattern(gf, 0), build_cartesian_pattern(gc, 0))
    for b, hb in enumerate(fine):
        assert np.abs(hb.w - w0[b]).max() > 0
        assert np.array_equal(got[b], hb.w), (b, np.abs(got[b] - hb.w).max())
