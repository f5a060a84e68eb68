"""Host logic of the halo exchange on CPU: pattern builder + numpy restatement of
whalo1to1RealGeneric, single process (internal copies) and world_size 2 over gloo."""
import os

import numpy as np
import pytest

from adflow_b200 import make_params
from adflow_b200.halo import BlockGrid, build_cartesian_pattern, comm_vars, exchange_numpy, make_grid_blocks


def global_field(grid, blocks_by_id, name, comp=None):
    """assemble the owned cells of all blocks into one global array"""
    n = grid.n
    G = np.zeros(tuple(grid.nb[a] * n[a] for a in range(3)))
    for b, hb in blocks_by_id.items():
        c = grid.coords[b]
        a = getattr(hb, name)
        if comp is not None:
            a = a[..., comp]
        G[c[0] * n[0]:(c[0] + 1) * n[0], c[1] * n[1]:(c[1] + 1) * n[1], c[2] * n[2]:(c[2] + 1) * n[2]] = a[hb.d.owned()]
    return G


def check_halos_against_global(grid, blocks_by_id, name, comp=None):
    G = global_field(grid, blocks_by_id, name, comp)
    n = grid.n
    nchecked = 0
    for b, hb in blocks_by_id.items():
        c = grid.coords[b]
        a = getattr(hb, name)
        if comp is not None:
            a = a[..., comp]
        for i in range(n[0] + 4):
            for j in range(n[1] + 4):
                for k in range(n[2] + 4):
                    g = (c[0] * n[0] + i - 2, c[1] * n[1] + j - 2, c[2] * n[2] + k - 2)
                    if all(0 <= g[q] < G.shape[q] for q in range(3)):
                        assert a[i, j, k] == G[g], (b, i, j, k)
                        nchecked += 1
    return nchecked


def test_pattern_counts_and_symmetry():
    grid = BlockGrid((2, 2, 1), (6, 5, 4), nranks=2)
    p0 = build_cartesian_pattern(grid, 0)
    p1 = build_cartesian_pattern(grid, 1)
    assert list(p0["nbrRank"]) == [1] and list(p1["nbrRank"]) == [0]
    assert p0["sendCount"][0] == p1["recvCount"][0] and p0["recvCount"][0] == p1["sendCount"][0]
    # every receive entry is a halo cell, every send entry an owned cell
    n = np.array(grid.n)
    own = lambda l: np.all((l[:, 1:] >= 2) & (l[:, 1:] <= n + 1), axis=1)  # noqa: E731
    assert own(p0["sendList"]).all() and not own(p0["recvList"]).any()
    assert own(p0["donorList"]).all() and not own(p0["haloList"]).any()


def test_single_rank_multiblock_exchange_fills_all_interior_halos():
    prm = make_params()
    grid = BlockGrid((2, 2, 2), (5, 4, 3), nranks=1)
    blocks = make_grid_blocks(grid, 0, prm)
    pat = build_cartesian_pattern(grid, 0)
    assert len(pat["nbrRank"]) == 0
    vars_of = lambda hb: comm_vars(hb, 1, 6, True, True, True, True)  # noqa: E731
    exchange_numpy(blocks, pat, vars_of)
    by_id = dict(zip(grid.local_blocks(0), blocks))
    for comp in range(6):
        assert check_halos_against_global(grid, by_id, "w", comp) > 0
    for name in ("p", "rlv", "rev"):
        check_halos_against_global(grid, by_id, name)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import torch

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prm = make_params()
    grid = BlockGrid((2, 2, 1), (5, 4, 3), nranks=world)
    blocks = make_grid_blocks(grid, rank, prm)
    pat = build_cartesian_pattern(grid, rank)

    def sendrecv(peer, sendbuf, rshape):
        recv = torch.empty(rshape, dtype=torch.float64)
        ops = [dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(sendbuf)), peer),
               dist.P2POp(dist.irecv, recv, peer)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        return recv.numpy()

    vars_of = lambda hb: comm_vars(hb, 1, 6, True, True, True, True)  # noqa: E731
    exchange_numpy(blocks, pat, vars_of, sendrecv)
    out = {b: (hb.w.copy(), hb.p.copy()) for b, hb in zip(grid.local_blocks(rank), blocks)}
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_exchange_matches_single_rank():
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, out = q.get(timeout=180)
        got.update(out)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank reference of the same global problem
    prm = make_params()
    grid = BlockGrid((2, 2, 1), (5, 4, 3), nranks=1)
    blocks = make_grid_blocks(grid, 0, prm)
    exchange_numpy(blocks, build_cartesian_pattern(grid, 0), lambda hb: comm_vars(hb, 1, 6, True, True, True, True))
    for b, hb in enumerate(blocks):
        assert np.array_equal(got[b][0], hb.w), b
        assert np.array_equal(got[b][1], hb.p), b
