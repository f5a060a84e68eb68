"""Device overset exchange (adfb_comm_set_overset + adfb_halo_exchange) against the host model of wOversetGeneric."""
import os

import numpy as np
import pytest

from adflow_b200 import make_params
from adflow_b200.halo import build_overset_pattern, exchange_numpy_overset
from adflow_b200.solver import ADFLOW_B200

from test_overset_host import VARS, overset_entries, two_blocks

pytestmark = pytest.mark.gpu


def _close(a, b):
    # the device evaluates the 8-term weighted sum with FMA contraction: a few ulp
    return np.abs(a - b).max() <= 4e-15 * max(1.0, np.abs(b).max())


def test_overset_exchange_single_gpu(cuda_lib):
    prm = make_params()
    blocks = two_blocks(prm)
    ref = [b.copy() for b in blocks]
    pat = build_overset_pattern(overset_entries())
    exchange_numpy_overset(ref, pat, VARS)
    s = ADFLOW_B200(prm)
    try:
        for hb in blocks:
            s.addBlock(hb)
        s.setOversetPattern(pat)
        s.haloExchange(1, 6, True, True, True)
        for q, hb in enumerate(blocks):
            w, p, rlv, rev = s.downloadState(q)
            ow = hb.d.owned()
            mask = np.ones(hb.d.box, bool)
            # whalo2 recomputes rhoE of the owned cells afterwards (computeEtotBlock): compare it separately
            for l in (0, 1, 2, 3, 5):
                assert _close(w[..., l], ref[q].w[..., l]), (q, l)
            mask[ow] = False
            assert _close(w[..., 4][mask], ref[q].w[..., 4][mask])
            assert _close(p, ref[q].p) and _close(rlv, ref[q].rlv) and _close(rev, ref[q].rev)
            assert np.abs(p - hb.p).max() > 0   # fringes changed
    finally:
        s.close()


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adflow_b200 import _lib

    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        import ctypes as C
        buf = (C.c_char * 128)()
        assert _lib.load().adfb_get_unique_id(buf) == 0
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    dist.broadcast(uid, 0)
    prm = make_params()
    blocks = two_blocks(prm)
    pat = build_overset_pattern(overset_entries(), rank=rank, owner=[0, 1])
    s = ADFLOW_B200(prm, device=rank, rank=rank, nranks=world, unique_id=bytes(uid.numpy().tobytes()))
    s.addBlock(blocks[rank])
    s.setOversetPattern(pat)
    s.haloExchange(1, 6, True, True, True)
    w, p, rlv, rev = s.downloadState(0)
    s.close()
    q.put((rank, w, p))
    dist.barrier()
    dist.destroy_process_group()


def test_overset_exchange_two_gpus_nccl(cuda_lib):
    if cuda_lib.adfb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, w, p_ = q.get(timeout=600)
        got[rank] = (w, p_)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    prm = make_params()
    ref = two_blocks(prm)
    exchange_numpy_overset(ref, build_overset_pattern(overset_entries()), VARS)
    for b in range(2):
        for l in (0, 1, 2, 3, 5):
            assert _close(got[b][0][..., l], ref[b].w[..., l]), (b, l)
        assert _close(got[b][1], ref[b].p)


def test_residual_with_overset_pattern_matches_oracle(cuda_lib):
    """blocketteRes with overset blocks present (src/NKSolver/blockette.F90:213-262): p/rlv/rev, BCs, whalo2 = overset
    interpolation followed by computeEtotBlock on the owned cells (the fringe rhoE follows the interpolated p, rho, u),
    then the turbulence and flow BCs AGAIN on every block, then the residual core -- composed from the oracle's pieces
    and compared with adfb_residual."""
    import ctypes as C

    from adflow_b200.solver import RES_FLOW, RES_TURB
    from oracle.pyoracle import Oracle

    from util import rel_l2

    prm = make_params()
    blocks = two_blocks(prm)
    pat = build_overset_pattern(overset_entries())
    ref = [b.copy() for b in blocks]
    orcs = [Oracle(b, prm) for b in ref]
    for o in orcs:
        o.pressure(False); o.lam_viscosity(False); o.eddy_viscosity(False)
        o.apply_turb_bc(True); o.apply_flow_bc(True)
    exchange_numpy_overset(ref, pat, VARS)
    for o, b in zip(orcs, ref):
        d = b.d
        o.L.orc_etot(C.byref(o.ob), C.byref(prm), 2, d.il, 2, d.jl, 2, d.kl)
        o.apply_turb_bc(True); o.apply_flow_bc(True)
        o.residual_core(RES_FLOW | RES_TURB)
    s = ADFLOW_B200(prm)
    try:
        for hb in blocks:
            s.addBlock(hb)
        s.setOversetPattern(pat)
        s.residual(RES_FLOW | RES_TURB)
        for q, b in enumerate(ref):
            dw = s.downloadResidual(q)
            ow = b.d.owned()
            for l in range(6):
                err = rel_l2(dw[ow + (l,)], b.dw[ow + (l,)])
                assert err < 1e-11, (q, l, err)   # the 8-term interpolation sums differ by FMA contraction (a few ulp)
    finally:
        s.close()


def test_orphan_average_on_device_matches_oracle(cuda_lib):
    """whalo2 with an orphan list (adfb_block_set_orphans): the exchange ends with orphanAverage
    (src/utils/haloExchange.F90:201-354) -- against the oracle's routine, which is pinned bit for bit against the
    reference's (tests/test_oracle_vs_reference_orphans.py)."""
    import ctypes as C

    from oracle.pyoracle import Oracle

    from test_oracle_vs_reference_orphans import orphan_case

    prm, hb, orph = orphan_case()
    mu_inf, ratio = 1.7e-3, 0.009
    ho = hb.copy()
    o = Oracle(ho, prm)
    flat = np.ascontiguousarray(orph.reshape(-1))
    o.L.orc_orphan_average(C.byref(o.ob), C.byref(prm), len(orph), flat.ctypes.data_as(C.c_void_p), 1, 6, 1, 1, 1, C.c_double(mu_inf),
                           C.c_double(ratio))
    d = hb.d
    o.L.orc_etot(C.byref(o.ob), C.byref(prm), 2, d.il, 2, d.jl, 2, d.kl)   # whalo2 ends with computeEtotBlock on the owned cells
    s = ADFLOW_B200(prm)
    try:
        s.addBlock(hb)
        s.setOrphans(0, orph, mu_inf, ratio)
        s.haloExchange(1, 6, True, True, True)
        w, p, rlv, rev = s.downloadState(0)
    finally:
        s.close()
    assert np.abs(w - hb.w).max() > 0
    assert np.array_equal(w[..., (0, 1, 2, 3, 5)], ho.w[..., (0, 1, 2, 3, 5)])   # sums of <= 6 terms in the reference's order
    assert _close(w[..., 4], ho.w[..., 4])
    assert np.array_equal(p, ho.p) and np.array_equal(rlv, ho.rlv) and np.array_equal(rev, ho.rev)
