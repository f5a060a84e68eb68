"""Host-side multigrid logic (no GPU): cycle strategies (extractMgInfo / setEntriesWcycle,
src/inputParam/inputParamRoutines.F90:880-945,1088-1180) and the restriction / interpolation tables of
createCoarseBlocks (src/preprocessing/coarseUtils.F90:254-420) as the synthetic generator builds them."""
import numpy as np
import pytest

from adflow_b200 import make_params
from adflow_b200 import synthetic as syn
from adflow_b200.solver import ADFLOW_B200


def n_steps_w(levels):   # computeNstepsWcycle
    return 4 if levels == 2 else 4 + 2 * n_steps_w(levels - 1)


@pytest.mark.parametrize("n", [2, 3, 4, 5])
def test_cycle_strategies(n):
    v = ADFLOW_B200.cycleStrategy("%dv" % n)
    w = ADFLOW_B200.cycleStrategy("%dw" % n)
    assert len(v) == 4 * n - 4 and len(w) == n_steps_w(n)
    for cyc in (v, w):
        lev = np.cumsum(cyc)
        assert lev[-1] == 0 and lev.min() == 0 and lev.max() == n - 1      # back on the ground level, n levels visited
        assert cyc[0] == 0 and all(a != b or a == 0 for a, b in zip(cyc, cyc[1:]) if a != 0)
    assert v == [0, 1] * (n - 1) + [0, -1] * (n - 1)
    assert ADFLOW_B200.cycleStrategy("sg") == [0]
    with pytest.raises(ValueError):
        ADFLOW_B200.cycleStrategy("1v")


def test_full_multigrid_schedule():
    """ground levels of the start-up (solver loop, src/solver/solvers.F90:63) and the cycle each of them runs"""
    assert ADFLOW_B200.fmgSchedule(1, "3w") == []                       # mgStartlevel = 1: nothing to do
    assert ADFLOW_B200.fmgSchedule(3, "3w") == [(3, "sg"), (2, "2w")]   # coarsest level single grid, then a 2-level W cycle
    assert ADFLOW_B200.fmgSchedule(2, "4v") == [(2, "3v")]
    assert ADFLOW_B200.fmgSchedule(4, "4V") == [(4, "sg"), (3, "2v"), (2, "3v")]
    for _, spec in ADFLOW_B200.fmgSchedule(4, "4w"):
        cyc = ADFLOW_B200.cycleStrategy(spec)
        assert cyc[0] == 0 and np.cumsum(cyc)[-1] == 0
    with pytest.raises(ValueError):
        ADFLOW_B200.fmgSchedule(3, "2v")                                 # start level beyond the cycle's levels
    with pytest.raises(ValueError):
        ADFLOW_B200.fmgSchedule(2, "sg")


@pytest.mark.parametrize("nx", [2, 5, 8, 9])
def test_transfer_tables(nx):
    keep = syn.mg_kept_nodes(nx)
    nc = int(keep.sum()) - 1
    fine, wgt, coarse = syn.mg_tables_1d(keep, nc + 2, nx + 2, nx + 3)
    # every fine cell belongs to exactly one coarse cell; irregular (single-cell) coarse cells carry weight 1/2
    seen = np.zeros(nx + 2, dtype=int)
    for ii in range(2, nc + 2):
        a, b = fine[ii]
        assert b in (a, a + 1) and 2 <= a <= nx + 1
        for c in {a, b}:
            seen[c] += 1
        assert wgt[ii] == (0.5 if a == b else 1.0)
    assert (seen[2:nx + 2] == 1).all()
    assert tuple(fine[1]) == (0, 1) and tuple(fine[nc + 2]) == (nx + 2, nx + 3)
    # interpolation: the nearest coarse cell is the one the fine cell was restricted into, the second one its neighbour
    for i in range(2, nx + 2):
        near, far = coarse[i]
        assert i in fine[near] and abs(int(far) - int(near)) <= 1 and 1 <= far <= nc + 2


def test_coarse_block_geometry():
    prm = make_params()
    fine = syn.make_block(8, 6, 5, prm)
    c = syn.make_coarse_block(fine, prm)
    assert (c.d.nx, c.d.ny, c.d.nz) == (4, 3, 3) and c.level == 2
    ow, owf = c.d.owned(), fine.d.owned()
    assert abs(c.vol[ow].sum() - fine.vol[owf].sum()) < 1e-12 * fine.vol[owf].sum()   # nested meshes: same total volume
    assert [s["bcType"] for s in c.subfaces] == [s["bcType"] for s in fine.subfaces]
    assert set(fine.mg) == {"mgICoarse", "mgJCoarse", "mgKCoarse"}
