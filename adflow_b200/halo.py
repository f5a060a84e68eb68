"""Host logic of the halo exchange: 1-to-1 communication patterns.

The reference builds per-level send/recv index lists in preprocessing
(``src/preprocessing/pointMatchedCommPattern.F90``; types ``commType`` /
``internalCommType`` in ``src/modules/communication.F90:85-168``) and
``whalo1to1RealGeneric`` (``src/utils/haloExchange.F90:553-719``) walks them.  That
preprocessing (CGNS connectivity) is out of scope; for the synthetic multi-block
configurations this module produces the same kind of lists for a Cartesian
arrangement of equally sized blocks with identity orientation, *including* the
indirect (edge / corner) halos the reference's pattern also carries
(``pointMatchedCommPattern.F90:54-95``).

``build_cartesian_pattern`` is pure host logic (unit-tested on CPU, world_size 2 over
gloo); the product moves the data with device pack/unpack kernels + NCCL
(``adfb_comm_set_pattern`` / ``adfb_halo_exchange``).  ``exchange_numpy`` is the
numpy restatement of ``whalo1to1RealGeneric`` used by the tests as the checker.
"""
import itertools

import numpy as np


class BlockGrid:
    """nb = (NBi, NBj, NBk) blocks of n = (nx, ny, nz) cells; block b <-> rank owner[b]."""

    def __init__(self, nb, n, nranks=1, owner=None):
        self.nb, self.n = tuple(nb), tuple(n)
        self.nblocks = nb[0] * nb[1] * nb[2]
        self.coords = list(itertools.product(range(nb[0]), range(nb[1]), range(nb[2])))
        # block id: i fastest
        self.coords.sort(key=lambda c: (c[2], c[1], c[0]))
        self.id_of = {c: q for q, c in enumerate(self.coords)}
        if owner is None:
            per = -(-self.nblocks // nranks)
            owner = [q // per for q in range(self.nblocks)]
        self.owner = list(owner)
        self.nranks = nranks

    def local_blocks(self, rank):
        return [q for q in range(self.nblocks) if self.owner[q] == rank]

    def physical_faces(self, bid):
        """block faces on the boundary of the global domain (iMin=1 .. kMax=6)."""
        c = self.coords[bid]
        faces = []
        for ax in range(3):
            if c[ax] == 0:
                faces.append(2 * ax + 1)
            if c[ax] == self.nb[ax] - 1:
                faces.append(2 * ax + 2)
        return faces


def _halo_pairs(grid):
    """all (recvBlock, i,j,k) <- (sendBlock, i,j,k) pairs as int arrays (n,8),
    ordered by receiving block, then k, j, i (i fastest) -- vectorised."""
    n = np.array(grid.n)
    ext = n + 4
    kk, jj, ii = np.meshgrid(np.arange(ext[2]), np.arange(ext[1]), np.arange(ext[0]), indexing="ij")
    loc = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)  # i fastest
    owned = np.all((loc >= 2) & (loc <= n + 1), axis=1)
    loc = loc[~owned]
    nbarr = np.array(grid.nb)
    out = []
    for rb in range(grid.nblocks):
        c = np.array(grid.coords[rb])
        gidx = c * n + (loc - 2)
        inside = np.all((gidx >= 0) & (gidx < nbarr * n), axis=1)
        gi, lo = gidx[inside], loc[inside]
        sc = gi // n
        sl = gi % n + 2
        sb = np.array([grid.id_of[tuple(x)] for x in np.unique(sc, axis=0)])
        # map block coords -> id without a python loop over cells
        key = (sc[:, 2] * nbarr[1] + sc[:, 1]) * nbarr[0] + sc[:, 0]
        lut = np.full(nbarr.prod(), -1, dtype=np.int64)
        for cc, q in grid.id_of.items():
            lut[(cc[2] * nbarr[1] + cc[1]) * nbarr[0] + cc[0]] = q
        sbid = lut[key]
        assert sb.size and (sbid >= 0).all() or gi.shape[0] == 0
        out.append(np.column_stack([np.full(len(lo), rb), lo, sbid, sl]))
    return np.concatenate(out) if out else np.zeros((0, 8), dtype=np.int64)


def build_cartesian_pattern(grid, rank):
    """Pattern of `rank`: dict with nbrRank, sendCount, recvCount, sendList, recvList
    (int32 (n,4): LOCAL block index, i, j, k), donorList, haloList (same-rank copies)."""
    local = grid.local_blocks(rank)
    lidx = np.full(grid.nblocks, -1, dtype=np.int64)
    lidx[local] = np.arange(len(local))
    owner = np.array(grid.owner)
    P = _halo_pairs(grid)
    ro, so = owner[P[:, 0]], owner[P[:, 4]]
    recv_side = np.column_stack([lidx[P[:, 0]], P[:, 1:4]])
    send_side = np.column_stack([lidx[P[:, 4]], P[:, 5:8]])
    as_arr = lambda a: np.ascontiguousarray(a, dtype=np.int32).reshape(-1, 4)  # noqa: E731
    internal = (ro == rank) & (so == rank)
    nbrs = sorted(set(so[(ro == rank) & (so != rank)].tolist()) | set(ro[(so == rank) & (ro != rank)].tolist()))
    sends = [send_side[(so == rank) & (ro == r)] for r in nbrs]
    recvs = [recv_side[(ro == rank) & (so == r)] for r in nbrs]
    cat = lambda lst: np.concatenate(lst) if lst else np.zeros((0, 4), dtype=np.int64)  # noqa: E731
    return {
        "nbrRank": np.array(nbrs, dtype=np.int32),
        "sendCount": np.array([len(x) for x in sends], dtype=np.int32),
        "recvCount": np.array([len(x) for x in recvs], dtype=np.int32),
        "sendList": as_arr(cat(sends)),
        "recvList": as_arr(cat(recvs)),
        "donorList": as_arr(send_side[internal]),
        "haloList": as_arr(recv_side[internal]),
    }


def comm_vars(hb, start, end, comm_pressure, comm_viscous, viscous, eddy):
    """setCommPointers (haloExchange.F90:356-470): list of (array, component) views."""
    v = [hb.w[..., l - 1] for l in range(start, min(end, hb.nw) + 1)]
    if comm_pressure:
        v.append(hb.p)
    if viscous and comm_viscous:
        v.append(hb.rlv)
    if eddy and comm_viscous:
        v.append(hb.rev)
    return v


def exchange_numpy(blocks, pat, vars_of, sendrecv=None):
    """whalo1to1RealGeneric on numpy blocks.  `blocks`: local HostBlocks in local order;
    `vars_of(hb)` -> list of 3-D arrays; `sendrecv(peer, sendbuf, nrecv)` -> recvbuf moves
    one message pair (None for single-rank patterns)."""
    V = [vars_of(b) for b in blocks]
    nvar = len(V[0]) if V else 0
    so = 0
    ro = 0
    recvbufs = []
    for m, peer in enumerate(pat["nbrRank"]):
        ns, nr = int(pat["sendCount"][m]), int(pat["recvCount"][m])
        sl = pat["sendList"][so:so + ns]
        buf = np.empty((nvar, ns))
        for v in range(nvar):
            for b in np.unique(sl[:, 0]):
                sel = sl[:, 0] == b
                buf[v, sel] = V[b][v][sl[sel, 1], sl[sel, 2], sl[sel, 3]]
        recvbufs.append(sendrecv(int(peer), buf, (nvar, nr)))
        so += ns
        ro += nr
    dl, hl = pat["donorList"], pat["haloList"]
    if len(dl):
        vals = np.empty((nvar, len(dl)))
        for v in range(nvar):
            for b in np.unique(dl[:, 0]):
                sel = dl[:, 0] == b
                vals[v, sel] = V[b][v][dl[sel, 1], dl[sel, 2], dl[sel, 3]]
        for v in range(nvar):
            for b in np.unique(hl[:, 0]):
                sel = hl[:, 0] == b
                V[b][v][hl[sel, 1], hl[sel, 2], hl[sel, 3]] = vals[v, sel]
    ro = 0
    for m, _peer in enumerate(pat["nbrRank"]):
        nr = int(pat["recvCount"][m])
        rl = pat["recvList"][ro:ro + nr]
        buf = recvbufs[m]
        for v in range(nvar):
            for b in np.unique(rl[:, 0]):
                sel = rl[:, 0] == b
                V[b][v][rl[sel, 1], rl[sel, 2], rl[sel, 3]] = buf[v, sel]
        ro += nr


def make_grid_blocks(grid, rank, prm, seed=314):
    """Synthetic blocks of one rank with a globally consistent mesh and state."""
    from . import synthetic as syn

    gshape = tuple(grid.nb[a] * grid.n[a] for a in range(3))
    out = []
    for b in grid.local_blocks(rank):
        c = grid.coords[b]
        origin = tuple(c[a] * grid.n[a] for a in range(3))
        hb = syn.make_block(*grid.n, prm, origin=origin, global_n=gshape, seed=seed, origin_tag=b,
                            physical_faces=tuple(grid.physical_faces(b)))
        out.append(hb)
    return out


# ---------------------------------------------------------------------------------------------
# overset (interpolating) exchange: wOversetGeneric, src/utils/haloExchange.F90:1471-1654
def trilinear_weights(frac):
    """the 8 weights of sendList%interp / donorInterp from the fractional position (u, v, w) in the donor
    stencil, in the reference's order (i fastest): the reference stores the weights themselves, computed by
    fracToWeights (src/overset/oversetUtilities.F90) = tensor product of (1-u, u), (1-v, v), (1-w, w)."""
    u, v, w = frac
    return np.array([(1 - u) * (1 - v) * (1 - w), u * (1 - v) * (1 - w), (1 - u) * v * (1 - w), u * v * (1 - w),
                     (1 - u) * (1 - v) * w, u * (1 - v) * w, (1 - u) * v * w, u * v * w])


def build_overset_pattern(entries, rank=0, owner=None):
    """entries: list of (donorBlock, (i, j, k) low corner, frac (u,v,w), fringeBlock, (i, j, k)) with GLOBAL block
    ids; owner[b] = rank of block b (default all on rank 0).  Returns the pattern of `rank` with local block ids
    (position among the rank's blocks in ascending global id), message entries ordered as they appear."""
    nblk = 1 + max(max(e[0], e[3]) for e in entries)
    owner = [0] * nblk if owner is None else list(owner)
    local = {}
    for b in range(nblk):
        local[b] = sum(1 for q in range(b) if owner[q] == owner[b])
    send, recv, don, halo, dw, sw = {}, {}, [], [], [], {}
    for db, dijk, frac, fb, fijk in entries:
        w = trilinear_weights(frac)
        if owner[db] == rank and owner[fb] == rank:
            don.append((local[db],) + tuple(dijk)); halo.append((local[fb],) + tuple(fijk)); dw.append(w)
        elif owner[db] == rank:
            send.setdefault(owner[fb], []).append((local[db],) + tuple(dijk)); sw.setdefault(owner[fb], []).append(w)
        elif owner[fb] == rank:
            recv.setdefault(owner[db], []).append((local[fb],) + tuple(fijk))
    peers = sorted(set(send) | set(recv))
    i4 = lambda rows: np.array(rows, dtype=np.int32).reshape(-1, 4)  # noqa: E731
    return {
        "nbrRank": np.array(peers, dtype=np.int32),
        "sendCount": np.array([len(send.get(q, [])) for q in peers], dtype=np.int32),
        "recvCount": np.array([len(recv.get(q, [])) for q in peers], dtype=np.int32),
        "sendList": i4([r for q in peers for r in send.get(q, [])]),
        "sendInterp": np.array([w for q in peers for w in sw.get(q, [])], dtype=np.float64).reshape(-1, 8),
        "recvList": i4([r for q in peers for r in recv.get(q, [])]),
        "donorList": i4(don), "donorInterp": np.array(dw, dtype=np.float64).reshape(-1, 8), "haloList": i4(halo),
    }


def _interp8(a, rows, wts):
    i, j, k = rows[:, 1], rows[:, 2], rows[:, 3]
    return (wts[:, 0] * a[i, j, k] + wts[:, 1] * a[i + 1, j, k] + wts[:, 2] * a[i, j + 1, k] + wts[:, 3] * a[i + 1, j + 1, k] +
            wts[:, 4] * a[i, j, k + 1] + wts[:, 5] * a[i + 1, j, k + 1] + wts[:, 6] * a[i, j + 1, k + 1] +
            wts[:, 7] * a[i + 1, j + 1, k + 1])


def exchange_numpy_overset(blocks, pat, vars_of, sendrecv=None):
    """wOversetGeneric on numpy blocks (same conventions as exchange_numpy)."""
    V = [vars_of(b) for b in blocks]
    nvar = len(V[0]) if V else 0
    so = 0
    recvbufs = []
    for m, peer in enumerate(pat["nbrRank"]):
        ns, nr = int(pat["sendCount"][m]), int(pat["recvCount"][m])
        sl, sw = pat["sendList"][so:so + ns], pat["sendInterp"][so:so + ns]
        buf = np.empty((nvar, ns))
        for v in range(nvar):
            for b in np.unique(sl[:, 0]):
                sel = sl[:, 0] == b
                buf[v, sel] = _interp8(V[b][v], sl[sel], sw[sel])
        recvbufs.append(sendrecv(int(peer), buf, (nvar, nr)))
        so += ns
    dl, hl, dwt = pat["donorList"], pat["haloList"], pat["donorInterp"]
    if len(dl):
        vals = np.empty((nvar, len(dl)))
        for v in range(nvar):
            for b in np.unique(dl[:, 0]):
                sel = dl[:, 0] == b
                vals[v, sel] = _interp8(V[b][v], dl[sel], dwt[sel])
        for v in range(nvar):
            for b in np.unique(hl[:, 0]):
                sel = hl[:, 0] == b
                V[b][v][hl[sel, 1], hl[sel, 2], hl[sel, 3]] = vals[v, sel]
    ro = 0
    for m, _peer in enumerate(pat["nbrRank"]):
        nr = int(pat["recvCount"][m])
        rl = pat["recvList"][ro:ro + nr]
        for v in range(nvar):
            for b in np.unique(rl[:, 0]):
                sel = rl[:, 0] == b
                V[b][v][rl[sel, 1], rl[sel, 2], rl[sel, 3]] = recvbufs[m][v, sel]
        ro += nr
