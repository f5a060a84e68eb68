"""Block extents and host-array conventions of the hot path.

The reference stores every per-block array as a Fortran (column-major) array with
its own lower bounds (``src/modules/block.F90:205-752``, allocation in
``src/initFlow/initializeFlow.F90:457-530,686-722``).  ``REF_EXTENTS`` restates
those bounds; they are what the C ABI (``include/adflow_b200.h``) accepts.

On the Python side a ``HostBlock`` keeps every array in one *uniform box*
``(0:ib, 0:jb, 0:kb)`` (numpy ``order='F'``), so the Fortran index ``(i,j,k)`` is the
numpy index; ``ref(name)`` cuts out the contiguous reference-extent array that is
handed to the C ABI, exactly what ``c_loc`` of the Fortran allocatable would give.
"""
import numpy as np


class BlockDims:
    """nx,ny,nz owned cells; il=nx+1, ie=nx+2, ib=nx+3 (src/modules/block.F90:209-223)."""

    def __init__(self, nx, ny, nz):
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.il, self.jl, self.kl = self.nx + 1, self.ny + 1, self.nz + 1
        self.ie, self.je, self.ke = self.nx + 2, self.ny + 2, self.nz + 2
        self.ib, self.jb, self.kb = self.nx + 3, self.ny + 3, self.nz + 3
        self.box = (self.ib + 1, self.jb + 1, self.kb + 1)

    @property
    def ncells(self):
        return self.nx * self.ny * self.nz

    def owned(self):
        return (slice(2, self.il + 1), slice(2, self.jl + 1), slice(2, self.kl + 1))

    def ref_slices(self, name):
        """Reference (lower:upper) bounds of array `name` as numpy slices into the box."""
        d = self
        c2 = (slice(0, d.ib + 1), slice(0, d.jb + 1), slice(0, d.kb + 1))
        c1 = (slice(1, d.ie + 1), slice(1, d.je + 1), slice(1, d.ke + 1))
        c0 = d.owned()
        table = {
            "w": c2, "p": c2, "rlv": c2, "rev": c2, "vol": c2, "volRef": c2, "dw": c2, "fw": c2,
            "iblank": c2, "aa": c2,
            "dtl": c1, "radI": c1, "radJ": c1, "radK": c1,
            "d2Wall": c0, "wn": c0, "pn": c0,
            "x": (slice(0, d.ie + 1), slice(0, d.je + 1), slice(0, d.ke + 1)),
            "si": (slice(0, d.ie + 1), slice(1, d.je + 1), slice(1, d.ke + 1)),
            "sj": (slice(1, d.ie + 1), slice(0, d.je + 1), slice(1, d.ke + 1)),
            "sk": (slice(1, d.ie + 1), slice(1, d.je + 1), slice(0, d.ke + 1)),
            "porI": (slice(1, d.il + 1), slice(2, d.jl + 1), slice(2, d.kl + 1)),
            "porJ": (slice(2, d.il + 1), slice(1, d.jl + 1), slice(2, d.kl + 1)),
            "porK": (slice(2, d.il + 1), slice(2, d.jl + 1), slice(1, d.kl + 1)),
        }
        return table[name]


# number of trailing components of each array class
NCOMP = {"x": 3, "si": 3, "sj": 3, "sk": 3, "fw": 5, "wn": 5, "dss": 3, "grad": 12, "scratch": 10, "wallTau": 27,
         "wr": 5, "w1": 5}


class HostBlock:
    """All per-block host arrays in uniform boxes (numpy, Fortran order)."""

    REAL = ["p", "rlv", "rev", "vol", "volRef", "d2Wall", "ss", "aa", "radI", "radJ", "radK", "dtl", "pn", "shock", "p1"]
    VEC = ["x", "si", "sj", "sk", "fw", "wn", "dss", "grad", "scratch", "wallTau", "wr", "w1"]

    def __init__(self, nx, ny, nz, nw=6, right_handed=True):
        self.d = BlockDims(nx, ny, nz)
        self.nw = nw
        self.right_handed = bool(right_handed)
        box = self.d.box
        self.w = np.zeros(box + (nw,), order="F")
        self.dw = np.zeros(box + (nw,), order="F")
        for n in self.REAL:
            setattr(self, n, np.zeros(box, order="F"))
        for n in self.VEC:
            setattr(self, n, np.zeros(box + (NCOMP[n],), order="F"))
        self.porI = np.full(box, 1, dtype=np.int8, order="F")
        self.porJ = np.full(box, 1, dtype=np.int8, order="F")
        self.porK = np.full(box, 1, dtype=np.int8, order="F")
        self.iblank = np.ones(box, dtype=np.int32, order="F")
        self.d2Wall[...] = 1.0
        self.subfaces = []  # list of dicts: bcType, faceId, icBeg.., norm (ndarray), ...
        # multigrid (src/modules/block.F90 mgIFine ... mgKCoarse): level 1 = finest; a coarse block carries the
        # restriction tables mg{I,J,K}Fine (1:ie, 2) / mg{I,J,K}Weight (2:il) towards ITS fine block, a fine block the
        # interpolation tables mg{I,J,K}Coarse (2:il, 2) towards its coarse block; stored with the Fortran index as the
        # numpy index (row 0 / rows 0-1 unused)
        self.level = 1
        self.mg = {}

    def ref(self, name):
        """Contiguous Fortran-order copy of `name` with the reference's extents."""
        a = getattr(self, name)
        sl = self.d.ref_slices(name)
        if a.ndim == 4:
            sl = sl + (slice(None),)
        return np.asfortranarray(a[sl])

    def set_ref(self, name, arr):
        """Inverse of ref(): scatter a reference-extent array back into the box."""
        a = getattr(self, name)
        sl = self.d.ref_slices(name)
        if a.ndim == 4:
            sl = sl + (slice(None),)
        a[sl] = arr

    def copy(self):
        import copy

        return copy.deepcopy(self)
