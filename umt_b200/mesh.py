"""Synthetic mesh generators that hand the sweep library the arrays Teton's
``Geometry`` module would hold.

The reference gets its meshes from third-party code that is not in its tree
(Conduit ``blueprint::mesh::examples::tiled`` at test_driver.cc:1787,1816 and
MFEM at test_driver.cc:1000-1156) and turns them into corner connectivity in
``TetonBlueprint.cc:774-853,1178-1221`` + ``aux/setTetonZone.F90:89-191`` +
``aux/setSharedFace.F90:40-72`` + ``aux/setOppositeFace.F90``.  This module is
the stand-in for that input side: it produces the *same kind* of records
(1-based ids, group of arrays named like ``Geometry_mod.F90:19-78``), vectorised
with numpy so a 192 000-zone domain takes seconds.

All integer ids are 1-based exactly as Teton's Fortran holds them; 2-D arrays
are stored C-contiguous with the Fortran *first* index last, e.g. Fortran
``cFP(maxcf, ncornr)`` is ``cFP[ncornr, maxcf]`` here, so ``.ravel()`` is the
Fortran memory image.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

BC_REFL = 32   # flags_mod.F90 bcType_refl
BC_SHARED = 33
BC_TEMP = 34
BC_VAC = 35    # flags_mod.F90:30

# ---------------------------------------------------------------------------
# The 24-quad tile (img/tile.jpg; README.md:75,79).  33 points on a 20x20 grid.
# ---------------------------------------------------------------------------
TILE_X = np.array([0, 3, 10, 17, 20, 0, 3, 17, 20, 5, 15, 7, 10, 13,
                   0, 7, 10, 13, 20, 7, 10, 13, 5, 15, 0, 3, 17, 20,
                   0, 3, 10, 17, 20], dtype=np.int64)
TILE_Y = np.array([0, 0, 0, 0, 0, 3, 3, 3, 3, 5, 5, 7, 7, 7,
                   10, 10, 10, 10, 10, 13, 13, 13, 15, 15, 17, 17, 17, 17,
                   20, 20, 20, 20, 20], dtype=np.int64)
TILE_QUADS = np.array([
    [0, 1, 6, 5], [1, 2, 9, 6], [2, 12, 11, 9], [5, 6, 9, 14], [9, 11, 15, 14], [11, 12, 16, 15],
    [2, 3, 7, 10], [3, 4, 8, 7], [7, 8, 18, 10], [2, 10, 13, 12], [12, 13, 17, 16], [10, 18, 17, 13],
    [14, 22, 25, 24], [14, 15, 19, 22], [15, 16, 20, 19], [24, 25, 29, 28], [22, 30, 29, 25], [19, 20, 30, 22],
    [16, 17, 21, 20], [18, 23, 21, 17], [18, 27, 26, 23], [20, 21, 23, 30], [23, 26, 31, 30], [26, 27, 32, 31],
], dtype=np.int64)
TILE_W = 20

# Local topology templates.  Zone-local corner i sits on zone node i.
# Hex nodes: 0-3 bottom ring counter-clockwise seen from +z, 4-7 the ring above.
# Faces are listed with the right-hand-rule normal pointing OUT of the zone; the
# Teton record wants the opposite sense (TetonBlueprint.cc:1178-1206), which
# ``_teton_face_lists`` applies and ``build_teton_mesh`` verifies numerically.
HEX_FACES_OUT = np.array([
    [0, 4, 7, 3],  # x-
    [1, 2, 6, 5],  # x+
    [0, 1, 5, 4],  # y-
    [3, 7, 6, 2],  # y+
    [0, 3, 2, 1],  # z-
    [4, 5, 6, 7],  # z+
], dtype=np.int64)
# Quad nodes counter-clockwise; face k joins node k and k+1.
QUAD_FACES = np.array([[0, 1], [1, 2], [2, 3], [3, 0]], dtype=np.int64)


@dataclass
class Boundary:
    """One entry of Teton's BoundaryList (mods/Boundary_mod.F90)."""
    bc_type: int
    n_elem: int
    first_elem: int            # 1-based BdyElem1
    neighbor: int = -1         # rank of the neighbour for shared boundaries
    side: int = -1             # which box side (0..5) it came from


@dataclass
class TetonMesh:
    ndim: int
    nzones: int
    ncornr: int
    nbelem: int
    maxcf: int
    maxCorner: int
    maxFaces: int
    numCorner: np.ndarray      # (nz,)
    cOffSet: np.ndarray        # (nz,) 0-based offset of the zone's first corner
    zoneFaces: np.ndarray      # (nz,)
    zoneOpp: np.ndarray        # (nz, maxFaces)  1-based, <0 on a boundary
    faceOpp: np.ndarray        # (nz, maxFaces)  1-based, -1 on a boundary
    nCFacesArray: np.ndarray   # (nc,)
    cFP: np.ndarray            # (nc, maxcf) global corner (1-based) or nc + b
    cEZ: np.ndarray            # (nc, maxcf) zone-local corner (1-based)
    CToFace: np.ndarray        # (nc, maxcf) zone-local face (1-based)
    CToZone: np.ndarray        # (nc,) 1-based
    px: np.ndarray             # (nc, ndim) corner (= node) coordinates
    BoundaryZone: np.ndarray   # (nz,) bool
    boundaries: List[Boundary]
    BdyToC: np.ndarray         # (nb,) global corner (1-based) of boundary element b
    BdyToZone: np.ndarray      # (nb,)
    BdyToBC: np.ndarray        # (nb,) index into boundaries (0-based)
    corner_node: np.ndarray    # (nc,) node id, for tests
    node_key: Optional[np.ndarray] = None  # (nnodes, ndim) integer lattice key (global)
    info: dict = field(default_factory=dict)


# ---------------------------------------------------------------------------
# generic builder
# ---------------------------------------------------------------------------

def _local_templates(ndim: int):
    """Per-zone-local cEZ / CToFace / (face, slot) tables for the uniform cell,
    following the chaining rule of setTetonZone.F90:124-191."""
    if ndim == 2:
        ncl, nfl, maxcf = 4, 4, 2
        # Teton face record (c1, c2): cFP(2,c1), cFP(1,c2), cEZ(1,c1)=c2, cEZ(2,c2)=c1
        # (setTetonZone.F90:124-141).  With the geometry of geometryUCBrz.F90
        # (A_fp(:,2,c) = 90deg CCW rotation of p(cEZ1)-p(c)) the outward normal
        # requires the face to be traversed clockwise around the zone.
        faces = QUAD_FACES[:, ::-1].copy()   # (c1, c2) = (k+1, k): clockwise
        cEZ = np.zeros((ncl, maxcf), np.int64)
        CToFace = np.zeros((ncl, maxcf), np.int64)
        slot = np.zeros((ncl, maxcf), np.int64)   # position in the face record
        for f in range(nfl):
            c1, c2 = faces[f]
            cEZ[c1, 0] = c2 + 1
            cEZ[c2, 1] = c1 + 1
            CToFace[c1, 1] = f + 1
            CToFace[c2, 0] = f + 1
            slot[c1, 1] = 0
            slot[c2, 0] = 1
        return faces, cEZ, CToFace, slot
    # 3-D hex
    faces = HEX_FACES_OUT[:, [0, 3, 2, 1]].copy()   # reversed sense: normal points into the zone
    ncl, maxcf = 8, 3
    numC = 4
    ent = [[] for _ in range(ncl)]   # per corner: (cCW, cCCW, face, slot)
    for f in range(6):
        for i in range(numC):
            iCCW = (i - 1) % numC
            iCW = (i + 1) % numC
            c = faces[f, i]
            ent[c].append((faces[f, iCW], faces[f, iCCW], f, i))
    cEZ = np.zeros((ncl, maxcf), np.int64)
    CToFace = np.zeros((ncl, maxcf), np.int64)
    slot = np.zeros((ncl, maxcf), np.int64)
    for c in range(ncl):
        e = ent[c]
        assert len(e) == 3
        cEZ[c, 0] = e[0][1] + 1
        CToFace[c, 0] = e[0][2] + 1
        slot[c, 0] = e[0][3]
        last = e[0][0]
        for i in range(1, 3):
            for ii in range(1, 3):
                if e[ii][1] == last:
                    cEZ[c, i] = e[ii][1] + 1
                    CToFace[c, i] = e[ii][2] + 1
                    slot[c, i] = e[ii][3]
                    last = e[ii][0]
                    break
            else:  # pragma: no cover
                raise RuntimeError("corner faces do not chain")
    return faces, cEZ, CToFace, slot


def build_teton_mesh(coords: np.ndarray, zones: np.ndarray,
                     bface_side_fn, sides: Sequence[Tuple[int, int]],
                     node_key: Optional[np.ndarray] = None) -> TetonMesh:
    """coords (nn, ndim); zones (nz, 4|8) node ids in the template order.

    ``bface_side_fn(face_center (m, ndim), face_nodes (m, k)) -> side id (m,)``
    classifies boundary half-faces; ``sides[s] = (bc_type, neighbor_rank)``.
    Non-shared boundaries are numbered first (by side id), then shared ones.
    """
    ndim = coords.shape[1]
    nz, ncl = zones.shape
    faces, cEZ_l, CToFace_l, slot_l = _local_templates(ndim)
    nfl, nfc = faces.shape
    maxcf = 2 if ndim == 2 else 3
    nc = nz * ncl

    # orientation check on the first zone; flip a left-handed template input
    zn = coords[zones]                                  # (nz, ncl, ndim)
    if ndim == 2:
        a = zn[:, 1] - zn[:, 0]
        b = zn[:, 3] - zn[:, 0]
        area = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
        if np.median(area) <= 0:
            raise ValueError("quads must be counter-clockwise")
    else:
        a = zn[:, 1] - zn[:, 0]
        b = zn[:, 3] - zn[:, 0]
        c = zn[:, 4] - zn[:, 0]
        vol = np.einsum('ij,ij->i', np.cross(a, b), c)
        if np.median(vol) <= 0:   # (median: strongly warped test meshes may have a few skewed corners)
            raise ValueError("hexes must be right-handed (VTK order)")

    # half-faces: (zone, local face) -> node list in Teton order
    hf_nodes = zones[:, faces]                          # (nz, nfl, nfc)
    key = np.sort(hf_nodes.reshape(nz * nfl, nfc), axis=1)
    order = np.lexsort(key.T[::-1])
    ks = key[order]
    same = np.all(ks[1:] == ks[:-1], axis=1)
    if np.any(same[1:] & same[:-1]):
        raise ValueError("a face is shared by more than two zones")
    opp = np.full(nz * nfl, -1, np.int64)
    i0 = order[:-1][same]
    i1 = order[1:][same]
    opp[i0] = i1
    opp[i1] = i0
    zoneOpp = np.where(opp >= 0, opp // nfl + 1, -1).reshape(nz, nfl)
    faceOpp = np.where(opp >= 0, opp % nfl + 1, -1).reshape(nz, nfl)

    # corner across each half-face slot: the neighbour's corner on the same node
    cornerOpp = np.zeros((nz * nfl, nfc), np.int64)     # global 1-based
    has = opp >= 0
    hfn = hf_nodes.reshape(nz * nfl, nfc)
    nb_nodes = hfn[opp[has]]                            # (m, nfc) neighbour's list
    nb_zone = opp[has] // nfl
    nb_face = opp[has] % nfl
    mine = hfn[has]
    match = mine[:, :, None] == nb_nodes[:, None, :]    # (m, nfc, nfc)
    assert np.all(match.sum(axis=2) == 1)
    j = np.argmax(match, axis=2)                        # slot in neighbour's list
    nb_local_corner = faces[nb_face[:, None], j]        # (m, nfc)
    cornerOpp[has] = nb_zone[:, None] * ncl + nb_local_corner + 1

    # boundary half-faces -> boundaries
    bidx = np.nonzero(~has)[0]                          # sorted: zone-major, face-minor
    fc = coords[hfn[bidx]].mean(axis=1)
    side = np.asarray(bface_side_fn(fc, hfn[bidx]), np.int64)
    used = np.unique(side)
    nonshared = [s for s in used if sides[s][0] != BC_SHARED]
    shared = [s for s in used if sides[s][0] == BC_SHARED]
    boundaries: List[Boundary] = []
    nbelem = 0
    BdyToC, BdyToZone, BdyToBC = [], [], []
    for s in nonshared + shared:
        sel = bidx[side == s]
        if sides[s][0] == BC_SHARED:
            # both domains must list shared corner-faces in the same order
            # (checkSharedBoundary.F90): sort by the global lattice key.
            assert node_key is not None, "shared boundaries need node_key"
            fk = np.sort(_pack_key(node_key[hfn[sel]]), axis=1)
            fo = np.lexsort(fk.T[::-1])
            sel = sel[fo]
            nk = _pack_key(node_key[hfn[sel]])          # (m, nfc)
            so = np.argsort(nk, axis=1, kind='stable')
        else:
            so = np.tile(np.arange(nfc), (len(sel), 1))  # face-record order (setTetonZone.F90:106-112)
        m = len(sel)
        elem = nbelem + 1 + np.arange(m * nfc).reshape(m, nfc)   # 1-based global bdy element
        rows = np.arange(m)[:, None]
        cornerOpp[sel[:, None], so] = nc + elem
        zl = sel // nfl
        fl = sel % nfl
        cg = zl[:, None] * ncl + faces[fl[:, None], so] + 1
        BdyToC.append(cg.ravel())
        BdyToZone.append(np.repeat(zl + 1, nfc))
        BdyToBC.append(np.full(m * nfc, len(boundaries)))
        boundaries.append(Boundary(sides[s][0], m * nfc, nbelem + 1, sides[s][1], int(s)))
        nbelem += m * nfc
        del rows

    # per-corner tables from the local templates
    zi = np.arange(nz)
    cFP = np.zeros((nz, ncl, maxcf), np.int64)
    for c in range(ncl):
        for k in range(maxcf):
            f = CToFace_l[c, k] - 1
            cFP[:, c, k] = cornerOpp.reshape(nz, nfl, nfc)[zi, f, slot_l[c, k]]
    cEZ = np.broadcast_to(cEZ_l, (nz, ncl, maxcf)).reshape(nc, maxcf).copy()
    CToFace = np.broadcast_to(CToFace_l, (nz, ncl, maxcf)).reshape(nc, maxcf).copy()

    mesh = TetonMesh(
        ndim=ndim, nzones=nz, ncornr=nc, nbelem=nbelem, maxcf=maxcf, maxCorner=ncl, maxFaces=nfl,
        numCorner=np.full(nz, ncl, np.int32), cOffSet=(zi * ncl).astype(np.int32),
        zoneFaces=np.full(nz, nfl, np.int32),
        zoneOpp=zoneOpp.astype(np.int32), faceOpp=faceOpp.astype(np.int32),
        nCFacesArray=np.full(nc, maxcf, np.int32),
        cFP=cFP.reshape(nc, maxcf).astype(np.int32), cEZ=cEZ.astype(np.int32),
        CToFace=CToFace.astype(np.int32), CToZone=np.repeat(zi + 1, ncl).astype(np.int32),
        px=np.ascontiguousarray(coords[zones].reshape(nc, ndim), dtype=np.float64),
        BoundaryZone=np.any(zoneOpp < 0, axis=1),
        boundaries=boundaries,
        BdyToC=(np.concatenate(BdyToC) if BdyToC else np.zeros(0)).astype(np.int32),
        BdyToZone=(np.concatenate(BdyToZone) if BdyToZone else np.zeros(0)).astype(np.int32),
        BdyToBC=(np.concatenate(BdyToBC) if BdyToBC else np.zeros(0)).astype(np.int32),
        corner_node=zones.reshape(nc).astype(np.int64),
        node_key=node_key,
    )
    return mesh


def _pack_key(k: np.ndarray) -> np.ndarray:
    """(…, ndim) small non-negative lattice coordinates -> one int64 key."""
    out = np.zeros(k.shape[:-1], np.int64)
    for d in range(k.shape[-1]):
        out = out * (1 << 20) + k[..., d]
    return out


# ---------------------------------------------------------------------------
# domain decomposition of the driver (test_driver.cc:1741-1756)
# ---------------------------------------------------------------------------

def _factor(n: int) -> List[int]:
    f, p = [], 2
    while n > 1:
        while n % p == 0:
            f.append(p)
            n //= p
        p += 1
    return f


def decompose(rank: int, size: int, ndims: int):
    domains = [1, 1, 1]
    fac = _factor(size)
    for i in range(len(fac)):
        dim = (ndims - 1) - (i % ndims)
        domains[dim] *= fac[len(fac) - 1 - i]
    domainid = [rank % domains[0],
                (rank % (domains[0] * domains[1])) // domains[0],
                rank // (domains[0] * domains[1])]
    return domainid, domains


def _box_sides(ndim, domainid, domains, lo, hi, tol, reflecting=()):
    """side ids 0..2*ndim-1 = (x-, x+, y-, y+, z-, z+); shared where a neighbour exists,
    reflecting (bcType_refl) where asked on an outer side, vacuum otherwise."""
    sides = []
    for d in range(ndim):
        for s in (0, 1):
            nb = list(domainid)
            nb[d] += -1 if s == 0 else 1
            if 0 <= nb[d] < domains[d]:
                r = nb[0] + domains[0] * (nb[1] + domains[1] * nb[2])
                sides.append((BC_SHARED, r))
            else:
                sides.append((BC_REFL if (2 * d + s) in reflecting else BC_VAC, -1))

    def classify(fc, _nodes):
        out = np.full(len(fc), -1, np.int64)
        for d in range(ndim):
            out[np.abs(fc[:, d] - lo[d]) < tol] = 2 * d
            out[np.abs(fc[:, d] - hi[d]) < tol] = 2 * d + 1
        assert np.all(out >= 0)
        return out
    return sides, classify


# ---------------------------------------------------------------------------
# concrete meshes
# ---------------------------------------------------------------------------

def tiled_mesh(dims: Sequence[int], rank: int = 0, size: int = 1,
               extents=(0., 1., 0., 1., 0., 1.), reflecting=()) -> TetonMesh:
    """The driver's ``-B local -d nx,ny,nz`` mesh (test_driver.cc:1787-1816):
    every domain holds nx x ny tiles (x nz layers); the unit box is split
    between ``size`` domains by ``decompose``.  nz == 0 gives the 2-D (r,z) mesh."""
    nx, ny, nzl = int(dims[0]), int(dims[1]), int(dims[2]) if len(dims) > 2 else 0
    ndim = 3 if nzl > 0 else 2
    domainid, domains = decompose(rank, size, ndim)
    lo = np.array([extents[0], extents[2], extents[4]][:ndim])
    full = np.array([extents[1] - extents[0], extents[3] - extents[2], extents[5] - extents[4]][:ndim])
    side_len = full / np.array(domains[:ndim], float)
    dlo = lo + np.array(domainid[:ndim]) * side_len
    dhi = dlo + side_len

    # 2-D lattice of tile points, de-duplicated on integer coordinates
    tx = (np.arange(nx)[:, None] * TILE_W + TILE_X[None, :])          # (nx, 33)
    ty = (np.arange(ny)[:, None] * TILE_W + TILE_Y[None, :])          # (ny, 33)
    gx = np.broadcast_to(tx[None, :, :], (ny, nx, 33)).reshape(-1)
    gy = np.broadcast_to(ty[:, None, :], (ny, nx, 33)).reshape(-1)
    key = gy * (TILE_W * nx + 1) + gx
    ukey, inv = np.unique(key, return_inverse=True)
    n2 = len(ukey)
    ix = ukey % (TILE_W * nx + 1)
    iy = ukey // (TILE_W * nx + 1)
    quads = inv.reshape(ny * nx, 33)[:, TILE_QUADS].reshape(-1, 4)    # tile-major
    x2 = dlo[0] + ix * (side_len[0] / (TILE_W * nx))
    y2 = dlo[1] + iy * (side_len[1] / (TILE_W * ny))
    gkx = ix + domainid[0] * TILE_W * nx
    gky = iy + domainid[1] * TILE_W * ny
    tol = 1e-9 * float(side_len.min())
    if ndim == 2:
        coords = np.stack([x2, y2], axis=1)
        nkey = np.stack([gkx, gky], axis=1)
        sides, classify = _box_sides(2, domainid, domains, dlo, dhi, tol, reflecting)
        m = build_teton_mesh(coords, quads, classify, sides, nkey)
    else:
        zs = dlo[2] + np.arange(nzl + 1) * (side_len[2] / nzl)
        coords = np.stack([np.tile(x2, nzl + 1), np.tile(y2, nzl + 1), np.repeat(zs, n2)], axis=1)
        nkey = np.stack([np.tile(gkx, nzl + 1), np.tile(gky, nzl + 1),
                         np.repeat(np.arange(nzl + 1) + domainid[2] * nzl, n2)], axis=1)
        lay = np.arange(nzl)[:, None, None] * n2
        hexes = np.concatenate([quads[None] + lay, quads[None] + lay + n2], axis=2).reshape(-1, 8)
        sides, classify = _box_sides(3, domainid, domains, dlo, dhi, tol, reflecting)
        m = build_teton_mesh(coords, hexes, classify, sides, nkey)
    m.info.update(kind="tiled", dims=(nx, ny, nzl), rank=rank, size=size,
                  domainid=domainid, domains=domains)
    return m


def box_mesh(n: Sequence[int], lengths=None, rank: int = 0, size: int = 1,
             warp: float = 0.0, seed: int = 0, reflecting=()) -> TetonMesh:
    """Structured quad/hex box (n cells per side).  ``warp`` > 0 displaces the
    interior nodes randomly (seeded) by that fraction of a cell so that faces
    become non-planar and corner faces on one zone face can change sign — the
    way to exercise the cycle-list and Jacobi branches (snneed.F90:156-183,
    getDownStreamData.F90:124-149)."""
    n = [int(v) for v in n if int(v) > 0]
    ndim = len(n)
    domainid, domains = decompose(rank, size, ndim)
    lengths = np.array(lengths if lengths is not None else [1.0] * ndim, float)
    side_len = lengths / np.array(domains[:ndim], float)
    dlo = np.array(domainid[:ndim]) * side_len
    dhi = dlo + side_len
    ax = [np.arange(k + 1) for k in n]
    if ndim == 2:
        J, I = np.meshgrid(ax[1], ax[0], indexing='ij')
        ik = np.stack([I.ravel(), J.ravel()], axis=1)
    else:
        K, J, I = np.meshgrid(ax[2], ax[1], ax[0], indexing='ij')
        ik = np.stack([I.ravel(), J.ravel(), K.ravel()], axis=1)
    h = side_len / np.array(n, float)
    coords = dlo + ik * h
    nkey = ik + np.array(domainid[:ndim]) * np.array(n)
    if warp > 0:
        # displacement is a function of the global lattice key so neighbouring
        # domains move shared nodes identically; box surface nodes stay put.
        gmax = np.array(n) * np.array(domains[:ndim])
        interior = np.all((nkey > 0) & (nkey < gmax), axis=1)
        packed = _pack_key(nkey).astype(np.uint64)
        rng_vals = np.empty((len(packed), ndim))
        with np.errstate(over='ignore'):
            for d in range(ndim):
                x = packed * np.uint64(6364136223846793005) + np.uint64(1442695040888963407) * np.uint64(seed * 3 + d + 1)
                x ^= x >> np.uint64(29)
                x = x * np.uint64(2862933555777941757) + np.uint64(3037000493)
                x ^= x >> np.uint64(32)
                rng_vals[:, d] = (x >> np.uint64(11)).astype(np.float64) / float(1 << 53) - 0.5
        coords = coords + warp * h * rng_vals * interior[:, None]

    def nid(*idx):
        if ndim == 2:
            i, j = idx
            return j * (n[0] + 1) + i
        i, j, k = idx
        return (k * (n[1] + 1) + j) * (n[0] + 1) + i
    if ndim == 2:
        J, I = np.meshgrid(np.arange(n[1]), np.arange(n[0]), indexing='ij')
        I, J = I.ravel(), J.ravel()
        zones = np.stack([nid(I, J), nid(I + 1, J), nid(I + 1, J + 1), nid(I, J + 1)], axis=1)
    else:
        K, J, I = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing='ij')
        I, J, K = I.ravel(), J.ravel(), K.ravel()
        zones = np.stack([nid(I, J, K), nid(I + 1, J, K), nid(I + 1, J + 1, K), nid(I, J + 1, K),
                          nid(I, J, K + 1), nid(I + 1, J, K + 1), nid(I + 1, J + 1, K + 1), nid(I, J + 1, K + 1)], axis=1)
    sides, classify = _box_sides(ndim, domainid, domains, dlo, dhi, 1e-9 * float(h.min()), reflecting)
    m = build_teton_mesh(coords, zones, classify, sides, nkey)
    m.info.update(kind="box", dims=tuple(n), rank=rank, size=size, domainid=domainid, domains=domains, warp=warp)
    return m


# The 12-hex / 36-vertex base of unstructBox3D.mesh (driver/makeUnstructuredBox.cc:39-110):
# twelve 2-D vertices a..l (in sixths of the box width), six quads per layer,
# two layers of thickness 1/3.  Vertex f has valence 5, so the mesh is not a lattice.
_UB_XY6 = np.array([[0, 0], [3, 0], [6, 0], [6, 2], [0, 3], [2, 3],
                    [4, 4], [6, 4], [0, 6], [2, 6], [4, 6], [6, 6]], dtype=np.int64)
_UB_QUADS = np.array([[0, 1, 5, 4], [1, 2, 3, 5], [4, 5, 9, 8],
                      [5, 6, 10, 9], [5, 3, 7, 6], [6, 7, 11, 10]], dtype=np.int64)
_UB_LAYERS = 2


def unstruct_box_mesh(refine: int = 6) -> TetonMesh:
    """unstructBox3D base mesh with every hex split ``refine`` times per edge by
    (bi/tri)linear subdivision — what MFEM's ``Mesh::MakeRefined`` with a closed
    uniform basis does to a hex (test_driver.cc:1000-1156).  Single domain."""
    R = int(refine)
    i = np.arange(R + 1)
    U, V = np.meshgrid(i, i, indexing='xy')             # U varies fastest along axis 1
    pts, quads = [], []
    for q in _UB_QUADS:
        A, B, C, D = _UB_XY6[q]
        # exact integer coordinates in units of 1/(6 R^2)
        P = ((R - U)[..., None] * (R - V)[..., None] * A + U[..., None] * (R - V)[..., None] * B
             + U[..., None] * V[..., None] * C + (R - U)[..., None] * V[..., None] * D)
        base = len(pts) * (R + 1) ** 2
        pts.append(P.reshape(-1, 2))
        jj, ii = np.meshgrid(np.arange(R), np.arange(R), indexing='ij')
        n00 = base + jj * (R + 1) + ii
        quads.append(np.stack([n00, n00 + 1, n00 + R + 2, n00 + R + 1], axis=-1).reshape(-1, 4))
    P = np.concatenate(pts)
    quads = np.concatenate(quads)
    S = 6 * R * R
    key = P[:, 1] * (S + 1) + P[:, 0]
    ukey, inv = np.unique(key, return_inverse=True)
    quads = inv[quads]
    ix = ukey % (S + 1)
    iy = ukey // (S + 1)
    n2 = len(ukey)
    nl = _UB_LAYERS * R
    x2 = ix / float(S)
    y2 = iy / float(S)
    zs = np.arange(nl + 1) * (1.0 / 3.0 / R)
    coords = np.stack([np.tile(x2, nl + 1), np.tile(y2, nl + 1), np.repeat(zs, n2)], axis=1)
    ik = np.stack([np.tile(ix, nl + 1), np.tile(iy, nl + 1), np.repeat(np.arange(nl + 1), n2)], axis=1)
    lay = np.arange(nl)[:, None, None] * n2
    hexes = np.concatenate([quads[None] + lay, quads[None] + lay + n2], axis=2).reshape(-1, 8)
    gmax = np.array([S, S, nl])

    def classify(fc, nodes):
        k = ik[nodes]
        out = np.full(len(fc), -1, np.int64)
        for d in range(3):
            out[np.all(k[:, :, d] == 0, axis=1)] = 2 * d
            out[np.all(k[:, :, d] == gmax[d], axis=1)] = 2 * d + 1
        assert np.all(out >= 0)
        return out
    sides = [(BC_VAC, -1)] * 6
    m = build_teton_mesh(coords, hexes, classify, sides, ik)
    m.info.update(kind="unstructBox3D", refine=R, base_hexes=12)
    return m
