"""TEST INFRASTRUCTURE (oracle): CPU restatement, in numpy, of the mesh extraction th_marching_cubes performs.

The reference calls the third-party `mcubes.marching_cubes(cube, cfg.mesh_th)` (if_mesh_renderer.py:98-104; PyMCubes,
not under /root/reference, not pinned in requirements.txt, absent from this image): **parity with it is unpinned**.
What is restated is marching cubes itself -- one vertex per cut lattice edge at the linear interpolation
`a + (iso - v_a) / (v_b - v_a)` in index coordinates, triangles from a 256-case table -- with the table DERIVED by
tools/gen_mc_table.py (face rule there), since PyMCubes' own table is not available.  Surfaces agree with any
marching-cubes implementation except in how ambiguous faces are joined and how polygons are fanned.

Only tests/ may import this module."""
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_TABLE_H = os.path.join(os.path.dirname(_HERE), "transhuman_b200", "csrc", "mc_table.h")

# edge e of a cube = (offset of its lower lattice point, axis)
EDGE_SLOT = [((0, 0, 0), 0), ((1, 0, 0), 1), ((0, 1, 0), 0), ((0, 0, 0), 1), ((0, 0, 1), 0), ((1, 0, 1), 1),
             ((0, 1, 1), 0), ((0, 0, 1), 1), ((0, 0, 0), 2), ((1, 0, 0), 2), ((1, 1, 0), 2), ((0, 1, 0), 2)]
CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]


def load_table():
    """(MC_NTRI (256,), MC_TRI (256,16)) parsed from the generated header -- the very table the CUDA code compiles."""
    src = open(_TABLE_H).read()
    ntri = np.array([int(x) for x in re.search(r"MC_NTRI_HOST\[256\] = \{([^}]*)\}", src).group(1).split(",")], dtype=np.int64)
    body = src[src.index("MC_TRI_HOST[256][16]"):]
    rows = re.findall(r"\{([-\d,\s]+)\},", body)
    tri = np.array([[int(x) for x in r.split(",")] for r in rows], dtype=np.int64)
    assert ntri.shape == (256,) and tri.shape == (256, 16)
    return ntri, tri


def marching_cubes(vol: np.ndarray, iso: float):
    """vol (nx,ny,nz) float32 -> (vertices (n,3) float32 in index coordinates, triangles (m,3) int32).
    Vertex order: by (lattice point in C order, axis); triangle order: by (cube in C order, table order)."""
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    iso = np.float32(iso)
    nx, ny, nz = vol.shape
    inside = vol > iso
    ntri, tri = load_table()
    # vertex ids: slot = 3 * point + axis
    cut = np.zeros((nx, ny, nz, 3), dtype=bool)
    cut[:-1, :, :, 0] = inside[:-1] != inside[1:]
    cut[:, :-1, :, 1] = inside[:, :-1] != inside[:, 1:]
    cut[:, :, :-1, 2] = inside[:, :, :-1] != inside[:, :, 1:]
    flat = cut.reshape(-1)
    vid = np.full(flat.shape, -1, dtype=np.int64)
    slots = np.nonzero(flat)[0]
    vid[slots] = np.arange(len(slots))
    p, axis = slots // 3, slots % 3
    i, j, k = p // (ny * nz), (p // nz) % ny, p % nz
    va = vol[i, j, k]
    vb = vol[i + (axis == 0), j + (axis == 1), k + (axis == 2)]
    t = ((iso - va) / (vb - va)).astype(np.float32)
    verts = np.stack([i, j, k], 1).astype(np.float32)
    verts[np.arange(len(slots)), axis] += t
    # triangles
    case = np.zeros((nx - 1, ny - 1, nz - 1), dtype=np.int64)
    for c, (dx, dy, dz) in enumerate(CORNERS):
        case |= inside[dx:nx - 1 + dx, dy:ny - 1 + dy, dz:nz - 1 + dz].astype(np.int64) << c
    cubes = np.nonzero(ntri[case].reshape(-1))[0]
    ci, cj, ck = cubes // ((ny - 1) * (nz - 1)), (cubes // (nz - 1)) % (ny - 1), cubes % (nz - 1)
    cc = case.reshape(-1)[cubes]
    vid4 = vid.reshape(nx, ny, nz, 3)
    out = []
    for q in range(len(cubes)):
        for tt in range(ntri[cc[q]]):
            ids = []
            for e in tri[cc[q], 3 * tt:3 * tt + 3]:
                (dx, dy, dz), ax = EDGE_SLOT[e]
                ids.append(vid4[ci[q] + dx, cj[q] + dy, ck[q] + dz, ax])
            out.append(ids)
    tris = np.array(out, dtype=np.int32).reshape(-1, 3)
    return verts, tris
