#!/usr/bin/env python
"""Generates transhuman_b200/csrc/mc_table.h: the 256-case triangle table of marching cubes.

The reference extracts its mesh with the third-party `mcubes.marching_cubes` (if_mesh_renderer.py:98-104), which is not
available here (nor are its tables), so the table is DERIVED instead of copied:

  * corner / edge numbering: the usual one (corner i at (x,y,z) = CORNERS[i], edge e between EDGES[e]),
  * a corner is "inside" when its value is > the iso level; an edge is cut when its two corners differ,
  * on each of the 6 faces the cut edges are joined pairwise; a face with four cut edges (inside corners on a
    diagonal) is resolved by cutting off each INSIDE corner.  The rule reads only the face's own four corners, so
    the two cubes sharing a face join its cut edges identically: the surface has no cracks,
  * the joins form closed loops through the cut edges; each loop is oriented so that its normal points from the
    inside corners to the outside corners, and triangulated without any diagonal that joins two cut edges of one
    cube face (such a diagonal lies in the face plane and the neighbouring cube may create it too: four triangles
    on one edge).

oracle/marching_cubes.py (numpy) and csrc/mesh.cu (CUDA) both read this one table; tests/test_marching_cubes.py
checks it (every cut edge used, closed loops, watertight meshes on random volumes, Euler characteristic of a sphere).
"""
import os

import numpy as np

CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
# faces as corner cycles
FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (3, 2, 6, 7), (0, 3, 7, 4), (1, 2, 6, 5)]
EDGE_OF = {frozenset(e): i for i, e in enumerate(EDGES)}


def case_triangles(mask: int):
    inside = [(mask >> i) & 1 for i in range(8)]
    nbr = {}   # cut edge -> the cut edges it is joined to (one per face it lies on)

    def join(a, b):
        nbr.setdefault(a, []).append(b)
        nbr.setdefault(b, []).append(a)

    for face in FACES:
        cyc = [(face[i], face[(i + 1) % 4]) for i in range(4)]
        cut = [EDGE_OF[frozenset(c)] for c in cyc if inside[c[0]] != inside[c[1]]]
        if len(cut) == 2:
            join(cut[0], cut[1])
        elif len(cut) == 4:
            # inside corners on a diagonal: cut off each inside corner (join the two face edges that meet in it)
            for i in range(4):
                c = face[i]
                if inside[c]:
                    e_prev = EDGE_OF[frozenset((face[(i - 1) % 4], c))]
                    e_next = EDGE_OF[frozenset((c, face[(i + 1) % 4]))]
                    join(e_prev, e_next)
    for e, n in nbr.items():
        assert len(n) == 2, (mask, e, n)
    # closed loops
    loops, seen = [], set()
    for start in sorted(nbr):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            a, b = nbr[cur]
            nxt = a if a != prev else b
            if len(loop) == 1 and prev is None:
                nxt = a
            if nxt == start:
                break
            loop.append(nxt)
            seen.add(nxt)
            prev, cur = cur, nxt
        loops.append(loop)
    mid = {e: (np.array(CORNERS[EDGES[e][0]], float) + np.array(CORNERS[EDGES[e][1]], float)) / 2 for e in range(12)}
    tris = []
    for loop in loops:
        assert len(loop) >= 3, (mask, loop)
        # Newell normal of the loop against the inside -> outside direction of its edges
        n = np.zeros(3)
        for i in range(len(loop)):
            p, q = mid[loop[i]], mid[loop[(i + 1) % len(loop)]]
            n += np.cross(p, q)
        d = np.zeros(3)
        for e in loop:
            a, b = EDGES[e]
            ins, out = (a, b) if inside[a] else (b, a)
            d += np.array(CORNERS[out], float) - np.array(CORNERS[ins], float)
        assert abs(n @ d) > 1e-9, (mask, loop)
        if n @ d < 0:
            loop = loop[::-1]
        tris.extend(triangulate(loop))
    return tris


FACE_EDGES = [frozenset(EDGE_OF[frozenset((f[i], f[(i + 1) % 4]))] for i in range(4)) for f in FACES]


def on_common_face(a, b):
    return any(a in fe and b in fe for fe in FACE_EDGES)


def triangulate(loop):
    """A triangulation of the loop whose interior diagonals never join two cut edges of one cube face: such a
    diagonal would lie in the face plane, the neighbouring cube could create the same one, and four triangles
    would share it.  Polygons have at most 7 corners here, so all triangulations are enumerated."""
    n = len(loop)

    def rec(i, j):      # triangulations of the sub-polygon loop[i..j] (a chain closed by the diagonal i-j)
        if j - i < 2:
            return [[]]
        out = []
        for k in range(i + 1, j):
            ok = True
            for (u, v) in ((i, k), (k, j)):
                if v - u > 1 and on_common_face(loop[u], loop[v]):
                    ok = False
            if not ok:
                continue
            for left in rec(i, k):
                for right in rec(k, j):
                    out.append(left + [(loop[i], loop[k], loop[j])] + right)
        return out

    sols = rec(0, n - 1)   # every triangulation has a triangle on the side (0, n-1): this enumerates them all
    assert sols, ("no admissible triangulation", loop)
    return sols[0]


def build():
    table = np.full((256, 16), -1, dtype=np.int8)
    count = np.zeros(256, dtype=np.int8)
    for mask in range(256):
        tris = case_triangles(mask)
        assert len(tris) <= 5, (mask, len(tris))
        count[mask] = len(tris)
        flat = [e for t in tris for e in t]
        table[mask, :len(flat)] = flat
    return table, count


def main():
    table, count = build()
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "transhuman_b200", "csrc", "mc_table.h")
    with open(out, "w") as f:
        f.write("// GENERATED by tools/gen_mc_table.py -- do not edit.  Marching-cubes triangle table derived from the face\n"
                "// rule described there (corner inside = value > iso; ambiguous faces cut off the inside corners).\n"
                "// MC_TRI[case][3 t .. 3 t + 2] = the cube edges of triangle t, -1 terminated; MC_NTRI[case] = count.\n"
                "#pragma once\n#include <stdint.h>\n\n")
        f.write("static const int8_t MC_NTRI_HOST[256] = {" + ", ".join(str(int(c)) for c in count) + "};\n\n")
        f.write("static const int8_t MC_TRI_HOST[256][16] = {\n")
        for m in range(256):
            f.write("    {" + ", ".join(f"{int(v):2d}" for v in table[m]) + "},\n")
        f.write("};\n")
    print(out, "max triangles", int(count.max()), "total", int(count.sum()))


if __name__ == "__main__":
    main()
