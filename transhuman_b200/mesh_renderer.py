"""Drop-in for ``lib/networks/renderer/if_mesh_renderer.py`` (configs/
reconstruction.yaml selects it through ``renderer_module``/``renderer_path``).

``render(batch)`` runs the density-grid query (if_mesh_renderer.py:46-96) on the
GPU and returns ``{'cube': np.ndarray (X+20,Y+20,Z+20), 'mesh': trimesh or None}``
(111).  Marching cubes and the mesh object (98-109) come after the path and
stay on the CPU third-party packages when those are installed.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .renderer import Renderer as _BaseRenderer


class Renderer(_BaseRenderer):

    def query_cube(self, batch) -> torch.Tensor:
        """alpha_raw on the voxel grid, (X,Y,Z) fp32 on the device; 0 where culled."""
        pts = batch['pts']
        sh = pts.shape  # (1, X, Y, Z, 3)
        with torch.no_grad():
            frame = self.prepare_frame({**batch, 'ray_o': pts})
            alpha, _ = ops.query_density(frame, pts.reshape(-1, 3))
        return alpha.view(*sh[1:4])

    def render(self, batch):
        cube = self.query_cube(batch).detach().cpu().numpy()
        cube = np.pad(cube, 10, mode='constant')
        mesh = None
        try:  # if_mesh_renderer.py:98-109 (CPU, third party; out of scope of the CUDA path)
            import mcubes
            import trimesh
            if not (getattr(mcubes, "__file__", None) and getattr(trimesh, "__file__", None)):
                raise ImportError("mcubes / trimesh are placeholders")      # e.g. the test shim's stub modules
            voxel_size = np.array(self.cfg.voxel_size)
            vertices, triangles = mcubes.marching_cubes(cube, self.cfg.mesh_th)
            can_bounds = batch['can_bounds'][0].cpu().numpy()
            LB = (can_bounds[0] - 10 * voxel_size)[None, ...]
            mesh = trimesh.Trimesh(vertices * voxel_size[None, ...] + LB, triangles)
        except ImportError:
            pass
        return {'cube': cube, 'mesh': mesh}
