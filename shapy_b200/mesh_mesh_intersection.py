"""MeshMeshIntersection module, mirror of
mesh-mesh-intersection/mesh_mesh_intersection/mesh_mesh_intersection.py:32-62."""
import torch
import torch.nn as nn

from . import mesh_mesh_intersect_cuda


class MeshMeshIntersection(nn.Module):
    def __init__(self, max_collisions=32):
        super().__init__()
        self.max_collisions = max_collisions

    @torch.no_grad()
    def forward(self, query_triangles, target_triangles, print_timings=False):
        faces, bcs = mesh_mesh_intersect_cuda.mesh_to_mesh_forward(
            query_triangles, target_triangles, print_timings=print_timings, max_collisions=self.max_collisions)
        return faces, bcs
