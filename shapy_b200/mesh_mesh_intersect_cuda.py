"""Drop-in for the reference's only native extension module, `mesh_mesh_intersect_cuda`
(mesh-mesh-intersection/src/mesh_mesh_intersect.cpp:59-64):

    mesh_to_mesh_forward(query_triangles, target_triangles, max_collisions=16, print_timings=False)
        -> [collision_faces (B, Q*M) int64, collision_bcs (B, Q*M, 2, 3)]

Batched over B in one set of launches on the CURRENT stream, no device synchronisation, errors as
RuntimeError (the reference printf()s and calls exit(0), op.cu:76-86).
"""
from .ops import mesh_to_mesh_forward  # noqa: F401
