"""psi-release_b200: B200-native geometry hot path of PSI scene fitting.

SMPL-X LBS forward/backward, scene-SDF trilinear lookup and the Chamfer / nearest
neighbour contact term as hand-written sm_100a CUDA kernels behind a C ABI
(include/psi_b200.h, lib/libpsi_b200.so), exposed through the reference's own Python call
signatures (smplx.create, chamfer_pytorch.dist_chamfer.chamferDist, F.grid_sample,
GeometryTransformer.verts_transform, FittingOP).  Import this package as
`psi_release_b200`.
"""
__version__ = "0.1.0"
