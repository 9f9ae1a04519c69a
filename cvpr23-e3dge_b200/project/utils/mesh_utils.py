from e3dge_b200.mesh_utils import align_volume, extract_mesh_with_marching_cubes  # noqa: F401
