"""Surface path of the renderer (partial): what `render(..., return_mesh=True)` does after the
fused kernel has produced the per-sample SDF (project/utils/volume_renderer.py:1703-1727) —

    frostum_aligned_sdf = align_volume(sdf)            project/utils/mesh_utils.py:17-44   (3-D grid_sample)
    verts, faces = marching_cubes(sdf_vol, 0)          :48-69, volume_renderer.py:1733-1758 (skimage, on the host)

`align_volume` is device-side PyTorch here as it is there (one `grid_sample` over a [B,1,S,H,W] volume).  The
reference extracts the surface with scikit-image's marching cubes on the CPU and wraps it in a trimesh object;
this module does the same WHEN those two packages are importable and raises an ImportError naming them when not
(neither is installed in the build container, so the extraction itself is untested here; a GPU extractor is not
built).  The depth-mesh Delaunay / pytorch3d rasteriser of trainer.py:2252-2346 is not provided."""
from collections import namedtuple

import torch
import torch.nn.functional as F

Mesh = namedtuple("Mesh", ["vertices", "faces"])  # stand-in when trimesh is not installed


def align_volume(volume, near=0.88, far=1.12):
    """Resamples a volume given on the camera frustum's sample lattice [B,H,W,D,C] onto the regular grid of its
    far plane: slice k is magnified by linspace(far/near, 1, D)[k] in x and y; cells that fall outside the
    frustum are set to 1 (positive = outside the surface) — mesh_utils.py:17-44."""
    b, h, w, d, c = volume.shape
    dev, dt = volume.device, volume.dtype
    yy, xx, zz = torch.meshgrid(torch.linspace(-1, 1, h, device=dev, dtype=dt), torch.linspace(-1, 1, w, device=dev, dtype=dt),
                                torch.linspace(-1, 1, d, device=dev, dtype=dt), indexing="ij")
    coeffs = torch.linspace(far / near, 1, d, device=dev, dtype=dt).reshape(1, 1, d)
    grid = torch.stack([xx * coeffs, yy * coeffs, zz], -1).unsqueeze(0)            # [1,H,W,D,3]
    outside = ((grid < -1) | (grid > 1)).any(-1, keepdim=True)                     # [1,H,W,D,1]
    sampled = F.grid_sample(volume.permute(0, 4, 3, 1, 2).contiguous(),            # [B,C,D,H,W]
                            grid.permute(0, 3, 1, 2, 4).expand(b, d, h, w, 3).contiguous(),
                            padding_mode="border", align_corners=True)
    sampled = sampled.permute(0, 3, 4, 2, 1).contiguous()                          # [B,H,W,D,C]
    return torch.where(outside, torch.ones((), device=dev, dtype=dt), sampled)


def extract_mesh_with_marching_cubes(sdf, shading=False):
    """Zero level set of sdf [1,H,W,D,1] as (mesh, verts, faces) in scene units — mesh_utils.py:48-69,
    volume_renderer.py:1733-1758.  Needs scikit-image (and uses trimesh when present), like the reference."""
    try:
        from skimage.measure import marching_cubes
    except ImportError as exc:  # the reference imports it at module level and fails the same way
        raise ImportError("mesh extraction uses skimage.measure.marching_cubes on the host, as the reference does "
                          "(project/utils/mesh_utils.py:10); scikit-image is not installed") from exc
    _, h, w, d, _ = sdf.shape
    sdf_vol = sdf[0, ..., 0].permute(1, 0, 2).detach().cpu().numpy()  # (y, x, z) -> (x, y, z)
    verts, faces, _, _ = marching_cubes(sdf_vol, 0)
    verts[:, 0] = (verts[:, 0] / float(w) - 0.5) * 0.24  # back to the scene's [-0.12, 0.12] box
    verts[:, 1] = (verts[:, 1] / float(h) - 0.5) * 0.24
    verts[:, 2] = (verts[:, 2] / float(d) - 0.5) * 0.24
    verts[:, 2] *= -1  # normal direction
    verts[:, 1] *= -1
    try:
        import trimesh
        mesh = trimesh.Trimesh(verts, faces)
    except ImportError:
        mesh = Mesh(verts, faces)
    return mesh, verts, faces
