"""render() / render_chn(): the reference's Render API (model/renderer.py:20-130 and :134-246) over
the B200 rasterizer.  Signatures, argument semantics and returned dict keys are the reference's;
``pc`` is any object with the GaussianModel getters the reference reads (get_xyz, get_opacity,
get_scaling, get_rotation, get_features, get_covariance[_rotation], active_sh_degree,
max_sh_degree) and ``pipe`` any object with convert_shs_python / compute_cov3d_python / debug.

Differences from the reference, all on purpose:
  * tensors are created on ``pc.get_xyz.device`` instead of the literal "cuda";
  * render_chn() honours ``pipe.debug`` instead of hard-coding debug=True
    (model/renderer.py:181), which made every call deep-copy all inputs to the CPU.
"""
import math

import torch

from . import channel_rasterization as chn_rasterize
from .rgbd_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from .sh_utils import eval_sh


def _prepare(viewpoint_camera, pc, pipe, scaling_modifier, override_color, override_shape, foreground, world_rotate):
    xyz = pc.get_xyz
    # zero tensor whose .grad receives the screen-space mean gradients (model/renderer.py:36-41)
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    if override_shape is not None:
        image_height, image_width = override_shape[1], override_shape[0]
    else:
        image_height, image_width = int(viewpoint_camera.image_height), int(viewpoint_camera.image_width)

    means3D = xyz
    opacity = pc.get_opacity
    if foreground is not None:
        opacity[~foreground] = 0  # model/renderer.py:74-75

    scales = rotations = cov3D_precomp = None
    if pipe.compute_cov3d_python:
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales, rotations = pc.get_scaling, pc.get_rotation
    if world_rotate is not None:  # model/renderer.py:88-93
        scales = rotations = None
        world_rotate = torch.from_numpy(world_rotate).float().to(xyz.device)
        means3D = means3D @ world_rotate
        cov3D_precomp = pc.get_covariance_rotation(scaling_modifier, world_rotate)

    shs = colors_precomp = None
    if override_color is None:
        if pipe.convert_shs_python:  # model/renderer.py:100-105
            shs_view = pc.get_features.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
            dir_pp = pc.get_xyz - viewpoint_camera.camera_center.repeat(pc.get_features.shape[0], 1)
            dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
            sh2rgb = eval_sh(pc.active_sh_degree, shs_view, dir_pp_normalized)
            colors_precomp = torch.clamp_min(sh2rgb + 0.5, 0.0)
        else:
            shs = pc.get_features
    else:
        colors_precomp = override_color
    common = dict(image_height=image_height, image_width=image_width, tanfovx=tanfovx, tanfovy=tanfovy,
                  scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
                  projmatrix=viewpoint_camera.full_proj_transform, sh_degree=pc.active_sh_degree,
                  campos=viewpoint_camera.camera_center, prefiltered=False)
    call = dict(means3D=means3D, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp,
                opacities=opacity, scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp)
    return screenspace_points, common, call


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
           override_shape=None, foreground=None, world_rotate=None):
    """RGB + median depth (rgbd rasterizer).  Background tensor (bg_color) must be on the GPU."""
    screenspace_points, common, call = _prepare(viewpoint_camera, pc, pipe, scaling_modifier, override_color,
                                                override_shape, foreground, world_rotate)
    raster_settings = GaussianRasterizationSettings(bg=bg_color, debug=pipe.debug, **common)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    rendered_image, radii, depth = rasterizer(**call)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii, "depth": depth}


def render_chn(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, num_channels=3,
               override_color=None, override_shape=None, foreground=None, world_rotate=None):
    """C-channel feature raster (channel rasterizer)."""
    screenspace_points, common, call = _prepare(viewpoint_camera, pc, pipe, scaling_modifier, override_color,
                                                override_shape, foreground, world_rotate)
    raster_settings = chn_rasterize.GaussianRasterizationSettings(
        bg=bg_color, debug=bool(getattr(pipe, "debug", False)), num_channels=num_channels, **common)
    rasterizer = chn_rasterize.GaussianRasterizer(raster_settings=raster_settings)
    rendered_image, radii = rasterizer.forward(**call)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii}
