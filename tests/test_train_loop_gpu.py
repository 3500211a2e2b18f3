"""GPU integration: the pieces either side of the rasterizer work together the way train.py uses them
(train.py:100-180): create_from_pcd (distCUDA2) -> training_setup -> render() -> L1 loss -> backward ->
add_densification_stats (viewspace_points.grad, radii) -> Adam step -> densify_and_prune / reset_opacity."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


def test_short_rgb_fit_with_density_control():
    from semantic_gaussians_b200.gaussian_model import GaussianModel
    from semantic_gaussians_b200.renderer import render
    from semantic_gaussians_b200.scene_synth import make_scene, orbit_cameras
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    scene = make_scene(4000, seed=21, sh=True, scale_mean=0.05)
    gt = GaussianModel.from_activated(scene.xyz, scene.scales, scene.rotations, scene.opacity, shs=scene.shs, device=dev)

    class Pipe:
        convert_shs_python = False
        compute_cov3d_python = False
        debug = False

    views = []
    for c in orbit_cameras(6, 160, 120):
        v = SimpleNamespace(image_width=c.image_width, image_height=c.image_height, FoVx=c.FoVx, FoVy=c.FoVy,
                            world_view_transform=torch.as_tensor(c.world_view_transform, device=dev),
                            full_proj_transform=torch.as_tensor(c.full_proj_transform, device=dev),
                            camera_center=torch.as_tensor(c.camera_center, device=dev))
        views.append(v)
    bg = torch.zeros(3, device=dev)
    with torch.no_grad():
        targets = [render(v, gt, Pipe, bg)["render"].clone() for v in views]

    rng = np.random.default_rng(0)
    keep = rng.choice(4000, 1500, replace=False)
    pts = scene.xyz[keep] + rng.normal(0, 0.01, (1500, 3)).astype(np.float32)
    m = GaussianModel(3).create_from_pcd(pts, rng.uniform(0.3, 0.7, (1500, 3)), spatial_lr_scale=1.0, device=dev)
    m.active_sh_degree = 0
    args = SimpleNamespace(percent_dense=0.01, position_lr_init=1.6e-4, position_lr_final=1.6e-6, position_lr_delay_mult=0.01,
                           position_lr_max_steps=300, feature_lr=2.5e-3, opacity_lr=0.05, scaling_lr=5e-3, rotation_lr=1e-3)
    m.training_setup(args)
    P0 = m._xyz.shape[0]
    losses, counts = [], []
    for it in range(1, 241):
        m.update_learning_rate(it)
        v = views[it % len(views)]
        out = render(v, m, Pipe, bg)
        loss = (out["render"] - targets[it % len(views)]).abs().mean()
        loss.backward()
        losses.append(float(loss))
        with torch.no_grad():
            vis, radii = out["visibility_filter"], out["radii"]
            m.max_radii2D[vis] = torch.max(m.max_radii2D[vis], radii[vis].float())
            m.add_densification_stats(out["viewspace_points"], vis)
            if it % 60 == 0:
                counts.append(m.densify_and_prune(0.0002, 0.005, 3.0, None))
            m.optimizer.step()
            m.optimizer.zero_grad(set_to_none=True)
    first, last = float(np.mean(losses[:10])), float(np.mean(losses[-10:]))
    assert np.isfinite(losses).all()
    assert last < 0.8 * first, (first, last)
    assert sum(c["cloned"] + c["split"] for c in counts) > 0          # the screen-space gradients drove a densification
    assert m._xyz.shape[0] != P0 and m._xyz.shape[0] == m.max_radii2D.shape[0] == m.denom.shape[0]
    for g in m.optimizer.param_groups:
        assert g["params"][0].shape[0] == m._xyz.shape[0]
    # opacity reset (train.py:175) and one more step through the swapped parameter
    m.reset_opacity()
    assert float(m.get_opacity.max()) <= 0.0100001
    out = render(views[0], m, Pipe, bg)
    (out["render"] - targets[0]).abs().mean().backward()
    m.optimizer.step()
    assert torch.isfinite(m._opacity).all() and m._opacity.grad is not None
