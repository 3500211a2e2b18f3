"""render() / render_chn() drop-in behaviour (reference model/renderer.py) on the GPU."""
import math

import numpy as np
import pytest
import torch

from semantic_gaussians_b200.gaussian_model import GaussianModel
from semantic_gaussians_b200.renderer import render, render_chn
from semantic_gaussians_b200.scene_synth import make_scene, orbit_cameras

pytestmark = pytest.mark.gpu


class Pipe:
    convert_shs_python = False
    compute_cov3d_python = False
    debug = False


class Cam:
    def __init__(self, c, dev):
        self.image_width, self.image_height, self.FoVx, self.FoVy = c.image_width, c.image_height, c.FoVx, c.FoVy
        self.world_view_transform = torch.as_tensor(c.world_view_transform, device=dev)
        self.full_proj_transform = torch.as_tensor(c.full_proj_transform, device=dev)
        self.camera_center = torch.as_tensor(c.camera_center, device=dev)


@pytest.fixture(scope="module")
def setup():
    dev = torch.device("cuda:0")
    scene = make_scene(30000, seed=6, sh=True, channels=16)
    pc = GaussianModel.from_activated(scene.xyz, scene.scales, scene.rotations, scene.opacity, scene.shs, device=dev)
    cam = Cam(orbit_cameras(3, 320, 240)[1], dev)
    return dev, scene, pc, cam


def test_render_returns_reference_dict_and_trains(setup):
    dev, scene, pc, cam = setup
    for t in (pc._xyz, pc._features_dc, pc._features_rest, pc._scaling, pc._rotation, pc._opacity):
        t.requires_grad_(True)
    out = render(cam, pc, Pipe, torch.zeros(3, device=dev))
    assert set(out) == {"render", "viewspace_points", "visibility_filter", "radii", "depth"}
    assert out["render"].shape == (3, 240, 320) and out["depth"].shape == (1, 240, 320)
    assert out["radii"].dtype == torch.int32 and out["visibility_filter"].dtype == torch.bool
    loss = (out["render"] - 0.5).abs().mean()
    loss.backward()
    g = out["viewspace_points"].grad                      # densification statistic source (gaussian_model.py:608-612)
    assert g is not None and g.shape == (30000, 3) and float(g[:, :2].norm(dim=-1).max()) > 0
    assert torch.all(g[:, 2] == 0)
    assert torch.all(g[~out["visibility_filter"]] == 0)
    for t in (pc._xyz, pc._features_dc, pc._features_rest, pc._scaling, pc._rotation, pc._opacity):
        assert t.grad is not None and torch.isfinite(t.grad).all() and float(t.grad.abs().max()) > 0
        t.grad = None
        t.requires_grad_(False)


def test_pipeline_flags_agree_with_native_paths(setup):
    dev, scene, pc, cam = setup
    base = render(cam, pc, Pipe, torch.zeros(3, device=dev))["render"]

    class P2(Pipe):
        convert_shs_python = True
    class P3(Pipe):
        compute_cov3d_python = True
    a = render(cam, pc, P2, torch.zeros(3, device=dev))["render"]
    b = render(cam, pc, P3, torch.zeros(3, device=dev))["render"]
    assert float((a - base).abs().max()) < 2e-5            # python SH eval vs in-kernel SH
    assert float((b - base).abs().max()) < 2e-3            # python cov3D (normalised quats, different rounding)


def test_render_chn_override_color_shape_foreground_and_world_rotate(setup):
    dev, scene, pc, cam = setup
    feats = torch.as_tensor(scene.features, device=dev)
    bg = torch.full((16,), 0.3, device=dev)
    out = render_chn(cam, pc, Pipe, bg, num_channels=16, override_color=feats, override_shape=[160, 120])
    assert set(out) == {"render", "viewspace_points", "visibility_filter", "radii"}
    assert out["render"].shape == (16, 120, 160)
    fg = torch.zeros(30000, dtype=torch.bool, device=dev)
    out2 = render_chn(cam, pc, Pipe, bg, num_channels=16, override_color=feats, foreground=fg)
    assert torch.allclose(out2["render"], bg[:, None, None].expand_as(out2["render"]))   # all opacity zeroed
    R = np.eye(3, dtype=np.float32)
    out3 = render_chn(cam, pc, Pipe, bg, num_channels=16, override_color=feats, world_rotate=R)
    ref = render_chn(cam, pc, Pipe, bg, num_channels=16, override_color=feats)
    assert float((out3["render"] - ref["render"]).abs().max()) < 5e-3


def test_render_on_non_default_stream_and_debug_flag(setup):
    dev, scene, pc, cam = setup
    base = render(cam, pc, Pipe, torch.zeros(3, device=dev))["render"]

    class PD(Pipe):
        debug = True
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        o = render(cam, pc, PD, torch.zeros(3, device=dev))["render"]
    s.synchronize()
    assert torch.equal(o, base)
