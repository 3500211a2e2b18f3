"""N > 1 on real GPUs: view-sharded forward/backward + NCCL all-reduce equals the single-GPU sum.
Needs >= 2 CUDA devices (skipped otherwise; the round-end scaling run exercises bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from semantic_gaussians_b200 import distributed as D
        g = _grads(dev, list(D.shard_range(4, rank, world)))
        D.allreduce_sums(g)
        q.put((rank, [t.cpu().numpy() for t in g]))
    finally:
        dist.destroy_process_group()


def _worker_overlap(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    from semantic_gaussians_b200.distributed import nccl_overlap_options
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, pg_options=nccl_overlap_options())
    try:
        from semantic_gaussians_b200 import distributed as D
        ov = D.OverlappedFeatureGradReduce(dev)
        outs = []
        for rep in range(3):                       # the event is re-recorded by every backward
            g = _grads(dev, [rank])                # one view per rank, then the overlapped exchange
            # rep 0/1: the gradient buffer was adopted as .grad (fresh) -> early exchange on the native event;
            # rep 2: conservative ordering after the current stream (what a pre-existing .grad requires)
            ov.start(g[0], fresh=rep < 2)
            D.allreduce_sums(g[1:])
            ov.finish()
            outs.append([t.cpu().numpy() for t in g])
        ov.close()
        q.put((rank, outs))
    finally:
        dist.destroy_process_group()


def _grads(dev, view_ids):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import dev_cam, dev_scene, run_ours
    from semantic_gaussians_b200.scene_synth import make_scene, orbit_cameras
    scene = make_scene(20000, seed=3, channels=24)
    sc = dev_scene(scene, dev, requires_grad=True)
    cams = orbit_cameras(4, 256, 192)
    rng = np.random.default_rng(0)
    dLs = [torch.as_tensor(rng.standard_normal((24, 192, 256)).astype(np.float32), device=dev) for _ in range(4)]
    for i in view_ids:
        o = run_ours("chn", sc, dev_cam(cams[i], dev), torch.zeros(24, device=dev), use_features=True)
        o["color"].backward(dLs[i])
    out = []
    for k in ("features", "means3D", "scales", "rotations", "opacities"):
        out.append(sc[k].grad if sc[k].grad is not None else torch.zeros_like(sc[k]))
    return out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_view_sharded_gradients_match_single_gpu():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = [t.cpu().numpy() for t in _grads(torch.device("cuda:0"), [0, 1, 2, 3])]
    for a, b, s in zip(res[0][1], res[1][1], single):
        assert np.array_equal(a, b)
        scale = np.abs(s).max()
        assert np.abs(a - s).max() <= 1e-4 * scale + 1e-7


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_overlapped_feature_grad_allreduce_matches_single_gpu():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_overlap, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = [t.cpu().numpy() for t in _grads(torch.device("cuda:0"), [0, 1])]
    for rep in range(3):
        for a, b, s in zip(res[0][1][rep], res[1][1][rep], single):
            assert np.array_equal(a, b)
            scale = np.abs(s).max()
            assert np.abs(a - s).max() <= 1e-4 * scale + 1e-7


# ------------------------------------------------------------------ batched views (K4 path) and sharded fusion (K5 path)
def _batch_setup(dev):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from semantic_gaussians_b200.gaussian_model import GaussianModel
    from semantic_gaussians_b200.scene_synth import make_scene, orbit_cameras
    C, W, H = 40, 256, 192
    scene = make_scene(20000, seed=3, channels=C)
    pc = GaussianModel.from_activated(scene.xyz, scene.scales, scene.rotations, scene.opacity, device=dev)
    pc.active_sh_degree = 0
    feats = torch.as_tensor(scene.features, device=dev).contiguous().requires_grad_(True)
    leaves = [feats, pc._xyz, pc._scaling, pc._rotation, pc._opacity]
    for t in leaves[1:]:
        t.requires_grad_(True)

    class Cam:
        pass
    cams = []
    for c in orbit_cameras(4, W, H):
        v = Cam()
        v.image_width, v.image_height, v.FoVx, v.FoVy = c.image_width, c.image_height, c.FoVx, c.FoVy
        v.world_view_transform = torch.as_tensor(c.world_view_transform, device=dev)
        v.full_proj_transform = torch.as_tensor(c.full_proj_transform, device=dev)
        v.camera_center = torch.as_tensor(c.camera_center, device=dev)
        cams.append(v)
    rng = np.random.default_rng(0)
    dLs = [torch.as_tensor(rng.standard_normal((C, H, W)).astype(np.float32), device=dev) for _ in range(4)]
    return pc, feats, leaves, cams, dLs, C


class _Pipe:
    convert_shs_python = False
    compute_cov3d_python = False
    debug = False


def _batch_grads(dev, view_ids, overlap=None):
    from semantic_gaussians_b200.renderer import render_chn_batch
    pc, feats, leaves, cams, dLs, C = _batch_setup(dev)
    outs = render_chn_batch([cams[i] for i in view_ids], pc, _Pipe, torch.zeros(C, device=dev), num_channels=C,
                            override_color=feats)
    if overlap is not None:
        overlap.arm(feats)
    torch.autograd.backward([o["render"] for o in outs], [dLs[i] for i in view_ids])
    return [t.grad for t in leaves]


def _worker_batch(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    from semantic_gaussians_b200.distributed import nccl_overlap_options
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, pg_options=nccl_overlap_options())
    try:
        from semantic_gaussians_b200 import distributed as D
        ov = D.OverlappedFeatureGradReduce(dev)
        g = _batch_grads(dev, list(D.shard_range(4, rank, world)), overlap=ov)
        ov.start(g[0])                      # armed before the backward: the batch's (P, C) buffer IS .grad -> early exchange
        D.allreduce_sums(g[1:])
        ov.finish()
        ov.close()
        q.put((rank, [t.cpu().numpy() for t in g]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_batched_view_shards_with_overlapped_exchange_match_single_gpu():
    """K4 path: every rank renders its contiguous shard of the view batch through render_chn_batch (feature gradient
    summed in place over the local views), the (P, C) gradient is exchanged on the native event while the chain
    kernels of the batch still run; result == one GPU rendering all views."""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_batch, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = [t.cpu().numpy() for t in _batch_grads(torch.device("cuda:0"), [0, 1, 2, 3])]
    for a, b, s in zip(res[0][1], res[1][1], single):
        assert np.array_equal(a, b)
        # both sides sum per-tile partial gradients with red.global in scheduling order (two runs of the same view
        # already differ by ~1e-5 of the scale); the cross-rank sum adds fp32 re-association
        assert np.abs(a - s).max() <= 1e-4 * np.abs(s).max() + 1e-9


def _fusion_partial(dev, view_ids):
    from semantic_gaussians_b200.fusion import PointCloudToImageMapper
    from semantic_gaussians_b200.scene_synth import make_scene, room_cameras
    P, C, w, h, nv = 50000, 64, 160, 120, 6
    scene = make_scene(P, 1, kind="room")
    cams = room_cameras(nv, w, h)
    rng = np.random.default_rng(5)
    maps = [rng.standard_normal((C, h, w)).astype(np.float16) for _ in range(nv)]
    xyz = torch.as_tensor(scene.xyz, device=dev)
    fs = torch.zeros((P, C), device=dev)
    cnt = torch.zeros(P, device=dev)

    def acc(i):
        m = PointCloudToImageMapper([w, h], 0.5, 5, cams[i].intrinsics(), device=dev)
        m.accumulate(cams[i].world_view_transform, xyz, torch.from_numpy(maps[i]).to(dev), fs, cnt,
                     torch.full((h, w), 2.5, device=dev))
    return nv, acc, fs, cnt


def _worker_fusion(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from semantic_gaussians_b200 import distributed as D
        from semantic_gaussians_b200.fusion import normalize_fused
        nv, acc, fs, cnt = _fusion_partial(dev, None)
        D.fuse_views_sharded(nv, acc, fs, cnt, normalize_fused)
        q.put((rank, fs.cpu().numpy(), cnt.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_fuse_views_sharded_under_nccl_matches_single_gpu():
    """K5 path: views strided over the ranks, one NCCL all-reduce of the (P, C) sums and the counts, then normalise:
    counts exact, means within fp32 re-association of the cross-rank sum."""
    import torch.multiprocessing as mp
    from semantic_gaussians_b200.fusion import normalize_fused
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_fusion, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nv, acc, fs, cnt = _fusion_partial(torch.device("cuda:0"), None)
    for i in range(nv):
        acc(i)
    seen = (cnt > 0).cpu().numpy()
    normalize_fused(fs, cnt)
    want, want_cnt = fs.cpu().numpy(), cnt.cpu().numpy()
    assert seen.sum() > 1000
    for _, got, got_cnt in res:
        assert np.array_equal(got_cnt, want_cnt)
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max() + 1e-9
    assert np.array_equal(res[0][1], res[1][1])
