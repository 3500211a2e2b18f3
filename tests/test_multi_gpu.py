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
            ov.start(g[0])
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
