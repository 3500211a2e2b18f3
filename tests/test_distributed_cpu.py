"""World-size-2 gloo tests of the N>1 path: view sharding + the single sum all-reduce
(semantic-gaussians_b200/distributed.py).  The per-view work on each rank is done by the CPU fusion
oracle here (test infrastructure); on the GPU box tests/test_multi_gpu.py runs the real kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fuse_single(scene, cams, feats, P, C, w, h):
    from oracle import fusion_oracle as fo
    fs, cnt = np.zeros((P, C), np.float32), np.zeros(P, np.float32)
    for i, cam in enumerate(cams):
        K = fo.rescale_intrinsics(cam.intrinsics(), [w, h])
        m = fo.compute_mapping(cam.world_view_transform, scene.xyz, [w, h], K, 0.05, 2, None)
        fo.accumulate(feats[i], m, fs, cnt)
    return fs, cnt


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from make_fusion_golden import fusion_inputs
        from oracle import fusion_oracle as fo
        from semantic_gaussians_b200 import distributed as D
        scene, cams, feats, _ = fusion_inputs(seed=3, P=3000, w=64, h=48, C=8, nviews=5)
        P, C = scene.P, 8
        fs = torch.zeros((P, C), dtype=torch.float32)
        cnt = torch.zeros(P, dtype=torch.float32)

        def acc(i):
            K = fo.rescale_intrinsics(cams[i].intrinsics(), [64, 48])
            m = fo.compute_mapping(cams[i].world_view_transform, scene.xyz, [64, 48], K, 0.05, 2, None)
            fo.accumulate(feats[i], m, fs.numpy(), cnt.numpy())

        def norm(a, b):
            fo.normalize(a.numpy(), b.numpy())
        D.fuse_views_sharded(len(cams), acc, fs, cnt, norm)

        # gradient-style reduction: every rank contributes its shard of "views"
        g = torch.zeros((1000, 4))
        small = torch.zeros(7)
        for i in D.shard_range(9, rank, world):
            g += float(i + 1)
            small += 1.0
        D.allreduce_sums([g, small], bucket_bytes=4096)
        q.put((rank, fs.numpy().copy(), cnt.numpy().copy(), float(g[0, 0]), float(small[0])))
    finally:
        dist.destroy_process_group()


def test_shard_helpers():
    from semantic_gaussians_b200 import distributed as D
    for n, w in ((32, 8), (9, 2), (5, 8), (0, 3)):
        got = sorted(i for r in range(w) for i in D.shard_range(n, r, w))
        assert got == list(range(n))
        got = sorted(i for r in range(w) for i in D.shard_strided(n, r, w))
        assert got == list(range(n))
    assert list(D.shard_range(32, 3, 8)) == [12, 13, 14, 15]       # K4: 4 views per GPU
    with pytest.raises(ValueError):
        D.shard_range(4, 2, 2)


def test_allreduce_noop_without_process_group():
    from semantic_gaussians_b200 import distributed as D
    t = torch.ones(10)
    D.allreduce_sums([t])
    assert torch.all(t == 1)


@pytest.mark.timeout(180)
def test_world2_fusion_and_grad_reduction_match_single_process():
    from make_fusion_golden import fusion_inputs
    from oracle import fusion_oracle as fo
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=150) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    scene, cams, feats, _ = fusion_inputs(seed=3, P=3000, w=64, h=48, C=8, nviews=5)
    fs, cnt = _fuse_single(scene, cams, feats, scene.P, 8, 64, 48)
    fo.normalize(fs, cnt)
    for rank, fs_r, cnt_r, g00, s0 in res:
        assert np.array_equal(cnt_r, cnt)                               # counts are exact integers
        np.testing.assert_allclose(fs_r, fs, rtol=1e-5, atol=1e-6)      # fp32 re-association only
        assert g00 == sum(range(1, 10)) and s0 == 9.0
    assert np.array_equal(res[0][1], res[1][1])                         # ranks agree bit for bit
