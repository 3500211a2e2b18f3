#!/usr/bin/env python
"""bench.py — headline benchmark of the rasterizer hot path (see BASELINE.json / DESIGN.md §Measurement).

Workload K3: 1 M Gaussians, 256-channel semantic features, 1920x1080, forward + backward of one
view per step (synthetic scene).  One JSON line on stdout (rank 0).

  value          Mviews/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e            same metric through the public API render_chn(): per step the camera matrices and
                 a per-pixel label map come from pinned host memory, the loss scalar goes back
  roofline       dominant kernel: algorithmic bytes / CUDA-event duration vs measured HBM peak
  cpu_baseline   the CPU oracle (a port: the reference has no CPU rasterizer) on a bounded sample
  --impl reference   times that CPU port alone (rank 0 only), same JSON contract
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

P_GAUSS, CHANNELS, WIDTH, HEIGHT, NVIEWS = 1_000_000, 256, 1920, 1080, 8
NUM_CLASSES = 20
METRIC = "Mviews/s + HBM GB/s, 1M Gaussians, 256-ch features, 1080p, fwd+bwd"
WORKLOAD = "K3: 1M Gaussians x 256-ch features, 1920x1080, 1 view/step, fwd+bwd (configs[2])"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


# ------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ------------------------------------------------------------------------------ CPU arm
def cpu_sample(scene, cam, band_tile_rows=2, channels=CHANNELS):
    """One bounded sample of the K3 workload on the host cores with the CPU oracle (a port of the
    reference's algorithm): full per-Gaussian preprocess, then binning + blend forward + backward
    of a band of tile rows at the image centre; returns seconds per stage and the scale factor."""
    from oracle import oracle as orc
    W, H = cam.image_width, cam.image_height
    gy = (H + 15) // 16
    r0 = max(0, gy // 2 - band_tile_rows // 2)
    r1 = min(gy, r0 + band_tile_rows)
    rows = (r0 * 16, min(H, r1 * 16))
    cd = orc.cam_dict(cam)
    bg = np.zeros(channels, np.float32)
    t0 = time.perf_counter()
    pre = orc.preprocess(scene.xyz, scene.scales, scene.rotations, scene.opacity, cd["viewmatrix"],
                         cd["projmatrix"], cd["campos"], W, H, cd["tanfovx"], cd["tanfovy"],
                         colors_precomp=scene.features)
    t1 = time.perf_counter()
    b = orc.bin_instances(pre, W, H, tile_rows=(r0, r1))
    f = orc.render_forward(pre, b, scene.features, bg, W, H, rows=rows)
    fwd = dict(pre=pre, bin=b, colors=scene.features, **f)
    dL = np.full((channels, H, W), 1.0 / (H * W), np.float32)
    orc.backward(fwd, orc.scene_dict(scene), cd, W, H, bg, dL, features=scene.features, rows=rows)
    t2 = time.perf_counter()
    frac = (rows[1] - rows[0]) / H
    full = (t1 - t0) + (t2 - t1) / frac
    return dict(seconds_sample=t2 - t0, seconds_full_view_est=full, rows=rows, frac=frac,
                threads=orc.num_threads())


def run_cpu_reference(args, rank, world):
    if rank != 0:
        return
    from semantic_gaussians_b200.scene_synth import make_scene, orbit_cameras
    scene = make_scene(P_GAUSS, seed=0, channels=CHANNELS)
    cams = orbit_cameras(NVIEWS, WIDTH, HEIGHT)
    for i in range(args.warmup):
        cpu_sample(scene, cams[i % NVIEWS])
    t0 = time.perf_counter()
    est = 0.0
    last = None
    for i in range(args.steps):
        last = cpu_sample(scene, cams[i % NVIEWS])
        est += last["seconds_full_view_est"]
    wall = time.perf_counter() - t0
    views_per_s = args.steps / est
    value = views_per_s * 1e-6
    sample = (f"per step: full preprocess of 1M Gaussians + binning/blend fwd+bwd of image rows "
              f"[{last['rows'][0]},{last['rows'][1]}) ({last['frac']:.1%} of pixels), scaled to the full view")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mviews/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * est / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU port of the reference algorithm (oracle/raster_oracle.c); "
                   "the reference itself has no CPU rasterizer", "wall_s": wall},
        "cpu_baseline": {"value": value, "unit": "Mviews/s", "cores": last["threads"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "Mviews/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ GPU arm
def algorithmic_bytes(P, P_vis, R, C, W, H):
    """SURVEY.md §8(d) / BASELINE.md §4 compulsory traffic, split by kernel."""
    px = W * H
    # per kernel: every input read once, every output written once
    alpha = 4 * R + 32 * P_vis + 8 * px                       # ids + splat records in, final_T / n_contrib out
    fwd_blend = 4 * C * P_vis + 4 * C * px + 4 * px           # features in, image out (+ final_T in)
    chain = 4 * C * px + 4 * C * P_vis + 32 * P_vis + 4 * px + 28 * P_vis   # dL/dout + features in, 7 geometry grads out
    dfeat = 4 * C * px + 4 * C * P_vis                        # dL/dout in, dL/dfeature out
    fwd_total = 44 * P + 4 * C * P_vis + 4 * C * px + 8 * px + 24 * R
    bwd_total = 4 * C * px + 8 * C * P_vis + 4 * R + 8 * px + 80 * P + 44 * P
    return dict(alpha_pass=alpha, blend_fwd=fwd_blend, blend_bwd=chain, dfeature=dfeat, fwd=fwd_total, bwd=bwd_total)


FMA_PEAK_TFLOPS = 70.5   # dependent-free FFMA loop on this pool's B200s (tools/microbench.cu)


def fma_roofline(C, blended_pairs, per_stage):
    """fp32 CUDA-core roof of the three C-wide contractions: achieved = algorithmic flops
    (2*C per blended pair and contraction, zero-weight padding not counted) / stage time."""
    out = {"peak_tflops": FMA_PEAK_TFLOPS, "peak_source": "measured FFMA micro-benchmark (tools/microbench.cu)",
           "algorithmic_flops_per_contraction": 2 * C * blended_pairs, "kernels": {}}
    for k in ("blend_fwd", "blend_bwd", "dfeature"):
        ms = per_stage.get(k)
        if ms:
            tf = 2 * C * blended_pairs / (ms * 1e-3) * 1e-12
            out["kernels"][k] = {"ms": ms, "achieved_tflops": tf, "frac": tf / FMA_PEAK_TFLOPS}
    tot = sum(per_stage.get(k, 0.0) for k in ("blend_fwd", "blend_bwd", "dfeature", "alpha_pass"))
    if tot:
        out["achieved_tflops"] = 6 * C * blended_pairs / (tot * 1e-3) * 1e-12
        out["frac"] = out["achieved_tflops"] / FMA_PEAK_TFLOPS
    return out


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from semantic_gaussians_b200 import _lib
    from semantic_gaussians_b200.gaussian_model import GaussianModel
    from semantic_gaussians_b200.renderer import render_chn
    from semantic_gaussians_b200.scene_synth import make_scene, orbit_cameras
    from semantic_gaussians_b200.semantic import distill_loss_and_grad

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the rasterizer has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        from semantic_gaussians_b200.distributed import nccl_overlap_options
        dist.init_process_group("nccl", device_id=dev, pg_options=nccl_overlap_options())
    _lib.load()

    scene = make_scene(P_GAUSS, seed=0, channels=CHANNELS)
    cams_np = orbit_cameras(NVIEWS, WIDTH, HEIGHT)
    pc = GaussianModel.from_activated(scene.xyz, scene.scales, scene.rotations, scene.opacity, device=dev)
    pc.active_sh_degree = 0
    feats = torch.as_tensor(scene.features, device=dev).contiguous().requires_grad_(True)
    for t in (pc._xyz, pc._scaling, pc._rotation, pc._opacity):
        t.requires_grad_(True)
    params = [feats, pc._xyz, pc._scaling, pc._rotation, pc._opacity]
    bg = torch.zeros(CHANNELS, device=dev)

    class Pipe:
        convert_shs_python = False
        compute_cov3d_python = False
        debug = False

    class Cam:
        pass

    def dev_cam(c):
        v = Cam()
        v.image_width, v.image_height, v.FoVx, v.FoVy = c.image_width, c.image_height, c.FoVx, c.FoVy
        v.world_view_transform = torch.as_tensor(c.world_view_transform, device=dev)
        v.full_proj_transform = torch.as_tensor(c.full_proj_transform, device=dev)
        v.camera_center = torch.as_tensor(c.camera_center, device=dev)
        return v

    cams = [dev_cam(c) for c in cams_np]
    # host side of the e2e arm: pinned camera blocks (35 floats) and per-view label maps
    host_cam = [torch.from_numpy(np.concatenate([c.world_view_transform.ravel(), c.full_proj_transform.ravel(),
                                                 c.camera_center.ravel()]).astype(np.float32)).pin_memory()
                for c in cams_np]
    rng = np.random.default_rng(1234 + rank)
    host_labels = [torch.from_numpy(rng.integers(0, NUM_CLASSES, size=(HEIGHT, WIDTH), dtype=np.int64)
                                    .astype(np.int32)).pin_memory() for _ in range(2)]
    class_emb = torch.nn.functional.normalize(torch.randn(NUM_CLASSES, CHANNELS, device=dev), dim=1)
    class_emb_t = (-class_emb.t() / (CHANNELS * HEIGHT * WIDTH)).contiguous()   # (C, K), pre-scaled
    dL_fixed = torch.randn((CHANNELS, HEIGHT, WIDTH), device=dev) / (HEIGHT * WIDTH)
    flat_small = None

    def zero_grads():
        for p in params:
            p.grad = None

    overlap = None
    if world > 1:
        from semantic_gaussians_b200.distributed import OverlappedFeatureGradReduce
        overlap = OverlappedFeatureGradReduce(dev)        # dL/dfeature is final before the chain kernels run

    def allreduce_grads():
        if world == 1:
            return
        overlap.start(feats.grad)                         # (P, C) fp32, 1 GB: reduced under the chain/geometry kernels
        small = torch.cat([p.grad.reshape(-1) for p in params[1:]])
        dist.all_reduce(small)
        overlap.finish()

    def step_device(i):
        cam = cams[(i * world + rank) % NVIEWS]           # views shard across ranks
        out = render_chn(cam, pc, Pipe, bg, num_channels=CHANNELS, override_color=feats)
        out["render"].backward(dL_fixed)
        allreduce_grads()
        zero_grads()

    cam_dev = Cam()
    cam_dev.image_width, cam_dev.image_height = WIDTH, HEIGHT
    cam_dev.FoVx, cam_dev.FoVy = cams_np[0].FoVx, cams_np[0].FoVy
    cam_buf = torch.empty(35, device=dev)
    label_buf = torch.empty((HEIGHT, WIDTH), dtype=torch.int32, device=dev)

    # the loss is read back the way training loops do it: an asynchronous 4-byte copy into pinned memory each
    # step, consumed one step later (and the last one inside the timed region), so the host keeps launching
    loss_host = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    e2e_losses = []

    def step_e2e(i):
        k = (i * world + rank) % NVIEWS
        cam_buf.copy_(host_cam[k], non_blocking=True)                      # H2D 140 B
        label_buf.copy_(host_labels[i % 2], non_blocking=True)             # H2D 8.3 MB
        cam_dev.world_view_transform = cam_buf[0:16].view(4, 4)
        cam_dev.full_proj_transform = cam_buf[16:32].view(4, 4)
        cam_dev.camera_center = cam_buf[32:35]
        out = render_chn(cam_dev, pc, Pipe, bg, num_channels=CHANNELS, override_color=feats)
        # open-vocabulary distillation loss  L = -mean <render[:, p], E[label(p)]>  and its gradient, one fused pass
        loss, dL = distill_loss_and_grad(out["render"], class_emb, label_buf)
        loss_host[i % 2].copy_(loss.reshape(1), non_blocking=True)            # D2H 8 B (float64 scalar)
        loss_ev[i % 2].record()
        out["render"].backward(dL)
        allreduce_grads()
        zero_grads()
        if i > 0:
            finish_e2e(i - 1)

    def finish_e2e(i):
        loss_ev[i % 2].synchronize()
        e2e_losses.append(float(loss_host[i % 2]))

    def timed(fn, steps, finish=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish(steps - 1)      # the last step's result is read inside the timed region
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    stream = torch.cuda.current_stream(dev).cuda_stream
    ctx = _lib.ctx_for(local_rank, stream)

    # ---- warm-up, then the device-resident timed region (with stage tracing + clock sampling)
    for i in range(max(args.warmup, 3)):
        step_device(i)
    torch.cuda.synchronize(dev)
    # per-view statistics of the scene (reported with every timing)
    with torch.no_grad():
        from semantic_gaussians_b200.rasterizer import _C_chn
        e = torch.Tensor([])
        c0 = cams[rank % NVIEWS]
        Rn, _, radii, _, _, img = _C_chn.rasterize_gaussians(
            bg, pc.get_xyz, feats.detach(), pc.get_opacity, pc.get_scaling, pc.get_rotation, 1.0, e,
            c0.world_view_transform, c0.full_proj_transform, math.tan(c0.FoVx / 2), math.tan(c0.FoVy / 2), HEIGHT, WIDTH,
            e, 0, c0.camera_center, False, False, CHANNELS)
        P_vis = int((radii > 0).sum())
        nc = torch.zeros(HEIGHT * WIDTH, dtype=torch.int32, device=dev)
        _lib.load().sgb_state_field(b"n_contrib", P_GAUSS, Rn, WIDTH, HEIGHT, None, None, img.data_ptr(),
                                    nc.data_ptr(), stream)
        ncontrib_mean = float(nc.float().mean())
        blended_pairs = _lib.view_stat(ctx, 0)        # (pixel, Gaussian) pairs blended in this view
        pool_chunks = _lib.view_stat(ctx, 1)
        del img, nc
    launches0 = _lib.launch_count(ctx)
    _lib.profile_enable(ctx, True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    stages = _lib.profile_read(ctx)
    _lib.profile_enable(ctx, False)
    launches1 = _lib.launch_count(ctx)

    # ---- end-to-end arm through the public API with host buffers
    for i in range(2):
        step_e2e(i)
    finish_e2e(1)
    e2e_losses.clear()
    ms_e2e = timed(step_e2e, args.steps, finish=finish_e2e)
    assert len(e2e_losses) == args.steps and all(math.isfinite(v) for v in e2e_losses)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    views = args.steps * world
    value = views / (ms_dev * 1e-3) * 1e-6
    e2e_value = views / (ms_e2e * 1e-3) * 1e-6
    peak, peak_src = measured_peaks()
    ab = algorithmic_bytes(P_GAUSS, P_vis, Rn, CHANNELS, WIDTH, HEIGHT)
    per_stage = {k: (v[0] / max(v[1], 1)) for k, v in stages.items() if v[1] > 0}
    dom = max(("blend_fwd", "blend_bwd", "dfeature", "alpha_pass"), key=lambda k: per_stage.get(k, 0.0))
    dom_ms = per_stage.get(dom, float("nan"))
    achieved = ab[dom] / (dom_ms * 1e-3) * 1e-9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")   # filled from an ncu --set full capture
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except Exception:
            traffic = None
    kernel_ms = sum(per_stage.values())
    eff_gbs = (ab["fwd"] + ab["bwd"]) / (ms_dev / args.steps * 1e-3) * 1e-9

    line = {
        "metric": METRIC, "value": value, "unit": "Mviews/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "P": P_GAUSS, "C": CHANNELS, "W": WIDTH, "H": HEIGHT,
                   "views_per_step": world, "parallelism": f"view-sharded x{world}" + (" + NCCL all-reduce of per-Gaussian grads (feature grad overlapped with the chain backward)" if world > 1 else ""),
                   "l2": "inputs larger than L2 (1.0 GB feature table, 2.1 GB dL/dout, 2.1 GB output per step; 8 cycling views)",
                   "P_vis": P_vis, "R": int(Rn), "gaussians_per_tile_mean": Rn / (((WIDTH + 15) // 16) * ((HEIGHT + 15) // 16)),
                   "n_contrib_mean": ncontrib_mean, "n_blended_mean": blended_pairs / (WIDTH * HEIGHT),
                   "tile_entries_mean": pool_chunks * 16 / (((WIDTH + 15) // 16) * ((HEIGHT + 15) // 16))},
        "views_per_s": value * 1e6, "hbm_gbs_effective": eff_gbs, "hbm_frac_effective": eff_gbs / peak,
        "stage_ms": per_stage, "kernel_ms_per_step": kernel_ms,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes": ab[dom], "kernel_ms": dom_ms,
                     "peak_source": peak_src,
                     "note": "C=256 blend is fp32-FMA bound by design (no tensor cores, north_star); see DESIGN.md"},
        # the C = 256 blend is three fp32 contractions on the CUDA cores (north_star rules out tensor cores):
        # algorithmic flops = 2*C per blended (pixel, Gaussian) pair for each of forward, s-pass, dL/dfeature
        "fma_roofline": fma_roofline(CHANNELS, blended_pairs, per_stage),
        "e2e": {"value": e2e_value, "unit": "Mviews/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": 35 * 4 + HEIGHT * WIDTH * 4, "d2h_bytes_per_step": 8,
                "api": "render_chn() + semantic.distill_loss_and_grad() + backward; camera + label map from pinned host memory; "
                       "loss read back every step (async 4-byte copy to pinned memory, consumed one step later)"},
        "gpu_launches": int(launches1[0] - launches0[0]), "cub_calls": int(launches1[1] - launches0[1]),
        "clocks": clocks,
    }

    # ---- reference CUDA path on the same GPU (compiled unmodified reference, oracle/_ref) and CPU port
    if world == 1 and not args.no_baselines:
        line["reference_cuda"] = reference_cuda_times(torch, dev, scene, cams_np[0], dL_fixed)
        try:
            smp = cpu_sample(scene, cams_np[0])
            v = 1.0 / smp["seconds_full_view_est"] * 1e-6
            line["cpu_baseline"] = {
                "value": v, "unit": "Mviews/s", "cores": smp["threads"], "kind": "port",
                "sample": f"full preprocess of 1M Gaussians + binning/blend fwd+bwd of rows [{smp['rows'][0]},{smp['rows'][1]}) "
                          f"({smp['frac']:.1%} of pixels) scaled to the view; {smp['seconds_sample']:.1f} s of CPU work"}
        except Exception as ex:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "error": repr(ex)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def reference_cuda_times(torch, dev, scene, cam, dL):
    """The reference's channel-rasterization CUDA path recompiled for sm_100a, timed on this GPU:
    forward by the stock library, backward by the NUM_CHANNELS=256 rebuild."""
    out = {}
    try:
        from oracle import ref as refmod
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from util import dev_cam, dev_scene
        sc, cm = dev_scene(scene, dev), dev_cam(cam, dev)
        bg = torch.zeros(CHANNELS, device=dev)
        kw = dict(bg=bg, means3D=sc["means3D"], opacities=sc["opacities"], viewmatrix=cm["viewmatrix"],
                  projmatrix=cm["projmatrix"], campos=cm["campos"], tanfovx=cm["tanfovx"], tanfovy=cm["tanfovy"],
                  W=WIDTH, H=HEIGHT, colors_precomp=sc["features"], scales=sc["scales"], rotations=sc["rotations"],
                  num_channels=CHANNELS)

        def ev_time(fn, n):
            fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / n
        if refmod.available("chn"):
            r = refmod.RefRasterizer("chn")
            out["fwd_ms"] = ev_time(lambda: r.forward(**kw), 3)
        if refmod.available("chn_c256"):
            r2 = refmod.RefRasterizer("chn_c256")
            r2.forward(**kw)
            out["bwd_ms"] = ev_time(lambda: r2.backward(dL), 1)
        out["kind"] = "unmodified reference cuda_rasterizer compiled for sm_100a (oracle/_ref), debug=False"
    except Exception as ex:  # pragma: no cover
        out["error"] = repr(ex)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-baselines", action="store_true", help="skip the reference-CUDA and CPU-port legs")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_cpu_reference(args, rank, world)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
