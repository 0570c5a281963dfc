#!/usr/bin/env python
"""bench.py -- faces/s of the GazeNeRF render hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mlp-impl tc|simt] [--faces-per-gpu F]
    python bench.py --workload train [--faces-per-gpu 2]      # BASELINE config[4]: full train step (not the default line)
    python bench.py --workload hier                            # BASELINE config[2]: coarse 64 + FineSample 64 -> 128 samples/ray
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one full drop-in forward("test") -- rays -> both radiance MLPs (fused tcgen05 kernel) -> composite ->
compose -> 2-D neural renderer -> four 512x512 images -- for F faces per GPU (BASELINE config[1]: 64x64 rays x 64 samples,
512x512 output, face + eyes branches, random codes, random-init weights).  Faces are batch-sharded across ranks (weights
replicated); for N > 1 the step ends with the single all-gather of the rendered images (config[3]) -> "scaling": "weak".

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` goes through the same public call with
pinned host buffers (H2D of the inputs and D2H of the four images inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MLP_FLOP_PER_POINT_PER_BRANCH = 3030144          # 2 x 1 515 072 MAC, layers as written in the reference
EXEC_MAC_PER_POINT_PER_BRANCH = 3 * (64 * 384 + 7 * 384 * 384 + 64 * 384 + 384 * 208)  # bf16x3 UMMAs actually issued (folded)
N_RAYS, N_SAMPLES = 64 * 64, 64
METRIC = "faces/s (512x512, 64 samp/ray)"
WORKLOAD = "config[1]: 512x512 render (64x64 rays), 64 samples/ray, face+eye branches, random codes"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d.get("bf16_tflops_sustained"), "bf16_burst": d.get("bf16_tflops"), "hbm": d.get("hbm_gbs"), "src": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


def keep_load_for_sampler(torch, dist, world, dev, ms_per_step, run_step, min_seconds=1.0):
    """The timed region of the default run lasts ~60 ms, shorter than nvidia-smi's start-up + sampling period: keep running the SAME
    step (untimed, after the timed region has been closed) until the sampler has had `min_seconds` of this load to look at.  The
    number of trailing steps is decided by rank 0 and broadcast so that every rank issues the same collectives."""
    n = torch.tensor([max(0, int(min_seconds * 1e3 / max(ms_per_step, 1e-3)) + 1)], device=dev, dtype=torch.int64)
    if world > 1:
        dist.broadcast(n, src=0)
    for _ in range(int(n.item())):
        run_step()
    torch.cuda.synchronize()
    return int(n.item())


def synthetic_inputs(torch, G, opt, faces, seed, device="cpu"):
    """SURVEY §8(d) synthetic inputs: codes ~ N(0, 0.3^2), gaze ~ U(-.5,.5), orbit cameras, the reference's pixel grid."""
    ru = G.RenderUtils(45, "cpu", opt)
    g = torch.Generator().manual_seed(seed)
    shape = torch.randn(faces, 179, generator=g) * 0.3
    appea = torch.randn(faces, 127, generator=g) * 0.3
    gaze = torch.rand(faces, 2, generator=g) - 0.5
    cams = [ru.cam_info_list[(seed * faces + i) % 45] for i in range(faces)]
    kw = dict(batch_xy=ru.ray_xy.expand(faces, -1, -1).contiguous(), batch_uv=None, bg_code=None, shape_code=shape, appea_code=appea,
              gaze_code=gaze, batch_Rmats=torch.cat([c["batch_Rmats"] for c in cams], 0), batch_Tvecs=torch.cat([c["batch_Tvecs"] for c in cams], 0),
              batch_inv_inmats=torch.cat([c["batch_inv_inmats"] for c in cams], 0))
    return kw


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_reference_step(torch, O, sd, oo, kw, ray_step):
    """One bounded sample of the workload on the host CPU with the oracle port of the reference forward:
    both MLPs + composite on every `ray_step`-th ray, and compose + 4 neural-render calls at full size.
    Returns (t_mlp_sample, t_rest) seconds."""
    xy = kw["batch_xy"][:1, :, ::ray_step].contiguous()
    t0 = time.perf_counter()
    smp = O.sample_points(xy, kw["batch_Rmats"][:1], kw["batch_Tvecs"][:1], kw["batch_inv_inmats"][:1], oo.num_sample_coarse, oo.world_z1, oo.world_z2)
    O.render_branches(sd, oo, smp["pts"], smp["z_dists"], smp["zvals"], kw["shape_code"][:1], kw["appea_code"][:1], kw["gaze_code"][:1])
    t1 = time.perf_counter()
    s, c = oo.featmap_size, oo.featmap_nc
    g = torch.Generator().manual_seed(1)
    ff, fe = torch.randn(1, c, s, s, generator=g), torch.randn(1, c, s, s, generator=g)
    af, ae = torch.rand(1, 1, s, s, generator=g), torch.rand(1, 1, s, s, generator=g)
    mf, ep, mg = O.compose_featmaps(ff, af, fe, ae, sd["neural_render.bg_featmap"], kw["gaze_code"][:1])
    for x in (mf, ep, mg, sd["neural_render.bg_featmap"]):
        O.neural_render(sd, x, oo.n_blocks)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def cpu_baseline(torch, G, steps, warmup, ray_step=16):
    from oracle import gazenerf_oracle as O  # test infrastructure; allowed here as the reported CPU baseline only
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    opt = G.BaseOptions()
    torch.manual_seed(45)
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    oo = O.OracleOptions()
    kw = synthetic_inputs(torch, G, opt, 1, 0)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            tm, tr = cpu_reference_step(torch, O, sd, oo, kw, ray_step)
            if i >= warmup:
                times.append(tm * ray_step + tr)  # extrapolate the MLP sample to the full 4096 rays
    t_face = statistics.median(times)
    return {"value": 1.0 / t_face, "unit": "faces/s", "cores": cores, "kind": "port",
            "sample": "oracle port of the reference forward (torch CPU fp32, %d threads): both MLPs + composite on %d of 4096 rays x 64 samples "
                      "(scaled x%d) + compose + 4 neural-render calls at full 512x512; median of %d" % (cores, N_RAYS // ray_step, ray_step, len(times)),
            "s_per_face": t_face}


# ------------------------------------------------------------------------------------------------ train step (BASELINE config[4])
def synthetic_targets(torch, faces, size, seed, device="cpu"):
    """gt ~ U(0,1); head rectangle + two eye rectangles as masks."""
    g = torch.Generator().manual_seed(1000 + seed)
    gt = torch.rand(faces, 3, size, size, generator=g)
    head = torch.zeros(faces, 1, size, size)
    head[:, :, size // 8: size - size // 8, size // 6: size - size // 6] = 1.0
    le, re = torch.zeros_like(head), torch.zeros_like(head)
    le[:, :, size * 3 // 8: size * 7 // 16, size * 5 // 16: size * 7 // 16] = 1.0
    re[:, :, size * 3 // 8: size * 7 // 16, size * 9 // 16: size * 11 // 16] = 1.0
    return {"gt": gt, "head": head, "left_eye": le, "right_eye": re, "full_eye": ((le + re) > 0).float()}


def run_train(args, torch, G, rank, local_rank, world, dev, dist):
    L = G.lib()
    F = args.faces_per_gpu
    opt = G.BaseOptions()
    torch.manual_seed(45)
    from gazenerf_b200.dist import allreduce_gradients
    from gazenerf_b200.trainer_utils import build_code_and_cam

    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False).to(dev).train()
    # the reference's fitting step (trainer/gazenerf_trainer.py:338-528): learnable code offsets + camera deltas next to the network
    off = {"iden": torch.zeros(F, 100, device=dev, requires_grad=True), "expr": torch.zeros(F, 79, device=dev, requires_grad=True),
           "appea": torch.zeros(F, 127, device=dev, requires_grad=True)}
    d_eul = torch.zeros(F, 3, device=dev, requires_grad=True)
    d_tv = torch.zeros(F, 3, 1, device=dev, requires_grad=True)
    lr = 1e-4   # README.md:30 of the reference
    optim = torch.optim.Adam([{"params": list(net.parameters()), "lr": lr}, {"params": list(off.values()), "lr": lr * 1.5},
                              {"params": [d_eul, d_tv], "lr": lr * 0.1}])
    params = [p for p in net.parameters()]   # shared by all ranks (averaged); code offsets / camera deltas belong to this rank's faces
    loss_fn = G.GazeNeRFLoss(eye_loss_importance=1.0, vgg_importance=1.0, use_vgg_loss=False, use_l1_loss=True)
    host_kw = synthetic_inputs(torch, G, opt, F, seed=rank)
    host_tg = synthetic_targets(torch, F, opt.pred_img_size, seed=rank)
    pin = lambda d: {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in d.items()}
    todev = lambda d: {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in d.items()}
    pinned_kw, pinned_tg = pin(host_kw), pin(host_tg)
    dev_kw, dev_tg = todev(host_kw), todev(host_tg)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step(kw, tg):
        base = {"iden": kw["shape_code"][:, :100], "expr": kw["shape_code"][:, 100:], "text": kw["appea_code"][:, :100],
                "illu": kw["appea_code"][:, 100:], "gaze": kw["gaze_code"]}
        cam = {k: kw[k] for k in ("batch_Rmats", "batch_Tvecs", "batch_inv_inmats")}
        code_info, opt_code, cam_info, delta_cam = build_code_and_cam(base, off, cam, 0, F, d_eul, d_tv)
        pred = net("train", kw["batch_xy"], None, **code_info, **cam_info)
        loss = loss_fn.calc_total_loss(delta_cam, opt_code, pred, tg["gt"], tg["head"], tg["full_eye"], tg["left_eye"], tg["right_eye"],
                                       None, None, 0, 0)["total_loss"]
        optim.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:   # data-parallel: one flat all-reduce of the 20 MB of gradients
            allreduce_gradients(params)
        optim.step()
        return loss.detach()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(dev_kw, dev_tg)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = L.gnrf_launch_count()
    barrier()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        step(dev_kw, dev_tg)
        ev[i][1].record()
    barrier()
    launches = (L.gnrf_launch_count() - launches0) // max(args.steps, 1)
    tail = keep_load_for_sampler(torch, dist, world, dev, sum(a.elapsed_time(b) for a, b in ev) / max(args.steps, 1), lambda: step(dev_kw, dev_tg))
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed steps + %d identical untimed trailing steps" % tail
    total_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())

    h2d = sum(v.numel() * v.element_size() for d in (pinned_kw, pinned_tg) for v in d.values() if torch.is_tensor(v))

    def e2e_step():
        loss = step(todev(pinned_kw), todev(pinned_tg))
        loss_host.copy_(loss, non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    if rank == 0:
        peaks = measured_peaks()
        faces_total = world * F * args.steps
        step_s = total_ms * 1e-3 / args.steps
        # forward + input-gradient + weight-gradient GEMMs of both MLPs = 3 x the forward's algorithmic FLOPs
        algo_flop = 3 * 2 * F * N_RAYS * N_SAMPLES * MLP_FLOP_PER_POINT_PER_BRANCH
        achieved = algo_flop / step_s / 1e12
        peak = peaks["bf16_sustained"]
        line = {
            "metric": "train faces/s (512x512, 64 samp/ray; forward + loss + backward + Adam)", "value": faces_total / (total_ms * 1e-3),
            "unit": "faces/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 split (fp32 accumulate) on tensor cores for every GEMM (forward, dX, dW); f32 elsewhere", "data": "synthetic",
            "config": {"workload": "config[4]: full train step (trainer perform_fitting): build_code_and_cam -> two-branch render + neural renderer -> GazeNeRFLoss (l1, no VGG) -> backward to weights, code offsets, camera deltas -> Adam",
                       "faces_per_gpu_per_step": F, "rays": N_RAYS, "samples_per_ray": N_SAMPLES, "l2": "256 MiB memset between timed steps (untimed)",
                       "multi_gpu": "data parallel, one flat gradient all-reduce per step" if world > 1 else "single GPU",
                       "precision_note": "config[4] allows bf16; this path keeps the bf16x3 split so gradients match fp32 autograd to 5e-3"},
            "roofline": {"bound": "tensor", "kernel": "whole step (conv_tc_kernel forward/dX + wgrad_tc_kernel dW dominate)", "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "peak_source": "%s bf16 dense, sustained" % peaks["src"],
                         "traffic": None, "algorithmic_flop_per_step": algo_flop,
                         "note": "MLP GEMM FLOPs as written in the reference, x3 for forward + both gradients; step time, not kernel time"},
            "e2e": {"value": faces_total / e2e_s, "unit": "faces/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mlp-impl", default="tc", choices=["tc", "simt"])
    ap.add_argument("--faces-per-gpu", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="render", choices=["render", "train", "hier"])
    ap.add_argument("--graph", action="store_true", help="single-GPU render: replay the forward from a captured CUDA graph (net.graphed); "
                    "measured +0.7 %% over eager launches -- the host keeps ahead of the GPU anyway -- so eager is the default")
    args = ap.parse_args()
    if args.workload == "train" and args.faces_per_gpu == 1:
        args.faces_per_gpu = 2   # config[4]: batch = 2
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import gazenerf_b200 as G

    if args.impl == "reference":
        # the reference's own CPU implementation of the path (oracle port; the reference is Python and cannot travel),
        # all host threads, rank 0 only
        if rank != 0:
            return
        cb = cpu_baseline(torch, G, args.steps, max(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "faces/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * cb["s_per_face"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "device": "host CPU"},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback of the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from gazenerf_b200.dist import all_gather_images

    if args.workload == "train":
        run_train(args, torch, G, rank, local_rank, world, dev, dist if world > 1 else None)
        if world > 1:
            dist.destroy_process_group()
        return

    L = G.lib()
    F = args.faces_per_gpu
    opt = G.BaseOptions()
    hier = args.workload == "hier"
    if hier:
        opt.num_sample_fine = 64   # 64 coarse + 64(+1) fine -> 128 sorted samples per ray
    torch.manual_seed(45)  # the reference's seed (train.py:53); identical weights on every rank
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=hier, mlp_impl=args.mlp_impl).to(dev).eval()
    host_kw = synthetic_inputs(torch, G, opt, F, seed=rank)
    pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host_kw.items()}
    dev_kw = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host_kw.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # N > 1: the all-gather of the rendered images is fused into the last neural-render kernel (multimem.st over the NVSwitch multicast
    # address / peer stores over NVLink + one device barrier per step, gazenerf_b200/dist.py PeerAllGather); NCCL is the fallback when
    # symmetric memory cannot be set up (and for the hierarchical workload, which renders two image sets)
    peer, gather_mode = None, "single GPU"
    if world > 1:
        gather_mode = "batch-sharded, one NCCL all-gather of the rendered images per step"
        if not hier:
            try:
                from gazenerf_b200.dist import PeerAllGather
                peer = PeerAllGather(F, opt.pred_img_size, dev)
                gather_mode = ("batch-sharded; all-gather fused into the last neural-render kernel (%s) + one device barrier per step"
                               % ("multimem.st via NVSwitch multicast" if peer.use_multicast else "peer stores over NVLink"))
            except Exception as e:  # noqa: BLE001 - any rendezvous failure -> plain NCCL
                peer = None
                gather_mode += " (symmetric memory unavailable: %s)" % str(e)[:80]

    def step(kw, pre_finish=None):
        if peer is not None:   # `peer` is re-bound to None if the self-check below fails
            net.gather_ctx = peer
            local = net("test", **kw)["coarse_dict"]
            net.gather_ctx = None
            if pre_finish is not None:
                pre_finish()
            out = peer.finish()
            out["bg_img"] = local["bg_img"]
            return out
        out = graphed(**kw) if graphed is not None else net("test", **kw)
        out = out["fine_dict"] if hier else out["coarse_dict"]
        if world > 1:
            out = all_gather_images(out, world * F)
        return out

    graphed = None   # single GPU: the whole forward captured once into a CUDA graph (net.graphed), replayed per step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            step(dev_kw)
        barrier()
        if peer is not None:
            # one-time check of the fused gather against the NCCL all-gather of the same local images; every rank must agree,
            # otherwise the run falls back to NCCL (and says so) instead of reporting a number on a wrong result
            net.gather_ctx = peer
            local = net("test", **dev_kw)["coarse_dict"]
            net.gather_ctx = None
            fused = {k: v.clone() for k, v in peer.finish().items()}
            ref = all_gather_images(local, world * F)
            ok = torch.tensor([1 if all(torch.equal(fused[k], ref[k]) for k in fused) else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) != 1:
                peer = None
                gather_mode = "batch-sharded, one NCCL all-gather of the rendered images per step (fused-gather self-check FAILED, not used)"
            barrier()
        use_graph = world == 1 and args.graph
        if use_graph:
            graphed = net.graphed("test", **dev_kw)
            for _ in range(2):
                step(dev_kw)
            barrier()
        # ---------------- device-resident timing: K steps, L2 flushed (untimed) between steps, CUDA events per step
        net.mlp_events = []
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        launches0 = L.gnrf_launch_count()
        barrier()
        for i in range(args.steps):
            flush.zero_()
            ev[i][0].record()
            step(dev_kw)
            ev[i][1].record()
        barrier()
        launches = (L.gnrf_launch_count() - launches0) // max(args.steps, 1)
        step_ms = [a.elapsed_time(b) for a, b in ev]
        if use_graph:
            # kernels replayed from the graph are not seen by the launch counter / the per-kernel events: take the launch count from
            # the capture and time the fused MLP kernel in a few eager steps right after the timed region
            launches = graphed.launches_per_replay
            net.mlp_events = []
            for _ in range(5):
                flush.zero_()
                net("test", **dev_kw)
            torch.cuda.synchronize()
        mlp_ms = [a.elapsed_time(b) for a, b in net.mlp_events]
        net.mlp_events = None
        tail = keep_load_for_sampler(torch, dist if world > 1 else None, world, dev, sum(step_ms) / max(args.steps, 1), lambda: step(dev_kw))
        clocks = sampler.stop() if rank == 0 else None
        if clocks is not None:
            clocks["window"] = "timed steps + %d identical untimed trailing steps" % tail
        total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
        mlp_avg = torch.tensor([sum(mlp_ms) / max(len(mlp_ms), 1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(mlp_avg, op=dist.ReduceOp.MAX)
        total_ms, mlp_avg = float(total_ms.item()), float(mlp_avg.item())

        # ---------------- end-to-end: pinned host inputs -> H2D -> forward -> D2H of the four images, every step.
        # The D2H of step i (12.6 MB per face) runs on a copy stream behind an event and overlaps the compute of step i+1
        # (double-buffered pinned host images); every step's result still reaches the host inside the timed region.
        # N > 1: the gathered batch is identical on every rank and all ranks share one host, so each rank copies ITS faces' slice
        # (and rank 0 the rank-independent bg_img): the host receives every image of the batch exactly once per step
        def host_part(out):
            if world == 1:
                return out
            part = {k: v[rank * F:(rank + 1) * F] for k, v in out.items() if k != "bg_img"}
            if rank == 0:
                part["bg_img"] = out["bg_img"]
            return part

        out0 = host_part(step(dev_kw))
        host_out = [{k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out0.items()} for _ in range(2)]
        h2d = sum(v.numel() * v.element_size() for v in pinned.values() if torch.is_tensor(v))
        d2h = sum(v.numel() * v.element_size() for v in host_out[0].values())   # this rank's share (rank 0: + bg_img)
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream()

        copied = [None]   # event: the previous step's D2H has drained its (symmetric) source buffer
        dev_stage, stage_free = [None, None], [None, None]

        def e2e_step(i):
            kw = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in pinned.items()}
            # fused gather: peers rewrite the buffer of step i-1 once they pass the barrier of step i+1 -> my D2H of step i-1 must be
            # complete before I enter this step's barrier
            out = step(kw, pre_finish=(lambda: main_stream.wait_event(copied[0])) if (peer is not None and copied[0] is not None) else None)
            if graphed is not None:
                # the graph's static outputs are overwritten by the next replay: stage them (D2D, 12.6 MB) into one of two device
                # buffers whose previous D2H has completed
                if dev_stage[i & 1] is None:
                    dev_stage[i & 1] = {k: torch.empty_like(v) for k, v in out.items()}
                    stage_free[i & 1] = None
                if stage_free[i & 1] is not None:
                    main_stream.wait_event(stage_free[i & 1])
                for k, v in out.items():
                    dev_stage[i & 1][k].copy_(v, non_blocking=True)
                out = dev_stage[i & 1]
            done = torch.cuda.Event()
            done.record(main_stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                for k, v in host_part(out).items():
                    host_out[i & 1][k].copy_(v, non_blocking=True)
                    v.record_stream(copy_stream)
                ev_c = torch.cuda.Event()
                ev_c.record(copy_stream)
            copied[0] = ev_c
            stage_free[i & 1] = ev_c

        for i in range(2):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i)
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_s = float(e2e_s.item())

    if rank == 0:
        peaks = measured_peaks()
        faces_total = world * F * args.steps
        value = faces_total / (total_ms * 1e-3)
        algo_flop = 2 * F * N_RAYS * N_SAMPLES * MLP_FLOP_PER_POINT_PER_BRANCH   # per fused-MLP launch (both branches)
        mlp_ms = mlp_avg
        if hier:
            # the coarse pass (64 samples) is the timed launch; the fine pass (128 samples) is a second launch of the same kernel with
            # twice the points: report the step-level rate over all three passes' worth of work (4.766e12 FLOP/face)
            algo_flop *= 3
            mlp_ms = total_ms / args.steps
        achieved = algo_flop / (mlp_ms * 1e-3) / 1e12
        exec_tflops = (3 if hier else 1) * 2 * F * N_RAYS * N_SAMPLES * 2 * EXEC_MAC_PER_POINT_PER_BRANCH / (mlp_ms * 1e-3) / 1e12
        peak = peaks["bf16_sustained"]
        line = {
            "metric": METRIC if not hier else "faces/s (512x512, hierarchical 64 coarse + 64 fine samp/ray)", "value": value, "unit": "faces/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 split (fp32 accumulate) on tensor cores; f32 elsewhere" if args.mlp_impl == "tc" else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD if not hier else "config[2]: hierarchical coarse(64)+fine(64) sampling at 512x512 (FineSample path), face+eye branches",
                       "faces_per_gpu_per_step": F, "rays": N_RAYS, "samples_per_ray": N_SAMPLES, "mlp_impl": args.mlp_impl,
                       "weights": "reference init, torch.manual_seed(45)", "l2": "256 MiB memset between timed steps (untimed)",
                       "launch": "one CUDA graph replay per step (net.graphed)" if use_graph else "eager launches",
                       "multi_gpu": gather_mode},
            "roofline": {"bound": "tensor", "kernel": ("whole step: coarse + fine mlp_tc_kernel launches, fine_depths, 2x neural renderer" if hier else "mlp_tc_kernel (+fold, rgb_head)") if args.mlp_impl == "tc" else "mlp_simt_kernel+composite",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": "%s bf16 dense, sustained (kernel timed inside a long step)" % peaks["src"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of one mlp_tc_kernel launch at F=1, ncu --set full
                         # (profiles/r1_mlp_tc_kernel.md): the packed weights of both branches, once; everything else stays on chip / in L2
                         "traffic": 11.0e6 if (args.mlp_impl == "tc" and F == 1 and not hier) else None, "traffic_unit": "bytes/launch",
                         "kernel_ms": mlp_ms, "algorithmic_flop_per_launch": algo_flop,
                         "executed_mma_tflops": exec_tflops, "executed_frac": exec_tflops / peak,
                         "note": "achieved = reference-as-written FLOPs (3 030 144/point/branch) / time; executed = bf16x3 UMMA FLOPs after exact folds"},
            "e2e": {"value": faces_total / e2e_s, "unit": "faces/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "per rank; with N > 1 every rank copies its own faces of the gathered batch, so the host receives each image once"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(torch, G, steps=3, warmup=1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
