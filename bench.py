#!/usr/bin/env python
"""bench.py -- faces/s of the GazeNeRF render hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mlp-impl tc|simt] [--faces-per-gpu F]
    python bench.py --workload train [--faces-per-gpu 2]      # BASELINE config[4]: full train step
    python bench.py --workload hier                            # BASELINE config[2]: coarse 64 + FineSample 64 -> 128 samples/ray
    python bench.py --workload c0                              # BASELINE config[0] on the GPU: 64x64 rays x 32 samples/ray
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one full drop-in forward("test") -- rays -> both radiance MLPs (fused tcgen05 kernel) -> composite ->
compose -> 2-D neural renderer -> four 512x512 images -- for F faces per GPU (BASELINE config[1]: 64x64 rays x 64 samples,
512x512 output, face + eyes branches, random codes, random-init weights).  Faces are batch-sharded across ranks (weights
replicated); for N > 1 the step ends with the single all-gather of the rendered images (config[3]) -> "scaling": "weak".

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` goes through the same public call with
pinned host buffers (H2D of the inputs and D2H of the four images inside the timed region).  At N = 1 the default line also carries
`cpu_baseline` (the UNMODIFIED reference forward from baseline/_ref on the host cores) and `aux` = short runs of the hierarchical
(config[2]) and train-step (config[4]) workloads, each with its own clock record.

`--impl reference`: the reference's own GazeNeRFNet.forward (baseline/_ref, staged by baseline/stage_ref.py; the kornia filter2d shim
is the only foreign code) on the host CPU, every step one FULL 4096-ray forward -- nothing extrapolated.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MLP_FLOP_PER_POINT_PER_BRANCH = 3030144          # 2 x 1 515 072 MAC, layers as written in the reference
EXEC_MAC_PER_POINT_PER_BRANCH = 3 * (64 * 384 + 7 * 384 * 384 + 64 * 384 + 384 * 208)  # bf16x3 UMMAs actually issued (folded)
N_RAYS = 64 * 64
METRIC = "faces/s (512x512, 64 samp/ray)"
WORKLOAD = "config[1]: 512x512 render (64x64 rays), 64 samples/ray, face+eye branches, random codes"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d.get("bf16_tflops_sustained"), "bf16_burst": d.get("bf16_tflops"), "hbm": d.get("hbm_gbs"), "src": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs (B200_PROFILING.md)."""

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


class Env(object):
    """rank / device / process-group handles shared by the workloads."""

    def __init__(self, torch, G, rank, local_rank, world, dev, dist):
        self.torch, self.G, self.rank, self.local_rank, self.world, self.dev, self.dist = torch, G, rank, local_rank, world, dev, dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def keep_load_for_sampler(env, ms_per_step, run_step, min_seconds=1.0):
    """The timed region of the default run lasts ~60 ms, shorter than nvidia-smi's start-up + sampling period: keep running the SAME
    step (untimed, after the timed region has been closed) until the sampler has had `min_seconds` of this load to look at.  The
    number of trailing steps is decided by rank 0 and broadcast so that every rank issues the same collectives."""
    torch = env.torch
    n = torch.tensor([max(0, int(min_seconds * 1e3 / max(ms_per_step, 1e-3)) + 1)], device=env.dev, dtype=torch.int64)
    if env.world > 1:
        env.dist.broadcast(n, src=0)
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


# ------------------------------------------------------------------------------------------------ CPU side: reference / oracle port
def cpu_port_sample(torch, G, steps, warmup, ray_step=16):
    """Fallback when baseline/_ref is not staged: the oracle port (torch CPU fp32) on a bounded ray sample.  kind = "port"."""
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
            if i >= warmup:
                times.append((t1 - t0) * ray_step + (t2 - t1))
    t_face = statistics.median(times)
    return {"value": 1.0 / t_face, "unit": "faces/s", "cores": cores, "kind": "port",
            "sample": "baseline/_ref NOT staged -> oracle port (torch CPU fp32, %d threads): both MLPs + composite on %d of 4096 rays x 64 samples "
                      "(EXTRAPOLATED x%d) + compose + 4 neural-render calls at 512x512; median of %d" % (cores, N_RAYS // ray_step, ray_step, len(times)),
            "s_per_face": t_face}


def cpu_baseline(torch, G, n_forwards=2):
    """`cpu_baseline` of the default line: the UNMODIFIED reference forward (baseline/_ref) on all host cores, `n_forwards` FULL
    forwards of the same workload (~10-30 s of CPU work), nothing extrapolated."""
    from baseline import ref_arm
    if not ref_arm.available():
        return cpu_port_sample(torch, G, steps=3, warmup=1)
    cores = os.cpu_count() or 1
    rf = ref_arm.ReferenceForward(torch, faces=1)
    ts = rf.timed(n_forwards, 0, cores)
    t_face = min(ts)   # no separate warm-up forward (each costs ~10 s): the first one pays the allocator, report the faster
    return {"value": 1.0 / t_face, "unit": "faces/s", "cores": cores, "kind": "reference",
            "sample": "%d full forwards of the unmodified reference GazeNeRFNet('test', B=1, 4096 rays x 64 samples, 4 x 512x512 images) from "
                      "baseline/_ref, torch CPU fp32, %d threads (%s); best of %s s" % (n_forwards, cores, ref_arm.cpu_model(), ["%.2f" % t for t in ts]),
            "s_per_face": t_face}


def reference_arm(args, torch, G):
    """--impl reference: every step = ONE full forward of the unmodified reference on the host CPU (all cores), timed by wall clock;
    plus one single-thread forward (the reference's own setting, train.py:62-64)."""
    from baseline import ref_arm
    base = {"impl": "reference", "metric": METRIC, "unit": "faces/s", "n_gpus": args.gpus, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "gpu_launches": 0}
    if not ref_arm.available():
        cb = cpu_port_sample(torch, G, args.steps, max(args.warmup, 1))
        base.update({"value": cb["value"], "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cb["s_per_face"],
                     "config": {"workload": WORKLOAD, "device": "host CPU", "note": "baseline/_ref missing: oracle port, extrapolated sample"},
                     "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                     "e2e": {"value": cb["value"], "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return base
    cores = os.cpu_count() or 1
    rf = ref_arm.ReferenceForward(torch, faces=1)
    torch.set_num_threads(cores)
    t_start = time.perf_counter()
    for _ in range(args.warmup):
        rf.step()
    ts = []
    for i in range(args.steps):
        t0 = time.perf_counter()
        rf.step()
        ts.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > args.ref_budget_s and i + 1 < args.steps:
            break   # safety valve for a slow host: report the steps actually run
    total = sum(ts)
    value = len(ts) / total
    t1 = rf.timed(1, 0, 1)[0]   # single thread, one full forward
    torch.set_num_threads(cores)
    cb = {"value": value, "unit": "faces/s", "cores": cores, "kind": "reference",
          "sample": "%d full forwards of the unmodified reference GazeNeRFNet('test', B=1, 4096 rays x 64 samples, 4 x 512x512 images) from "
                    "baseline/_ref, torch CPU fp32, %d threads (%s), nothing extrapolated" % (len(ts), cores, ref_arm.cpu_model()),
          "single_thread": {"value": 1.0 / t1, "unit": "faces/s", "cores": 1, "s_per_face": t1,
                            "note": "torch.set_num_threads(1), the reference's own setting (train.py:62-64); one full forward"}}
    base.update({"value": value, "steps": len(ts), "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts),
                 "config": {"workload": WORKLOAD, "device": "host CPU", "faces_per_step": 1, "rays": N_RAYS, "samples_per_ray": 64,
                            "weights": "reference init, torch.manual_seed(45)", "bg_img": "rendered every step (stock reference path)",
                            "third_party": "kornia.filters.filter2d provided by a shim (kornia 0.6.4 not installed)"},
                 "cpu_baseline": cb,
                 "e2e": {"value": value, "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    if len(ts) != args.steps:
        base["config"]["note"] = "stopped after %d of %d steps (--ref-budget-s %d)" % (len(ts), args.steps, args.ref_budget_s)
    return base


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


def run_train(env, steps, warmup, F, with_e2e=True, graph=True, precision="bf16x3"):
    torch, G, dev, world, rank = env.torch, env.G, env.dev, env.world, env.rank
    L = G.lib()
    opt = G.BaseOptions()
    torch.manual_seed(45)
    from gazenerf_b200.dist import allreduce_gradients
    from gazenerf_b200.trainer_utils import build_code_and_cam

    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False).to(dev).train()
    net.train_precision = precision
    # the reference's fitting step (trainer/gazenerf_trainer.py:338-528): learnable code offsets + camera deltas next to the network
    off = {"iden": torch.zeros(F, 100, device=dev, requires_grad=True), "expr": torch.zeros(F, 79, device=dev, requires_grad=True),
           "appea": torch.zeros(F, 127, device=dev, requires_grad=True)}
    d_eul = torch.zeros(F, 3, device=dev, requires_grad=True)
    d_tv = torch.zeros(F, 3, 1, device=dev, requires_grad=True)
    lr = 1e-4   # README.md:30 of the reference
    optim = torch.optim.Adam([{"params": list(net.parameters()), "lr": lr}, {"params": list(off.values()), "lr": lr * 1.5},
                              {"params": [d_eul, d_tv], "lr": lr * 0.1}], capturable=bool(graph))
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

    def split(kw):
        base = {"iden": kw["shape_code"][:, :100], "expr": kw["shape_code"][:, 100:], "text": kw["appea_code"][:, :100],
                "illu": kw["appea_code"][:, 100:], "gaze": kw["gaze_code"]}
        return base, {k: kw[k] for k in ("batch_Rmats", "batch_Tvecs", "batch_inv_inmats")}

    gs = None
    graph_note = "eager launches"
    if graph:
        # the whole step (code / camera assembly, render, loss, backward, [gradient all-reduce,] Adam, on-device jitter) as ONE CUDA graph
        try:
            from gazenerf_b200.trainer_utils import GraphedTrainStep
            b0, c0 = split(dev_kw)
            gs = GraphedTrainStep(net, loss_fn, optim, dev_kw["batch_xy"], {k: v.contiguous() for k, v in b0.items()}, off, c0, dev_tg, d_eul, d_tv,
                                  post_backward=(lambda: allreduce_gradients(params)) if world > 1 else None)
            graph_note = "one CUDA graph replay per step (GraphedTrainStep: %d libgnrf launches + the eager-torch glue captured)" % gs.launches_per_replay
        except Exception as e:  # noqa: BLE001 - fall back to eager launches and say so
            gs = None
            graph_note = "eager launches (graph capture failed: %s)" % str(e)[:120]

    def step(kw, tg):
        if gs is not None:
            b, c = split(kw)
            return gs.step(b, c, tg)["total_loss"]
        base, cam = split(kw)
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

    for _ in range(warmup):
        step(dev_kw, dev_tg)
    env.barrier()
    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches0 = L.gnrf_launch_count()
    env.barrier()
    for i in range(steps):
        flush.zero_()
        ev[i][0].record()
        step(dev_kw, dev_tg)
        ev[i][1].record()
    env.barrier()
    launches = (L.gnrf_launch_count() - launches0) // max(steps, 1)
    if gs is not None:
        launches = gs.launches_per_replay   # replayed kernels are not seen by the launch counter
    ms = sum(a.elapsed_time(b) for a, b in ev)
    tail = keep_load_for_sampler(env, ms / max(steps, 1), lambda: step(dev_kw, dev_tg))
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed steps + %d identical untimed trailing steps" % tail
    total_ms = env.max_over_ranks(ms)

    e2e = None
    if with_e2e:
        h2d = sum(v.numel() * v.element_size() for d in (pinned_kw, pinned_tg) for v in d.values() if torch.is_tensor(v))

        def e2e_step():
            loss = step(todev(pinned_kw), todev(pinned_tg))
            loss_host.copy_(loss, non_blocking=True)

        for _ in range(2):
            e2e_step()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        env.barrier()
        e2e_s = env.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * F * steps / e2e_s, "unit": "faces/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}
    del net, optim
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peaks = measured_peaks()
    faces_total = world * F * steps
    step_s = total_ms * 1e-3 / steps
    # forward + input-gradient + weight-gradient GEMMs of both MLPs = 3 x the forward's algorithmic FLOPs
    algo_flop = 3 * 2 * F * N_RAYS * 64 * MLP_FLOP_PER_POINT_PER_BRANCH
    achieved = algo_flop / step_s / 1e12
    peak = peaks["bf16_sustained"]
    return {
        "metric": "train faces/s (512x512, 64 samp/ray; forward + loss + backward + Adam)", "value": faces_total / (total_ms * 1e-3),
        "unit": "faces/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16x3 split (fp32 accumulate) on tensor cores for every GEMM (forward, dX, dW): per-point activations stored pre-split as bf16 hi/lo planes; f32 elsewhere",
                  "mixed": "forward: bf16x3 split on pre-split bf16 hi/lo planes (as the default); backward (dX, dW of the per-point MLP layers): single-pass bf16 on one bf16 plane per gradient tensor; neural renderer bf16x3; f32 elsewhere",
                  "bf16": "single-pass bf16 (fp32 accumulate) for the per-point MLP GEMMs on one bf16 activation plane; bf16x3 for the neural renderer; f32 elsewhere",
                  "f32": "bf16x3 split (fp32 accumulate) on tensor cores for every GEMM, fp32 activations re-split in-kernel (r1 path); f32 elsewhere"}[precision],
        "data": "synthetic",
        "config": {"workload": "config[4]: full train step (trainer perform_fitting): build_code_and_cam -> two-branch render + neural renderer -> GazeNeRFLoss (l1, no VGG) -> backward to weights, code offsets, camera deltas -> Adam",
                   "faces_per_gpu_per_step": F, "rays": N_RAYS, "samples_per_ray": 64, "l2": "256 MiB memset between timed steps (untimed)",
                   "launch": graph_note,
                   "multi_gpu": "data parallel, one flat gradient all-reduce per step" if world > 1 else "single GPU",
                   "train_precision": precision,
                   "precision_note": "config[4] names bf16; the default keeps the bf16x3 split so gradients match fp32 autograd to 5e-3 (--train-precision bf16 = single pass, tolerance in tests/test_train_grad.py)"},
        "roofline": {"bound": "tensor", "kernel": "whole step (lin_hl_kernel forward/dX + wgrad_hl_kernel dW of the radiance MLPs, conv_tc / wgrad_tc of the neural renderer)", "achieved": achieved,
                     "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "peak_source": "%s bf16 dense, sustained" % peaks["src"],
                     "traffic": None, "algorithmic_flop_per_step": algo_flop,
                     "note": "MLP GEMM FLOPs as written in the reference, x3 for forward + both gradients; step time, not kernel time"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }


# ------------------------------------------------------------------------------------------------ render (configs 0, 1, 2, 3)
def run_render(env, args, steps, warmup, F, workload="render", with_e2e=True):
    torch, G, dev, world, rank, dist = env.torch, env.G, env.dev, env.world, env.rank, env.dist
    from gazenerf_b200.dist import all_gather_images
    L = G.lib()
    opt = G.BaseOptions()
    hier = workload == "hier"
    n_s = 32 if workload == "c0" else 64
    opt.num_sample_coarse = n_s
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
    peer, gather_mode, gather_check = None, "single GPU", None
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
                gather_check = "not run (symmetric memory unavailable)"
                gather_mode += " (symmetric memory unavailable: %s)" % str(e)[:80]

    graphed = None   # the whole forward (incl. the fused gather at N > 1) captured once into a CUDA graph (net.graphed), replayed per step

    def step(kw, pre_finish=None):
        if peer is not None:   # `peer` is re-bound to None if the self-check below fails
            if graphed is not None:
                local = graphed(**kw)["coarse_dict"]      # replays the graph of the symmetric buffer this step writes
            else:
                net.gather_ctx = peer
                local = net("test", **kw)["coarse_dict"]
                net.gather_ctx = None
            if pre_finish is not None:
                pre_finish()
            out = peer.finish(alias=True)
            out["bg_img"] = local["bg_img"]
            return out
        out = graphed(**kw) if graphed is not None else net("test", **kw)
        out = out["fine_dict"] if hier else out["coarse_dict"]
        if world > 1:
            out = all_gather_images(out, world * F)
        return out

    with torch.no_grad():
        for _ in range(warmup):
            step(dev_kw)
        env.barrier()
        if peer is not None:
            # one-time check of the fused gather against the NCCL all-gather of the same local images; every rank must agree,
            # otherwise the run falls back to NCCL (and says so) instead of reporting a number on a wrong result
            net.gather_ctx = peer
            local = net("test", **dev_kw)["coarse_dict"]
            net.gather_ctx = None
            fused = peer.finish()
            ref = all_gather_images(local, world * F)
            ok = torch.tensor([1 if all(torch.equal(fused[k], ref[k]) for k in fused) else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            gather_check = "bit-identical" if int(ok.item()) == 1 else "failed"
            if int(ok.item()) != 1:
                peer = None
                gather_mode = "batch-sharded, one NCCL all-gather of the rendered images per step (fused-gather self-check FAILED, not used)"
            env.barrier()
        use_graph = (not args.no_graph) and not hier and (world == 1 or peer is not None)
        graph_note = "eager launches"
        if use_graph:
            try:
                graphed = net.graphed("test", gather=peer, **dev_kw)
                graph_note = "one CUDA graph replay per step (net.graphed%s)" % (": one graph per rotating symmetric buffer, fused gather inside" if peer is not None else "")
            except Exception as e:  # noqa: BLE001 - every rank fails alike (same code path); fall back to eager launches and say so
                graphed, use_graph = None, False
                net.gather_ctx = None
                graph_note = "eager launches (graph capture failed: %s)" % str(e)[:100]
            for _ in range(2):
                step(dev_kw)
            env.barrier()
        # ---------------- device-resident timing: K steps, L2 flushed (untimed) between steps, CUDA events per step
        net.mlp_events = []
        sampler = ClockSampler(env.local_rank)
        if rank == 0:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        launches0 = L.gnrf_launch_count()
        env.barrier()
        for i in range(steps):
            flush.zero_()
            ev[i][0].record()
            step(dev_kw)
            ev[i][1].record()
        env.barrier()
        launches = (L.gnrf_launch_count() - launches0) // max(steps, 1)
        step_ms = [a.elapsed_time(b) for a, b in ev]
        if use_graph:
            # kernels replayed from the graph are not seen by the launch counter / the per-kernel events: take the launch count from
            # the capture and time the fused MLP kernel in a few eager steps right after the timed region
            launches = graphed.launches_per_replay
            net.mlp_events = []
            g_keep, graphed = graphed, None        # a few EAGER steps (same collectives on every rank) to time the fused MLP launch
            for _ in range(5):
                flush.zero_()
                step(dev_kw)
            graphed = g_keep
            torch.cuda.synchronize()
        mlp_ms = [a.elapsed_time(b) for a, b in net.mlp_events]
        net.mlp_events = None
        tail = keep_load_for_sampler(env, sum(step_ms) / max(steps, 1), lambda: step(dev_kw))
        clocks = sampler.stop() if rank == 0 else None
        if clocks is not None:
            clocks["window"] = "timed steps + %d identical untimed trailing steps" % tail
        total_ms = env.max_over_ranks(sum(step_ms))
        mlp_avg = env.max_over_ranks(sum(mlp_ms) / len(mlp_ms) if mlp_ms else sum(step_ms) / max(steps, 1))   # simt path: whole step

        e2e = None
        if with_e2e:
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
            copied = [None, None]   # events of the last two steps' D2H copies (they read the aliased symmetric buffers)
            dev_stage, stage_free = [None, None], [None, None]

            # graph replay: the step's inputs travel as ONE pinned flat buffer in the layout of the graph's static input buffer
            # (GraphedForward.pack_inputs) -> one H2D copy per step instead of nine small ones
            packed = graphed.pack_inputs(pinned) if graphed is not None else None
            if packed is not None:
                h2d = packed.numel() * packed.element_size()

            # ... and that copy is issued one step AHEAD on its own stream into one of two device staging buffers (the step then starts
            # with a 34 KB device-to-device copy), so the PCIe read of step i+1's inputs never sits on the critical path behind the
            # 12.6 MB D2H burst of step i.  Every step's inputs still cross PCIe inside the timed region (the first timed step copies
            # its own inputs at its start).
            in_stream = torch.cuda.Stream(device=dev)
            in_stage = [torch.empty_like(graphed.flat_input) for _ in range(2)] if packed is not None else None
            in_ready, in_read = [None, None], [None, None]   # H2D into the staging buffer done / its last D2D read done

            def prefetch_inputs(i):
                with torch.cuda.stream(in_stream):
                    if in_read[i & 1] is not None:
                        in_stream.wait_event(in_read[i & 1])   # step i-2's D2D read of this staging buffer (long complete)
                    in_stage[i & 1].copy_(packed, non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record(in_stream)
                in_ready[i & 1] = ev_in

            def e2e_step(i, last=False):
                if packed is not None:
                    if in_ready[i & 1] is None:
                        prefetch_inputs(i)
                    main_stream.wait_event(in_ready[i & 1])
                    graphed.flat_input.copy_(in_stage[i & 1], non_blocking=True)
                    in_read[i & 1] = torch.cuda.Event()
                    in_read[i & 1].record(main_stream)
                    in_ready[i & 1] = None
                    kw = {}
                else:
                    kw = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in pinned.items()}
                # fused gather, three rotating buffers: peers rewrite the buffer of step i-2 once they pass the barrier of step i+1, and
                # they cannot pass it before I enter it -> my D2H of step i-2 must be complete before I enter the barrier of step i
                # (two whole steps of slack: this wait never blocks in practice)
                out = step(kw, pre_finish=(lambda: main_stream.wait_event(copied[0])) if (peer is not None and copied[0] is not None) else None)
                if packed is not None and not last:
                    prefetch_inputs(i + 1)
                if graphed is not None and peer is None:
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
                copied[0], copied[1] = copied[1], ev_c
                stage_free[i & 1] = ev_c

            for i in range(2):
                e2e_step(i, last=(i == 1))
            env.barrier()
            t0 = time.perf_counter()
            for i in range(steps):
                e2e_step(i, last=(i == steps - 1))
            env.barrier()
            e2e_s = env.max_over_ranks(time.perf_counter() - t0)
            e2e = {"value": world * F * steps / e2e_s, "unit": "faces/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "note": "per rank; with N > 1 every rank copies its own faces of the gathered batch, so the host receives each image once"}
    del net, graphed
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peaks = measured_peaks()
    faces_total = world * F * steps
    value = faces_total / (total_ms * 1e-3)
    algo_flop = 2 * F * N_RAYS * n_s * MLP_FLOP_PER_POINT_PER_BRANCH   # per fused-MLP launch (both branches)
    mlp_ms = mlp_avg
    if hier:
        # the coarse pass (64 samples) is the timed launch; the fine pass (128 samples) is a second launch of the same kernel with
        # twice the points: report the step-level rate over all three passes' worth of work (4.766e12 FLOP/face)
        algo_flop *= 3
        mlp_ms = total_ms / steps
    achieved = algo_flop / (mlp_ms * 1e-3) / 1e12
    exec_tflops = (3 if hier else 1) * 2 * F * N_RAYS * n_s * 2 * EXEC_MAC_PER_POINT_PER_BRANCH / (mlp_ms * 1e-3) / 1e12
    # the timed region is K x ~3 ms: the kernel is "timed alone" in MEASURED_PEAKS terms -> the BURST peak is the denominator
    # (the fraction against the sustained figure is kept beside it; under >= 1 s of this load the board sits at its power cap)
    peak = peaks["bf16_burst"]
    metric = {"render": METRIC, "hier": "faces/s (512x512, hierarchical 64 coarse + 64 fine samp/ray)",
              "c0": "faces/s (512x512 from 64x64 rays, 32 samp/ray)"}[workload]
    wl = {"render": WORKLOAD, "hier": "config[2]: hierarchical coarse(64)+fine(64) sampling at 512x512 (FineSample path), face+eye branches",
          "c0": "config[0] on the GPU: 64x64 low-res feature render, 32 samples/ray, random latent+gaze codes (+ neural renderer to 512x512)"}[workload]
    line = {
        "metric": metric, "value": value, "unit": "faces/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 split (fp32 accumulate) on tensor cores; f32 elsewhere" if args.mlp_impl == "tc" else "f32",
        "data": "synthetic",
        "config": {"workload": wl, "faces_per_gpu_per_step": F, "rays": N_RAYS, "samples_per_ray": n_s, "mlp_impl": args.mlp_impl,
                   "weights": "reference init, torch.manual_seed(45)", "l2": "256 MiB memset between timed steps (untimed)",
                   "bg_img": "cached per weight version (parameter-only input, models/gaze_nerf.py:175-176): the timed step renders 3 of the 4 images",
                   "launch": graph_note,
                   "multi_gpu": gather_mode},
        "roofline": {"bound": "tensor", "kernel": ("whole step: coarse + fine mlp_tc_kernel launches, fine_depths, 2x neural renderer" if hier else "mlp_tc_kernel (+rgb_head_kernel, launched back to back by the same C call)") if args.mlp_impl == "tc" else "mlp_simt_kernel+composite",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "peak_source": "%s bf16 dense, burst (cuBLAS best-of-10; the timed region is K x %.1f ms)" % (peaks["src"], total_ms / steps),
                     "frac_vs_sustained_peak": achieved / peaks["bf16_sustained"], "sustained_peak": peaks["bf16_sustained"],
                     # dram__bytes_read.sum + dram__bytes_write.sum of one mlp_tc_kernel launch at F=1, ncu --set full
                     # (profiles/r1_mlp_tc_kernel.md): the packed weights of both branches, once; everything else stays on chip / in L2
                     "traffic": 11.0e6 if (args.mlp_impl == "tc" and F == 1 and workload == "render") else None, "traffic_unit": "bytes/launch",
                     "kernel_ms": mlp_ms, "algorithmic_flop_per_launch": algo_flop,
                     "executed_mma_tflops": exec_tflops, "executed_frac": exec_tflops / peak,
                     "note": "achieved = reference-as-written FLOPs (3 030 144/point/branch) / time; executed = bf16x3 UMMA FLOPs after exact folds"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if world > 1:
        line["gather_check"] = gather_check
    return line


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
    ap.add_argument("--no-aux", action="store_true", help="skip the short hier / train runs attached to the default N=1 line")
    ap.add_argument("--ref-budget-s", type=int, default=300, help="--impl reference: stop issuing timed steps after this many seconds")
    ap.add_argument("--workload", default="render", choices=["render", "train", "hier", "c0"])
    ap.add_argument("--no-graph", action="store_true", help="render workloads: eager launches instead of replaying the forward (incl. the fused "
                    "all-gather at N > 1) from a captured CUDA graph (net.graphed)")
    ap.add_argument("--train-precision", default="bf16x3", choices=["bf16x3", "mixed", "bf16", "f32"],
                    help="train workload: storage of the per-point MLP activations (gazenerf_b200/train.py)")
    ap.add_argument("--no-train-graph", action="store_true", help="train workload: eager launches instead of the captured step (GraphedTrainStep)")
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
        if rank != 0:   # rank 0 alone runs the CPU arm
            return
        print(json.dumps(reference_arm(args, torch, G)))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback of the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    env = Env(torch, G, rank, local_rank, world, dev, dist)

    if args.workload == "train":
        line = run_train(env, args.steps, args.warmup, args.faces_per_gpu, graph=not args.no_train_graph, precision=args.train_precision)
    else:
        line = run_render(env, args, args.steps, args.warmup, args.faces_per_gpu, workload=args.workload)
        if args.workload == "render" and world == 1 and not args.no_aux and args.mlp_impl == "tc":
            # BASELINE configs 2 and 4 measured by the same driver-run command (short: 5 steps each, own clock record)
            aux = {}
            for name, fn in (("hier", lambda: run_render(env, args, 5, 3, 1, workload="hier", with_e2e=False)),
                             ("c0", lambda: run_render(env, args, 5, 3, 1, workload="c0", with_e2e=False)),
                             ("train", lambda: run_train(env, 5, 3, 2, with_e2e=False, graph=not args.no_train_graph)),
                             ("train_mixed", lambda: run_train(env, 5, 3, 2, with_e2e=False, graph=not args.no_train_graph,
                                                               precision="mixed"))):
                try:
                    r = fn()
                    aux[name] = {k: r[k] for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "dtype", "config", "roofline",
                                                   "gpu_launches", "clocks")}
                except Exception as e:  # noqa: BLE001 - the headline line must still be printed
                    aux[name] = {"error": str(e)[:300]}
            line["aux"] = aux
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(torch, G)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
