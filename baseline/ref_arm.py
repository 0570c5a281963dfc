"""The reference arm of bench.py: the UNMODIFIED reference ``GazeNeRFNet.forward`` (staged by baseline/stage_ref.py into the
git-ignored ``baseline/_ref/``) run on the host CPU through the reference's own public API and stock code path.

Nothing of gazenerf_b200 (module, kernels, oracle) is on this path: the network is the reference's ``models/gaze_nerf.py`` class, the
rays / cameras come from the reference's ``utils/render_utils.py`` ``RenderUtils``, the weights from the reference's own initialisers
under ``torch.manual_seed(45)`` (``train.py:53``).  The only foreign code is the ``kornia.filters.filter2d`` shim (kornia==0.6.4 is a
pinned, un-vendored, uninstalled dependency of ``models/pixel_shuffle_upsample.py:4``): normalised kernel, reflect border, depthwise
cross-correlation -- kornia's documented semantics, the same shim oracle/gen_golden.py uses to import the reference.
"""
import os
import statistics
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "models", "gaze_nerf.py"))


def _install_kornia_shim(torch):
    if "kornia" in sys.modules and hasattr(sys.modules["kornia"], "filters"):
        return
    F = torch.nn.functional
    k = types.ModuleType("kornia")
    kf = types.ModuleType("kornia.filters")

    def filter2d(x, kernel, border_type="reflect", normalized=False, padding="same"):
        b, c, h, w = x.shape
        tmp = kernel.unsqueeze(1).to(x)
        if normalized:
            tmp = tmp / tmp.abs().sum(dim=(-2, -1), keepdim=True)
        kh, kw = kernel.shape[-2:]
        tmp = tmp.expand(-1, c, -1, -1).reshape(-1, 1, kh, kw)
        xp = F.pad(x, [kw // 2, kw // 2, kh // 2, kh // 2], mode=border_type)
        return F.conv2d(xp, tmp, groups=c)

    kf.filter2d = filter2d
    k.filters = kf
    sys.modules["kornia"] = k
    sys.modules["kornia.filters"] = kf


class ReferenceForward(object):
    """Builds the reference network + inputs once; ``step()`` is one full ``net("test", ...)`` at BASELINE config[1]
    (B = 1, 64x64 rays x 64 samples, 512x512 output, face + eyes branches, four images)."""

    def __init__(self, torch, faces: int = 1, seed: int = 0, num_sample_coarse: int = 64):
        if not available():
            raise RuntimeError("baseline/_ref is not staged (run `python baseline/stage_ref.py` where /root/reference exists)")
        _install_kornia_shim(torch)
        if REF not in sys.path:
            sys.path.insert(0, REF)
        cwd = os.getcwd()
        os.chdir(REF)   # RenderUtils opens configs/config_files/... relative to the cwd (utils/render_utils.py:36)
        try:
            from configs.gazenerf_options import BaseOptions
            from models.gaze_nerf import GazeNeRFNet
            from utils.render_utils import RenderUtils

            self.torch = torch
            opt = BaseOptions()
            opt.num_sample_coarse = num_sample_coarse
            torch.manual_seed(45)   # train.py:53
            self.net = GazeNeRFNet(opt, include_vd=False, hier_sampling=False).eval()
            ru = RenderUtils(45, "cpu", opt)
        finally:
            os.chdir(cwd)
        g = torch.Generator().manual_seed(seed)   # SURVEY §8(d) synthetic inputs (same draws as bench.synthetic_inputs)
        self.shape = torch.randn(faces, 179, generator=g) * 0.3
        self.appea = torch.randn(faces, 127, generator=g) * 0.3
        self.gaze = torch.rand(faces, 2, generator=g) - 0.5
        cams = [ru.cam_info_list[(seed * faces + i) % 45] for i in range(faces)]
        self.cam = {k: torch.cat([c[k] for c in cams], 0) for k in ("batch_Rmats", "batch_Tvecs", "batch_inv_inmats")}
        self.xy = ru.ray_xy.expand(faces, -1, -1)
        self.uv = ru.ray_uv.expand(faces, -1, -1)
        self.faces = faces

    def step(self):
        torch = self.torch
        with torch.no_grad():
            out = self.net("test", self.xy, self.uv, None, self.shape, self.appea, self.gaze, **self.cam)
        return out["coarse_dict"]

    def timed(self, steps: int, warmup: int, threads: int):
        """-> list of seconds per full forward with `threads` intra-op threads."""
        torch = self.torch
        torch.set_num_threads(max(1, threads))
        for _ in range(warmup):
            self.step()
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            self.step()
            ts.append(time.perf_counter() - t0)
        return ts


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:  # noqa: BLE001
        pass
    return "unknown"


def measure(torch, steps: int, warmup: int, threads: int, faces: int = 1):
    rf = ReferenceForward(torch, faces=faces)
    ts = rf.timed(steps, warmup, threads)
    return {"s_per_step": ts, "median_s": statistics.median(ts), "faces": faces, "threads": threads}
