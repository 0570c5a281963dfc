"""Build + load libgnrf.so (the C-ABI shared library declared in include/gnrf.h) and bind it with ctypes.

The library is compiled in-tree with nvcc for sm_100a only.  There is NO fallback: if the library is missing or
the device is not a B200, every entry point raises (the product path must fail loudly, never route to a CPU
or eager-PyTorch implementation).
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from typing import List, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(_ROOT, "include")
LIB_PATH = os.path.join(_HERE, "libgnrf.so")

SOURCES = ["abi.cu", "geometry.cu", "mlp_simt.cu", "mlp_tc.cu", "compose.cu", "neural_render.cu", "conv_tc.cu", "wgrad_tc.cu", "lin_hl.cu",
           "train_ops.cu", "nr_train.cu", "loss.cu", "data.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v", "-diag-suppress=20013,20015",
]
OBJ_DIR = os.path.join(_HERE, "build")

SYMBOLS = [
    "gnrf_abi_version", "gnrf_last_error", "gnrf_device_check", "gnrf_launch_count",
    "gnrf_ray_setup", "gnrf_coarse_depths", "gnrf_fine_depths",
    "gnrf_mlp_simt_fwd", "gnrf_composite_fwd",
    "gnrf_mlp_tc_packed_bytes", "gnrf_mlp_tc_bias_floats", "gnrf_mlp_tc_pack", "gnrf_mlp_tc_fold",
    "gnrf_mlp_tc_workspace_bytes", "gnrf_mlp_tc_fwd", "gnrf_mlp_tc_fwd_debug",
    "gnrf_mlp_tc_pack_vd", "gnrf_mlp_tc_vd_bias", "gnrf_mlp_tc_fwd_vd",
    "gnrf_compose_fwd", "gnrf_nr_workspace_bytes", "gnrf_neural_render_fwd",
    "gnrf_nr_tc_packed_bytes", "gnrf_nr_tc_pack", "gnrf_neural_render_tc_fwd", "gnrf_neural_render_tc_layerwise_fwd",
    "gnrf_neural_render_tc_fwd_gather",
    # training path
    "gnrf_conv_tc_packed_bytes", "gnrf_conv_tc_pack", "gnrf_conv_tc", "gnrf_wgrad_tc_workspace_bytes", "gnrf_wgrad_tc",
    "gnrf_pe_fwd", "gnrf_pe_bwd", "gnrf_composite_cm_fwd", "gnrf_composite_cm_bwd", "gnrf_geom_bwd",
    "gnrf_lin_hl_packed_bytes", "gnrf_lin_hl_pack", "gnrf_lin_hl", "gnrf_wgrad_hl_workspace_bytes", "gnrf_wgrad_hl", "gnrf_pe_fwd_hl",
    "gnrf_composite_cm_bwd_hl",
    "gnrf_compose_bwd_blocks", "gnrf_compose_bwd_groups", "gnrf_compose_bwd",
    "gnrf_nr_train_saved_bytes", "gnrf_nr_train_fwd", "gnrf_nr_train_bwd_workspace_bytes", "gnrf_nr_train_bwd",
    "gnrf_data_loss_workspace_floats", "gnrf_data_loss_fwd", "gnrf_data_loss_bwd",
    # dataset sample -> device tensors
    "gnrf_sample_images_to_device", "gnrf_sample_meta_to_device",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libgnrf.so")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "gnrf.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(src: str, obj: str, headers_mtime: float, force: bool):
    """nvcc -c one source (skipped when the object is newer than the source and every header)."""
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), headers_mtime):
        return src, 0, "(up to date)\n", None
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-c", "-o", obj, src]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, proc.returncode, proc.stdout, " ".join(cmd)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo; cross-compiles without a GPU)
    into gazenerf_b200/libgnrf.so: one object per source, compiled in parallel and re-used while the source and headers are
    unchanged, then one link."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [os.path.join(INCLUDE, "gnrf.h")]
    hm = max(os.path.getmtime(h) for h in headers)
    jobs = [(os.path.join(CSRC, s), os.path.join(OBJ_DIR, s[:-3] + ".o")) for s in SOURCES]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda j: _compile_one(j[0], j[1], hm, force), jobs))
    text = ""
    for src, rc, out, cmd in results:
        text += "==== %s\n%s\n%s" % (os.path.basename(src), cmd or "", out)
        if rc != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", tmp] + [o for _, o in jobs]
    proc = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout)
    os.replace(tmp, LIB_PATH)
    with open(os.path.join(_HERE, "build.log"), "w") as f:
        f.write(text + "==== link\n" + " ".join(link) + "\n" + proc.stdout)
    if verbose:
        print(text)
    return LIB_PATH


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    """Load (building first if needed and possible) and return the bound library."""
    global _lib
    if _lib is not None:
        return _lib
    if _stale():
        try:
            build()
        except Exception as e:  # on a box without nvcc a prebuilt, possibly "stale by mtime" .so is still fine
            if not os.path.exists(LIB_PATH):
                raise RuntimeError("libgnrf.so is missing and cannot be built: %s" % e)
    L = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i32, f32, sz = c.c_void_p, c.c_int, c.c_float, c.c_size_t
    L.gnrf_abi_version.restype = i32
    L.gnrf_last_error.restype = c.c_char_p
    L.gnrf_device_check.restype = i32
    L.gnrf_launch_count.restype = c.c_ulonglong
    L.gnrf_ray_setup.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    L.gnrf_coarse_depths.argtypes = [vp, vp, vp, i32, i32, i32, f32, f32, vp, vp]
    L.gnrf_fine_depths.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    L.gnrf_mlp_simt_fwd.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    L.gnrf_composite_fwd.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    L.gnrf_mlp_tc_packed_bytes.restype = sz
    L.gnrf_mlp_tc_bias_floats.restype = sz
    L.gnrf_mlp_tc_pack.argtypes = [vp, vp, vp]
    L.gnrf_mlp_tc_fold.argtypes = [vp, vp, vp, i32, vp, vp]
    L.gnrf_mlp_tc_workspace_bytes.restype = sz
    L.gnrf_mlp_tc_workspace_bytes.argtypes = [i32, i32, i32]
    L.gnrf_mlp_tc_fwd.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, sz, vp]
    L.gnrf_mlp_tc_pack_vd.argtypes = [vp, i32, vp, vp]
    L.gnrf_mlp_tc_vd_bias.argtypes = [vp, vp, i32, i32, vp, vp]
    L.gnrf_mlp_tc_fwd_vd.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, sz, vp]
    L.gnrf_mlp_tc_fwd_debug.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, sz, vp, vp, i32, vp]
    L.gnrf_compose_fwd.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]
    L.gnrf_nr_workspace_bytes.restype = sz
    L.gnrf_nr_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.gnrf_neural_render_fwd.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, sz, vp]
    L.gnrf_nr_tc_packed_bytes.restype = sz
    L.gnrf_nr_tc_packed_bytes.argtypes = [i32, i32, i32]
    L.gnrf_nr_tc_pack.argtypes = [vp, i32, i32, i32, vp, vp]
    L.gnrf_neural_render_tc_fwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, sz, vp]
    L.gnrf_neural_render_tc_layerwise_fwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, sz, vp]
    L.gnrf_neural_render_tc_fwd_gather.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, sz, vp, vp, i32, i32, i32, i32, vp]
    i64 = c.c_longlong
    L.gnrf_conv_tc_packed_bytes.restype = sz
    L.gnrf_conv_tc_packed_bytes.argtypes = [i32, i32]
    L.gnrf_conv_tc_pack.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.gnrf_conv_tc.argtypes = [vp, i32, i32, vp, i64, vp, vp, i64, i32, vp, i64, i32, f32, vp, i64, i32, i32, i32, vp]
    L.gnrf_wgrad_tc_workspace_bytes.restype = sz
    L.gnrf_wgrad_tc_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.gnrf_wgrad_tc.argtypes = [vp, i64, vp, i64, i32, i32, i32, i32, vp, vp, i32, i32, vp, sz, vp]
    L.gnrf_pe_fwd.argtypes = [vp, vp, vp, i32, i32, i32, vp, i64, vp]
    L.gnrf_pe_fwd_hl.argtypes = [vp, vp, vp, i32, i32, i32, vp, i64, vp, i64, i64, i32, vp]
    L.gnrf_lin_hl_packed_bytes.restype = sz
    L.gnrf_lin_hl_packed_bytes.argtypes = [i32, i32, i32]
    L.gnrf_lin_hl_pack.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.gnrf_lin_hl.argtypes = [vp, i32, i32, i32, vp, i64, i64, vp, i32, i32, vp, i64, i64, i32, vp, i64, vp, i64, i32, vp, i64, i32, i32, i32, vp]
    L.gnrf_wgrad_hl_workspace_bytes.restype = sz
    L.gnrf_wgrad_hl_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.gnrf_wgrad_hl.argtypes = [vp, i64, i64, vp, i64, i64, i32, i32, i32, i32, i32, vp, vp, i32, vp, sz, vp]
    L.gnrf_composite_cm_bwd_hl.argtypes = [vp, vp, vp, i64, vp, i64, vp, vp, vp, i32, i32, i32, i32, vp, i64, i64, vp, i64, i64, i32, vp,
                                           vp, vp]
    L.gnrf_pe_bwd.argtypes = [vp, i64, vp, i64, vp, i64, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    L.gnrf_composite_cm_fwd.argtypes = [vp, i64, vp, i64, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    L.gnrf_composite_cm_bwd.argtypes = [vp, vp, vp, i64, vp, i64, vp, vp, vp, i32, i32, i32, i32, vp, i64, vp, i64, vp, vp, vp]
    L.gnrf_geom_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]
    L.gnrf_compose_bwd_blocks.argtypes = [i32, i32]
    L.gnrf_compose_bwd_groups.argtypes = [i32]
    L.gnrf_compose_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.gnrf_nr_train_saved_bytes.restype = sz
    L.gnrf_nr_train_saved_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.gnrf_nr_train_fwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, sz, vp]
    L.gnrf_nr_train_bwd_workspace_bytes.restype = sz
    L.gnrf_nr_train_bwd_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.gnrf_nr_train_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, sz, vp]
    L.gnrf_data_loss_workspace_floats.restype = sz
    L.gnrf_data_loss_fwd.argtypes = [vp] * 9 + [i32, i32, i32, f32, vp, vp, vp, vp]
    L.gnrf_data_loss_bwd.argtypes = [vp] * 9 + [i32, i32, i32, f32, vp, vp, vp, vp, vp, vp, vp]
    L.gnrf_sample_images_to_device.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    L.gnrf_sample_meta_to_device.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is c.c_int and name not in ("gnrf_abi_version",):
            fn.restype = i32
    _lib = L
    return L


def check(code: int, what: str = "") -> None:
    """Raise RuntimeError carrying gnrf_last_error() for a non-zero status (the reference's train loop catches
    Exception per batch, trainer/gazenerf_trainer.py:576-582)."""
    if code != 0:
        msg = lib().gnrf_last_error().decode("utf-8", "replace")
        raise RuntimeError("libgnrf %s failed (code %d): %s" % (what, code, msg))


def ptr_array(ptrs: List[int]):
    """Host array of device pointers (const float* const*)."""
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    return arr
