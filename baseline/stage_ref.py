"""Stage the UNMODIFIED reference hot-path files into git-ignored ``baseline/_ref/`` so that ``bench.py --impl reference`` can run the
reference's own ``GazeNeRFNet.forward`` on the GPU box's host cores (``/root/reference`` does not exist there; git-ignored files
travel with the gpurun snapshot the same way ``libgnrf.so`` does).

    python baseline/stage_ref.py [--src /root/reference]

The reference is a script tree (no setup.py / pyproject), so ``pip install --target baseline/_ref /root/reference`` is not applicable;
this copy of the seven files SURVEY.md §7-1 lists (+ the intrinsics JSON ``utils/render_utils.py:36`` opens) is the install.  Nothing is
edited, nothing is committed (``baseline/_ref/`` is in .gitignore), and a manifest with the sha256 of every staged file is written next
to them so the run can state exactly what it timed.  The one missing third-party op, ``kornia.filters.filter2d`` (kornia==0.6.4,
requirements.txt:9), is provided at import time by ``baseline/ref_arm.py`` (same shim as oracle/gen_golden.py).
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

FILES = [
    "configs/gazenerf_options.py",
    "configs/config_files/cam_inmat_info_32x32.json",
    "models/gaze_nerf.py",
    "models/mlp_nerf.py",
    "models/neural_renderer.py",
    "models/pixel_shuffle_upsample.py",
    "utils/__init__.py",
    "utils/model_utils.py",
    "utils/render_utils.py",
]


def stage(src: str = "/root/reference", dst: str = DST, quiet: bool = False) -> bool:
    """Copy FILES from `src` to `dst`; returns False (and leaves `dst` alone) when the reference tree is not present."""
    if not os.path.isdir(src):
        return False
    manifest = {"source": src, "files": {}}
    for rel in FILES:
        s = os.path.join(src, rel)
        d = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(d, "rb") as f:
            manifest["files"][rel] = hashlib.sha256(f.read()).hexdigest()
    sub = os.path.join(src, ".SUBMODULES.json")
    if os.path.exists(sub):
        try:
            manifest["submodules"] = json.load(open(sub))
        except Exception:  # noqa: BLE001
            pass
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if not quiet:
        print("staged %d reference files into %s" % (len(FILES), dst))
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("GNRF_REFERENCE", "/root/reference"))
    a = ap.parse_args()
    sys.exit(0 if stage(a.src) else 1)
