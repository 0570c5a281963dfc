#!/usr/bin/env python
"""Turn ncu output into the small, committed summaries under profiles/.

  python profiles/summarize_ncu.py launches gpurun_out/launches.csv            > profiles/<round>_launches.md
  python profiles/summarize_ncu.py kernel   gpurun_out/prof.ncu-rep  <label>   > profiles/<round>_<label>.md
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
]


def launches_between(path, marker="ray_setup", which=-2):
    """Launch list of one step of a multi-step run: the launches between two consecutive `marker` kernels."""
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    names = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in rows]
    starts = [i for i, (n, _) in enumerate(names) if marker in n]
    step = names[starts[which]:starts[which + 1]] if len(starts) >= 2 and which + 1 != 0 else names[starts[-1]:]
    tot = sum(t for _, t in step)
    agg = collections.OrderedDict()
    for n, t in step:
        k = n.split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    print("| kernel | launches | time (us) | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% |" % (k[:90], c, t / 1e3, 100 * t / tot))
    print("| **total** | %d | %.1f | 100 %% |" % (len(step), tot / 1e3))


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    names = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in rows]
    starts = [i for i, (n, _) in enumerate(names) if "ray_setup" in n]
    step = names[starts[-2]:starts[-1]] if len(starts) >= 2 else names
    tot = sum(t for _, t in step)
    agg = collections.OrderedDict()
    for n, t in step:
        k = n.split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    print("# ncu launch list — one steady-state step (`ncu --metrics gpu__time_duration.sum --clock-control none`)\n")
    print("Per-launch times are cold-cache and serialised under the profiler: compare SHARES, not absolutes.\n")
    print("| kernel | launches | time (us) | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% |" % (k, c, t / 1e3, 100 * t / tot))
    print("| **total** | %d | %.1f | 100 %% |" % (len(step), tot / 1e3))


def kernel(rep, label):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full — %s\n" % label)
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("## %s  (launch id %s)\n" % (d.get("Kernel Name", "?"), d.get("ID", "?")))
        print("| metric | value | unit |\n|---|---:|---|")
        for k in hdr:
            if any(k.endswith(w) or k == w for w in KEYS):
                print("| %s | %s | %s |" % (k, d[k], u[k]))
        print()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "train_launches":
    launches_between(sys.argv[2])
    sys.exit(0)

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
