"""ncu launch-list CSV (`--metrics gpu__time_duration.sum --csv`) -> per-launch table of the LAST full step (between the last two
mlp_tc_kernel launches) + per-kernel totals.  usage: python profiles/launch_table.py gpurun_out/launches.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, mi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
data = [r for r in rows[hi + 2:] if len(r) > mi]
idx = [i for i, r in enumerate(data) if "mlp_tc_kernel" in r[ki]]
anchor = "ray_setup_kernel"
starts = [i for i, r in enumerate(data) if anchor in r[ki]]
s, e = starts[-2], starts[-1]
tot = collections.OrderedDict()
print("| # | kernel | grid | time (us) |\n|---:|---|---|---:|")
for n, r in enumerate(data[s:e]):
    name = r[ki].split("(")[0].replace("void ", "")
    t = float(r[mi].replace(",", "")) / 1e3
    tot.setdefault(name, [0, 0.0])
    tot[name][0] += 1; tot[name][1] += t
    print("| %d | `%s` | %s | %.1f |" % (n, name[:60], r[gi], t))
all_t = sum(v[1] for v in tot.values())
print("\n| kernel | launches | time (us) | share |\n|---|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% |" % (k[:60], v[0], v[1], 100 * v[1] / all_t))
print("| **total** | %d | %.1f | 100 %% |" % (sum(v[0] for v in tot.values()), all_t))
