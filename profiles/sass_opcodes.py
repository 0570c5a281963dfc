"""Per-kernel SASS opcode histogram of the in-tree libgnrf.so (evidence that the hot kernels are Blackwell-native):
    python profiles/sass_opcodes.py > profiles/r2_sass_opcodes.md
Counts the tcgen05 / TMA / TMEM / multicast mnemonics B200_PROFILING.md names."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gazenerf_b200", "libgnrf.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTCATOMSWS", "ELECT", "MULTIMEM", "STG.E.128.STRONG.SYS", "HMMA", "FFMA"]
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    if kern is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    total[kern] += 1
    for o in OPS:
        if op.startswith(o) or (o == "MULTIMEM" and "MULTIMEM" in line.upper()):
            counts[kern][o] += 1
            if o == "UBLKCP" and "MULTICAST" in op:
                counts[kern]["UBLKCP.MULTICAST"] += 1
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n
print("# r2 — SASS opcode histogram of `gazenerf_b200/libgnrf.so` (`cuobjdump -sass`, sm_100a)\n")
print("`UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st (TMEM), `UBLKCP` = cp.async.bulk (TMA bulk copy; `.MULTICAST` = cluster multicast),")
print("`UTMALDG` / `UTMASTG` = cp.async.bulk.tensor load / store (tensor-map TMA), `UTCBAR` = tcgen05.commit -> mbarrier, `SYNCS` = mbarrier ops, `ELECT` = elect.sync.\n")
cols = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UBLKCP.MULTICAST", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "ELECT", "FFMA"]
print("| kernel | SASS instr | " + " | ".join(cols) + " |\n|---|---:|" + "---:|" * len(cols))
for k, c in counts.items():
    if total[k] < 50:
        continue
    name = demangle(k)
    hot = any(c[o] for o in ("UTCHMMA", "UBLKCP", "UTMALDG", "LDTM"))
    print("| `%s` | %d | " % (name[:70], total[k]) + " | ".join(str(c[o]) if c[o] else ("·") for o in cols) + " |")
# multimem store of the fused all-gather
mm = [l for l in out.splitlines() if "STG.E.128.STRONG.SYS" in l]
print("\nFused all-gather store (`multimem.st.relaxed.sys.global.v4.f32` in `blur_lrelu_rgb_kernel`): %d `STG.E.128.STRONG.SYS` instruction(s) in the library." % len(mm))
