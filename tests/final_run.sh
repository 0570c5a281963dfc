# scratch: the end-of-round validation sequence (full GPU suite, default bench line, train bench in every precision mode, launch lists)
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.err
rm -f gpurun_out/r2_bench_train.json
for p in f32 bf16x3 mixed bf16; do timeout 250 python bench.py --workload train --faces-per-gpu 2 --steps 10 --warmup 3 --train-precision $p 2> gpurun_out/train_$p.err | tail -n 1 >> gpurun_out/r2_bench_train.json; done
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print("render", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["clocks"]["reasons"])
for k,v in d["aux"].items(): print(" aux", k, v.get("value"), v.get("ms_per_step"), v.get("roofline",{}).get("frac"), v.get("error"))
for l in open("gpurun_out/r2_bench_train.json"):
    t=json.loads(l); print("train", t["config"]["train_precision"], t["value"], t["ms_per_step"], t["roofline"]["frac"], t["e2e"]["value"])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/train_launches.csv python tests/train_probe.py 2 > gpurun_out/train_probe.log 2>&1; tail -1 gpurun_out/train_probe.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/train_launches_mixed.csv python tests/prof_train.py mixed 3 > gpurun_out/prof_train2.log 2>&1; tail -1 gpurun_out/prof_train2.log
