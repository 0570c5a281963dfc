# scratch: unit tests of the plane kernels + the train bench in two precisions (used while tuning csrc/lin_hl.cu)
timeout 300 python -m pytest tests/test_lin_hl.py -x -q 2>&1 | tail -2
for p in bf16x3 mixed; do timeout 250 python bench.py --workload train --faces-per-gpu 2 --steps 5 --warmup 3 --train-precision $p > gpurun_out/train_$p.json 2> gpurun_out/train_$p.err; python -c "
import json,sys
d=json.load(open('gpurun_out/train_$p.json')); print('$p', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'])"; tail -2 gpurun_out/train_$p.err; done
