timeout 300 python -m pytest tests/test_lin_hl.py -x -q 2>&1 | tail -3
for p in bf16x3 bf16; do timeout 250 python bench.py --workload train --faces-per-gpu 2 --steps 5 --warmup 3 --train-precision $p > gpurun_out/train_$p.json 2> gpurun_out/train_$p.err; python -c "
import json,sys
d=json.load(open('gpurun_out/train_$p.json')); print('$p', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'])"; tail -2 gpurun_out/train_$p.err; done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:"lin_hl_kernel|wgrad_hl_kernel" -s 60 -c 60 --csv --log-file gpurun_out/hl_launches.csv python tests/prof_train.py bf16x3 2 > gpurun_out/prof_train.log 2>&1; tail -1 gpurun_out/prof_train.log
