#!/bin/bash
# parameter sweep: lanes x gate mode on the bench workload (GPU box)
for G in 0 1 2; do for L in 2 3 4; do
  LMG_BENCH_CPU_S=0 LMG_GATE=$G LMG_LANES=$L python bench.py --steps 4 --warmup 3 > gpurun_out/sw_g${G}_l${L}.json 2> gpurun_out/sw_g${G}_l${L}.err
  python -c "
import json;d=json.load(open('gpurun_out/sw_g${G}_l${L}.json'));print('gate',$G,'lanes',$L, round(d['ms_per_step'],1), round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],1), round(d['roofline']['frac'],3))"
done; done
LMG_BENCH_CPU_S=0 LMG_LANES=1 python bench.py --steps 4 --warmup 3 > gpurun_out/sw_l1.json 2> gpurun_out/sw_l1.err
python -c "
import json;d=json.load(open('gpurun_out/sw_l1.json'));print('lanes 1', round(d['ms_per_step'],1), round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],1), round(d['roofline']['frac'],3))"
