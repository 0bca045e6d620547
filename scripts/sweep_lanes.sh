#!/bin/bash
# parameter sweep: lanes on the bench workload (GPU box)
for L in ${LANES:-2 3 4 5 6}; do
  LMG_BENCH_CPU_S=0 LMG_LANES=$L python bench.py --steps 5 --warmup 3 > gpurun_out/sw_l${L}.json 2> gpurun_out/sw_l${L}.err
  python -c "
import json;d=json.load(open('gpurun_out/sw_l${L}.json'));print('lanes',$L, round(d['ms_per_step'],1), round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],1), round(d['roofline']['frac'],3))"
done
