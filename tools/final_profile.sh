#!/bin/bash
# round-end measurements on the GPU box: bench lines (DTU with all legs, Tanks & Temples), then the ncu launch list of one forward
# usage: bash tools/final_profile.sh <tag>   -> gpurun_out/<tag>_*
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "dtu rc $?"
timeout 600 python bench.py --shape tanks --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_tanks_n1.json 2> gpurun_out/${tag}_bench_tanks_n1.err; echo "tanks rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_forward_launches.csv python bench.py --profile-one > gpurun_out/${tag}_ncu_one.log 2>&1; echo "ncu rc $?"
python tools/summarize_launches.py gpurun_out/${tag}_forward_launches.csv 60 > gpurun_out/${tag}_forward_summary.txt
head -4 gpurun_out/${tag}_forward_summary.txt
python -c "
import json,sys
for f in ('gpurun_out/${tag}_bench_n1.json','gpurun_out/${tag}_bench_tanks_n1.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],2), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],2), d['roofline']['frac'] if d.get('roofline') else None, d['clocks'])
"
