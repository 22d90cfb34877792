#!/bin/bash
# A/B of programmatic dependent launch (EFFIMVS_PDL) on the bench forward + the GPU suite with it switched on.
# usage (on the GPU box): bash tools/pdl_ab.sh  -> gpurun_out/pdl_ab.log
mkdir -p gpurun_out
{
for v in 0 1 0 1; do
  EFFIMVS_PDL=$v EFFIMVS_BENCH_STRICT=0 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-scene 2>gpurun_out/pdl_bench_$v.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL=$v', round(d['value'],2), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],2), 'roof', d['roofline']['ms'], d['clocks'])"
done
date
EFFIMVS_PDL=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
date
} > gpurun_out/pdl_ab.log 2>&1
cat gpurun_out/pdl_ab.log
