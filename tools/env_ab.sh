#!/bin/bash
# A/B of one environment switch on the short bench line: tools/env_ab.sh NAME valueA valueB [reps]
name=$1; a=$2; b=$3; reps=${4:-2}
for i in $(seq $reps); do for v in "$a" "$b"; do
  env $name=$v EFFIMVS_BENCH_STRICT=0 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-scene 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name=$v', round(d['ms_per_step'],4), round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches_per_step'])"
done; done
