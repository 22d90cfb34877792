#!/bin/bash
# one GPU-box round trip: the GPU suite, then a short bench line (no CPU / scene / strict legs) -> gpurun_out/quick.log
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
EFFIMVS_BENCH_STRICT=0 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-scene 2>gpurun_out/quick_bench.err > gpurun_out/quick_bench.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/quick_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 2), d["clocks"])
print("roofline", d["roofline"]["ms"], d["roofline"]["frac"])
for k, v in d["kernels"].items():
    print(" ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in ("ms", "gbs", "frac")})
PY
} > gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
