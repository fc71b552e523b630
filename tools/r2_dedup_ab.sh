#!/bin/bash
# cache-pass A/B on one GPU: parity tests, single-image latency and the 4K / 1080p bench legs per variant
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_headline.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${tag}_pytest.log
for v in "$@"; do
  name=$(echo "$v" | tr ' =' '__')
  echo "== $v"
  env $v timeout 300 python tools/latency.py --sizes 1080x1920,2160x3840 --reps 20 > gpurun_out/${tag}_lat_${name}.json 2> gpurun_out/${tag}_lat_${name}.err
  python - <<P
import json
d=json.load(open("gpurun_out/${tag}_lat_${name}.json"))
for k,v in d.items(): print("  lat", k, v["ms_median"], "dedup", v["stage_ms_per_call"]["dedup"], "kp", v["keypoints"])
P
  env $v timeout 600 python bench.py --legs extract,extract_4k --steps 3 --warmup 3 --no-cpu 2> gpurun_out/${tag}_bench_${name}.err > gpurun_out/${tag}_bench_${name}.json
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${name}.json").read().strip().splitlines()[-1])
    print("  bench 1080p", round(d["value"],1), round(d["e2e"]["value"],1), " 4k", round(d["extract_4k"]["value"],1), d.get("parity_check"))
except Exception as e:
    print("  bench failed", e)
P
done
