#!/bin/bash
# latency + bench A/B of environment switches on one GPU:  tools/r2_lat_ab.sh TAG "ENV=1" "ENV2=1 ENV3=1" ...
tag=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -2 gpurun_out/${tag}_pytest.log
for v in "$@"; do
  name=$(echo "$v" | tr ' =' '__')
  echo "== $v"
  env $v timeout 300 python tools/latency.py --sizes 1080x1920,2160x3840,600x800 --reps 30 > gpurun_out/${tag}_lat_${name}.json 2> gpurun_out/${tag}_lat_${name}.err
  python - <<P
import json
d=json.load(open("gpurun_out/${tag}_lat_${name}.json"))
for k,v in d.items(): print("  lat", k, v["ms_median"], v["ms_min"], "kp", v["keypoints"])
P
  env $v timeout 600 python bench.py --legs extract,extract_4k --steps 3 --warmup 3 --no-cpu 2> gpurun_out/${tag}_bench_${name}.err > gpurun_out/${tag}_bench_${name}.json
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${name}.json").read().strip().splitlines()[-1])
    print("  bench 1080p", round(d["value"],1), round(d["e2e"]["value"],1), " 4k", round(d["extract_4k"]["value"],1), round(d["extract_4k"]["e2e"]["value"],1))
except Exception as e:
    print("  bench failed", e)
P
done
