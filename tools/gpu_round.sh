#!/bin/bash
# One GPU-box visit: parity tests, both bench workloads, the ncu launch list and the full ncu capture.
# usage: tools/gpu_round.sh <tag> [what...]   what in: tests extract match launches full   (default: all)
# gpurun copies back at most 64 MiB of gpurun_out/, so the big .ncu-rep stays on the box and only its CSV export returns.
tag=${1:-x}; shift
what=${*:-tests extract match launches full}
mkdir -p gpurun_out
has() { [[ " $what " == *" $1 "* ]]; }
if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
  tail -5 gpurun_out/${tag}_pytest.log
fi
if has extract; then
  timeout 600 python bench.py > gpurun_out/${tag}_bench_extract.json 2> gpurun_out/${tag}_bench_extract.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_extract.json").read().strip().splitlines()[-1])
    print("extract: value %.1f e2e %.1f img/s" % (d["value"], d["e2e"]["value"]), {k: round(v["ms_per_image"], 4) for k, v in d["stages"].items()}, d["cpu_baseline"] and d["cpu_baseline"]["value"])
except Exception as e:
    print("extract bench failed", e); print(open("gpurun_out/${tag}_bench_extract.err").read()[-2000:])
PY
fi
if has match; then
  timeout 600 python bench.py --workload match > gpurun_out/${tag}_bench_match.json 2> gpurun_out/${tag}_bench_match.err; tail -c 1200 gpurun_out/${tag}_bench_match.json; tail -5 gpurun_out/${tag}_bench_match.err
fi
# the profiling runs reuse the synthetic images cached by tools/stage_ab.py (image synthesis takes ~10 s per image)
if has launches || has full; then python tools/stage_ab.py --images 8 "" > /dev/null 2>&1; fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --images 256 --unique 4 --no-e2e --no-cpu > gpurun_out/${tag}_launches.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_match --csv --log-file gpurun_out/${tag}_launches_match.csv python tools/profile_run.py --no-extract --match 65536 >> gpurun_out/${tag}_launches.log 2>&1
fi
if has full; then
  timeout 900 ncu --set full --clock-control none -k regex:k_ -c 400 -o /tmp/${tag}_full python tools/profile_run.py --images 64 --match 32768 > gpurun_out/${tag}_full.log 2>&1
  ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>> gpurun_out/${tag}_full.log
fi
du -sh gpurun_out
