#!/bin/bash
# One GPU-box visit: parity tests, both bench workloads, the ncu launch list and ncu captures.
# usage: tools/gpu_round.sh <tag> [what...]   what in: tests extract match launches full src   (default: all)
# gpurun copies back at most 64 MiB of gpurun_out/, so .ncu-rep files are converted to CSV on the box and
# only small reports are kept.
tag=${1:-x}; shift
what=${*:-tests extract match launches full src srcmatch}
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
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_run.py --images 32 --match 65536 > gpurun_out/${tag}_launches.log 2>&1
fi
if has full; then
  timeout 900 ncu --set full --clock-control none -k regex:k_ -c 260 -o /tmp/${tag}_full python tools/profile_run.py --images 4 --match 32768 > gpurun_out/${tag}_full.log 2>&1
  ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>> gpurun_out/${tag}_full.log
fi
if has src; then
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"${NCU_SRC_REGEX:-k_detector_stream}" -c ${NCU_SRC_COUNT:-4} -o gpurun_out/${tag}_src python tools/profile_run.py --images ${NCU_SRC_IMAGES:-8} > gpurun_out/${tag}_src.log 2>&1
  ls -la gpurun_out/${tag}_src.ncu-rep
  if [ $(stat -c %s gpurun_out/${tag}_src.ncu-rep 2>/dev/null || echo 0) -gt 30000000 ]; then rm -f gpurun_out/${tag}_src.ncu-rep; fi
fi
if has srcmatch; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_match_tc -c 1 -o gpurun_out/${tag}_srcmatch python tools/profile_run.py --no-extract --match 131072 > gpurun_out/${tag}_srcmatch.log 2>&1
  ls -la gpurun_out/${tag}_srcmatch.ncu-rep
fi
du -sh gpurun_out
