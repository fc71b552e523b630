#!/bin/bash
# round 2, first GPU visit: full GPU test suite, smoke, the default bench (1080p + 4K + match), launch list
tag=${1:-r2a}
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt; free -g >> gpurun_out/${tag}_host.txt; nvidia-smi -L >> gpurun_out/${tag}_host.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
tail -3 gpurun_out/${tag}_smoke.log
( time timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2> gpurun_out/${tag}_bench.time
echo "bench exit $?"; tail -3 gpurun_out/${tag}_bench.time; tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("extract: value %.1f e2e %.1f img/s" % (d["value"], d["e2e"]["value"]), {k: round(v["ms_per_image"], 4) for k, v in d["stages"].items()})
    print("cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "parity", d["parity_check"])
    x = d["extract_4k"]; print("4k: value %.1f e2e %.1f" % (x["value"], x["e2e"]["value"]), {k: round(v["ms_per_image"], 4) for k, v in x["stages"].items()}, x["roofline_pipeline"]["frac"], x["run"])
    m = d["match"]; print("match: %.3e pairs/s e2e %.3e" % (m["value"], m["e2e"]["value"]), m["roofline"]["frac"], m["parity_check"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --images 256 --unique 4 --legs extract --no-e2e --no-cpu > gpurun_out/${tag}_launches.log 2>&1
echo "launches exit $?"
du -sh gpurun_out
