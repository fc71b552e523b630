#!/bin/bash
# round 2 GPU visit: usage tools/r2_round.sh <tag> [what...]   what in: tests smoke bench bench2 launches ab
tag=${1:-r2x}; shift
what=${*:-tests smoke bench}
mkdir -p gpurun_out
has() { [[ " $what " == *" $1 "* ]]; }
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
  tail -12 gpurun_out/${tag}_pytest.log
fi
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
  tail -3 gpurun_out/${tag}_smoke.log
fi
summ() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("extract: value %.1f e2e %.1f img/s n_gpus %d" % (d["value"], d["e2e"]["value"], d["n_gpus"]), {k: round(v["ms_per_image"], 4) for k, v in d["stages"].items()})
    print("  cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 3), "parity", d["parity_check"] and (d["parity_check"]["ok"], d["parity_check"]["bit_exact_keypoints"], d["parity_check"]["descriptor_bits_differing"]))
    if "single_image" in d:
        print("  single image: %.3f ms median, %.3f min, %d keypoints, %d launches" % (d["single_image"]["ms_median"], d["single_image"]["ms_min"], d["single_image"]["keypoints"], d["single_image"]["gpu_launches_per_call"]))
    if "extract_4k" in d:
        x = d["extract_4k"]; print("  4k: value %.1f e2e %.1f frac %.3f" % (x["value"], x["e2e"]["value"], x["roofline_pipeline"]["frac"]), {k: round(v["ms_per_image"], 4) for k, v in x["stages"].items()})
    if "match" in d:
        m = d["match"]; print("  match: %.3e pairs/s frac %.3f ranks_agree %s parity %s" % (m["value"], m["roofline"]["frac"], m["ranks_agree"], m["parity_check"] and m["parity_check"]["ok"]))
except Exception as e:
    print("bench parse failed", e)
PY
}
if has bench; then
  ( time timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2> gpurun_out/${tag}_bench.time
  echo "bench exit $?"; tail -3 gpurun_out/${tag}_bench.time; tail -5 gpurun_out/${tag}_bench.err; summ gpurun_out/${tag}_bench.json
fi
if has bench2; then
  N=$(nvidia-smi -L | wc -l)
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err ) 2> gpurun_out/${tag}_bench_${N}gpu.time
  echo "bench $N gpu exit $?"; tail -3 gpurun_out/${tag}_bench_${N}gpu.time; tail -8 gpurun_out/${tag}_bench_${N}gpu.err; summ gpurun_out/${tag}_bench_${N}gpu.json
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --images 256 --unique 4 --legs extract --no-e2e --no-cpu > gpurun_out/${tag}_launches.log 2>&1
  echo "launches exit $?"
fi
if has ab; then
  python tools/stage_ab.py --images 256 -- $AB_VARIANTS > gpurun_out/${tag}_ab.log 2>&1; cat gpurun_out/${tag}_ab.log | tail -30
fi
du -sh gpurun_out
