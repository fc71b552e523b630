mkdir -p gpurun_out
python tools/stage_ab.py --images 8 "" > /dev/null 2>&1
timeout 600 python bench.py > gpurun_out/r1zz_bench_extract.json 2> gpurun_out/r1zz_bench_extract.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r1zz_bench_extract.json").read().strip().splitlines()[-1])
print("extract: value %.1f e2e %.1f img/s" % (d["value"], d["e2e"]["value"]), {k: round(v["ms_per_image"], 4) for k, v in d["stages"].items()}, d["cpu_baseline"] and d["cpu_baseline"]["value"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/r1zz_launches.csv python bench.py --steps 1 --warmup 1 --images 256 --unique 4 --no-e2e --no-cpu > gpurun_out/r1zz_launches.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"k_detector_tmem|k_fed_pp|k_prep_stream|k_contrast|k_flow_ew|k_level0" -c 150 -o /tmp/r1zz_stencil python tools/profile_run.py --images 64 > gpurun_out/r1zz_full.log 2>&1
ncu -i /tmp/r1zz_stencil.ncu-rep --page raw --csv > gpurun_out/r1zz_stencil_raw.csv 2>> gpurun_out/r1zz_full.log
ls -la gpurun_out | grep r1zz
