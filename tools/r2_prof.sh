#!/bin/bash
# round 2 profiling visit: full-load ncu (--set full) of every kernel of one 64-image 1080p extraction, and of the
# tensor-core matcher at the bench size (1M x 1M). usage: tools/r2_prof.sh <tag> [extract] [match] [src]
tag=${1:-r2p}; shift
what=${*:-extract match}
mkdir -p gpurun_out
has() { [[ " $what " == *" $1 "* ]]; }
if has extract; then
  AKZ_NO_RAMP=1 AKZ_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none -k regex:k_ -c 140 -o /tmp/${tag}_full python tools/profile_run.py --images 64 --unique 8 > gpurun_out/${tag}_full.log 2>&1
  echo "ncu extract exit $?"
  ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>> gpurun_out/${tag}_full.log
  python tools/ncu_rows.py gpurun_out/${tag}_full_raw.csv > gpurun_out/${tag}_ncu_fullload.txt; head -3 gpurun_out/${tag}_ncu_fullload.txt | cut -c1-250
  python tools/stage_traffic.py gpurun_out/${tag}_full_raw.csv 64 > gpurun_out/${tag}_stage_traffic.json
fi
if has match; then
  timeout 900 ncu --set full --clock-control none -k regex:k_match_tc -c 1 -o /tmp/${tag}_match python tools/profile_run.py --no-extract --match 1048576 > gpurun_out/${tag}_match.log 2>&1
  echo "ncu match exit $?"
  ncu -i /tmp/${tag}_match.ncu-rep --page raw --csv > gpurun_out/${tag}_match_raw.csv 2>> gpurun_out/${tag}_match.log
  python tools/ncu_rows.py gpurun_out/${tag}_match_raw.csv > gpurun_out/${tag}_ncu_match.txt; cat gpurun_out/${tag}_ncu_match.txt | cut -c1-400
fi
if has src; then
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"${SRC_KERNELS:-k_prep_stream|k_fed_pp|k_detector_tmem|k_descriptor}" -c ${SRC_COUNT:-8} -o gpurun_out/${tag}_src python tools/profile_run.py --images 64 --unique 8 > gpurun_out/${tag}_src.log 2>&1
  echo "ncu src exit $?"
fi
du -sh gpurun_out
