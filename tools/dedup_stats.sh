#!/bin/bash
# builds the instrumented cache pass (-DAKZ_DEDUP_STATS: per-level step counts and cycle split, printed for image 0) into
# akaze-rust_b200/build/libakaze_dstat.so; on the GPU box: cp it over libakaze_b200.so and run tools/latency.py --reps 1
set -e
cd "$(dirname "$0")/../akaze-rust_b200/csrc"
mkdir -p /tmp/dstat
for f in akaze_api scale_space detector keypoints matcher matcher_tc ransac; do
  extra=""; [ $f = keypoints ] && extra="-DAKZ_DEDUP_STATS"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC $extra -c $f.cu -o /tmp/dstat/$f.o 2>/dev/null &
done
wait
nvcc -shared -o ../build/libakaze_dstat.so /tmp/dstat/*.o -ldl 2>/dev/null
ls -la ../build/libakaze_dstat.so
