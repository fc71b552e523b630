# full-load ncu of the keypoint-stage kernels and the (new) stencil kernels: 64 x 1080p, one launch each
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_descriptor|k_orientation|k_filter_refine" -c 3 -o gpurun_out/${1:-r2a}_kp python tools/profile_run.py --images 64 > gpurun_out/${1:-r2a}_kp.log 2>&1
ncu -i gpurun_out/${1:-r2a}_kp.ncu-rep --page raw --csv > gpurun_out/${1:-r2a}_kp_raw.csv
ls -la gpurun_out | tail -5
