mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:k_ -c 70 -o /tmp/r1z_full python tools/profile_run.py --images 64 > gpurun_out/r1z_full.log 2>&1
ncu -i /tmp/r1z_full.ncu-rep --page raw --csv > gpurun_out/r1z_full64_raw.csv 2>> gpurun_out/r1z_full.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_detector_stream -c 2 -o gpurun_out/r1z_src python tools/profile_run.py --images 64 > gpurun_out/r1z_src.log 2>&1
ls -la gpurun_out
