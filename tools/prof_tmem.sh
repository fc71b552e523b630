mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_detector_tmem" -c 4 -o gpurun_out/r2f_tmem python tools/profile_run.py --images 64 > gpurun_out/r2f_tmem.log 2>&1
ncu -i gpurun_out/r2f_tmem.ncu-rep --page raw --csv > gpurun_out/r2f_tmem_raw.csv
ls -la gpurun_out | tail -3
