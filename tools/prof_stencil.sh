# full-load ncu (64 x 1080p) of the first octave's streaming kernels, with source correlation
mkdir -p gpurun_out
tag=${1:-r2d}
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_detector_stream|k_fed_pp|k_prep_stream" -c ${2:-11} -o gpurun_out/${tag}_stencil python tools/profile_run.py --images 64 > gpurun_out/${tag}_stencil.log 2>&1
ncu -i gpurun_out/${tag}_stencil.ncu-rep --page raw --csv > gpurun_out/${tag}_stencil_raw.csv
ls -la gpurun_out | tail -4
