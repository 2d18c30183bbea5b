mkdir -p gpurun_out
MS=104 timeout 300 ncu --set full --import-source on --clock-control none -k regex:small_eigh -c 2 -o gpurun_out/eigh104 -f python tools/check_eigh.py > gpurun_out/E_ncu.log 2>&1
ls -la gpurun_out/ | head; tail -5 gpurun_out/E_ncu.log
