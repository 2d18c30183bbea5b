mkdir -p gpurun_out
MS=104 XT_EIG_GRID=148 timeout 200 ncu --set full --import-source on --clock-control none -k regex:small_eigh -c 1 -o gpurun_out/eigh104g -f python tools/check_eigh.py > gpurun_out/E_ncu.log 2>&1
tail -3 gpurun_out/E_ncu.log
