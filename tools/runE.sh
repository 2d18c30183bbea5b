mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:expand_fused -s 9 -c 1 -o gpurun_out/expand10 -f python tools/trace_c2.py > gpurun_out/E_ncu.log 2>&1
tail -3 gpurun_out/E_ncu.log
