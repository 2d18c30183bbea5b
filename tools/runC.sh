mkdir -p gpurun_out
./tools/micro/lat > gpurun_out/C_lat.log 2>&1
timeout 120 python tests/gpu_eigh_phases.py > gpurun_out/C_eigh_phases.log 2>&1
cat gpurun_out/C_lat.log gpurun_out/C_eigh_phases.log
