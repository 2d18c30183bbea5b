mkdir -p gpurun_out
MS=104 XT_EIG_DEBUG=8 timeout 100 python tools/check_eigh.py > gpurun_out/D_check.log 2>&1
grep -E "xt-eig|mode=0" gpurun_out/D_check.log
timeout 100 python tests/gpu_eigh_phases.py 2>&1 | grep -v phases
