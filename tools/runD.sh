mkdir -p gpurun_out
for v in v1 v2 v3; do echo "== variant $v"; MS=32,33,48,104 XT_LIB_OVERRIDE=$PWD/tools/variants/lib_$v.so timeout 120 python tools/check_eigh.py 2>&1 | grep -E "mode=0|rror"; done > gpurun_out/D_check.log 2>&1
cat gpurun_out/D_check.log
