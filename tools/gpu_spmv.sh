#!/bin/bash
# gpurun --timeout 1200 -- 'bash tools/gpu_spmv.sh <tag>' : implicit-operator parity tests + isolated SpMV timings
tag=${1:-s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_implicit.py -m gpu -x -q > gpurun_out/${tag}_imp_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/${tag}_imp_pytest.log
timeout 300 python tools/bench_imp_spmv.py 30 1024 20 > gpurun_out/${tag}_spmv30.log 2>&1; grep "^kernel" gpurun_out/${tag}_spmv30.log
KERNELS=3 timeout 300 python tools/bench_imp_spmv.py 30 1024 20 homog > gpurun_out/${tag}_spmv30_homog.log 2>&1; grep "^kernel" gpurun_out/${tag}_spmv30_homog.log
KERNELS=3 timeout 300 python tools/bench_imp_spmv.py 50 256 10 > gpurun_out/${tag}_spmv50.log 2>&1; grep "^kernel" gpurun_out/${tag}_spmv50.log
