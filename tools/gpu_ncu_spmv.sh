#!/bin/bash
# gpurun --timeout 1500 -- 'bash tools/gpu_ncu_spmv.sh <tag> <kernel regex> <case> [n] [ngp]' : ncu --set full of the SpMV micro-bench
tag=${1:-n}; rx=${2:-k_spmv_dot_roll}; cs=${3:-homog}; n=${4:-30}; ngp=${5:-256}
mkdir -p gpurun_out
KERNELS=4 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 2 -f -o gpurun_out/${tag}_ncu \
   python tools/bench_imp_spmv.py $n $ngp 2 $cs > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
ncu -i gpurun_out/${tag}_ncu.ncu-rep --page details > gpurun_out/${tag}_ncu_details.txt 2>&1
grep -E "Duration|Registers Per|Theoretical Occ|Achieved Occ|FP64|Issue Slots Busy|No Eligible|Stall|L1/TEX Hit|Executed Ipc|pipe" gpurun_out/${tag}_ncu_details.txt | head -40
