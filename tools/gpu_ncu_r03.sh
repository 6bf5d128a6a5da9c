#!/bin/bash
# gpurun --timeout 2400 -- 'bash tools/gpu_ncu_r03.sh <tag>' : ncu evidence for the RVEs with a damage / plastic phase
#   (hybrid operator + restructured k_elem_ctan), 1 GPU, plain stream launches (MICROPP_GRAPHS=0)
tag=${1:-n}
O=gpurun_out
mkdir -p $O
for wl in damage50 plastic40; do
MICROPP_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
    --log-file $O/${tag}_launches_${wl}_ngp8.csv \
    python bench.py --workload $wl --ngp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $O/${tag}_ncu_launch_$wl.log 2>&1
done
# DRAM traffic of the hybrid operator's kernels, 2 RVEs of plastic40 (every launch holds 1 or 2 active RVEs)
MICROPP_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"k_spmv_hyb|k_spmv_dot_tmac|k_spmv_fix" -s 6000 -c 600 --csv --log-file $O/${tag}_dram_hybrid_plastic40_ngp2.csv \
    python bench.py --workload plastic40 --ngp 2 --steps 1 --warmup 0 --no-cpu-baseline --no-extra > $O/${tag}_ncu_dram_p.log 2>&1
MICROPP_GRAPHS=0 timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"k_spmv_hyb|k_spmv_dot_tmac|k_spmv_fix|k_elem_ctan|k_asm_mat_general|k_probe_lin|k_hyb_list" -s 2500 -c 14 -f -o $O/${tag}_ncu_plastic40_ngp64 \
    python bench.py --workload plastic40 --ngp 64 --steps 1 --warmup 0 --no-cpu-baseline --no-extra > $O/${tag}_ncu_full_p.log 2>&1
for f in ${tag}_ncu_plastic40_ngp64; do
  ncu -i $O/$f.ncu-rep --page details > $O/${f}_details.txt 2>&1
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>&1
done
tail -2 $O/${tag}_ncu_launch_*.log $O/${tag}_ncu_dram_p.log $O/${tag}_ncu_full_p.log | cut -c1-300
ls -la $O | grep ${tag}_
