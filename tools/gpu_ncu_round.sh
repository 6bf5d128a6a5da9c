#!/bin/bash
# gpurun --timeout 2400 -- 'bash tools/gpu_ncu_round.sh <tag>' : ncu evidence of the bench command (1 GPU)
#   launch lists (gpu__time_duration, plain stream launches: MICROPP_GRAPHS=0) of small batches of elastic30 / damage50,
#   one `ncu --set full` capture of the DPCG kernels of elastic30 at the bench size (1024 RVEs) and of damage50 (64 RVEs)
tag=${1:-n}
O=gpurun_out
mkdir -p $O
MICROPP_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file $O/${tag}_launches_elastic30_ngp64.csv \
    python bench.py --ngp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-assembled --no-extra > $O/${tag}_ncu_launch_e.log 2>&1
MICROPP_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file $O/${tag}_launches_damage50_ngp8.csv \
    python bench.py --workload damage50 --ngp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $O/${tag}_ncu_launch_d.log 2>&1
MICROPP_GRAPHS=0 timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"k_spmv_dot_tmac|k_spmv_fix|k_cg_update_imp|k_cg_pupdate_imp" -s 40 -c 8 -f -o $O/${tag}_ncu_dpcg_elastic30_ngp1024 \
    python bench.py --ngp 1024 --steps 1 --warmup 1 --no-cpu-baseline --no-assembled --no-extra > $O/${tag}_ncu_full_e.log 2>&1
MICROPP_GRAPHS=0 timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"k_spmv_dot|k_elem_ctan|k_asm_mat_general" -s 3000 -c 6 -f -o $O/${tag}_ncu_damage50_ngp64 \
    python bench.py --workload damage50 --ngp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > $O/${tag}_ncu_full_d.log 2>&1
for f in ${tag}_ncu_dpcg_elastic30_ngp1024 ${tag}_ncu_damage50_ngp64; do
  ncu -i $O/$f.ncu-rep --page details > $O/${f}_details.txt 2>&1
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>&1
done
ls -la $O | grep ${tag}_
