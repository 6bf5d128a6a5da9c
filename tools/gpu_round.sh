#!/bin/bash
# One gpurun call: GPU parity tests, the default bench (both arms), a damage50 bench, the ncu launch list of the
# bench command and one full ncu capture of the dominant kernel.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r01c'
tag=${1:-run}
what=${2:-all}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/${tag}_gpu.txt 2>&1

if [[ $what == all || $what == *tests* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> $O/${tag}_pytest.log
  tail -3 $O/${tag}_pytest.log
fi
if [[ $what == all || $what == *bench* ]]; then
  timeout 600 python bench.py > $O/${tag}_bench_elastic30.json 2> $O/${tag}_bench_elastic30.err
  tail -c 600 $O/${tag}_bench_elastic30.json
  timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $O/${tag}_bench_ref_elastic30.json 2> $O/${tag}_bench_ref.err
  timeout 900 python bench.py --workload damage50 --ngp 128 --steps 2 --warmup 3 --cpu-sample 8 > $O/${tag}_bench_damage50_ngp128.json 2> $O/${tag}_bench_damage50.err
  tail -c 600 $O/${tag}_bench_damage50_ngp128.json
fi
if [[ $what == all || $what == *ncu* ]]; then
  # launch list of the bench command itself (small batch so that ncu's serialisation stays short)
  MICROPP_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
      --log-file $O/${tag}_launches_elastic30_ngp64.csv \
      python bench.py --ngp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-assembled > $O/${tag}_ncu_launch.log 2>&1
  MICROPP_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_dot -s 40 -c 2 \
      -f -o $O/${tag}_ncu_spmv_imp_elastic30_ngp64 \
      python bench.py --ngp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-assembled > $O/${tag}_ncu_full.log 2>&1
  MICROPP_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
      --log-file $O/${tag}_launches_damage50_ngp8.csv \
      python bench.py --workload damage50 --ngp 8 --steps 1 --warmup 1 --no-cpu-baseline > $O/${tag}_ncu_launch_dmg.log 2>&1
fi
ls -la $O | tail -20
