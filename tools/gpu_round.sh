#!/bin/bash
# One gpurun call: the -m gpu suite, smoke(), the default bench (both arms).  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r02m'
tag=${1:-run}
mkdir -p gpurun_out
bash tools/gpu_tests.sh $tag
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/${tag}_smoke.log
bash tools/gpu_bench.sh $tag
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; tail -c 400 gpurun_out/${tag}_bench_ref.json
