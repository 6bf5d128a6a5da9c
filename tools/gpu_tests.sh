#!/bin/bash
# gpurun --timeout 2400 -- 'bash tools/gpu_tests.sh <tag> [pytest args]'   : the -m gpu suite, log under gpurun_out/
tag=${1:-t}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 2200 python -m pytest tests -m gpu -q -s -rA --durations=15 "$@" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|error|FAILED|ERROR|worst|slab200:" gpurun_out/${tag}_pytest.log | tail -40
