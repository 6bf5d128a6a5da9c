for wl in damage50 plastic40; do timeout 600 python bench.py --workload $wl --ngp 128 --steps 2 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; print('$wl', d['value'], d['ms_per_step'], r['other_kernels_ms'], r['kernel_ms'], r['instrumented_step_ms'])"; done
