#!/bin/bash
# A/B helper: bench line per build (trajectory + e2e), search stats for a stats build if present
#   bash scripts/r2_ab.sh <tag> build/variants/lib_x.so ...
out=gpurun_out; mkdir -p $out; tag=$1; shift
for lib in visma_b200/libvisma_b200.so "$@"; do
  VISMA_B200_LIB=$PWD/$lib timeout 150 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra 2> $out/${tag}_err.txt | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=j['config']
print('$lib', 'value %.0f step %.4f pass %.4f solve %.4f' % (j['value'], j['ms_per_step'], c['pass_ms'], c['solve_ms']), 'first6 %.3f settled %.4f' % (sum(c['trajectory_ms_per_step'][:6]), c['settled_regime']['ms_per_step']), 'e2e', j['e2e']['note'].split(';')[1][:14], 'default %.2f' % j['e2e_default_criteria']['ms_per_call_32_objects'], 'with_build %.1f' % j['e2e_with_scene_build']['ms_per_call'])
print('   ', c['trajectory_ms_per_step'])
"
done 2>&1 | tee $out/${tag}_ab.txt
if [ -f build/variants/lib_stats.so ]; then VISMA_B200_LIB=$PWD/build/variants/lib_stats.so timeout 200 python scripts/search_stats.py 8 2>&1 | grep -v "^$" | cut -c1-700 > $out/${tag}_stats.txt; fi
