#!/bin/bash
for v in "X=0" "RDDL_PLANE_MINB=4" "RDDL_PLANE_MINB=2" "RDDL_PLANE_WPT=2" "RDDL_PLANE_WPT=2 RDDL_PLANE_MINB=4" "RDDL_PLANE_WPT=2 RDDL_PLANE_MINB=5" "RDDL_PLANE_WPT=2 RDDL_PLANE_MINB=6"; do
  env $v RDDL_B200_JIT_CACHE=/tmp/jit_$RANDOM python bench.py --workload wildfire1024 --envs 512 --steps 5 --warmup 3 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', 'kernel_us', round(d['roofline']['kernel_us'],1), 'value', int(d['value']), 'e2e', int(d['e2e']['value']), 'e2e dev us', round(d['e2e']['device_us_per_rollout']/4,1))"
done
