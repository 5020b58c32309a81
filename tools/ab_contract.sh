#!/bin/bash
# A/B of the FP64 contraction's split-K factor (pair contraction 512 x 512; 0 = the library's choice)
for cfg in "1024 0" "128 1" "128 4" "128 8" "128 0" "256 0" "512 0"; do
  set -- $cfg
  if [ "$2" = "0" ]; then unset RDDL_CONTRACT_SPLITK; else export RDDL_CONTRACT_SPLITK=$2; fi
  python bench.py --workload pair512 --envs $1 --steps 5 --warmup 3 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d['roofline']['tensor']
print('envs $1 splitk $2', 'f64 kernel_us', round(t['fp64_dmma']['kernel_us'],1), 'transition(graph) us', {k: round(x,1) for k,x in t['transition_us_graph_replay'].items()}, 'value', int(d['value']))"
done
