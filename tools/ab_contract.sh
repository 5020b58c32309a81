#!/bin/bash
# A/B of the FP64 contraction (pair contraction 512 x 512 x 1024 envs): N tile width and split-K factor
for v in "X=0" "RDDL_CONTRACT_BN=64" "RDDL_CONTRACT_SPLITK=2" "RDDL_CONTRACT_BN=64 RDDL_CONTRACT_SPLITK=2" "RDDL_CONTRACT_BN=64 RDDL_CONTRACT_SPLITK=3"; do
  env $v python bench.py --workload pair512 --envs 1024 --steps 5 --warmup 3 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d['roofline']['tensor']
print('$v', 'f64 kernel_us', round(t['fp64_dmma']['kernel_us'],1), 'transition(graph) us', {k: round(x,1) for k,x in t['transition_us_graph_replay'].items()}, 'value', int(d['value']))"
done
