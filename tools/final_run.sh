#!/bin/bash
# Round-end measurement pass on one B200: GPU tests, the bench line, the reference arm, the ncu launch list of the
# bench command and `--set full` captures of the hot kernels (outputs under gpurun_out/final_*)
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/final_gputest.log 2>&1; echo "pytest rc=$?" >> $O/final_gputest.log
tail -n 3 $O/final_gputest.log
python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_reference.json 2> $O/final_bench_reference.err; echo "reference rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/final_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-packed > $O/final_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:rddl_plane_static -s 2 -c 1 -f -o $O/final_plane1024 python tools/prof_run.py 1024 4096 4 > $O/final_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:rddl_level_fused -s 2 -c 1 -f -o $O/final_resv python tools/prof_generic.py reservoir 1000 4096 4 > $O/final_ncu2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:rddl_contract_f64 -s 2 -c 1 -f -o $O/final_ct python tools/prof_generic.py pair 512 1024 4 > $O/final_ncu3.log 2>&1
ncu --set full --clock-control none -k regex:rddl_plane_static -s 2 -c 1 -f -o $O/final_policy1024 python tools/prof_policy.py 1024 4096 3 > $O/final_ncu4.log 2>&1
ls -la $O | tail -n 20
