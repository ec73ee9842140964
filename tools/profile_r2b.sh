#!/bin/bash
# Round-2 profile artefacts, second batch (tensor-core VJP, fused K=1 stack kernel, launch list of the final bench command).
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cheb_backward_f16 -s 2 -c 1 -o $O/r2_bwd_f16 python tools/bwd_once.py 4 > $O/r2_ncu_bwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cheb_mlp_f16 -s 2 -c 1 -o $O/r2_mlp_f16 python tools/stack_once.py 4 > $O/r2_ncu_mlp.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:grads_sum_fused -s 2 -c 1 -o $O/r2_grads_sum python tools/bwd_once.py 4 > $O/r2_ncu_sum.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2_launches_bench_b.csv python bench.py --steps 20 --warmup 5 --no-cpu --sweep off --replays 3 > $O/r2_launches_bench_b.out 2>&1
ls -la $O | tail -8
