#!/bin/bash
# Round-2 profile artefacts (run on the GPU box through gpurun; everything lands in gpurun_out/).
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
# 1. the bench line itself (no profiler)
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_line.json 2> $O/r2_bench_line.err
# 2. every launch of the bench command with its device time (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu --sweep off --replays 3 > $O/r2_launches_bench.out 2>&1
# 3. full capture of the headline kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cheb_f16_kernel -s 2 -c 1 -o $O/r2_f16_k5 python tools/probe_once.py > $O/r2_ncu_f16.log 2>&1
# 4. full capture of one K = 10 / n = 512 sweep point (CSR-walk kernel)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cheb_forward_kernel -s 1 -c 1 -o $O/r2_walk_k10_n512 python tools/sweep_point.py 512 10 4096 2 > $O/r2_ncu_walk.log 2>&1
# 5. error table of the deep K > 1 stacks (first-generation dense kernel) against numpy fp32
timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -q -s -k deep_k_stacks 2>&1 | grep "^K " > $O/r2_deep_stack_errors.txt
# 6. clock marks of one CTA (instrumented build)
MHO_LIB=$PWD/multihop_offload_b200/libmho_probe.so timeout 100 python tools/probe_once.py > $O/r2_probe_f16.txt 2>&1
ls -la $O | tail -12
