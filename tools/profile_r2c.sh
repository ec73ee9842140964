#!/bin/bash
# Round-2 profile artefacts, third batch: the warp-specialised forward (headline kernel at the end of the round).
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cheb_f16ws -s 2 -c 1 -o $O/r2_f16ws_k5 python tools/probe_once.py > $O/r2_ncu_f16ws.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2_launches_bench_c.csv python bench.py --steps 20 --warmup 5 --no-cpu --sweep off --replays 3 > $O/r2_launches_bench_c.out 2>&1
MHO_LIB=$PWD/multihop_offload_b200/libmho_probe.so timeout 100 python tools/probe_once.py > $O/r2_probe_f16ws.txt 2>&1
timeout 60 ./tools/umma_probe3 > $O/r2_umma_probe3.txt 2>&1
ls -la $O | tail -6
