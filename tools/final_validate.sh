#!/bin/bash
# Round-end validation on the GPU box: GPU suite, smoke, headline bench, C1-C5, per-set and stem benches.
# Everything lands in gpurun_out/.
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_gpu.log
timeout -s KILL 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/smoke.log
timeout -s KILL 300 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo bench rc=$?
timeout -s KILL 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo reference rc=$?
timeout -s KILL 300 python tools/bench_configs.py --out gpurun_out/r2_configs.json > gpurun_out/configs.log 2>&1; echo configs rc=$?
timeout -s KILL 100 python tools/pset_bench.py > gpurun_out/r2_psets.txt 2>&1; echo psets rc=$?
timeout -s KILL 100 python tools/bench_stem.py --json gpurun_out/r2_stem_bench.json > gpurun_out/stem_bench.log 2>&1; echo stembench rc=$?
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:bhstem -s 9 -c 3 -o gpurun_out/prof_r2_stem_split_final -f python tools/run_stem_split_once.py 46 6 > gpurun_out/ncu_stem_split_final.log 2>&1; echo ncu-split rc=$?
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:bhstem -s 6 -c 2 -o gpurun_out/prof_r2_stem_final -f python tools/run_stem_once.py 46 6 > gpurun_out/ncu_stem_final.log 2>&1; echo ncu-stem rc=$?
