#!/bin/bash
# Round-end validation on the GPU box: GPU suite, smoke, headline bench, C1-C5, stem bench, ncu of the stem.
# Everything lands in gpurun_out/.
timeout -s KILL 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_gpu.log
timeout -s KILL 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/smoke.log
timeout -s KILL 300 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo bench rc=$?
timeout -s KILL 300 python tools/bench_configs.py --out gpurun_out/r1_configs.json > gpurun_out/configs.log 2>&1; echo configs rc=$?
timeout -s KILL 100 python tools/bench_stem.py --json gpurun_out/stem_bench.json > gpurun_out/stem_bench.log 2>&1; echo stembench rc=$?
if [ -f build/libbhstem_prof.so ]; then BHSTEM_LIB=$PWD/build/libbhstem_prof.so timeout -s KILL 100 python tools/stem_roles.py 16 > gpurun_out/stem_roles.log 2>&1; fi
timeout -s KILL 60 python tools/run_stem_once.py 16 10 > gpurun_out/stem_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:bhstem -s 4 -c 2 -o gpurun_out/prof_stem_r1b \
      python tools/run_stem_once.py 16 4 > gpurun_out/ncu_stem.log 2>&1
echo ncu rc=$?
