#!/bin/bash
# A/B helper for the conv stem on the GPU box: tests, bench table, role waits -> gpurun_out/$1
out=gpurun_out/${1:-stem_ab.log}
{
  timeout -s KILL 200 python -m pytest tests/test_gpu_stem.py -m gpu -q 2>&1 | tail -3
  timeout -s KILL 100 python tools/bench_stem.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['batch'], round(r['ours_ms'], 4), round(r['conv1_tflops']), round(r['conv2_tflops']), round(r['ours_tflops']), round(r['speedup_vs_torch'], 2))
"
  if [ -f build/libbhstem_prof.so ]; then BHSTEM_LIB=$PWD/build/libbhstem_prof.so timeout -s KILL 100 python tools/stem_roles.py 16 2>&1; fi
} > "$out" 2>&1
tail -5 "$out"
