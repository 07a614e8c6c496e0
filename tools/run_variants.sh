#!/bin/bash
# usage: run_variants.sh tag variant...   (on the GPU box; "base" = the in-tree library, others = build/<variant>/)
# "name+t" also runs the GPU parity module for that variant; every variant gets a 12 000-pair bench (GCUPS and phase times)
tag=$1; shift
for spec in "$@"; do
  v=${spec%+t}
  if [ "$v" = base ]; then lib=cpecan_b200/lib/libcpecan_b200.so; else lib=build/$v/libcpecan_b200.so; fi
  echo "== $v" >> gpurun_out/${tag}.log
  if [ "$spec" != "$v" ]; then
    CPB_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/${tag}.log
  fi
  CPB_LIB=$PWD/$lib timeout 300 python bench.py --pairs 12000 --skip-e2e --skip-cpu --steps 3 --warmup 2 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        j=json.loads(l); print('$v', round(j['value'],3), {k: round(x,2) for k,x in j['phase_ms_per_step'].items()})
" >> gpurun_out/${tag}.log
done
cat gpurun_out/${tag}.log
