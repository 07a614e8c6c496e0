#!/bin/bash
# usage: run_variants.sh tag variant...   (on the GPU box; "base" = the in-tree library, others = build/<variant>/)
# each variant: the GPU parity module, then a 12 000-pair bench (GCUPS and phase times)
tag=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then lib=cpecan_b200/lib/libcpecan_b200.so; else lib=build/$v/libcpecan_b200.so; fi
  echo "== $v" >> gpurun_out/${tag}.log
  CPB_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2 >> gpurun_out/${tag}.log
  CPB_LIB=$PWD/$lib timeout 300 python bench.py --pairs 12000 --skip-e2e --skip-cpu --steps 3 --warmup 2 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        j=json.loads(l); print('$v', round(j['value'],3), j['phase_ms_per_step'])
" >> gpurun_out/${tag}.log
done
cat gpurun_out/${tag}.log
