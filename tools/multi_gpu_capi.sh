#!/bin/bash
# The C library over several GPUs of one box (gpurun --gpus N): the two-GPU parity tests, then the reference-named batch call and an
# EM iteration over 1 .. N devices on fixed batches (strong scaling: 20 000 x 1 kb, 400 x 100 kb, 50 000 x 2 kb), written to gpurun_out/<tag>_multi_gpu_capi.jsonl
tag=${1:-r2}; n=${2:-2}
out=gpurun_out/${tag}_multi_gpu_capi.jsonl
python -m pytest "tests/test_host_c_api.py::test_c_api_matches_the_oracle[0-0,1]" "tests/test_host_c_api.py::test_c_api_matches_the_oracle[2-0,1]" -m gpu -q 2>&1 | tail -2
python tools/write_workload.py /tmp/capi_c2.bin 20000 1000 14 20
python tools/write_workload.py /tmp/capi_c4.bin 50000 2000 0 10
python tools/write_workload.py /tmp/capi_c3.bin ${C3_PAIRS:-400} 100000 14 20 64
: > $out
for d in 1 2 4 8; do
  [ $d -gt $n ] && break
  CPECAN_DEVICES=$d cpecan_b200/lib/bench_capi /tmp/capi_c2.bin 2 1 | sed "s/^{/{\"config\": \"C2 subset: 20000 x 1 kb through getAlignedPairsUsingAnchorsBatch\", /" >> $out
  CPECAN_DEVICES=$d cpecan_b200/lib/bench_capi /tmp/capi_c3.bin 1 1 | sed "s/^{/{\"config\": \"C3 subset: ${C3_PAIRS:-400} x 100 kb through getAlignedPairsUsingAnchorsBatch\", /" >> $out
  CPECAN_DEVICES=$d cpecan_b200/lib/bench_capi /tmp/capi_c4.bin 3 1 em 2>> gpurun_out/${tag}_multi_gpu_capi.err | sed "s/^{/{\"config\": \"C4: 50000 x 2 kb expectation pass + NCCL all-reduce (cpecanResidentBatch)\", /" >> $out
done
cat $out
