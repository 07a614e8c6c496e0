#!/bin/bash
# One B200: the bench itself (both arms), then -- only after the same command exited 0 without ncu -- its ncu launch list and one
# full-set capture of the wavefront kernels.   usage: bash tools/profile_round.sh <tag>     outputs: gpurun_out/<tag>_*
tag=${1:-r2}
out=gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench.err
python bench.py > $out/${tag}_bench.json 2>> $out/${tag}_bench.err || { tail -5 $out/${tag}_bench.err; exit 1; }
small="python bench.py --pairs 12000 --skip-e2e --skip-cpu --steps 1 --warmup 1"
$small > $out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $small > $out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_forward_strip|k_backward_strip' --launch-skip 2 -c 2 \
    -f -o $out/${tag}_prof $small > $out/${tag}_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:'k_posterior|k_totals|k_band' --launch-skip 5 -c 4 -f -o $out/${tag}_prof_small $small > $out/${tag}_ncu_small.log 2>&1
tail -2 $out/${tag}_ncu_full.log
cat $out/${tag}_bench.json
