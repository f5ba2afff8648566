#!/bin/bash
# ncu evidence of a round on the GPU box (run only after the plain commands exit 0):
#   gpurun --timeout 1500 -- 'bash tools/ncu_round.sh r02'
# 1. launch list (gpu__time_duration) of the bench command and of one pass over every hot-path kernel
# 2. one --set full capture of every seir_ kernel once
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_plain.json 2> $out/${tag}_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_ncu.log 2>&1
SEIR_SWEEP_GROUPS=1 python tools/profile_once.py > $out/${tag}_once_plain.log 2>&1 || { tail -3 $out/${tag}_once_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/${tag}_launches.csv \
    env SEIR_SWEEP_GROUPS=1 python tools/profile_once.py > $out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:seir_ -o $out/${tag}_full \
    env SEIR_SWEEP_GROUPS=1 python tools/profile_once.py > $out/${tag}_ncu_full.log 2>&1
tail -2 $out/${tag}_ncu_full.log
