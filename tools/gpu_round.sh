#!/bin/bash
# One GPU-box visit: parity tests, bench, per-launch timing list, full ncu capture of every kernel once.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r1f'
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
tail -c 600 $out/${tag}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
python tools/profile_sweep.py --sweeps 5 > $out/${tag}_sweep.json 2> $out/${tag}_sweep.err; cat $out/${tag}_sweep.json
if [ "$2" != "noncu" ]; then
SEIR_SWEEP_GROUPS=1 python tools/profile_once.py > $out/${tag}_once_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/${tag}_launches.csv \
    env SEIR_SWEEP_GROUPS=1 python tools/profile_once.py > $out/${tag}_ncu_list.log 2>&1
SEIR_SWEEP_GROUPS=1 python tools/profile_once.py > $out/${tag}_once_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:seir_ -o $out/${tag}_full \
    env SEIR_SWEEP_GROUPS=1 python tools/profile_once.py > $out/${tag}_ncu_full.log 2>&1
tail -2 $out/${tag}_ncu_full.log
fi
