#!/bin/bash
# Last check of a round on the GPU box: every GPU test, smoke(), both bench arms, the host-path probe.
#   gpurun --timeout 1200 -- "bash tools/final_check.sh"
python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; tail -2 gpurun_out/r2z_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2z_bench_ref.json 2>> gpurun_out/r2z_bench.err
python tools/e2e_probe.py --quick > gpurun_out/r2z_probe.json 2> gpurun_out/r2z_probe.err; tail -2 gpurun_out/r2z_probe.err | cut -c1-400
