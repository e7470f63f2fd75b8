#!/bin/bash
# Final GPU check of the round: the whole GPU suite as the driver runs it, then the bench line
# (with the partition sub-run on the 5 M-tet mesh of BASELINE configs[4]).
mkdir -p gpurun_out
timeout 160 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2_final_gputest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_final_gputest.log
tail -4 gpurun_out/r2_final_gputest.log
timeout 200 python bench.py --partition-tets 5e6 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_final.json; tail -3 gpurun_out/r2_bench_final.err
