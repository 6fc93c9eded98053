#!/bin/bash
# One ncu --set full capture of the fused kernel on the default bench workload (run on the GPU box through gpurun;
# the same command must have exited 0 without ncu first).  Output: gpurun_out/$1.ncu-rep
set -e
name=${1:-prof_closed_loop}
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${name}_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:closed_loop --launch-skip 3 --launch-count 1 \
    -o gpurun_out/${name} -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${name}_ncu.log 2>&1
tail -1 gpurun_out/${name}_plain.log | cut -c1-200
