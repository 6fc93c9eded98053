#!/bin/bash
# ncu --set full capture of the fused loop on one config (GPU box, through gpurun): tools/ncu_loop.sh <name> <config> <S>
# Writes gpurun_out/<name>.ncu-rep (+ the plain run's line, which must exit 0 first).
name=$1; cfg=$2; S=$3; shift 3   # further arguments go to tools/run_loop.py (inner policy, state-row mode)
python tools/run_loop.py $cfg $S "$@" > gpurun_out/${name}_plain.log 2>&1 || { tail -5 gpurun_out/${name}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:closed_loop --launch-skip 1 --launch-count 1 \
    -o gpurun_out/${name} -f python tools/run_loop.py $cfg $S "$@" > gpurun_out/${name}_ncu.log 2>&1
tail -1 gpurun_out/${name}_plain.log
ls -la gpurun_out/${name}.ncu-rep
