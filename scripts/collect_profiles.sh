#!/bin/bash
# Run on the GPU box (gpurun): launch list + one full ncu capture of the solve kernel for the bench command.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${1:-r01}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${1:-r01}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cuipm_solve -s 1 -c 1 -o gpurun_out/${1:-r01}_full \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${1:-r01}_full_bench.log 2>&1
tail -2 gpurun_out/${1:-r01}_full_bench.log
