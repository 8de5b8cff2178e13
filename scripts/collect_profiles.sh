#!/bin/bash
# Run on the GPU box (gpurun): launch list of the bench command + one full ncu capture of the dominant kernel.
#   scripts/collect_profiles.sh <tag> [kernel-regex] [launches of that kernel to skip]
# r02: the dominant kernel is the ring loop of the throughput kernel, the second cuipm_fast launch of a solve.
set -x
tag=${1:-r02}; kre=${2:-cuipm_fast}; skip=${3:-1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-plugin > gpurun_out/${tag}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${kre} -s ${skip} -c 1 -f -o gpurun_out/${tag}_full \
    python scripts/dev_fast.py ncu c2 4096 > gpurun_out/${tag}_full_bench.log 2>&1
tail -2 gpurun_out/${tag}_full_bench.log
