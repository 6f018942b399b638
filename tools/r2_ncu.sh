#!/bin/bash
# tools/r2_ncu.sh TAG KERNEL_REGEX [extra bench args]: one `ncu --set full` capture of a kernel inside the C2 bench
cd "$(dirname "$0")/.."
TAG=$1; K=$2; shift; shift
ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 8 --warmup 3 --no-cpu "$@" > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/${TAG}_prof.ncu-rep
