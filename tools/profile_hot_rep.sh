#!/bin/bash
# GPU box: one `ncu --set full --import-source on` capture of ONE kernel (default hot_kernel), report kept
# (gpurun brings back <= 64 MiB).  SPACEGYM_B200_LIB selects a dev build.  tools/profile_hot_rep.sh <outprefix> [kernel-regex]
P=${1:-gpurun_out/hotrep}; K=${2:-hot_kernel}
mkdir -p $(dirname $P)
CMD="python bench.py --steps 20 --warmup 5 --batches 1 --no-cpu-baseline --no-f32 --e2e-seconds 0.1"
$CMD > /dev/null 2> ${P}_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 135 -c 1 -f -o $P $CMD > /dev/null 2>&1
ls -la $P.ncu-rep
