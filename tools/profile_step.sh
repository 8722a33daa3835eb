#!/bin/bash
# GPU box: launch list + one `ncu --set full` capture per kernel of the split step in steady state.
#   tools/profile_step.sh <prefix>   -> gpurun_out/<prefix>_{launches.csv,bench.json,<kernel>.ncu-rep,<kernel>_raw.csv,<kernel>_summary.txt}
# Each ncu pass runs only after the same command exited 0 without ncu (B200_PROFILING.md).  Numbers printed by a
# run under ncu are never bench values; <prefix>_bench.json comes from the plain run.
P=${1:-prof}
EXTRA=${SG_PROFILE_ARGS:-}
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 5 --batches 1 --no-cpu-baseline --no-f32 --e2e-seconds 0.1 $EXTRA"
python bench.py --no-cpu-baseline $EXTRA > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err || exit 1
$CMD > /dev/null 2> gpurun_out/${P}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 380 -c 90 --csv \
    --log-file gpurun_out/${P}_launches.csv $CMD > /dev/null 2>&1
for k in hot_kernel cont_kernel tail2_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 135 -c 1 -f \
      -o gpurun_out/${P}_$k $CMD > /dev/null 2>&1
  ncu -i gpurun_out/${P}_$k.ncu-rep --page raw --csv > gpurun_out/${P}_${k}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${P}_$k.ncu-rep 30 > gpurun_out/${P}_${k}_summary.txt 2>&1
  ncu -i gpurun_out/${P}_$k.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${P}_${k}_source.csv.gz
  ls -la gpurun_out/${P}_$k.ncu-rep
  # gpurun brings back at most 64 MiB: only the hot kernel's report is kept (tools/ncu_by_line.py reads it)
  if [ "$k" != "hot_kernel" ] && [ -z "$SG_KEEP_REP" ]; then rm -f gpurun_out/${P}_$k.ncu-rep; fi
  head -12 gpurun_out/${P}_${k}_summary.txt
done
