#!/bin/bash
# GPU box with G GPUs: host-link ceiling, the bench line under torchrun, and the configs[4] sweep.
#   tools/multi_gpu_round.sh G OUTDIR
G=$1; OUT=${2:-gpurun_out/mg$G}
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
python tools/host_link_peak.py --gpus $G --mb 18 --seconds 1.0 --out $OUT/host_link.json > $OUT/host_link.log 2>&1
if [ "$G" = "1" ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_s20.json 2> $OUT/bench_s20.err
  python bench.py --gpus 1 --steps 4000 --warmup 20 > $OUT/bench_s3000.json 2> $OUT/bench_s3000.err
  python tools/sweep_config5.py --out $OUT/sweep.jsonl > $OUT/sweep.log 2>&1
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29501 \
    bench.py --gpus $G --steps 20 --warmup 5 > $OUT/bench_s20.json 2> $OUT/bench_s20.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus $G --steps 3000 --warmup 20 > $OUT/bench_s3000.json 2> $OUT/bench_s3000.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29503 \
    tools/sweep_config5.py --out $OUT/sweep.jsonl > $OUT/sweep.log 2>&1
fi
tail -c 400 $OUT/host_link.log; echo; tail -3 $OUT/sweep.log | cut -c1-400
