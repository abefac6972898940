#!/bin/bash
# Round 2, third session: the bench line on 2 GPUs with the two-method gather (device-resident step with the
# NCCL all-reduce inside the CUDA graph, e2e through rectify_band_stream).
#   gpurun --gpus 2 --timeout 200 -- 'bash tools/r2e_pass.sh r2e'
set -u
TAG=${1:-r2e}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-configs --no-cpu > "$OUT/bench_n2.json" 2> "$OUT/bench_n2.err"
echo "bench n2 rc=$?" | tee -a "$OUT/status.txt"
tail -c 1500 "$OUT/bench_n2.json" | cut -c1-600
tail -5 "$OUT/bench_n2.err"
