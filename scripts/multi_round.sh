#!/bin/bash
# usage (gpurun --gpus N): scripts/multi_round.sh TAG N  -- the 2-rank GPU tests, then the N-GPU bench the driver runs
tag=$1; n=$2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/${tag}_multitests.log 2>&1; echo "multi tests rc=$? $(tail -1 gpurun_out/${tag}_multitests.log)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
echo "bench rc=$?"; python scripts/bench_digest.py gpurun_out/${tag}_bench_n$n.json | grep -v "^   [a-z_]*kernel" | cut -c1-260
