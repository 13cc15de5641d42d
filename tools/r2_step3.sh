#!/bin/bash
# GPU box, 2 GPUs: chain sanity, the row-sharded group (peer-mapped history) bit-exactness tests, a short N = 2 bench line
cd "$(dirname "$0")/.."
nvidia-smi -L | head -3
python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -25
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/step3_n2.json 2> gpurun_out/step3_n2.err
tail -c 3000 gpurun_out/step3_n2.json; tail -15 gpurun_out/step3_n2.err
