#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the NCCL parity tests and the strong-scaling benchmark line.
N=${1:-2}; tag=${2:-runN}; steps=${3:-10}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  python -m pytest tests/test_multi_partition.py tests/test_gpu_cg.py -m gpu -q --timeout 1800 > gpurun_out/${tag}_tests.log 2>&1
  echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
  tail -3 gpurun_out/${tag}_tests.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps $steps --warmup 3 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/${tag}_bench_n$N.json; tail -5 gpurun_out/${tag}_bench_n$N.err
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
