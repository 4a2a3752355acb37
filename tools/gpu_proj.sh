#!/bin/bash
# 2-GPU check of the partitioned projection solvers (ChoCG / LohCG) against the oracle's 2-chare runs
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multi_partition.py -m gpu -x -q -s -k "projection" > gpurun_out/r2m_proj.log 2>&1
echo "rc=$?" >> gpurun_out/r2m_proj.log
tail -40 gpurun_out/r2m_proj.log
