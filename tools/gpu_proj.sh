#!/bin/bash
# 2-GPU checks against the oracle's 2-chare runs: $1 = pytest -k expression
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_multi_partition.py -m gpu -x -q -s -k "${1:-projection}" > gpurun_out/r2m_proj.log 2>&1
echo "rc=$?" >> gpurun_out/r2m_proj.log
grep -v "^$" gpurun_out/r2m_proj.log | cut -c1-400 | tail -40
