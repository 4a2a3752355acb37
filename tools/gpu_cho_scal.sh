#!/bin/bash
# single-GPU check of the projection solvers incl. the transported-scalar cases
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_scalars.py tests/test_gpu_zalcg.py -q -s -k "zalcg" > gpurun_out/r2n_cho.log 2>&1
echo "rc=$?" >> gpurun_out/r2n_cho.log
grep -v "^$" gpurun_out/r2n_cho.log | tail -60
