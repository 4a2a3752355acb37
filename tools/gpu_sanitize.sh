#!/bin/bash
# compute-sanitizer memcheck over the projection solvers' small GPU tests (new scalar / frozen-flow paths)
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/r2p_memcheck.txt \
  python -m pytest tests/test_gpu_scalars.py tests/test_gpu_lohcg.py -q -x -k "chocg or lohcg" > gpurun_out/r2p_sanitize.log 2>&1
echo "rc=$?" >> gpurun_out/r2p_sanitize.log
tail -5 gpurun_out/r2p_sanitize.log; tail -8 gpurun_out/r2p_memcheck.txt
