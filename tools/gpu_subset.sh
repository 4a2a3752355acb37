#!/bin/bash
# quick single-GPU regression subset after changes to the linear solver set-up, halo lists and host problems
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cg.py tests/test_gpu_scalars.py tests/test_gpu_errors.py tests/test_gpu_parity.py tests/test_gpu_shim.py -q -x > gpurun_out/r2o_subset.log 2>&1
echo "rc=$?" >> gpurun_out/r2o_subset.log
grep -v "^$" gpurun_out/r2o_subset.log | tail -15
