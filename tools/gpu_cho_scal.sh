#!/bin/bash
# single-GPU run of the transported-scalar / frozen-flow tests with the regression tests of one solver family:
#   tools/gpu_cho_scal.sh [pytest -k expression] [extra test files...]
mkdir -p gpurun_out
K="${1:-chocg or lohcg or zalcg or kozcg}"; shift
timeout 500 python -m pytest tests/test_gpu_scalars.py "$@" -q -s -k "$K" > gpurun_out/r2n_cho.log 2>&1
echo "rc=$?" >> gpurun_out/r2n_cho.log
grep -v "^$" gpurun_out/r2n_cho.log | tail -60
