#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
XYST_LOOKBACK=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solver.py tests/test_gpu_laxcg.py -m gpu -q --timeout 900 > gpurun_out/r2b_tests_lb.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests_lb.log
python tools/variants.py run 150 > gpurun_out/r2b_variants.log 2>&1
XYST_LOOKBACK=1 ncu --set full --clock-control none --import-source on -k regex:"k_stage_tile|k_grad_node" -s 8 -c 2 \
    -o gpurun_out/prof_r2b python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu_full.log 2>&1
tail -3 gpurun_out/r2b_tests.log; tail -3 gpurun_out/r2b_tests_lb.log; cat gpurun_out/r2b_variants.log
