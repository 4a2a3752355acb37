#!/bin/bash
mkdir -p gpurun_out
XYST_FLUX_MODE=1 XYST_REORDER=0 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solver.py tests/test_gpu_laxcg.py -m gpu -q --timeout 900 > gpurun_out/r2d_tests_m1.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2d_tests_m1.log
python tools/variants.py run 150 > gpurun_out/r2d_variants.log 2>&1
XYST_FLUX_MODE=1 XYST_REORDER=0 XYST_B200_LIB=tools/_lib/lib_own_nr_sint.so ncu --set full --clock-control none --import-source on -k regex:"k_flux_own|k_grad_node" -s 8 -c 2 \
    -o gpurun_out/prof_r2d python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2d_ncu_full.log 2>&1
tail -3 gpurun_out/r2d_tests_m1.log; cat gpurun_out/r2d_variants.log
