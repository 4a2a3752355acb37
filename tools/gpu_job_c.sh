#!/bin/bash
mkdir -p gpurun_out
XYST_FLUX_MODE=3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solver.py tests/test_gpu_laxcg.py -m gpu -q --timeout 900 > gpurun_out/r2c_tests_m3.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c_tests_m3.log
python tools/variants.py run 150 > gpurun_out/r2c_variants.log 2>&1
XYST_FLUX_MODE=3 XYST_B200_LIB=tools/_lib/lib_own2_sint.so ncu --set full --clock-control none --import-source on -k regex:"k_flux_own2|k_update_in" -s 8 -c 2 \
    -o gpurun_out/prof_r2c python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2c_ncu_full.log 2>&1
tail -3 gpurun_out/r2c_tests_m3.log; cat gpurun_out/r2c_variants.log
