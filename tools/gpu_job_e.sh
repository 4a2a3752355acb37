#!/bin/bash
mkdir -p gpurun_out
XYST_FLUX_MODE=1 XYST_REORDER=0 XYST_GRAD_MODE=1 XYST_B200_LIB=tools/_lib/lib_gsm4.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_solver.py tests/test_gpu_laxcg.py -m gpu -q --timeout 900 > gpurun_out/r2e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2e_tests.log
python tools/variants.py run 150 > gpurun_out/r2e_variants.log 2>&1
XYST_FLUX_MODE=1 XYST_REORDER=0 XYST_GRAD_MODE=1 XYST_B200_LIB=tools/_lib/lib_gsm4.so ncu --set full --clock-control none --import-source on -k regex:"k_flux_own|k_grad_node_p" -s 8 -c 2 \
    -o gpurun_out/prof_r2e python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2e_ncu_full.log 2>&1
tail -3 gpurun_out/r2e_tests.log; cat gpurun_out/r2e_variants.log
