#!/bin/bash
# GPU job: parity tests, variant timings, ncu captures (run under gpurun from the repo root)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
python tools/variants.py run 150 > gpurun_out/r2a_variants.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_stage_tile|k_grad_node" -s 8 -c 2 \
    -o gpurun_out/prof_r2a python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2a_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2a.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2a_ncu_list.log 2>&1
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_variants.log
