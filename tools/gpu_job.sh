#!/bin/bash
# One GPU session: parity tests, the benchmark line, the ncu launch list and a full capture of the
# three main kernels. Run from the repo root under gpurun; results land in gpurun_out/<tag>_*.
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1800 > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_flux_own|k_grad_node_p|k_update_in" -s 9 -c 3 \
    -o gpurun_out/${tag}_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_full.log 2>&1
tail -4 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_bench_n1.json | cut -c1-1500; cat gpurun_out/${tag}_bench_ref.json | cut -c1-600
