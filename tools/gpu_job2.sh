#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_shim.py tests/test_gpu_solver.py tests/test_gpu_parity.py -m gpu -q --timeout 1800 > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --workload tg_strong --steps 10 --no-cpu-baseline > gpurun_out/${tag}_bench_tg322_n1.json 2> gpurun_out/${tag}_bench_tg322_n1.err
tail -4 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_bench_n1.json | cut -c1-200; python -c "
import json
for f in ('gpurun_out/${tag}_bench_n1.json','gpurun_out/${tag}_bench_tg322_n1.json'):
    try:
        j=json.load(open(f)); print(f, j['ms_per_step'], j['value'], j['roofline_stage'], j['setup_s'], j['e2e'])
    except Exception as e: print(f, 'FAILED', e)
"; tail -3 gpurun_out/${tag}_bench_tg322_n1.err
