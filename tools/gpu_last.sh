#!/bin/bash
# last short single-GPU sanity run of the round: smoke + the RieCG scalar tests + one ZalCG / KozCG regression each
mkdir -p gpurun_out
( python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ;
  timeout 200 python -m pytest tests/test_gpu_scalars.py tests/test_gpu_zalcg.py tests/test_gpu_kozcg.py -q -x ) > gpurun_out/r2q_last.log 2>&1
echo "rc=$?" >> gpurun_out/r2q_last.log
grep -v "^$" gpurun_out/r2q_last.log | tail -8
