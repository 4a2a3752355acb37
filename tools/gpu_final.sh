#!/bin/bash
# Final verification of a round on one GPU: the whole -m gpu suite, smoke(), the default benchmark line,
# the reference arm, and the FCT solvers' secondary rates.
tag=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1800 > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
python bench_secondary.py --only zalcg,kozcg > gpurun_out/${tag}_secondary.jsonl 2> gpurun_out/${tag}_secondary.err
tail -4 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_smoke.log | tail -2; cut -c1-400 gpurun_out/${tag}_bench_n1.json; cut -c1-300 gpurun_out/${tag}_secondary.jsonl
