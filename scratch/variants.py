"""Build tuning variants of the device library into scratch/ and (on the GPU box) time them."""
import os, sys, subprocess, json
sys.path.insert(0, '.')
from xyst_b200 import build as B
VAR = {
 "g7r7": [],
 "g14r7": ["GRAD_UNROLL=14"],
 "g7r14": ["RHS_UNROLL=14"],
 "g5r7": ["GRAD_UNROLL=5"],
 "g7r5": ["RHS_UNROLL=5"],
 "g7r7n128": ["NODE_THREADS=128", "GRAD_MINB=8", "RHS_MINB=8"],
 "g7r7f7": ["FLUX_MINB=7"],
}
if len(sys.argv) > 3:
    VAR = {k: v for k, v in VAR.items() if k in sys.argv[3].split(",")}
if sys.argv[1] == "build":
    for k, d in VAR.items():
        B.build_device(force=True, out="scratch/lib_%s.so" % k, defines=d, verbose=False)
        print("built", k)
else:
    for k in VAR:
        env = dict(os.environ, XYST_B200_LIB="scratch/lib_%s.so" % k)
        r = subprocess.run([sys.executable, "bench.py", "--steps", "8", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--n", sys.argv[2] if len(sys.argv) > 2 else "150"],
                           env=env, capture_output=True, text=True)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            print(k, "ms/step %.3f" % j["ms_per_step"], j["roofline_stage"]["kernel_ms_per_stage"], flush=True)
        except Exception as e:
            print(k, "FAILED", r.stdout[-300:], r.stderr[-300:])
