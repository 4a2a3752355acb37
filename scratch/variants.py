"""Build tuning variants of the device library into scratch/ and (on the GPU box) time them."""
import os, sys, subprocess, json
sys.path.insert(0, '.')
from xyst_b200 import build as B
VAR = {
 "edge": ["FLUX_SLICE=0"],
 "slice2": ["FLUX_SLICE=1", "FSLICE_MINB=2"],
 "slice3": ["FLUX_SLICE=1", "FSLICE_MINB=3"],
 "slice4": ["FLUX_SLICE=1", "FSLICE_MINB=4"],
}
if sys.argv[1] == "build":
    for k, d in VAR.items():
        B.build_device(force=True, out="scratch/lib_%s.so" % k, defines=d, verbose=True)
        print("built", k)
else:
    for k in VAR:
        env = dict(os.environ, XYST_B200_LIB="scratch/lib_%s.so" % k)
        r = subprocess.run([sys.executable, "bench.py", "--steps", "5", "--warmup", "2", "--no-cpu-baseline", "--no-e2e", "--n", sys.argv[2] if len(sys.argv) > 2 else "150"],
                           env=env, capture_output=True, text=True)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            print(k, "ms/step %.3f" % j["ms_per_step"], j["roofline_stage"]["kernel_ms_per_stage"], flush=True)
        except Exception as e:
            print(k, "FAILED", r.stdout[-300:], r.stderr[-300:])
