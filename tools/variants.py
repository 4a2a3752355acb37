"""Build tuning variants of the device library (compile-time knobs) and time them with bench.py
on the GPU box; runtime knobs go through the environment (XYST_REORDER, XYST_TILE_WX, XYST_GRAD_MODE).

    python tools/variants.py build [-] [names]   # here (nvcc cross-compiles), writes tools/_lib/*.so
    python tools/variants.py run [n] [names]     # on the GPU box
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xyst_b200 import build as B

LIBDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib")
# name -> (defines, environment)
VAR = {
    "default": ([], {}),
    "grad_oneshot": ([], {"XYST_GRAD_MODE": "0"}),
    "tile_order": ([], {"XYST_REORDER": "1"}),
    "tile_order_rows": ([], {"XYST_REORDER": "1", "XYST_TILE_WX": "0.03"}),
    "own_regs": (["OWN_GSMEM=0", "OWN_MINB=3"], {}),
    "own_sint": (["MUSCL_SIGN_INT=1"], {}),
}


def main():
    names = sys.argv[3].split(",") if len(sys.argv) > 3 else list(VAR)
    if sys.argv[1] == "build":
        os.makedirs(LIBDIR, exist_ok=True)
        done = {}
        for k in names:
            d = tuple(VAR[k][0])
            out = os.path.join(LIBDIR, "lib_%s.so" % k)
            if os.path.lexists(out):
                os.remove(out)
            if d not in done:
                done[d] = out
                B.build_device(force=True, out=out, defines=d)
            else:
                os.symlink(os.path.basename(done[d]), out)
            print("built", k, flush=True)
        return
    n = sys.argv[2] if len(sys.argv) > 2 else "150"
    for k in names:
        env = dict(os.environ, XYST_B200_LIB=os.path.join(LIBDIR, "lib_%s.so" % k), **VAR[k][1])
        r = subprocess.run([sys.executable, "bench.py", "--steps", "8", "--warmup", "3", "--no-cpu-baseline",
                            "--no-e2e", "--no-strong-base", "--reforder", "0", "--n", n], env=env, capture_output=True, text=True)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            print(k, "ms/step %.3f" % j["ms_per_step"], "finite", j["finite"],
                  {a: (round(b, 4) if b else b) for a, b in j["roofline_stage"]["kernel_ms_per_stage"].items()},
                  "setup", {a: round(b, 1) for a, b in j["setup_s"].items()}, flush=True)
        except Exception:
            print(k, "FAILED", r.stdout[-300:], r.stderr[-600:], flush=True)


if __name__ == "__main__":
    main()
