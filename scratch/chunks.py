import os, sys, subprocess, json
for k in sys.argv[1].split(","):
    env = dict(os.environ, XYST_CHUNKS=k)
    r = subprocess.run([sys.executable, "bench.py", "--steps", "10", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--n", "150"],
                       env=env, capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        print("chunks", k, "ms/step %.3f" % j["ms_per_step"], j["roofline_stage"]["kernel_ms_per_stage"], flush=True)
    except Exception as e:
        print(k, "FAILED", r.stdout[-300:], r.stderr[-600:])
