"""The reference's kernels (oracle/_ref) on the FULL 20M-tet Sedov box of the headline benchmark, one
partition, one host core: the same configuration as the GPU line, timed once (set-up takes minutes).
Writes profiles/r2_cpu_full_n150.json."""
import json, os, sys, time, resource
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import oraclelib as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h = 1.2 / 150.0
flavour = "reference" if O.lib("reference") is not None else "port"
t0 = time.perf_counter()
mesh = bench.kuhn_box(n, n, n, n * h, n * h, n * h)
o = O.Oracle(mesh, O.make_cfg(**bench.sedov_kw(h)), flavour)
t_setup = time.perf_counter() - t0
o.step(1)
t0 = time.perf_counter(); o.step(steps); sec = time.perf_counter() - t0
E = bench.box_edges(n, n, n)
out = {"what": "RieCG Sedov, %d^3-cell box (%d tets, %d edges), ONE partition on ONE host core, %s kernels; 1 warm + %d timed steps"
               % (n, 6 * n ** 3, E, flavour, steps),
       "value": E * 3 * steps / sec, "unit": "edge-updates/s", "cores": 1, "kind": flavour, "seconds_per_step": sec / steps,
       "setup_seconds": t_setup, "max_rss_gb": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6,
       "cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t") if os.path.exists("/proc/cpuinfo") else None}
print(json.dumps(out))
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_cpu_full_n%d.json" % n), "w"), indent=1)
