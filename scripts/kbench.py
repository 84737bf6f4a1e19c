"""Per-kernel timings of the spectral path (L2 flushed between reps) for one workload:
   python scripts/kbench.py [cfg2|cfg3|cfg1] ; tunables are read from the environment by the library."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
peak = 6547.2
if os.environ.get('SB200_TC_MODE'):
    from dlwp_benchmark_b200 import _lib
    _lib.set_tc_mode(int(os.environ['SB200_TC_MODE']))
kr = bench.kernel_rooflines(wl, peak)
env = {k: v for k, v in os.environ.items() if k.startswith("SB200_")}
print(json.dumps({"env": env, "kernels": {k: {"us": round(v["ms"] * 1e3, 1), "frac": round(v["frac"], 3)} for k, v in kr.items()}}))
