import os, sys, json, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for var in ["4,3", "2,3", "3,3", "6,2", "2,6", "4,6", "0"]:
    env = dict(os.environ, MT_GATHER_BULK=var, MT_GATHER="1r,16")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "10", "--no-cpu-baseline", "--e2e-steps", "1"],
                         env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        ks = {k["kernel"]: k["ms_per_step"] for k in d["roofline"]["kernels"]}
        print(var, "step_ms=%.3f" % d["ms_per_step"], " ".join(f"{k}={v:.3f}" for k, v in ks.items() if v > 0.015), flush=True)
    except Exception as e:
        print(var, "FAILED", e, out.stderr[-800:])
