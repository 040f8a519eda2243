import os, sys, json, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for var in ["2r,16", "1r,16", "1c,16", "2c,16", "4r,16", "4c,16", "2r,8", "1r,24", "2c,8", "1c,32", "2r,32"]:
    env = dict(os.environ, MT_GATHER=var)
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "10", "--no-cpu-baseline", "--e2e-steps", "1"],
                         env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        ks = {k["kernel"]: k["ms_per_step"] for k in d["roofline"]["kernels"]}
        print(var, "step_ms=%.3f" % d["ms_per_step"], " ".join(f"{k}={v:.3f}" for k, v in ks.items() if v > 0.03), flush=True)
    except Exception as e:
        print(var, "FAILED", e, out.stderr[-500:])
