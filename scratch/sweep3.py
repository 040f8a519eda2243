import os, sys, json, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for pad in ["0", "64"]:
    env = dict(os.environ, MT_EXP_PAD=pad)
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "20", "--no-cpu-baseline", "--e2e-steps", "1"],
                         env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        ks = {k["kernel"]: (k["ms_per_step"], k["achieved_gbs"]) for k in d["roofline"]["kernels"]}
        print("pad", pad, "step_ms=%.3f" % d["ms_per_step"], " ".join(f"{k}={v[0]:.3f}" for k, v in ks.items() if v[0] > 0.2), flush=True)
    except Exception as e:
        print(pad, "FAILED", e, out.stderr[-1500:])
