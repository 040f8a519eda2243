import torch, time
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (77_760_000, 400_000_000):
    a = torch.empty(n, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
    gb = n * 8 / 1e9
    ms = t(lambda: a.fill_(1.0)); print(f"n={n}: fill   {gb/ms*1e3:8.0f} GB/s ({ms:.3f} ms)")
    ms = t(lambda: torch.cuda.memset if False else a.zero_()); print(f"n={n}: zero_  {gb/ms*1e3:8.0f} GB/s ({ms:.3f} ms)")
    ms = t(lambda: b.copy_(a)); print(f"n={n}: copy   {2*gb/ms*1e3:8.0f} GB/s r+w ({ms:.3f} ms)")
    ms = t(lambda: a.sum()); print(f"n={n}: sum    {gb/ms*1e3:8.0f} GB/s read ({ms:.3f} ms)")
    ms = t(lambda: torch.add(a, 1.0, out=b)); print(f"n={n}: add    {2*gb/ms*1e3:8.0f} GB/s r+w ({ms:.3f} ms)")
    del a, b
