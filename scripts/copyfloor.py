import torch, time
n = 148 * 1000 * 1000
a = torch.randn(n, device="cuda"); b = torch.empty_like(a)
for nm, f in [("copy a->b", lambda: b.copy_(a)), ("alternate", None)]:
    if f is None:
        def f():
            b.copy_(a); a.copy_(b)
        reps, per = 50, 2
    else:
        reps, per = 100, 1
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * per)
    print(nm, "ms per 592 MB copy", round(ms, 4), "GB/s", round(2 * n * 4 / ms / 1e6, 1))
