import torch, time
x = torch.empty(256, 524160, pin_memory=True); d = torch.empty(256, 524160, device="cuda")
y = torch.empty(256, 4096, 80, pin_memory=True); dy = torch.empty(256, 4096, 80, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("h2d", lambda: d.copy_(x, non_blocking=True)), ("d2h", lambda: y.copy_(dy, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    nbytes = (x if name == "h2d" else y).numel() * 4
    print(name, f"{nbytes / dt / 1e9:.1f} GB/s")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(dy, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
print("both directions:", f"h2d {x.numel()*4/dt/1e9:.1f} GB/s + d2h {y.numel()*4/dt/1e9:.1f} GB/s, {dt*1e3:.2f} ms per step-equivalent")
