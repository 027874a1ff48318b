"""Raw pinned-memory copy bandwidth of the GPU box: H2D alone, D2H alone, both at once (4.42 GB each)."""
import time, torch
n = 4_423_680_000
h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
for name, fn in (("H2D", h2d), ("D2H", d2h), ("H2D + D2H concurrently", both)):
    dt = t(fn)
    print(f"{name:26s} {dt*1e3:7.1f} ms  {n/dt/1e9:6.1f} GB/s per direction")
