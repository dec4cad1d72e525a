"""HBM bandwidth of a write-only stream (torch fill_) and a read-only stream (sum) next to the copy figure in
MEASURED_PEAKS.json: what a store-bound kernel such as the tail can expect."""
import torch
n = 1 << 29   # 2 GiB of float32
x = torch.empty(n, device="cuda")
y = torch.empty(n, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
w = t(lambda: x.fill_(1.0))
c = t(lambda: y.copy_(x))
r = t(lambda: x.sum())
print(f"write-only  {n*4/w/1e6:8.1f} GB/s   copy (r+w) {2*n*4/c/1e6:8.1f} GB/s   read-only {n*4/r/1e6:8.1f} GB/s")
