"""Measured HBM rates of simple torch kernels on this box (context for the assembly roofline):
pure write (fill), pure read (sum), copy (read + write)."""
import json, torch
n = 1 << 28  # 2 GiB of fp64
a = torch.empty(n, dtype=torch.float64, device='cuda')
b = torch.empty(n, dtype=torch.float64, device='cuda')
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = {}
out['write_GBs'] = 8 * n / timed(lambda: a.fill_(1.0)) / 1e6
out['read_GBs'] = 8 * n / timed(lambda: a.sum()) / 1e6
out['copy_GBs'] = 16 * n / timed(lambda: b.copy_(a)) / 1e6
print(json.dumps({k: round(v, 1) for k, v in out.items()}))
