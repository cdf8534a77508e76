"""Read-only and copy bandwidth of this GPU with plain torch kernels (calibration of the HBM roof for read-dominated kernels)."""
import torch
x = torch.empty(2 * 1024 ** 3 // 4, dtype=torch.float32, device="cuda").normal_()
y = torch.empty_like(x)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: x.sum())
print("read-only  (sum of 2 GiB fp32): %.3f ms -> %.0f GB/s" % (ms, x.numel() * 4 / ms / 1e6))
ms = t(lambda: torch.max(x))
print("read-only  (max of 2 GiB fp32): %.3f ms -> %.0f GB/s" % (ms, x.numel() * 4 / ms / 1e6))
ms = t(lambda: y.copy_(x))
print("copy       (2 GiB -> 2 GiB):    %.3f ms -> %.0f GB/s read+write" % (ms, 2 * x.numel() * 4 / ms / 1e6))
ms = t(lambda: y.fill_(1.0))
print("write-only (fill 2 GiB):        %.3f ms -> %.0f GB/s" % (ms, x.numel() * 4 / ms / 1e6))
