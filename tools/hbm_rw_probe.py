"""HBM read-only / write-only / copy bandwidth on this GPU (torch kernels, CUDA events, best of 10)."""
import torch

n = 1 << 30  # 1 Gi elements
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")


def best(fn, nbytes, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(b))
    return nbytes / min(t) * 1e-6, nbytes / (sum(t) / len(t)) * 1e-6


print("write  (fill_)   GB/s best / mean: %.0f / %.0f" % best(lambda: x.fill_(1.0), 4 * n))
print("write  (zero_)   GB/s best / mean: %.0f / %.0f" % best(lambda: x.zero_(), 4 * n))
print("read   (sum)     GB/s best / mean: %.0f / %.0f" % best(lambda: x.sum(), 4 * n))
print("copy   (r + w)   GB/s best / mean: %.0f / %.0f" % best(lambda: y.copy_(x), 8 * n))
h = x.view(torch.int32)[: n // 2]
print("cast f32->f16 (4B read + 2B write) GB/s: %.0f / %.0f" % best(lambda: torch.empty(n, dtype=torch.float16, device="cuda").copy_(x), 6 * n))
