"""HBM throughput for read:write mixes (torch elementwise kernels, CUDA events): what a write-heavy kernel can reach.
    python tools/hbm_mix.py"""
import torch
dev = torch.device("cuda:0")
N = 1 << 28                      # 1 GiB of fp32
a = torch.randn(N, device=dev)
b = torch.empty(2, N, device=dev)
c = torch.empty(N, device=dev)


def timeit(fn, nbytes, name):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:40s} {nbytes / best / 1e6:8.1f} GB/s  ({best:.3f} ms)")


timeit(lambda: c.copy_(a), 8 * N, "copy 1 read : 1 write")
timeit(lambda: b.copy_(a.expand(2, N)), 12 * N, "expand copy 1 read : 2 writes")
timeit(lambda: c.fill_(1.0), 4 * N, "fill (write only)")
timeit(lambda: torch.sum(a), 4 * N, "sum (read only)")
