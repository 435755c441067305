"""Write-bandwidth probe: time torch fills / copies of observation-sized buffers (CUDA events)."""
import torch
n, d = 131072, 107
bufs = [torch.empty((n, d), dtype=torch.float32, device="cuda") for _ in range(5)]
src = torch.randn((n, d), device="cuda")
def timeit(fn, iters=200):
    for _ in range(10): fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
mb = n * d * 4 / 1e6
t = timeit(lambda i: bufs[i % 5].fill_(1.0)); print(f"fill ring5: {t:.2f} us  {mb / t * 1e-3:.2f} TB/s write")
t = timeit(lambda i: bufs[0].fill_(1.0)); print(f"fill same : {t:.2f} us  {mb / t * 1e-3:.2f} TB/s write")
t = timeit(lambda i: bufs[i % 5].copy_(src)); print(f"copy ring5: {t:.2f} us  {2 * mb / t * 1e-3:.2f} TB/s r+w")
big = torch.empty(1 << 28, dtype=torch.float32, device="cuda"); big2 = torch.empty_like(big)
t = timeit(lambda i: big.fill_(0.0), 20); print(f"fill 1GiB : {t:.1f} us  {big.numel() * 4 / 1e6 / t * 1e-3:.2f} TB/s write")
t = timeit(lambda i: big2.copy_(big), 20); print(f"copy 1GiB : {t:.1f} us  {2 * big.numel() * 4 / 1e6 / t * 1e-3:.2f} TB/s r+w")
