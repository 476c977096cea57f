import torch, sys
n = 108_000_000
x = torch.empty(n, dtype=torch.float64, device='cuda')
y = torch.empty(n, dtype=torch.float64, device='cuda')
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: x.fill_(1.5)); print('fill  %.1f us  %.0f GB/s (write only)' % (ms * 1e3, n * 8 / ms / 1e6))
ms = t(lambda: y.copy_(x)); print('copy  %.1f us  %.0f GB/s (read+write)' % (ms * 1e3, 2 * n * 8 / ms / 1e6))
ms = t(lambda: x.sum()); print('sum   %.1f us  %.0f GB/s (read only)' % (ms * 1e3, n * 8 / ms / 1e6))
