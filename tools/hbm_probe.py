import torch
x = torch.empty(102400, 768, device="cuda"); y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
us = t(lambda: x.fill_(1.0)); print(f"fill 315 MB: {us:.1f} us  {x.numel()*4/us/1e6:.2f} TB/s write")
us = t(lambda: y.copy_(x)); print(f"copy 315 MB: {us:.1f} us  {2*x.numel()*4/us/1e6:.2f} TB/s r+w")
us = t(lambda: x.sum()); print(f"sum  315 MB: {us:.1f} us  {x.numel()*4/us/1e6:.2f} TB/s read")
# strided writes: 1 KB segments at a 3 KB pitch (one third of the rows' bytes), and the three thirds one after the other
us = t(lambda: x[:, :256].fill_(1.0)); print(f"fill [:, :256] (1 KB of every 3 KB): {us:.1f} us  {x.numel()*4/3/us/1e6:.2f} TB/s write")
def thirds():
    x[:, :256].fill_(1.0); x[:, 256:512].fill_(2.0); x[:, 512:].fill_(3.0)
us = t(thirds); print(f"three strided fills (all bytes): {us:.1f} us  {x.numel()*4/us/1e6:.2f} TB/s write")
z = torch.empty(102400, 64, device="cuda")
us = t(lambda: x.copy_(z.repeat(1, 12))); print(f"repeat+copy: {us:.1f} us")
