import time, torch
n = 64*2*1025*512
h = torch.empty(n, dtype=torch.complex64).pin_memory()
h2 = torch.empty(n, dtype=torch.complex64).pin_memory()
d = torch.empty(n, dtype=torch.complex64, device='cuda')
d2 = torch.empty(n, dtype=torch.complex64, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, rep=3):
    best = 1e9
    for _ in range(rep):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter()-t0)
    return best
gb = n*8/1e9
a = t(lambda: d.copy_(h, non_blocking=True)); print('H2D %.2f ms %.1f GB/s' % (a*1e3, gb/a))
b = t(lambda: h2.copy_(d2, non_blocking=True)); print('D2H %.2f ms %.1f GB/s' % (b*1e3, gb/b))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print('both %.2f ms' % (c*1e3))
