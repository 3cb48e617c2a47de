import torch, time
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device='cuda')
a = torch.randn(8192, 8192, device='cuda', dtype=torch.bfloat16)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def work():
    with torch.cuda.stream(s1):
        for _ in range(60): a @ a
def copy():
    with torch.cuda.stream(s2): d.copy_(h, non_blocking=True)
for name, fns in (('kernels only', [work]), ('copy only', [copy]), ('both', [work, copy])):
    torch.cuda.synchronize(); t = time.perf_counter()
    for f in fns: f()
    torch.cuda.synchronize(); print(name, '%.1f ms' % ((time.perf_counter() - t) * 1e3))
