import torch, time
n = 4 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device='cuda')
for name, a, b in (('H2D', d, h), ('D2H', h, d)):
    for _ in range(2):
        torch.cuda.synchronize(); t = time.perf_counter(); a.copy_(b, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(name, 'pinned 4 GiB: %.1f GB/s' % (n / dt / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize(); t = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print('H2D + D2H concurrently: %.1f GB/s each direction' % (n / dt / 1e9))
