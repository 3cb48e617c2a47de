import torch, time
for mb in (1, 4, 8.7, 16, 64, 256):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(3)
    d = torch.empty(n, dtype=torch.uint8, device='cuda')
    s = torch.cuda.Stream()
    best = 1e9
    for _ in range(20):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(); d.copy_(h, non_blocking=True); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    h.fill_(5)   # dirty the CPU caches, then copy once
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print('%6.1f MB: best %.3f ms = %.1f GB/s; right after a CPU write %.3f ms' % (mb, best, n / best / 1e6, e0.elapsed_time(e1)))
