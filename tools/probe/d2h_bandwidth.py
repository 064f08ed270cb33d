"""pinned host <-> device copy bandwidth of the box (torch, 16.8 MB = one ensemble state, and 168 MB):
measured 55 GB/s each way on this pool's B200 boxes, i.e. the per-step snapshot copy of the e2e path
(16.8 MB, 0.3 ms) hides behind a 1.2 ms step on one GPU"""
import torch, time
for nbytes in (16_800_000, 168_000_000):
    n = nbytes//8
    d = torch.empty(n, dtype=torch.float64, device='cuda')
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    s = torch.cuda.Stream()
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for k in range(20):
                h.copy_(d, non_blocking=True)
        s.synchronize()
        t = time.perf_counter() - t0
        print('D2H', nbytes, 'bytes:', 20*nbytes/t/1e9, 'GB/s')
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for k in range(20):
                d.copy_(h, non_blocking=True)
        s.synchronize()
        t = time.perf_counter() - t0
        print('H2D', nbytes, 'bytes:', 20*nbytes/t/1e9, 'GB/s')
