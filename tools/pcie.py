"""development: raw pinned-memory PCIe bandwidth on this box (D2H alone, H2D alone, both at once) — the ceiling of the e2e leg"""
import torch, time
n = 49152000 * 4
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True); d_out = torch.empty(n, dtype=torch.uint8, device='cuda')
h_in = torch.empty(n // 2, dtype=torch.uint8, pin_memory=True); d_in = torch.empty(n // 2, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(d2h, h2d, reps=10):
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(reps):
        if d2h:
            with torch.cuda.stream(s1): h_out.copy_(d_out, non_blocking=True)
        if h2d:
            with torch.cuda.stream(s2): d_in.copy_(h_in, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    return (n * reps / dt / 1e9 if d2h else 0.0), (n // 2 * reps / dt / 1e9 if h2d else 0.0)
run(True, True, 2)
print('D2H alone %.1f GB/s | H2D alone %.1f GB/s | concurrent: D2H %.1f + H2D %.1f GB/s (bytes in the e2e ratio 2:1)' % (run(True, False)[0], run(False, True)[1], *run(True, True)))
