"""tools/h2d_bw.py -- pinned host->device bandwidth of one frame's inputs (45.7 MB), as the 12 tensors ClipRunner copies
and as one packed buffer; tells whether the e2e number of bench.py is compute- or PCIe-bound."""
import time, torch
from itertools import chain
shapes = ((100, 168), (50, 84), (25, 42), (13, 21))
host = [torch.randn(256, h * w).pin_memory() for h, w in shapes] * 2 + [torch.zeros(h * w, dtype=torch.uint8).pin_memory() for h, w in shapes]
dev = [torch.empty_like(t, device="cuda") for t in host]
nbytes = sum(t.numel() * t.element_size() for t in host)
big_h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
big_d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
def run(fn, n=20):
    with torch.cuda.stream(s):
        fn(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(n): fn()
        e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) / n
t12 = run(lambda: [d.copy_(h, non_blocking=True) for d, h in zip(dev, host)])
t1 = run(lambda: big_d.copy_(big_h, non_blocking=True))
td = run(lambda: big_h.copy_(big_d, non_blocking=True))
print(f"bytes/frame {nbytes}: 12 copies {t12:.3f} ms = {nbytes/t12/1e6:.1f} GB/s; 1 copy {t1:.3f} ms = {nbytes/t1/1e6:.1f} GB/s; D2H {nbytes/td/1e6:.1f} GB/s")
