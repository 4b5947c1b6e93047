#!/usr/bin/env python3
"""PCIe copy rates of the box (pinned memory), to put the e2e number in context."""
import torch, time
dev = torch.device("cuda:0")
h_in = torch.empty(97 << 20, dtype=torch.uint8).pin_memory(); d_in = torch.empty_like(h_in, device=dev)
d_out = torch.empty(268 << 20, dtype=torch.uint8, device=dev); h_out = torch.empty(268 << 20, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, n=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print("H2D 97 MiB %.2f ms (%.1f GB/s)  D2H 268 MiB %.2f ms (%.1f GB/s)  both %.2f ms" % (a, (97 << 20) / a / 1e6, b, (268 << 20) / b / 1e6, c))
