"""PCIe probe: H2D / D2H time of one step's bytes as 4 separate copies vs 1 packed copy, alone and both directions at once."""
import torch, time
N = 32768
sizes_in = [N * 52, N * 96, N * 156, N * 48]
sizes_out = [N * 168, N * 72, N * 4, N * 1]
dev = "cuda:0"
def bufs(sizes):
    return [torch.empty(s, dtype=torch.uint8).pin_memory() for s in sizes], [torch.empty(s, dtype=torch.uint8, device=dev) for s in sizes]
hin, din = bufs(sizes_in); hout, dout = bufs(sizes_out)
hin1, din1 = bufs([sum(sizes_in)]); hout1, dout1 = bufs([sum(sizes_out)])
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, k=200):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        if h2d:
            with torch.cuda.stream(s1):
                for h, d in zip(*h2d): d.copy_(h, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                for h, d in zip(*d2h): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e6
for name, a, b in (("H2D 4 copies", (hin, din), None), ("H2D 1 copy", (hin1, din1), None), ("D2H 4 copies", None, (hout, dout)),
                   ("D2H 1 copy", None, (hout1, dout1)), ("both, 4+4 copies", (hin, din), (hout, dout)), ("both, 1+1 copies", (hin1, din1), (hout1, dout1))):
    run(a, b, 20)
    us = run(a, b)
    nb = (sum(sizes_in) if a else 0) + (sum(sizes_out) if b else 0)
    print("%-18s %7.1f us per step  (%5.1f GB/s)" % (name, us, nb / us / 1e3))
