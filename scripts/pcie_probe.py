"""PCIe probe: H2D alone, D2H alone, both at once (pinned buffers, two streams)."""
import time
import torch

dev = torch.device("cuda", 0)
N = 88_000_000
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
d_in = torch.empty(N, dtype=torch.uint8, device=dev)
d_out = torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, chunks=1, reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    c = N // chunks
    for _ in range(reps):
        for i in range(chunks):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[i * c:(i + 1) * c].copy_(d_out[i * c:(i + 1) * c], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return dt * 1e3, N * (h2d + d2h) / dt / 1e9


for name, a, b, ch in (("h2d", 1, 0, 1), ("d2h", 0, 1, 1), ("both", 1, 1, 1), ("both x8 chunks", 1, 1, 8),
                       ("h2d x16 chunks", 1, 0, 16)):
    run(a, b, ch, 2)
    ms, gbs = run(a, b, ch)
    print(f"{name:16s} {ms:7.3f} ms  {gbs:6.1f} GB/s total")
