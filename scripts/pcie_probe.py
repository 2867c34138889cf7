"""PCIe probe for the e2e path: pinned H2D alone, D2H alone, both concurrently (two streams), in the sizes of one bench
step (201 MB up, 135 MB down) -- tells whether mfa_attention_run_host's 5.1 ms is the link or the implementation."""
import time
import torch

up = torch.empty(201326592, dtype=torch.uint8, pin_memory=True)
down = torch.empty(135266304, dtype=torch.uint8, pin_memory=True)
d_up = torch.empty_like(up, device="cuda")
d_down = torch.empty_like(down, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def h2d():
    with torch.cuda.stream(s1):
        d_up.copy_(up, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        down.copy_(d_down, non_blocking=True)


def both():
    h2d(); d2h()


def chunked(n):
    def f():
        cu, cd = up.numel() // n, down.numel() // n
        for i in range(n):
            with torch.cuda.stream(s1):
                d_up[i * cu:(i + 1) * cu].copy_(up[i * cu:(i + 1) * cu], non_blocking=True)
            with torch.cuda.stream(s2):
                down[i * cd:(i + 1) * cd].copy_(d_down[i * cd:(i + 1) * cd], non_blocking=True)
    return f


a, b, c = run(h2d), run(d2h), run(both)
print(f"H2D 201 MB alone {a:.2f} ms ({201.3 / a:.1f} GB/s); D2H 135 MB alone {b:.2f} ms ({135.3 / b:.1f} GB/s); "
      f"concurrent {c:.2f} ms; 16 chunks each way {run(chunked(16)):.2f} ms; 48+32 chunks {run(chunked(48)):.2f} ms")
