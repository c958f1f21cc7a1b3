"""H2D, D2H and both at once between pinned host memory and the GPU: what bounds bench.py's e2e variants (640 MB each way per step at 10 M entities)."""
import json
import time

import torch

n = 640_000_000
h_up = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_down = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_a.copy_(h_up, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_down.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


run(True, True, 1)
out = {"bytes_each_way": n, "h2d_ms": run(True, False), "d2h_ms": run(False, True), "both_ms": run(True, True)}
out["h2d_gbs"] = n / out["h2d_ms"] / 1e6
out["d2h_gbs"] = n / out["d2h_ms"] / 1e6
out["both_aggregate_gbs"] = 2 * n / out["both_ms"] / 1e6
print(json.dumps(out))
