#!/usr/bin/env python
"""PCIe capability of the box for the e2e numbers: pinned-memory H2D alone, D2H alone, and both directions concurrently
on two streams, with the byte counts of bench.py's C3 e2e step (209 MB up, 229 MB down).  Prints one JSON line."""
import json

import torch

dev = torch.device("cuda:0")
up_b, down_b = 209_472_152, 228_883_200
h_up = torch.empty(up_b, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(down_b, dtype=torch.uint8).pin_memory()
d_up = torch.empty(up_b, dtype=torch.uint8, device=dev)
d_dn = torch.empty(down_b, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(up, down, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(n):
        if up:
            with torch.cuda.stream(s1):
                d_up.copy_(h_up, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_dn.copy_(d_dn, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


run(True, True, 2)
r = {"h2d_only_ms": run(True, False), "d2h_only_ms": run(False, True), "both_ms": run(True, True)}
r["h2d_gbs"] = up_b / r["h2d_only_ms"] / 1e6
r["d2h_gbs"] = down_b / r["d2h_only_ms"] / 1e6
r["duplex_aggregate_gbs"] = (up_b + down_b) / r["both_ms"] / 1e6
print(json.dumps(r))
