#!/usr/bin/env python
"""Pinned H2D bandwidth on GPU 0: one stream vs several concurrent streams (copy engines), 2-D vs 1-D copies."""
import os, sys, time
import torch

torch.cuda.init()
n = 192 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda:0")
for ns in (1, 2, 3, 4):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    chunk = n // ns
    best = 0.0
    for _ in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, n / (time.perf_counter() - t0) / 1e9)
    print(f"h2d {ns} stream(s): {best:.1f} GB/s", flush=True)
# many small-ish copies on one stream (like 32 planes per tick)
st = torch.cuda.Stream()
for pieces in (32, 8):
    chunk = n // pieces
    best = 0.0
    for _ in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            for i in range(pieces):
                d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, n / (time.perf_counter() - t0) / 1e9)
    print(f"h2d 1 stream, {pieces} copies: {best:.1f} GB/s", flush=True)
# simultaneous d2h on another stream
s2 = torch.cuda.Stream()
h2 = torch.empty(16 << 20, dtype=torch.uint8).pin_memory()
best = 0.0
for _ in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(st): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d[:16 << 20], non_blocking=True)
    torch.cuda.synchronize()
    best = max(best, n / (time.perf_counter() - t0) / 1e9)
print(f"h2d with concurrent 16MB d2h: {best:.1f} GB/s")
