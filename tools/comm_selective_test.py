#!/usr/bin/env python
"""2+ GPUs under torchrun: smr_comm_exchange_inputs sends a frame only to the ranks that consume it.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/comm_selective_test.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import smelter_b200 as s
from smelter_b200 import _ffi as F

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
r = s.Renderer(s.RendererOptions(cuda_device=local))
uid = [r.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
r.comm_init(uid[0], rank, world)
w, h = 640, 360
n = 3
# frame 0: root 0 -> rank 1 only; frame 1: root 1 -> nobody else; frame 2: root 0 -> everyone (broadcast path)
roots = [0, 1 % world, 0]
masks = [1 << (1 % world), 1 << (1 % world), (1 << world) - 1]
ys, uvs = [], []
arr = (F.InputFrame * n)()
keep = []
for i in range(n):
    fill = 10 * (i + 1) + (100 if rank == roots[i] else 0)        # the root's copy is marked by +100
    y = torch.full((h, w), fill, dtype=torch.uint8, device=dev)
    uv = torch.full((h // 2, w // 2, 2), fill + 1, dtype=torch.uint8, device=dev)
    ys.append(y); uvs.append(uv)
    b = f"input_{i}".encode(); keep.append(b)
    arr[i].input_id = b
    arr[i].format = F.FRAME_NV12
    arr[i].width, arr[i].height = w, h
    arr[i].mem_kind = F.MEM_DEVICE
    arr[i].planes[0], arr[i].planes[1] = y.data_ptr(), uv.data_ptr()
r.comm_exchange_inputs(arr, n, roots, masks, pooled=False)
torch.cuda.synchronize()
torch.cuda.current_stream().synchronize()
import ctypes as C
C.CDLL("libcudart.so", mode=C.RTLD_GLOBAL) if False else None
torch.cuda.synchronize(dev)
# the exchange runs on the handle's communication stream: a device-wide sync covers it
ok = True
for i in range(n):
    consumer = (masks[i] >> rank) & 1 or rank == roots[i]
    want = 10 * (i + 1) + 100 if consumer else 10 * (i + 1)
    got = int(ys[i][0, 0].item())
    if got != want or int(uvs[i][0, 0, 0].item()) != want + 1:
        ok = False
        print(f"rank {rank} frame {i}: got {got}, expected {want} (consumer={bool(consumer)})")
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("selective exchange:", "ok" if int(t.item()) == 1 else "FAILED")
r.comm_destroy()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
