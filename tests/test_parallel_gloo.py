"""N > 1 host path on CPU: output sharding + shared-input replication plan, exercised with 2 processes
over the gloo backend (the GPU path runs the same plan through smr_comm_broadcast_inputs / NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smelter_b200 import parallel


def test_shard_plan_config4_shape():
    """BASELINE config 4: 64 outputs, 4 inputs each out of a pool of 8, 8 GPUs -> 8 outputs per GPU and all
    8 pool inputs replicated everywhere"""
    outs = {f"out_{k:02d}": [f"in_{(k + j) % 8}" for j in range(4)] for k in range(64)}
    plan = parallel.shard_outputs(outs, 8)
    assert [len(plan.outputs_of(r)) for r in range(8)] == [8] * 8
    assert sorted(plan.input_root) == [f"in_{i}" for i in range(8)]
    assert len(plan.broadcasts) == 8
    frame = 1920 * 1080 * 3 // 2
    assert parallel.broadcast_bytes(plan, {f"in_{i}": frame for i in range(8)}) == 8 * 7 * frame


def test_private_inputs_are_not_broadcast():
    outs = {"a": ["x", "s"], "b": ["y", "s"]}
    plan = parallel.shard_outputs(outs, 2)
    assert plan.output_rank == {"a": 0, "b": 1}
    assert plan.broadcasts == [("s", 0)]          # only the shared input crosses GPUs
    assert plan.rank_inputs == [["x", "s"], ["y", "s"]]
    one = parallel.shard_outputs(outs, 1)
    assert one.broadcasts == []


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        outs = {f"out_{k}": [f"in_{(k + j) % 3}" for j in range(2)] for k in range(4)}
        outs["out_9"] = ["private"]
        plan = parallel.shard_outputs(outs, world)
        # every rank ingests only the inputs it is root of; the rest arrive by broadcast
        frames = {}
        for i, root in plan.input_root.items():
            seed = sum(map(ord, i))
            frames[i] = torch.from_numpy(np.random.default_rng(seed).integers(0, 256, 4096, dtype=np.uint8)) \
                if root == rank else torch.zeros(4096, dtype=torch.uint8)
        for i, root in plan.broadcasts:   # the per-tick exchange step (ncclBroadcast group on GPUs)
            dist.broadcast(frames[i], src=root)
        ok = True
        for i in plan.rank_inputs[rank]:
            seed = sum(map(ord, i))
            exp = torch.from_numpy(np.random.default_rng(seed).integers(0, 256, 4096, dtype=np.uint8))
            ok = ok and bool(torch.equal(frames[i], exp))
        # a private input of another rank was never transferred here
        if "private" not in plan.rank_inputs[rank]:
            ok = ok and int(frames["private"].sum()) == 0
        # the union of the shards is every output exactly once
        mine = plan.outputs_of(rank)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        flat = sorted(o for g in gathered for o in g)
        ok = ok and flat == sorted(outs)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_replication_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
