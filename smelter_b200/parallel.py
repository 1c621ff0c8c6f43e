"""Output sharding and shared-input replication plan for a multi-GPU box.

The reference is single-device: one `Renderer`, outputs rendered in a serial loop over read-only shared
inputs (smelter-render/src/state/render_loop.rs:232-236).  Outputs are independent, so they shard across
GPUs with NO data-path collective; the only exchange is replicating an input frame to every GPU that hosts
an output referencing it.  This module is the host-side plan (pure Python, backend-agnostic); the transfers
are `smr_comm_broadcast_inputs` (NCCL over NVLink) on GPUs, or any torch.distributed backend in tests.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple


@dataclass
class ShardPlan:
    n_ranks: int
    output_rank: Dict[str, int]                       # output id -> rank that composites it
    input_root: Dict[str, int]                        # input id -> rank that ingests (decodes / uploads) it
    rank_inputs: List[List[str]] = field(default_factory=list)    # inputs each rank needs resident
    broadcasts: List[Tuple[str, int]] = field(default_factory=list)  # (input id, root): inputs needed off-root

    def outputs_of(self, rank: int) -> List[str]:
        return [o for o, r in self.output_rank.items() if r == rank]

    def needs(self, rank: int, input_id: str) -> bool:
        return input_id in self.rank_inputs[rank]


def shard_outputs(output_inputs: Dict[str, Sequence[str]], n_ranks: int, input_root: Dict[str, int] = None) -> ShardPlan:
    """Static assignment: outputs in id order, balanced by count (rebalanced on scene change by calling
    again); an input's root defaults to the rank hosting the first output that uses it.

    output_inputs: output id -> the input ids its scene references."""
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    outs = sorted(output_inputs)
    per = -(-len(outs) // n_ranks) if outs else 0
    output_rank = {o: (i // per if per else 0) for i, o in enumerate(outs)}
    rank_inputs: List[List[str]] = [[] for _ in range(n_ranks)]
    roots: Dict[str, int] = dict(input_root or {})
    for o in outs:
        r = output_rank[o]
        for i in output_inputs[o]:
            if i not in rank_inputs[r]:
                rank_inputs[r].append(i)
            roots.setdefault(i, r)
    for i, r in roots.items():
        if not (0 <= r < n_ranks):
            raise ValueError(f"root of {i} out of range")
    broadcasts = []
    for i in sorted(roots):
        users = [r for r in range(n_ranks) if i in rank_inputs[r]]
        if any(r != roots[i] for r in users):
            broadcasts.append((i, roots[i]))
    return ShardPlan(n_ranks, output_rank, roots, rank_inputs, broadcasts)


def broadcast_bytes(plan: ShardPlan, frame_bytes: Dict[str, int]) -> int:
    """bytes that cross NVLink per tick under `plan` (each broadcast reaches n_ranks - 1 peers at most;
    NCCL replicates to the whole communicator, so every non-root rank receives the frame)."""
    return sum(frame_bytes[i] * (plan.n_ranks - 1) for i, _ in plan.broadcasts)
