"""Small host-side collectives the stores and samplers need (counts, maxima, batch-id offsets).

They run on whatever device the process group's backend communicates with -- CUDA tensors under NCCL, CPU tensors
under gloo -- so the multi-rank bookkeeping can be exercised by world_size-2 gloo tests without a GPU
(tests/test_dist_cpu.py)."""
from typing import List, Sequence, Tuple

import torch


def initialized() -> bool:
    return torch.distributed.is_available() and torch.distributed.is_initialized()


def world() -> Tuple[int, int]:
    """(rank, world size); (0, 1) without a process group."""
    if not initialized():
        return 0, 1
    return torch.distributed.get_rank(), torch.distributed.get_world_size()


def coll_device() -> torch.device:
    if initialized() and torch.distributed.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_ints(values: Sequence[int]) -> List[List[int]]:
    """Every rank contributes `values` (same length everywhere); returns [world][len(values)]."""
    rank, size = world()
    if size == 1:
        return [list(int(v) for v in values)]
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=coll_device())
    out = torch.empty(size * t.numel(), dtype=torch.int64, device=t.device)  # flat: gloo rejects a 2-D output
    torch.distributed.all_gather_into_tensor(out, t)
    return out.view(size, t.numel()).tolist()


def all_reduce_max(value: int) -> int:
    if world()[1] == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=coll_device())
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return int(t)


def batch_id_start(local_num_batches: int, assume_equal_input_size: bool = False) -> Tuple[int, bool]:
    """First global batch id of this rank and whether every rank holds the same number of batches
    (reference: cugraph_pyg/sampler/distributed_sampler.py:301-329)."""
    rank, size = world()
    if size == 1:
        return 0, True
    if assume_equal_input_size:
        return rank * int(local_num_batches), True
    counts = [c[0] for c in all_gather_ints([local_num_batches])]
    return int(sum(counts[:rank])), len(set(counts)) == 1


def equalized_call_count(local_calls: int, equal: bool) -> int:
    """Number of sampling calls every rank makes: ranks with fewer call groups pad with empty ones so that collectives
    issued per call stay in step (reference: distributed_sampler.py:180-214)."""
    return int(local_calls) if equal else all_reduce_max(local_calls)


def edge_id_starts(local_counts: Sequence[int]) -> List[int]:
    """Per edge type: first edge id of this rank's partition (ids run over the ranks in rank order,
    reference: cugraph_pyg/data/graph_store.py:578-607)."""
    rank, size = world()
    if size == 1:
        return [0] * len(local_counts)
    table = all_gather_ints(local_counts)
    return [int(sum(table[r][k] for r in range(rank))) for k in range(len(local_counts))]
