"""Call-group sizing of the distributed sampler (host logic, no GPU): a seed EDGE contributes two seed vertices to the native
call, so an edge call group holds local_seeds_per_call // 2 edges -- the vertex count of a native call (and with it the
sampler's per-hop edge bound, 2^28) is the same on the node path and on the link path (distributed_sampler.py)."""
import pytest
import torch


@pytest.fixture
def sampler_cls(monkeypatch):
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    from cugraph_pyg.sampler.distributed_sampler import BaseDistributedSampler

    class Recorder(BaseDistributedSampler):
        def __init__(self, per_call):
            super().__init__(object(), per_call)
            self.calls = []

        def sample_batches(self, seeds, seed_times, batch_id_offsets, random_state=0, metadata=None, return_seed_local_ids=False):
            self.calls.append((int(seeds.numel()), batch_id_offsets.clone()))
            n = int(seeds.numel())
            out = {"renumber_map": seeds.clone(), "renumber_map_offsets": batch_id_offsets.clone()}
            if return_seed_local_ids:
                out["seed_local_ids"] = torch.zeros(n, dtype=torch.int32)
            return out

    return Recorder


def test_edge_call_groups_hold_half_as_many_edges_as_node_call_groups_hold_nodes(sampler_cls):
    per_call, batch = 64, 8
    s = sampler_cls(per_call)
    list(s.sample_from_nodes(torch.arange(200), batch_size=batch))
    assert max(n for n, _ in s.calls) == per_call
    s = sampler_cls(per_call)
    edges = torch.stack([torch.arange(200), torch.arange(200) + 1000])
    list(s.sample_from_edges(edges, batch_size=batch))
    # 2 seed vertices per edge: no native call sees more than local_seeds_per_call vertices
    assert max(n for n, _ in s.calls) == per_call
    assert sum(n for n, _ in s.calls) == 2 * 200
    for n, off in s.calls:
        assert int(off[-1]) == n and ((off[1:] - off[:-1]) <= 2 * batch).all()


def test_edge_call_group_never_smaller_than_one_batch(sampler_cls):
    s = sampler_cls(4)  # smaller than one batch of edges
    edges = torch.stack([torch.arange(20), torch.arange(20) + 100])
    list(s.sample_from_edges(edges, batch_size=8))
    assert [n for n, _ in s.calls] == [16, 16, 8]
