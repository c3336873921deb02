"""Call-group batching around the native sampler (role of the reference's
cugraph_pyg/sampler/distributed_sampler.py:25-908, node path).

A "call group" is the set of mini-batches handed to ONE native call; its size is `local_seeds_per_call`
(default: the reference's memory heuristic, clamped to what one fused call accepts).  Each rank samples its own
seeds from its own (replicated) CSR, so no collective is issued while sampling; the only collectives are the
ones that keep uneven ranks in step (all_reduce of the call-group count), as in the reference (:301-329).
"""
from functools import reduce
from math import ceil
from typing import Dict, Iterator, List, Optional, Tuple, Union

import numpy as np
import torch

import pylibcugraph
from cugraph_pyg.utils import dist as dist_utils
from .sampler_utils import verify_metadata

TensorType = Union[torch.Tensor, np.ndarray, list]


class BaseDistributedSampler:
    def __init__(self, graph, local_seeds_per_call: int, retain_original_seeds: bool = False):
        self.__graph = graph
        self.__local_seeds_per_call = int(local_seeds_per_call)
        self.__handle = None
        self.__retain_original_seeds = retain_original_seeds

    def sample_batches(self, seeds, seed_times, batch_id_offsets, random_state: int = 0, metadata=None) -> Dict[str, torch.Tensor]:
        raise NotImplementedError("Must be implemented by subclass")

    @property
    def is_multi_gpu(self) -> bool:
        return isinstance(self.__graph, pylibcugraph.MGGraph)

    @property
    def _local_seeds_per_call(self) -> int:
        return self.__local_seeds_per_call

    @property
    def _graph(self):
        return self.__graph

    @property
    def _resource_handle(self):
        if self.__handle is None:
            self.__handle = pylibcugraph.ResourceHandle()
        return self.__handle

    @property
    def _retain_original_seeds(self) -> bool:
        return self.__retain_original_seeds

    def get_start_batch_offset(self, local_num_batches: int, assume_equal_input_size: bool = False) -> Tuple[int, bool]:
        """First global batch id of this rank, and whether every rank has the same number of batches."""
        if not self.is_multi_gpu:
            return 0, True
        return dist_utils.batch_id_start(local_num_batches, assume_equal_input_size)

    def _call_group(self, seeds: torch.Tensor, index: torch.Tensor, batch_id_start: int, batch_size: int,
                    random_state: int, metadata, times: Optional[torch.Tensor] = None) -> Tuple[Dict[str, torch.Tensor], int, int]:
        n = int(seeds.numel())
        num_full, last = divmod(n, batch_size)
        sizes = [batch_size] * num_full + ([last] if last else [])
        input_offsets = torch.tensor([0] + sizes, dtype=torch.int64).cumsum(0)
        out = self.sample_batches(seeds=seeds, seed_times=times, batch_id_offsets=input_offsets.cuda(non_blocking=True),
                                  random_state=random_state, metadata=metadata)
        out["input_index"] = index.cuda(non_blocking=True)
        out["input_offsets"] = input_offsets  # host: readers slice with python ints
        out["map"] = out.pop("renumber_map")
        out = {k: v for k, v in out.items() if v is not None}
        return out, batch_id_start, batch_id_start + len(sizes) - 1

    def sample_from_nodes(self, nodes: TensorType, *, batch_size: int = 16, random_state: int = 62,
                          assume_equal_input_size: bool = False, input_id: Optional[TensorType] = None,
                          input_time: Optional[TensorType] = None, metadata=None
                          ) -> Iterator[Tuple[Dict[str, torch.Tensor], int, int]]:
        """Lazily yields (raw call-group dict, first batch id, last batch id)."""
        verify_metadata(metadata)
        nodes = torch.as_tensor(nodes).cuda()
        num_seeds = int(nodes.numel())
        if input_time is not None:
            input_time = torch.as_tensor(input_time).reshape(-1).to(device=nodes.device, dtype=torch.int64)
            if input_time.numel() != num_seeds:
                raise ValueError("input_time must have one entry per input node")
        input_id = torch.arange(num_seeds, dtype=torch.int64) if input_id is None else torch.as_tensor(input_id).cpu()
        batches_per_call = max(1, self._local_seeds_per_call // batch_size)
        seeds_per_call = batches_per_call * batch_size
        local_num_batches = int(ceil(num_seeds / batch_size))
        batch_id_start, equal = self.get_start_batch_offset(local_num_batches, assume_equal_input_size)
        seed_groups = list(torch.split(nodes, seeds_per_call))
        index_groups = list(torch.split(input_id, seeds_per_call))
        time_groups = list(torch.split(input_time, seeds_per_call)) if input_time is not None else [None] * len(seed_groups)
        if self.is_multi_gpu and not equal:
            # every rank makes the same number of calls (uneven ranks sample empty groups)
            pad = dist_utils.equalized_call_count(len(seed_groups), equal) - len(seed_groups)
            seed_groups += [nodes[:0]] * pad
            index_groups += [input_id[:0]] * pad
            time_groups += [None if input_time is None else input_time[:0]] * pad

        def gen():
            start = batch_id_start
            for call_id, (s, ix, tm) in enumerate(zip(seed_groups, index_groups, time_groups)):
                raw, first, last = self._call_group(s, ix, start, batch_size, random_state + call_id, metadata, times=tm)
                start = last + 1
                yield raw, first, last

        return gen()

    def _edge_call_group(self, edges: torch.Tensor, index: torch.Tensor, label, batch_id_start: int, batch_size: int,
                         random_state: int, metadata, times: Optional[torch.Tensor] = None) -> Tuple[Dict[str, torch.Tensor], int, int]:
        """One call group of seed EDGES.  A batch's seeds are the endpoints [sources | destinations] of its edges;
        they are handed to the native sampler sorted inside each batch (ONE device sort per call group), which
        deduplicates them in first-occurrence (= ascending) order -- the order the reference produces with a python
        loop of sort + unique_consecutive per batch (distributed_sampler.py:487-533) -- and reports the local id of
        every input seed, i.e. `edge_inverse`."""
        n = int(edges.shape[1])
        num_full, last = divmod(n, batch_size)
        sizes = [batch_size] * num_full + ([last] if last else [])
        input_offsets = torch.tensor([0] + sizes, dtype=torch.int64).cumsum(0)
        dev = edges.device
        # per batch [src ... | dst ...]: position of endpoint (side, e) = 2 * first(batch) + side * size(batch) + (e - first(batch))
        e = torch.arange(n, device=dev)
        batch_of = e // batch_size
        first = batch_of * batch_size
        size_of = torch.full((n,), batch_size, device=dev, dtype=torch.int64)
        if last:
            size_of[num_full * batch_size:] = last
        pos_src = 2 * first + (e - first)
        seeds = torch.empty(2 * n, dtype=torch.int64, device=dev)
        seeds[pos_src] = edges[0]
        seeds[pos_src + size_of] = edges[1]
        seed_batch = torch.empty(2 * n, dtype=torch.int64, device=dev)
        seed_batch[pos_src] = batch_of
        seed_batch[pos_src + size_of] = batch_of
        span = int(seeds.max()) + 1 if n else 1
        order = torch.argsort(seed_batch * span + seeds, stable=True)  # batch-major, ascending id inside a batch
        seed_times = None
        if times is not None:
            # both endpoints of a seed edge start at the edge's time (reference :490-493); a vertex that is an endpoint of
            # several edges of the batch keeps the time of its first occurrence in the sorted seed list
            seed_times = torch.empty(2 * n, dtype=torch.int64, device=dev)
            seed_times[pos_src] = times
            seed_times[pos_src + size_of] = times
            seed_times = seed_times[order]
        out = self.sample_batches(seeds=seeds[order], seed_times=seed_times, batch_id_offsets=(2 * input_offsets).to(dev),
                                  random_state=random_state, metadata=metadata, return_seed_local_ids=True)
        inverse = torch.empty(2 * n, dtype=torch.int64, device=dev)
        inverse[order] = out.pop("seed_local_ids").to(torch.int64)
        out["edge_inverse"] = inverse  # 2 * batch entries per batch: local ids of the sources, then of the destinations
        out["input_index"] = index.to(dev)
        if label is not None:
            out["input_label"] = label.to(dev)
        out["input_offsets"] = input_offsets
        out["map"] = out.pop("renumber_map")
        out = {k: v for k, v in out.items() if v is not None}
        return out, batch_id_start, batch_id_start + len(sizes) - 1

    def sample_from_edges(self, edges: TensorType, *, batch_size: int = 16, random_state: int = 62,
                          assume_equal_input_size: bool = False, input_id: Optional[TensorType] = None,
                          input_time: Optional[TensorType] = None, input_label: Optional[TensorType] = None, metadata=None
                          ) -> Iterator[Tuple[Dict[str, torch.Tensor], int, int]]:
        """Sampling that starts from seed edges ([2, n], sources first): lazily yields (raw call-group dict, first batch
        id, last batch id).  Role of the reference's distributed_sampler.py:428-726."""
        verify_metadata(metadata)
        edges = torch.as_tensor(edges).cuda()
        n = int(edges.shape[-1])
        if input_time is not None:
            input_time = torch.as_tensor(input_time).reshape(-1).to(device=edges.device, dtype=torch.int64)
            if input_time.numel() != n:
                raise ValueError("input_time must have one entry per seed edge")
        input_id = torch.arange(n, dtype=torch.int64) if input_id is None else torch.as_tensor(input_id).cpu()
        label = None if input_label is None else torch.as_tensor(input_label)
        # every seed edge contributes TWO seed vertices to the native call: a call group holds local_seeds_per_call // 2
        # edges, so that its vertex count (and with it the per-hop edge bound of the native sampler) stays what the
        # node path gets
        batches_per_call = max(1, (self._local_seeds_per_call // 2) // batch_size)
        per_call = batches_per_call * batch_size
        local_num_batches = int(ceil(n / batch_size))
        batch_id_start, equal = self.get_start_batch_offset(local_num_batches, assume_equal_input_size)
        groups = [(edges[:, lo:lo + per_call], input_id[lo:lo + per_call], None if label is None else label[lo:lo + per_call],
                   None if input_time is None else input_time[lo:lo + per_call]) for lo in range(0, n, per_call)]
        if self.is_multi_gpu and not equal:
            pad = dist_utils.equalized_call_count(len(groups), equal) - len(groups)
            groups += [(edges[:, :0], input_id[:0], None if label is None else label[:0], None if input_time is None else input_time[:0])] * pad

        def gen():
            start = batch_id_start
            for call_id, (e, ix, lb, tm) in enumerate(groups):
                raw, first, last = self._edge_call_group(e, ix, lb, start, batch_size, random_state + call_id, metadata, times=tm)
                start = last + 1
                yield raw, first, last

        return gen()


class DistributedNeighborSampler(BaseDistributedSampler):
    BASE_VERTICES_PER_BYTE = 0.1107662486009992  # reference heuristic ("based on benchmarking", :755-757)
    UNKNOWN_VERTICES_DEFAULT = 32768
    MAX_EDGES_PER_HOP = (1 << 28) - 1  # one fused call packs (step, rank) references in 32 bits

    def __init__(self, graph, *, local_seeds_per_call: Optional[int] = None, retain_original_seeds: bool = False,
                 fanout: List[int] = [-1], prior_sources_behavior: str = "exclude", deduplicate_sources: bool = True,
                 compression: str = "COO", compress_per_hop: bool = False, with_replacement: bool = False,
                 disjoint: bool = False, biased: bool = False, heterogeneous: bool = False, temporal: bool = False,
                 temporal_comparison: Optional[str] = None, vertex_type_offsets=None, num_edge_types: int = 1):
        if num_edge_types > 1 and not heterogeneous:
            raise ValueError("Heterogeneous sampling must be selected if there is > 1 edge type.")
        self.__fanout = [int(f) for f in np.asarray(fanout).reshape(-1)]
        self.__heterogeneous = bool(heterogeneous)
        self.__num_edge_types = int(num_edge_types)
        self.__temporal = bool(temporal)
        table = {
            (False, False, False): pylibcugraph.homogeneous_uniform_neighbor_sample,
            (False, True, False): pylibcugraph.homogeneous_biased_neighbor_sample,
            (True, False, False): pylibcugraph.heterogeneous_uniform_neighbor_sample,
            (True, True, False): pylibcugraph.heterogeneous_biased_neighbor_sample,
            (False, False, True): pylibcugraph.homogeneous_uniform_temporal_neighbor_sample,
            (True, False, True): pylibcugraph.heterogeneous_uniform_temporal_neighbor_sample,
            (False, True, True): pylibcugraph.homogeneous_biased_temporal_neighbor_sample,
            (True, True, True): pylibcugraph.heterogeneous_biased_temporal_neighbor_sample,
        }
        self.__func = table[(self.__heterogeneous, bool(biased), self.__temporal)]
        self.__func_kwargs = {
            "h_fan_out": np.asarray(self.__fanout, dtype="int32"),
            "prior_sources_behavior": prior_sources_behavior,
            "retain_seeds": retain_original_seeds,
            "deduplicate_sources": deduplicate_sources,
            "compress_per_hop": compress_per_hop,
            "compression": compression,
            "with_replacement": with_replacement,
            "disjoint_sampling": disjoint,
        }
        if temporal:
            self.__func_kwargs["temporal_property_name"] = "time"
            self.__func_kwargs["temporal_sampling_comparison"] = temporal_comparison or "monotonically_decreasing"
        if heterogeneous:
            if vertex_type_offsets is None:
                raise ValueError("Heterogeneous sampling requires vertex type offsets.")
            if len(self.__fanout) % self.__num_edge_types != 0:
                raise ValueError(f"Illegal fanout for {num_edge_types} edge types.")
            self.__func_kwargs["num_edge_types"] = self.__num_edge_types
            self.__func_kwargs["vertex_type_offsets"] = torch.as_tensor(vertex_type_offsets).cpu()
        super().__init__(graph, self.__calc_local_seeds_per_call(local_seeds_per_call), retain_original_seeds)

    def __calc_local_seeds_per_call(self, local_seeds_per_call: Optional[int]) -> int:
        if local_seeds_per_call is not None:
            return int(local_seeds_per_call)
        fanout = self.__fanout
        if self.__heterogeneous:
            # loaders lay the vector out [hop * T + etype] (neighbor_loader.py:192-201); the bound is the per-hop sum
            T = self.__num_edge_types
            fanout = [sum(fanout[h * T + t] for t in range(T)) if all(fanout[h * T + t] >= 0 for t in range(T)) else -1
                      for h in range(len(fanout) // T)]
        if any(f <= 0 for f in fanout):
            return DistributedNeighborSampler.UNKNOWN_VERTICES_DEFAULT
        total_memory = torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory
        prod = reduce(lambda x, y: x * y, fanout)
        by_memory = int(DistributedNeighborSampler.BASE_VERTICES_PER_BYTE * total_memory / prod)
        return max(1, min(by_memory, DistributedNeighborSampler.MAX_EDGES_PER_HOP // prod))

    def sample_batches(self, seeds, seed_times, batch_id_offsets, random_state: int = 0, metadata=None,
                       return_seed_local_ids: bool = False) -> Dict[str, torch.Tensor]:
        rank = torch.distributed.get_rank() if (self.is_multi_gpu and torch.distributed.is_initialized()) else 0
        kwargs = {
            "resource_handle": self._resource_handle,
            "input_graph": self._graph,
            "start_vertex_list": seeds,
            "starting_vertex_label_offsets": batch_id_offsets,
            "renumber": True,
            "return_hops": True,
            "do_expensive_check": False,
            "random_state": random_state + rank,
        }
        kwargs.update(self.__func_kwargs)
        if return_seed_local_ids:
            kwargs["return_seed_local_ids"] = True
        if seed_times is not None:
            if not self.__temporal:
                raise ValueError("seed times were given to a sampler that was built with temporal=False")
            kwargs["starting_vertex_times"] = seed_times
        out = self.__func(**kwargs)
        out["fanout"] = torch.tensor(self.__fanout, dtype=torch.int32)
        out["rank"] = rank
        if metadata is not None:
            out.update(metadata)
        return out
