"""NodeLoader (role of the reference's cugraph_pyg/loader/node_loader.py:14-178)."""
import warnings
from typing import Callable, Optional

import torch

import cugraph_pyg
from cugraph_pyg._pyg_compat import NodeSamplerInput, get_input_nodes
from .utils import generate_seed


class NodeLoader:
    """Iterable over mini-batches: seeds are (optionally shuffled and) split into batches, sampled in call groups
    by the given BaseSampler, and joined with their features by SampleIterator."""

    def __init__(self, data, node_sampler, input_nodes=None, input_time=None, transform: Optional[Callable] = None,
                 transform_sampler_output: Optional[Callable] = None, filter_per_worker: Optional[bool] = None,
                 custom_cls=None, input_id=None, batch_size: int = 1, shuffle: bool = False, drop_last: bool = False,
                 **kwargs):
        if not isinstance(data, (list, tuple)) or not isinstance(data[1], cugraph_pyg.data.GraphStore):
            raise NotImplementedError("Currently can't accept non-cugraph graphs")
        if not isinstance(node_sampler, cugraph_pyg.sampler.BaseSampler):
            raise NotImplementedError("Must provide a cuGraph sampler")
        for name, value in (("filter_per_worker", filter_per_worker), ("custom_cls", custom_cls), ("transform", transform),
                            ("transform_sampler_output", transform_sampler_output)):
            if value:
                warnings.warn(f"{name} is currently ignored")
        explicit = not (input_nodes is None or isinstance(input_nodes, str))
        if isinstance(input_nodes, (list, tuple)) and len(input_nodes) == 2 and isinstance(input_nodes[0], str):
            explicit = input_nodes[1] is not None
        self.__has_explicit_input_nodes = explicit
        input_type, input_nodes, input_id = get_input_nodes(data, input_nodes, input_id)
        input_nodes = input_nodes.detach().clone()
        if input_nodes.numel() < batch_size and drop_last:
            raise ValueError("The number of input nodes is less than the batch size and drop_last is True. "
                             "This will result in all batches being dropped.")
        if input_type is not None:
            input_nodes += data[1]._vertex_offsets[input_type]
        self.__input_data = NodeSamplerInput(
            input_id=torch.arange(len(input_nodes), dtype=torch.int64) if input_id is None else input_id,
            node=input_nodes, time=None if input_time is None else torch.as_tensor(input_time).reshape(-1),
            input_type=input_type)
        if self.__input_data.time is not None and self.__input_data.time.numel() != input_nodes.numel():
            raise ValueError("input_time must have one entry per input node")
        self.__data = data
        self.__node_sampler = node_sampler
        self.__batch_size = batch_size
        self.__shuffle = shuffle
        self.__drop_last = drop_last

    def __iter__(self):
        n = self.__input_data.node.numel()
        perm = torch.randperm(n) if self.__shuffle else torch.arange(n)
        if self.__drop_last and n % self.__batch_size:
            perm = perm[: n - n % self.__batch_size]
        node = self.__input_data.node
        perm_dev = perm.to(node.device)
        input_data = NodeSamplerInput(input_id=self.__input_data.input_id[perm.to(self.__input_data.input_id.device)],
                                      node=node[perm_dev],
                                      time=None if self.__input_data.time is None else self.__input_data.time[perm.to(self.__input_data.time.device)],
                                      input_type=self.__input_data.input_type)
        return cugraph_pyg.sampler.SampleIterator(
            self.__data, self.__node_sampler.sample_from_nodes(input_data, random_state=generate_seed()))

    def __len__(self):
        if not self.__has_explicit_input_nodes:
            raise ValueError("len(loader) is only supported when the loader was constructed with an explicit "
                             "number of seeds via input_nodes for now.")
        n = self.__input_data.node.numel()
        return n // self.__batch_size if self.__drop_last else (n + self.__batch_size - 1) // self.__batch_size
