"""LinkLoader (role of the reference's cugraph_pyg/loader/link_loader.py:17-234)."""
import warnings
from typing import Callable, Optional

import torch

import cugraph_pyg
from cugraph_pyg._pyg_compat import EdgeSamplerInput, NegativeSampling, get_edge_label_index
from .utils import generate_seed


class LinkLoader:
    """Iterable over mini-batches of seed EDGES (link prediction): edges are (optionally shuffled and) split into
    batches, their endpoints sampled in call groups by the given BaseSampler, negatives added when asked for."""

    def __init__(self, data, link_sampler, edge_label_index=None, edge_label=None, edge_label_time=None, neg_sampling=None,
                 neg_sampling_ratio=None, transform: Optional[Callable] = None, transform_sampler_output: Optional[Callable] = None,
                 filter_per_worker: Optional[bool] = None, custom_cls=None, input_id=None, batch_size: int = 1,
                 shuffle: bool = False, drop_last: bool = False, **kwargs):
        if not isinstance(data, (list, tuple)) or not isinstance(data[1], cugraph_pyg.data.GraphStore):
            raise NotImplementedError("Currently can't accept non-cugraph graphs")
        if not isinstance(link_sampler, cugraph_pyg.sampler.BaseSampler):
            raise NotImplementedError("Must provide a cuGraph sampler")
        for name, value in (("filter_per_worker", filter_per_worker), ("custom_cls", custom_cls), ("transform", transform),
                            ("transform_sampler_output", transform_sampler_output)):
            if value:
                warnings.warn(f"{name} is currently ignored")
        if neg_sampling_ratio is not None:
            warnings.warn("The 'neg_sampling_ratio' argument is deprecated in PyG and is not supported in cuGraph-PyG.")
        neg_sampling = NegativeSampling.cast(neg_sampling)
        if edge_label_time is not None:
            if neg_sampling is not None:
                raise NotImplementedError("temporal negative sampling is not implemented (DESIGN.md §10)")
            edge_label_time = torch.as_tensor(edge_label_time).reshape(-1)
        explicit = edge_label_index is not None
        if isinstance(edge_label_index, (list, tuple)):
            if len(edge_label_index) == 3 and all(isinstance(v, str) for v in edge_label_index):
                explicit = False
            elif len(edge_label_index) == 2 and not torch.is_tensor(edge_label_index[0]):
                explicit = edge_label_index[1] is not None
        self.__has_explicit_edge_label_index = explicit
        input_type, edge_label_index = get_edge_label_index(data, edge_label_index)
        edge_label_index = edge_label_index.detach().clone()
        if edge_label_index.shape[1] < batch_size and drop_last:
            raise ValueError("The number of input edges is less than the batch size and drop_last is True. This will result "
                             "in all batches being dropped. Either set drop_last to False or increase the number of edges "
                             "in edge_label_index.")
        if input_type is not None:  # note the reverse of the usual convention: row = source, col = destination
            edge_label_index[0] += data[1]._vertex_offsets[input_type[0]]
            edge_label_index[1] += data[1]._vertex_offsets[input_type[2]]
        if neg_sampling is not None and neg_sampling.is_binary() and edge_label is not None and edge_label.min() == 0:
            edge_label = edge_label + 1
        if neg_sampling is not None and neg_sampling.is_triplet() and edge_label is not None:
            raise ValueError("'edge_label' needs to be undefined for 'triplet'-based negative sampling. Please use `src_index`, "
                             "`dst_pos_index` and `neg_pos_index` of the returned mini-batch instead to differentiate between "
                             "positive and negative samples.")
        self.__input_data = EdgeSamplerInput(
            input_id=torch.arange(edge_label_index[0].numel(), dtype=torch.int64) if input_id is None else input_id,
            row=edge_label_index[0], col=edge_label_index[1], label=edge_label, time=edge_label_time, input_type=input_type)
        if edge_label_time is not None and edge_label_time.numel() != edge_label_index.shape[1]:
            raise ValueError("edge_label_time must have one entry per seed edge")
        self.__data = data
        self.__link_sampler = link_sampler
        self.__neg_sampling = neg_sampling
        self.__batch_size = batch_size
        self.__shuffle = shuffle
        self.__drop_last = drop_last

    def __iter__(self):
        n = self.__input_data.row.numel()
        perm = torch.randperm(n) if self.__shuffle else torch.arange(n)
        if self.__drop_last and n % self.__batch_size:
            perm = perm[: n - n % self.__batch_size]
        d = self.__input_data
        input_data = EdgeSamplerInput(input_id=d.input_id[perm.to(d.input_id.device)], row=d.row[perm.to(d.row.device)],
                                      col=d.col[perm.to(d.col.device)],
                                      label=None if d.label is None else d.label[perm.to(d.label.device)],
                                      time=None if d.time is None else d.time[perm.to(d.time.device)], input_type=d.input_type)
        return cugraph_pyg.sampler.SampleIterator(
            self.__data, self.__link_sampler.sample_from_edges(input_data, neg_sampling=self.__neg_sampling,
                                                               random_state=generate_seed()))

    def __len__(self):
        if not self.__has_explicit_edge_label_index:
            raise ValueError("len(loader) is only supported when the loader was constructed with an explicit number of seeds "
                             "via edge_label_index for now.")
        n = self.__input_data.row.numel()
        return n // self.__batch_size if self.__drop_last else (n + self.__batch_size - 1) // self.__batch_size
