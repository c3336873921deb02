"""FeatureStore (role of the reference's cugraph_pyg/data/feature_store.py:24-239).

PyG FeatureStore interface over WholeMemory: every rank puts ITS slice of a feature, any rank reads any rows.
2-D features become DistEmbedding tables (row-sharded, read by in-kernel P2P), 1-D features DistTensors.
"""
from typing import Dict, List, Optional, Tuple

import torch

from cugraph_pyg._pyg_compat import FeatureStoreBase, TensorAttr
from cugraph_pyg.tensor import DistEmbedding, DistTensor, is_empty


def _world():
    d = torch.distributed
    if d.is_available() and d.is_initialized():
        return d.get_world_size(), d.get_rank()
    return 1, 0


def _all_gather_scalar(v: int) -> torch.Tensor:
    world, _ = _world()
    t = torch.tensor([v], dtype=torch.int64, device="cuda")
    if world == 1:
        return t
    out = torch.empty((world,), dtype=torch.int64, device="cuda")
    torch.distributed.all_gather_into_tensor(out, t)
    return out


_DTYPES = [torch.float32, torch.float64, torch.int32, torch.int64, torch.int16, torch.float16, torch.int8, torch.bfloat16]


class FeatureStore(FeatureStoreBase):
    def __init__(self, memory_type=None, location: str = "cpu"):
        super().__init__()
        self.__features: Dict[Tuple, DistTensor] = {}
        self.__location = location
        self.__backend = "vmm"

    def __make_wg_tensor(self, tensor: torch.Tensor, ix: Optional[torch.Tensor] = None):
        world, rank = _world()
        dims = _all_gather_scalar(tensor.dim())
        if not bool((dims == dims[0]).all()):
            raise ValueError("Tensor dimension must be the same across ranks")
        sizes = _all_gather_scalar(0 if is_empty(tensor) else tensor.shape[0])
        dtype_ids = _all_gather_scalar(_DTYPES.index(tensor.dtype))[sizes > 0]
        if dtype_ids.numel() == 0:
            raise ValueError("Tensor is empty")
        if not bool((dtype_ids == dtype_ids[0]).all()):
            raise ValueError("Tensor dtype must be the same across ranks")
        dtype = _DTYPES[int(dtype_ids[0])]
        total = int(sizes.sum())
        if tensor.dim() == 1:
            tx = DistTensor(None, shape=[total], dtype=dtype, device=self.__location, backend=self.__backend,
                            partition_book=[int(s) for s in sizes] if ix is None else None)
        elif tensor.dim() == 2:
            widths = _all_gather_scalar(-1 if is_empty(tensor) else tensor.shape[1])
            widths = widths[widths > 0]
            if widths.numel() == 0:
                raise ValueError("Tensor is empty")
            if not bool((widths == widths[0]).all()):
                raise ValueError("Trailing dimensions must be the same across ranks")
            width = int(widths[0])
            # each rank keeps exactly the rows it put: the partition follows the puts, no data moves between GPUs
            tx = DistEmbedding(None, shape=[total, width], dtype=dtype, device=self.__location, backend=self.__backend,
                               partition_book=[int(s) for s in sizes] if ix is None else None)
            if is_empty(tensor):
                tensor = tensor.reshape((-1, width))
        else:
            raise ValueError("Tensor must be 1D or 2D.")
        if ix is None:
            if tensor.shape[0]:
                tx.get_local_tensor().copy_(tensor.to(dtype).cuda())
            torch.cuda.current_stream().synchronize()
            tx.get_comm().barrier()
        else:
            if tensor.shape[0] != ix.shape[0]:
                raise ValueError("Shape mismatch")
            if ix.dim() != 1:
                raise ValueError("Index must be 1D")
            tx[ix] = tensor
            torch.cuda.current_stream().synchronize()
            tx.get_comm().barrier()
        return tx

    def replicate_hot_rows(self, graph_store, ratio: float = 0.1) -> Dict[Tuple, int]:
        """B200 extension: on every GPU, keep a local copy of the `ratio` most frequently reachable rows of each 2-D node
        feature (those of the vertices that are the PyG *source* of the most edges, i.e. the ones neighbour sampling
        returns most often), so that most feature reads stop crossing NVLink (role of the reference's device cache of a
        remote table; here the set is static).  Collective when the stores are multi-GPU (degree counts are summed over
        the ranks).  Call after the features and edges have been put; returns {(group, attr): rows replicated}."""
        world, _ = _world()
        counts: Dict[str, torch.Tensor] = {}
        nv = graph_store._num_vertices()
        for attr in graph_store.get_all_edge_attrs():
            src_type = attr.edge_type[0]
            row, _ = graph_store.get_edge_index(attr.edge_type, "coo")
            c = torch.bincount(row.cuda().long(), minlength=nv[src_type])
            counts[src_type] = counts[src_type] + c if src_type in counts else c
        done = {}
        for (group, name), tx in self.__features.items():
            if isinstance(group, tuple) or not isinstance(tx, DistEmbedding) or group not in counts:
                continue
            c = counts[group]
            if world > 1:
                c = c.clone()
                torch.distributed.all_reduce(c)
            k = min(int(ratio * tx.shape[0]), int((c > 0).sum()))
            if k <= 0:
                continue
            tx.set_hot_rows(torch.topk(c, k).indices)
            done[(group, name)] = k
        return done

    def _put_tensor(self, tensor, attr: TensorAttr) -> bool:
        key = (attr.group_name, attr.attr_name)
        if attr.is_set("index") and attr.index is not None:
            if key not in self.__features:
                self.__features[key] = self.__make_wg_tensor(tensor, ix=attr.index)
            else:
                self.__features[key][attr.index] = tensor
        else:
            self.__features[key] = self.__make_wg_tensor(tensor)
        return True

    def _get_tensor(self, attr: TensorAttr):
        key = (attr.group_name, attr.attr_name)
        if key not in self.__features:
            return None
        emb = self.__features[key]
        if attr.is_set("index") and attr.index is not None:
            return emb[attr.index]
        return emb

    def _remove_tensor(self, attr: TensorAttr) -> bool:
        key = (attr.group_name, attr.attr_name)
        if key not in self.__features:
            return False
        del self.__features[key]
        return True

    def _get_tensor_size(self, attr: TensorAttr) -> Tuple:
        return self.__features[attr.group_name, attr.attr_name].shape

    def get_all_tensor_attrs(self) -> List[TensorAttr]:
        return [TensorAttr(group_name=g, attr_name=a) for g, a in self.__features.keys()]
