"""Minimal stand-ins for the torch_geometric symbols this package executes at run time.

torch_geometric is optional: when it is installed the real classes are used,
otherwise these duck-typed equivalents provide the same attribute names and the store "sugar"
(``store[group, attr, index] = tensor`` -> ``put_tensor`` -> ``_put_tensor`` ...).
"""
import enum
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import torch

try:  # pragma: no cover - not installed in the build container
    import torch_geometric  # noqa: F401

    HAS_PYG = True
except Exception:  # ModuleNotFoundError or a broken install
    HAS_PYG = False

if HAS_PYG:  # pragma: no cover
    from torch_geometric.data import Data, HeteroData, FeatureStore as FeatureStoreBase, GraphStore as GraphStoreBase
    from torch_geometric.data.feature_store import TensorAttr
    from torch_geometric.data.graph_store import EdgeAttr, EdgeLayout
    from torch_geometric.sampler import NodeSamplerInput, EdgeSamplerInput, NegativeSampling, SamplerOutput, HeteroSamplerOutput
    from torch_geometric.edge_index import ptr2index
else:

    class _Unset(enum.Enum):
        UNSET = 0

    UNSET = _Unset.UNSET

    class EdgeLayout(enum.Enum):
        COO = "coo"
        CSC = "csc"
        CSR = "csr"

    @dataclass
    class TensorAttr:
        group_name: Any = UNSET
        attr_name: Any = UNSET
        index: Any = UNSET

        def is_set(self, key: str) -> bool:
            return getattr(self, key) is not UNSET

        def is_fully_specified(self) -> bool:
            return all(self.is_set(k) for k in ("group_name", "attr_name", "index"))

        def fully_specify(self):
            for k in ("group_name", "attr_name", "index"):
                if not self.is_set(k):
                    setattr(self, k, None)
            return self

    @dataclass
    class EdgeAttr:
        edge_type: Any
        layout: Any
        is_sorted: bool = False
        size: Optional[Tuple[int, int]] = None

        def __post_init__(self):
            if isinstance(self.layout, str):
                self.layout = EdgeLayout(self.layout)

    class FeatureStoreBase:
        """put/get/remove sugar of torch_geometric.data.FeatureStore."""

        def _attr(self, *args, **kwargs):
            if len(args) == 1 and isinstance(args[0], TensorAttr):
                return args[0]
            if len(args) == 1 and isinstance(args[0], (tuple, list)):
                args = tuple(args[0])
            return TensorAttr(*args, **kwargs)

        def put_tensor(self, tensor, *args, **kwargs) -> bool:
            attr = self._attr(*args, **kwargs)
            if not attr.is_set("index"):
                attr.index = None
            return self._put_tensor(tensor, attr)

        def get_tensor(self, *args, **kwargs):
            attr = self._attr(*args, **kwargs)
            if not attr.is_set("index"):
                attr.index = None
            out = self._get_tensor(attr)
            if out is None:
                raise KeyError(f"A tensor corresponding to '{attr}' was not found")
            return out

        def multi_get_tensor(self, attrs):
            return [self.get_tensor(a) for a in attrs]

        def remove_tensor(self, *args, **kwargs) -> bool:
            return self._remove_tensor(self._attr(*args, **kwargs))

        def get_tensor_size(self, *args, **kwargs):
            return self._get_tensor_size(self._attr(*args, **kwargs))

        def update_tensor(self, tensor, *args, **kwargs) -> bool:
            attr = self._attr(*args, **kwargs)
            self.remove_tensor(attr)
            return self.put_tensor(tensor, attr)

        def __setitem__(self, key, value):
            self.put_tensor(value, self._attr(key))

        def __getitem__(self, key):
            return self.get_tensor(self._attr(key))

        def __delitem__(self, key):
            self.remove_tensor(self._attr(key))

    class GraphStoreBase:
        """put/get/remove sugar of torch_geometric.data.GraphStore."""

        def _eattr(self, *args, **kwargs):
            if len(args) == 1 and isinstance(args[0], EdgeAttr):
                return args[0]
            if len(args) == 1 and isinstance(args[0], (tuple, list)) and not (len(args[0]) == 3 and all(isinstance(x, str) for x in args[0])):
                args = tuple(args[0])
            return EdgeAttr(*args, **kwargs)

        def put_edge_index(self, edge_index, *args, **kwargs) -> bool:
            return self._put_edge_index(edge_index, self._eattr(*args, **kwargs))

        def get_edge_index(self, *args, **kwargs):
            out = self._get_edge_index(self._eattr(*args, **kwargs))
            if out is None:
                raise KeyError("edge index not found")
            return out

        def remove_edge_index(self, *args, **kwargs) -> bool:
            return self._remove_edge_index(self._eattr(*args, **kwargs))

        def __setitem__(self, key, value):
            self.put_edge_index(value, self._eattr(key))

        def __getitem__(self, key):
            return self.get_edge_index(self._eattr(key))

        def __delitem__(self, key):
            self.remove_edge_index(self._eattr(key))

    class Data:
        """Attribute bag with torch_geometric.data.Data's basic protocol."""

        def __init__(self, **kwargs):
            object.__setattr__(self, "_store", dict(kwargs))

        def __getattr__(self, k):
            store = object.__getattribute__(self, "_store")
            if k in store:
                return store[k]
            raise AttributeError(k)

        def __setattr__(self, k, v):
            self._store[k] = v

        def __getitem__(self, k):
            return self._store[k]

        def __setitem__(self, k, v):
            self._store[k] = v

        def __contains__(self, k):
            return k in self._store

        def keys(self):
            return list(self._store.keys())

        @property
        def num_nodes(self):
            if "n_id" in self._store:
                return int(self._store["n_id"].numel())
            if "x" in self._store:
                return int(self._store["x"].shape[0])
            return None

        def __repr__(self):
            parts = []
            for k, v in self._store.items():
                parts.append(f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={v}")
            return "Data(" + ", ".join(parts) + ")"

    class HeteroData:
        """Per-type attribute bags with torch_geometric.data.HeteroData's basic protocol: ``data[node_type]`` /
        ``data[(src, rel, dst)]`` return (creating on first use) the storage of that type; ``set_value_dict`` spreads a
        {type: value} dict over the storages."""

        def __init__(self):
            object.__setattr__(self, "_stores", {})

        def __getitem__(self, key):
            if isinstance(key, list):
                key = tuple(key)
            stores = object.__getattribute__(self, "_stores")
            if key not in stores:
                stores[key] = Data()
            return stores[key]

        def __contains__(self, key):
            return key in self._stores

        @property
        def node_types(self):
            return [k for k in self._stores if isinstance(k, str)]

        @property
        def edge_types(self):
            return [k for k in self._stores if isinstance(k, tuple)]

        def set_value_dict(self, key, value_dict):
            for k, v in (value_dict or {}).items():
                self[k][key] = v
            return self

        def __repr__(self):
            return "HeteroData(" + ", ".join(f"{k}={v!r}" for k, v in self._stores.items()) + ")"

    @dataclass
    class NodeSamplerInput:
        input_id: Any
        node: Any
        time: Any = None
        input_type: Any = None

    @dataclass
    class EdgeSamplerInput:
        input_id: Any
        row: Any
        col: Any
        label: Any = None
        time: Any = None
        input_type: Any = None

    class NegativeSampling:
        """torch_geometric.sampler.NegativeSampling: mode 'binary' | 'triplet', amount = negatives per positive."""

        def __init__(self, mode, amount=1, src_weight=None, dst_weight=None):
            mode = getattr(mode, "value", mode)
            if mode not in ("binary", "triplet"):
                raise ValueError(f"unknown negative sampling mode '{mode}'")
            if amount <= 0:
                raise ValueError(f"The attribute 'amount' needs to be positive (got {amount})")
            if mode == "triplet" and amount != int(amount):
                raise ValueError("'amount' needs to be an integer for triplet negative sampling")
            self.mode, self.amount, self.src_weight, self.dst_weight = mode, amount, src_weight, dst_weight

        def is_binary(self) -> bool:
            return self.mode == "binary"

        def is_triplet(self) -> bool:
            return self.mode == "triplet"

        @classmethod
        def cast(cls, value):
            if value is None or isinstance(value, cls):
                return value
            if isinstance(value, str):
                return cls(value)
            if isinstance(value, (tuple, list)):
                return cls(*value)
            if isinstance(value, dict):
                return cls(**value)
            raise ValueError(f"cannot cast {value!r} to NegativeSampling")

    @dataclass
    class SamplerOutput:
        node: Any
        row: Any
        col: Any
        edge: Any
        batch: Any = None
        num_sampled_nodes: Any = None
        num_sampled_edges: Any = None
        orig_row: Any = None
        orig_col: Any = None
        metadata: Any = None

    @dataclass
    class HeteroSamplerOutput:
        node: Dict[str, Any]
        row: Dict[Any, Any]
        col: Dict[Any, Any]
        edge: Dict[Any, Any]
        batch: Any = None
        num_sampled_nodes: Any = None
        num_sampled_edges: Any = None
        orig_row: Any = None
        orig_col: Any = None
        metadata: Any = None

    def ptr2index(ptr: torch.Tensor, output_size: Optional[int] = None) -> torch.Tensor:
        """CSR pointer -> row index per nonzero (torch_geometric.edge_index.ptr2index)."""
        index = torch.arange(ptr.numel() - 1, dtype=ptr.dtype, device=ptr.device)
        return index.repeat_interleave(ptr.diff(), output_size=output_size)


def get_input_nodes(data, input_nodes, input_id=None):
    """(node_type | None, node ids, input_id) -- torch_geometric.loader.utils.get_input_nodes for store tuples."""
    node_type = None
    if isinstance(input_nodes, (tuple, list)) and len(input_nodes) == 2 and isinstance(input_nodes[0], str):
        node_type, input_nodes = input_nodes
    elif isinstance(input_nodes, str):
        node_type, input_nodes = input_nodes, None
    if input_nodes is None:
        feature_store, graph_store = data
        n = graph_store._num_vertices()[node_type if node_type is not None else "_N"]
        input_nodes = torch.arange(n)
    if input_nodes.dtype == torch.bool:
        input_nodes = input_nodes.nonzero().view(-1)
        if input_id is not None:
            input_id = input_id[input_nodes.cpu()] if input_id.numel() != input_nodes.numel() else input_id
    return node_type, input_nodes, input_id


def get_edge_label_index(data, edge_label_index):
    """(edge_type | None, [2, n] tensor) -- torch_geometric.loader.utils.get_edge_label_index for store tuples."""
    feature_store, graph_store = data
    edge_type = None
    if isinstance(edge_label_index, (tuple, list)) and len(edge_label_index) == 3 and all(isinstance(x, str) for x in edge_label_index):
        edge_type, edge_label_index = tuple(edge_label_index), None
    elif isinstance(edge_label_index, (tuple, list)) and len(edge_label_index) == 2 and not torch.is_tensor(edge_label_index[0]):
        edge_type, edge_label_index = edge_label_index
        edge_type = None if edge_type is None else tuple(edge_type)
    if edge_label_index is None:
        attrs = graph_store.get_all_edge_attrs()
        if edge_type is None:
            if len(attrs) != 1:
                raise ValueError("edge_label_index needs an edge type on a graph with several edge types")
            edge_type = attrs[0].edge_type
        row, col = graph_store.get_edge_index(edge_type, "coo")
        edge_label_index = torch.stack([row, col])
    elif not torch.is_tensor(edge_label_index):
        edge_label_index = torch.stack([torch.as_tensor(edge_label_index[0]), torch.as_tensor(edge_label_index[1])])
    if edge_type is None and not graph_store.is_homogeneous:
        raise ValueError("edge_label_index needs an edge type on a heterogeneous graph")
    if graph_store.is_homogeneous and len(graph_store.get_all_edge_attrs()) == 1:
        edge_type = None  # homogeneous loaders carry no input type (the reference's readers branch on it)
    return edge_type, edge_label_index
