"""GraphStore (role of the reference's cugraph_pyg/data/graph_store.py:50-631).

PyG GraphStore interface; each rank puts ITS partition of every edge type, the sampler graph
(pylibcugraph-shaped SGGraph/MGGraph over the B200 kernels) is built lazily or by finalize().
Conventions kept from the reference:
  * cuGraph src = PyG edge_index[1], cuGraph dst = PyG edge_index[0] (:509-533): the sampler follows PyG in-edges;
  * vertex ids are offset by vertex type in lexicographic type order (:372-383);
  * edge ids are running indices per edge type, global across ranks (:578-607).
"""
from typing import Dict, List, Optional, Tuple, Union

import torch

import pylibcugraph
from cugraph_pyg._pyg_compat import EdgeAttr, EdgeLayout, GraphStoreBase
from cugraph_pyg.tensor import DistMatrix
from cugraph_pyg.utils import dist as dist_utils


class GraphStore(GraphStoreBase):
    def __init__(self, location: str = "cpu"):
        if location not in ("cpu", "cuda"):
            raise ValueError("location must be 'cpu' or 'cuda'")
        self.__edge_indices: Dict[Tuple[str, str, str], DistMatrix] = {}
        self.__sizes: Dict[Tuple[str, str, str], Optional[Tuple[int, int]]] = {}
        self.__finalized = False
        self.__handle = None
        self.__clear_graph()
        super().__init__()

    def __clear_graph(self):
        if self.__finalized:
            raise NotImplementedError("Modifying a finalized GraphStore is not supported.")
        self.__graph = None
        self.__vertex_offsets = None
        self.__weight_attr = None
        self.__time_attr = None
        self.__numeric_edge_types = None
        self.__num_vertices_cache = None

    def finalize(self, weight_attr=None, time_attr=None):
        """Build the sampler graph now and drop the COO copies; the store becomes read-only."""
        if self.__finalized:
            raise RuntimeError("This GraphStore object has already been finalized.")
        if weight_attr is not None:
            self._set_weight_attr(weight_attr)
        if time_attr is not None:
            self._set_time_attr(time_attr)
        self.__construct_graph(finalize=True)
        self.__finalized = True
        return self

    # ---- PyG GraphStore interface -------------------------------------------------------------------
    def _put_edge_index(self, edge_index, edge_attr: EdgeAttr) -> bool:
        if edge_attr.layout != EdgeLayout.COO:
            raise ValueError("Only COO format supported")
        if isinstance(edge_index, (tuple, list)):
            edge_index = torch.stack([torch.as_tensor(edge_index[0]), torch.as_tensor(edge_index[1])])
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError("edge_index must have shape [2, num_edges]")
        self.__clear_graph()
        self.__edge_indices[edge_attr.edge_type] = DistMatrix(edge_index, shape=edge_attr.size, dtype=torch.int64)
        self.__sizes[edge_attr.edge_type] = tuple(edge_attr.size) if edge_attr.size is not None else None
        return True

    def _get_edge_index(self, edge_attr: EdgeAttr):
        m = self.__edge_indices.get(edge_attr.edge_type)
        if m is None:
            return None
        if self.__finalized:
            raise RuntimeError("edge indices were released by finalize()")
        row, col = m.local_row, m.local_col
        if edge_attr.layout == EdgeLayout.COO:
            return row, col
        n_dst, n_src = self.__sizes[edge_attr.edge_type] or (int(row.max()) + 1, int(col.max()) + 1)
        if edge_attr.layout == EdgeLayout.CSR:
            order = torch.argsort(row, stable=True)
            ptr = torch.zeros(n_dst + 1, dtype=torch.int64, device=row.device)
            ptr[1:] = torch.bincount(row, minlength=n_dst).cumsum(0)
            return ptr, col[order]
        order = torch.argsort(col, stable=True)
        ptr = torch.zeros(n_src + 1, dtype=torch.int64, device=row.device)
        ptr[1:] = torch.bincount(col, minlength=n_src).cumsum(0)
        return row[order], ptr

    def _remove_edge_index(self, edge_attr: EdgeAttr) -> bool:
        if edge_attr.edge_type not in self.__edge_indices:
            return False
        self.__clear_graph()
        del self.__edge_indices[edge_attr.edge_type]
        del self.__sizes[edge_attr.edge_type]
        return True

    def get_all_edge_attrs(self) -> List[EdgeAttr]:
        return [EdgeAttr(et, EdgeLayout.COO, is_sorted=False, size=self.__sizes[et]) for et in self.__edge_indices]

    # ---- internals the loaders use ----------------------------------------------------------------------
    @property
    def is_multi_gpu(self) -> bool:
        return torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1

    @property
    def _resource_handle(self):
        if self.__handle is None:
            self.__handle = pylibcugraph.ResourceHandle()
        return self.__handle

    def _num_vertices(self) -> Dict[str, int]:
        if self.__num_vertices_cache is not None:
            return dict(self.__num_vertices_cache)
        nv: Dict[str, int] = {}

        def bump(k, v):
            nv[k] = max(nv.get(k, 0), int(v))

        for et, m in self.__edge_indices.items():
            size = self.__sizes[et]
            if size is not None:
                bump(et[0], size[0])
                bump(et[2], size[1])
            elif m.local_row.numel():
                if et[0] != et[2]:
                    bump(et[0], m.local_row.max() + 1)
                    bump(et[2], m.local_col.max() + 1)
                else:
                    bump(et[0], m.local_coo.max() + 1)
        if self.is_multi_gpu:
            for k in sorted(nv):
                nv[k] = dist_utils.all_reduce_max(nv[k])
        self.__num_vertices_cache = dict(nv)
        return nv

    @property
    def _vertex_offsets(self) -> Dict[str, int]:
        if self.__vertex_offsets is None:
            nv = self._num_vertices()
            self.__vertex_offsets, off = {}, 0
            for k in sorted(nv):
                self.__vertex_offsets[k] = off
                off += nv[k]
        return dict(self.__vertex_offsets)

    @property
    def _vertex_offset_array(self) -> torch.Tensor:
        offs = self._vertex_offsets
        total = sum(self._num_vertices().values())
        return torch.tensor([offs[k] for k in sorted(offs)] + [total], dtype=torch.int64, device="cuda")

    @property
    def is_homogeneous(self) -> bool:
        return len(self._vertex_offsets) == 1

    def _set_weight_attr(self, attr):
        if attr != self.__weight_attr:
            time_attr = self.__time_attr
            self.__clear_graph()
            self.__weight_attr = attr
            self.__time_attr = time_attr

    def _set_time_attr(self, attr):
        """(feature_store, attribute name) of the EDGE times used by temporal sampling (reference :410-416, 621-629:
        edge-based temporal sampling only)."""
        if attr != self.__time_attr:
            weight_attr = self.__weight_attr
            self.__clear_graph()
            self.__time_attr = attr
            self.__weight_attr = weight_attr

    def _get_ntime_func(self):
        if self.__time_attr is None:
            return None
        fs, name = self.__time_attr
        return lambda node_type, node_id: fs[node_type, name][node_id]

    @property
    def _numeric_edge_types(self):
        if self.__numeric_edge_types is None:
            keys = sorted(self.__edge_indices.keys())
            vt = {k: i for i, k in enumerate(sorted(self._vertex_offsets))}
            self.__numeric_edge_types = (
                keys,
                torch.tensor([vt[k[0]] for k in keys], device="cuda", dtype=torch.int32),
                torch.tensor([vt[k[2]] for k in keys], device="cuda", dtype=torch.int32),
            )
        return self.__numeric_edge_types

    def __edge_list(self, finalize: bool):
        keys = sorted(self.__edge_indices.keys())
        starts = dist_utils.edge_id_starts([self.__edge_indices[k].local_row.numel() for k in keys])
        offs = self._vertex_offsets
        dst, src, eid, etp, wgt, tme = [], [], [], [], [], []
        for i, k in enumerate(keys):
            m = self.__edge_indices[k]
            n = m.local_row.numel()
            dst.append(m.local_row + offs[k[0]])   # PyG row 0 = cuGraph dst
            src.append(m.local_col + offs[k[2]])   # PyG row 1 = cuGraph src
            eid.append(torch.arange(int(starts[i]), int(starts[i]) + n, dtype=torch.int64, device="cuda"))
            etp.append(torch.full((n,), i, dtype=torch.int32, device="cuda"))
            if self.__weight_attr is not None:
                fs, name = self.__weight_attr
                w = fs[k, name, None]
                ids = torch.arange(int(starts[i]), int(starts[i]) + n, device="cuda")
                wgt.append(w[ids].reshape(-1).float() if n else torch.empty(0, device="cuda"))
            if self.__time_attr is not None:
                fs, name = self.__time_attr
                tm = fs[k, name, None]
                if tm is None:
                    raise ValueError("Time property must be present for all edge types.")
                ids = torch.arange(int(starts[i]), int(starts[i]) + n, device="cuda")
                tme.append(tm[ids].reshape(-1).to(device="cuda", dtype=torch.int64) if n else torch.empty(0, dtype=torch.int64, device="cuda"))
        d = {"dst": torch.cat(dst) if dst else torch.empty(0, dtype=torch.int64, device="cuda"),
             "src": torch.cat(src) if src else torch.empty(0, dtype=torch.int64, device="cuda"),
             "eid": torch.cat(eid) if eid else torch.empty(0, dtype=torch.int64, device="cuda"),
             "etp": torch.cat(etp) if etp else torch.empty(0, dtype=torch.int32, device="cuda")}
        if wgt:
            d["wgt"] = torch.cat(wgt)
        if tme:
            d["etime"] = torch.cat(tme)
        return d

    def __construct_graph(self, finalize: bool = False):
        if finalize:
            nv = self._num_vertices()
            for et, size in self.__sizes.items():
                if size is None:
                    self.__sizes[et] = (nv[et[0]], nv[et[2]])
        if self.__graph is None:
            d = self.__edge_list(finalize)
            num_vertices = sum(self._num_vertices().values())
            props = pylibcugraph.GraphProperties(is_multigraph=True, is_symmetric=False)
            cls = pylibcugraph.MGGraph if self.is_multi_gpu else pylibcugraph.SGGraph
            self.__graph = cls(self._resource_handle, props, d["src"], d["dst"], weight_array=d.get("wgt"),
                               edge_id_array=d["eid"], edge_type_array=d["etp"], num_vertices=num_vertices,
                               edge_start_time_array=d.get("etime"))
        if finalize:
            for k in list(self.__edge_indices.keys()):
                self.__edge_indices[k] = DistMatrix(None, shape=self.__sizes[k])
        return self.__graph

    @property
    def _graph(self) -> Union[pylibcugraph.SGGraph, pylibcugraph.MGGraph]:
        return self.__construct_graph(finalize=False)
