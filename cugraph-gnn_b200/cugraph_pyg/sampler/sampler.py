"""Readers that turn a sampled call group into per-mini-batch PyG sampler outputs, and the iterator that attaches
features (role of the reference's cugraph_pyg/sampler/sampler.py:17-225, 505-797).

Vectorised decode: the reference pays ~30 small torch ops and >= 6 `.cpu()` syncs PER
MINI-BATCH (sampler.py:560-575, 632-639, 676-683, 722); here the small offset arrays of a call group are
copied to the host ONCE, every per-batch quantity (hop sizes, node counts, slices) is host arithmetic on
them, and a mini-batch costs a handful of tensor views plus one tiny kernel.
"""
from typing import Dict, Iterator, Tuple, Union

import torch

from typing import List

from cugraph_pyg._pyg_compat import SamplerOutput, HeteroSamplerOutput, NodeSamplerInput, ptr2index
from .sampler_utils import filter_cugraph_pyg_store, filter_cugraph_pyg_hetero_store


class SampleIterator:
    """Combines sampler outputs with their features into mini-batches a GNN can consume."""

    def __init__(self, data, output_iter: Iterator[SamplerOutput]):
        self.__feature_store, self.__graph_store = data
        self.__output_iter = output_iter

    def __next__(self):
        s = next(self.__output_iter)
        if isinstance(s, HeteroSamplerOutput):
            return self.__hetero(s)
        if not isinstance(s, SamplerOutput):
            raise ValueError("Invalid output type")
        n_edges = int(s.edge.numel())
        if s.col.numel() == n_edges and s.metadata_is_coo:
            col = s.col
        else:
            col = ptr2index(s.col, n_edges)  # CSR major offsets -> COO (what PyG layers take)
        data = filter_cugraph_pyg_store(self.__feature_store, self.__graph_store, s.node, s.row, col, s.edge, None)
        if "n_id" not in data:
            data.n_id = s.node
        if s.edge is not None and "e_id" not in data:
            data.e_id = s.edge.to(torch.long)
        data.batch = s.batch
        data.num_sampled_nodes = s.num_sampled_nodes
        data.num_sampled_edges = s.num_sampled_edges
        data.input_id = s.metadata[0]
        data.batch_size = int(data.input_id.size(0))
        data.seed_time = s.metadata[1]
        if s.csr is not None:
            # extension: the sampler's CSR block, for aggregation kernels that consume it directly
            data.csr_indptr, data.csr_indices = s.csr
        return data

    def __hetero(self, s: HeteroSamplerOutput):
        """sampler.py:118-163 of the reference: HeteroData with per-type n_id / e_id / counts."""
        data = filter_cugraph_pyg_hetero_store(self.__feature_store, self.__graph_store, s.node, s.row, s.col, s.edge, None)
        for key, node in s.node.items():
            if "n_id" not in data[key]:
                data[key].n_id = node
        for key, edge in (s.edge or {}).items():
            if edge is not None and "e_id" not in data[key]:
                data[key].e_id = edge.to(torch.long)
        data.set_value_dict("batch", s.batch)
        data.set_value_dict("num_sampled_nodes", s.num_sampled_nodes)
        data.set_value_dict("num_sampled_edges", s.num_sampled_edges)
        input_type, input_id = s.metadata[0]
        data[input_type].input_id = input_id
        data[input_type].batch_size = int(input_id.size(0))
        data[input_type].seed_time = s.metadata[1]
        return data

    def __iter__(self):
        return self


class _Output(SamplerOutput):
    """SamplerOutput + two private fields used by SampleIterator."""

    metadata_is_coo: bool = True
    csr = None


class SampleReader:
    """Iterates the mini-batches of successive call groups."""

    def __init__(self, base_reader: Iterator[Tuple[Dict[str, torch.Tensor], int, int]]):
        self.__base_reader = base_reader
        self.__remaining = 0
        self.__index = 0
        self.__raw = None

    def __next__(self):
        while self.__remaining == 0:
            self.__raw, first, last = next(self.__base_reader)
            self._prepare(self.__raw)
            self.__remaining = last - first + 1
            self.__index = 0
        out = self._decode(self.__raw, self.__index)
        self.__index += 1
        self.__remaining -= 1
        return out

    def __iter__(self):
        return self

    def _prepare(self, raw):
        pass

    def _decode(self, raw, index: int):
        raise NotImplementedError("Must be implemented by subclass")


class HomogeneousSampleReader(SampleReader):
    def _prepare(self, raw: Dict[str, torch.Tensor]):
        """One host copy per call group of everything the per-batch decode needs."""
        lho = raw["label_hop_offsets"]
        rmo = raw["renumber_map_offsets"]
        B = rmo.numel() - 1
        L = (lho.numel() - 1) // max(B, 1)
        pieces = [lho, rmo, raw["label_step_base"].reshape(-1).to(torch.int64)]
        if "major_offsets" in raw:
            pieces.append(raw["major_offsets"][lho])  # edge offsets at the (label, hop) row boundaries
        host = torch.cat(pieces).cpu()  # the only device->host sync of the call group
        n_lho = lho.numel()
        raw["_L"], raw["_B"] = L, B
        raw["_lho"] = host[:n_lho].tolist()
        raw["_rmo"] = host[n_lho:n_lho + B + 1].tolist()
        raw["_base"] = host[n_lho + B + 1:n_lho + B + 1 + (L + 1) * B].view(L + 1, B).tolist()
        raw["_edge_lho"] = host[n_lho + B + 1 + (L + 1) * B:].tolist() if "major_offsets" in raw else raw["_lho"]
        raw["_input_offsets"] = raw["input_offsets"].tolist()

    def _decode(self, raw: Dict[str, torch.Tensor], index: int):
        L, B = raw["_L"], raw["_B"]
        lho, elho = raw["_lho"], raw["_edge_lho"]
        e0, e1 = elho[index * L], elho[(index + 1) * L]
        n0, n1 = raw["_rmo"][index], raw["_rmo"][index + 1]
        node = raw["map"][n0:n1]
        minors = raw["minors"][e0:e1]
        edge = raw["edge_id"][e0:e1]
        num_sampled_edges = torch.tensor([elho[index * L + h + 1] - elho[index * L + h] for h in range(L)])
        base = [raw["_base"][t][index] for t in range(L + 1)] + [n1 - n0]
        num_sampled_nodes = torch.tensor([base[t + 1] - base[t] for t in range(L + 1)])
        i0, i1 = raw["_input_offsets"][index], raw["_input_offsets"][index + 1]
        input_index = raw["input_index"][i0:i1]
        num_seeds = base[1]
        out = _Output(node=node, row=minors, col=None, edge=edge, batch=node[:num_seeds],
                      num_sampled_nodes=num_sampled_nodes, num_sampled_edges=num_sampled_edges,
                      metadata=(input_index, None))
        if "major_offsets" in raw:
            r0, r1 = lho[index * L], lho[(index + 1) * L]
            out.col = raw["major_offsets"][r0:r1 + 1] - e0
            out.metadata_is_coo = False
            out.csr = (out.col, minors)
        else:
            out.col = raw["majors"][e0:e1]
            out.metadata_is_coo = True
        return out


class HeterogeneousSampleReader(SampleReader):
    """Heterogeneous call groups (role of the reference's sampler.py:231-502).  Same vectorised decode as the
    homogeneous reader: the (label, edge type, hop) offsets, the map offsets and the per-step bases are copied to the
    host once per call group; a mini-batch is then slices, one subtraction per vertex type and no host sync (the
    reference reduces `majors[:lho].max()` on the device and calls `.cpu()` per type and hop, :349-362, 404-410)."""

    def __init__(self, base_reader, src_types: torch.Tensor, dst_types: torch.Tensor, vertex_offsets: torch.Tensor,
                 edge_types: List[Tuple[str, str, str]], vertex_types: List[str]):
        self.__src_types = [int(x) for x in src_types.tolist()]
        self.__dst_types = [int(x) for x in dst_types.tolist()]
        self.__edge_types = list(edge_types)
        self.__vertex_types = list(vertex_types)
        self.__vertex_offsets = [int(x) for x in vertex_offsets.tolist()]
        super().__init__(base_reader)

    def _prepare(self, raw: Dict[str, torch.Tensor]):
        if "major_offsets" in raw:
            raise ValueError("CSR format not currently supported for heterogeneous graphs")
        if raw.get("input_type") is None:
            raise ValueError("No input type found!")
        T, Vt = len(self.__edge_types), len(self.__vertex_types)
        lto, rmo = raw["label_type_hop_offsets"], raw["renumber_map_offsets"]
        B = (rmo.numel() - 1) // Vt
        L = (lto.numel() - 1) // max(B * T, 1)
        host = torch.cat([lto, rmo, raw["label_type_step_base"].reshape(-1).to(torch.int64)]).cpu()  # the only sync
        raw["_T"], raw["_Vt"], raw["_L"], raw["_B"] = T, Vt, L, B
        raw["_lto"] = host[:lto.numel()].tolist()
        raw["_rmo"] = host[lto.numel():lto.numel() + rmo.numel()].tolist()
        raw["_base"] = host[lto.numel() + rmo.numel():].view(L + 1, Vt, B).tolist()
        raw["_input_offsets"] = raw["input_offsets"].tolist()

    def _decode(self, raw: Dict[str, torch.Tensor], index: int):
        T, Vt, L = raw["_T"], raw["_Vt"], raw["_L"]
        lto, rmo, base = raw["_lto"], raw["_rmo"], raw["_base"]
        node, num_sampled_nodes = {}, {}
        for vt, name in enumerate(self.__vertex_types):
            n0, n1 = rmo[index * Vt + vt], rmo[index * Vt + vt + 1]
            node[name] = raw["map"][n0:n1] - self.__vertex_offsets[vt]
            b = [base[s][vt][index] for s in range(L + 1)] + [n1 - n0]
            num_sampled_nodes[name] = torch.tensor([b[s + 1] - b[s] for s in range(L + 1)])
        row, col, edge, num_sampled_edges = {}, {}, {}, {}
        for t, et in enumerate(self.__edge_types):
            g = (index * T + t) * L
            e0, e1 = lto[g], lto[g + L]
            row[et] = raw["minors"][e0:e1]
            col[et] = raw["majors"][e0:e1]
            # edge_id is the position inside the (label, edge type) group and the group's slice of
            # edge_renumber_map starts at its first edge, so emap[edge_id] (sampler.py:334-341) is this slice
            edge[et] = raw["edge_renumber_map"][e0:e1]
            num_sampled_edges[et] = torch.tensor([lto[g + h + 1] - lto[g + h] for h in range(L)])
        input_type = raw["input_type"]
        if not (isinstance(input_type, str) and input_type in self.__vertex_types):
            raise ValueError("Input type did not match any vertex type!")
        i0, i1 = raw["_input_offsets"][index], raw["_input_offsets"][index + 1]
        input_index = raw["input_index"][i0:i1]
        return HeteroSamplerOutput(node=node, row=row, col=col, edge=edge, batch=None, num_sampled_nodes=num_sampled_nodes,
                                   num_sampled_edges=num_sampled_edges, metadata=((input_type, input_index), None))


class BaseSampler:
    def __init__(self, sampler, data, batch_size: int = 16):
        self.__sampler = sampler
        self.__feature_store, self.__graph_store = data
        self.__batch_size = batch_size

    def sample_from_nodes(self, index: NodeSamplerInput, **kwargs) -> Iterator[SamplerOutput]:
        metadata = {"input_type": index.input_type} if index.input_type is not None else None
        reader = self.__sampler.sample_from_nodes(index.node, batch_size=self.__batch_size, input_id=index.input_id,
                                                  input_time=index.time, metadata=metadata, **kwargs)
        attrs = self.__graph_store.get_all_edge_attrs()
        if len(attrs) == 1 and attrs[0].edge_type[0] == attrs[0].edge_type[2]:
            return HomogeneousSampleReader(reader)
        edge_types, src_types, dst_types = self.__graph_store._numeric_edge_types
        return HeterogeneousSampleReader(reader, src_types=src_types, dst_types=dst_types,
                                         vertex_offsets=self.__graph_store._vertex_offset_array, edge_types=edge_types,
                                         vertex_types=sorted(self.__graph_store._vertex_offsets.keys()))

    def sample_from_edges(self, *args, **kwargs):
        raise NotImplementedError("link-prediction sampling is outside the B200 hot path (SURVEY.md §8f row 2)")
