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

from cugraph_pyg._pyg_compat import Data, HeteroData, SamplerOutput, HeteroSamplerOutput, NodeSamplerInput, EdgeSamplerInput, ptr2index
from .sampler_utils import filter_cugraph_pyg_store, filter_cugraph_pyg_hetero_store, neg_sample, neg_cat


class SampleIterator:
    """Combines sampler outputs with their features into mini-batches a GNN can consume."""

    # Node features are fetched ONCE per call group -- one gather of the concatenated renumber maps, the shape the gather
    # kernels run at full bandwidth -- and every mini-batch of the group receives a row slice of that block (the reference
    # gathers per mini-batch, sampler.py:51-76: one small launch each).  Set to False to fetch per mini-batch.
    prefetch_call_group_features = True

    def __init__(self, data, output_iter: Iterator[SamplerOutput]):
        self.__feature_store, self.__graph_store = data
        self.__output_iter = output_iter

    def __homogeneous_data(self, s, col):
        grp = getattr(s, "group", None)
        pre = getattr(s, "edge_index", None)
        if grp is None or not self.prefetch_call_group_features:
            data = filter_cugraph_pyg_store(self.__feature_store, self.__graph_store, s.node, s.row, col if pre is None else pre[1], s.edge, None)
            return data
        blocks = grp.get("_features")
        if blocks is None:
            node_attrs = [a for a in self.__feature_store.get_all_tensor_attrs() if not isinstance(a.group_name, tuple)]
            for a in node_attrs:
                a.index = grp["map"]
            blocks = {a.attr_name: t for a, t in zip(node_attrs, self.__feature_store.multi_get_tensor(node_attrs))} if node_attrs else {}
            grp["_features"] = blocks
        n0, n1 = s.node_span
        data = Data()
        data.edge_index = pre if pre is not None else torch.stack([s.row, col], dim=0)
        data.num_nodes = int(s.node.numel())
        for name, block in blocks.items():
            data[name] = block[n0:n1]
        edge_attrs = [a for a in self.__feature_store.get_all_tensor_attrs() if isinstance(a.group_name, tuple)]
        if edge_attrs:
            for a in edge_attrs:
                a.index = s.edge
            for a, t in zip(edge_attrs, self.__feature_store.multi_get_tensor(edge_attrs)):
                data[a.attr_name] = t
        return data

    def __next__(self):
        s = next(self.__output_iter)
        if isinstance(s, HeteroSamplerOutput):
            return self.__hetero(s)
        if not isinstance(s, SamplerOutput):
            raise ValueError("Invalid output type")
        if getattr(s, "edge_index", None) is not None:
            col = None  # the reader already holds the call group's edge_index block
        else:
            n_edges = int(s.edge.numel())
            if s.col.numel() == n_edges and s.metadata_is_coo:
                col = s.col
            else:
                col = ptr2index(s.col, n_edges)  # CSR major offsets -> COO (what PyG layers take)
        data = self.__homogeneous_data(s, col)
        if "n_id" not in data:
            data.n_id = s.node
        if s.edge is not None and "e_id" not in data:
            data.e_id = s.edge.to(torch.long)
        data.batch = s.batch
        data.num_sampled_nodes = s.num_sampled_nodes
        data.num_sampled_edges = s.num_sampled_edges
        data.input_id = s.metadata[0]
        data.batch_size = int(data.input_id.size(0))
        if len(s.metadata) == 2:
            data.seed_time = s.metadata[1]
        elif len(s.metadata) == 4:
            data.edge_label_index, data.edge_label, data.seed_time = s.metadata[1:]
        else:
            raise ValueError("Invalid metadata")
        if s.csr is not None:
            # extension: the sampler's CSR block, for aggregation kernels that consume it directly
            data.csr_indptr, data.csr_indices = s.csr
        return data

    def __hetero_data(self, s):
        grp = getattr(s, "group", None)
        if grp is None or not self.prefetch_call_group_features:
            return filter_cugraph_pyg_hetero_store(self.__feature_store, self.__graph_store, s.node, s.row, s.col, s.edge, None)
        reader, b = s.reader, s.batch_index
        vtypes = list(s.node.keys())
        blocks = grp.get("_features")
        if blocks is None:  # one gather per (vertex type, feature) for the whole call group
            attrs = []
            for a in self.__feature_store.get_all_tensor_attrs():
                if not isinstance(a.group_name, tuple) and a.group_name in vtypes:
                    a.index = reader.type_index(grp, vtypes.index(a.group_name))
                    attrs.append(a)
            blocks = {(a.group_name, a.attr_name): t for a, t in zip(attrs, self.__feature_store.multi_get_tensor(attrs))} if attrs else {}
            grp["_features"] = blocks
        data = HeteroData()
        for et, ei in s.edge_index.items():
            data[et].edge_index = ei
        for nt, node in s.node.items():
            data[nt].num_nodes = int(node.numel())
        for (nt, name), block in blocks.items():
            t0, t1 = grp["_type_start"][vtypes.index(nt)][b], grp["_type_start"][vtypes.index(nt)][b + 1]
            data[nt][name] = block[t0:t1]
        edge_attrs = [a for a in self.__feature_store.get_all_tensor_attrs() if isinstance(a.group_name, tuple) and a.group_name in s.edge]
        if edge_attrs:
            for a in edge_attrs:
                a.index = s.edge[a.group_name]
            for a, t in zip(edge_attrs, self.__feature_store.multi_get_tensor(edge_attrs)):
                data[a.group_name][a.attr_name] = t
        return data

    def __hetero(self, s: HeteroSamplerOutput):
        """sampler.py:118-163 of the reference: HeteroData with per-type n_id / e_id / counts."""
        data = self.__hetero_data(s)
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
        if len(s.metadata) == 2:
            data[input_type].seed_time = s.metadata[1]
        elif len(s.metadata) == 4:
            data[input_type].edge_label_index, data[input_type].edge_label, data[input_type].seed_time = s.metadata[1:]
        else:
            raise ValueError("Invalid metadata")
        return data

    def __iter__(self):
        return self


class _Output(SamplerOutput):
    """SamplerOutput + private fields used by SampleIterator."""

    metadata_is_coo: bool = True
    csr = None
    edge_index = None  # [2, E] view of the call group's edge_index block (row = source, col = destination, batch-local ids)
    group = None      # the call group's raw sampler output (shared by its mini-batches)
    node_span = None  # (first, last + 1) position of this mini-batch's vertices in the call group's concatenated renumber map


class _HeteroOutput(HeteroSamplerOutput):
    """HeteroSamplerOutput + private fields used by SampleIterator."""

    edge_index = None  # {edge type: [2, E] view of the call group's edge_index block}
    group = None       # the call group's raw sampler output
    batch_index = None  # position of this mini-batch in its call group


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

    @staticmethod
    def _seed_metadata(raw, index: int):
        """(input_index, ...) of mini-batch `index`: node seeds -> (input_index, None); edge seeds ->
        (input_index, edge_label_index, edge_label, None) with negatives (input id -1) dropped from input_index and
        labelled 0 (reference: sampler.py:576-628)."""
        i0, i1 = raw["_input_offsets"][index], raw["_input_offsets"][index + 1]
        input_index = raw["input_index"][i0:i1]
        if "edge_inverse" not in raw:
            return (input_index, None)
        num_seeds = i1 - i0
        input_index = input_index[input_index >= 0]
        num_pos = int(input_index.numel())
        if num_seeds - num_pos > 0:
            edge_label = torch.cat([torch.full((num_pos,), 1.0), torch.full((num_seeds - num_pos,), 0.0)])
        elif "input_label" in raw:
            edge_label = raw["input_label"][i0:i1]
        else:
            edge_label = None
        edge_inverse = raw["edge_inverse"][2 * i0:2 * i1].view(2, -1)
        return (input_index, edge_inverse, edge_label, None)


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
        # per-hop counts of every mini-batch as two host tensors per call group (a mini-batch takes a row)
        el = torch.tensor(raw["_edge_lho"], dtype=torch.int64)
        raw["_num_sampled_edges"] = (el[1:] - el[:-1]).view(B, L)
        rm = torch.tensor(raw["_rmo"], dtype=torch.int64)
        steps = torch.cat([torch.tensor(raw["_base"], dtype=torch.int64).view(L + 1, B), (rm[1:] - rm[:-1]).view(1, B)], dim=0)
        raw["_num_sampled_nodes"] = (steps[1:] - steps[:-1]).t().contiguous()
        # edge_index of every mini-batch of the call group in one block, built by a handful of launches per CALL GROUP; a
        # mini-batch takes a column slice (the per-mini-batch form -- offsets minus base, ptr2index, stack -- was ~7 launches each)
        minors = raw["minors"]
        E = int(minors.numel())
        if "major_offsets" in raw:
            dev = minors.device
            mo = raw["major_offsets"]
            elho, rlho = raw["_edge_lho"], raw["_lho"]
            e0 = torch.tensor([elho[b * L] for b in range(B)], dtype=mo.dtype).to(dev, non_blocking=True)
            r0 = torch.tensor([rlho[b * L] for b in range(B)], dtype=torch.int64)
            n_rows = torch.tensor([rlho[(b + 1) * L] - rlho[b * L] for b in range(B)], dtype=torch.int64)
            n_edges = torch.tensor([elho[(b + 1) * L] - elho[b * L] for b in range(B)], dtype=torch.int64).to(dev, non_blocking=True)
            total_rows = int(n_rows.sum())
            # batch-local indptr of every mini-batch, boundaries duplicated: batch b at [r0_b + b, r0_b + b + rows_b]
            bid = torch.repeat_interleave(torch.arange(B, device=dev), (n_rows + 1).to(dev, non_blocking=True), output_size=total_rows + B)
            raw["_indptr_local"] = mo[torch.arange(total_rows + B, device=dev) - bid] - e0[bid]
            raw["_indptr_at"] = (r0 + torch.arange(B)).tolist()
            row_of_edge = ptr2index(mo[: total_rows + 1], E)  # row of the call group per edge
            majors_local = row_of_edge - torch.repeat_interleave(r0.to(dev, non_blocking=True), n_edges, output_size=E)
            raw["_edge_index"] = torch.stack([minors, majors_local.to(minors.dtype)], dim=0)
        else:
            raw["_edge_index"] = torch.stack([minors, raw["majors"]], dim=0)

    def _decode(self, raw: Dict[str, torch.Tensor], index: int):
        L, B = raw["_L"], raw["_B"]
        lho, elho = raw["_lho"], raw["_edge_lho"]
        e0, e1 = elho[index * L], elho[(index + 1) * L]
        n0, n1 = raw["_rmo"][index], raw["_rmo"][index + 1]
        node = raw["map"][n0:n1]
        minors = raw["minors"][e0:e1]
        edge = raw["edge_id"][e0:e1]
        num_sampled_edges = raw["_num_sampled_edges"][index]
        num_sampled_nodes = raw["_num_sampled_nodes"][index]
        num_seeds = raw["_base"][1][index]
        out = _Output(node=node, row=minors, col=None, edge=edge, batch=node[:num_seeds],
                      num_sampled_nodes=num_sampled_nodes, num_sampled_edges=num_sampled_edges,
                      metadata=self._seed_metadata(raw, index))
        out.group, out.node_span = raw, (n0, n1)
        out.edge_index = raw["_edge_index"][:, e0:e1]
        if "major_offsets" in raw:
            r0, r1 = lho[index * L], lho[(index + 1) * L]
            at = raw["_indptr_at"][index]
            out.col = raw["_indptr_local"][at:at + (r1 - r0) + 1]
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
        # Everything a mini-batch needs is built per CALL GROUP by a handful of launches and sliced per mini-batch (as in the
        # homogeneous reader): type-local vertex ids, the edge_index block, per-hop counts.
        dev = raw["map"].device
        n_total = int(raw["map"].numel())
        rm = host[lto.numel():lto.numel() + rmo.numel()]
        seg_len = rm[1:] - rm[:-1]                                   # segments [label][vertex type]
        seg_off = torch.tensor(self.__vertex_offsets[:Vt], dtype=torch.int64).repeat(B)
        raw["_map_local"] = raw["map"] - torch.repeat_interleave(seg_off.to(dev, non_blocking=True), seg_len.to(dev, non_blocking=True),
                                                                 output_size=n_total)
        raw["_edge_index"] = torch.stack([raw["minors"], raw["majors"]], dim=0)
        steps = torch.cat([host[lto.numel() + rmo.numel():].view(L + 1, Vt, B), seg_len.view(B, Vt).t().reshape(1, Vt, B)], dim=0)
        raw["_num_sampled_nodes"] = (steps[1:] - steps[:-1]).permute(2, 1, 0).contiguous()        # [B, Vt, L + 1]
        lt = host[:lto.numel()]
        raw["_num_sampled_edges"] = (lt[1:] - lt[:-1]).view(B, T, L)
        # where the vertices of (mini-batch b, type vt) sit in the call group's per-type concatenation (feature prefetch)
        seg = seg_len.view(B, Vt)
        raw["_type_start"] = torch.cat([torch.zeros((1, Vt), dtype=torch.int64), seg.cumsum(0)], dim=0).t().tolist()  # [Vt][B + 1]

    def type_index(self, raw, vt: int):
        """type-local ids of every vertex of type vt in the call group, mini-batch after mini-batch (no host sync)."""
        key = ("_type_index", vt)
        if key not in raw:
            Vt, B = raw["_Vt"], raw["_B"]
            dev = raw["map"].device
            rmo, tstart = raw["_rmo"], raw["_type_start"][vt]
            total = tstart[B]
            lens = torch.tensor([tstart[b + 1] - tstart[b] for b in range(B)], dtype=torch.int64)
            shift = torch.tensor([rmo[b * Vt + vt] - tstart[b] for b in range(B)], dtype=torch.int64)
            pos = torch.arange(total, device=dev) + torch.repeat_interleave(shift.to(dev, non_blocking=True), lens.to(dev, non_blocking=True),
                                                                            output_size=total)
            raw[key] = raw["_map_local"][pos]
        return raw[key]

    def _decode(self, raw: Dict[str, torch.Tensor], index: int):
        T, Vt, L = raw["_T"], raw["_Vt"], raw["_L"]
        lto, rmo, base = raw["_lto"], raw["_rmo"], raw["_base"]
        node, num_sampled_nodes = {}, {}
        for vt, name in enumerate(self.__vertex_types):
            n0, n1 = rmo[index * Vt + vt], rmo[index * Vt + vt + 1]
            node[name] = raw["_map_local"][n0:n1]
            num_sampled_nodes[name] = raw["_num_sampled_nodes"][index, vt]
        row, col, edge, num_sampled_edges, edge_index = {}, {}, {}, {}, {}
        for t, et in enumerate(self.__edge_types):
            g = (index * T + t) * L
            e0, e1 = lto[g], lto[g + L]
            edge_index[et] = raw["_edge_index"][:, e0:e1]
            row[et] = edge_index[et][0]
            col[et] = edge_index[et][1]
            # edge_id is the position inside the (label, edge type) group and the group's slice of
            # edge_renumber_map starts at its first edge, so emap[edge_id] (sampler.py:334-341) is this slice
            edge[et] = raw["edge_renumber_map"][e0:e1]
            num_sampled_edges[et] = raw["_num_sampled_edges"][index, t]
        input_type = raw["input_type"]
        if isinstance(input_type, (list, tuple)):
            input_type = tuple(input_type)
            if input_type not in self.__edge_types:
                raise ValueError("Input type did not match any edge type!")
        elif not (isinstance(input_type, str) and input_type in self.__vertex_types):
            raise ValueError("Input type did not match any vertex type!")
        meta = self._seed_metadata(raw, index)
        # edge_inverse already holds ids local to (label, vertex type): no de-offsetting by the other type's count
        # (the reference subtracts `max + 1` of the lower type, sampler.py:452-461)
        out = _HeteroOutput(node=node, row=row, col=col, edge=edge, batch=None, num_sampled_nodes=num_sampled_nodes,
                            num_sampled_edges=num_sampled_edges, metadata=((input_type, meta[0]),) + tuple(meta[1:]))
        out.edge_index, out.group, out.batch_index = edge_index, raw, index
        out.reader = self
        return out


class BaseSampler:
    def __init__(self, sampler, data, batch_size: int = 16):
        self.__sampler = sampler
        self.__feature_store, self.__graph_store = data
        self.__batch_size = batch_size

    def sample_from_nodes(self, index: NodeSamplerInput, **kwargs) -> Iterator[SamplerOutput]:
        metadata = {"input_type": index.input_type} if index.input_type is not None else None
        reader = self.__sampler.sample_from_nodes(index.node, batch_size=self.__batch_size, input_id=index.input_id,
                                                  input_time=index.time, metadata=metadata, **kwargs)
        return self.__reader(reader)

    def __reader(self, reader):
        attrs = self.__graph_store.get_all_edge_attrs()
        if len(attrs) == 1 and attrs[0].edge_type[0] == attrs[0].edge_type[2]:
            return HomogeneousSampleReader(reader)
        edge_types, src_types, dst_types = self.__graph_store._numeric_edge_types
        return HeterogeneousSampleReader(reader, src_types=src_types, dst_types=dst_types,
                                         vertex_offsets=self.__graph_store._vertex_offset_array, edge_types=edge_types,
                                         vertex_types=sorted(self.__graph_store._vertex_offsets.keys()))

    def sample_from_edges(self, index: EdgeSamplerInput, neg_sampling=None, **kwargs) -> Iterator[SamplerOutput]:
        """Link-prediction sampling (role of the reference's sampler.py:798-896): optional negative edges are drawn for
        the whole epoch at once and interleaved batch by batch; negatives carry input id -1."""
        src, dst, input_id = index.row, index.col, index.input_id
        if index.time is not None and neg_sampling:
            raise NotImplementedError("temporal negative sampling is not implemented (DESIGN.md §10)")
        neg_batch_size = 0
        if neg_sampling:
            src_neg, dst_neg = neg_sample(self.__graph_store, index.row, index.col, index.input_type, self.__batch_size, neg_sampling)
            if neg_sampling.is_binary():
                src, _ = neg_cat(src.cuda(), src_neg, self.__batch_size)
            else:
                # triplet: repeat a random subset of the sources so that the lengths line up (same unique vertices)
                scu = src.cuda()
                per = torch.randint(0, scu.numel(), (dst_neg.numel(),), device=scu.device)
                src, _ = neg_cat(scu, scu[per], self.__batch_size)
            dst, neg_batch_size = neg_cat(dst.cuda(), dst_neg, self.__batch_size)
            input_id, _ = neg_cat(input_id, torch.full((dst_neg.numel(),), -1, dtype=torch.int64, device=input_id.device),
                                  self.__batch_size)
        metadata = {"input_type": index.input_type} if index.input_type is not None else None
        reader = self.__sampler.sample_from_edges(torch.stack([src.cuda(), dst.cuda()]), input_id=input_id, input_time=index.time,
                                                  input_label=index.label, batch_size=self.__batch_size + neg_batch_size,
                                                  metadata=metadata, **kwargs)
        return self.__reader(reader)
