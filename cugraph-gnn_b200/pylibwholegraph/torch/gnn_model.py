"""Homogeneous GNN model of the WholeGraph examples (role of the reference's pylibwholegraph/torch/gnn_model.py).

The reference builds its layers from torch_geometric ("pyg") or the retired wg_torch package ("wg"); here the layers are
this project's own: the neighbourhood aggregation is the CSR kernel of csrc/aggregate.cu, consumed straight from the
sampler's CSR blocks (framework "wg": sub-graph = [csr_row_ptr, csr_col_ind]).  GAT is not provided."""
from pylibwholegraph.utils.imports import import_optional
from .aggregate import csr_aggregate
from .common_options import parse_max_neighbors
from .embedding import WholeMemoryEmbedding, WholeMemoryEmbeddingModule
from .graph_structure import GraphStructure

torch = import_optional("torch")

framework_name = None


def set_framework(framework: str):
    global framework_name
    if framework not in ("wg",):
        raise NotImplementedError("only the built-in 'wg' layers exist in this build (torch_geometric / DGL are optional and absent)")
    framework_name = framework


class WGSAGEConv(torch.nn.Module):
    """aggregator 'mean': W_n * mean_{j in N(i)} x_j + W_s * x_i;  aggregator 'gcn': W * (sum_{j in N(i)} x_j + x_i) / (deg_i + 1)."""

    def __init__(self, in_feats, out_feats, aggregator="mean"):
        super().__init__()
        assert aggregator in ("mean", "gcn")
        self.aggregator = aggregator
        self.lin_neigh = torch.nn.Linear(in_feats, out_feats, bias=True)
        self.lin_self = torch.nn.Linear(in_feats, out_feats, bias=False) if aggregator == "mean" else None

    def forward(self, csr_row_ptr, csr_col_ind, x_feat, x_target_feat):
        if self.aggregator == "mean":
            agg = csr_aggregate(csr_row_ptr, csr_col_ind, x_feat, "mean")
            return self.lin_neigh(agg.to(x_feat.dtype)) + self.lin_self(x_target_feat)
        agg = csr_aggregate(csr_row_ptr, csr_col_ind, x_feat, "sum")
        deg = (csr_row_ptr[1:] - csr_row_ptr[:-1]).to(x_feat.dtype).unsqueeze(1)
        return self.lin_neigh((agg.to(x_feat.dtype) + x_target_feat) / (deg + 1))


def create_gnn_layers(in_feat_dim, hidden_feat_dim, class_count, num_layer, num_head, model_type):
    if model_type == "gat":
        raise NotImplementedError("GAT layers are not provided")
    assert model_type in ("sage", "gcn")
    layers = torch.nn.ModuleList()
    for i in range(num_layer):
        out_dim = hidden_feat_dim // num_head if i != num_layer - 1 else class_count
        in_dim = in_feat_dim if i == 0 else hidden_feat_dim
        layers.append(WGSAGEConv(in_dim, out_dim, aggregator="mean" if model_type == "sage" else "gcn"))
    return layers


def create_sub_graph(target_gid, target_gid_1, edge_data, csr_row_ptr, csr_col_ind, max_num_neighbors: int, add_self_loop: bool):
    return [csr_row_ptr, csr_col_ind]


def layer_forward(layer, x_feat, x_target_feat, sub_graph):
    return layer(sub_graph[0], sub_graph[1], x_feat, x_target_feat)


class HomoGNNModel(torch.nn.Module):
    """ids -> layered sampling (GraphStructure) -> feature gather (WholeMemoryEmbeddingModule) -> num_layer GNN layers."""

    def __init__(self, graph_structure: GraphStructure, node_embedding: WholeMemoryEmbedding, args):
        super().__init__()
        global framework_name
        if framework_name is None:
            set_framework(getattr(args, "framework", "wg"))
        self.graph_structure = graph_structure
        self.node_embedding = node_embedding
        self.num_layer = args.layernum
        self.hidden_feat_dim = args.hiddensize
        num_head = args.heads if args.model == "gat" else 1
        assert self.hidden_feat_dim % num_head == 0
        self.gnn_layers = create_gnn_layers(self.node_embedding.shape[1], self.hidden_feat_dim, args.classnum, args.layernum,
                                            num_head, args.model)
        self.add_self_loop = False
        self.gather_fn = WholeMemoryEmbeddingModule(self.node_embedding)
        self.dropout = args.dropout
        self.max_neighbors = parse_max_neighbors(args.layernum, args.neighbors)
        self.max_inference_neighbors = parse_max_neighbors(args.layernum, args.inferencesample)

    def forward(self, ids):
        max_neighbors = self.max_neighbors if self.training else self.max_inference_neighbors
        ids = ids.to(self.graph_structure.csr_col_ind.dtype).cuda()
        target_gids, edge_indice, csr_row_ptrs, csr_col_inds = self.graph_structure.multilayer_sample_without_replacement(ids, max_neighbors)
        x_feat = self.gather_fn(target_gids[0], force_dtype=torch.float32)
        for i in range(self.num_layer):
            x_target_feat = x_feat[: target_gids[i + 1].numel()]
            sub_graph = create_sub_graph(target_gids[i], target_gids[i + 1], edge_indice[i], csr_row_ptrs[i], csr_col_inds[i],
                                         max_neighbors[self.num_layer - 1 - i], self.add_self_loop)
            x_feat = layer_forward(self.gnn_layers[i], x_feat, x_target_feat, sub_graph)
            if i != self.num_layer - 1:
                x_feat = torch.nn.functional.relu(x_feat)
                x_feat = torch.nn.functional.dropout(x_feat, self.dropout, training=self.training)
        return x_feat
