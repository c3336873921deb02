"""pylibwholegraph.torch: the public names of the reference's torch layer that the hot path uses."""
from .comm import (
    WholeMemoryCommunicator,
    set_world_info,
    create_group_communicator,
    destroy_communicator,
    get_global_communicator,
    get_local_node_communicator,
    get_local_device_communicator,
    comm_set_distributed_backend,
    split_communicator,
    reset_communicators,
)
from .embedding import (
    WholeMemoryEmbedding,
    WholeMemoryEmbeddingModule,
    create_embedding,
    create_embedding_from_filelist,
    destroy_embedding,
    create_builtin_cache_policy,
    WholeMemoryOptimizer,
    EmbeddingLookupFn,
    create_wholememory_optimizer,
    destroy_wholememory_optimizer,
)
from .graph_structure import GraphStructure
from .initialize import init, init_torch_env, init_torch_env_and_create_wm_comm, finalize
from .tensor import (
    WholeMemoryTensor,
    create_wholememory_tensor,
    create_wholememory_tensor_from_filelist,
    destroy_wholememory_tensor,
)
from . import wholememory_ops, wholegraph_ops, graph_ops
from .multihop import MultiHopSampler, multihop_neighbor_sample
from .aggregate import csr_aggregate, csr_aggregate_forward, csr_transpose, sage_layer_forward, SAGEConv

from .common_options import (
    add_training_options,
    add_common_graph_options,
    add_common_model_options,
    add_common_sampler_options,
    add_node_classfication_options,
    add_dataloader_options,
    parse_max_neighbors,
)
from .data_loader import (
    NodeClassificationDataset,
    create_node_classification_datasets,
    get_train_dataloader,
    get_valid_test_dataloader,
)
from .distributed_launch import (
    add_distributed_launch_options,
    distributed_launch,
    get_rank,
    get_world_size,
    get_local_rank,
    get_local_size,
)
from .gnn_model import set_framework, create_gnn_layers, create_sub_graph, HomoGNNModel
from .utils import get_part_file_name, get_part_file_list
