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
from .aggregate import csr_aggregate, csr_aggregate_forward, SAGEConv
