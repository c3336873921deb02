"""WholeMemoryEmbedding, feature-fetch path (role of the reference's pylibwholegraph/torch/embedding.py).

Only the non-cached embedding exists here: its gather is the mini-batch feature fetch
(reference: noncached_embedding::gather, cpp/src/wholememory/embedding.cpp:545-554).  Cache policies
and sparse optimizers belong to trainable embeddings and raise.
"""
from typing import List, Union

import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.utils.imports import import_optional
from .comm import WholeMemoryCommunicator
from .tensor import WholeMemoryTensor
from .utils import (
    torch_dtype_to_wholememory_dtype,
    str_to_wmb_wholememory_memory_type,
    str_to_wmb_wholememory_location,
    get_file_size,
    get_part_file_list,
    get_part_file_name,
)
from .wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor

torch = import_optional("torch")


class WholeMemoryCachePolicy(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("embedding caches are outside the B200 hot path (tables live in HBM)")


def create_builtin_cache_policy(builtin_cache_type: str, *args, **kwargs):
    if builtin_cache_type == "none":
        return None
    raise NotImplementedError("embedding caches are outside the B200 hot path (tables live in HBM)")


class WholeMemoryEmbedding(object):
    r"""Row-sharded feature table with a gather that reads peer rows by P2P."""

    def __init__(self, wmb_embedding: wmb.PyWholeMemoryEmbedding, wmb_cache_policy=None):
        self.wmb_embedding = wmb_embedding
        self.wmb_cache_policy = None
        self.embedding_tensor = None
        self.adjust_cache = False

    def dim(self):
        return self.get_embedding_tensor().dim()

    @property
    def shape(self):
        return self.get_embedding_tensor().shape

    def set_adjust_cache(self, adjust_cache: bool):
        self.adjust_cache = False

    def need_grad(self):
        return False

    def gather(self, indice: "torch.Tensor", *, is_training: bool = False,
               force_dtype: Union["torch.dtype", None] = None):
        assert indice.dim() == 1
        table = self.get_embedding_tensor()
        out = torch.empty(
            (indice.shape[0], table.shape[1]),
            device="cuda:%d" % torch.cuda.current_device(),
            dtype=table.dtype if force_dtype is None else force_dtype,
        )
        wmb.EmbeddingGatherForward(
            self.wmb_embedding, wrap_torch_tensor(indice), wrap_torch_tensor(out), False, get_wholegraph_env_fns(), get_stream()
        )
        return out

    def get_embedding_tensor(self):
        if self.embedding_tensor is None:
            self.embedding_tensor = WholeMemoryTensor(self.wmb_embedding.get_embedding_tensor())
        return self.embedding_tensor

    def get_optimizer_state_names(self):
        return []

    def writeback_all_cache(self):
        pass

    def drop_all_cache(self):
        pass

    def save(self, file_prefix: str):
        self.get_embedding_tensor().to_file_prefix(file_prefix + "_embedding_tensor")

    def load(self, file_prefix: str, *, ignore_embedding: bool = False, part_count: Union[int, None] = None):
        if not ignore_embedding:
            self.get_embedding_tensor().from_file_prefix(file_prefix + "_embedding_tensor", part_count)


def create_embedding(
    comm: WholeMemoryCommunicator,
    memory_type: str,
    memory_location: str,
    dtype: "torch.dtype",
    sizes: List[int],
    *,
    cache_policy=None,
    embedding_entry_partition: Union[List[int], None] = None,
    random_init: bool = False,
    gather_sms: int = -1,
    round_robin_size: int = 0,
):
    r"""Collective: allocate a [rows, dim] embedding table striped over the communicator."""
    if cache_policy is not None:
        raise NotImplementedError("embedding caches are outside the B200 hot path (tables live in HBM)")
    assert len(sizes) == 2
    if embedding_entry_partition is not None and round_robin_size != 0:
        print("round_robin_size is ignored because embedding_entry_partition is specified")
        round_robin_size = 0
    desc = wmb.PyWholeMemoryTensorDescription()
    desc.set_dtype(torch_dtype_to_wholememory_dtype(dtype))
    desc.set_shape(sizes)
    desc.set_stride([sizes[1], 1])
    emb = WholeMemoryEmbedding(
        wmb.create_embedding(
            desc,
            comm.wmb_comm,
            str_to_wmb_wholememory_memory_type(memory_type),
            str_to_wmb_wholememory_location(memory_location),
            None,
            embedding_entry_partition=embedding_entry_partition,
            user_defined_sms=gather_sms,
            round_robin_size=round_robin_size,
        )
    )
    if random_init:
        local, _ = emb.get_embedding_tensor().get_local_tensor()
        if local.numel() > 0:
            torch.nn.init.xavier_uniform_(local)
    comm.barrier()
    return emb


def create_embedding_from_filelist(
    comm: WholeMemoryCommunicator,
    memory_type: str,
    memory_location: str,
    filelist: Union[List[str], str],
    dtype: "torch.dtype",
    last_dim_size: int,
    *,
    cache_policy=None,
    embedding_entry_partition: Union[List[int], None] = None,
    gather_sms: int = -1,
    round_robin_size: int = 0,
):
    r"""Collective: size the table from raw row-major binary files and load them."""
    if isinstance(filelist, str):
        filelist = [filelist]
    assert last_dim_size > 0
    row = torch.tensor([], dtype=dtype).element_size() * last_dim_size
    total = sum(get_file_size(f) for f in filelist)
    if total % row != 0:
        raise ValueError("total file size %d is not a multiple of the row size %d" % (total, row))
    emb = create_embedding(
        comm, memory_type, memory_location, dtype, [total // row, last_dim_size],
        cache_policy=cache_policy, embedding_entry_partition=embedding_entry_partition,
        gather_sms=gather_sms, round_robin_size=round_robin_size,
    )
    emb.get_embedding_tensor().from_filelist(filelist, round_robin_size)
    return emb


def destroy_embedding(wm_embedding: WholeMemoryEmbedding):
    wm_embedding.embedding_tensor = None
    wm_embedding.wmb_embedding.destroy_embedding()
    wm_embedding.wmb_embedding = None


class WholeMemoryEmbeddingModule(torch.nn.Module):
    """torch.nn.Module wrapper: forward(indices) = embedding.gather(indices)."""

    def __init__(self, wm_embedding: WholeMemoryEmbedding):
        super().__init__()
        self.wm_embedding = wm_embedding

    def forward(self, indice: "torch.Tensor", force_dtype: Union["torch.dtype", None] = None):
        return self.wm_embedding.gather(indice, is_training=self.training, force_dtype=force_dtype)
