"""WholeMemoryEmbedding (role of the reference's pylibwholegraph/torch/embedding.py).

The non-cached embedding: its gather is the mini-batch feature fetch (reference: noncached_embedding::gather,
cpp/src/wholememory/embedding.cpp:545-554); with a WholeMemoryOptimizer attached it is trainable -- gradients of the
gathered rows are collected by EmbeddingLookupFn.backward and applied by WholeMemoryOptimizer.step(lr)
(wholememory_embedding_gather_gradient_apply, csrc/embedding_optimizer.cu).  Cache policies are not provided: tables
live in HBM.
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
    str_to_wmb_wholememory_optimizer_type,
)
from .comm import get_global_communicator
from .wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor

torch = import_optional("torch")


class WholeMemoryCachePolicy(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("embedding caches are outside the B200 hot path (tables live in HBM)")


def create_builtin_cache_policy(builtin_cache_type: str, *args, **kwargs):
    if builtin_cache_type == "none":
        return None
    raise NotImplementedError("embedding caches are outside the B200 hot path (tables live in HBM)")


class WholeMemoryOptimizer(object):
    """Sparse optimizer for WholeMemoryEmbedding; several embeddings may share one.  Use create_wholememory_optimizer."""

    def __init__(self, global_comm: WholeMemoryCommunicator):
        super().__init__()
        self.wmb_opt = wmb.WholeMemoryOptimizer()
        self.embeddings = []
        self.global_comm = global_comm

    def add_embedding(self, wm_embedding):
        """Collective: allocates the per-row optimizer states, partitioned like the table."""
        assert isinstance(wm_embedding, WholeMemoryEmbedding)
        if wm_embedding.wmb_optimizer is not None:
            raise ValueError("optimizer can only be set once.")
        wm_embedding.wmb_optimizer = self.wmb_opt
        wm_embedding.dummy_input.requires_grad_(True)
        self.wmb_opt.add_embedding(wm_embedding.wmb_embedding)
        self.embeddings.append(wm_embedding)

    def step(self, lr: float):
        r"""Apply the collected gradients to every embedding of this optimizer (collective)."""
        for wm_embedding in self.embeddings:
            if wm_embedding.need_apply:
                wm_embedding.apply_gradients(lr)
        self.global_comm.barrier()


class EmbeddingLookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, indice, dummy_input, wm_embedding, is_training: bool = False, force_dtype=None):
        output_tensor = wm_embedding.gather(indice, is_training=is_training, force_dtype=force_dtype)
        if is_training and wm_embedding.need_grad():
            ctx.save_for_backward(indice, output_tensor, dummy_input)
            ctx.wm_embedding = wm_embedding
        return output_tensor

    @staticmethod
    def backward(ctx, grad_outputs):
        if not ctx.saved_tensors:  # eval mode or no optimizer: nothing to collect
            return None, None, None, None, None
        indice, output_tensor, dummy_input = ctx.saved_tensors
        wm_embedding = ctx.wm_embedding
        wm_embedding.add_gradients(indice, grad_outputs)
        ctx.wm_embedding = None
        return None, torch.zeros_like(dummy_input), None, None, None


class WholeMemoryEmbedding(object):
    r"""Row-sharded table with a gather that reads peer rows by P2P; trainable once an optimizer is attached."""

    def __init__(self, wmb_embedding: wmb.PyWholeMemoryEmbedding, wmb_cache_policy=None):
        self.wmb_embedding = wmb_embedding
        self.wmb_cache_policy = None
        self.embedding_tensor = None
        self.optimizer_states = dict()
        self.adjust_cache = False
        self.wmb_optimizer = None
        self.dummy_input = torch.nn.Parameter(torch.zeros(1), requires_grad=False)
        self.need_apply = False
        self.sparse_indices = []
        self.sparse_grads = []

    def dim(self):
        return self.get_embedding_tensor().dim()

    @property
    def shape(self):
        return self.get_embedding_tensor().shape

    def set_adjust_cache(self, adjust_cache: bool):
        self.adjust_cache = False

    def need_grad(self):
        return self.wmb_optimizer is not None

    def gather(self, indice: "torch.Tensor", *, is_training: bool = False,
               force_dtype: Union["torch.dtype", None] = None):
        assert indice.dim() == 1
        table = self.get_embedding_tensor()
        need_grad = self.need_grad() and is_training
        out = torch.empty(
            (indice.shape[0], table.shape[1]),
            device="cuda:%d" % torch.cuda.current_device(),
            dtype=table.dtype if force_dtype is None else force_dtype,
            requires_grad=need_grad,
        )
        if need_grad:
            self.need_apply = True
        wmb.EmbeddingGatherForward(
            self.wmb_embedding, wrap_torch_tensor(indice), wrap_torch_tensor(out), False, get_wholegraph_env_fns(), get_stream()
        )
        return out

    def set_hot_rows(self, hot_indices: Union["torch.Tensor", None]):
        """Replicate the given rows (int64 CUDA tensor; None drops the replica) on THIS GPU: later gathers serve them
        locally instead of over NVLink, with identical results.  Typical choice: the vertices of highest degree
        (`torch.bincount(col).topk(k).indices`).  Costs len(hot_indices) * row bytes + 4 bytes per table row of HBM;
        rank-local; a snapshot of the rows (call again after changing the table)."""
        if hot_indices is None:
            self.wmb_embedding.set_hot_rows(None, get_stream())
            return
        hot_indices = hot_indices.to(device="cuda:%d" % torch.cuda.current_device(), dtype=torch.int64).contiguous()
        self.wmb_embedding.set_hot_rows(wrap_torch_tensor(hot_indices), get_stream())

    def hot_row_count(self) -> int:
        return self.wmb_embedding.hot_row_count()

    def add_gradients(self, indice: "torch.Tensor", grad_outputs: "torch.Tensor"):
        self.sparse_indices.append(indice)
        self.sparse_grads.append(grad_outputs)

    def apply_gradients(self, lr: float):
        """Collective over the embedding's communicator (every rank passes its own gradients, possibly none)."""
        dim = self.get_embedding_tensor().shape[1]
        if self.sparse_indices:
            sparse_indices = torch.cat(self.sparse_indices)
            sparse_grads = torch.cat(self.sparse_grads).float().contiguous()
        else:
            dev = "cuda:%d" % torch.cuda.current_device()
            sparse_indices = torch.empty(0, dtype=torch.int64, device=dev)
            sparse_grads = torch.empty((0, dim), dtype=torch.float32, device=dev)
        wmb.EmbeddingGatherGradientApply(
            self.wmb_embedding, wrap_torch_tensor(sparse_indices), wrap_torch_tensor(sparse_grads), False, lr,
            get_wholegraph_env_fns(), get_stream(),
        )
        self.sparse_indices = []
        self.sparse_grads = []
        self.need_apply = False

    def get_embedding_tensor(self):
        if self.embedding_tensor is None:
            self.embedding_tensor = WholeMemoryTensor(self.wmb_embedding.get_embedding_tensor())
        return self.embedding_tensor

    def get_optimizer_state_names(self):
        return self.wmb_embedding.get_optimizer_state_names()

    def get_optimizer_state(self, state_name):
        if state_name not in self.optimizer_states:
            self.optimizer_states[state_name] = WholeMemoryTensor(self.wmb_embedding.get_optimizer_state(state_name))
        return self.optimizer_states[state_name]

    def writeback_all_cache(self):
        pass

    def drop_all_cache(self):
        pass

    def save(self, file_prefix: str):
        self.get_embedding_tensor().to_file_prefix(file_prefix + "_embedding_tensor")
        for state_name in self.get_optimizer_state_names():
            self.get_optimizer_state(state_name).to_file_prefix(file_prefix + "_" + state_name)

    def load(self, file_prefix: str, *, ignore_embedding: bool = False, part_count: Union[int, None] = None):
        if not ignore_embedding:
            self.get_embedding_tensor().from_file_prefix(file_prefix + "_embedding_tensor", part_count)
        for state_name in self.get_optimizer_state_names():
            self.get_optimizer_state(state_name).from_file_prefix(file_prefix + "_" + state_name, part_count)


class WholeMemoryEmbeddingModule(torch.nn.Module):
    """torch.nn.Module wrapper of WholeMemoryEmbedding (training mode routes gradients to the sparse optimizer)."""

    def __init__(self, wm_embedding: WholeMemoryEmbedding):
        super().__init__()
        self.wm_embedding = wm_embedding
        self.embedding_gather_fn = EmbeddingLookupFn.apply

    def forward(self, indice: "torch.Tensor", force_dtype: Union["torch.dtype", None] = None):
        return self.embedding_gather_fn(indice, self.wm_embedding.dummy_input, self.wm_embedding, self.training, force_dtype)


def create_wholememory_optimizer(embeddings: Union[WholeMemoryEmbedding, List[WholeMemoryEmbedding]], optimizer_type: str,
                                 param_dict: dict):
    """optimizer_type: sgd | adam (lazy) | adagrad | rmsprop; param_dict: weight_decay, epsilon, beta1, beta2, adam_w, alpha."""
    wm_optimizer = WholeMemoryOptimizer(get_global_communicator())
    wm_optimizer.wmb_opt.create_optimizer(str_to_wmb_wholememory_optimizer_type(optimizer_type), param_dict)
    if isinstance(embeddings, WholeMemoryEmbedding):
        wm_optimizer.add_embedding(embeddings)
    else:
        for em in embeddings:
            wm_optimizer.add_embedding(em)
    return wm_optimizer


def destroy_wholememory_optimizer(optimizer: WholeMemoryOptimizer):
    optimizer.wmb_opt.destroy_optimizer()
    optimizer.wmb_opt = None


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


