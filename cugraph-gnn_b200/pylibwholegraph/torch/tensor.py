"""WholeMemoryTensor (role of the reference's pylibwholegraph/torch/tensor.py)."""
from typing import List, Union

import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.utils.imports import import_optional
from .comm import WholeMemoryCommunicator
from .utils import (
    torch_dtype_to_wholememory_dtype,
    wholememory_dtype_to_torch_dtype,
    str_to_wmb_wholememory_memory_type,
    str_to_wmb_wholememory_location,
    get_file_size,
    get_part_file_name,
    get_part_file_list,
    view_as_torch,
)
from .wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor

torch = import_optional("torch")


class WholeMemoryTensor(object):
    r"""A 1-D or 2-D tensor striped over the GPUs of the communicator."""

    def __init__(self, wmb_tensor: wmb.PyWholeMemoryTensor):
        self.wmb_tensor = wmb_tensor

    @property
    def dtype(self):
        return wholememory_dtype_to_torch_dtype(self.wmb_tensor.dtype)

    def dim(self):
        return self.wmb_tensor.dim()

    @property
    def shape(self):
        return self.wmb_tensor.shape

    def stride(self):
        return self.wmb_tensor.stride()

    def storage_offset(self):
        return self.wmb_tensor.storage_offset()

    def get_comm(self):
        return WholeMemoryCommunicator(self.wmb_tensor.get_wholememory_handle().get_communicator())

    def gather(self, indice: "torch.Tensor", *, force_dtype: Union["torch.dtype", None] = None):
        """out[i] = self[indice[i]] (rows with a negative index are left uninitialised)."""
        assert indice.dim() == 1
        width = self.shape[1] if self.dim() == 2 else 1
        out = torch.empty(
            (indice.shape[0], width),
            device="cuda:%d" % torch.cuda.current_device(),
            dtype=self.dtype if force_dtype is None else force_dtype,
        )
        wmb.wholememory_gather_op(
            self.wmb_tensor, wrap_torch_tensor(indice), wrap_torch_tensor(out), get_wholegraph_env_fns(), get_stream()
        )
        return out if self.dim() == 2 else out.view(-1)

    def scatter(self, input_tensor: "torch.Tensor", indice: "torch.Tensor"):
        """self[indice[i]] = input_tensor[i]; input may live in pinned host memory."""
        assert indice.dim() == 1
        assert input_tensor.dim() == self.dim()
        assert indice.shape[0] == input_tensor.shape[0]
        if self.dim() == 2:
            assert input_tensor.shape[1] == self.shape[1]
        else:
            input_tensor = input_tensor.unsqueeze(1)
        wmb.wholememory_scatter_op(
            wrap_torch_tensor(input_tensor), wrap_torch_tensor(indice), self.wmb_tensor, get_wholegraph_env_fns(), get_stream()
        )

    def get_sub_tensor(self, starts, ends):
        """ends[i] == -1 means up to the last element of dim i."""
        return WholeMemoryTensor(self.wmb_tensor.get_sub_tensor(starts, ends))

    def get_local_tensor(self, host_view: bool = False):
        """(torch view of this rank's entries, index of its first entry)."""
        if host_view:
            raise NotImplementedError("host views need host-resident WholeMemory, which this build does not provide")
        view, start = self.wmb_tensor.get_local_view()
        return view_as_torch(view), start

    def get_global_tensor(self, host_view: bool = False):
        """Flat view of the whole tensor; only exists when one rank holds everything."""
        if host_view:
            raise NotImplementedError("host views need host-resident WholeMemory, which this build does not provide")
        if self.get_comm().get_size() != 1:
            raise ValueError("a flat global view needs continuous memory; use get_all_chunked_tensor on multi-GPU")
        view, _ = self.wmb_tensor.get_local_view()
        return view_as_torch(view), 0

    def get_all_chunked_tensor(self, host_view: bool = False):
        """([per-rank torch views], [first entry of each rank]) -- peer chunks are P2P mapped."""
        if host_view:
            raise NotImplementedError("host views need host-resident WholeMemory, which this build does not provide")
        world = self.get_comm().get_size()
        pairs = [self.wmb_tensor.get_rank_view(r) for r in range(world)]
        return [view_as_torch(v) for v, _ in pairs], [s for _, s in pairs]

    def from_filelist(self, filelist: Union[List[str], str], round_robin_size: int = 0):
        if isinstance(filelist, str):
            filelist = [filelist]
        self.wmb_tensor.from_filelist(filelist, round_robin_size)

    def from_file_prefix(self, file_prefix: str, part_count: Union[int, None] = None):
        if part_count is None:
            part_count = self.get_comm().get_size()
        self.from_filelist(get_part_file_list(file_prefix, part_count))

    def local_to_file(self, filename: str):
        self.wmb_tensor.to_file(filename)

    def to_file_prefix(self, file_prefix: str):
        comm = self.get_comm()
        self.local_to_file(get_part_file_name(file_prefix, comm.get_rank(), comm.get_size()))


def create_wholememory_tensor(
    comm: WholeMemoryCommunicator,
    memory_type: str,
    memory_location: str,
    sizes: List[int],
    dtype: "torch.dtype",
    strides: List[int],
    tensor_entry_partition: Union[List[int], None] = None,
):
    """Collective: allocate an uninitialised WholeMemory tensor (dim 1 or 2)."""
    ndim = len(sizes)
    if ndim not in (1, 2):
        raise ValueError("Only dim 1 or 2 is supported now.")
    if strides is None:
        strides = [sizes[1], 1] if ndim == 2 else [1]
    else:
        assert len(strides) == ndim and strides[-1] == 1
        assert ndim == 1 or strides[0] >= sizes[1]
    desc = wmb.PyWholeMemoryTensorDescription()
    desc.set_shape(sizes)
    desc.set_stride(strides)
    desc.set_dtype(torch_dtype_to_wholememory_dtype(dtype))
    return WholeMemoryTensor(
        wmb.create_wholememory_tensor(
            desc,
            comm.wmb_comm,
            str_to_wmb_wholememory_memory_type(memory_type),
            str_to_wmb_wholememory_location(memory_location),
            tensor_entry_partition,
        )
    )


def create_wholememory_tensor_from_filelist(
    comm: WholeMemoryCommunicator,
    memory_type: str,
    memory_location: str,
    filelist: Union[List[str], str],
    dtype: "torch.dtype",
    last_dim_size: int = 0,
    last_dim_strides: int = -1,
    tensor_entry_partition: Union[List[int], None] = None,
):
    """Collective: size the tensor from the files (raw row-major binary), then load them."""
    if isinstance(filelist, str):
        filelist = [filelist]
    assert last_dim_size >= 0
    elt = torch.tensor([], dtype=dtype).element_size()
    total = sum(get_file_size(f) for f in filelist)
    if last_dim_size == 0:
        assert total % elt == 0
        sizes, strides = [total // elt], [1]
    else:
        row = elt * last_dim_size
        if total % row != 0:
            raise ValueError("total file size %d is not a multiple of the row size %d" % (total, row))
        if last_dim_strides == -1:
            last_dim_strides = last_dim_size
        sizes, strides = [total // row, last_dim_size], [last_dim_strides, 1]
    t = create_wholememory_tensor(comm, memory_type, memory_location, sizes, dtype, strides, tensor_entry_partition)
    t.from_filelist(filelist)
    return t


def destroy_wholememory_tensor(wm_tensor: WholeMemoryTensor):
    wmb.destroy_wholememory_tensor(wm_tensor.wmb_tensor)
    wm_tensor.wmb_tensor = None
