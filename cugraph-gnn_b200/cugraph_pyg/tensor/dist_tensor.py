"""DistTensor / DistEmbedding (role of the reference's cugraph_pyg/tensor/dist_tensor.py:20-534)."""
from typing import List, Optional, Union

import torch

import pylibwholegraph.torch as wgth
from .utils import create_wg_dist_tensor, get_comm


class DistTensor:
    """A 1-D or 2-D tensor row-sharded over the GPUs of the box; indexing gathers / scatters rows, remote rows
    travel over NVLink from inside the kernel (no collective: only creation is collective)."""

    def __init__(self, src: Optional[torch.Tensor] = None, shape: Optional[Union[list, tuple]] = None,
                 dtype: Optional[torch.dtype] = None, device: Optional[str] = "cpu",
                 partition_book: Optional[List[int]] = None, backend: Optional[str] = "vmm", *args, **kwargs):
        self._tensor = None
        self._device = "cuda"
        self._dtype = dtype
        if src is None:
            if shape is not None:
                if dtype is None:
                    raise ValueError("dtype must be given together with shape")
                self._tensor = create_wg_dist_tensor(list(shape), dtype, device, backend, partition_book)
        else:
            if isinstance(src, str):
                raise NotImplementedError("loading from files goes through DistTensor.from_file")
            self._dtype = src.dtype if dtype is None else dtype
            self._tensor = create_wg_dist_tensor(list(src.shape), self._dtype, device, backend, partition_book)
            self.load_from_global_tensor(src)

    # -- the WholeMemoryTensor behind a 1-D tensor or an embedding
    def _wm(self) -> wgth.WholeMemoryTensor:
        t = self._tensor
        return t.get_embedding_tensor() if isinstance(t, wgth.WholeMemoryEmbedding) else t

    def load_from_global_tensor(self, tensor: torch.Tensor):
        """Every rank holds the full source tensor and copies its own stripe."""
        assert tuple(tensor.shape) == tuple(self.shape)
        local, start = self._wm().get_local_tensor()
        local.copy_(tensor[start:start + local.shape[0]].to(local.device, non_blocking=True))
        torch.cuda.current_stream().synchronize()
        get_comm().barrier()

    def load_from_local_tensor(self, tensor: torch.Tensor):
        """Every rank holds exactly the rows of its own stripe."""
        local, _ = self._wm().get_local_tensor()
        assert tuple(tensor.shape) == tuple(local.shape), "local tensor does not match this rank's partition"
        local.copy_(tensor.to(local.device, non_blocking=True))
        torch.cuda.current_stream().synchronize()
        get_comm().barrier()

    @classmethod
    def from_tensor(cls, tensor: torch.Tensor, device: Optional[str] = "cpu", partition_book=None, backend: Optional[str] = "vmm"):
        return cls(src=tensor, device=device, partition_book=partition_book, backend=backend)

    @classmethod
    def from_file(cls, file_path: str, shape, dtype, device: Optional[str] = "cpu", partition_book=None, backend: Optional[str] = "vmm"):
        out = cls(shape=shape, dtype=dtype, device=device, partition_book=partition_book, backend=backend)
        out._wm().from_filelist([file_path])
        return out

    def __setitem__(self, idx: torch.Tensor, val: torch.Tensor):
        """Rank-local scatter (the reference requires every rank to call it; here it is optional)."""
        assert self._tensor is not None, "Please create WholeGraph tensor first."
        idx = idx.cuda()
        if not val.is_cuda and not val.is_pinned():
            val = val.pin_memory()
        if val.dtype != self.dtype:
            val = val.to(self.dtype)
        self._wm().scatter(val, idx)

    def __getitem__(self, idx: torch.Tensor) -> torch.Tensor:
        assert self._tensor is not None, "Please create WholeGraph tensor first."
        return self._wm().gather(idx if idx.is_cuda else idx.cuda())

    def get_local_tensor(self, host_view: bool = False):
        local, _ = self._wm().get_local_tensor(host_view=host_view)
        return local

    def get_local_offset(self) -> int:
        _, offset = self._wm().get_local_tensor()
        return int(offset)

    def get_comm(self):
        return self._wm().get_comm()

    def dim(self) -> int:
        return self._wm().dim()

    @property
    def shape(self):
        return tuple(self._wm().shape)

    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return self._wm().dtype

    def __repr__(self):
        if self._tensor is None:
            return "<DistTensor: No tensor created>"
        return f"DistTensor(shape={self.shape}, dtype={self.dtype}, device='{self._device}')"


class DistEmbedding(DistTensor):
    """2-D feature table [rows, dim] backed by a WholeMemory embedding (the mini-batch feature fetch)."""

    def __init__(self, src: Optional[torch.Tensor] = None, shape=None, dtype=None, device: Optional[str] = "cpu",
                 partition_book=None, backend: Optional[str] = "vmm", cache_policy=None, gather_sms: Optional[int] = -1,
                 round_robin_size: int = -1, name: Optional[str] = None):
        if cache_policy is not None:
            raise NotImplementedError("embedding caches are outside the B200 hot path")
        self._name = name
        if src is not None and src.dim() != 2:
            raise ValueError("The embedding must be 2D.")
        if shape is not None and len(shape) != 2:
            raise ValueError("The shape of the embedding must be 2D.")
        super().__init__(src, shape, dtype, device, partition_book, backend)
        self._embedding = self._tensor

    @classmethod
    def from_tensor(cls, tensor: torch.Tensor, device: Optional[str] = "cpu", partition_book=None, name=None,
                    cache_policy=None, **kwargs):
        return cls(tensor, device=device, partition_book=partition_book, name=name, cache_policy=cache_policy, **kwargs)

    def __getitem__(self, idx: torch.Tensor) -> torch.Tensor:
        assert self._tensor is not None, "Please create WholeGraph embedding first."
        # the index stays on the device: the reference moves renumber maps to the host and straight back
        # (sampler.py:632 -> dist_tensor.py:513)
        return self._embedding.gather(idx if idx.is_cuda else idx.cuda())

    def set_hot_rows(self, hot_indices: Optional[torch.Tensor]):
        """Replicate the given rows on this GPU (WholeMemoryEmbedding.set_hot_rows): reads of those rows stop crossing
        NVLink.  Rank-local; a snapshot of the rows -- call again after writing to the table.  None drops the replica."""
        assert self._tensor is not None, "Please create WholeGraph embedding first."
        self._embedding.set_hot_rows(hot_indices)

    @property
    def name(self):
        return self._name

    def __repr__(self):
        if self._tensor is None:
            return "<DistEmbedding: No embedding created>"
        return f"DistEmbedding(name={self._name}, shape={self.shape}, dtype={self.dtype}, device='{self._device}')"
