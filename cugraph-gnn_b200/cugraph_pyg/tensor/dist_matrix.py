"""DistMatrix: a COO matrix whose nonzeros are spread over the ranks (role of the reference's
cugraph_pyg/tensor/dist_matrix.py).  Each rank keeps the nonzeros it was given; GraphStore.finalize()
assembles the CSR the sampler reads."""
from typing import Optional, Tuple

import torch


class DistMatrix:
    def __init__(self, src=None, shape: Optional[Tuple[int, int]] = None, dtype: Optional[torch.dtype] = torch.int64,
                 device: Optional[str] = "cuda", backend: Optional[str] = "vmm", format: str = "coo"):
        if format != "coo":
            raise ValueError("Only the COO format is supported")
        self._shape = shape
        self._dtype = dtype
        self._row = torch.empty(0, dtype=dtype, device="cuda")
        self._col = torch.empty(0, dtype=dtype, device="cuda")
        if src is not None:
            self[None] = src

    def __setitem__(self, idx, val):
        """val: [2, nnz] tensor or a (row, col) pair; idx is accepted for API compatibility and ignored
        (nonzeros are appended to this rank's partition)."""
        row, col = (val[0], val[1])
        self._row = torch.cat([self._row, row.to(device="cuda", dtype=self._dtype)])
        self._col = torch.cat([self._col, col.to(device="cuda", dtype=self._dtype)])

    def __getitem__(self, idx: torch.Tensor):
        idx = idx.cuda()
        return self._row[idx], self._col[idx]

    def get_local_tensor(self):
        return self._row, self._col

    @property
    def local_row(self):
        return self._row

    @property
    def local_col(self):
        return self._col

    @property
    def local_coo(self):
        return torch.stack([self._row, self._col])

    @property
    def shape(self):
        return self._shape

    @property
    def dtype(self):
        return self._dtype
