from .dist_tensor import DistTensor, DistEmbedding
from .dist_matrix import DistMatrix
from .utils import empty, is_empty, has_nvlink_network

__all__ = ["DistTensor", "DistEmbedding", "DistMatrix", "empty", "is_empty", "has_nvlink_network"]
