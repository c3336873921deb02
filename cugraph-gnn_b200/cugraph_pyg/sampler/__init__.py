from .sampler import BaseSampler, SampleIterator, SampleReader, HomogeneousSampleReader
from .distributed_sampler import BaseDistributedSampler, DistributedNeighborSampler

__all__ = ["BaseSampler", "SampleIterator", "SampleReader", "HomogeneousSampleReader", "BaseDistributedSampler",
           "DistributedNeighborSampler"]
