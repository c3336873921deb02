from .sampler import BaseSampler, SampleIterator, SampleReader, HomogeneousSampleReader, HeterogeneousSampleReader
from .distributed_sampler import BaseDistributedSampler, DistributedNeighborSampler

__all__ = ["BaseSampler", "SampleIterator", "SampleReader", "HomogeneousSampleReader", "HeterogeneousSampleReader", "BaseDistributedSampler",
           "DistributedNeighborSampler"]
