"""pylibwholegraph API surface on top of libwholegraph_b200 (B200-native hot path).

Mirrors the subset of rapidsai/cugraph-gnn's ``pylibwholegraph`` that the sampler / gather hot
path uses.  The native code lives in ``../csrc`` and is reached through
the C ABI declared in ``/include/wholememory``; there is no CPU fallback: importing
``pylibwholegraph.binding.wholememory_binding`` fails loudly if the shared library is missing.
"""
__version__ = "26.10.00+b200"
