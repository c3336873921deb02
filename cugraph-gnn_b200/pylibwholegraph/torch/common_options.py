"""argparse option groups of the WholeGraph example scripts (role of the reference's
pylibwholegraph/torch/common_options.py: same flags, destinations and defaults)."""
from argparse import ArgumentParser

_TRAINING = [
    (("-e", "--epochs"), dict(type=int, dest="epochs", default=24, help="number of epochs")),
    (("-b", "--batchsize"), dict(type=int, dest="batchsize", default=1024, help="batch size")),
    (("--lr",), dict(type=float, dest="lr", default=0.003, help="learning rate")),
    (("--embedding-memory-type",), dict(dest="embedding_memory_type", default="chunked",
                                        help="Embedding memory type, should be: continuous, chunked, distributed, hierarchy")),
    (("--cache-type",), dict(dest="cache_type", default="none",
                             help="Embedding cache type, should be: none, local_device, local_node or all_devices")),
    (("--cache-ratio",), dict(type=float, dest="cache_ratio", default=0.5, help="cache ratio")),
    (("--use-cpp-ext",), dict(action="store_true", dest="use_cpp_ext", default=False,
                              help="Whether to use cpp extension for pytorch (no-op here: callbacks go through ctypes)")),
    (("--train-embedding",), dict(action="store_true", dest="train_embedding", default=False, help="Whether to train embedding")),
    (("--distributed-backend-type",), dict(dest="distributed_backend_type", default="nccl",
                                           help="Distributed backend type, should be: nccl, nvshmem")),
    (("--log-level",), dict(dest="log_level", default="info", help="Logging level of wholegraph, should be: trace, debug, info, warn, error")),
]
_GRAPH = [
    (("-r", "--root-dir"), dict(dest="root_dir", default="dataset", help="graph dataset root directory.")),
    (("--use-global-embedding",), dict(action="store_true", dest="use_global_embedding", default=False,
                                       help="Store embedding across all ranks or only in local node.")),
    (("--feat-dim",), dict(type=int, dest="feat_dim", default=100, help="default feature dim")),
    (("--round-robin-size",), dict(type=int, dest="round_robin_size", default=0, help="continuous embedding size of a rank using round robin shard strategy")),
]
_MODEL = [
    (("--hiddensize",), dict(type=int, dest="hiddensize", default=256, help="hidden size")),
    (("-l", "--layernum"), dict(type=int, dest="layernum", default=3, help="layer number")),
    (("-m", "--model"), dict(dest="model", default="sage", help="model type, valid values are: sage, gcn, gat")),
    (("-f", "--framework"), dict(dest="framework", default="wg", help="framework type, valid values are: pyg, wg")),
    (("--heads",), dict(type=int, dest="heads", default=4, help="num heads")),
    (("-d", "--dropout"), dict(type=float, dest="dropout", default=0.5, help="dropout")),
]
_SAMPLER = [
    (("-n", "--neighbors"), dict(dest="neighbors", default="30,30,30", help="train neighboor sample count")),
    (("-s", "--inferencesample"), dict(type=str, dest="inferencesample", default="30", help="inference sample count, -1 is all")),
]


def _add(argparser: ArgumentParser, table):
    for flags, kw in table:
        argparser.add_argument(*flags, **kw)


def add_training_options(argparser: ArgumentParser):
    _add(argparser, _TRAINING)


def add_common_graph_options(argparser: ArgumentParser):
    _add(argparser, _GRAPH)


def add_common_model_options(argparser: ArgumentParser):
    _add(argparser, _MODEL)


def add_common_sampler_options(argparser: ArgumentParser):
    _add(argparser, _SAMPLER)


def add_node_classfication_options(argparser: ArgumentParser):
    argparser.add_argument("-c", "--classnum", type=int, dest="classnum", default=172, help="class number")


def add_dataloader_options(argparser: ArgumentParser):
    argparser.add_argument("--pickle-data-path", dest="pickle_data_path", default="", help="training data file path, should be pickled dict")
    argparser.add_argument("-w", "--dataloaderworkers", type=int, dest="dataloaderworkers", default=0, help="number of workers for dataloader")


def parse_max_neighbors(num_layer, neighbor_str):
    """'30,30,30' -> [30, 30, 30]; a single number is repeated for every layer."""
    values = [int(x) for x in neighbor_str.split(",")]
    assert len(values) in (1, num_layer)
    return values * num_layer if len(values) == 1 and num_layer > 1 else values
