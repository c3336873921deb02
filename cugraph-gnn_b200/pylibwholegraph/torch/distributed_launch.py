"""Process launch helpers of the WholeGraph examples (role of the reference's pylibwholegraph/torch/distributed_launch.py):
ranks come from the command line, from the launcher's environment (torchrun / mpirun / srun) or from a local spawn."""
import os
from argparse import ArgumentParser

from pylibwholegraph.utils.imports import import_optional

torch = import_optional("torch")


class _Config(object):
    rank = world_size = local_rank = local_size = -1
    master_addr = ""
    master_port = -1


distributed_config = _Config()


def get_rank():
    return distributed_config.rank


def get_world_size():
    return distributed_config.world_size


def get_local_rank():
    return distributed_config.local_rank


def get_local_size():
    return distributed_config.local_size


def get_master_addr():
    return distributed_config.master_addr


def get_master_port():
    return distributed_config.master_port


def is_main_process():
    return get_rank() == 0


def add_distributed_launch_options(parser: ArgumentParser):
    parser.add_argument("--launch-agent", dest="launch_agent", default="mpi", help="launch agent used, mpi, pytorch or spawn")
    parser.add_argument("--rank", dest="rank", type=int, default=-1, help="command line flag for rank")
    parser.add_argument("--world-size", dest="world_size", type=int, default=-1, help="command line flag for world_size")
    parser.add_argument("--local-rank", dest="local_rank", type=int, default=-1, help="command line flag for local_rank")
    parser.add_argument("--local-size", dest="local_size", type=int, default=-1, help="command line flag for local_size")
    parser.add_argument("--master-addr", dest="master_addr", default="", help="command line flag for master_addr")
    parser.add_argument("--master-port", dest="master_port", type=int, default=-1, help="command line flag for master_port")
    for name, env in (("world-rank", "RANK"), ("world-size", "WORLD_SIZE"), ("local-rank", "LOCAL_RANK"),
                      ("local-size", "LOCAL_WORLD_SIZE"), ("master-addr", "MASTER_ADDR"), ("master-port", "MASTER_PORT")):
        parser.add_argument("--launch-env-name-" + name, dest="launch_env_name_" + name.replace("-", "_"), default=env,
                            help="environment variable name for " + name)


def get_value_from_env(env_name, fill_default=None):
    value = os.environ.get(env_name, fill_default)
    if value is None:
        raise ValueError("Environment variable %s is not set and no default is given" % env_name)
    return value


def get_value_from_option_and_env(option_value, env_name, not_set_value, fill_default=None):
    return option_value if option_value != not_set_value else get_value_from_env(env_name, fill_default)


def _fill_from_launcher(args, env_fallbacks):
    c = distributed_config
    first = lambda names, default=None: next((os.environ[n] for n in names if n in os.environ), default)  # noqa: E731
    c.rank = int(get_value_from_option_and_env(args.rank, args.launch_env_name_world_rank, -1, first(env_fallbacks["rank"], "0")))
    c.world_size = int(get_value_from_option_and_env(args.world_size, args.launch_env_name_world_size, -1, first(env_fallbacks["size"], "1")))
    c.local_rank = int(get_value_from_option_and_env(args.local_rank, args.launch_env_name_local_rank, -1, first(env_fallbacks["local_rank"], str(c.rank))))
    c.local_size = int(get_value_from_option_and_env(args.local_size, args.launch_env_name_local_size, -1, first(env_fallbacks["local_size"], str(c.world_size))))
    c.master_addr = get_value_from_option_and_env(args.master_addr, args.launch_env_name_master_addr, "", "127.0.0.1")
    c.master_port = int(get_value_from_option_and_env(args.master_port, args.launch_env_name_master_port, -1, "12335"))
    os.environ.update(RANK=str(c.rank), WORLD_SIZE=str(c.world_size), LOCAL_RANK=str(c.local_rank), LOCAL_WORLD_SIZE=str(c.local_size),
                      MASTER_ADDR=c.master_addr, MASTER_PORT=str(c.master_port))


def distributed_launch_mpi(args, main_func):
    _fill_from_launcher(args, {"rank": ["OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"], "size": ["OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS"],
                               "local_rank": ["OMPI_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "SLURM_LOCALID"],
                               "local_size": ["OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS", "SLURM_NTASKS_PER_NODE"]})
    main_func()


def distributed_launch_pytorch(args, main_func):
    _fill_from_launcher(args, {"rank": [], "size": [], "local_rank": [], "local_size": []})
    main_func()


def main_spawn_routine(local_rank, main_func, config):
    c = distributed_config
    c.world_size, c.local_size, c.master_addr, c.master_port, node_rank = config
    c.local_rank = local_rank
    c.rank = node_rank * c.local_size + local_rank
    os.environ.update(RANK=str(c.rank), WORLD_SIZE=str(c.world_size), LOCAL_RANK=str(c.local_rank), LOCAL_WORLD_SIZE=str(c.local_size),
                      MASTER_ADDR=c.master_addr, MASTER_PORT=str(c.master_port))
    main_func()


def distributed_launch_spawn(args, main_func):
    local_size = args.local_size if args.local_size > 0 else torch.cuda.device_count()
    world_size = args.world_size if args.world_size > 0 else local_size
    node_rank = (args.rank // local_size) if args.rank >= 0 else 0
    addr = args.master_addr or "127.0.0.1"
    port = args.master_port if args.master_port > 0 else 12335
    config = (world_size, local_size, addr, port, node_rank)
    if local_size > 1:
        torch.multiprocessing.spawn(main_spawn_routine, nprocs=local_size, args=(main_func, config), join=True)
    else:
        main_spawn_routine(0, main_func, config)


def distributed_launch(args, main_func):
    """launch_agent: 'mpi' (ranks from the MPI / Slurm environment), 'pytorch' (torchrun environment) or 'spawn'."""
    assert args.launch_agent in ("mpi", "pytorch", "spawn")
    {"mpi": distributed_launch_mpi, "pytorch": distributed_launch_pytorch, "spawn": distributed_launch_spawn}[args.launch_agent](args, main_func)
