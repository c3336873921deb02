"""Spawn one process per (virtual) rank -- test helper with the reference's signature
(pylibwholegraph/utils/multiprocess.py: multiprocess_run(world_size, func, inline_single_process))."""
import multiprocessing as mp


def _entry(rank, world_size, func, q):
    try:
        func(rank, world_size)
        q.put((rank, None))
    except BaseException as e:  # noqa: BLE001 - report any failure to the parent
        import traceback

        q.put((rank, traceback.format_exc() + repr(e)))
        raise


def multiprocess_run(world_size: int, func, inline_single_process=False):
    """Run func(rank, world_size) in world_size spawned processes; raises if any rank failed."""
    assert world_size > 0
    if world_size == 1 and inline_single_process:
        func(0, 1)
        return
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entry, args=(r, world_size, func, q)) for r in range(world_size)]
    for p in procs:
        p.start()
    errors = []
    # a rank that dies leaves its peers waiting in a barrier: stop them as soon as one has failed
    import time

    while any(p.is_alive() for p in procs):
        if any((not p.is_alive()) and p.exitcode not in (0, None) for p in procs):
            time.sleep(1.0)
            for p in procs:
                if p.is_alive():
                    p.terminate()
            break
        time.sleep(0.05)
    for p in procs:
        p.join()
    while not q.empty():
        rank, err = q.get()
        if err is not None:
            errors.append((rank, err))
    for r, p in enumerate(procs):
        if p.exitcode != 0 and not any(e[0] == r for e in errors):
            errors.append((r, "exit code %s" % p.exitcode))
    if errors:
        raise RuntimeError("ranks failed: %s" % errors)
