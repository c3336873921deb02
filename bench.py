#!/usr/bin/env python
"""Hot-path benchmark: multi-hop neighbour sampling + renumbering -> feature gather.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the CPU implementation of the same path)

Workload = the configuration BASELINE.json's metric is quoted on ("C4", ogbn-papers100M shape): synthetic RMAT
|V| = 111 M, |E| = 1.6 B (a,b,c,d = .57,.19,.19,.05, seed 42, vertex ids scrambled by an affine bijection as Graph500
does), CSR by destination, int64 row_ptr / int32 col_idx (7.3 GB, replicated per GPU); features fp32 [|V|, 128] with the
closed form table[i][d] = (i + d) & 0xFFFF (56.8 GB, striped over the GPUs of the box, WholeGraph "chunked"); fan-out
[25, 10]; sampler seed 62; pylibcugraph output shape (int64 majors / minors / edge ids).  It fits one B200 (64 GB of 180),
so N = 1 runs it too.  --workload c2 | headline | c5 select the other shapes BASELINE.json names (c5: heterogeneous,
with a DDP-wrapped 2-layer SAGE step so that the NCCL all-reduce of the dense weights is on the timeline).

One step = one call group of 148 mini-batches ("labels": one per SM of a B200) x 1024 seeds per GPU, i.e. what
cugraph_pyg.sampler.DistributedNeighborSampler hands to the native sampler in one call
(python/cugraph-pyg/cugraph_pyg/sampler/distributed_sampler.py:877-908) followed by the feature fetch of
every mini-batch (sampler/sampler.py:51-165 -> FeatureStore -> WholeMemoryEmbedding.gather):
    1. fused multi-hop sampler: per label, hop 1 samples 25 neighbours of the seeds, hop 2 samples 10 neighbours
       of the vertices new in hop 1; per-label renumbering (seeds first, first-occurrence order); COO output
    2. gather the feature row of every vertex of every sampled sub-graph (the concatenated renumber maps).
Every step uses a different seed set; the feature table and the bytes gathered per step exceed L2 many times over.
Before timing, rank 0 runs one 4-label call group through the CPU oracle on the same graph and asserts that the GPU result
is identical (COO, edge ids, renumber map, offsets, gathered rows): "parity_checked" in the line.

metric = sampled edges / s for the whole step (sampling + renumbering + feature gather), whole job over all GPUs.
`value`: K call groups with device-resident seeds, software-pipelined the way the loader runs them (call group k+1 is
begun on a second sampler object before k is finished; for N > 1 the NVLink-bound gather runs on its own stream so the
next call group's sampling kernels execute underneath it).  `e2e`: the same loop with pinned-host seeds in and the
step's result read back to the host every step.
Extra keys: gather_gbs (reference definition: gathered output bytes / gather time,
cpp/bench/wholememory_ops/gather_scatter_bench.cu:352-355), stages (per-stage device times), roofline
(dominant kernel = the gather), roofline_sampler (the sampling + renumbering stage against the same HBM peak),
cpu_baseline (the oracle on the host cores, bounded sample); for N > 1 also striped_only (the same timed loop without the
replicated hot rows: every remote row crosses NVLink) with its NVLink roofline.
"""
import argparse
import contextlib
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))

NUM_NODES = 111_000_000
NUM_EDGES = 1_600_000_000
FEAT_DIM = 128
FANOUT = [25, 10]
BATCH = 1024
LABELS_PER_STEP = 148  # mini-batches per call group (the reference's `local_seeds_per_call` / batch size): one label per SM -- the fused sampler
                       # then runs one CTA per label (profiles/run_r2p.sh: 3.95 us per label against 5.3 at 64 labels per call)
SAMPLER_SEED = 62
RMAT = (0.57, 0.19, 0.19, 0.05)
GATHER_TRAFFIC_FILE = {"register": "r1_gather_traffic.json", "bulk": "r2_gather_bulk_traffic.json"}  # ncu captures (per-unit DRAM bytes) per gather kernel
SAMPLER_TRAFFIC_FILE = "r2_sampler_traffic.json"
SCRAMBLE_MUL, SCRAMBLE_ADD = 7_919_717, 1_234_567  # multiplier coprime with every |V| used here (odd, not a multiple of 3 or 5)
WORKLOAD = "C4 ogbn-papers100M-shape synthetic RMAT |V|=111M |E|=1.6B fanout=[25,10] feat_dim=128 fp32, sampler+renumber+gather"
METRIC = "sampled_edges_per_sec (multi-hop sample + renumber + feature gather, fanout [25,10])"
# --workload: the other shapes BASELINE.json names for this metric (same step, same kernels; not the default bench line).
# c4 (the shape the metric is quoted on) is the default and is what the constants above say; the others overwrite them in main().
WORKLOADS = {
    "c4": None,
    "c2": {"NUM_NODES": 10_000_000, "NUM_EDGES": 160_000_000, "FEAT_DIM": 128,
           "WORKLOAD": "C2 synthetic RMAT |V|=10M |E|=160M fanout=[25,10] feat_dim=128 fp32, sampler+renumber+gather"},
    "tiny": {"NUM_NODES": 200_000, "NUM_EDGES": 3_200_000, "FEAT_DIM": 128,  # CPU-side self-test of the harness (tests/test_bench_cpu.py)
             "WORKLOAD": "tiny synthetic RMAT |V|=200k |E|=3.2M fanout=[25,10] feat_dim=128 fp32 (harness self-test, not a bench line)"},
    "headline": {"NUM_NODES": 100_000_000, "NUM_EDGES": 1_000_000_000, "FEAT_DIM": 256,
                 "WORKLOAD": "north-star synthetic RMAT |V|=100M |E|=1B fanout=[25,10] feat_dim=256 fp32, sampler+renumber+gather"},
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def rmat_csr(torch, num_nodes, num_edges, seed, device, rows=None, cols=None):
    """RMAT edge list generated on `device`, folded into [0, num_nodes), returned as CSR by destination.
    rows / cols = (first id, count): destinations / sources live in those id ranges of a num_nodes-wide id space
    (heterogeneous graphs: one call per edge type); default: both span [0, num_nodes)."""
    r0, rn = rows if rows is not None else (0, num_nodes)
    c0, cn = cols if cols is not None else (0, num_nodes)
    scale = max(1, (max(rn, cn) - 1).bit_length())
    a, b, c, d = RMAT
    g = torch.Generator(device=device).manual_seed(seed)
    chunk = 1 << 24
    keys = torch.empty(num_edges, dtype=torch.int64, device=device)
    for lo in range(0, num_edges, chunk):
        n = min(chunk, num_edges - lo)
        src = torch.zeros(n, dtype=torch.int64, device=device)
        dst = torch.zeros(n, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(n, device=device, generator=g)
            src_bit = (r >= a + b).to(torch.int64)  # quadrants c, d
            dst_bit = (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)  # quadrants b, d
            src = (src << 1) | src_bit
            dst = (dst << 1) | dst_bit
        # RMAT concentrates the edges on small ids; scramble the ids with an affine bijection of [0, V) (Graph500
        # scrambles too) so that hubs are spread over the contiguous row partitions of the feature table instead of all
        # living on rank 0 (measured at N=2 before this: rank 0 gathered 0.54 ms/step, rank 1 1.31 ms/step)
        src = ((src % cn) * SCRAMBLE_MUL + SCRAMBLE_ADD) % cn + c0
        dst = ((dst % rn) * SCRAMBLE_MUL + SCRAMBLE_ADD) % rn + r0
        keys[lo:lo + n] = (dst << 32) | src
        del src, dst
    keys, _ = torch.sort(keys)
    rows = keys >> 32
    col = (keys & 0xFFFFFFFF).to(torch.int32)
    del keys
    row_ptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=device)
    row_ptr[1:] = torch.bincount(rows, minlength=num_nodes).cumsum(0)
    return row_ptr, col


def seed_sets(torch, num_sets, labels, rank=0):
    """num_sets call groups of `labels` mini-batches of BATCH distinct seeds (int64, like NodeLoader's randperm)."""
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    out = []
    if NUM_NODES > 20_000_000 and num_sets * labels * BATCH <= NUM_NODES:
        # the large shapes: ONE permutation cut into call groups (a permutation of 10^8 ids per call group takes seconds)
        perm = torch.randperm(NUM_NODES, generator=g)
        return [perm[i * labels * BATCH:(i + 1) * labels * BATCH].clone() for i in range(num_sets)]
    for _ in range(num_sets):
        out.append(torch.randperm(NUM_NODES, generator=g)[: labels * BATCH].contiguous())
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            time.sleep(0.3)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

        def num(x):
            try:
                return float(x)
            except Exception:
                return None

        sm = sorted(v for v in (num(r[0]) for r in self.rows if r) if v is not None)
        mx = [v for v in (num(r[1]) for r in self.rows if len(r) > 1) if v is not None]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 4 + k and r[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    import pylibwholegraph.binding.wholememory_binding as wmb
    import pylibwholegraph.torch as wgth

    wgth.init(rank, world, local_rank, world)
    comm = wgth.get_global_communicator()
    launch_count = wmb.native_symbol("wholememory_b200_kernel_launch_count")
    launch_count.restype = ctypes.c_ulonglong

    t0 = time.time()
    row_ptr, col = rmat_csr(torch, NUM_NODES, NUM_EDGES, 42, dev)
    torch.cuda.synchronize()
    log("[rank %d] RMAT CSR built in %.1fs, |E|=%d, max degree %d" % (rank, time.time() - t0, col.numel(), int((row_ptr[1:] - row_ptr[:-1]).max())))

    # graph: replicated per GPU (C4: 7.3 GB); features: striped over the GPUs of the box and read by P2P
    one = wgth.create_group_communicator(1, 1) if world > 1 else comm
    wm_rp = wgth.create_wholememory_tensor(one, "chunked", "cuda", [NUM_NODES + 1], torch.int64, [1])
    wm_rp.get_local_tensor()[0].copy_(row_ptr)
    wm_col = wgth.create_wholememory_tensor(one, "chunked", "cuda", [col.numel()], torch.int32, [1])
    wm_col.get_local_tensor()[0].copy_(col)
    del col
    emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [NUM_NODES, FEAT_DIM], gather_sms=args.gather_sms)
    local, start = emb.get_embedding_tensor().get_local_tensor()
    ar = torch.arange(FEAT_DIM, device=dev)[None, :]
    for lo in range(0, local.shape[0], 1 << 20):
        hi = min(local.shape[0], lo + (1 << 20))
        local[lo:hi] = ((torch.arange(start + lo, start + hi, device=dev)[:, None] + ar) & 0xFFFF).float()
    comm.barrier()
    hot_rows = 0
    if world > 1 and args.hot_ratio > 0:
        # replicate the hottest rows on every GPU (role of the reference's device cache for remote tables, static here):
        # hotness of a vertex = how often it appears as a CSR column = how often the sampler can reach it
        hot_rows = int(args.hot_ratio * NUM_NODES)
        hotness = torch.zeros(NUM_NODES, dtype=torch.int64, device=dev)
        col_local = wm_col.get_local_tensor()[0]
        for lo in range(0, col_local.numel(), 1 << 28):  # chunked: a 1.6 B-element int64 copy of col_idx would take 12.8 GB
            hotness += torch.bincount(col_local[lo:lo + (1 << 28)].long(), minlength=NUM_NODES)
        emb.set_hot_rows(torch.argsort(hotness, descending=True)[:hot_rows].contiguous())
        del hotness, col_local
        comm.barrier()
    sampler = wgth.MultiHopSampler()
    labels = args.labels
    label_offsets = (torch.arange(labels + 1, dtype=torch.int64) * BATCH).to(dev)

    # ---- parity, outside every timed region: one 4-label call group on THIS graph through the CPU oracle (test
    # infrastructure, used here as the checker only) must equal what the GPU path returns, bit for bit
    host_graph = None
    parity = None
    if rank == 0 and not args.no_parity_check:
        host_graph = (row_ptr.cpu().numpy(), wm_col.get_local_tensor()[0].cpu().numpy())
        parity = parity_check(torch, wgth, sampler, emb, wm_rp, wm_col, host_graph, dev)
        log("[rank 0] parity vs oracle on the bench graph: %s" % parity)

    n_sets = args.steps + args.warmup
    host_seeds = [s.pin_memory() for s in seed_sets(torch, n_sets, labels, rank)]

    def step(seeds_dev, seed, ev=None):
        if ev:
            ev[0].record()
        res = sampler.sample(wm_rp, wm_col, seeds_dev, label_offsets, FANOUT, seed, int64_ids=True)
        if ev:
            ev[1].record()
        x = emb.gather(res["renumber_map"])
        if ev:
            ev[2].record()
        return int(res["minors"].numel()), int(res["renumber_map"].numel()), x, res

    # End to end = what a loader does per call group through the public API (pylibwholegraph.torch.MultiHopSampler +
    # WholeMemoryEmbedding.gather): pinned host seeds -> H2D -> sampler -> gather -> the step's result read back to the
    # host.  The result that crosses back is what the reference's loader reads on the host per call group
    # (per-batch offsets, sampler/sampler.py:570-575) plus the gathered feature rows of the block's first `labels` vertices
    # (proves the gather ran; the full [n, F] block stays in HBM for the model, as in the reference).
    # The loop is software-pipelined the way cugraph_pyg's loader runs it: call group k+1 is enqueued
    # (sample_async on the second sampler object) before the host waits for the sizes of call group k.
    samplers = [sampler, wgth.MultiHopSampler()]

    copy_stream = torch.cuda.Stream(device=dev)

    def e2e_begin(k):
        # H2D of the step's input, from pinned memory, on a copy stream of its own: it has no dependency, so it runs while the
        # previous call group's sampler kernel is still busy instead of queueing behind it
        with torch.cuda.stream(copy_stream):
            sd = host_seeds[k].to(dev, non_blocking=True)
        torch.cuda.current_stream().wait_stream(copy_stream)
        sd.record_stream(torch.cuda.current_stream())
        return samplers[k & 1].sample_async(wm_rp, wm_col, sd, label_offsets, FANOUT, SAMPLER_SEED + 7 * k, int64_ids=True)

    # Feature fetch on its own stream: the next call group's sampling kernels (latency / random-access bound, little
    # bandwidth) run underneath the bandwidth-bound gather.  Measured at N=1 with 20 steps: 0.845 ms/step against 0.881 on
    # one stream; for N > 1 the gather additionally waits on NVLink while HBM and the SMs idle.  (An earlier measurement that
    # found the overlap slower was taken before the per-stream allocator priming below and had cudaMalloc stalls in it.)
    use_side = True if args.gather_stream < 0 else bool(args.gather_stream)
    side = torch.cuda.Stream(device=dev) if use_side else None
    # what crosses back per step: label_hop_offsets and renumber_map_offsets (int64) + the first gathered row of every label (fp32);
    # pinned, allocated once (cudaHostAlloc inside the loop stalls), three DMA copies per step and no conversion kernels
    host_ring = [(torch.empty(labels * len(FANOUT) + 1, dtype=torch.int64, pin_memory=True), torch.empty(labels + 1, dtype=torch.int64, pin_memory=True),
                  torch.empty((labels, FEAT_DIM), dtype=torch.float32, pin_memory=True)) for _ in range(4)]
    ring_pos = [0]

    dbg_events = []

    def mark():
        if dbg:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            dbg_events.append(e)

    def e2e_end(pending):
        mark()
        res = pending.result()
        mark()
        if side is not None:  # feature fetch on its own stream: overlaps the next call group's sampling kernels
            side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
            x = emb.gather(res["renumber_map"])
            mark()
            host = host_ring[ring_pos[0] % len(host_ring)]
            ring_pos[0] += 1
            host[0].copy_(res["label_hop_offsets"], non_blocking=True)
            host[1].copy_(res["renumber_map_offsets"], non_blocking=True)
            # the gathered rows of the first `labels` vertices of the block (seeds of the first mini-batch): a plain DMA slice.
            # (An index_select of every label's first row is a kernel, and a kernel waits for an SM while the next call group's
            # sampler holds all of them -- measured 0.36 ms of queueing per step on the gather's stream.)
            host[2].copy_(x[:labels], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            mark()
        return int(res["minors"].numel()), (host, res, x), ev

    dbg = os.environ.get("BENCH_E2E_DEBUG")

    def e2e_loop(first_k, n):
        """n call groups, one sampler call in flight ahead of the gather; returns (edges, d2h bytes per step)."""
        edges, nbytes = 0, 0
        pend = e2e_begin(first_k)
        done = []
        for i in range(n):
            t0 = time.perf_counter()
            nxt = e2e_begin(first_k + i + 1) if i + 1 < n else None
            t1 = time.perf_counter()
            e, host, ev = e2e_end(pend)
            if dbg:
                log("[rank %d] step %d begin %.3f ms  end %.3f ms" % (rank, i, 1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t1)))
            edges += e
            done.append((host, ev))
            if len(done) > 1:  # read the previous step's result on the host while this one runs
                h0, ev0 = done.pop(0)
                ev0.synchronize()
                nbytes = sum(t.numel() * t.element_size() for t in h0[0])
                assert int(h0[0][0][0]) == 0 and int(h0[0][1][0]) == 0
            pend = nxt
        for h0, ev0 in done:
            ev0.synchronize()
            nbytes = sum(t.numel() * t.element_size() for t in h0[0])
        return edges, nbytes

    dev_seeds = [s.to(dev) for s in host_seeds]
    torch.cuda.synchronize()
    # clocks / throttle reasons are sampled from before the warm-up to after the last timed region: nvidia-smi takes a few
    # hundred ms to come up and stalls CUDA calls while it does (measured in the C5 step: 300 ms on the first timed step)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    n_max = e_max = 0
    for w in range(args.warmup):
        e, n, x, _ = step(dev_seeds[w], SAMPLER_SEED + 7 * w)
        n_max, e_max = max(n_max, n), max(e_max, e)
    del x
    e2e_loop(0, args.warmup)
    torch.cuda.synchronize()
    # Output sizes vary by ~1 % from call group to call group.  torch's caching allocator serves a request from a cached
    # block that is large enough, but the first request ABOVE everything seen so far goes to cudaMalloc, which takes
    # 3-40 ms for a GB-sized block once NCCL has enabled peer access (measured on 2 GPUs: always the step with the
    # largest sample).  A loader that runs for more than a few steps never sees this; K = 10 timed steps would.  Park
    # blocks 25 % above the warm-up maxima in the allocator (4 of each: up to three call groups are alive in the pipelined
    # loop -- one being built, two whose results the host has not read yet).
    prime = []
    block_bytes = int(n_max * 1.25) * FEAT_DIM * 4
    free_bytes = torch.cuda.mem_get_info(dev)[0] + torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
    # (the north-star shape leaves ~30 GB beside its 102 GB table: parking more feature blocks than fit makes torch's allocator
    # release and re-request them every other step, profiles/r2aw_bench_headline_n1_e2e_debug.json)
    n_feature_blocks = max(2, min(4, int(0.6 * free_bytes / max(block_bytes, 1))))
    for i in range(4):
        # the feature block is allocated on the stream that runs the gather: torch keeps one pool per stream
        if i < n_feature_blocks:
            with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
                prime.append(torch.empty((int(n_max * 1.25), FEAT_DIM), dtype=torch.float32, device=dev))
        prime.append(torch.empty(int(n_max * 1.25), dtype=torch.int64, device=dev))
        prime += [torch.empty(int(e_max * 1.25), dtype=torch.int64, device=dev) for _ in range(3)]
    if side is not None and n_feature_blocks == 4:  # the synchronous per-stage pass gathers on the main stream: its pool gets a block too
        prime.append(torch.empty((int(n_max * 1.25), FEAT_DIM), dtype=torch.float32, device=dev))
    del prime
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    # ---- device-resident timing: `value` -------------------------------------------------------------

    def begin_dev(k):
        return samplers[k & 1].sample_async(wm_rp, wm_col, dev_seeds[k], label_offsets, FANOUT, SAMPLER_SEED + 7 * k, int64_ids=True)

    def value_loop():
        """The timed region of `value`: K call groups, device-resident seeds, software-pipelined like the loader (call group
        k+1 is begun before k is finished); gather events are recorded on the stream the gather is launched on."""
        launches0 = int(launch_count())
        gev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
        t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_begin.record()
        tot_edges = tot_nodes = 0
        alive = []
        pend = begin_dev(args.warmup)
        for i in range(args.steps):
            nxt = begin_dev(args.warmup + i + 1) if i + 1 < args.steps else None
            res = pend.result()
            if side is not None:
                side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
                gev[i][0].record()
                x = emb.gather(res["renumber_map"])
                gev[i][1].record()
            alive.append((res, x))  # outputs live until the stream that reads them has passed (3 call groups in flight at most)
            if len(alive) > 3:
                alive.pop(0)
            tot_edges += int(res["minors"].numel())
            tot_nodes += int(res["renumber_map"].numel())
            pend = nxt
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        t_end.record()
        torch.cuda.synchronize()
        del alive
        return {"ms": t_begin.elapsed_time(t_end), "gather_ms": sum(ev[0].elapsed_time(ev[1]) for ev in gev),
                "edges": tot_edges, "nodes": tot_nodes, "launches": int(launch_count()) - launches0}

    main_run = value_loop()
    ms_total, gather_ms, tot_edges, tot_nodes, launches = (main_run[k] for k in ("ms", "gather_ms", "edges", "nodes", "launches"))
    # (b) per-stage times, one synchronous pass over the same call groups (not part of `value`)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for k in range(args.steps):
        step(dev_seeds[args.warmup + k], SAMPLER_SEED + 7 * (args.warmup + k), evs[k])
    torch.cuda.synchronize()
    sample_ms = sum(ev[0].elapsed_time(ev[1]) for ev in evs)
    gather_alone_ms = sum(ev[1].elapsed_time(ev[2]) for ev in evs)

    # ---- end-to-end: pinned host seeds in, per-batch sizes + feature checksum out, every step -----------
    if world > 1:
        dist.barrier()
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e_begin.record()
    e2e_edges, d2h_bytes = e2e_loop(args.warmup, args.steps)
    e_end.record()
    torch.cuda.synchronize()
    e2e_ms = e_begin.elapsed_time(e_end)
    if dbg:
        ev4 = dbg_events[-4 * args.steps:]
        log("[rank %d] e2e device phases per step (before-result | result->gather done | gather->metric done): %s" % (
            rank, " ".join("%.2f|%.2f|%.2f" % (ev4[4 * i].elapsed_time(ev4[4 * i + 1]), ev4[4 * i + 1].elapsed_time(ev4[4 * i + 2]),
                                               ev4[4 * i + 2].elapsed_time(ev4[4 * i + 3])) for i in range(args.steps))))
        log("[rank %d] e2e %.3f ms for %d steps; value loop %.3f ms (sample %.3f, gather %.3f)" % (rank, e2e_ms, args.steps, ms_total, sample_ms, gather_ms))
    clock_info = clocks.stop() if rank == 0 else None

    # ---- N > 1: the same timed loop with the replica dropped (every remote row crosses NVLink), reported beside `value`
    striped = None
    if world > 1 and hot_rows > 0:
        emb.set_hot_rows(None)
        comm.barrier()
        value_loop()  # one untimed pass: allocator / clocks settle on the new gather durations
        striped = value_loop()
    stats = torch.tensor([ms_total, e2e_ms, sample_ms, gather_ms, gather_alone_ms] + ([striped["ms"], striped["gather_ms"]] if striped else [0.0, 0.0]),
                         dtype=torch.float64, device=dev)
    counts = torch.tensor([tot_edges, tot_nodes, e2e_edges] + ([striped["edges"], striped["nodes"]] if striped else [0, 0]), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms, sample_ms, gather_ms, gather_alone_ms, striped_ms, striped_gather_ms = stats.tolist()
    tot_edges, tot_nodes, e2e_edges, striped_edges, striped_nodes = counts.tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        row_bytes = FEAT_DIM * 4

        def captured(name, per_key, units):
            """DRAM bytes per unit of work from a committed `ncu --set full` capture (profiles/<name>), scaled to this run."""
            try:
                cap = json.load(open(os.path.join(ROOT, "profiles", name)))
                per = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) / cap[per_key]
                return per * units, "ncu dram__bytes_read+write per %s (%s) x this run's count per step" % (per_key.split("_")[0], cap["source"])
            except Exception:
                return None, None

        env_bulk = os.environ.get("WGB_GATHER_BULK", "")
        gather_is_bulk = env_bulk != "0"  # the library's dispatch rule (gather_scatter.cu: bulk_enabled)
        rows_per_rank = tot_nodes / world
        edges_per_rank = tot_edges / world
        traffic, traffic_src = captured(GATHER_TRAFFIC_FILE["bulk" if gather_is_bulk else "register"], "rows_in_launch", rows_per_rank / args.steps)
        gather_alg_bytes = (2 * row_bytes + 8) * rows_per_rank  # per rank, all steps
        gather_achieved = gather_alg_bytes / (gather_ms * 1e-3) / 1e9
        # sampler algorithmic bytes (pylibcugraph shape = what this run returns): per sampled edge col 8 (4 read here:
        # int32 col_idx) + minor 8 + major 8 + edge id 8 = 32 B, per frontier vertex id 8 + row_ptr 16 = 24 B
        sample_alg_bytes = 32.0 * edges_per_rank + 24.0 * rows_per_rank
        s_traffic, s_traffic_src = captured(SAMPLER_TRAFFIC_FILE, "edges_in_step", edges_per_rank / args.steps)
        sampler_achieved = sample_alg_bytes / (sample_ms * 1e-3) / 1e9
        step_alg_bytes = gather_alg_bytes + sample_alg_bytes
        out = {
            "metric": METRIC,
            "value": tot_edges / (ms_total * 1e-3),
            "unit": "edges/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "int64 ids (int32 col_idx in the CSR), fp32 features",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "labels_per_step_per_gpu": labels, "seeds_per_label": BATCH,
                       "l2": "inputs larger than L2 (%.1f GB table, ~1 GB gathered per step); new seed set every step" % (NUM_NODES * FEAT_DIM * 4 / 1e9),
                       "graph": "replicated per GPU", "features": "chunked over %d GPU(s), in-kernel P2P gather" % world,
                       "output_shape": "pylibcugraph (int64 majors, minors, edge ids, renumber map)",
                       "hot_rows_replicated_per_gpu": hot_rows if world > 1 else "n/a at N=1 (every row is local)",
                       "hot_rows_bytes_per_gpu": hot_rows * FEAT_DIM * 4 + (4 * NUM_NODES if hot_rows else 0),
                       "pipeline": "call group k+1 begun before k is finished; gather on %s" % ("its own stream" if use_side else "the same stream"),
                       "gather_sms": args.gather_sms},
            "parity_checked": bool(parity and parity.get("ok")),
            "parity": parity,
            "gather_gbs": row_bytes * tot_nodes / (gather_ms * 1e-3) / 1e9,
            "stages": {
                "sample_renumber_ms_per_step": sample_ms / args.steps,
                "gather_ms_per_step": gather_ms / args.steps,
                "gather_alone_ms_per_step": gather_alone_ms / args.steps,
                "note": "gather_ms: inside the timed (pipelined) region; sample_renumber_ms and gather_alone_ms: separate synchronous pass",
                "edges_per_step_per_gpu": tot_edges / args.steps / world,
                "nodes_gathered_per_step_per_gpu": tot_nodes / args.steps / world,
                "sample_stage_edges_per_sec_per_gpu": tot_edges / world / (sample_ms * 1e-3),
            },
            "roofline": {"kernel": "feature gather: %s" % ("rows_bulk_gather_kernel (cp.async.bulk tile rings; local rows, replica rows and peer rows over NVLink)" if gather_is_bulk
                                                             else "rows_copy_kernel (register path)"),
                         "bound": "hbm", "achieved": gather_achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": gather_achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": gather_alg_bytes / args.steps, "peak_source": peak_src,
                         "timed": "CUDA events on the gather's stream inside the timed region, where the next call group's sampler kernel runs beside it",
                         "achieved_alone": gather_alg_bytes / (gather_alone_ms * 1e-3) / 1e9, "frac_alone": gather_alg_bytes / (gather_alone_ms * 1e-3) / 1e9 / hbm_peak,
                         "alone": "the same launches in the separate synchronous pass (nothing else on the GPU)"},
            "roofline_sampler": {"kernel": "sampling + renumbering stage (all kernels of one MultiHopSampler call, synchronous pass)", "bound": "hbm",
                                 "achieved": sampler_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": sampler_achieved / hbm_peak,
                                 "traffic": s_traffic, "traffic_source": s_traffic_src,
                                 "wasted_traffic_ratio": (s_traffic / (sample_alg_bytes / args.steps)) if s_traffic else None,
                                 "algorithmic_bytes_per_step": sample_alg_bytes / args.steps,
                                 "algorithmic_bytes": "32 B per sampled edge + 24 B per frontier vertex"},
            "roofline_step": {"bound": "hbm", "achieved": step_alg_bytes / (ms_total * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                              "frac": step_alg_bytes / (ms_total * 1e-3) / 1e9 / hbm_peak,
                              "note": "algorithmic bytes of sampler + gather over the pipelined step time"},
            "e2e": {"value": e2e_edges / (e2e_ms * 1e-3), "unit": "edges/s", "h2d_bytes_per_step": labels * BATCH * 8, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "clocks": clock_info,
        }
        if striped:
            # per-GPU NVLink bound of a striped table: (W-1)/W of the gathered bytes arrive over NVLink at the measured 770 GB/s
            out_bytes_rank = row_bytes * striped_nodes / world
            nvl_ms = out_bytes_rank * (world - 1) / world / 770e9 * 1e3
            out["striped_only"] = {"value": striped_edges / (striped_ms * 1e-3), "unit": "edges/s", "ms_per_step": striped_ms / args.steps,
                                   "gather_ms_per_step": striped_gather_ms / args.steps,
                                   "gather_gbs_per_gpu": out_bytes_rank / (striped_gather_ms * 1e-3) / 1e9,
                                   "roofline": {"bound": "nvlink", "peak": 770.0 * world / (world - 1), "unit": "GB/s of gathered output per GPU",
                                                "achieved": out_bytes_rank / (striped_gather_ms * 1e-3) / 1e9,
                                                "frac": nvl_ms / striped_gather_ms, "peak_source": "measured peer copy 770 GB/s per direction (B200_PROFILING.md)"},
                                   "note": "same timed loop, replica dropped: every remote row crosses NVLink"}
        if world == 1 and not args.no_cpu_baseline:
            if host_graph is None:
                host_graph = (row_ptr.cpu().numpy(), wm_col.get_local_tensor()[0].cpu().numpy())
            out["cpu_baseline"] = cpu_baseline(host_graph[0], host_graph[1], budget_s=15.0, labels=labels)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# --workload c5: heterogeneous call groups + a DDP-wrapped GNN step (BASELINE.json configs[4])
# ----------------------------------------------------------------------------------------------------
C5_NODES = (25_000_000, 25_000_000)                    # vertex types "a", "b": global ids [0, 25 M) and [25 M, 50 M)
C5_EDGE_TYPES = (("a", "a"), ("b", "a"), ("a", "b"))     # (source type, destination type); CSR rows are destinations
C5_EDGES = 800_000_000
C5_FANOUT = [10, 10]                                     # per edge type and hop -> h_fan_out [hop * T + etype] (neighbor_loader.py:192-201)
C5_HIDDEN, C5_CLASSES = 256, 47
C5_WORKLOAD = ("C5 heterogeneous synthetic: 3 edge types over |V|=50M (2 vertex types), |E|=800M, per-type fan-out [10,10], feat_dim=128 fp32, "
               "hetero sampler + striped P2P gather + 2-layer SAGE step (NCCL all-reduce of the dense weights)")


def run_c5(args):
    """One step per GPU = one heterogeneous call group (args.labels mini-batches x 1024 seeds of type "a") through the
    fused hetero sampler, the feature rows of every sampled vertex from the table striped over the GPUs (in-kernel P2P,
    no collective), then forward + backward of a 2-layer GraphSAGE (pylibwholegraph.torch.SAGEConv: CSR aggregation kernel +
    dense layers) over the sampled block, wrapped in DistributedDataParallel so that the NCCL all-reduce of the dense
    weights is on the timeline (reference loop: examples/gcn_dist_mnmg.py:233-251, 427).  Reports per-stage device times."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    import pylibwholegraph.binding.wholememory_binding as wmb
    import pylibwholegraph.torch as wgth
    from pylibwholegraph.torch.aggregate import SAGEConv, csr_transpose

    wgth.init(rank, world, local_rank, world)
    comm = wgth.get_global_communicator()
    launch_count = wmb.native_symbol("wholememory_b200_kernel_launch_count")
    launch_count.restype = ctypes.c_ulonglong
    one = wgth.create_group_communicator(1, 1) if world > 1 else comm
    V = sum(C5_NODES)
    vto = [0, C5_NODES[0], V]
    rng_of = {"a": (0, C5_NODES[0]), "b": (C5_NODES[0], C5_NODES[1])}
    T = len(C5_EDGE_TYPES)
    t0 = time.time()
    wm_rps, wm_cols = [], []
    for t, (st, dt) in enumerate(C5_EDGE_TYPES):
        rp, col = rmat_csr(torch, V, C5_EDGES // T, 42 + t, dev, rows=rng_of[dt], cols=rng_of[st])
        w_rp = wgth.create_wholememory_tensor(one, "chunked", "cuda", [V + 1], torch.int64, [1])
        w_rp.get_local_tensor()[0].copy_(rp)
        w_col = wgth.create_wholememory_tensor(one, "chunked", "cuda", [col.numel()], torch.int32, [1])
        w_col.get_local_tensor()[0].copy_(col)
        wm_rps.append(w_rp)
        wm_cols.append(w_col)
        del rp, col
    torch.cuda.synchronize()
    log("[rank %d] 3 typed RMAT CSRs built in %.1fs" % (rank, time.time() - t0))
    emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [V, FEAT_DIM])
    local, start = emb.get_embedding_tensor().get_local_tensor()
    ar = torch.arange(FEAT_DIM, device=dev)[None, :]
    for lo in range(0, local.shape[0], 1 << 20):
        hi = min(local.shape[0], lo + (1 << 20))
        local[lo:hi] = ((torch.arange(start + lo, start + hi, device=dev)[:, None] + ar) & 0xFF).float() / 256.0
    comm.barrier()
    hot_rows = 0
    if world > 1 and args.hot_ratio > 0:
        # as in the homogeneous bench: the rows reached most often (in-degree as a CSR column, over the three edge types) are
        # replicated on every GPU, the rest is read from its owner over NVLink
        hot_rows = int(args.hot_ratio * V)
        hotness = torch.zeros(V, dtype=torch.int64, device=dev)
        for w_col in wm_cols:
            c_local = w_col.get_local_tensor()[0]
            for lo in range(0, c_local.numel(), 1 << 28):
                hotness += torch.bincount(c_local[lo:lo + (1 << 28)].long(), minlength=V)
        emb.set_hot_rows(torch.argsort(hotness, descending=True)[:hot_rows].contiguous())
        del hotness
        comm.barrier()

    labels = args.labels
    fan = [f for f in C5_FANOUT for _ in range(T)]  # [hop * T + etype]
    L, Vt = len(C5_FANOUT), 2
    lo_dev = (torch.arange(labels + 1, dtype=torch.int64) * BATCH).to(dev)
    g = torch.Generator(device="cpu").manual_seed(4321 + rank)
    n_sets = args.steps + args.warmup
    seed_sets_ = [torch.randint(0, C5_NODES[0], (labels * BATCH,), generator=g).to(dev) for _ in range(n_sets)]
    sampler = wgth.MultiHopSampler()
    row_t = torch.tensor([0 if dt == "a" else 1 for _, dt in C5_EDGE_TYPES], device=dev)  # vertex type of the majors (CSR rows)
    col_t = torch.tensor([0 if st == "a" else 1 for st, _ in C5_EDGE_TYPES], device=dev)

    # ---- parity, outside every timed region: one 4-label heterogeneous call group on THIS graph through the CPU oracle (checker
    # only) must equal the GPU result bit for bit -- typed COO, edge ids, renumber maps, every offset array -- and the gathered rows
    parity = None
    if rank == 0 and not args.no_parity_check:
        import numpy as np

        oracle = _oracle()
        h_rps = [w.get_local_tensor()[0].cpu().numpy() for w in wm_rps]
        h_cols = [w.get_local_tensor()[0].cpu().numpy() for w in wm_cols]
        gp = torch.Generator(device="cpu").manual_seed(99)
        p_seeds = torch.randint(0, C5_NODES[0], (4 * BATCH,), generator=gp)
        p_lo = (np.arange(5) * BATCH).astype(np.int64)
        got = sampler.sample_hetero(wm_rps, wm_cols, vto, p_seeds.to(dev), torch.from_numpy(p_lo).to(dev), fan, SAMPLER_SEED, int64_ids=True)
        exp = oracle.hetero_multihop_sample(h_rps, h_cols, np.asarray(vto, dtype=np.int64), p_seeds.numpy(), p_lo, fan, SAMPLER_SEED)
        keys = [k for k in exp if k in got]
        bad = [k for k in keys if not np.array_equal(got[k].cpu().numpy(), exp[k])]
        xg = emb.gather(got["renumber_map"]).cpu().numpy()
        ids = exp["renumber_map"].astype(np.int64)
        want = (((ids[:, None] + np.arange(FEAT_DIM)[None, :]) & 0xFF).astype(np.float32)) / 256.0
        if not np.array_equal(xg, want):
            bad.append("gathered_features")
        assert not bad and len(keys) >= 6, "GPU result differs from the oracle on the C5 graph: %s (compared %s)" % (bad, keys)
        parity = {"ok": True, "labels": 4, "edges": int(exp["minors"].shape[0]), "nodes": int(ids.shape[0]), "compared": keys + ["gathered_features"]}
        log("[rank 0] parity vs oracle on the C5 graph: %s" % parity)
        del h_rps, h_cols, got, xg

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.c1 = SAGEConv(FEAT_DIM, C5_HIDDEN)
            self.c2 = SAGEConv(C5_HIDDEN, C5_CLASSES)

        def forward(self, x, indptr, indices, indptr_t, indices_t):
            h = torch.relu(self.c1(x, indptr, indices, (indptr_t, indices_t)))
            return self.c2(h, indptr, indices, (indptr_t, indices_t))

    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = True  # the dense layers are the model's business (cuBLAS), not this repo's kernels
    net = Net().to(dev)
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local_rank]) if world > 1 else net
    opt = torch.optim.SGD(model.parameters(), lr=0.01)
    n_params = sum(p.numel() for p in net.parameters())

    def build_block(res):
        """typed local ids -> row numbers of the gathered feature matrix (= positions in the concatenated renumber maps,
        segment (label, vertex type) at renumber_map_offsets[label * Vt + vtype]), then a CSR by destination."""
        ltho, rmo = res["label_type_hop_offsets"], res["renumber_map_offsets"]
        E = int(res["minors"].numel())
        group = torch.searchsorted(ltho[1:].contiguous(), torch.arange(E, device=dev), right=True)  # edge -> (label, type, hop)
        lab, typ = group // (T * L), (group // L) % T
        dst = rmo[lab * Vt + row_t[typ]] + res["majors"]
        src = rmo[lab * Vt + col_t[typ]] + res["minors"]
        n = int(res["renumber_map"].numel())
        order = torch.argsort(dst)
        indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        indptr[1:] = torch.bincount(dst, minlength=n).cumsum(0)
        indices = src[order].to(torch.int32).contiguous()
        return (indptr, indices) + csr_transpose(indptr, indices, n) + (n,)  # + the reversed block for the backward pass

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def step(k, marks=None, sync_grads=True):
        def mark():
            if marks is not None:
                e = ev()
                e.record()
                marks.append(e)

        mark()
        res = sampler.sample_hetero(wm_rps, wm_cols, vto, seed_sets_[k], lo_dev, fan, SAMPLER_SEED + 7 * k, int64_ids=True)
        mark()
        x = emb.gather(res["renumber_map"])
        mark()
        indptr, indices, indptr_t, indices_t, n = build_block(res)
        mark()
        out = model(x, indptr, indices, indptr_t, indices_t)
        # "labels": a closed form of the global id, on the seed rows of every mini-batch (first rows of each (label, "a") segment)
        seed_rows = (res["renumber_map_offsets"][0::Vt][:-1, None] if Vt > 1 else res["renumber_map_offsets"][:-1, None])
        seed_rows = (seed_rows + torch.arange(BATCH, device=dev)[None, :]).reshape(-1)
        seed_rows = seed_rows[seed_rows < n]
        target = res["renumber_map"][seed_rows] % C5_CLASSES
        loss = torch.nn.functional.cross_entropy(out[seed_rows], target)
        mark()
        opt.zero_grad(set_to_none=True)
        with contextlib.nullcontext() if (sync_grads or world == 1) else model.no_sync():
            loss.backward()  # DDP: the reducer's bucketed NCCL all-reduce of the dense weights runs inside backward
        opt.step()
        mark()
        return int(res["minors"].numel()), n, loss

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()  # before the warm-up: nvidia-smi takes a few hundred ms to come up and must not do that inside the timed region
    for w in range(args.warmup):
        step(w)
    # Block sizes differ from call group to call group; the first time a step needs a larger activation block than any
    # before, torch's caching allocator goes to cudaMalloc, which takes tens of ms once NCCL has enabled peer access
    # (measured: 40 ms backward on first sight of a shape against 10.5 ms afterwards).  A training run sees every shape
    # within a few steps; a 10-step measurement would time the allocator, so the timed call groups pass once untimed.
    for _ in range(2):  # (twice: one pass did not always leave every block the backward pass asks for in the cache -- a 70 ms
        for k in range(args.steps):  # cudaMalloc showed up inside one timed step of a 2-GPU run, profiles/r2bb_bench_c5_n2.json)
            step(args.warmup + k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = int(launch_count())
    marks = []
    t_b, t_e = ev(), ev()
    torch.cuda.synchronize()
    t_b.record()
    edges = nodes = 0
    for k in range(args.steps):
        e, n, loss = step(args.warmup + k, marks)
        edges += e
        nodes += n
    t_e.record()
    torch.cuda.synchronize()
    ms_total = t_b.elapsed_time(t_e)
    launches = int(launch_count()) - launches0
    names = ["sample_renumber", "gather", "block_build_torch", "forward_aggregation_dense", "backward_allreduce_optimizer"]
    stage = [0.0] * 5
    per_step = sorted(marks[6 * k].elapsed_time(marks[6 * k + 5]) for k in range(args.steps))
    median_step_ms = per_step[len(per_step) // 2]
    for k in range(args.steps):
        m = marks[6 * k: 6 * k + 6]
        for j in range(5):
            stage[j] += m[j].elapsed_time(m[j + 1])
        log("[rank %d] step %d: %s" % (rank, k, " ".join("%.2f" % m[j].elapsed_time(m[j + 1]) for j in range(5))))
    # the same steps without the gradient all-reduce (DDP no_sync): the difference is what the collective costs on the timeline
    nosync_bwd = 0.0
    if world > 1:
        dist.barrier()
        m2 = []
        for k in range(args.steps):
            step(args.warmup + k, m2, sync_grads=False)
        torch.cuda.synchronize()
        nosync_bwd = sum(m2[6 * k + 4].elapsed_time(m2[6 * k + 5]) for k in range(args.steps)) / args.steps
    # the aggregation kernel by itself on the last step's block (first layer: 128-wide rows)
    from pylibwholegraph.torch.aggregate import csr_aggregate_forward
    res = sampler.sample_hetero(wm_rps, wm_cols, vto, seed_sets_[-1], lo_dev, fan, SAMPLER_SEED, int64_ids=True)
    xg = emb.gather(res["renumber_map"])
    indptr, indices, indptr_t, indices_t, n_blk = build_block(res)
    gh = torch.randn((n_blk, C5_HIDDEN), device=dev)
    csr_aggregate_forward(indptr_t, indices_t, gh, "sum")
    b0, b1 = ev(), ev()
    b0.record()
    for _ in range(3):
        csr_aggregate_forward(indptr_t, indices_t, gh, "sum")
    b1.record()
    torch.cuda.synchronize()
    agg_bwd_ms = b0.elapsed_time(b1) / 3
    max_in, max_out = int((indptr[1:] - indptr[:-1]).max()), int((indptr_t[1:] - indptr_t[:-1]).max())
    del gh, indptr_t, indices_t
    csr_aggregate_forward(indptr, indices, xg, "mean")
    g0, g1 = ev(), ev()
    torch.cuda.synchronize()
    g0.record()
    for _ in range(10):
        csr_aggregate_forward(indptr, indices, xg, "mean")
    g1.record()
    torch.cuda.synchronize()
    agg_ms = g0.elapsed_time(g1) / 10
    agg_bytes = int(indices.numel()) * (FEAT_DIM * 4 + 4) + n_blk * FEAT_DIM * 4
    del xg, indptr, indices
    # the all-reduce by itself: a tensor of the model's size, K times
    ar_ms = 0.0
    if world > 1:
        buf = torch.zeros(n_params, device=dev)
        dist.all_reduce(buf)
        a0, a1 = ev(), ev()
        torch.cuda.synchronize()
        a0.record()
        for _ in range(args.steps):
            dist.all_reduce(buf)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / args.steps
    clock_info = clocks.stop() if rank == 0 else None
    stats = torch.tensor([ms_total, ar_ms, nosync_bwd] + stage, dtype=torch.float64, device=dev)
    counts = torch.tensor([edges, nodes], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    ms_total, ar_ms, nosync_bwd = stats[0].item(), stats[1].item(), stats[2].item()
    stage = stats[3:].tolist()
    edges, nodes = counts.tolist()
    if rank == 0:
        out = {
            "metric": "sampled_edges_per_sec (heterogeneous multi-hop sample + renumber + feature gather + SAGE step, per-type fanout [10,10])",
            "value": edges / (ms_total * 1e-3), "unit": "edges/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "median_step_ms_rank0": median_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64 ids (int32 col_idx in the CSRs), fp32 features and model", "data": "synthetic",
            "config": {"workload": C5_WORKLOAD, "labels_per_step_per_gpu": labels, "seeds_per_label": BATCH, "edge_types": ["%s->%s" % et for et in C5_EDGE_TYPES],
                       "h_fan_out": fan, "graph": "3 CSRs over one id space, replicated per GPU", "features": "chunked over %d GPU(s), in-kernel P2P gather" % world,
                       "hot_rows_replicated_per_gpu": hot_rows,
                       "model": "2-layer GraphSAGE %d-%d-%d (pylibwholegraph.torch.SAGEConv), %d dense parameters, %s" % (
                           FEAT_DIM, C5_HIDDEN, C5_CLASSES, n_params, "DistributedDataParallel (NCCL)" if world > 1 else "single process"),
                       "nccl_ranks": world},
            "parity_checked": bool(parity and parity.get("ok")), "parity": parity,
            "stages_ms_per_step": {nm: v / args.steps for nm, v in zip(names, stage)},
            "aggregation_kernel_alone": {"ms": agg_ms, "algorithmic_gbs": agg_bytes / (agg_ms * 1e-3) / 1e9,
                                         "note": "csr_aggregate (mean) over the whole sampled block, 128-wide fp32 rows, rank 0's last block",
                                         "backward_as_gather_over_reversed_block_ms_256_wide": agg_bwd_ms,
                                         "max_row_degree": max_in, "max_row_degree_reversed": max_out},
            "allreduce_alone_ms": ar_ms, "allreduce_bytes": n_params * 4, "backward_ms_without_allreduce (no_sync)": nosync_bwd,
            "edges_per_step_per_gpu": edges / args.steps / world, "nodes_gathered_per_step_per_gpu": nodes / args.steps / world,
            "gather_gbs": FEAT_DIM * 4 * nodes / (stage[1] * 1e-3) / 1e9,
            "final_loss": float(loss.detach()), "gpu_launches": launches, "clocks": clock_info,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# --loader: the same hot path driven through the reference-facing Python stack, one mini-batch at a time
# ----------------------------------------------------------------------------------------------------
def run_loader(args):
    """cugraph_pyg.loader.NeighborLoader(batch_size=1024, num_neighbors=[25, 10]) over GraphStore / FeatureStore on the bench
    graph (role of the reference's sampler/sampler.py:51-165 -> loader iteration): mini-batches / s and sampled edges / s as
    a training loop sees them (every batch arrives with its features gathered), next to the device time the native calls of
    the same call groups take (the share of the wall clock the hot-path kernels keep the GPU busy).  One GPU.
    Not the driver's line: `python bench.py --loader [--workload c2]` prints its own JSON line."""
    import torch

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product path has no CPU fallback)"
    if NUM_EDGES > 400_000_000:
        # the PyG stores are fed the way a user feeds them -- an int64 COO edge_index (25.6 GB on C4) and a dense feature tensor
        # (56.8 GB) that the stores then copy -- which does not fit one GPU next to the stores' own copies on the large shapes
        sys.exit("bench.py --loader builds PyG-style stores from whole tensors: use --workload c2 (the large shapes do not fit one GPU twice)")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    os.environ.setdefault("LOCAL_WORLD_SIZE", "1")
    import pylibwholegraph.torch as wgth
    from cugraph_pyg.data import FeatureStore, GraphStore
    from cugraph_pyg.loader import NeighborLoader
    from cugraph_pyg.sampler.sampler import SampleIterator

    if os.environ.get("WGB_LOADER_PREFETCH") == "0":  # A/B: fetch features per mini-batch as the reference does
        SampleIterator.prefetch_call_group_features = False
    t0 = time.time()
    row_ptr, col = rmat_csr(torch, NUM_NODES, NUM_EDGES, 42, dev)
    dst = torch.repeat_interleave(torch.arange(NUM_NODES, device=dev), row_ptr[1:] - row_ptr[:-1])
    graph_store = GraphStore()
    graph_store[("n", "e", "n"), "coo", False, (NUM_NODES, NUM_NODES)] = torch.stack([col.long(), dst])  # PyG: [source; destination]
    del dst, col, row_ptr
    feature_store = FeatureStore(location="cuda")
    x = torch.empty((NUM_NODES, FEAT_DIM), dtype=torch.float32, device=dev)
    ar = torch.arange(FEAT_DIM, device=dev)[None, :]
    for lo in range(0, NUM_NODES, 1 << 20):
        hi = min(NUM_NODES, lo + (1 << 20))
        x[lo:hi] = ((torch.arange(lo, hi, device=dev)[:, None] + ar) & 0xFFFF).float()
    feature_store["n", "x", None] = x
    del x
    torch.cuda.empty_cache()
    log("[loader] stores built in %.1fs" % (time.time() - t0))
    labels = args.labels
    per_group = labels * BATCH
    n_groups = args.steps + args.warmup
    seeds = torch.randperm(NUM_NODES, generator=torch.Generator().manual_seed(1234))[: n_groups * per_group]

    def one_pass(nodes):
        loader = NeighborLoader((feature_store, graph_store), FANOUT, input_nodes=nodes, batch_size=BATCH, shuffle=False,
                                local_seeds_per_call=per_group)
        nb = ne = nn = 0
        acc = torch.zeros((), device=dev)
        for batch in loader:
            nb += 1
            ne += int(batch.edge_index.shape[1])
            nn += int(batch.n_id.shape[0])
            acc += batch.x[-1, -1]  # the features are there (device-side use, no host sync per batch)
        return nb, ne, nn, acc

    one_pass(seeds[: args.warmup * per_group])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    nb, ne, nn, acc = one_pass(seeds[args.warmup * per_group:])
    e1.record()
    torch.cuda.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - w0)
    dev_ms = e0.elapsed_time(e1)
    float(acc)
    out = {
        "mode": "loader",
        "metric": "loader_minibatches_per_sec (NeighborLoader over GraphStore / FeatureStore, batch 1024, fan-out [25,10], features gathered)",
        "value": nb / (wall_ms * 1e-3), "unit": "mini-batches/s",
        "edges_per_sec": ne / (wall_ms * 1e-3), "nodes_gathered_per_sec": nn / (wall_ms * 1e-3),
        "gather_gbs": nn * FEAT_DIM * 4 / (wall_ms * 1e-3) / 1e9,
        "minibatches": nb, "call_groups": args.steps, "ms_per_call_group": wall_ms / args.steps, "ms_per_minibatch": wall_ms / nb,
        "device_span_ms": dev_ms, "wall_ms": wall_ms,
        "n_gpus": 1, "higher_is_better": True, "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_size": BATCH, "minibatches_per_call_group": labels,
                   "features": "one gather per call group, row slices per mini-batch" if SampleIterator.prefetch_call_group_features else "one gather per mini-batch",
                   "stack": "cugraph_pyg.loader.NeighborLoader -> DistributedNeighborSampler -> pylibcugraph shim -> libwholegraph_b200 (fused sampler) ; "
                            "FeatureStore -> WholeMemoryEmbedding.gather"},
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a C++/OpenMP restatement of the reference's own host reference algorithms; the
# reference's implementation cannot be compiled or installed here -- SURVEY.md §8c) on the host cores.
# ----------------------------------------------------------------------------------------------------
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import wg_oracle as oracle

    # the host arm always gets every core of the box: torchrun exports OMP_NUM_THREADS=1 into each rank's environment
    oracle.set_num_threads(os.cpu_count() or 1)
    return oracle


def parity_check(torch, wgth, sampler, emb, wm_rp, wm_col, host_graph, dev, labels=4):
    """One `labels`-label call group on the bench graph through the GPU path and through the CPU oracle (checker only, outside
    every timed region): COO, edge ids, renumber map, offsets and the gathered feature rows must be identical."""
    import numpy as np

    oracle = _oracle()
    row_ptr, col = host_graph
    g = torch.Generator(device="cpu").manual_seed(99)
    seeds = torch.randperm(NUM_NODES, generator=g)[: labels * BATCH].contiguous()
    lo = (np.arange(labels + 1) * BATCH).astype(np.int64)
    res = sampler.sample(wm_rp, wm_col, seeds.to(dev), torch.from_numpy(lo).to(dev), FANOUT, SAMPLER_SEED, int64_ids=True)
    x = emb.gather(res["renumber_map"])
    exp = oracle.multihop_sample(row_ptr, col, seeds.numpy(), lo, FANOUT, SAMPLER_SEED)
    keys = ("majors", "minors", "edge_id", "renumber_map", "renumber_map_offsets", "label_hop_offsets")
    bad = [k for k in keys if not np.array_equal(res[k].cpu().numpy(), exp[k])]
    ids = exp["renumber_map"].astype(np.int64)
    want = ((ids[:, None] + np.arange(FEAT_DIM)[None, :]) & 0xFFFF).astype(np.float32)  # closed form of the table
    if not np.array_equal(x.cpu().numpy(), want):
        bad.append("gathered_features")
    assert not bad, "GPU result differs from the oracle on the bench graph: %s" % bad
    return {"ok": True, "labels": labels, "edges": int(exp["minors"].shape[0]), "nodes": int(ids.shape[0]), "compared": list(keys) + ["gathered_features"]}


PROXY_ROWS = 4_000_000  # CPU arm: rows of the host feature table the gather reads (ids mod PROXY_ROWS)


def cpu_baseline(row_ptr, col, budget_s=15.0, labels=4, steps=None):
    import numpy as np

    oracle = _oracle()
    # Feature fetch on the host: random rows of a real fp32 table with the bench's row size.  The large shapes' tables
    # (56.8 / 102 GB) are not replicated in host memory; the gather reads a 4 M-row proxy (2-4 GB, far beyond any LLC) at
    # ids mod 4 M -- the same number of random row copies per step.
    rows = min(NUM_NODES, PROXY_ROWS)
    table = np.empty((rows, FEAT_DIM), dtype=np.float32)
    for lo_ in range(0, rows, 1 << 20):
        hi_ = min(rows, lo_ + (1 << 20))
        table[lo_:hi_] = (np.arange(lo_, hi_, dtype=np.int64)[:, None] + np.arange(FEAT_DIM)[None, :]) & 0xFFFF

    def feat(ids):
        return oracle.gather(table, ids if rows == NUM_NODES else ids % rows)

    rng = np.random.default_rng(0)
    lo = (np.arange(labels + 1) * BATCH).astype(np.int64)
    edges, busy, n_steps = 0, 0.0, 0
    t0 = time.time()
    while True:
        seeds = rng.choice(NUM_NODES, size=labels * BATCH, replace=False).astype(np.int64)
        t1 = time.time()
        res = oracle.multihop_sample(row_ptr, col, seeds, lo, FANOUT, SAMPLER_SEED + n_steps)
        feat(res["renumber_map"])
        dt = time.time() - t1
        if n_steps > 0:  # first pass warms the page cache / OpenMP pool
            edges += int(res["minors"].shape[0])
            busy += dt
        n_steps += 1
        if steps is not None:
            if n_steps > steps:
                break
        elif time.time() - t0 > budget_s and n_steps >= 3:
            break
    return {"value": edges / busy, "unit": "edges/s", "cores": oracle.num_threads(), "kind": "port",
            "sample": "%d timed steps of %d labels x %d seeds (same graph, fan-out, seed rule as the GPU arm; 1 untimed warm-up step; "
                      "feature rows from a %d-row host table)" % (n_steps - 1, labels, BATCH, rows),
            "ms_per_step": 1e3 * busy / max(1, n_steps - 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    row_ptr, col = rmat_csr(torch, NUM_NODES, NUM_EDGES, 42, dev)  # input generation only; nothing timed runs on the GPU
    row_ptr, col = row_ptr.cpu().numpy(), col.cpu().numpy()
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    # same step as the GPU arm: one call group of args.labels labels x 1024 seeds; W untimed + K timed steps
    base = cpu_baseline(row_ptr, col, steps=args.steps + args.warmup - 1, labels=args.labels)
    out = {
        "impl": "reference",
        "metric": METRIC,
        "value": base["value"], "unit": "edges/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64 ids (int32 col_idx in the CSR), fp32 features", "data": "synthetic",
        "config": {"workload": WORKLOAD, "labels_per_step": args.labels, "seeds_per_label": BATCH,
                   "host_threads": base["cores"], "note": "CPU arm runs on rank 0's host cores only, whatever N is"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--labels", type=int, default=LABELS_PER_STEP)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--loader", action="store_true",
                    help="time the same hot path through cugraph_pyg.loader.NeighborLoader + FeatureStore, per mini-batch (one GPU; c2 recommended)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the oracle comparison on the bench graph (profiling runs)")
    ap.add_argument("--hot-ratio", type=float, default=0.1,
                    help="N > 1: fraction of the feature rows (the highest-degree vertices) replicated on every GPU; 0 = none")
    ap.add_argument("--gather-stream", type=int, default=-1, help="run the feature gather on a second stream (1, default) or in line (0)")
    ap.add_argument("--gather-sms", type=int, default=-1,
                    help="SM budget of the feature gather (reference knob `gather_sms`; grid = 8 CTAs x this many SMs, spread over all SMs); "
                         "-1 = every SM (default).  Experiment: 74 leaves half of every SM to the sampling kernels of the next call group")
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["c5"],
                    help="c4 (default, the shape the metric is quoted on: papers100M) | c2 (|V|=10M, |E|=160M) | headline (|V|=100M, |E|=1B, F=256) | tiny (harness self-test)")
    args = ap.parse_args()
    if args.workload == "c5":
        if args.labels == LABELS_PER_STEP:
            args.labels = 32
        args.warmup = max(args.warmup, 3)
        return run_c5(args)
    if WORKLOADS[args.workload] is not None:
        globals().update(WORKLOADS[args.workload])
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.loader:
        run_loader(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
