#!/usr/bin/env python
"""bench.py -- edge updates/sec of the sampled-SGD hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm: CUDA hot path through the C ABI
  python bench.py --impl reference --steps K --warmup W    # reference arm: the reference's own CPU code

N = 1 (config.workload): BASELINE.json configs[1] -- LINE 2nd order, dim 128, K = 5 negatives, synthetic power-law graph
with 1M vertices / 10M weighted edges (undirected -> 20M CSR entries), C++-tree semantics, fp32 tables. A "step" is one
Hogwild pass of `--batch` edge updates (default 2^24). Tables (2 x 512 MB) plus graph (~0.5 GB) exceed the 126 MB L2, so
no L2 flush is needed between steps.

N > 1 (default --parallelism rotating): BASELINE.json configs[4] shape -- 12.5M vertices and 250M weighted edge lines PER
GPU (N = 8: 100M vertices / 2B edges), generated on the device (smore_graph_create_synthetic_rotating), both tables
row-sharded, the vertex shards travelling around the ring of GPUs in 2N sub-parts ("rotating shards"): every
update runs out of local HBM with the single-GPU kernel while the sub-part trained in the previous episode moves to the
next GPU over NVLink (copy engines). A "step" is one full cycle of 2N episodes of `--episode-batch` updates per GPU.

value : updates/s with graph + tables resident in HBM (timed: K steps, CUDA events on the launching stream, barrier +
        synchronize on both sides, max over ranks).
e2e   : the same metric through the C ABI the way a host Train() call sees it: every step uploads both embedding tables
        (this rank's shards) from pinned HOST memory, trains, and reads the vertex table back to the host. N = 1: two
        model instances alternate, so the uploads of step i+1 and the read-back of step i-1 run on the copy engines while
        step i trains (an input pipeline: every step still moves all of its bytes). N > 1: the read-back of step i
        overlaps the context-table upload of step i+1.
roofline : algorithmic bytes per update (2*(K+2)*D*4 + 76 = 7244 B; with --split-samples one more row:
        2*(K+3)*D*4 + 84 = 8276 B) x updates / kernel time (the library's own CUDA events around the kernel) against
        MEASURED_PEAKS.json hbm_gbs.
cpu_baseline : the reference's CPU implementation timed on this box's host cores on a bounded sample (N = 1 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIM, K, V_TARGET, N_EDGES, GRAPH_SEED = 128, 5, 1_000_000, 10_000_000, 20261018
ALGO_BYTES = 2 * (K + 2) * DIM * 4 + 76
ALGO_BYTES_SPLIT = 2 * (K + 3) * DIM * 4 + 84  # split samples (sharded modes): second vertex row + its alias draw
V_PER_GPU, E_PER_GPU = 12_500_000, 250_000_000  # configs[4] / 8
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def physical_cores():
    """(physical cores, logical cpus) of this host."""
    logical = os.cpu_count() or 1
    try:
        cores = set()
        phys = core = None
        for line in open("/proc/cpuinfo"):
            if line.startswith("physical id"):
                phys = line.split(":")[1].strip()
            elif line.startswith("core id"):
                core = line.split(":")[1].strip()
            elif not line.strip():
                if phys is not None and core is not None:
                    cores.add((phys, core))
                phys = core = None
        if cores:
            return len(cores), logical
    except Exception:
        pass
    return logical, logical


def graph_shape(args, world):
    """(V, E_lines) the bench graph is generated with (V is the label range; configs[1] ends up with slightly fewer used)."""
    if world == 1 or args.parallelism != "rotating":
        return int(V_TARGET * args.scale), int(N_EDGES * args.scale)
    return int(V_PER_GPU * args.scale) * world, int(E_PER_GPU * args.scale) * world


def bench_config(args, world, V=None, E=None):
    """The `config` object of the JSON line: identical for both arms (--impl ours / reference) at the same flags."""
    lv, le = graph_shape(args, world)
    if world == 1:
        wl = (f"LINE-2 dim={DIM} K={K} Hogwild, synthetic power-law graph V={V if V else lv} E_lines={le} "
              f"(undirected, {2 * le} CSR entries) [BASELINE configs[1]]")
        par = "single GPU"
        upd = args.batch
        l2 = "working set per GPU (1.0 GB of table rows + >= 0.5 GB graph) exceeds the 126 MB L2; no flush between steps"
    elif args.parallelism == "rotating":
        wl = (f"LINE-2 dim={DIM} K={K} Hogwild, synthetic power-law graph V={lv} E_lines={le} (undirected), generated on the "
              f"device: {lv // world} vertices / {le // world} edge lines per GPU [BASELINE configs[4] shape: 100M vertices / 2B edges at 8 GPUs]")
        par = (f"tables row-sharded over {world} GPUs; context shards fixed, vertex shards rotating around the ring in {2 * world} "
               f"sub-parts (block-cyclic episodes, every update out of local HBM, sub-parts moved by copy engines over NVLink "
               f"behind the update kernel); shard-local negatives"
               + (" applied to an independently drawn vertex (split samples)" if args.split_samples else "")
               + "; host barrier per episode, no data-path collective")
        upd = args.episode_batch * 2 * world
        l2 = f"working set per GPU ({lv // world * DIM * 4 * 2.5 / 1e9:.1f} GB of table rows + block tables) exceeds the 126 MB L2; no flush between steps"
    else:
        wl = (f"LINE-2 dim={DIM} K={K} Hogwild, synthetic power-law graph V={V if V else lv} E_lines={le} "
              f"(undirected, {2 * le} CSR entries) [BASELINE configs[1] graph at every N]")
        par = {"sharded": f"tables row-sharded over {world} GPUs, context-owner computes, remote vertex rows over NVLink peer "
                          f"mappings (CUDA IPC), no data-path collective",
               "sharded-replica": f"row-sharded over {world} GPUs, vertex rows read from a local replica refreshed every step, "
                                  f"deltas pushed with red.global.add",
               "sharded-exchange": f"row-sharded over {world} GPUs, remote vertex rows moved in super-batches of {args.superbatch} "
                                   f"samples/GPU by NCCL all-to-all",
               "replicas": f"{world} independent replicas (weak scaling)"}[args.parallelism]
        upd = args.batch
        l2 = "working set per GPU exceeds the 126 MB L2; no flush between steps"
    return {"workload": wl, "updates_per_step_per_gpu": upd, "parallelism": par, "l2": l2}


def make_graph(scale=1.0):
    from smore_b200 import synth

    t0 = time.time()
    nv, ne = int(V_TARGET * scale), int(N_EDGES * scale)
    src, dst, w = synth.power_law_edges(nv, ne, GRAPH_SEED)
    off, col, ww, labels = synth.csr_from_edges(src, dst, w, undirected=True)
    log(f"[bench] graph: V={len(off) - 1} E={len(col)} generated in {time.time() - t0:.1f}s")
    return (src, dst, w), (off, col, ww)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
        out["reasons"] = sorted(reasons)
        return out


class CpuReference:
    """Reference CPU path on this box's host cores: the compiled reference (kind "reference": unmodified sources, its own
    flags -Ofast -fopenmp and its own RNG) when oracle/_ref travelled here, else the oracle restatement (kind "port").
    The graph is ingested once; timed() then runs bounded samples of LINE::Train. Threads = all logical CPUs the box
    offers (the reference's -threads flag); the physical core count is stated next to it."""

    def __init__(self, edges, csr, threads=None):
        from oracle import bindings as B

        self.B = B
        self.phys, self.logical = physical_cores()
        self.threads = threads or self.logical
        self.kind = "reference" if B.ref_available(fast=True) else "port"
        if self.kind == "reference":
            src, dst, w = edges
            tmp = tempfile.NamedTemporaryFile(suffix=".txt", delete=False)
            tmp.close()
            t0 = time.time()
            B.write_edge_list_fast(tmp.name, src, dst, w)
            self.ref = B.Ref(B.K_LINE, tmp.name, True, DIM, order=2, fast=True)
            os.unlink(tmp.name)
            log(f"[bench] reference ingested the graph in {time.time() - t0:.1f}s (V={self.ref.V})")
            t0 = time.time()
            self.ref.train(1, K, alpha=0.025, workers=self.threads)  # LINE::Train granularity: sample_times x 1e6 updates
            self.sec_per_million = time.time() - t0
        else:
            off, col, ww = csr
            self.g = B.OracleGraph(B.SEM_CPP, off, col, ww)
            rng = np.random.RandomState(0)
            self.Wv = (rng.random_sample((self.g.V, DIM)) - 0.5) / DIM
            self.Wc = np.zeros((self.g.V, DIM))
            t = self.g.time_line_cpp(self.Wv, self.Wc, K, 0.025, 1_000_000, 1, self.threads)
            self.sec_per_million = t

    def timed(self, seconds_target):
        s = int(max(1, min(256, seconds_target / max(self.sec_per_million, 1e-3))))
        updates = s * 1_000_000
        where = f"{self.threads} OpenMP threads on {self.phys} physical cores ({self.logical} logical CPUs)"
        if self.kind == "reference":
            t0 = time.time()
            self.ref.train(s, K, alpha=0.025, workers=self.threads)
            dt = time.time() - t0
            what = (f"LINE::Train(sample_times={s}) = {updates} updates on the configs[1] graph (V=1M / 10M edge lines, fp64 tables), "
                    f"{where}, {dt:.1f}s (compiled from the unmodified reference sources with its own flags -Ofast -fopenmp)")
        else:
            dt = self.g.time_line_cpp(self.Wv, self.Wc, K, 0.025, updates, 1, self.threads)
            what = (f"{updates} updates of the oracle restatement (oracle/smore_oracle.cpp, -O2 -fopenmp) on the configs[1] graph, "
                    f"{where}, {dt:.1f}s")
        return {"value": updates / dt, "unit": "updates/s", "cores": self.phys, "threads": self.threads, "kind": self.kind,
                "sample": what}


def cpu_baseline(edges, csr, seconds_target=12.0):
    return CpuReference(edges, csr).timed(seconds_target)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    # the CPU reference always trains the configs[1] graph: it cannot load the N > 1 workload (30 M-vertex hash limit,
    # src/proNet.h:34); its per-update cost on the largest graph it loads is what the line reports (`cpu_baseline.sample`)
    edges, csr = make_graph(args.scale if world == 1 else 1.0)
    cpu = CpuReference(edges, csr)
    per_step = max(3.0, min(15.0, 100.0 / max(1, args.steps + args.warmup)))
    vals, secs = [], []
    base = None
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        base = cpu.timed(per_step)
        if i >= args.warmup:
            vals.append(base["value"])
            secs.append(time.time() - t0)
    v = float(np.mean(vals)) if vals else base["value"]
    base["value"] = v
    V = len(csr[0]) - 1
    print(json.dumps({
        "impl": "reference", "metric": "edge_updates_per_sec", "value": v, "unit": "updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)) if secs else None,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": bench_config(args, world, V=V),
        "cpu_baseline": base, "e2e": {"value": v, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from smore_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: smore_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    # stdout must carry exactly one JSON line: NCCL (torch's communicator and the library's own) prints its version banner
    # with printf, so fd 1 points at stderr until the line is written
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    capi.check(capi.lib().smore_init(local))

    mode = "single" if world == 1 else args.parallelism
    rotating = mode == "rotating"
    sharded = mode in ("sharded", "sharded-replica", "sharded-exchange")
    exchange, replica = mode == "sharded-exchange", mode == "sharded-replica"
    edges = csr = None
    t0 = time.time()
    if rotating:
        lv, le = graph_shape(args, world)
        g = capi.Graph.synthetic_rotating(lv, le, GRAPH_SEED, rank, world, semantics=capi.SEM_CPP)
        ri = g.rotation_info()
        log(f"[bench] rank {rank}: synthetic graph V={lv} E_lines={le} -> {int(ri['block_edges'].sum())} entries in {ri['n_sub']} blocks, "
            f"mass {ri['block_mass'].sum():.4f}, built on the device in {time.time() - t0:.1f}s")
    else:
        edges, csr = make_graph(args.scale)
        off, col, ww = csr
        t0 = time.time()
        g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP)
        if sharded:
            info = g.set_shard(rank, world)
            log(f"[bench] rank {rank}: shard {info}")
        log(f"[bench] rank {rank}: alias tables built + uploaded in {time.time() - t0:.1f}s")
    m = capi.Model(g, DIM, 2, capi.F32)
    m.init(0, True, seed=1)
    m.init(1, False, seed=1)
    from smore_b200 import dist as sdist

    if rotating:
        m.enable_rotation()
        sdist.connect_rotation(m)
        dist.barrier()
    elif sharded:
        if exchange:
            sdist.init_exchange()
            m.enable_exchange(args.superbatch, args.hot_threshold)
            if args.hot_threshold >= 0:
                sdist.connect_peers(m)
        else:
            sdist.connect_peers(m)
        if replica:
            m.enable_replica(0)
        dist.barrier()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    p = capi.default_params()
    p.semantics, p.mode, p.seed, p.alpha = capi.SEM_CPP, capi.MODE_HOGWILD, 1, 0.025
    p.negative_samples, p.order = K, 2
    p.neg_mode = capi.PAIRING_SPLIT if args.split_samples else capi.PAIRING_COUPLED
    step_no = [0]
    episodes_per_step = 2 * world

    def step():
        """-> (samples, kernel ms) of this rank"""
        i = step_no[0]
        step_no[0] += 1
        if rotating:
            # one cycle: every vertex sub-part meets every context shard once; total = samples of all ranks per episode
            p.total = args.episode_batch * world
            p.stream_base = 0  # (train_line_rotating derives the sub-streams from the absolute episode number and the rank)
            done, ms = sdist.train_line_rotating([m], p, episodes_per_step, first_episode=i * episodes_per_step,
                                                 barrier=dist.barrier, world=world, rank0=rank)
            return done[0], ms[0]
        # sharded: `total` is the GLOBAL update count of the step; each rank runs its mass share (~ batch per GPU)
        p.total = args.batch * (world if sharded else 1)
        p.stream_base = (i * world + rank) * (1 << 20)  # fresh Philox sub-streams every step / rank
        if sharded:
            dist.barrier()  # ranks enter the step together (control plane only)
        if replica:
            m.refresh_replica(0)
            dist.barrier()
        st = m.train_line(p)
        return st["samples"], st["kernel_ms"]

    # ---- value: inputs resident in HBM ----
    for _ in range(args.warmup):
        step()
    launches0 = capi.kernel_launches()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    # nvidia-smi needs ~0.2 s to deliver its first sample: keep the device under the same load meanwhile (untimed steps, the
    # same number on every rank), so that short timed regions are still covered by samples taken under load
    for _ in range(16 if world == 1 else 2):
        step()
    barrier()
    launches0 = capi.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    updates = 0
    kernel_ms = 0.0
    for _ in range(args.steps):
        n, kms = step()
        updates += n
        kernel_ms += kms
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    launches = capi.kernel_launches() - launches0

    # ---- e2e: host buffers in, host buffers out, every step ----
    V = m.rows  # local rows (== g.V unless sharded)
    e2e_updates, e2e_ms = 0, float("nan")
    e2e_skip = None
    if not args.no_e2e:
        try:
            import psutil

            need = (4 if world == 1 else 3) * V * DIM * 4 * world  # pinned buffers per rank, all ranks on this host
            avail = psutil.virtual_memory().available
            if need > 0.6 * avail:
                e2e_skip = f"skipped: {need / 1e9:.0f} GB of pinned host buffers needed, {avail / 1e9:.0f} GB of host RAM available"
        except Exception:
            pass
    if e2e_skip:
        args.no_e2e = True
        log(f"[bench] e2e leg {e2e_skip}")
    if not args.no_e2e:
        hv = torch.empty((V, DIM), dtype=torch.float32, pin_memory=True)
        hc = torch.empty((V, DIM), dtype=torch.float32, pin_memory=True)
        m.get_rows(0, out=hv.numpy())
        m.get_rows(1, out=hc.numpy())
        if world == 1:
            # Two model instances on the one graph, used alternately: the uploads of step i + 1 (into the idle instance) and
            # the read-back of step i - 1 run on the copy engines WHILE step i trains -- an input pipeline, nothing more:
            # every step still uploads both of its tables from pinned host memory and reads its result back. The library
            # orders an instance's uploads after its pending read-backs on the device, so the host never blocks on a copy
            # before launching the next train call.
            m2 = capi.Model(g, DIM, 2, capi.F32)
            inst = [m, m2]
            houts = [torch.empty((V, DIM), dtype=torch.float32, pin_memory=True) for _ in inst]
            e2e_no = [0]

            def upload(mm):
                mm.set_rows_async(1, hc.numpy())
                mm.set_rows_async(0, hv.numpy())

            def train_on(mm):
                i = step_no[0]
                step_no[0] += 1
                p.total = args.batch
                p.stream_base = (i * world + rank) * (1 << 20)
                return mm.train_line(p)["samples"]

            def e2e_step():
                i = e2e_no[0]
                e2e_no[0] += 1
                cur, nxt = inst[i & 1], inst[(i + 1) & 1]
                upload(nxt)                                      # step i + 1's inputs: in flight while step i trains
                cur.wait_copies(uploads=True, readbacks=False)   # step i's inputs (enqueued one step ago)
                n = train_on(cur)
                cur.get_rows_async(0, houts[i & 1].numpy())      # step i's result: in flight while step i + 1 trains
                return n

            upload(inst[0])

            def wait_all():
                for mm in inst:
                    mm.wait_copies()
        else:
            hout = torch.empty((V, DIM), dtype=torch.float32, pin_memory=True)

            def e2e_step():
                # pipeline: the context-table upload overlaps the previous step's read-back of the vertex table (PCIe is full
                # duplex); the vertex-table upload has to wait for that read-back (same device rows)
                m.set_rows_async(1, hc.numpy())
                m.wait_copies(uploads=False, readbacks=True)
                m.set_rows_async(0, hv.numpy())
                m.wait_copies(uploads=True, readbacks=False)
                n, _ = step()
                m.get_rows_async(0, hout.numpy())
                return n

            def wait_all():
                m.wait_copies()

        e2e_step()
        e2e_step()
        wait_all()
        barrier()  # (world == 1: the next step's inputs were enqueued one step ago, as for every step of the timed region)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            e2e_updates += e2e_step()
        wait_all()
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)

    stats = torch.tensor([ms, 0.0 if args.no_e2e else e2e_ms, float(updates), float(e2e_updates), kernel_ms, float(launches)],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms, kernel_ms = mx[0].item(), mx[1].item(), mx[4].item()
        updates, e2e_updates, launches = sm[2].item(), sm[3].item(), sm[5].item()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
    # dominant kernel = k_line: per-rank algorithmic bytes / its own CUDA-event time (max over ranks)
    split = args.split_samples and (rotating or mode == "sharded")
    algo = ALGO_BYTES_SPLIT if split else ALGO_BYTES
    per_rank_updates = updates / world
    achieved = per_rank_updates * algo / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_k_line.json")
    if world == 1 and os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    out = {
        "metric": "edge_updates_per_sec", "value": updates / (ms * 1e-3), "unit": "updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, world, V=g.V),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_line<float,4,1>" + (" (split samples)" if split else ""),
                     "algorithmic_bytes_per_update": algo, "peak_source": peak_src},
        "e2e": {"value": None if args.no_e2e else e2e_updates / (e2e_ms * 1e-3), "unit": "updates/s",
                "h2d_bytes_per_step": 2 * V * DIM * 4 * world, "d2h_bytes_per_step": V * DIM * 4 * world,
                "note": e2e_skip or ("per step and GPU: both table shards uploaded from pinned host memory, Train call(s), vertex "
                                     "shard read back; " +
                                     ("two model instances alternate: the uploads of step i+1 and the read-back of step i-1 run "
                                      "on the copy engines while step i trains" if world == 1 else
                                      "the read-back overlaps the next step's context-table upload"))},
        "gpu_launches": int(launches), "clocks": clk,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            out["cpu_baseline"] = cpu_baseline(edges, csr)
        except Exception as ex:  # the baseline is reporting only; never lose the GPU line
            out["cpu_baseline"] = {"value": None, "unit": "updates/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 24, help="edge updates per step per GPU (N = 1 and the peer-access modes)")
    ap.add_argument("--episode-batch", type=int, default=1 << 23, help="rotating: edge updates per episode per GPU (a step = 2N episodes)")
    ap.add_argument("--scale", type=float, default=1.0, help="graph size multiplier (1.0 = configs[1] / configs[4] per-GPU share)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="experiments only: skip the host-buffer (e2e) leg")
    ap.add_argument("--parallelism", default="rotating",
                    choices=["rotating", "sharded", "sharded-replica", "sharded-exchange", "replicas"], help="N>1 only")
    ap.add_argument("--split-samples", action="store_true",
                    help="sharded modes: apply the negatives to an independently drawn vertex (one more row per update)")
    ap.add_argument("--superbatch", type=int, default=1 << 20, help="sharded-exchange: samples per GPU and super-batch")
    ap.add_argument("--hot-threshold", type=float, default=64.0,
                    help="sharded-exchange: expected source draws per super-batch above which a vertex keeps a single copy (<0: none)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
