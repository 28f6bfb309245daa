#!/usr/bin/env python
"""bench.py -- edge updates/sec of the sampled-SGD hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm: CUDA hot path through the C ABI
  python bench.py --impl reference --steps K --warmup W    # reference arm: the reference's own CPU code

Workload (config.workload): BASELINE.json configs[1] -- LINE 2nd order, dim 128, K = 5 negatives, synthetic power-law
graph with 1M vertices / 10M weighted edges (undirected -> 20M CSR entries), C++-tree semantics, fp32 tables.
A "step" is one Hogwild pass of `--batch` edge updates (default 2^24) over that graph. Tables (2 x 512 MB) plus graph
(~0.5 GB) exceed the 126 MB L2, so no L2 flush is needed between steps.

value : updates/s with graph + tables resident in HBM (timed: K train calls, CUDA events on the launching stream,
        barrier + synchronize on both sides, max over ranks).
e2e   : the same metric through the C ABI the way a host Train() call sees it: every step uploads both embedding
        tables from pinned HOST memory, trains, and reads the vertex table back to the host.
roofline : algorithmic bytes per update (2*(K+2)*D*4 + 76 = 7244 B) x updates / kernel time (the
        library's own CUDA events around the kernel) against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline : the reference's CPU implementation timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIM, K, V_TARGET, N_EDGES, GRAPH_SEED = 128, 5, 1_000_000, 10_000_000, 20261018
ALGO_BYTES = 2 * (K + 2) * DIM * 4 + 76
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_name(V, E):
    return (f"LINE-2 dim={DIM} K={K} Hogwild fp32, synthetic power-law graph V={V} E_lines={E // 2} "
            f"(undirected, {E} CSR entries) [BASELINE configs[1]]")


def make_graph(scale=1.0):
    from smore_b200 import synth

    t0 = time.time()
    nv, ne = int(V_TARGET * scale), int(N_EDGES * scale)
    src, dst, w = synth.power_law_edges(nv, ne, GRAPH_SEED)
    off, col, ww, labels = synth.csr_from_edges(src, dst, w, undirected=True)
    log(f"[bench] graph: V={len(off) - 1} E={len(col)} generated in {time.time() - t0:.1f}s")
    return (src, dst, w), (off, col, ww)


def make_graph_torch(scale, device):
    """Same generator as smore_b200.synth.power_law_edges + csr_from_edges(undirected), on the GPU (multi-GPU runs scale
    the graph with the number of GPUs; 80M edges are generated in about a second instead of minutes of numpy). Vertex ids
    are the labels themselves (no first-appearance relabelling: nothing is compared with the reference at N > 1)."""
    import torch

    t0 = time.time()
    nv, ne = int(V_TARGET * scale), int(N_EDGES * scale)
    gen = torch.Generator(device=device)
    gen.manual_seed(GRAPH_SEED)
    p = torch.arange(1, nv + 1, device=device, dtype=torch.float64) ** (-1.0 / 1.5)
    cdf = torch.cumsum(p, 0)
    cdf /= cdf[-1].clone()
    perm = torch.randperm(nv, device=device, generator=gen)

    def endpoints():
        u = torch.rand(ne, device=device, dtype=torch.float64, generator=gen)
        return perm[torch.searchsorted(cdf, u, right=True).clamp_(max=nv - 1)]

    src, dst = endpoints(), endpoints()
    w = torch.randint(1, 6, (ne,), device=device, generator=gen).to(torch.float64)
    es = torch.stack([src, dst], 1).reshape(-1)
    ed = torch.stack([dst, src], 1).reshape(-1)
    ew = torch.stack([w, w], 1).reshape(-1)
    order = torch.argsort(es, stable=True)
    col = ed[order].to(torch.int32).cpu().numpy()
    ww = ew[order].cpu().numpy()
    off = torch.zeros(nv + 1, dtype=torch.int64, device=device)
    off[1:] = torch.cumsum(torch.bincount(es, minlength=nv), 0)
    off = off.cpu().numpy()
    del src, dst, w, es, ed, ew, order, perm, cdf, p
    torch.cuda.empty_cache()
    log(f"[bench] graph (torch/GPU): V={nv} E={len(col)} generated in {time.time() - t0:.1f}s")
    return None, (off, col, ww)


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
    The graph is ingested once; timed() then runs bounded samples of LINE::Train."""

    def __init__(self, edges, csr, cores=None):
        from oracle import bindings as B

        self.B = B
        self.cores = cores or os.cpu_count() or 1
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
            self.ref.train(1, K, alpha=0.025, workers=self.cores)  # LINE::Train granularity: sample_times x 1e6 updates
            self.sec_per_million = time.time() - t0
        else:
            off, col, ww = csr
            self.g = B.OracleGraph(B.SEM_CPP, off, col, ww)
            rng = np.random.RandomState(0)
            self.Wv = (rng.random_sample((self.g.V, DIM)) - 0.5) / DIM
            self.Wc = np.zeros((self.g.V, DIM))
            t = self.g.time_line_cpp(self.Wv, self.Wc, K, 0.025, 1_000_000, 1, self.cores)
            self.sec_per_million = t

    def timed(self, seconds_target):
        s = int(max(1, min(256, seconds_target / max(self.sec_per_million, 1e-3))))
        updates = s * 1_000_000
        if self.kind == "reference":
            t0 = time.time()
            self.ref.train(s, K, alpha=0.025, workers=self.cores)
            dt = time.time() - t0
            what = (f"LINE::Train(sample_times={s}) = {updates} updates on the same graph, {self.cores} OpenMP threads, "
                    f"{dt:.1f}s (compiled from the unmodified reference sources with its own flags -Ofast -fopenmp)")
        else:
            dt = self.g.time_line_cpp(self.Wv, self.Wc, K, 0.025, updates, 1, self.cores)
            what = (f"{updates} updates of the oracle restatement (oracle/smore_oracle.cpp, -O2 -fopenmp), "
                    f"{self.cores} threads, {dt:.1f}s")
        return {"value": updates / dt, "unit": "updates/s", "cores": self.cores, "kind": self.kind, "sample": what}


def cpu_baseline(edges, csr, seconds_target=12.0):
    return CpuReference(edges, csr).timed(seconds_target)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    edges, csr = make_graph(args.scale)
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
    V, E = len(csr[0]) - 1, len(csr[1])
    print(json.dumps({
        "impl": "reference", "metric": "edge_updates_per_sec", "value": v, "unit": "updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)) if secs else None,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": workload_name(V, E)},
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
        # NCCL writes its version / debug lines to stdout by default; stdout must carry exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    capi.check(capi.lib().smore_init(local))

    sharded = world > 1 and args.parallelism in ("sharded", "sharded-replica", "sharded-exchange")
    exchange = sharded and args.parallelism == "sharded-exchange"
    # N > 1, sharded: by default every N runs the SAME configs[1] graph (updates per GPU fixed: weak scaling in work).
    # --grow-graph grows the graph with N instead (1M vertices / 10M edges per GPU, generated on the GPU). Measured
    # (DESIGN.md section 7): growing helps at 2 GPUs (hub rows are hit half as often per second) but the randomly accessed
    # PEER footprint then grows too, and fine-grained peer access falls off a cliff beyond ~1 GB of peer rows.
    grow = sharded and args.grow_graph
    if grow:
        edges, csr = make_graph_torch(args.scale * world, torch.device("cuda", local))
    else:
        edges, csr = make_graph(args.scale)
    off, col, ww = csr
    t0 = time.time()
    g = capi.Graph.from_csr(off, col, ww, semantics=capi.SEM_CPP)
    replica = sharded and args.parallelism == "sharded-replica"
    if sharded:
        # row-sharded store: this rank owns vertices v % world == rank, computes the samples whose positive
        # context it owns, and reaches remote vertex rows over NVLink peer mappings. No data-path collective.
        info = g.set_shard(rank, world)
        log(f"[bench] rank {rank}: shard {info}")
    log(f"[bench] rank {rank}: alias tables built + uploaded in {time.time() - t0:.1f}s")
    m = capi.Model(g, DIM, 2, capi.F32)
    m.init(0, True, seed=1)
    m.init(1, False, seed=1)
    if sharded:
        from smore_b200 import dist as sdist

        if exchange:
            # bulk-exchange mode: no peer mappings; the library's own NCCL communicator carries the row batches
            sdist.init_exchange()
            m.enable_exchange(args.superbatch, args.hot_threshold)
            if args.hot_threshold >= 0:
                sdist.connect_peers(m)  # hot (hub) rows stay single-copy behind the peer mappings
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
    # sharded: `total` is the GLOBAL update count of the step; each rank runs its mass share (~ batch per GPU)
    p.negative_samples, p.order, p.total = K, 2, args.batch * (world if sharded else 1)
    step_no = [0]

    def step():
        # fresh Philox sub-streams every step (stream ids never repeat across steps / ranks)
        p.stream_base = (step_no[0] * world + rank) * (1 << 20)
        step_no[0] += 1
        if sharded:
            dist.barrier()  # ranks enter the step together (control plane only)
        if replica:
            m.refresh_replica(0)  # pull the authoritative vertex rows (inside the timed region), then everyone trains
            dist.barrier()
        return m.train_line(p)

    # ---- value: inputs resident in HBM ----
    for _ in range(args.warmup):
        step()
    launches0 = capi.kernel_launches()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    updates = 0
    kernel_ms = 0.0
    for _ in range(args.steps):
        st = step()
        updates += st["samples"]
        kernel_ms += st["kernel_ms"]
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    launches = capi.kernel_launches() - launches0

    # ---- e2e: host buffers in, host buffers out, every step ----
    V = m.rows  # local rows (== g.V unless sharded)
    if not args.no_e2e:
        hv = torch.empty((V, DIM), dtype=torch.float32).pin_memory()
        hc = torch.empty((V, DIM), dtype=torch.float32).pin_memory()
        m.get_rows(0, out=hv.numpy())
        m.get_rows(1, out=hc.numpy())
        hout = torch.empty((V, DIM), dtype=torch.float32).pin_memory()

    def e2e_step():
        m.set_rows(0, hv.numpy())
        m.set_rows(1, hc.numpy())
        st = step()
        m.get_rows(0, out=hout.numpy())
        return st

    e2e_updates, e2e_ms = 0, float("nan")
    if not args.no_e2e:
        e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            e2e_updates += e2e_step()["samples"]
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)

    stats = torch.tensor([ms, 0.0 if args.no_e2e else e2e_ms, float(updates), float(e2e_updates), kernel_ms, float(launches)], dtype=torch.float64,
                         device="cuda")
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
    # dominant kernel = k_line: per-rank algorithmic bytes / its own CUDA-event time
    per_rank_updates = updates / world
    achieved = per_rank_updates * ALGO_BYTES / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_k_line.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    out = {
        "metric": "edge_updates_per_sec", "value": updates / (ms * 1e-3), "unit": "updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(g.V, g.E) + (
                       f" -- graph grown with the GPU count ({V_TARGET} vertices / {N_EDGES} edge lines per GPU)" if grow else ""),
                   "updates_per_step_per_gpu": args.batch,
                   "parallelism": "single GPU" if world == 1 else (
                       f"tables row-sharded over {world} GPUs, context-owner computes, remote vertex rows over NVLink "
                       f"peer mappings (CUDA IPC), no data-path collective"
                       + (" [replica mode: vertex rows read from a local replica refreshed every step, deltas pushed with "
                          "red.global.add]" if replica else "") if sharded and not exchange
                       else f"tables row-sharded over {world} GPUs, context-owner computes, remote vertex rows moved in "
                            f"super-batches of {args.superbatch} samples/GPU by NCCL all-to-all (request lists, rows out, "
                            f"rows back), owner applies the deltas; vertices expected >= {args.hot_threshold} times per super-batch "
                            f"stay single-copy behind NVLink peer mappings ({m.exchange_stats()['hot_vertices']} of {g.V})" if exchange
                       else f"{world} independent replicas (weak scaling)"),
                   "l2": "working set per GPU (1.0 GB of table rows + >= 0.5 GB graph) exceeds the 126 MB L2; no flush between steps"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_line<float,4,1>", "algorithmic_bytes_per_update": ALGO_BYTES,
                     "peak_source": peak_src},
        "e2e": {"value": None if args.no_e2e else e2e_updates / (e2e_ms * 1e-3), "unit": "updates/s",
                "h2d_bytes_per_step": 2 * V * DIM * 4, "d2h_bytes_per_step": V * DIM * 4,
                "note": "per step: both tables uploaded from pinned host memory, Train call, vertex table read back"},
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
    ap.add_argument("--batch", type=int, default=1 << 24, help="edge updates per step per GPU")
    ap.add_argument("--scale", type=float, default=1.0, help="graph size multiplier (1.0 = configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="experiments only: skip the host-buffer (e2e) leg")
    ap.add_argument("--grow-graph", action="store_true", help="N>1: grow the graph with N (1M vertices per GPU)")
    ap.add_argument("--parallelism", default="sharded", choices=["sharded", "sharded-replica", "sharded-exchange", "replicas"],
                    help="N>1 only")
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
