#!/usr/bin/env python
"""bench.py — LightGCN propagation throughput (edges/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C5|C5/8|C4|C1] [--impl reference]

One "step" = one full L-layer LightGCN propagation (the reference's LightGCN.forward,
lightgcn.py:62-104) over the whole synthetic graph; value = L * nnz(Â) / t, nnz = 2E directed
edge traversals per layer, whole job over all N GPUs.  Default workload = BASELINE.json
configs[4] ("C5": 20M users x 5M items, 500M edges, d=128, L=4); it fits one 180 GB GPU, so N=1
runs the full graph and N>1 is STRONG scaling of the same graph, row-partitioned with one
all-gather of the layer's rows per layer.  The printed JSON line is described in DESIGN.md §6.
"""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (n_users, n_items, n_edges, d, L)
    "C5": (20_000_000, 5_000_000, 500_000_000, 128, 4),
    "C5/8": (2_500_000, 625_000, 62_500_000, 128, 4),
    "C5/64": (312_500, 78_125, 7_812_500, 128, 4),
    "C4": (52_643, 91_599, 2_984_108, 64, 3),
    "C1": (6_040, 3_706, 1_000_209, 64, 3),
}
CPU_SAMPLE = "C5/64"


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def alg_bytes_per_layer(nnz, n_rows, d):
    """per directed edge 4 (col) + 4 (val) + 4d (gathered row); per output row
    4 (indptr) + 4d (write)."""
    return nnz * (8 + 4 * d) + n_rows * (4 + 4 * d)


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # NVML missing: report nulls rather than fail the bench
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's LightGCN.forward on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_forward_setup(workload, seed=42):
    from gnn_recommendations_b200.synthetic import synth_pairs_host
    from oracle import pyoracle as po

    nu, ni, e, d, L = WORKLOADS[workload]
    u, i = synth_pairs_host(nu, ni, e, seed)
    adj = po.build_norm_adj(u, i, nu, ni)
    coo = po.to_torch_coo(adj)
    g = torch.Generator().manual_seed(seed)
    uw = torch.randn(nu, d, generator=g) * 0.1
    iw = torch.randn(ni, d, generator=g) * 0.1
    nnz = int(adj["vals"].size)
    return (coo, uw, iw, L), L * nnz, po


def cpu_forward_time(setup, po, steps, warmup):
    coo, uw, iw, L = setup
    with torch.no_grad():
        for _ in range(warmup):
            po.lightgcn_forward(coo, uw, iw, L)
        t0 = time.perf_counter()
        for _ in range(steps):
            po.lightgcn_forward(coo, uw, iw, L)
        return (time.perf_counter() - t0) / steps


def run_reference(args):
    """--impl reference: the reference's CPU LightGCN.forward (oracle port: same torch ops, the
    reference itself does not travel to the GPU box) on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nu, ni, e, d, L = WORKLOADS[args.workload]
    sample = CPU_SAMPLE if e > WORKLOADS[CPU_SAMPLE][2] else args.workload
    setup, edges, po = cpu_forward_setup(sample)
    sec = cpu_forward_time(setup, po, args.steps, args.warmup)
    value = edges / sec
    snu, sni, se, sd, sL = WORKLOADS[sample]
    desc = (f"same generator law at {sample} scale ({snu}x{sni}, {se} edges, d={sd}, L={sL}), full forward, "
            f"torch {torch.__version__} CPU torch.sparse.mm")
    line = {
        "impl": "reference", "metric": "lightgcn_propagation_edges_per_s", "value": value, "unit": "edges/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": desc},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(name, n_gpus):
    nu, ni, e, d, L = WORKLOADS[name]
    return {"workload": f"{name}: LightGCN {L}-layer d={d} fp32 propagation, synthetic power-law graph "
                        f"{nu} users x {ni} items, {e} edges (nnz(A_hat)={2 * e})",
            "partition": "single GPU" if n_gpus == 1 else f"rows distributed cyclically over {n_gpus} ranks; layer rows "
                                                           "exchanged by P2P stores from the SpMM epilogue (fused "
                                                           "all-gather) or NCCL all-gather (--nccl-allgather)",
            "l2": f"inputs larger than L2 (table {(nu + ni) * d * 4 / 1e9:.2f} GB, CSR {2 * e * 8 / 1e9:.2f} GB)"
            if (nu + ni) * d * 4 + 2 * e * 8 > 2 * 126e6 else "L2 flushed between iterations (256 MB write)"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist

    import gnn_recommendations_b200 as g
    from gnn_recommendations_b200.dist import RowPartition, lightgcn_propagate_sharded
    from gnn_recommendations_b200.synthetic import synth_pairs_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    nu, ni, e, d, L = WORKLOADS[args.workload]
    n = nu + ni
    small = (n * d * 4 + 2 * e * 8) <= 2 * 126e6
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if small else None

    t_setup = time.perf_counter()
    u, i = synth_pairs_device(nu, ni, e, 42, dev)
    full = g.NormAdjCSR.from_pairs(u, i, nu, ni, device=dev)
    del u, i
    nnz = full.nnz
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup

    gen = torch.Generator(device=dev).manual_seed(1234)
    if world == 1:
        with torch.device(dev):
            model = g.LightGCN(nu, ni, embedding_dim=d, n_layers=L, init_scale=0.1)
        csr = full

        def step():
            with torch.no_grad():
                return model.get_all_embeddings(csr)
        n_rows_local = n
    else:
        part = RowPartition(full.indptr, world)
        csr = part.local_csr(full, rank)
        n_loc = part.n_local(rank)
        del full
        torch.cuda.empty_cache()
        x0_local = torch.randn(n_loc, d, device=dev, generator=gen) * 0.1
        exchange = None
        if not args.nccl_allgather:
            try:
                from gnn_recommendations_b200.dist import PeerExchange, lightgcn_propagate_fused

                exchange = PeerExchange(part, d, dev)
            except Exception as err:  # symmetric memory unavailable: NCCL all-gather path
                if rank == 0:
                    print(f"[bench] peer exchange unavailable ({type(err).__name__}: {err}); using NCCL all-gather",
                          file=sys.stderr)
        ok = torch.tensor([1 if exchange is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            exchange = None

        def step():
            if exchange is not None:
                return lightgcn_propagate_fused(csr, exchange, x0_local, L)
            return lightgcn_propagate_sharded(csr, part, rank, x0_local, L)
        n_rows_local = n_loc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
            if flush is not None:
                flush.zero_()
        barrier()
        ms = 0.0
        ctx = sampler if sampler is not None else _Null()
        with ctx:
            if flush is None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    fn()
                e1.record()
                barrier()
                ms = e0.elapsed_time(e1)
            else:
                evs = []
                for _ in range(steps):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn()
                    e1.record()
                    evs.append((e0, e1))
                barrier()
                ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    n_long = int(csr.n_long)
    # ---- value: inputs resident in HBM -------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    csr.timings = None
    for _ in range(args.warmup):
        step()
    csr.timings = []
    launches0 = csr.launches
    ms_step = timed(step, args.steps, 0, sampler)
    launches = csr.launches - launches0
    torch.cuda.synchronize()
    kernel_ms = [a.elapsed_time(b) for a, b in csr.timings]
    csr.timings = None
    value = L * nnz / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (one SpMM launch = one layer on this rank's rows) -----
    hbm_peak, peak_src = peaks()
    b_alg = alg_bytes_per_layer(csr.nnz, n_rows_local, d)
    avg_kernel_ms = float(np.mean(kernel_ms))
    achieved = b_alg / (avg_kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(REPO, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{args.workload}@{world}")
    roofline = {"bound": "hbm", "kernel": f"gr_spmm_csr_f32 (spmm_stream_rows<{d}> + spmm_long_rows<{d}>)",
                "achieved": achieved, "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "algorithmic_bytes_per_launch": b_alg,
                "avg_launch_ms": avg_kernel_ms, "launches_timed": len(kernel_ms),
                "kernel_share_of_step": sum(kernel_ms) / (ms_step * args.steps)}

    # ---- e2e: evaluator.py:76-80 as a user runs it — adjacency arrives as the reference's torch
    # COO in pinned host memory, is uploaded, converted and propagated; embeddings return to host
    e2e = None
    if not args.no_e2e:
        rows_h = csr.row_ids().to(torch.int64)
        cols_h = csr.indices.to(torch.int64)
        idx_host = torch.empty((2, csr.nnz), dtype=torch.int64, pin_memory=True)
        idx_host[0].copy_(rows_h)
        idx_host[1].copy_(cols_h)
        del rows_h, cols_h
        val_host = torch.empty(csr.nnz, dtype=torch.float32, pin_memory=True)
        val_host.copy_(csr.vals)
        out_host = torch.empty((n_rows_local, d), dtype=torch.float32, pin_memory=True)
        shape = (csr.n_rows, csr.n_cols)
        torch.cuda.empty_cache()

        def e2e_step():
            idx_d = idx_host.to(dev, non_blocking=True)
            val_d = val_host.to(dev, non_blocking=True)
            adj = torch.sparse_coo_tensor(idx_d, val_d, shape, check_invariants=False)
            loc = g.NormAdjCSR.from_torch_coo(adj)        # what as_csr() does on a cache miss
            if world == 1:
                with torch.no_grad():
                    ue, ie = model.get_all_embeddings(loc)
                out_host[:nu].copy_(ue, non_blocking=True)
                out_host[nu:].copy_(ie, non_blocking=True)
            else:
                out = (lightgcn_propagate_fused(loc, exchange, x0_local, L) if exchange is not None
                       else lightgcn_propagate_sharded(loc, part, rank, x0_local, L))
                out_host.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(out_host[0, 0])

        ms_serial = timed(e2e_step, args.steps, min(args.warmup, 3))
        ms_e2e, mode = ms_serial, "one step at a time"

        if world == 1:
            # The same steps software-pipelined, as a serving loop runs them: the upload of step i+1
            # (copy stream) and the download of step i-1 (second copy stream) overlap the propagation
            # of step i.  Every step still uploads its own COO and downloads its own embeddings; one
            # device COO buffer is enough because the CSR conversion has consumed it (host-synchronous)
            # before the next upload is enqueued.
            s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
            cur = torch.cuda.current_stream()
            idx_d = torch.empty(idx_host.shape, dtype=torch.int64, device=dev)
            val_d = torch.empty(val_host.shape, dtype=torch.float32, device=dev)

            def pipelined_step():
                with torch.cuda.stream(s_h2d):
                    idx_d.copy_(idx_host, non_blocking=True)
                    val_d.copy_(val_host, non_blocking=True)
                    up = torch.cuda.Event()
                    up.record(s_h2d)
                cur.wait_event(up)
                adj = torch.sparse_coo_tensor(idx_d, val_d, shape, check_invariants=False)
                loc = g.NormAdjCSR.from_torch_coo(adj)      # synchronises: idx_d / val_d are free again
                with torch.no_grad():
                    ue, ie = model.get_all_embeddings(loc)
                done = torch.cuda.Event()
                done.record(cur)
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(done)
                    out_host[:nu].copy_(ue, non_blocking=True)
                    out_host[nu:].copy_(ie, non_blocking=True)
                ue.record_stream(s_d2h)
                ie.record_stream(s_d2h)

            def drain():
                cur.wait_stream(s_d2h)
                cur.wait_stream(s_h2d)

            for _ in range(min(args.warmup, 3)):
                pipelined_step()
            drain()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                pipelined_step()
            drain()
            e1.record()
            barrier()
            ms_e2e = e0.elapsed_time(e1) / args.steps
            mode = ("software-pipelined: upload of step i+1 and download of step i-1 on copy streams overlap the "
                    "propagation of step i; timed over all steps incl. pipeline fill and drain")
            assert float(out_host[0, 0]) == float(out_host[0, 0])
            del idx_d, val_d
        e2e = {"value": L * nnz / (ms_e2e * 1e-3), "unit": "edges/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(idx_host.numel() * 8 + val_host.numel() * 4),
               "d2h_bytes_per_step": int(out_host.numel() * 4),
               "schedule": mode,
               "serial_ms_per_step": ms_serial, "serial_value": L * nnz / (ms_serial * 1e-3),
               "path": "torch COO (int64 indices, f32 values) in pinned host memory -> .to(device) -> "
                       "model.get_all_embeddings(adj) -> embeddings copied to pinned host memory"}
        del idx_host, val_host, out_host

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample = CPU_SAMPLE if e > WORKLOADS[CPU_SAMPLE][2] else args.workload
        setup, edges, po = cpu_forward_setup(sample)
        sec = cpu_forward_time(setup, po, 3, 1)
        snu, sni, se, sd, sL = WORKLOADS[sample]
        cpu = {"value": edges / sec, "unit": "edges/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"same generator law at {sample} scale ({snu}x{sni}, {se} edges, d={sd}, L={sL}), "
                         f"3 full forwards after 1 warm-up, torch CPU torch.sparse.mm (oracle port of "
                         f"lightgcn.py:62-104)"}

    # ---- the rest of the path, small configs (not the headline; N=1 only) -------------------------
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        model = csr = full = None
        torch.cuda.empty_cache()
        extras = run_extras(g, dev)

    if rank == 0:
        line = {
            "metric": "lightgcn_propagation_edges_per_s", "value": value, "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, world),
            "clocks": sampler.summary() if sampler else None,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "setup_s": t_setup, "nnz": nnz, "long_rows": n_long, "extras": extras,
            "exchange": None if world == 1 else (("fused_multicast_stores" if exchange.multicast else "fused_peer_stores")
                                                  if exchange is not None else "nccl_all_gather"),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extras(g, dev):
    """Secondary measurements of the other path stages at the small dataset shapes: full-ranking
    top-20 evaluation (C4, Amazon-Book shape) and one full training epoch with the reference's
    semantics — a complete propagation forward + backward per 512-triple step — (C1, ML-1M shape)."""
    from gnn_recommendations_b200.evaluator import full_rank_topk, seen_csr
    from gnn_recommendations_b200.synthetic import synth_pairs_device

    out = {}
    # --- eval users/s at C4: scores + seen-mask + top-20 for every user, catalogue 91 599 items
    nu, ni, e, d, L = WORKLOADS["C4"]
    u, i = synth_pairs_device(nu, ni, e, 42, dev)
    csr = g.NormAdjCSR.from_pairs(u, i, nu, ni, device=dev)
    with torch.device(dev):
        model = g.LightGCN(nu, ni, embedding_dim=d, n_layers=L, init_scale=0.1)
    with torch.no_grad():
        ue, ie = model.get_all_embeddings(csr)
    eval_users = np.arange(nu)
    ip, it = seen_csr(eval_users, nu, (u.cpu().numpy(), i.cpu().numpy()))
    ip_d, it_d = torch.from_numpy(ip).to(dev), torch.from_numpy(it).to(dev)
    eu_d = torch.from_numpy(eval_users).to(dev)
    def time_topk(tc):
        stats = {}
        for _ in range(2):
            full_rank_topk(ue, ie, eu_d, ip_d, it_d, 20, tensor_cores=tc, stats=stats)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            full_rank_topk(ue, ie, eu_d, ip_d, it_d, 20, tensor_cores=tc, stats=stats)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 3, stats

    ms_exact, _ = time_topk(False)
    ms, st = time_topk(None)
    same = bool(torch.equal(full_rank_topk(ue, ie, eu_d, ip_d, it_d, 20, tensor_cores=False),
                            full_rank_topk(ue, ie, eu_d, ip_d, it_d, 20, tensor_cores=None)))
    out["eval_c4"] = {"users_per_s": nu / (ms * 1e-3), "ms": ms, "users": nu, "items": ni, "k": 20, "d": d,
                      "tensor_cores": bool(st.get("tensor_cores")), "rows_reranked_exactly": st.get("rows_reranked_exactly"),
                      "tf32_tflop_per_s": 2.0 * nu * ni * d / (ms * 1e-3) / 1e12,
                      "exact_only_ms": ms_exact, "exact_only_users_per_s": nu / (ms_exact * 1e-3),
                      "exact_only_fp32_tflop_per_s": 2.0 * nu * ni * d / (ms_exact * 1e-3) / 1e12,
                      "lists_identical_to_exact_kernel": same,
                      "what": "full-ranking top-20 for all users: tcgen05 TF32 nomination (K'=32, two CTAs per SM) + exact fp32 "
                              "re-scoring + exact re-rank of unproven rows; exact_only = FFMA kernel alone"}
    del csr, model, ue, ie
    # --- epoch time at C1: 1 954 steps of sample + propagate + fused BPR + backward + clip + Adam
    nu, ni, e, d, L = WORKLOADS["C1"]
    u, i = synth_pairs_device(nu, ni, e, 42, dev)
    empty = (np.zeros(0, np.int64), np.zeros(0, np.int64))
    ds = g.InteractionDataset((u.cpu().numpy(), i.cpu().numpy()), empty, empty, nu, ni, device=dev, name="C1")
    torch.manual_seed(42)
    model = g.LightGCN(nu, ni, embedding_dim=d, n_layers=L, init_scale=0.1)
    cfg = {"batch_size": 512, "learning_rate": 1e-3, "weight_decay": 1e-4, "use_scheduler": False,
           "checkpoint_dir": "/tmp/gr_bench_ckpt"}
    tr = g.Trainer(model, ds, cfg, device=dev)
    steps = len(u) // 512 + 1
    tr.batch_size = 512
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss = tr.train_epoch()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    out["epoch_c1"] = {"epoch_s": sec, "steps": steps, "ms_per_step": sec / steps * 1e3, "loss": loss,
                       "edge_traversals_per_s": steps * 2 * L * 2 * e / sec,
                       "what": "Trainer.train_epoch at the ML-1M shape (B=512, full propagation fwd+bwd per step, "
                               "host sampler included); reference CPU: 1410.6 s (BASELINE.md)"}
    return out


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GR_BENCH_WORKLOAD", "C5"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--nccl-allgather", action="store_true", help="N>1: NCCL all-gather per layer instead of the "
                    "fused peer-store exchange")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
