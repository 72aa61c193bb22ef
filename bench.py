#!/usr/bin/env python
"""Benchmark of the D-MPNN hot path (BASELINE.json): molecules/sec, forward+backward, of
BondMessagePassing(h=300, depth=3) + MeanAggregation on synthetic ~25-atom molecules.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (torchrun for N>1, NCCL).  A "step" is one pass of the hot path over one batch:
device layout build, forward, dummy scalar loss on the b x h output, backward (all weight
gradients), and for N>1 the flat-bucket gradient all-reduce.  Prints ONE JSON line (rank 0).

 value       whole-job molecules/s with the batch's tensors already resident in HBM
 e2e         same metric through the public API with HOST (pinned) buffers: H2D copies of
             V / E / edge_index / rev_edge_index / batch and the D2H read of the loss are in the timed region
 roofline    depth-step kernel: algorithmic bytes (3*E*h*s, t>=2) / CUDA-event duration vs measured HBM peak
 cpu_baseline  the oracle port (same algorithm as the reference's PyTorch CPU path) on the host cores

`--impl reference` times the reference's CPU implementation of the path (the oracle port; the
reference package itself cannot be imported on the GPU box) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(n_mols=10_000, d_h=300, depth=3, mean_atoms=25.0)
FALLBACK_HBM_GBS = 6650.0


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (nvidia-smi fields via NVML)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        self._nv = self._h = None
        try:                                    # NVML is initialised here, outside the timed region
            import pynvml as nv

            nv.nvmlInit()
            self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def _sample(self):
        nv, h = self._nv, self._h
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for k, bit in names.items():
            if r & bit:
                self.reasons.add(k)

    def _run(self):
        if self._nv is None:
            return
        try:
            while not self._stop.is_set():
                self._sample()
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        # called right after the closing synchronize: one more sample while the clocks are still at their load value
        try:
            if self._nv is not None:
                self._sample()
        except Exception:  # noqa: BLE001
            pass
        self._stop.set()
        self._t.join(timeout=2)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_host_batch(n_mols: int, seed: int, pin: bool, transfer_dtype=None, pack_tiles: bool = True):
    """The batch a loader hands over.  `pack_tiles`: the loader orders the molecules inside the batch so that the
    engine's 128-row tiles come out nearly full (dmpnn_tile_pack_order; same molecules, same work, the order inside a
    batch is the loader's choice -- the reference reshuffles it every epoch)."""
    from chemprop_b200.data import BatchMolGraph, make_molecules, tile_packing_order_of

    mgs = make_molecules(n_mols, seed=seed, mean_atoms=WORKLOAD["mean_atoms"])
    if pack_tiles:
        mgs = [mgs[i] for i in tile_packing_order_of(mgs)]
    return BatchMolGraph(mgs, pin_memory=pin, transfer_dtype=transfer_dtype), mgs


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port, all host threads
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(n_mols: int, seed: int):
    from oracle import restatement as R

    from chemprop_b200.data import make_molecules

    torch.manual_seed(seed)
    mgs = make_molecules(n_mols, seed=seed, mean_atoms=WORKLOAD["mean_atoms"])
    V, E, ei, rev, batch = (torch.from_numpy(x) for x in R.collate(mgs))
    h = WORKLOAD["d_h"]
    lin = lambda o, i: torch.nn.Linear(i, o).weight.detach().requires_grad_(True)  # noqa: E731
    Wi, Wh, Wo = lin(h, 86), lin(h, h), lin(h, 72 + h)
    bo = torch.zeros(h, requires_grad=True)

    def step():
        for p in (Wi, Wh, Wo, bo):
            p.grad = None
        H = R.message_passing_forward("bond", V, E, ei, rev, Wi, None, Wh, None, Wo, bo, WORKLOAD["depth"])
        loss = R.aggregate(H, batch, "mean").square().mean()
        loss.backward()
        return loss.item()

    return step


def run_cpu(n_mols: int, steps: int, warmup: int):
    """Times the oracle port on the host.  torch's CPU scatter / index kernels stop scaling (and regress) well
    before 128 threads, so the thread count is probed (8, 16, 32, all cores: one step each) and the fastest is
    used for the timed steps -- the baseline gets the best configuration the host offers."""
    ncpu = os.cpu_count() or 1
    step = cpu_step_fn(n_mols, seed=1)
    best_t, best_dt = ncpu, float("inf")
    for nt in sorted({min(8, ncpu), min(16, ncpu), min(32, ncpu), ncpu}):
        torch.set_num_threads(nt)
        step()                                   # warm-up at this thread count
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if dt < best_dt:
            best_t, best_dt = nt, dt
    torch.set_num_threads(best_t)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n_mols / dt, dt


def time_collate(n_mols: int = 4000):
    """Host-side batch assembly, molecules/s (the reference collate is timed beside the path): the
    reference's collate loop (oracle restatement of collate.py:37-62, numpy), our C collate of the same MolGraphs, and
    the packed data set's gather into a reused staging buffer.  CPU only, bounded (a few seconds)."""
    from oracle import restatement as R

    from chemprop_b200.data import BatchMolGraph, HostBatchBuffer, PackedMolGraphDataset, make_molecules

    mgs = make_molecules(n_mols, seed=3, mean_atoms=WORKLOAD["mean_atoms"])

    def best(f, n=3):
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            f()
            ts.append(time.perf_counter() - t0)
        return n_mols / min(ts)

    ds = PackedMolGraphDataset.from_molgraphs(mgs)
    ids = np.random.default_rng(0).permutation(n_mols)
    buf = HostBatchBuffer(ds.d_v, ds.d_e)
    ds.batch(ids, buffer=buf)
    return {"unit": "molecules/s", "sample": f"{n_mols} molecules, best of 3",
            "reference_collate_port": best(lambda: R.collate_torch(mgs)),
            "c_collate": best(lambda: BatchMolGraph(mgs)),
            "packed_dataset_gather": best(lambda: ds.batch(ids, buffer=buf))}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 1000
    v, dt = run_cpu(sample, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "molecules/sec fwd+bwd (h=300 d=3)", "value": v, "unit": "molecules/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: BondMessagePassing h=300 depth=3 + MeanAggregation, ~25-atom synthetic mols",
                   "timed_sample": f"{sample} molecules per step (bounded sample of the 10k-molecule batch)"},
        "cpu_baseline": {"value": v, "unit": "molecules/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} molecules x {args.steps} steps, torch CPU, {cores} threads "
                                   f"(fastest of 8/16/32/{os.cpu_count()} threads)"},
        "e2e": {"value": v, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main_gpu(args):
    import torch.distributed as dist

    from chemprop_b200 import _lib, engine
    from chemprop_b200.nn import BondMessagePassing, MeanAggregation
    from chemprop_b200.parallel import FlatGradAllReducer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    assert lib.dmpnn_device_ok() == 1, "bench needs an sm_100 device"

    n_mols = args.mols
    precision = args.precision
    torch.manual_seed(0)  # identical weights on every rank
    mp = BondMessagePassing(d_h=WORKLOAD["d_h"], depth=WORKLOAD["depth"], precision=precision).to(dev)
    mp.fused = not args.no_fused
    agg = MeanAggregation()
    params = list(mp.parameters())
    reducer = FlatGradAllReducer(params)

    # weak scaling: every rank its own batch.  The pinned host batch is what a loader hands over: for the bf16 tier
    # it carries the compact transfer copy (bf16 features, int32 indices -- result-identical, see BatchMolGraph);
    # `host_f32` is the same batch in the reference's f32 / int64 host format, timed as a second e2e figure.
    host_bmg, mgs = make_host_batch(n_mols, seed=1 + rank, pin=True,
                                    transfer_dtype=torch.bfloat16 if precision == "bf16" else None,
                                    pack_tiles=not args.no_pack)
    n_tiles = host_bmg._meta_host[_lib.META_N_TILES] if host_bmg._meta_host else None
    V_atoms, E_rows = host_bmg.V.shape[0], host_bmg.E.shape[0]
    h2d_bytes = host_bmg.transfer_nbytes()

    from chemprop_b200.data import BatchMolGraph

    host_f32 = BatchMolGraph(mgs, pin_memory=True) if precision == "bf16" else host_bmg
    del mgs
    src = {"bmg": host_bmg}

    def to_device():
        return src["bmg"].cuda_copy(dev, non_blocking=True)

    resident = to_device()

    def step(bmg):
        bmg._layout = None                       # the device layout build is part of every step
        for p in params:
            p.grad = None
        H = mp(bmg)
        loss = agg(H, bmg.batch).float().square().mean()
        loss.backward()
        reducer.allreduce_()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput --------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step(resident)
    engine.STEP_EVENTS = []
    l0 = lib.dmpnn_launch_count()
    with ClockSampler(local) as clocks:
        ms = timed(lambda: step(resident), args.steps)
    launches = lib.dmpnn_launch_count() - l0
    step_events, engine.STEP_EVENTS = engine.STEP_EVENTS, None
    ms_per_step = ms / args.steps
    value = world * n_mols / (ms_per_step * 1e-3)

    # ---- end to end: host buffers in, loss out ---------------------------------------------
    # Every step copies its own inputs from pinned host memory and reads its loss back.  As a training loop
    # with a pinned-memory loader does, the copy of step i+1 is issued on a side stream while step i computes;
    # all K copies and K loss reads happen inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)

    def issue_copy():
        with torch.cuda.stream(copy_stream):
            b = to_device()
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return b, ev

    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()

    def e2e_loop(k):
        # The loss of step i is copied to pinned host memory right after its backward is queued and READ by the host
        # one step later (after step i+1 is queued), as a logging training loop does: the GPU queue never drains on
        # the host's read.  All k H2D batches and all k D2H loss reads are inside the timed region.
        nxt = issue_copy()
        pending = None
        losses = []
        for i in range(k):
            bmg, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            for t in (bmg.V, bmg.E, bmg.edge_index, bmg.rev_edge_index, bmg.batch):
                t.record_stream(torch.cuda.current_stream())
            if i + 1 < k:
                nxt = issue_copy()
            loss = step(bmg)
            slot = loss_host[i & 1:(i & 1) + 1]
            slot.copy_(loss.detach().float().reshape(1), non_blocking=True)   # D2H of the step's result
            done = torch.cuda.Event()
            done.record()
            if pending is not None:
                pending[0].synchronize()
                losses.append(float(pending[1][0]))
            pending = (done, slot)
        pending[0].synchronize()
        losses.append(float(pending[1][0]))
        assert len(losses) == k and all(np.isfinite(losses))

    e2e_loop(2)
    ms_e2e = timed(lambda: e2e_loop(args.steps), 1) / args.steps
    e2e_value = world * n_mols / (ms_e2e * 1e-3)
    e2e_f32 = None
    if host_f32 is not host_bmg:
        src["bmg"] = host_f32
        e2e_loop(2)
        ms_f32 = timed(lambda: e2e_loop(args.steps), 1) / args.steps
        e2e_f32 = {"value": world * n_mols / (ms_f32 * 1e-3), "unit": "molecules/s", "ms_per_step": ms_f32,
                   "h2d_bytes_per_step": host_f32.transfer_nbytes(), "d2h_bytes_per_step": 4,
                   "host_format": "f32 features + int64 indices (the reference's BatchMolGraph dtypes)"}
        src["bmg"] = host_bmg

    # ---- roofline of the depth step ----------------------------------------------------------
    s = 2 if precision == "bf16" else 4
    h = WORKLOAD["d_h"]
    by_tag = {}
    for tag, a, b in step_events:
        by_tag.setdefault(tag, []).append(a.elapsed_time(b))
    tag = next((t for t in ("fused", "unfused") if t in by_tag), None)
    peak, peak_src = hbm_peak()
    roofline = None
    if tag:
        dur_ms = statistics.mean(by_tag[tag])
        alg_bytes = 3 * E_rows * h * s + 12 * E_rows + 4 * V_atoms
        ach = alg_bytes / (dur_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "fused_step_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tag == "fused" and int(tj.get("directed_edges", -1)) == E_rows and tj.get("precision") == precision:
                    traffic = tj["dram_bytes_per_launch"]
            except Exception:
                pass
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "kernel": f"bond depth step t>=2 ({tag})", "launch_ms": dur_ms,
                    "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                    "traffic_source": ("profiles/fused_step_traffic.json: ncu --set full capture of round 1 (same rows, "
                                       "generator molecule order)") if traffic is not None else None,
                    "first_step_ms": statistics.mean(by_tag.get(tag + "_first", [float("nan")]))}

    # ---- resident data set: ids in, loss out -----------------------------------------------
    # The data set's packed arrays live in HBM; a step uploads the molecule ids + output offsets (24 B / molecule),
    # assembles its BatchMolGraph with one gather launch and runs the same fwd+bwd.  Extra figure, N = 1 only, measured
    # last and guarded: it must never take the contract's numbers down with it.
    resident_ds = None
    if world == 1 and not args.no_dataset:
        try:
            from chemprop_b200.data import PackedMolGraphDataset, make_molecules

            pool = 3 * n_mols
            ds = PackedMolGraphDataset.from_molgraphs(make_molecules(pool, seed=7, mean_atoms=WORKLOAD["mean_atoms"])).to(dev)
            rng = np.random.default_rng(0)
            id_sets = [rng.permutation(pool)[:n_mols] for _ in range(4)]

            def ds_loop(k):
                losses = []
                for i in range(k):
                    ids = id_sets[i % len(id_sets)]
                    loss = step(ds.batch(ids if args.no_pack else ds.packed_order(ids)))
                    losses.append(loss.detach())
                vals = torch.stack(losses).float().cpu()          # one D2H read of the k losses, inside the timed region
                assert bool(torch.isfinite(vals).all())

            ds_loop(3)
            ms_ds = timed(lambda: ds_loop(args.steps), 1) / args.steps
            resident_ds = {"value": n_mols / (ms_ds * 1e-3), "unit": "molecules/s", "ms_per_step": ms_ds,
                           "h2d_bytes_per_step": 8 * (3 * n_mols + 2), "d2h_bytes_per_step": 4,
                           "dataset_molecules": pool, "dataset_bytes_in_hbm": ds.nbytes(),
                           "note": "batch assembled on the GPU from the HBM-resident packed data set (random ids per step)"}
        except Exception as e:  # noqa: BLE001
            resident_ds = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (rank 0, N=1 only; bounded sample) ---------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        sample = 1000
        v, dt = run_cpu(sample, 3, 1)
        cpu = {"value": v, "unit": "molecules/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{sample} molecules x 3 steps (1 warm-up), oracle restatement on torch CPU, "
                         f"{torch.get_num_threads()} threads (fastest of 8/16/32/{os.cpu_count()})"}

    host_assembly = None
    if world == 1 and not args.no_cpu:
        try:
            host_assembly = time_collate()
        except Exception as e:  # noqa: BLE001
            host_assembly = {"error": f"{type(e).__name__}: {e}"[:200]}

    line = {
        "metric": "molecules/sec fwd+bwd (h=300 d=3)", "value": value, "unit": "molecules/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32",
        "data": "synthetic",
        "config": {"workload": f"C2: {n_mols} synthetic mols/GPU (~25 atoms), BondMessagePassing h=300 depth=3 + "
                               "MeanAggregation, fwd+bwd incl. device layout build",
                   "atoms": V_atoms, "directed_edges": E_rows, "precision": precision,
                   "parallelism": f"dp{world}", "fused_depth_step": tag == "fused",
                   "molecule_order": ("loader tile packing (best-fit decreasing on edge counts, dmpnn_tile_pack_order)"
                                      if not args.no_pack else "generator order"),
                   "tiles": n_tiles, "tile_fill": (E_rows / (128.0 * n_tiles)) if n_tiles else None,
                   "step_sync_free": bool(engine.HOST_META),
                   "l2": "working set (>=0.3 GB hidden buffers per step) exceeds the 126 MB L2; no explicit flush"},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "molecules/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "host_format": ("bf16 features + int32 indices staged by BatchMolGraph(transfer_dtype=bfloat16); the bf16 "
                                "tier rounds V/E to bf16 on the GPU anyway, results are bit-identical"
                                if precision == "bf16" else "f32 features + int64 indices")},
        "e2e_f32_host": e2e_f32,
        "resident_dataset": resident_ds,
        "host_batch_assembly": host_assembly,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mols", type=int, default=WORKLOAD["n_mols"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-dataset", action="store_true", help="skip the resident-data-set figure")
    ap.add_argument("--no-pack", action="store_true", help="keep the generator's molecule order (no tile packing)")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
