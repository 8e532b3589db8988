#!/usr/bin/env python
"""bench.py -- edges/sec per layer (forward + backward) of the fused hot path on B200.

Workload at every N (BASELINE.json configs[1], the largest single-GPU configuration the metric is
quoted on): SchNet, 6 interactions, hidden 128, 50 RBF, cutoff 5 A, 4096 synthetic molecules x 32
atoms PER GPU (weak scaling, molecules sharded across ranks; the only collective is the gradient
all-reduce after backward).  One step = forward + backward of the whole model on one batch;
value = (edges of all ranks) x (6 layers) / (max-over-ranks device time per step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision fp32|bf16]

`--impl reference` times the CPU oracle port (the reference's PyTorch path with the missing wheels
restated, oracle/) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "edges/sec per layer fwd+bwd"
UNIT = "edges/s"
CFG = dict(hidden=128, filters=128, layers=6, gaussians=50, cutoff=5.0, molecules=4096, atoms=32, box=8.0,
           max_num_neighbors=32)
WORKLOAD = ("SchNet 6 interactions, hidden 128, 50 RBF, cutoff 5 A, 4096 molecules x 32 atoms per GPU "
            "(BASELINE.json configs[1])")


def synth(molecules: int, seed: int):
    """pos ~ U(0, 8 A)^3, atoms ~ randint(1, 10), CPU generator."""
    import torch
    g = torch.Generator().manual_seed(seed)
    n = molecules * CFG["atoms"]
    pos = torch.rand(n, 3, generator=g) * CFG["box"]
    atoms = torch.randint(1, 10, (n,), generator=g)
    batch = torch.arange(molecules).repeat_interleave(CFG["atoms"])
    return atoms, pos, batch


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(molecules: int):
    import torch
    from oracle import ref_layers as R
    from oracle.thirdparty import cluster
    torch.manual_seed(0)
    atoms, pos, batch = synth(molecules, 0)
    ei = torch.from_numpy(cluster.radius_graph(pos.numpy(), CFG["cutoff"], batch.numpy(), False, CFG["max_num_neighbors"]))
    model = R.SchNetModel(hidden_channels=CFG["hidden"], num_filters=CFG["filters"], num_layers=CFG["layers"],
                          num_gaussians=CFG["gaussians"], cutoff=CFG["cutoff"])
    b = Bag(atoms=atoms, pos=pos, edge_index=ei, batch=batch)

    def step():
        model.zero_grad(set_to_none=True)
        model(b).sum().backward()
    return step, ei.shape[1]


def run_cpu(steps: int, warmup: int, molecules: int = 256, budget_s: float = 25.0):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, E = cpu_step_fn(molecules)
    for _ in range(max(1, warmup)):
        step()
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        step()
        done += 1
        if time.perf_counter() - t0 > budget_s and done >= 3:
            break
    dt = (time.perf_counter() - t0) / done
    return dict(value=E * CFG["layers"] / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"oracle/ (PyTorch CPU port of the reference path), {molecules} molecules x 32 atoms, E={E}, "
                       f"{done} steps of fwd+bwd over 6 layers, {torch.get_num_threads()} threads",
                ms_per_step=dt * 1e3, steps=done)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu(args.steps, args.warmup)
    line = dict(impl="reference", metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=r["steps"],
                warmup=args.warmup, ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", config={"workload": WORKLOAD, "cpu_sample": r["sample"]},
                cpu_baseline={k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                e2e={"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main_gpu(args):
    import torch
    import torch.distributed as dist
    import gmp_b200
    from gmp_b200 import _lib

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"

    # bf16 mode: node-side library GEMMs on TF32 tensor cores (fp32-strict keeps full fp32 everywhere)
    gmp_b200.set_fast_matmul(args.precision == "bf16")
    torch.manual_seed(0)
    model = gmp_b200.SchNetModel(hidden_channels=CFG["hidden"], num_filters=CFG["filters"], num_layers=CFG["layers"],
                                 num_gaussians=CFG["gaussians"], cutoff=CFG["cutoff"], precision=args.precision).to(dev)
    params = [p for p in model.parameters()]
    atoms_h, pos_h, batch_h = synth(CFG["molecules"], seed=rank)  # graph-sharded: every rank owns its molecules
    atoms, pos, batch = atoms_h.to(dev), pos_h.to(dev), batch_h.to(dev)
    ei = gmp_b200.radius_graph(pos, CFG["cutoff"], batch, max_num_neighbors=CFG["max_num_neighbors"])
    E, N = ei.shape[1], pos.shape[0]
    # num_graphs: the field PyG's Batch carries; with it the pooling read-out needs no batch.max() read-back (a host sync)
    b = Bag(atoms=atoms, pos=pos, edge_index=ei, batch=batch, num_graphs=CFG["molecules"])

    def allreduce_grads(grads=None):
        if world > 1:
            grads = [p.grad for p in params if p.grad is not None] if grads is None else grads
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat)
            # (gradients stay in `flat`; an optimizer would consume them from here)

    def step_eager(batch_obj):
        for p in params:
            p.grad = None
        out = model(batch_obj)
        out.sum().backward()
        allreduce_grads()
        return out

    # The step is ~270 launches for ~11 ms of device work and the Python / ATen / ctypes launch path needs about as long
    # to issue them, so forward + backward are captured once into a CUDA graph (gmp_b200.GraphedStep) and replayed; the
    # gradient all-reduce stays outside the graph.  --eager (or a failed capture) falls back to launch-by-launch.
    graph_note = None
    gs = None
    if not args.eager:
        try:
            gs = gmp_b200.GraphedStep(model, b, warmup=max(3, args.warmup))
        except Exception as ex:  # noqa: BLE001 -- report and measure eagerly rather than lose the run
            graph_note = f"capture failed: {type(ex).__name__}: {str(ex)[:200]}"
            torch.cuda.synchronize()

    def step(batch_obj):
        if gs is None or batch_obj is not b:
            return step_eager(batch_obj)
        out = gs.replay()
        allreduce_grads(gs.grads)
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, K = max(3, args.warmup), args.steps
    for _ in range(W):
        step(b)
    barrier()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    launches0 = _lib.kernel_launches()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s.record()
    for _ in range(K):
        step(b)
    e.record()
    barrier()
    ms = s.elapsed_time(e) / K
    launches = gs.kernels_per_replay if gs is not None else (_lib.kernel_launches() - launches0) // K

    # ---- e2e: host buffers -> public API -> host result, copies inside the timed region ------------------
    pin = lambda t: t.pin_memory()
    atoms_p, pos_p, batch_p, ei_p = pin(atoms_h), pin(pos_h), pin(batch_h), pin(ei.cpu())
    h2d = sum(t.numel() * t.element_size() for t in (atoms_p, pos_p, batch_p, ei_p))
    out_host = torch.empty(CFG["molecules"], 1).pin_memory()

    host_batch = gmp_b200.Batch(atoms=atoms_p, pos=pos_p, batch=batch_p, edge_index=ei_p, num_graphs=CFG["molecules"])
    gs2 = None
    if gs is not None:
        # static device buffers that every step overwrites from the pinned host batch; the captured graph contains the
        # CSR sort of the freshly copied edge_index as well as forward + backward (nothing is reused between steps)
        try:
            static = host_batch.to(dev)
            torch.cuda.synchronize()
            gs2 = gmp_b200.GraphedStep(model, static, warmup=2, rebuild_graph=True)
        except Exception as ex:  # noqa: BLE001
            graph_note = f"e2e capture failed: {type(ex).__name__}: {str(ex)[:200]}"
            torch.cuda.synchronize()

    uploader = gmp_b200.DevicePrefetcher((), dev, static=True)

    def e2e_steps(n):
        # the user-level loop `for batch in loader: step(batch)`.  Graph mode: DevicePrefetcher(static=True) uploads host
        # batch i+1 into one of two staging sets on a side stream while step i runs; load() copies the staged batch into the
        # graph's inputs (device to device) and the replay sorts it into CSR form, runs forward + backward.  All n uploads
        # (the first one not overlapped) and n result read-backs are issued inside this call.
        if gs2 is not None:
            uploader.batches = (host_batch for _ in range(n))   # same uploader every time: its staging buffers persist
            for staged in uploader:
                gs2.load(staged)
                out = gs2.replay()
                allreduce_grads(gs2.grads)
                out_host.copy_(out.detach(), non_blocking=True)
        else:
            for _ in range(n):
                out = step_eager(host_batch.to(dev, non_blocking=True))
                out_host.copy_(out.detach(), non_blocking=True)

    e2e_steps(W + 3)  # also lets the caching allocator reach its steady state for the per-step buffers
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    e2e_steps(K)
    e2.record()
    barrier()
    ms_e2e = s2.elapsed_time(e2) / K
    clk = clocks.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel, timed alone with CUDA events on the launch stream ---------------
    roof = dominant_kernel_roofline(model, b, E, N, dev, args) if rank == 0 else None

    t = torch.tensor([ms, ms_e2e, float(E)], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ms_e2e, E_total = tmax[0].item(), tmax[1].item(), tsum[2].item()
    else:
        E_total = float(E)
    strict = None
    if world == 1 and args.precision == "bf16" and not args.no_strict:
        # the same step in the fp32-strict mode (1e-5 parity mode), for the record
        gmp_b200.set_fast_matmul(False)
        for blk in model.interactions:
            blk.conv.precision = "fp32"
        for _ in range(2):
            step_eager(b)
        barrier()
        s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s3.record()
        for _ in range(3):
            step_eager(b)
        e3.record()
        barrier()
        ms3 = s3.elapsed_time(e3) / 3
        strict = {"precision": "fp32 (FFMA, 1e-5 vs reference)", "ms_per_step": ms3, "value": float(E) * CFG["layers"] / (ms3 * 1e-3),
                  "unit": UNIT}
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = run_cpu(steps=10, warmup=1)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        L = CFG["layers"]
        line = dict(metric=METRIC, value=E_total * L / (ms * 1e-3), unit=UNIT, n_gpus=world, steps=K, warmup=W,
                    ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32" if args.precision == "fp32" else "bf16", data="synthetic",
                    config={"workload": WORKLOAD, "nodes_per_gpu": N, "edges_per_gpu": E, "layers": L,
                            "precision": args.precision,
                            "tolerance_vs_fp32_reference": 1e-2 if args.precision == "bf16" else 1e-5,
                            "node_side_gemms": "cuBLAS TF32 forward / dx; dW, db on tcgen05 (linear_wgrad_tc_kernel)" if args.precision == "bf16" else "cuBLAS fp32", "parallelism": f"graph-sharded x{world}",
                            "l2": "per-step working set (x1/agg/g rows of 6 layers, >2 GB) exceeds the 126 MB L2",
                            "cuda_graph": (gs is not None) if graph_note is None else graph_note,
                            "e2e_path": "pinned host batch -> staging buffers (upload of batch i+1 under step i) -> graph inputs -> one CUDA graph "
                                        "(CSR sort + forward + backward) -> host result" if gs2 is not None else "host batch -> model(batch) eagerly -> host result"},
                    roofline=roof, cpu_baseline=cpu, fp32_strict=strict,
                    e2e={"value": E_total * L / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                         "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e2e},
                    gpu_launches=int(launches), clocks=clk)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


TRAFFIC_FWD_KEEP = 124.33e6 + 477.71e6   # dram read + write of one launch (ncu, profiles/r01_ncu_hw_schnet_fwd_tc2_keep.csv)


def dominant_kernel_roofline(model, b, E, N, dev, args):
    """Time the fused CFConv forward kernel (the kernel launched twice per layer per step: messages and
    dL/dx1) alone: CUDA events around each launch on torch's current stream.  Algorithmic bytes per
    launch: E*(4F + 8) [x1 row gather + col + distance] + N*(4F + 4) [agg row + rowptr]."""
    import ctypes as C
    import torch
    import gmp_b200
    from gmp_b200._lib import SchnetFilter, call, ptr
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, which = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, which = 6650.0, "fallback (B200_PROFILING.md)"
    F = CFG["filters"]
    g = gmp_b200.get_graph(b.edge_index, N)
    csr = g.by_dst
    blk = model.interactions[0]
    sm = model.distance_expansion
    w1, b1, w2, b2 = blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias
    filt = SchnetFilter(ptr(w1), ptr(b1), ptr(w2), ptr(b2), CFG["gaussians"], F, CFG["cutoff"], ptr(sm.offset), sm.coeff)
    with torch.no_grad():
        ew = gmp_b200.edge_length(b.pos, g)
    x1 = torch.randn(N, F, device=dev)
    agg = torch.empty(N, F, device=dev)
    prec = 0 if args.precision == "fp32" else 1
    lib = gmp_b200._lib.lib()
    x1b = x1.to(torch.bfloat16)
    head = torch.empty(lib.gmp_schnet_tc2_num_chunks(E), F, device=dev)
    rowid = csr.row_ids()
    keep = torch.empty(E, F, dtype=torch.bfloat16, device=dev)   # training keeps the per-edge filter values for the backward,
    keep_row = g.by_src.inv_perm()                               # in the order its gather-multiply-reduce reads them
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times = []
    for it in range(3 + 10):
        flush.zero_()  # write 256 MB > L2 between launches
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if prec == 1:   # the entry point the bf16 model path calls (zeroes agg, runs the pipelined kernel and the boundary fix-up)
            call("gmp_schnet_cfconv_fwd_tc2_keep", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(rowid), N, E, ptr(ew), ptr(x1b),
                 C.byref(filt), ptr(agg), ptr(head), ptr(keep), ptr(keep_row))
        else:
            call("gmp_schnet_cfconv_fwd", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, N, E, ptr(ew), None, ptr(x1),
                 C.byref(filt), ptr(agg), prec)
        e.record()
        torch.cuda.synchronize()
        if it >= 3:
            times.append(s.elapsed_time(e))
    ms = sum(times) / len(times)
    alg = E * (4 * F + 8) + N * (4 * F + 4)
    achieved = alg / (ms * 1e-3) / 1e9
    flops = E * (2 * 64 * F + 2 * F * F + 2 * F)  # padded-G GEMM1 + GEMM2 + message product
    if prec == 1:
        kname = ("schnet_fwd_tc2_kernel (tcgen05, bf16, pipelined; the training variant that also stores the per-edge filter values; "
                 "time includes the agg memset and the boundary fix-up) via gmp_schnet_cfconv_fwd_tc2_keep")
        note = (f"{flops / (ms * 1e-3) / 1e12:.1f} TFLOP/s on the filter-MLP GEMMs (bf16 tcgen05); algorithmic bytes are SURVEY 8d's "
                "fp32 figure -- the kernel itself gathers x1 as bf16 rows (256 B/edge, L2-resident) and, in training, writes 256 B/edge of "
                "filter values that are not algorithmic bytes (the plain forward: 0.321 ms, 48.1 %); the kernel is bound by load/store-unit "
                "and MIO cycles (lane-per-row accesses, one ex2 per softplus / Gaussian), tensor pipe 12 %, not by HBM "
                "(profiles/r01f_summary.md)")
    else:
        kname = "schnet_fwd_kernel<128> (fp32 FFMA) via gmp_schnet_cfconv_fwd"
        note = f"{flops / (ms * 1e-3) / 1e12:.1f} TFLOP/s on the filter-MLP GEMMs (fp32 FFMA)"
    return {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of schnet_fwd_tc2_kernel, one launch, ncu hardware-counter pass
            # (profiles/r01_ncu_hw_schnet_fwd_tc2_keep.csv; without the kept filter values 124.3 + 28.6 MB,
            # profiles/r01_ncu_hw_schnet_fwd_tc2.csv)
            "traffic": TRAFFIC_FWD_KEEP if prec == 1 else None, "peak_source": which, "ms_per_launch": ms, "algorithmic_bytes": alg, "note": note}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gmp_b200", choices=["gmp_b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"],
                    help="bf16: filter-MLP GEMMs on tcgen05 with bf16 operands / fp32 accumulation (1e-2 vs the fp32 reference, "
                         "the tolerance BASELINE.json states for bf16 MLP inputs); fp32: strict FFMA path (1e-5)")
    ap.add_argument("--eager", action="store_true", help="launch the step kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strict", action="store_true", help="skip the secondary fp32-strict measurement")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_gpu(a)
