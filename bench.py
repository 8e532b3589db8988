#!/usr/bin/env python
"""bench.py -- edges/sec per layer (forward + backward) of the fused hot path on B200.

Headline workload at every N (BASELINE.json configs[1], the largest single-GPU configuration the metric is
quoted on): SchNet, 6 interactions, hidden 128, 50 RBF, cutoff 5 A, 4096 synthetic molecules x 32
atoms PER GPU (weak scaling, molecules sharded across ranks; the only collective is the gradient
all-reduce after backward).  One step = forward + backward of the whole model on one batch;
value = (edges of all ranks) x (6 layers) / (max-over-ranks device time per step).

The same JSON line carries, under "configs", BASELINE.json configs[2..4] measured by the same process(es):
  3_tfn    TFN l_max=2, 64 channels, 4 layers, 2048 clouds x 64 points  (graph-sharded over the ranks, strong)
  4_mace   MACE l_max=2, correlation 3, 128 channels, 2 interactions, 1024 clouds x 64 (graph-sharded, strong)
  5_egnn   EGNN 4 layers on ONE random radius graph (density 8, r = 1), destination-partitioned slabs with
           halo exchange (strong scaling: the graph is fixed, ranks split it); at N >= 2 preceded by a
           partitioned-vs-single parity self-check on a small cube.
Each block has ms_per_step, edges/s/layer, precision and a roofline fraction from SURVEY.md 8d's formulas.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision fp32|bf16] [--only 2,3,4,5]

`--impl reference` times the CPU oracle port (the reference's PyTorch path with the missing wheels
restated, oracle/) on the host cores on the full config-2 batch.
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
# BASELINE.json configs[2], [3]: point clouds of 64 nodes in a 4^3 box, r = 2
CLOUDS = {"tfn": dict(clouds=2048, layers=4, emb=64), "mace": dict(clouds=1024, layers=2, emb=128)}
CUBE_LOG2N = 22          # config 5 geometry at the largest node count whose 4-layer training step fits one B200 (2^24 does not)


# ------------------------------------------------------------------------------------------------
# synthetic workloads; tests/test_gpu_config_parity.py takes its subsets from these same batches
# ------------------------------------------------------------------------------------------------
def synth(molecules: int, seed: int):
    """config 2: pos ~ U(0, 8 A)^3, atoms ~ randint(1, 10), CPU generator."""
    import torch
    g = torch.Generator().manual_seed(seed)
    n = molecules * CFG["atoms"]
    pos = torch.rand(n, 3, generator=g) * CFG["box"]
    atoms = torch.randint(1, 10, (n,), generator=g)
    batch = torch.arange(molecules).repeat_interleave(CFG["atoms"])
    return atoms, pos, batch


def synth_clouds(clouds: int, seed: int = 0):
    """configs 3 / 4: `clouds` x 64 points, pos ~ U(0, 4)^3, atoms = 0 (in_dim = 1), r = 2."""
    import torch
    g = torch.Generator().manual_seed(seed)
    pos = torch.rand(clouds * 64, 3, generator=g) * 4.0
    batch = torch.arange(clouds).repeat_interleave(64)
    return torch.zeros(clouds * 64, dtype=torch.long), pos, batch


def synth_cube(log2n: int, seed: int = 0):
    """config 5 geometry: 2^log2n points, density 8 per unit volume, sorted along x (slab partition), r = 1."""
    import torch
    n = 2 ** log2n
    side = (n / 8.0) ** (1.0 / 3.0)
    g = torch.Generator().manual_seed(seed)
    pos = torch.rand(n, 3, generator=g) * side
    return pos[torch.argsort(pos[:, 0])].contiguous()


def make_model(which: str, precision: str):
    import gmp_b200
    if which == "schnet":
        return gmp_b200.SchNetModel(hidden_channels=CFG["hidden"], num_filters=CFG["filters"], num_layers=CFG["layers"],
                                    num_gaussians=CFG["gaussians"], cutoff=CFG["cutoff"], precision=precision)
    if which == "tfn":
        return gmp_b200.TFNModel(r_max=2.0, max_ell=2, emb_dim=64, num_layers=4, precision=precision)
    if which == "mace":
        return gmp_b200.MACEModel(r_max=2.0, max_ell=2, correlation=3, emb_dim=128, num_layers=2, precision=precision)
    raise ValueError(which)


def make_oracle_model(which: str):
    from oracle import ref_layers as R
    if which == "schnet":
        return R.SchNetModel(hidden_channels=CFG["hidden"], num_filters=CFG["filters"], num_layers=CFG["layers"],
                             num_gaussians=CFG["gaussians"], cutoff=CFG["cutoff"])
    if which == "tfn":
        return R.TFNModel(r_max=2.0, max_ell=2, emb_dim=64, num_layers=4)
    if which == "mace":
        return R.MACEModel(r_max=2.0, max_ell=2, correlation=3, emb_dim=128, num_layers=2)
    raise ValueError(which)


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d["bf16_tflops_sustained"], source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback (B200_PROFILING.md)")


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
# CPU arm: the oracle port on host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(molecules: int):
    import torch
    from oracle.thirdparty import cluster
    torch.manual_seed(0)
    atoms, pos, batch = synth(CFG["molecules"], 0)            # the bench batch of rank 0 ...
    n = molecules * CFG["atoms"]
    atoms, pos, batch = atoms[:n], pos[:n], batch[:n]         # ... or its first `molecules` molecules (graphs are independent)
    ei = torch.from_numpy(cluster.radius_graph(pos.numpy(), CFG["cutoff"], batch.numpy(), False, CFG["max_num_neighbors"]))
    model = make_oracle_model("schnet")
    b = Bag(atoms=atoms, pos=pos, edge_index=ei, batch=batch)

    def step():
        model.zero_grad(set_to_none=True)
        model(b).sum().backward()
    return step, ei.shape[1]


def run_cpu(steps: int, warmup: int, molecules: int, budget_s: float):
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
    full = molecules == CFG["molecules"]
    return dict(value=E * CFG["layers"] / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"oracle/ (PyTorch CPU port of the reference path), {'the full config-2 batch: ' if full else 'first '}{molecules} molecules x 32 atoms"
                       f"{'' if full else ' of the bench batch'}, E={E}, {done} steps of fwd+bwd over 6 layers, {torch.get_num_threads()} threads",
                ms_per_step=dt * 1e3, steps=done, full_size=full)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # config 2 is timed at full size on the CPU (about 10 s per step on 16 cores: a few steps)
    r = run_cpu(min(args.steps, 5), min(args.warmup, 1), molecules=args.cpu_molecules or CFG["molecules"], budget_s=100.0)
    line = dict(impl="reference", metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=r["steps"],
                warmup=min(args.warmup, 1), ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", config={"workload": WORKLOAD, "cpu_sample": r["sample"]},
                cpu_baseline={k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                e2e={"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Dist:
    """rank / world / device plus the three collectives the timing needs."""

    def __init__(self, gpus):
        import torch
        import torch.distributed as dist
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        assert self.world == gpus, f"--gpus {gpus} but WORLD_SIZE={self.world} (launch with torch.distributed.run)"

    def barrier(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(self, vals, op="max"):
        import torch
        import torch.distributed as dist
        t = torch.tensor(vals, device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return t.tolist()

    def timed(self, fn, warmup, steps):
        """ms per call of fn: barrier + synchronize on both sides, CUDA events, max over ranks."""
        import torch
        for _ in range(warmup):
            fn()
        self.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        self.barrier()
        return self.reduce([s.elapsed_time(e) / steps])[0]


def main_gpu(args):
    import torch
    import torch.distributed as dist
    import gmp_b200
    from gmp_b200 import _lib

    D = Dist(args.gpus)
    rank, world, dev = D.rank, D.world, D.dev
    only = set(int(x) for x in args.only.split(",")) if args.only else {2, 3, 4, 5}
    line = {}
    if 2 in only:
        line = bench_config2(D, args)
    else:
        line = dict(metric=METRIC, value=None, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, config={"workload": WORKLOAD})
    blocks = {}
    for cid, name, fn in ((3, "3_tfn", lambda: bench_clouds(D, "tfn", args)), (4, "4_mace", lambda: bench_clouds(D, "mace", args)),
                          (5, "5_egnn", lambda: bench_cube(D, args))):
        if cid not in only:
            continue
        try:
            blocks[name] = fn()
        except Exception as ex:  # noqa: BLE001 -- one failed block must not lose the line
            import traceback
            blocks[name] = {"error": f"{type(ex).__name__}: {str(ex)[:300]}", "trace": traceback.format_exc()[-600:]}
            torch.cuda.synchronize()
        torch.cuda.empty_cache()
    if 5 in only and world >= 4 and args.cube_log2n < 24:
        # the size BASELINE.json configs[4] names (2^24 nodes, ~5.5e8 edges): fits from 4 GPUs up (one GPU holds 2^22)
        try:
            blocks["5_egnn_2p24"] = bench_cube(D, args, log2n=24, self_check=False)
        except Exception as ex:  # noqa: BLE001
            import traceback
            blocks["5_egnn_2p24"] = {"error": f"{type(ex).__name__}: {str(ex)[:300]}", "trace": traceback.format_exc()[-600:]}
            torch.cuda.synchronize()
        torch.cuda.empty_cache()
    if rank == 0:
        line["configs"] = blocks
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_config2(D, args):
    import torch
    import torch.distributed as dist
    import gmp_b200
    from gmp_b200 import _lib
    rank, world, dev, local = D.rank, D.world, D.dev, D.local
    barrier = D.barrier

    # bf16 mode: node-side library GEMMs on TF32 tensor cores (fp32-strict keeps full fp32 everywhere)
    gmp_b200.set_fast_matmul(args.precision == "bf16")
    torch.manual_seed(0)
    model = make_model("schnet", args.precision).to(dev)
    params = [p for p in model.parameters()]
    atoms_h, pos_h, batch_h = synth(CFG["molecules"], seed=rank)  # graph-sharded: every rank owns its molecules
    atoms, pos, batch = atoms_h.to(dev), pos_h.to(dev), batch_h.to(dev)
    ei = gmp_b200.radius_graph(pos, CFG["cutoff"], batch, max_num_neighbors=CFG["max_num_neighbors"])
    E, N = ei.shape[1], pos.shape[0]
    # num_graphs: the field PyG's Batch carries; with it the pooling read-out needs no batch.max() read-back (a host sync)
    b = Bag(atoms=atoms, pos=pos, edge_index=ei, batch=batch, num_graphs=CFG["molecules"])

    def allreduce_grads(grads=None):
        # sums written back in place into the gradient tensors (what an optimizer consumes); no-op on one rank
        gmp_b200.allreduce_gradients(params, grads=grads)

    def step_eager(batch_obj):
        for p in params:
            p.grad = None
        out = model(batch_obj)
        out.sum().backward()
        allreduce_grads()
        return out

    # The step is ~270 launches for ~8 ms of device work and the Python / ATen / ctypes launch path needs about as long
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
        out = gs.replay()          # the replay leaves the gradients in gs.grads (static buffers of the graph's pool)
        allreduce_grads(gs.grads)
        return out

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
    graph_nodes = gs.kernel_nodes() if gs is not None else None

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

    # ---- graph construction (not part of the step: the reference receives edge_index ready-made): bit-exact radius_graph + both CSR sorts
    def build_graph():
        e = gmp_b200.radius_graph(pos, CFG["cutoff"], batch, max_num_neighbors=CFG["max_num_neighbors"])
        g_ = gmp_b200.Graph(e, N)
        return g_.by_dst.rowptr, g_.by_src.rowptr
    build_graph()
    torch.cuda.synchronize()
    s4, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s4.record()
    for _ in range(3):
        build_graph()
    e4.record()
    torch.cuda.synchronize()
    ms_graph = s4.elapsed_time(e4) / 3

    # ---- roofline of the dominant kernels, each timed alone with CUDA events on the launch stream ---------------
    roof = dominant_kernel_roofline(model, b, E, N, dev, args) if rank == 0 else None

    ms, ms_e2e = D.reduce([ms, ms_e2e])
    E_total = D.reduce([float(E)], "sum")[0]
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
        gmp_b200.set_fast_matmul(True)
    del gs, gs2, uploader
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = run_cpu(steps=10, warmup=1, molecules=512, budget_s=20.0)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        L = CFG["layers"]
        line = dict(metric=METRIC, value=E_total * L / (ms * 1e-3), unit=UNIT, n_gpus=world, steps=K, warmup=W,
                    ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32" if args.precision == "fp32" else "bf16", data="synthetic",
                    config={"workload": WORKLOAD, "nodes_per_gpu": N, "edges_per_gpu": E, "layers": L,
                            "precision": args.precision,
                            "tolerance_vs_fp32_reference": 1e-2 if args.precision == "bf16" else 1e-5,
                            "parity": "tests/test_gpu_config_parity.py::test_config2_* : 64 molecules of this batch, whole model, vs the oracle",
                            "node_side_gemms": ("tcgen05 chain kernel (node_chain_kernel: lin2 -> ssp -> lin -> + h -> next lin1 per launch, and the transposed chain "
                                                "backward); dW, db on tcgen05 (linear_wgrad_tc_kernel); no cuBLAS / ATen elementwise inside the interaction blocks")
                            if args.precision == "bf16" else "cuBLAS fp32", "parallelism": f"graph-sharded x{world}",
                            "l2": "per-step working set (x1/agg/g rows of 6 layers, >2 GB) exceeds the 126 MB L2",
                            "cuda_graph": (graph_nodes is not None) if graph_note is None else graph_note,
                            "e2e_path": "pinned host batch -> staging buffers (upload of batch i+1 under step i) -> graph inputs -> one CUDA graph "
                                        "(CSR sort + forward + backward) -> host result" if graph_nodes is not None else "host batch -> model(batch) eagerly -> host result"},
                    roofline=roof, cpu_baseline=cpu, fp32_strict=strict,
                    e2e={"value": E_total * L / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                         "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e2e},
                    graph_build={"ms": ms_graph, "edges_per_s": float(E) / (ms_graph * 1e-3),
                                 "what": "gmp_b200.radius_graph (torch_cluster order, bit-exact) + dst- and src-sorted CSR, per GPU, outside the timed step"},
                    gpu_launches=int(launches), graph_kernel_nodes=graph_nodes, clocks=clk)
    return line


def dominant_kernel_roofline(model, b, E, N, dev, args):
    """Time the two dominant kernels of the step alone -- the fused CFConv forward (training variant) and the fused
    filter-side backward -- with CUDA events around each launch on torch's current stream, the L2 flushed in between,
    and report the WORSE fraction.  Algorithmic bytes per launch:
      forward : E*(4F + 8) [x1 row gather + col + distance] + N*(4F + 4) [agg row + rowptr]
      backward: E*(4F + 8) + N*(8F + 4) [x1 gather; g row read once per destination row ...]   (2x fwd bytes for
                the whole backward, i.e. this kernel + the dL/dx1 reduction; each is held to the forward's figure)"""
    import ctypes as C
    import torch
    import gmp_b200
    from gmp_b200._lib import SchnetFilter, call, ptr
    pk = peaks()
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
    gout = torch.randn(N, F, device=dev)
    agg = torch.empty(N, F, device=dev)
    prec = 0 if args.precision == "fp32" else 1
    lib = gmp_b200._lib.lib()
    x1b = x1.to(torch.bfloat16)
    head = torch.empty(lib.gmp_schnet_tc2_num_chunks(E), F, device=dev)
    rowid = csr.row_ids()
    keep = torch.empty(E, F, dtype=torch.bfloat16, device=dev)   # training keeps the per-edge filter values for the backward,
    keep_row = g.by_src.inv_perm()                               # in the order its gather-multiply-reduce reads them
    nparts, plen = lib.gmp_schnet_bwd_num_parts(E), lib.gmp_schnet_bwd_part_len(CFG["gaussians"], F)
    parts = torch.empty(nparts, plen, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def fwd():
        if prec == 1:   # the entry point the bf16 model path calls (zeroes agg, runs the pipelined kernel and the boundary fix-up)
            call("gmp_schnet_cfconv_fwd_tc2_keep", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(rowid), N, E, ptr(ew), ptr(x1b),
                 C.byref(filt), ptr(agg), ptr(head), ptr(keep), ptr(keep_row))
        else:
            call("gmp_schnet_cfconv_fwd", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, N, E, ptr(ew), None, ptr(x1),
                 C.byref(filt), ptr(agg), prec)

    def bwd():
        if prec == 1:
            call("gmp_schnet_cfconv_bwd_tc2", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, ptr(rowid), N, E, ptr(ew), ptr(x1b),
                 C.byref(filt), ptr(gout), ptr(parts), nparts)
        else:
            call("gmp_schnet_cfconv_bwd", ptr(csr.rowptr), ptr(csr.col), csr.perm_ptr, N, E, ptr(ew), None, ptr(x1),
                 C.byref(filt), ptr(gout), ptr(parts), None, None, prec)

    def time_alone(fn):
        times = []
        for it in range(3 + 10):
            flush.zero_()  # write 256 MB > L2 between launches
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            if it >= 3:
                times.append(s.elapsed_time(e))
        return sum(times) / len(times)

    ms_f, ms_b = time_alone(fwd), time_alone(bwd)
    alg_f = E * (4 * F + 8) + N * (4 * F + 4)
    alg_b = E * (4 * F + 8) + N * (4 * F + 4)      # x1 row gather per edge + g row per destination row (read once)
    fr_f, fr_b = alg_f / (ms_f * 1e-3) / 1e9 / pk["hbm"], alg_b / (ms_b * 1e-3) / 1e9 / pk["hbm"]
    flops = E * (2 * 64 * F + 2 * F * F + 2 * F)  # padded-G GEMM1 + GEMM2 + message product
    worse_is_bwd = fr_b < fr_f
    ms, alg = (ms_b, alg_b) if worse_is_bwd else (ms_f, alg_f)
    names = (("schnet_fwd_tc2_kernel (tcgen05, bf16, pipelined; training variant that also stores the per-edge filter values; time includes "
              "the agg memset and the boundary fix-up) via gmp_schnet_cfconv_fwd_tc2_keep", "schnet_bwd_tc2_kernel via gmp_schnet_cfconv_bwd_tc2")
             if prec == 1 else ("schnet_fwd_kernel<128> (fp32 FFMA)", "schnet_bwd_kernel<128> (fp32 FFMA)"))
    traffic = _ncu_traffic("schnet_bwd_tc2_kernel" if worse_is_bwd else "schnet_fwd_tc2_kernel") if prec == 1 else None
    return {"bound": "hbm", "kernel": names[1] if worse_is_bwd else names[0], "achieved": alg / (ms * 1e-3) / 1e9, "peak": pk["hbm"],
            "unit": "GB/s", "frac": min(fr_f, fr_b), "traffic": traffic, "peak_source": pk["source"] + " hbm_gbs", "ms_per_launch": ms,
            "algorithmic_bytes": alg,
            "both": {"forward": {"kernel": names[0], "ms_per_launch": ms_f, "frac": fr_f, "algorithmic_bytes": alg_f,
                                 "tflops_filter_mlp": flops / (ms_f * 1e-3) / 1e12},
                     "backward": {"kernel": names[1], "ms_per_launch": ms_b, "frac": fr_b, "algorithmic_bytes": alg_b,
                                  "tflops_filter_mlp": 3 * flops / (ms_b * 1e-3) / 1e12}},
            "note": "the reported fraction is the worse of the two kernels; algorithmic bytes are SURVEY 8d's fp32 figure -- the kernels gather "
                    "x1 as bf16 rows (256 B/edge, mostly L2 hits) and the training forward also writes 256 B/edge of filter values, "
                    "which are not algorithmic bytes; `traffic` = dram read + write bytes per launch from profiles/r02_ncu_traffic.json "
                    "(ncu --set full of this same build; null when that file is absent)"}


def _ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture of the current build."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(kernel, {}).get("dram_bytes")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# configs 3 / 4: TFN / MACE on batched point clouds, graph-sharded (strong scaling: the cloud count is fixed)
# ------------------------------------------------------------------------------------------------
TP_FLOPS_FWD = {   # SURVEY.md 8d, FLOPs per edge forward: first layer (in = C x 0e), later layers
    "tfn": (10.6e6, 35.99e6),
    "mace": (25.4e6, 93.43e6),
}


def bench_clouds(D, which: str, args):
    import torch
    import torch.distributed as dist
    import gmp_b200
    rank, world, dev = D.rank, D.world, D.dev
    spec = CLOUDS[which]
    clouds, layers = spec["clouds"], spec["layers"]
    assert clouds % world == 0
    mine = clouds // world
    atoms_all, pos_all, _ = synth_clouds(clouds, 0)
    pos = pos_all[rank * mine * 64:(rank + 1) * mine * 64].contiguous().to(dev)
    batch = torch.arange(mine).repeat_interleave(64).to(dev)
    ei = gmp_b200.radius_graph(pos, 2.0, batch, max_num_neighbors=64)
    atoms = torch.zeros(pos.shape[0], dtype=torch.long, device=dev)
    gmp_b200.set_fast_matmul(args.precision == "bf16")
    torch.manual_seed(0)
    model = make_model(which, args.precision).to(dev)
    if world > 1:   # e3nn BatchNorm statistics over the whole (sharded) batch: same numbers as one process
        for m in model.modules():
            if isinstance(m, gmp_b200.tfn.BatchNorm):
                m.process_group = dist.group.WORLD
    params = list(model.parameters())
    b = Bag(atoms=atoms, pos=pos, edge_index=ei, batch=batch, num_graphs=mine)

    def step():
        for p in params:
            p.grad = None
        model(b).sum().backward()
        gmp_b200.allreduce_gradients(params)

    K = max(2, min(args.steps, 3))
    ms = D.timed(step, 2, K)
    E_tot, N_tot = D.reduce([float(ei.shape[1]), float(pos.shape[0])], "sum")
    pk = peaks()
    f0, f1 = TP_FLOPS_FWD[which]
    flops = 3.0 * E_tot * (f0 + (layers - 1) * f1)
    if which == "mace":
        flops += 3.0 * N_tot * layers * 2.18e6          # product block (symmetric contraction + o3.Linear)
    tf = flops / (ms * 1e-3) / 1e12
    out = {"workload": f"{which.upper()} model, BASELINE.json configs[{2 if which == 'tfn' else 3}]: {clouds} clouds x 64 points, {layers} layers, "
                       f"graph-sharded x{world}", "n_gpus": world, "scaling": "strong",
           "precision": "fp32-strict (1e-5)" if args.precision == "fp32" else "bf16 tcgen05 (1e-2 per layer)",
           "nodes": int(N_tot), "edges": int(E_tot), "steps": K, "ms_per_step": ms,
           "edges_per_s_per_layer": E_tot * layers / (ms * 1e-3),
           "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["tc_sustained"] * world, "unit": "TFLOP/s", "frac": tf / (pk["tc_sustained"] * world),
                        "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside a long step)" + (f" x {world} GPUs" if world > 1 else ""),
                        "algorithmic_flops": flops, "traffic": None,
                        "note": "whole step (all layers, forward + backward = 3x SURVEY 8d's forward FLOPs per edge; includes the node-side work and the gradient all-reduce)"}}
    # one hidden layer (layer >= 1: the shape SURVEY 8d's targets are quoted on) forward + backward alone, on rank 0's shard
    try:
        conv = model.convs[1]
        esh, eft = gmp_b200.edge_geometry(pos, ei, 2, model.radial_embedding)
        x = torch.randn(pos.shape[0], conv.in_irreps.dim, device=dev, requires_grad=True)
        cparams = list(conv.parameters())

        def layer_step():
            for p in cparams:
                p.grad = None
            x.grad = None
            conv(x, ei, esh, eft).sum().backward()

        for _ in range(2):
            layer_step()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(K):
            layer_step()
        e.record()
        torch.cuda.synchronize()
        lms = s.elapsed_time(e) / K
        ltf = 3.0 * ei.shape[1] * f1 / (lms * 1e-3) / 1e12
        out["hidden_layer"] = {"ms_fwd_bwd": lms, "edges": int(ei.shape[1]), "edges_per_s": ei.shape[1] / (lms * 1e-3), "gpus": 1,
                               "roofline": {"bound": "tensor", "achieved": ltf, "peak": pk["tc_sustained"], "unit": "TFLOP/s",
                                            "frac": ltf / pk["tc_sustained"], "algorithmic_flops": 3.0 * ei.shape[1] * f1, "traffic": None}}
    except Exception as ex:  # noqa: BLE001
        out["hidden_layer"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
    return out if rank == 0 else None


# ------------------------------------------------------------------------------------------------
# config 5: EGNN on one large radius graph, destination-partitioned
# ------------------------------------------------------------------------------------------------
def bench_cube(D, args, log2n: int = None, layers: int = 4, self_check: bool = True):
    import torch
    import gmp_b200
    rank, world, dev = D.rank, D.world, D.dev
    log2n = log2n or args.cube_log2n
    gmp_b200.set_fast_matmul(args.precision == "bf16")
    pos = synth_cube(log2n).to(dev)
    n = pos.shape[0]
    part = gmp_b200.slab_partition(pos[:, 0], 1.0, rank, world)
    t0 = time.perf_counter()
    ei = gmp_b200.distributed.local_radius_graph(pos[part.local_global].contiguous(), 1.0, part)
    torch.cuda.synchronize()
    t_graph = time.perf_counter() - t0
    torch.manual_seed(0)
    # halo rows: pulled out of the owners' symmetric-memory buffers by our own kernel (P2P loads over NVLink, csrc/halo.cu);
    # if symmetric memory cannot be set up on this box the grouped ncclSend / ncclRecv exchange runs instead (reported)
    halo, halo_note = (args.halo if world > 1 else "nccl"), None
    if halo == "fused" and args.precision != "bf16":
        halo = "peer"
    model = gmp_b200.PartitionedEGNN(num_layers=layers, emb_dim=128, precision=args.precision, halo=halo).to(dev)
    params = list(model.parameters())
    h_own = torch.randn(part.n_own, 128, device=dev)
    p_own = pos[part.own_lo:part.own_hi].clone()
    del pos

    def step():
        for p in params:
            p.grad = None
        ho, po = model(h_own, p_own, ei, part)
        (ho.sum() + po.sum()).backward()
        gmp_b200.allreduce_gradients(params)

    while halo != "nccl":       # fused -> peer -> nccl: every rank takes the same fallback
        ok = 1.0
        try:
            step()
            torch.cuda.synchronize()
        except Exception as ex:  # noqa: BLE001 -- e.g. no P2P mapping between the devices, slabs thinner than the radius
            ok, halo_note = 0.0, f"{halo}: {type(ex).__name__}: {str(ex)[:160]}"
        if D.reduce([-ok], "max")[0] == -1.0:
            break
        halo = "peer" if halo == "fused" else "nccl"
        model.halo, model._peer, model._peer_q = halo, None, None
    check = partition_self_check(D, halo=halo) if (world > 1 and self_check) else None
    K = max(2, min(args.steps, 3))
    ms = D.timed(step, 2, K)
    E_tot = D.reduce([float(ei.shape[1])], "sum")[0]
    halo_rows = D.reduce([float(part.n_left + part.n_right)], "max")[0]
    pk = peaks()
    flops = 3.0 * layers * (131584.0 * E_tot + 98304.0 * n)
    byts = 3.0 * layers * (528.0 * E_tot + 1052.0 * n)
    tf, gbs = flops / (ms * 1e-3) / 1e12, byts / (ms * 1e-3) / 1e9
    f_tc, f_hbm = tf / (pk["tc_sustained"] * world), gbs / (pk["hbm"] * world)
    size_note = ("the full BASELINE.json configs[4] size; needs >= 4 GPUs, so it has no N = 1 point" if log2n >= 24 else
                 f"BASELINE.json configs[4] geometry; 2^24 nodes do not fit one GPU's 180 GB for a training step, so the strong-scaling series runs 2^{log2n} at every N")
    out = {"workload": f"EGNN {layers} layers d=128 on one radius graph N=2^{log2n}, r=1, density 8 ({size_note}), destination-partitioned x{world}"
                       + (", halo exchange per layer" if world > 1 else ""),
           "n_gpus": world, "scaling": "strong", "precision": "fp32-strict (1e-5)" if args.precision == "fp32" else "bf16 tcgen05 (1e-2; ReLU gradients: see tests)",
           "nodes": n, "edges": int(E_tot), "halo_nodes_max": int(halo_rows),
           "halo_exchange": None if world == 1 else (
               {"fused": "features: none -- the edge kernels' gather reads halo sources' projected rows out of the neighbours' symmetric-memory "
                         "buffers (P2P loads over NVLink inside egnn_fwd_tc2 / egnn_bwd_tc); positions and returning gradients: pull kernel (csrc/halo.cu)",
                "peer": "peer-memory pull kernel (symmetric memory, P2P loads over NVLink; csrc/halo.cu)",
                "nccl": "grouped ncclSend / ncclRecv"}[halo] + (f" (fell back: {halo_note})" if halo_note else "")),
           "steps": K, "ms_per_step": ms,
           "edges_per_s_per_layer": E_tot * layers / (ms * 1e-3), "graph_build_s": t_graph,
           "roofline": {"bound": "tensor" if f_tc >= f_hbm else "hbm", "achieved": tf if f_tc >= f_hbm else gbs,
                        "peak": (pk["tc_sustained"] if f_tc >= f_hbm else pk["hbm"]) * world, "unit": "TFLOP/s" if f_tc >= f_hbm else "GB/s",
                        "frac": max(f_tc, f_hbm), "frac_tensor": f_tc, "frac_hbm": f_hbm, "algorithmic_flops": flops, "algorithmic_bytes": byts,
                        "traffic": None, "peak_source": pk["source"]},
           "partition_check": check}
    return out if rank == 0 else None


def partition_self_check(D, log2n: int = 15, halo: str = "nccl"):
    """N >= 2 only: the destination-partitioned 2-layer EGNN (fp32-strict) against the same model on the whole graph on one
    GPU -- owned rows of h and pos and the parameter gradients must agree to fp32 round-off (the local edge order equals
    the global one, so the sums are the same sums)."""
    import torch
    import torch.distributed as dist
    import gmp_b200
    rank, world, dev = D.rank, D.world, D.dev
    was = torch.backends.cuda.matmul.allow_tf32
    gmp_b200.set_fast_matmul(False)
    try:
        pos = synth_cube(log2n, seed=1).to(dev)
        n = pos.shape[0]
        torch.manual_seed(3)
        halo = "peer" if halo == "fused" else halo     # the fp32 check has no bf16 rows to read across: it validates the partition + pull kernels
        model = gmp_b200.PartitionedEGNN(num_layers=2, emb_dim=128, precision="fp32", halo=halo).to(dev)
        g = torch.Generator().manual_seed(5)
        h = torch.randn(n, 128, generator=g).to(dev)
        ch, cp = torch.randn(n, 128, generator=g).to(dev), torch.randn(n, 3, generator=g).to(dev)
        params = list(model.parameters())
        # whole graph, one GPU (every rank computes it: 32k nodes)
        whole = gmp_b200.slab_partition(pos[:, 0], 1.0, 0, 1)
        ei_w = gmp_b200.distributed.local_radius_graph(pos, 1.0, whole)
        ho, po = model(h, pos, ei_w, whole)
        ((ho * ch).sum() + (po * cp).sum()).backward()
        ref = [p.grad.clone() for p in params]
        for p in params:
            p.grad = None
        part = gmp_b200.slab_partition(pos[:, 0], 1.0, rank, world)
        ei = gmp_b200.distributed.local_radius_graph(pos[part.local_global].contiguous(), 1.0, part)
        own = slice(part.own_lo, part.own_hi)
        h2, p2 = model(h[own].clone(), pos[own].clone(), ei, part)
        ((h2 * ch[own]).sum() + (p2 * cp[own]).sum()).backward()
        gmp_b200.allreduce_gradients(params)
        rel = lambda a, b_: ((a - b_).abs().max() / b_.abs().max().clamp_min(1e-30)).item()
        e_out = max(rel(h2, ho[own]), rel(p2 - pos[own], po[own] - pos[own]))
        e_grad = max(rel(p.grad, r) for p, r in zip(params, ref))
        e_out, e_grad = D.reduce([e_out, e_grad])
        for p in params:
            p.grad = None
        return {"nodes": n, "layers": 2, "precision": "fp32", "halo": halo, "max_rel_err_outputs": e_out, "max_rel_err_param_grads": e_grad,
                "ok": bool(e_out <= 1e-5 and e_grad <= 1e-4)}
    finally:
        gmp_b200.set_fast_matmul(was)


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
    ap.add_argument("--only", default="", help="comma-separated config numbers to run (2 = headline SchNet, 3 = TFN, 4 = MACE, 5 = EGNN large graph)")
    ap.add_argument("--cube-log2n", type=int, default=CUBE_LOG2N, help="config 5: log2 of the node count of the radius graph")
    ap.add_argument("--halo", default="peer", choices=["fused", "peer", "nccl"],
                    help="config 5 at N > 1: halo rows through the peer-memory pull kernel (default: fastest), read inside the gather from the "
                         "neighbours' memory (fused: every halo row then crosses NVLink once per incident edge, ~16x, instead of once: "
                         "584.8 vs 568.5 ms at 2 GPUs, profiles/r02_summary.md), or NCCL send/recv")
    ap.add_argument("--cpu-molecules", type=int, default=0, help="--impl reference: molecules of the bench batch to time (default: all 4096)")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_gpu(a)
