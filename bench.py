#!/usr/bin/env python
"""bench.py -- KG edges aggregated / s (forward + backward) of the hetero-GNN convolution path.

    python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path, host cores)

A "step" = one full-graph pass of the hot path: L x (HeteroConv -> ReLU) -> Linear(h,1) -> ReLU ->
LDSC-weighted MSE over every SNP (kgwas/model.py:62-86, kgwas/kgwas.py:139-151), forward +
backward + Adam step, on the synthetic KG ``kgwas-synth-v1`` with features
already projected to ``hidden``.  One unit = one typed directed edge
processed by one conv layer, forward and backward together: a step is L * sum_r E_r units.
Prints ONE JSON line (see the task contract / DESIGN.md section "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=30)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--backbone", default="SAGE", choices=["SAGE", "GAT"])
    p.add_argument("--hidden", type=int, default=128)
    p.add_argument("--layers", type=int, default=2)
    p.add_argument("--scale", type=float, default=1.0, help="shrink kgwas-synth-v1 (tests only)")
    p.add_argument("--cpu-scale", type=float, default=0.1, help="graph scale of the bounded CPU sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cuda-graph", action="store_true", help="drive every step eagerly from Python")
    p.add_argument("--profile-range", action="store_true",
                   help="wrap the timed steps in cudaProfilerStart/Stop (ncu --profile-from-start off)")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------

def layer_bytes(edge_sizes, num_nodes, h):
    """compulsory HBM bytes of ONE layer, forward + backward, fp32 / int32."""
    b = 0
    for (s, _, t), e in edge_sizes.items():
        b += 8 * e + 4 * (num_nodes[s] + num_nodes[t] + 2) + 4 * h * (min(e, num_nodes[s]) + min(e, num_nodes[t]))
    return b + 16 * h * sum(num_nodes.values())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class LaunchProfiler:
    """CUDA-event timer around every launch of the dominant kernel (kgb_spmm), on the launching
    stream, with that launch's algorithmic bytes (col idx + rowptr + unique gathered rows
    + output rows [+ output read when accumulating])."""

    def __init__(self):
        self.items = []
        self.cur = None

    def begin(self, name, csr, h, beta):
        e, n_rows, n_cols = csr.n_edges, csr.n_rows, csr.n_cols
        if name == "sddmm":
            by = 4 * e + 4 * (n_rows + 1) + 4 * h * min(e, n_cols) + 4 * h * n_rows + 4 * e
        else:
            by = 4 * e + 4 * (n_rows + 1) + 4 * h * min(e, n_cols) + 4 * h * n_rows * (2 if beta != 0 else 1)
        ev0 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        self.cur = (name, by, ev0)

    def end(self):
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        name, by, ev0 = self.cur
        self.items.append((name, by, ev0, ev1))

    def summary(self, name="spmm"):
        tot_b = tot_ms = 0.0
        n = 0
        for nm, by, e0, e1 in self.items:
            if nm == name:
                tot_b += by
                tot_ms += e0.elapsed_time(e1)
                n += 1
        return n, tot_b, tot_ms


def make_problem(args, scale, device, seed=42):
    from kgwas_b200 import make_synth_kg
    data = make_synth_kg(scale=scale, seed=seed, hidden=args.hidden)
    g = torch.Generator().manual_seed(seed + 1)
    n_snp = data["SNP"].x.size(0)
    y = torch.rand(n_snp, generator=g) * 4.0                       # chi^2-like labels
    w = (0.5 + torch.rand(n_snp, generator=g, dtype=torch.float64))  # LDSC weights are float64 (kgwas.py:143)
    w = w / w.mean()
    return data, y, w


def edge_stats(data):
    sizes = {et: int(ei.size(1)) for et, ei in data.edge_index_dict.items()}
    nodes = {t: int(x.size(0)) for t, x in data.x_dict.items()}
    return sizes, nodes


# ------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port), also the cpu_baseline leg
# ------------------------------------------------------------------------------------------------

def run_cpu(args, steps, warmup, scale):
    from oracle import kgwas_oracle as O
    torch.set_num_threads(os.cpu_count())
    data, y, w = make_problem(args, scale, "cpu")
    sizes, nodes = edge_stats(data)
    h, L = args.hidden, args.layers
    torch.manual_seed(0)
    model = O.HeteroGNN(data, h, 1, L, args.backbone, "sum", h, h, h, 1)
    opt = None
    ei = data.edge_index_dict
    n_snp = nodes["SNP"]
    times = []
    for it in range(warmup + steps):
        x = {k: v.clone().requires_grad_() for k, v in data.x_dict.items()}
        t0 = time.perf_counter()
        if opt is not None:
            opt.zero_grad()
        xd = O.conv_stack_forward(model.convs, x, ei)
        pred = model.lin(xd["SNP"]).relu().reshape(-1)[:n_snp]
        loss = O.weighted_mse(pred, y, w)
        loss.backward()
        if opt is None:   # lazy weights exist only after the first forward (kgwas.py:116 relies on in-place materialise)
            opt = torch.optim.Adam([p for p in model.parameters() if not isinstance(p, torch.nn.parameter.UninitializedParameter)],
                                   lr=1e-4, weight_decay=5e-4)
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    edges_step = L * sum(sizes.values())
    t = sum(times) / len(times)
    return {"value": edges_step / t, "ms_per_step": t * 1e3, "edges_step": edges_step, "cores": os.cpu_count(),
            "sample": f"kgwas-synth-v1 at scale {scale} ({sum(sizes.values())} typed edges, {sum(nodes.values())} nodes), "
                      f"{L}-layer {args.backbone} h={h}, full-graph fwd+bwd+Adam, {steps} timed step(s) after {warmup} warm-up"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu(args, max(1, min(args.steps, 3)), max(1, min(args.warmup, 1)), args.cpu_scale)
    line = {
        "impl": "reference", "metric": "kg_edges_aggregated_per_s_fwd_bwd", "value": r["value"], "unit": "edges/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": r["value"], "unit": "edges/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "torch_geometric is not installable offline (un-vendored, un-pinned dependency of the reference), so the "
                "reference's device='cpu' path is its pure-PyTorch restatement oracle/kgwas_oracle.py on all host cores",
    }
    print(json.dumps(line))


def workload_config(args, n_gpus):
    return {"workload": f"kgwas-synth-v1 fast-mode KG (784256 SNP / 20371 Gene / 23211 GO nodes, 27 edge types, ~21.4M typed "
                        f"edges), {args.layers}-layer hetero-{args.backbone} hidden={args.hidden}, full-graph fwd+bwd+Adam",
            "baseline_config": "configs[1]" if args.backbone == "SAGE" else "configs[2]-like (GAT on the fast-mode KG)",
            "hidden": args.hidden, "layers": args.layers, "backbone": args.backbone, "graph_scale": args.scale,
            "parallelism": f"snp-shard x{n_gpus}" if n_gpus > 1 else "single",
            "l2_policy": "inputs larger than L2 (node features + CSR > 1 GB vs 126 MB L2)"}


# ------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------

def ours(args):
    import kgwas_b200
    from kgwas_b200 import _lib
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world > 1:
        from kgwas_b200 import dist as kdist
        return kdist.bench_sharded(args, rank, world, dev, {"ClockSampler": ClockSampler, "workload_config": workload_config})

    h, L = args.hidden, args.layers
    data, y, w = make_problem(args, args.scale, dev)
    sizes, nodes = edge_stats(data)
    edges_layer = sum(sizes.values())
    edges_step = L * edges_layer
    n_snp = nodes["SNP"]
    torch.manual_seed(0)
    model = kgwas_b200.HeteroGNN(data, h, 1, L, args.backbone, "sum", h, h, h, 1).to(dev)
    gdata = data.to(dev)
    ei = gdata.edge_index_dict
    x_dev = {k: v.clone().requires_grad_() for k, v in gdata.x_dict.items()}
    y_d, w_d = y.to(dev), w.to(dev)
    x_host = {k: v.pin_memory() for k, v in data.x_dict.items()}
    opt = None

    def step(x):
        nonlocal opt
        if opt is not None:
            opt.zero_grad(set_to_none=True)
        for v in x.values():
            v.grad = None           # the features stand for the MLP outputs: their gradient is produced, not accumulated
        pred = model.forward_from_hidden(x, ei, n_snp).reshape(-1)
        loss = torch.mean(w_d * (pred - y_d) ** 2)                       # kgwas.py:145
        loss.backward()
        if opt is None:
            # kgwas.py:116 Adam(lr, weight_decay); the fused implementation is the one that is both a single kernel per
            # step and legal inside a CUDA graph (the default foreach path is not capturable, and capturable=True
            # without fused=True falls back to ~170 per-tensor kernels for the step counters)
            opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=5e-4, fused=True, capturable=True)
        opt.step()
        return pred, loss

    for _ in range(max(args.warmup, 3)):
        step(x_dev)
    torch.cuda.synchronize()

    # The full-graph step has fixed shapes: capture it once in a CUDA graph (kgwas_b200.graphed) and replay it --
    # the eager step is bound by the host (~170 launches issued from Python take longer than the kernels run).
    graphed, graph_note = None, "off (--no-cuda-graph)"
    if not args.no_cuda_graph:
        try:
            from kgwas_b200.graphed import GraphedStep
            graphed = GraphedStep(step, x_dev, warmup=3)
            graph_note = "whole step (fwd + bwd + Adam, both scheduler streams) captured once, replayed per step"
        except Exception as e:                               # noqa: BLE001 -- report and fall back to the eager step
            graphed, graph_note = None, f"capture failed, eager steps: {type(e).__name__}: {e}"[:300]
            torch.cuda.synchronize()
    run_step = (lambda x: graphed(x)) if graphed is not None else step
    for _ in range(3):
        run_step(x_dev)
    torch.cuda.synchronize()

    # ---- device-resident throughput ("value")
    from kgwas_b200 import ops as _ops
    clocks = ClockSampler(local_rank)
    clocks.start()
    k0 = _lib.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if args.profile_range:
        torch.cuda.profiler.start()
    ev0.record()
    for _ in range(args.steps):
        run_step(x_dev)
    ev1.record()
    torch.cuda.synchronize()
    if args.profile_range:
        torch.cuda.profiler.stop()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = _lib.kernel_launch_count() - k0
    if graphed is not None:                                  # replays do not pass through the library's launch counter
        launches = graphed.kernels_per_replay * args.steps
    clk = clocks.stop()

    # ---- dominant-kernel roofline: CUDA events around every kgb_spmm launch, on the launching stream, in a second
    # timed region with the side stream switched off (concurrent kernels would share the machine and the per-launch
    # durations would not be attributable)
    prof = LaunchProfiler()
    _ops.MULTI_STREAM = False
    step(x_dev)
    torch.cuda.synchronize()
    _lib._prof = prof
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_steps = max(3, min(10, args.steps))
    pe0.record()
    for _ in range(prof_steps):
        step(x_dev)
    pe1.record()
    torch.cuda.synchronize()
    _lib._prof = None
    ms_serial = pe0.elapsed_time(pe1) / prof_steps
    serial_note = "eager, host-bound"
    if graphed is not None:
        # the eager single-stream step is bound by the host; the denominator of the kernel's share of the step is the
        # same single-stream step replayed from a CUDA graph (pure device time, kernels back to back)
        try:
            g1 = GraphedStep(step, x_dev, warmup=1)
            torch.cuda.synchronize()
            pe0.record()
            for _ in range(prof_steps):
                g1()
            pe1.record()
            torch.cuda.synchronize()
            ms_serial, serial_note = pe0.elapsed_time(pe1) / prof_steps, "CUDA-graph replay, device-bound"
            del g1
        except Exception:                                     # noqa: BLE001
            torch.cuda.synchronize()
    _ops.MULTI_STREAM = True
    n_spmm, spmm_bytes, spmm_ms = prof.summary("spmm")

    # ---- end to end: host (pinned) features -> H2D -> step -> D2H of logits + loss, every step
    e2e = None
    if not args.no_e2e:
        h2d = sum(v.numel() * 4 for v in x_host.values())
        out_host = torch.empty(n_snp, dtype=torch.float32).pin_memory()

        # Double-buffered input pipeline (what a DataLoader with pin_memory + non_blocking copies does): the H2D copy
        # of step i+1's features runs on a copy stream while step i computes.  Every step's inputs still cross PCIe
        # inside the timed region, and every step ends with the D2H read of its logits and loss.
        copy_stream = torch.cuda.Stream(dev)
        main_stream = torch.cuda.current_stream(dev)
        bufs = [{k: torch.empty_like(v, device=dev) for k, v in x_host.items()} for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            b = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])          # the step that last used this buffer is done
                for k, v in x_host.items():
                    bufs[b][k].copy_(v, non_blocking=True)
                ready[b].record(copy_stream)

        def e2e_step(i, last):
            b = i % 2
            main_stream.wait_event(ready[b])
            if not last:
                prefetch(i + 1)
            if graphed is not None:                                  # D2D into the graph's static inputs, then replay
                pred, loss = graphed(bufs[b])
            else:
                x = {k: v.detach().requires_grad_() for k, v in bufs[b].items()}
                pred, loss = step(x)
            consumed[b].record(main_stream)
            out_host.copy_(pred.detach(), non_blocking=True)
            return loss.item()                                           # D2H + sync

        for b in range(2):
            consumed[b].record(main_stream)
        prefetch(0)
        for i in range(2):
            e2e_step(i, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(2, 2 + args.steps):
            e2e_step(i, False)      # K steps, K host->device copies inside the timed region (steady-state pipeline)
        e1.record()
        torch.cuda.synchronize()
        ems = e0.elapsed_time(e1) / args.steps
        e2e = {"value": edges_step / (ems * 1e-3), "unit": "edges/s", "ms_per_step": ems, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": n_snp * 4 + 8,
               "api": "HeteroGNN.forward_from_hidden(x_dict, edge_index_dict, batch_size) + loss.backward() + Adam.step(); "
                      "node features copied from pinned host memory every step (double-buffered on a copy stream, overlapping the previous step), graph resident (data_to_cuda=True, kgwas.py:96-97)"}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    b_layer = layer_bytes(sizes, nodes, h)
    step_gbs = L * b_layer / (ms * 1e-3) / 1e9
    spmm_gbs = (spmm_bytes / (spmm_ms * 1e-3) / 1e9) if spmm_ms > 0 else 0.0

    # measured DRAM traffic of the same launches (ncu --set full, dram__bytes_read + dram__bytes_write), per launch like
    # `achieved`: written by scratch/ncu_traffic.py from the capture committed under profiles/
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_spmm_traffic.json")))
        if tj.get("hidden") == h and tj.get("backbone") == args.backbone and args.scale == 1.0:
            traffic = tj["dram_bytes_per_step"] / max(1, n_spmm // prof_steps)
            traffic_src = tj["source"]
    except Exception:
        pass

    cpu = None
    if not args.no_cpu_baseline:
        r = run_cpu(args, 2, 1, args.cpu_scale)
        cpu = {"value": r["value"], "unit": "edges/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    line = {
        "metric": "kg_edges_aggregated_per_s_fwd_bwd", "value": edges_step / (ms * 1e-3), "unit": "edges/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "edges_per_step": edges_step, "edges_per_layer": edges_layer, "num_nodes": nodes,
        "edge_counts": {"|".join(k): v for k, v in sizes.items()},
        "roofline": {"bound": "hbm", "kernel": f"lean::k_spmm_lean<{h // 128}> (segmented gather-reduce, all kgb_spmm launches of a step)",
                     "achieved": spmm_gbs, "peak": peak, "unit": "GB/s", "frac": spmm_gbs / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "launches_per_step": n_spmm // prof_steps,
                     "algorithmic_bytes_per_step": spmm_bytes / prof_steps,
                     "avg_launch_ms": spmm_ms / max(1, n_spmm), "kernel_share_of_step": spmm_ms / (ms_serial * prof_steps),
                     "measured_in": f"{prof_steps} extra eager steps, single stream, CUDA events around every kgb_spmm launch; "
                                    f"share = their sum / single-stream step ({ms_serial:.3f} ms/step, {serial_note})"},
        "roofline_step": {"formula": "SURVEY.md 8(d) B_layer(h)", "bytes_per_layer": b_layer,
                          "bytes_per_edge": b_layer / edges_layer, "achieved": step_gbs, "peak": peak,
                          "unit": "GB/s", "frac": step_gbs / peak, "frac_of_8000_spec": step_gbs / 8000.0},
        "clocks": clk, "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
        "cuda_graph": graph_note,
    }
    if e2e:
        line["e2e"] = e2e
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)
