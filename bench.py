#!/usr/bin/env python
"""bench.py -- KG edges aggregated / s (forward + backward) of the hetero-GNN convolution path.

    python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path, host cores)

A "step" = one full-graph pass of the hot path: L x (HeteroConv -> ReLU) -> Linear(h,1) -> ReLU ->
LDSC-weighted MSE over every SNP (kgwas/model.py:62-86, kgwas/kgwas.py:139-151), forward +
backward + Adam step, on the synthetic KG ``kgwas-synth-v1`` with features
already projected to ``hidden``.  One unit = one typed directed edge
processed by one conv layer, forward and backward together: a step is L * sum_r E_r units.

N = 1: BASELINE configs[1] (2-layer hetero-SAGE h=128 on the full fast-mode KG) is the headline; the same line carries
a ``gat`` block (configs[2] / configs[4] shaped: 2-layer GAT h=128, 3-layer GAT h=256 on the same KG).
N > 1: BASELINE configs[3] as written -- the ONE 784 256-SNP KG with its SNP axis split N ways (strong scaling,
headline) and, as a secondary ``weak`` block, N blocks of 784 256 SNPs against one shared gene graph.  Every N > 1 line
carries a ``parity`` block: the sharded logits and parameter gradients against the same graph run un-sharded on rank 0.

Prints ONE JSON line (see the task contract / DESIGN.md section "Measurement").
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "kg_edges_aggregated_per_s_fwd_bwd"


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
    p.add_argument("--scaling", default="both", choices=["strong", "weak", "both"],
                   help="N > 1: strong = the named KG split N ways (headline), weak = N SNP blocks; both = strong + a weak block")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-gat", action="store_true", help="N = 1: skip the GAT block (configs 3 / 5 shapes)")
    p.add_argument("--gat-steps", type=int, default=8)
    p.add_argument("--no-parity", action="store_true")
    p.add_argument("--no-cuda-graph", action="store_true", help="drive every step eagerly from Python")
    p.add_argument("--profile-range", action="store_true",
                   help="wrap the timed steps in cudaProfilerStart/Stop (ncu --profile-from-start off)")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------

def layer_bytes(edge_sizes, num_nodes, h, backbone="SAGE"):
    """compulsory HBM bytes of ONE layer, forward + backward, fp32 / int32.  GAT adds the four per-edge
    attention scalars (alpha written + read, d alpha, d u: 16 B per edge)."""
    b = 0
    for (s, _, t), e in edge_sizes.items():
        b += 8 * e + 4 * (num_nodes[s] + num_nodes[t] + 2) + 4 * h * (min(e, num_nodes[s]) + min(e, num_nodes[t]))
        if backbone == "GAT":
            b += 16 * e
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


def labels(n_snp, seed):
    g = torch.Generator().manual_seed(seed)
    y = torch.rand(n_snp, generator=g) * 4.0                            # chi^2-like labels
    w = (0.5 + torch.rand(n_snp, generator=g, dtype=torch.float64))     # LDSC weights are float64 (kgwas.py:143)
    return y, w / w.mean()


def edge_stats(data):
    sizes = {et: int(ei.size(1)) for et, ei in data.edge_index_dict.items()}
    nodes = {t: int(x.size(0)) for t, x in data.x_dict.items()}
    return sizes, nodes


def workload_config(args, n_gpus, sizes=None, nodes=None, backbone=None, hidden=None, layers=None, mode=None):
    backbone, hidden, layers = backbone or args.backbone, hidden or args.hidden, layers or args.layers
    if sizes is not None:
        go = sum(v for k, v in nodes.items() if k not in ("SNP", "Gene"))
        shape = (f"{nodes['SNP']} SNP / {nodes['Gene']} Gene / {go} GO nodes, {len(sizes)} edge types, "
                 f"{sum(sizes.values())} typed edges after ToUndirected (coalescing) + AddSelfLoops")
    else:
        shape = "784256 SNP / 20371 Gene / 23211 GO nodes, 27 edge types, 18400753 typed edges at scale 1.0"
    par = "single"
    if n_gpus > 1:
        par = (f"snp-shard x{n_gpus}, strong scaling: the one named KG, SNP rows split {n_gpus} ways" if mode != "weak"
               else f"snp-shard x{n_gpus}, weak scaling: {n_gpus} SNP blocks, shared gene/GO graph")
    base = {("SAGE", 2, 128): "configs[1]", ("GAT", 2, 128): "configs[2] shape (GAT h=128 L=2) on the fast-mode KG",
            ("GAT", 3, 256): "configs[4] shape (GAT h=256 L=3) on the fast-mode KG"}.get((backbone, layers, hidden), "custom")
    if n_gpus > 1 and backbone == "SAGE":
        base = "configs[3]"
    return {"workload": f"kgwas-synth-v1 fast-mode KG ({shape}), {layers}-layer hetero-{backbone} hidden={hidden}, "
                        f"full-graph fwd+bwd+Adam",
            "baseline_config": base, "hidden": hidden, "layers": layers, "backbone": backbone,
            "graph_scale": args.scale, "parallelism": par,
            "l2_policy": "inputs larger than L2 (node features + CSR > 1 GB vs 126 MB L2)"}


# ------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port), also the cpu_baseline leg
# ------------------------------------------------------------------------------------------------

def run_cpu(args, steps, warmup, scale, parity_dev=None):
    """``steps`` timed full-graph steps of the oracle port on all host cores after ``warmup`` untimed ones, on
    kgwas-synth-v1 at ``scale``.  ``parity_dev``: also run OUR engine on the same graph with the same weights on that
    CUDA device and report the relative error of the per-SNP logits (before the timed steps move the weights)."""
    from oracle import kgwas_oracle as O
    from kgwas_b200 import make_synth_kg
    torch.set_num_threads(os.cpu_count())
    h, L = args.hidden, args.layers
    data = make_synth_kg(scale=scale, seed=42, hidden=h)
    sizes, nodes = edge_stats(data)
    n_snp = nodes["SNP"]
    y, w = labels(n_snp, 43)
    torch.manual_seed(0)
    model = O.HeteroGNN(data, h, 1, L, args.backbone, "sum", h, h, h, 1)
    opt = None
    ei = data.edge_index_dict
    parity = None
    if parity_dev is not None:
        import kgwas_b200
        with torch.no_grad():
            xd = O.conv_stack_forward(model.convs, dict(data.x_dict), ei)      # materialises the lazy weights
            ref = model.lin(xd["SNP"]).reshape(-1)
        ours = kgwas_b200.HeteroGNN(data, h, 1, L, args.backbone, "sum", h, h, h, 1, no_relu=True)
        ours.load_state_dict(model.state_dict())
        ours = ours.to(parity_dev)
        gd = data.to(parity_dev)
        with torch.no_grad():
            got = ours.forward_from_hidden(gd.x_dict, gd.edge_index_dict, n_snp).reshape(-1).cpu()
        parity = {"logits_max_rel_err": float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)),
                  "tolerance": 1e-4, "n_logits": n_snp, "against": "oracle port, same graph and weights (pre-ReLU logits)"}
        del ours, gd
        kgwas_b200.plan.clear_plan_cache()
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
            "parity": parity, "graph_scale": scale,
            "sample": f"kgwas-synth-v1 at scale {scale} ({sum(sizes.values())} typed edges, {sum(nodes.values())} nodes), "
                      f"{L}-layer {args.backbone} h={h}, full-graph fwd+bwd+Adam, {steps} timed step(s) after {warmup} warm-up"}


def reference_arm(args):
    """The reference's own CPU implementation of the path (its pure-PyTorch restatement: torch_geometric cannot be
    installed offline) on all host cores.  Times EXACTLY --steps steps after --warmup warm-up steps; each step is a
    bounded sample of the workload (the same synthetic KG at --cpu-scale), and the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu(args, max(1, args.steps), max(0, args.warmup), args.cpu_scale)
    cfg = workload_config(args, args.gpus)
    cfg["reference_sample"] = f"each timed step runs the workload's graph at scale {args.cpu_scale}: {r['sample']}"
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "edges/s",
        "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": max(0, args.warmup), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 and args.scaling != "weak" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "sample_graph_scale": args.cpu_scale, "edges_per_step_of_sample": r["edges_step"],
        "cpu_baseline": {"value": r["value"], "unit": "edges/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "torch_geometric is not installable offline (un-vendored, un-pinned dependency of the reference), so the "
                "reference's device='cpu' path is its pure-PyTorch restatement oracle/kgwas_oracle.py on all host cores; "
                "edges/s is size-normalised (edges of the sample / time of the sample)",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# ours: one runner for N = 1, strong and weak sharding
# ------------------------------------------------------------------------------------------------

class Runner:
    """Problem + model + (captured) step for one configuration on this rank.

    mode 'single': the whole KG on one GPU.  'strong': the same KG, this rank owns SNP rows [lo, hi) and the edges that
    touch them (dist.shard_graph).  'weak': this rank owns its own block of 784 256 SNPs against the shared gene graph."""

    def __init__(self, args, dev, rank, world, mode, backbone=None, hidden=None, layers=None):
        import kgwas_b200
        from kgwas_b200 import make_synth_kg, dist as kdist
        self.args, self.dev, self.rank, self.world, self.mode = args, dev, rank, world, mode
        self.backbone, self.h, self.L = backbone or args.backbone, hidden or args.hidden, layers or args.layers
        h, L = self.h, self.L
        self.shard, self.full = None, None
        if mode == "weak":
            full = make_synth_kg(scale=args.scale, seed=42, hidden=h, snp_block=rank)
            local, self.shard, _ = kdist.shard_graph(full, rank, world, shard_snp=False)
            n_snp = local["SNP"].num_nodes
            y, w = labels(n_snp, 43 + rank)
            self.n_global = n_snp * world
        elif mode == "strong":
            full = make_synth_kg(scale=args.scale, seed=42, hidden=h)
            local, self.shard, (lo, hi) = kdist.shard_graph(full, rank, world)
            self.lo, self.hi = lo, hi
            self.n_global = full["SNP"].num_nodes
            y, w = labels(self.n_global, 43)
            self.y_full, self.w_full = y, w
            y, w = y[lo:hi].clone(), w[lo:hi].clone()
            if rank == 0 and not args.no_parity:
                self.full = full
        else:
            local = make_synth_kg(scale=args.scale, seed=42, hidden=h)
            self.n_global = local["SNP"].num_nodes
            y, w = labels(self.n_global, 43)
        self.sizes, self.nodes = edge_stats(local)
        self.n_snp = self.nodes["SNP"]
        self.edges_local = sum(self.sizes.values())
        self.y, self.w = y.to(dev), w.to(dev)
        torch.manual_seed(0)
        self.model = kgwas_b200.HeteroGNN(local, h, 1, L, self.backbone, "sum", h, h, h, 1).to(dev)
        if self.shard is not None:
            kdist.attach(self.model, self.shard)
        gdata = local.to(dev)
        self.ei = gdata.edge_index_dict
        self.x_dev = {k: v.clone().requires_grad_() for k, v in gdata.x_dict.items()}
        self.x_host = {k: v.pin_memory() for k, v in local.x_dict.items()}
        self.opt = None
        self.graphed, self.graph_note = None, "off (--no-cuda-graph)"

    # -- one optimiser step ---------------------------------------------------------------------
    def step(self, x):
        from kgwas_b200 import dist as kdist
        if self.opt is not None:
            self.opt.zero_grad(set_to_none=True)
        for v in x.values():
            v.grad = None           # the features stand for the MLP outputs: their gradient is produced, not accumulated
        pred = self.model.forward_from_hidden(x, self.ei, self.n_snp).reshape(-1)
        loss = torch.sum(self.w * (pred - self.y) ** 2) / self.n_global          # kgwas.py:145 (global mean)
        loss.backward()
        if self.world > 1:
            kdist.all_reduce_gradients([p for p in self.model.parameters()
                                        if not isinstance(p, torch.nn.parameter.UninitializedParameter)])
        if self.opt is None:
            # kgwas.py:116 Adam(lr, weight_decay); the fused implementation is the one that is both a single kernel per
            # step and legal inside a CUDA graph
            self.opt = torch.optim.Adam(self.model.parameters(), lr=1e-4, weight_decay=5e-4, fused=True, capturable=True)
        self.opt.step()
        return pred, loss

    # -- parity of the sharded run against the un-sharded one (rank 0 holds the full graph) --------------------------
    def parity_vs_unsharded(self):
        """Forward + backward (no optimiser step) of the sharded model on every rank and of an un-sharded copy with the
        same weights on rank 0: max relative error of all per-SNP logits and of the parameter gradients."""
        import torch.distributed as dist
        import kgwas_b200
        from kgwas_b200 import dist as kdist
        x = {k: v.detach().clone().requires_grad_() for k, v in self.x_dev.items()}
        self.model.zero_grad(set_to_none=True)
        pred = self.model.forward_from_hidden(x, self.ei, self.n_snp).reshape(-1)
        loss = torch.sum(self.w * (pred - self.y) ** 2) / self.n_global
        loss.backward()
        params = [(k, p) for k, p in self.model.named_parameters()
                  if not isinstance(p, torch.nn.parameter.UninitializedParameter)]
        kdist.all_reduce_gradients([p for _, p in params])
        loss_all = loss.detach().clone()
        dist.all_reduce(loss_all)
        n_max = (self.n_global + self.world - 1) // self.world + 1
        pad = torch.zeros(n_max, device=self.dev)
        pad[:self.n_snp] = pred.detach()
        gathered = [torch.empty_like(pad) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(pad, gathered, dst=0)
        res = None
        if self.rank == 0:
            sharded = torch.cat([gathered[r][:kdist.split_range(self.n_global, r, self.world)[1]
                                              - kdist.split_range(self.n_global, r, self.world)[0]]
                                 for r in range(self.world)])
            ref_model = kgwas_b200.HeteroGNN(self.full, self.h, 1, self.L, self.backbone, "sum", self.h, self.h, self.h, 1)
            ref_model.load_state_dict(self.model.state_dict())
            ref_model = ref_model.to(self.dev)
            fd = self.full.to(self.dev)
            xf = {k: v.clone().requires_grad_() for k, v in fd.x_dict.items()}
            pf = ref_model.forward_from_hidden(xf, fd.edge_index_dict, self.n_global).reshape(-1)
            lf = torch.sum(self.w_full.to(self.dev) * (pf - self.y_full.to(self.dev)) ** 2) / self.n_global
            lf.backward()
            torch.cuda.synchronize()
            err = float((sharded - pf.detach()).abs().max() / pf.detach().abs().max().clamp_min(1e-30))
            ref_g = dict(ref_model.named_parameters())
            gscale = max(float(p.grad.abs().max()) for p in ref_g.values()
                         if not isinstance(p, torch.nn.parameter.UninitializedParameter) and p.grad is not None)
            gerr, n_g, n_none_vs_zero, bad_keys = 0.0, 0, 0, []
            for k, p in params:
                rg = ref_g[k].grad
                if rg is None and p.grad is None:
                    continue
                if rg is None or p.grad is None:
                    # un-sharded: relations into node types the last layer does not use get NO gradient (None, as in
                    # the reference); sharded: their partial rows still pass through the cross-rank sum, so autograd
                    # hands back an all-zero tensor.  Same update either way (Adam skips None, and a zero gradient
                    # with zero moments is a zero step apart from weight decay -- reported, not hidden).
                    other = rg if rg is not None else p.grad
                    if float(other.abs().max()) != 0.0:
                        gerr = float("inf")
                        bad_keys.append(k)
                    n_none_vs_zero += 1
                    continue
                gerr = max(gerr, float((p.grad - rg).abs().max()) / max(gscale, 1e-30))
                n_g += 1
            res = {"err": err, "grad_err": gerr, "tolerance": 1e-4, "n_logits": int(pf.numel()), "n_grad_tensors": n_g, "n_grads_none_vs_zero": n_none_vs_zero, "grads_missing_on_one_side": bad_keys[:5],
                   "loss_sharded": float(loss_all.item()), "loss_unsharded": float(lf.item()),
                   "against": "the same KG and weights run un-sharded on rank 0 (forward + backward), all per-SNP logits"}
            del ref_model, fd, xf, pf, lf
            self.full = None
        # EVERY rank drops its plans (rank 0 also holds the un-sharded one): the sharded plan is rebuilt by the warm-up,
        # and building it runs collectives (global in-degrees) that all ranks must enter together
        kgwas_b200.plan.clear_plan_cache()
        self.model.zero_grad(set_to_none=True)
        gc.collect()
        torch.cuda.empty_cache()
        return res

    # -- warm-up, capture ---------------------------------------------------------------------------------------------
    def prepare(self):
        import torch.distributed as dist
        for _ in range(max(self.args.warmup, 3)):
            self.step(self.x_dev)
        torch.cuda.synchronize()
        if not self.args.no_cuda_graph:
            ok = torch.ones(1, device=self.dev)
            try:
                from kgwas_b200.graphed import GraphedStep
                self.graphed = GraphedStep(self.step, self.x_dev, warmup=3)
                self.graph_note = ("whole step (fwd + bwd + " + ("all-reduces + " if self.world > 1 else "")
                                   + "Adam, all scheduler streams) captured once, replayed per step")
            except Exception as e:                               # noqa: BLE001 -- report and fall back to eager steps
                self.graphed, self.graph_note = None, f"capture failed, eager steps: {type(e).__name__}: {e}"[:300]
                ok.zero_()
                torch.cuda.synchronize()
            if self.world > 1:                  # every rank must agree on graph vs eager (collectives inside)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if ok.item() == 0 and self.graphed is not None:
                    self.graphed, self.graph_note = None, "capture failed on another rank, eager steps"
        for _ in range(3):
            self.run(self.x_dev)
        torch.cuda.synchronize()

    def run(self, x=None):
        if self.graphed is not None:
            return self.graphed(x)
        return self.step(x if x is not None else self.x_dev)

    # -- timing: device time between two events, barrier + synchronize on both sides, max over ranks ----------------
    def timed(self, fn, steps, profile=False):
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if profile:
            torch.cuda.profiler.start()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        if profile:
            torch.cuda.profiler.stop()
        if self.world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def time_resident(self, steps, profile=False):
        from kgwas_b200 import _lib
        k0 = _lib.kernel_launch_count()
        ms = self.timed(lambda i: self.run(self.x_dev), steps, profile)
        launches = _lib.kernel_launch_count() - k0
        if self.graphed is not None:                    # replays do not pass through the library's launch counter
            launches = self.graphed.kernels_per_replay * steps
        return ms, launches

    def time_e2e(self, steps):
        """Host (pinned) features -> H2D -> step -> D2H of logits + loss, every step.  Double-buffered input pipeline
        (what a DataLoader with pin_memory + non_blocking copies does): the H2D copy of step i+1's features runs on a
        copy stream while step i computes; every step's inputs still cross PCIe inside the timed region and every step
        ends with the D2H read of its logits and loss.  Same code for N = 1 and N > 1."""
        dev = self.dev
        out_host = torch.empty(self.n_snp, dtype=torch.float32).pin_memory()
        copy_stream = torch.cuda.Stream(dev)
        main_stream = torch.cuda.current_stream(dev)
        bufs = [{k: torch.empty_like(v, device=dev) for k, v in self.x_host.items()} for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            b = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])          # the step that last used this buffer is done
                for k, v in self.x_host.items():
                    bufs[b][k].copy_(v, non_blocking=True)
                ready[b].record(copy_stream)

        def e2e_step(i):
            b = i % 2
            main_stream.wait_event(ready[b])
            prefetch(i + 1)
            if self.graphed is not None:                     # D2D into the graph's static inputs, then replay
                pred, loss = self.graphed(bufs[b])
            else:
                x = {k: v.detach().requires_grad_() for k, v in bufs[b].items()}
                pred, loss = self.step(x)
            consumed[b].record(main_stream)
            out_host.copy_(pred.detach(), non_blocking=True)
            return loss.item()                               # D2H + sync

        for b in range(2):
            consumed[b].record(main_stream)
        prefetch(0)
        for i in range(2):
            e2e_step(i)
        ms = self.timed(lambda i: e2e_step(i + 2), steps)
        torch.cuda.synchronize()
        h2d = sum(v.numel() * 4 for v in self.x_host.values())
        return ms, h2d, self.n_snp * 4 + 8

    def close(self):
        """Release the captured graph before anything tears NCCL down."""
        if self.graphed is not None:
            torch.cuda.synchronize()
            try:
                self.graphed.graph.reset()
            except Exception:                                 # noqa: BLE001
                pass
            self.graphed = None
        self.opt = None
        self.model = None
        self.x_dev = self.ei = None
        gc.collect()
        torch.cuda.empty_cache()


def gat_block(args, dev):
    """BASELINE configs[2] / configs[4] shapes on the same synthetic KG: 2-layer GAT h=128, 3-layer GAT h=256."""
    import kgwas_b200
    peak = _peak()[0]
    out = []
    for (L, h) in ((2, 128), (3, 256)):
        entry = {"backbone": "GAT", "layers": L, "hidden": h}
        try:
            r = Runner(args, dev, 0, 1, "single", backbone="GAT", hidden=h, layers=L)
            r.prepare()
            ms, launches = r.time_resident(args.gat_steps)
            edges_step = L * r.edges_local
            b_layer = layer_bytes(r.sizes, r.nodes, h, "GAT")
            gbs = L * b_layer / (ms * 1e-3) / 1e9
            entry.update({"ms_per_step": ms, "value": edges_step / (ms * 1e-3), "unit": "edges/s", "steps": args.gat_steps,
                          "edges_per_step": edges_step, "bytes_per_layer": b_layer,
                          "bytes_per_edge": b_layer / r.edges_local,
                          "roofline_step": {"achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                            "formula": "SURVEY.md 8(d) B_layer(h) + 16 B/edge attention scalars"},
                          "gpu_launches_per_step": launches / args.gat_steps, "cuda_graph": r.graph_note,
                          "config": workload_config(args, 1, r.sizes, r.nodes, "GAT", h, L)["baseline_config"]})
            r.close()
            del r
        except Exception as e:                                   # noqa: BLE001
            entry["error"] = f"{type(e).__name__}: {e}"[:400]
            torch.cuda.synchronize()
        kgwas_b200.plan.clear_plan_cache()
        gc.collect()
        torch.cuda.empty_cache()
        out.append(entry)
    return out


def _peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    return peak, src


def spmm_roofline(args, r, ms_step):
    """Dominant-kernel roofline: CUDA events around every kgb_spmm launch, on the launching stream, in a second timed
    region with the side streams switched off (concurrent kernels would share the machine and the per-launch durations
    would not be attributable)."""
    from kgwas_b200 import _lib, ops as _ops
    from kgwas_b200.graphed import GraphedStep
    prof = LaunchProfiler()
    _ops.MULTI_STREAM = False
    r.step(r.x_dev)
    torch.cuda.synchronize()
    _lib._prof = prof
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_steps = max(3, min(10, args.steps))
    pe0.record()
    for _ in range(prof_steps):
        r.step(r.x_dev)
    pe1.record()
    torch.cuda.synchronize()
    _lib._prof = None
    ms_serial = pe0.elapsed_time(pe1) / prof_steps
    serial_note = "eager, host-bound"
    if r.graphed is not None:
        # the eager single-stream step is bound by the host; the denominator of the kernel's share of the step is the
        # same single-stream step replayed from a CUDA graph (pure device time, kernels back to back)
        try:
            g1 = GraphedStep(r.step, r.x_dev, warmup=1)
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
    peak, peak_src = _peak()
    spmm_gbs = (spmm_bytes / (spmm_ms * 1e-3) / 1e9) if spmm_ms > 0 else 0.0
    # measured DRAM traffic of the same launches (ncu --set full, dram__bytes_read + dram__bytes_write), per launch like
    # `achieved`: written by scratch/ncu_traffic.py from the capture committed under profiles/ (newest round first)
    traffic, traffic_src = None, None
    for name in ("r02_spmm_traffic.json", "r01_spmm_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            if tj.get("hidden") == r.h and tj.get("backbone") == r.backbone and args.scale == 1.0:
                traffic = tj["dram_bytes_per_step"] / max(1, n_spmm // prof_steps)
                traffic_src = tj["source"]
                break
        except Exception:                                     # noqa: BLE001
            pass
    return {"bound": "hbm",
            "kernel": f"kgb_spmm = lean::k_spmm_lean<{r.h // 128}> (segmented gather-reduce, all kgb_spmm launches of a step; "
                      f"the opt-in hub-tile path hub::k_hub_tile is {'ON' if os.environ.get('KGB_SPMM_HUB') == '1' else 'off'})",
            "achieved": spmm_gbs, "peak": peak, "unit": "GB/s", "frac": spmm_gbs / peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src, "launches_per_step": n_spmm // prof_steps,
            "algorithmic_bytes_per_step": spmm_bytes / prof_steps,
            "avg_launch_ms": spmm_ms / max(1, n_spmm), "kernel_share_of_step": spmm_ms / (ms_serial * prof_steps),
            "measured_in": f"{prof_steps} extra eager steps, single stream, CUDA events around every kgb_spmm launch; "
                           f"share = their sum / single-stream step ({ms_serial:.3f} ms/step, {serial_note})"}


def ours(args):
    import kgwas_b200  # noqa: F401
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        # a mismatched collective should fail in two minutes, not hold N GPUs for NCCL's default ten
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    steps, warm = args.steps, max(args.warmup, 3)
    peak, peak_src = _peak()

    def measure(mode):
        r = Runner(args, dev, rank, world, mode)
        parity = None
        if mode == "strong" and not args.no_parity:
            r.step(r.x_dev)                                   # materialise the lazy weights
            r.opt = None
            parity = r.parity_vs_unsharded()
        r.prepare()
        clocks = ClockSampler(local_rank)
        clocks.start()
        ms, launches = r.time_resident(steps, args.profile_range)
        clk = clocks.stop()
        el = torch.tensor([r.edges_local], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el)
        edges_layer = float(el.item())
        res = {"ms": ms, "launches": launches, "clk": clk, "edges_layer": edges_layer, "parity": parity, "runner": r,
               "edges_step": r.L * edges_layer}
        if not args.no_e2e:
            ems, h2d, d2h = r.time_e2e(steps)
            res["e2e"] = {"value": res["edges_step"] / (ems * 1e-3), "unit": "edges/s", "ms_per_step": ems,
                          "h2d_bytes_per_step": _sum_all(h2d, dev, world),
                          "d2h_bytes_per_step": _sum_all(d2h, dev, world),
                          "api": "HeteroGNN.forward_from_hidden(x_dict, edge_index_dict, batch_size) + loss.backward() + "
                                 "Adam.step(); node features copied from pinned host memory every step (double-buffered on a "
                                 "copy stream, overlapping the previous step), graph resident (data_to_cuda=True, kgwas.py:96-97)"}
        return res

    if world == 1:
        m = measure("single")
        r = m["runner"]
        line = {"metric": METRIC, "value": m["edges_step"] / (m["ms"] * 1e-3), "unit": "edges/s", "n_gpus": 1,
                "steps": steps, "warmup": warm, "ms_per_step": m["ms"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, 1, r.sizes, r.nodes),
                "edges_per_step": m["edges_step"], "edges_per_layer": m["edges_layer"], "num_nodes": r.nodes,
                "edge_counts": {"|".join(k): v for k, v in r.sizes.items()}}
        line["roofline"] = spmm_roofline(args, r, m["ms"])
        b_layer = layer_bytes(r.sizes, r.nodes, r.h, r.backbone)
        step_gbs = r.L * b_layer / (m["ms"] * 1e-3) / 1e9
        line["roofline_step"] = {"formula": "SURVEY.md 8(d) B_layer(h)", "bytes_per_layer": b_layer,
                                 "bytes_per_edge": b_layer / m["edges_layer"], "achieved": step_gbs, "peak": peak,
                                 "unit": "GB/s", "frac": step_gbs / peak, "frac_of_8000_spec": step_gbs / 8000.0}
        line.update({"clocks": m["clk"], "gpu_launches": m["launches"], "gpu_launches_per_step": m["launches"] / steps,
                     "cuda_graph": r.graph_note})
        if "e2e" in m:
            line["e2e"] = m["e2e"]
        r.close()
        del r, m["runner"]
        kgwas_b200.plan.clear_plan_cache()
        gc.collect()
        torch.cuda.empty_cache()
        if not args.no_gat and args.backbone == "SAGE" and args.scale == 1.0:
            line["gat"] = gat_block(args, dev)
        if not args.no_cpu_baseline:
            c = run_cpu(args, 2, 1, args.cpu_scale, parity_dev=dev if not args.no_parity else None)
            line["cpu_baseline"] = {"value": c["value"], "unit": "edges/s", "cores": c["cores"], "kind": "port",
                                    "sample": c["sample"]}
            if c["parity"] is not None:
                line["parity"] = c["parity"]
        print(json.dumps(line), flush=True)
        return

    # ---- N > 1
    modes = {"strong": ["strong"], "weak": ["weak"], "both": ["strong", "weak"]}[args.scaling]
    results = {}
    for mode in modes:
        m = measure(mode)
        r = m.pop("runner")
        m["nodes"], m["sizes"], m["graph_note"] = r.nodes, r.sizes, r.graph_note
        r.close()
        del r
        kgwas_b200.plan.clear_plan_cache()
        gc.collect()
        torch.cuda.empty_cache()
        results[mode] = m
    head_mode = modes[0]
    m = results[head_mode]
    if rank == 0:
        line = {"metric": METRIC, "value": m["edges_step"] / (m["ms"] * 1e-3), "unit": "edges/s", "n_gpus": world,
                "steps": steps, "warmup": warm, "ms_per_step": m["ms"], "higher_is_better": True, "scaling": head_mode,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, world, mode=head_mode),
                "edges_per_step": m["edges_step"], "edges_per_layer_all_ranks": m["edges_layer"],
                "rank0_num_nodes": m["nodes"],
                "collectives_per_step": "per layer: one all-reduce(sum) of the shared node types' partial rows forward and "
                                        "one of their input gradients backward (both on a communication stream, overlapping the "
                                        "SNP-row kernels); one flat all-reduce of the parameter gradients",
                "clocks": m["clk"], "gpu_launches": m["launches"], "gpu_launches_per_step": m["launches"] / steps,
                "cuda_graph": m["graph_note"]}
        if m.get("parity") is not None:
            line["parity"] = m["parity"]
        if "e2e" in m:
            line["e2e"] = m["e2e"]
        if "weak" in results and head_mode != "weak":
            wk = results["weak"]
            line["weak"] = {"value": wk["edges_step"] / (wk["ms"] * 1e-3), "unit": "edges/s", "ms_per_step": wk["ms"],
                            "edges_per_step": wk["edges_step"], "scaling": "weak",
                            "workload": f"{world} SNP blocks of {wk['nodes']['SNP']} variants, shared gene/GO graph split by destination",
                            "e2e": wk.get("e2e")}
        print(json.dumps(line), flush=True)
    _teardown(dev)


def _sum_all(v, dev, world):
    if world == 1:
        return v
    import torch.distributed as dist
    t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
    dist.all_reduce(t)
    return int(t.item())


def _teardown(dev):
    """Captured graphs that hold NCCL kernels were reset in Runner.close(); tear the communicator down normally.  A
    watchdog turns a hang (seen in round 1 when a live graph still referenced the communicator) into a clean exit
    AFTER the result line is out, and says so on stderr."""
    import torch.distributed as dist

    def bail():
        sys.stderr.write("bench.py: destroy_process_group() did not return within 20 s; leaving with os._exit(0)\n")
        sys.stderr.flush()
        os._exit(0)

    t = threading.Timer(20.0, bail)
    t.daemon = True
    t.start()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    t.cancel()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)
