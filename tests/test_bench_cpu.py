"""CPU: bench.py's bookkeeping -- the SURVEY 8(d) byte formula on kgwas-synth-v1, the workload description, the
reference arm's JSON line (contract keys, honest step / sample fields)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_layer_bytes_formula_on_the_named_graph():
    from kgwas_b200.graph import SYNTH_NODES
    # edge counts of kgwas-synth-v1 at scale 1.0 as bench.py prints them (profiles/r02_bench_n1.json: edge_counts)
    line = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")).read().strip().splitlines()[-1])
    sizes = {tuple(k.split("|")): v for k, v in line["edge_counts"].items()}
    nodes = {k: int(v) for k, v in line["num_nodes"].items()}
    assert nodes == SYNTH_NODES and sum(sizes.values()) == 18400753 and len(sizes) == 27
    b128 = bench.layer_bytes(sizes, nodes, 128)
    assert b128 == line["roofline_step"]["bytes_per_layer"] == 5939400824
    assert abs(b128 / sum(sizes.values()) - 322.78) < 0.01
    # GAT adds the four per-edge attention scalars
    assert bench.layer_bytes(sizes, nodes, 128, "GAT") - b128 == 16 * sum(sizes.values())
    # the line itself: contract keys
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches", "gat", "parity"):
        assert key in line, key
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] == 4 * 128 * sum(nodes.values())
    assert abs(line["value"] - line["edges_per_step"] / (line["ms_per_step"] * 1e-3)) < 1e-3 * line["value"]


def test_reference_arm_line_says_what_it_ran():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-scale", "0.005"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1
    assert line["sample_graph_scale"] == 0.005 and "scale 0.005" in line["config"]["reference_sample"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == os.cpu_count()
    assert line["e2e"] == {"value": line["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(line["value"] - line["edges_per_step_of_sample"] / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]
