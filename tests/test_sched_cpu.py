"""CPU: the launch scheduler's pure logic (kgwas_b200.ops.issue_order) -- the dependency DAG derived from program order
and the critical-first topological issue order that ``_Sched.join`` turns into stream placement and event waits."""
import random

from kgwas_b200.ops import issue_order


def _simulate(ops, order):
    """Run the launches in ``order``; each one hashes the current contents of what it reads (and of what it writes:
    accumulating kernels) into what it writes.  Any legal reordering must end in the same buffer contents."""
    state = {}
    for i in order:
        _, rk, wk = ops[i]
        seen = tuple(state.get(k, 0) for k in list(rk) + list(wk))
        for k in wk:
            state[k] = hash((i, k, seen))
    return state


def test_issue_order_is_a_legal_reordering():
    rng = random.Random(7)
    for trial in range(200):
        n_buf = rng.randint(2, 9)
        ops = []
        for _ in range(rng.randint(1, 40)):
            rk = rng.sample(range(n_buf), rng.randint(0, min(3, n_buf)))
            wk = rng.sample(range(n_buf), rng.randint(1, min(2, n_buf)))
            ops.append((rng.random() < 0.3, rk, wk))
        preds, succs, order = issue_order(ops)
        assert sorted(order) == list(range(len(ops)))
        pos = {i: p for p, i in enumerate(order)}
        for i, ps in enumerate(preds):
            assert all(p < i for p in ps)                      # the DAG follows program order
            assert all(pos[p] < pos[i] for p in ps)            # and the issue order respects it
        for i, ss in enumerate(succs):
            assert all(i in preds[j] for j in ss)
        assert _simulate(ops, order) == _simulate(ops, range(len(ops))), trial


def test_hazards():
    # 0 writes A; 1 reads A (RAW on 0); 2 writes A (WAW on 0, WAR on 1); 3 reads B only (independent)
    ops = [(False, [], ["A"]), (False, ["A"], ["C"]), (False, [], ["A"]), (False, ["B"], ["D"])]
    preds, _, order = issue_order(ops)
    assert preds[0] == set() and preds[1] == {0} and preds[2] == {0, 1} and preds[3] == set()
    assert order == [0, 1, 2, 3]                               # nothing critical: program order


def test_critical_launches_go_first():
    # program order: three independent small launches, then a small one (3) that feeds the big one (4)
    ops = [(False, ["x"], ["a"]), (False, ["x"], ["b"]), (False, ["x"], ["c"]), (False, ["x"], ["z"]), (True, ["z"], ["out"])]
    _, _, order = issue_order(ops)
    assert order[:2] == [3, 4]                                 # the big kernel's producer, then the big kernel
    assert order[2:] == [0, 1, 2]                              # the rest keeps program order
    # a small launch that WAITS for a big result must not hold up independent small launches queued after it
    ops = [(True, ["x"], ["big"]), (False, ["big"], ["y"]), (False, ["x"], ["a"]), (False, ["a"], ["b"])]
    preds, _, order = issue_order(ops)
    assert preds[1] == {0} and preds[2] == set()
    assert order[0] == 0 and sorted(order) == [0, 1, 2, 3]


def test_longest_chain_first_among_ready_big_launches(monkeypatch):
    """Layer-forward shape: root GEMM (big) -> Gene->SNP gather (big, nothing waits for it); SNP->Gene gather (big) ->
    gene-sized GEMM (small).  The gather with work behind it is issued before the one without, so the small GEMM runs
    under a big kernel instead of after the last one."""
    ops = [(True, ["x_snp"], ["out_snp"]),               # 0 root term of the SNP rows
           (False, ["x_gene"], ["z"]),                    # 1 Z = X_gene W^T
           (True, ["z", "out_snp"], ["out_snp"]),         # 2 Gene -> SNP gather-reduce
           (False, ["x_gene"], ["out_gene"]),             # 3 root term of the gene rows
           (True, ["x_snp"], ["A"]),                      # 4 SNP -> Gene gather-reduce
           (False, ["A", "out_gene"], ["out_gene"])]      # 5 out_gene += A W^T
    from kgwas_b200 import ops as _ops
    monkeypatch.setattr(_ops, "CHAIN_PRIORITY", True)
    _, _, order = issue_order(ops)
    assert order.index(4) < order.index(2)
    assert order.index(0) < order.index(2) and order.index(1) < order.index(2) and order.index(4) < order.index(5)


def test_row_slices_of_one_buffer_are_independent_regions():
    """(storage, first byte, end byte) keys: launches on disjoint row-slices of one flat buffer do not depend on each
    other; a launch on the whole buffer (the all-reduce of every shared node type) depends on all of them, and whatever
    reads a slice afterwards depends on that launch."""
    buf, lo, mid, hi = 7, 0, 4096, 8192
    ops = [(False, ["x"], [(buf, lo, mid)]),            # 0 writes slice A
           (False, ["y"], [(buf, mid, hi)]),            # 1 writes slice B
           (False, [(buf, lo, hi)], [(buf, lo, hi)]),   # 2 all-reduce in place over the whole buffer
           (False, [(buf, lo, mid)], ["outA"]),         # 3 reads slice A
           (False, ["z"], [(buf, mid, hi)])]            # 4 overwrites slice B (WAR on 2 and nothing else of A)
    preds, _, order = issue_order(ops)
    assert preds[0] == set() and preds[1] == set()
    assert preds[2] == {0, 1}
    assert preds[3] == {2}
    assert 2 in preds[4] and 3 not in preds[4]
    assert sorted(order) == [0, 1, 2, 3, 4]
