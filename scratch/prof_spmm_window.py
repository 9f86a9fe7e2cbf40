import sys, time, torch
sys.path.insert(0, '.')
from kgwas_b200 import _lib, make_synth_kg
from kgwas_b200.plan import PairJob
d = torch.device('cuda')
data = make_synth_kg(scale=1.0, seed=42, hidden=128)
rels = [et for et in data.edge_types if et[0] == 'SNP' and et[2] == 'Gene']
eis = [data[et].edge_index.to(d) for et in rels]
job = PairJob('Gene', 'SNP', rels, list(range(len(rels))), eis, data['SNP'].num_nodes, data['Gene'].num_nodes)
print(job, 'hsegs', job.csr.n_hsegs, 'hrows', job.csr.n_hrows, 'E', job.n_edges)
x = torch.randn(job.n_src, 128, device=d)
g = torch.randn(job.n_dst * job.R, 128, device=d)   # for the transposed pass we use the xf bwd analogue below
A = torch.empty(job.n_dst * job.R, 128, device=d)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for win_mb in (0, 64, 48, 32, 24, 16, 8):
    job.csr.hseg_order = None; job.csr._refresh_struct()
    if win_mb: job.csr.schedule_for_l2(512, window_bytes=win_mb << 20)
    t = timeit(lambda: _lib.spmm(job.csr, x, A, 128, ew=job.w_mean))
    print('af fwd window MB', win_mb, 'us', round(t, 1))
# seg_len sweep (rebuild csr)
for seg in (32, 64, 128, 256):
    src = torch.cat([e[0] for e in eis]); 
    slot = torch.cat([torch.full((e.size(1),), k, dtype=torch.int64, device=d) for k, e in enumerate(eis)])
    dst = torch.cat([e[1] for e in eis]) * job.R + slot
    csr, eperm, _, _ = _lib.csr_build(src, dst, job.n_src, job.n_dst * job.R, transposed=False, seg_len=seg, sort_cols=True)
    w = job.w_mean  # same slot order (sort is identical)
    for win_mb in (0, 24):
        csr.hseg_order = None; csr._refresh_struct()
        if win_mb: csr.schedule_for_l2(512, window_bytes=win_mb << 20)
        t = timeit(lambda: _lib.spmm(csr, x, A, 128, ew=w))
        print('seg_len', seg, 'hsegs', csr.n_hsegs, 'window', win_mb, 'us', round(t, 1))
