import sys, copy, torch
sys.path.insert(0, '.')
from oracle import kgwas_oracle as O
import kgwas_b200
from kgwas_b200 import make_synth_kg
cuda = torch.device('cuda')
h=32
data = make_synth_kg(scale=0.002, seed=5, hidden=h)
torch.manual_seed(0)
ref = O.HeteroGNN(data, h, 1, 2, "GAT", "sum", h, h, h, 1, no_relu=True)
ref({k: v.clone() for k, v in data.x_dict.items()}, data.edge_index_dict, 4)
ours = kgwas_b200.HeteroGNN(data, h, 1, 2, "GAT", "sum", h, h, h, 1, no_relu=True)
ours.load_state_dict(ref.state_dict())
ref64 = copy.deepcopy(ref).double()
ours = ours.to(cuda); g = data.to(cuda)
bs=150
o64 = ref64({k: v.double() for k, v in data.x_dict.items()}, data.edge_index_dict, bs); o64.sum().backward()
og = ours(g.x_dict, g.edge_index_dict, bs); og.sum().backward()
print('fwd err', ((og.cpu().double()-o64).abs().max()/o64.abs().max()).item())
p64, pg = dict(ref64.named_parameters()), dict(ours.named_parameters())
bad = []
for k in p64:
    if isinstance(p64[k], torch.nn.parameter.UninitializedParameter): continue
    a, b = p64[k].grad, pg[k].grad
    if a is None or b is None:
        if (a is None) != (b is None): print('NONE mismatch', k, a is None, b is None)
        continue
    if a.abs().max() == 0: continue
    err = ((b.cpu().double()-a).abs().max()/a.abs().max()).item()
    if err > 1e-3: bad.append((err, k))
for e, k in sorted(bad, reverse=True)[:40]: print(f'{e:.3e} {k}')
print(len(bad), 'bad of', len(p64))
