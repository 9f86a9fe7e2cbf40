"""kgwas_b200 -- B200-native (sm_100a) message-passing engine for the KGWAS knowledge-graph
convolution: drop-in ``HeteroGNN`` / ``HeteroConv`` / ``SAGEConv`` / ``GATConv`` behind the
reference's Python surface, hand-written CUDA kernels behind a C ABI (include/kgwas_b200.h)."""
from .graph import AddSelfLoops, HeteroData, ToUndirected, make_synth_kg  # noqa: F401
from .conv import HeteroConv, Linear, SAGEConv  # noqa: F401
from .model import HeteroGNN, SimpleMLP  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name == "GATConv":
        from .gat import GATConv
        return GATConv
    if name in ("KGWAS", "KGWAS_Data"):
        from . import kgwas as _k, kgwas_data as _d
        return {"KGWAS": _k.KGWAS, "KGWAS_Data": _d.KGWAS_Data}[name]
    raise AttributeError(name)
