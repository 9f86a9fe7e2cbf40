import math

import torch
import torch.nn.functional as F
from torch import nn

from ..inits import glorot


class Linear(nn.Module):
    """torch_geometric.nn.dense.linear.Linear: lazy in_channels=-1, optional glorot initialiser."""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None, bias_initializer=None):
        super().__init__()
        self.in_channels, self.out_channels, self.weight_initializer = in_channels, out_channels, weight_initializer
        if in_channels > 0:
            self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        else:
            self.weight = nn.parameter.UninitializedParameter()
            self._hook = self.register_forward_pre_hook(self.initialize_parameters)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if isinstance(self.weight, nn.parameter.UninitializedParameter):
            return
        if self.weight_initializer == "glorot":
            glorot(self.weight)
        else:
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.weight.size(1)) if self.weight.size(1) > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    @torch.no_grad()
    def initialize_parameters(self, module, inputs):
        if isinstance(self.weight, nn.parameter.UninitializedParameter):
            self.in_channels = inputs[0].size(-1)
            self.weight.materialize((self.out_channels, self.in_channels))
            self.reset_parameters()
        self._hook.remove()

    def _save_to_state_dict(self, destination, prefix, keep_vars):     # PyG keeps a lazy weight as-is
        if isinstance(self.weight, nn.parameter.UninitializedParameter):
            destination[prefix + "weight"] = self.weight
            if self.bias is not None:
                destination[prefix + "bias"] = self.bias if keep_vars else self.bias.detach()
        else:
            super()._save_to_state_dict(destination, prefix, keep_vars)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)
