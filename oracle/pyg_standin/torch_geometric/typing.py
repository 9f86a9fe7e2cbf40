from typing import Optional, Tuple, Union

from torch import Tensor

NoneType = type(None)
Adj = Tensor
OptTensor = Optional[Tensor]
OptPairTensor = Tuple[Tensor, Optional[Tensor]]
PairTensor = Tuple[Tensor, Tensor]
Size = Optional[Tuple[int, int]]
