from typing import Optional, Union
from torch import Tensor

Adj = Union[Tensor, object]
OptTensor = Optional[Tensor]
Size = Optional[tuple]
