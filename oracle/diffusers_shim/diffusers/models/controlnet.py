from dataclasses import dataclass
from typing import Tuple

import torch

from ..utils import BaseOutput


@dataclass
class ControlNetOutput(BaseOutput):
    down_block_res_samples: Tuple[torch.Tensor] = None
    mid_block_res_sample: torch.Tensor = None
