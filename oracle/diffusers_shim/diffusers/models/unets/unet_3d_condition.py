from dataclasses import dataclass

import torch

from ...utils import BaseOutput


@dataclass
class UNet3DConditionOutput(BaseOutput):
    sample: torch.FloatTensor = None
