"""SchedulerMixin / SchedulerOutput / KarrasDiffusionSchedulers: the three names the reference's flow-matching UniPC
scheduler imports (fm_solvers_unipc.py:25-27).  Only what that file touches: `order`, `_compatibles`, the output record."""
from dataclasses import dataclass
from enum import Enum

import torch


class KarrasDiffusionSchedulers(Enum):
    DDIMScheduler = 1
    DDPMScheduler = 2
    PNDMScheduler = 3
    LMSDiscreteScheduler = 4
    EulerDiscreteScheduler = 5
    HeunDiscreteScheduler = 6
    EulerAncestralDiscreteScheduler = 7
    DPMSolverMultistepScheduler = 8
    DPMSolverSinglestepScheduler = 9
    KDPM2DiscreteScheduler = 10
    KDPM2AncestralDiscreteScheduler = 11
    DEISMultistepScheduler = 12
    UniPCMultistepScheduler = 13
    DPMSolverSDEScheduler = 14
    EDMEulerScheduler = 15


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor


class SchedulerMixin:
    config_name = "scheduler_config.json"
    _compatibles = []
    has_compatibles = True
