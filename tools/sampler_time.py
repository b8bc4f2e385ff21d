"""Point-cloud batch sampler alone and in the training loop (bench.py's aux_device_sampler, stand-alone)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import aux_device_sampler  # noqa: E402
from diffudf_b200 import SIREN  # noqa: E402
from diffudf_b200.train import FusedTrainer  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(123)
out = aux_device_sampler(dev, FusedTrainer(SIREN(3, 1, [256] * 8, w0=30).to(dev), precision="tcx3"))
for k, v in out.items():
    print(k, v)
