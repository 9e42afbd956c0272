"""`accelerate.utils.set_seed` as used at /root/reference/inference_ID-Booth.py:8,67: seeds the
python, numpy and torch (CPU + all CUDA devices) generators."""
import random

import numpy as np
import torch


def set_seed(seed: int, device_specific: bool = False, deterministic: bool = False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
