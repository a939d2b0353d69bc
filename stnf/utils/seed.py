"""Seeding (upstream stnf/utils/seed.py:9-27): python, numpy and torch generators, in that order."""
import random

import numpy as np
import torch


def set_seed(seed: int):
    for seeder in (random.seed, np.random.seed, torch.manual_seed):
        seeder(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
        torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    print(f"[INFO] Seed set to {seed}")
