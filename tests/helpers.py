"""Shared loaders for the golden fixtures (tests/golden/*.npz, written by
oracle/gen_golden.py from the unmodified reference)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["cont_fixture", "disc_fixture", "cont_ragged", "disc_ragged"]
LAMBDA_CASES = ["lambda_h15", "lambda_h3", "lambda_h1", "lambda_h64"]


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        if "dims" in self.z and len(self.z["dims"]) == 11:
            (self.B, self.H, self.T, self.D, self.S, self.C, self.Hd, self.E, self.A, self.Ha,
             self.La) = [int(v) for v in self.z["dims"]]
            self.variant = "continuous" if name.startswith("cont") else "discrete"
            self.S_in = self.S if self.variant == "continuous" else self.S * self.C

    def t(self, key, dtype=torch.float32):
        return torch.from_numpy(self.z[key]).to(dtype)

    def weights(self, prefix, dtype=torch.float32):
        return {k[len(prefix):]: torch.from_numpy(v).to(dtype)
                for k, v in self.z.items() if k.startswith(prefix)}

    def grads(self, prefix):
        return {k[len(prefix):]: torch.from_numpy(v) for k, v in self.z.items() if k.startswith(prefix)}
