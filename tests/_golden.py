"""Loader for the fixtures minted by tests/golden/make_golden.py (reference outputs)."""
import os

import numpy as np

from oracle import lstm_lm_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STEP_CASES = ["tiny_pytorch", "tiny_custom", "tiny_dropout", "tiny_carry3",
              "edge_T1_B1_L1", "odd_H40_custom_dropout", "mid_H72"]


class StepCase:
    """One `step_case` fixture: reference parameters, inputs and per-step outputs,
    with custom-path tensors already mapped to the pytorch-path names / gate order."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.z = z
        self.name = name
        self.V, self.H, self.L, self.T, self.B, self.steps = [int(v) for v in z["meta"]]
        self.lstm_type = str(z["lstm_type"])
        self.dropout = float(z["dropout"])
        self.lr = float(z["lr"])
        self.max_norm = float(z["max_norm"])
        self.winit = float(z["winit"])
        self.seed = int(z["seed"])
        self.names = O.param_names(self.L)

    def _group(self, prefix):
        d = {k[len(prefix):]: self.z[k] for k in self.z.files if k.startswith(prefix)}
        if self.lstm_type == "custom":
            d = O.custom_state_dict_to_pytorch(d)
        return d

    def params0(self, dtype=np.float32):
        return {k: v.astype(dtype) for k, v in self._group("param0/").items()}

    def states0(self, dtype=np.float32):
        return [(self.z[f"h0/{l}"].astype(dtype), self.z[f"c0/{l}"].astype(dtype)) for l in range(self.L)]

    def x(self, s):
        return self.z[f"s{s}/x"]

    def y(self, s):
        return self.z[f"s{s}/y"]

    def masks(self, s):
        if self.dropout == 0:
            return None
        n = self.T * self.B * self.H
        return [np.unpackbits(self.z[f"s{s}/mask/{i}"])[:n].reshape(self.T, self.B, self.H).astype(bool)
                for i in range(self.L + 1)]

    def scores(self, s):
        return self.z[f"s{s}/scores"]

    def loss(self, s):
        return float(self.z[f"s{s}/loss"])

    def norm(self, s):
        return float(self.z[f"s{s}/norm"])

    def grads(self, s):
        return self._group(f"s{s}/grad/")

    def params_after(self, s):
        return self._group(f"s{s}/param/")

    def states_after(self, s):
        return [(self.z[f"s{s}/h/{l}"], self.z[f"s{s}/c/{l}"]) for l in range(self.L)]
