# coding: utf-8
"""Shared helpers for the tests: load a golden case, build oracle inputs from it."""
import ast
import os

import numpy as np
import torch

from oracle import wavenet_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class GoldenCase:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.name = name
        self.kw = ast.literal_eval(str(z["kw"]))
        self.B, self.T, self.B_free = int(z["B"]), int(z["T"]), int(z["B_free"])
        self.seed = int(z["seed"])
        self.sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
        self.noise = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("noise.")}
        self.noise_tf = {k[9:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("noise_tf.")}
        self.arr = {k: z[k] for k in z.files if "." not in k and k != "kw"}
        kw = self.kw
        self.cfg = orc.PathConfig(
            out_channels=kw["out_channels"], layers=kw["layers"], stacks=kw["stacks"],
            residual_channels=kw["residual_channels"], gate_channels=kw["gate_channels"],
            skip_out_channels=kw["skip_out_channels"], kernel_size=kw.get("kernel_size", 3),
            cin_channels=kw.get("cin_channels", -1), gin_channels=kw.get("gin_channels", -1),
            scalar_input=kw.get("scalar_input", False),
            output_distribution=kw.get("output_distribution", "Logistic"))
        self.w = orc.weights_from_state_dict(self.cfg, self.sd)

    def t(self, key):
        return None if key not in self.arr else torch.from_numpy(self.arr[key])

    @property
    def x_tf(self):
        """Teacher-forcing input in the reference's (B,C,T) layout."""
        x = self.t("x_tf")
        if self.cfg.scalar_input:
            return x
        return torch.zeros(self.B, self.cfg.out_channels, self.T).scatter_(
            1, x.long().unsqueeze(1), 1.0)
