"""Helpers shared by the tests: golden fixture loading (tests/golden/*.npz, written by
oracle/make_golden.py from the unmodified reference)."""
import json
from pathlib import Path

import numpy as np
import torch

from oracle import vtmae_oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = sorted(p.stem for p in GOLDEN.glob("*.npz"))


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(GOLDEN / f"{name}.npz")
        cfgd = json.loads(bytes(self.z["config_json"]).decode())
        for k in ("image_size", "tactile_size"):
            cfgd[k] = tuple(cfgd[k])
        self.cfg = O.VTMAEConfig(**cfgd)
        self.batch = int(self.z["batch"])

    def t(self, key):
        return torch.from_numpy(np.array(self.z[key]))

    def has(self, key):
        return key in self.z.files

    def inputs(self):
        return {k[2:]: self.t(k) for k in self.z.files if k.startswith("x.")}

    def noise(self):
        return self.t("noise")

    def weights(self):
        """Alias-free state dict of the case (stored, or re-derived from the seed + checksum check)."""
        if any(k.startswith("w.") for k in self.z.files):
            return {k[2:]: self.t(k).clone() for k in self.z.files if k.startswith("w.")}
        sd = O.init_state_dict(self.cfg, seed=0)
        chk = float(sum(v.double().abs().sum() for k, v in sorted(sd.items())))
        assert abs(chk - float(self.z["weights_checksum"])) <= 1e-9 * abs(chk), "RNG drift: weights differ"
        return sd

    def grad_norms(self):
        return {k[6:]: float(self.z[k]) for k in self.z.files if k.startswith("gnorm.")}

    def grad_present(self):
        return {k[5:]: bool(self.z[k]) for k in self.z.files if k.startswith("ghas.")}

    def full_grads(self):
        return {k[5:]: self.t(k) for k in self.z.files if k.startswith("grad.")}

    def after(self):
        return {k[6:]: self.t(k) for k in self.z.files if k.startswith("after.")}


VTT_DINO_CASES = sorted(p.stem for p in (GOLDEN / "vtt_dino").glob("*.npz"))


class VttDinoGolden:
    """tests/golden/vtt_dino/*.npz (oracle/make_golden_vtt_dino.py): models/VTT.py::VTT.forward_features."""

    def __init__(self, name):
        from oracle import vtt_dino_oracle as VD
        self.z = np.load(GOLDEN / "vtt_dino" / f"{name}.npz")
        cfgd = json.loads(bytes(self.z["config_json"]).decode())
        for k in ("image_size", "tactile_size"):
            cfgd[k] = tuple(cfgd[k])
        self.cfg = VD.VTTDinoConfig(**cfgd)
        self.batch = int(self.z["batch"])

    def t(self, key):
        return torch.from_numpy(np.array(self.z[key]))

    def weights(self):
        return {k[2:]: self.t(k).clone() for k in self.z.files if k.startswith("w.")}

    def inputs(self):
        return {k[2:]: self.t(k) for k in self.z.files if k.startswith("x.")}

    def masks(self):
        return [self.t(f"mask.{i}") for i in range(int(self.z["n_masks"]))] or None

    def objective(self, out):
        w1, w2 = self.t("w1").to(out["x_prenorm"].device), self.t("w2").to(out["x_prenorm"].device)
        return (out["x_norm_patchtokens"] * w1).sum() + (out["x_norm_regtokens"] ** 2).sum() + 0.5 * (out["x_prenorm"] * w2).sum()

    def grad_present(self):
        return {k[5:]: bool(self.z[k]) for k in self.z.files if k.startswith("ghas.")}

    def grad_norms(self):
        return {k[6:]: float(self.z[k]) for k in self.z.files if k.startswith("gnorm.")}

    def full_grads(self):
        return {k[5:]: self.t(k) for k in self.z.files if k.startswith("grad.")}
