"""Shared test helpers: deterministic parameters, golden loading, oracle import (tests only)."""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def named_param_values(shapes: dict, seed: int, scale: float = 1.0) -> dict:
    """Portable deterministic parameter values: one numpy PCG64 stream per parameter NAME, so values do not
    depend on dict order, torch version or device. Weights ~ U(-a,a) with a = scale*sqrt(3/fan_in)
    (variance 1/fan_in keeps activations O(1) like the reference's Kaiming/Xavier inits); biases U(-.1,.1)."""
    out = {}
    for name, shape in shapes.items():
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        shape = tuple(int(s) for s in shape)
        if len(shape) >= 2:
            a = scale * np.sqrt(3.0 / shape[-1])
            v = rng.uniform(-a, a, size=shape)
        else:
            v = rng.uniform(-0.1, 0.1, size=shape)
        out[name] = v.astype(np.float32)
    return out


def load_golden(name: str) -> dict:
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def split_prefixed(d: dict, prefix: str) -> dict:
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def to_torch_sd(npd: dict, dtype=torch.float32, device="cpu") -> dict:
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype if v.dtype.kind == "f" else None).to(device)
            for k, v in npd.items()}


def import_oracle():
    from oracle import get_oracle
    return get_oracle


def keep_sets(idx) -> list:
    return [frozenset(int(i) for i in row) for row in np.asarray(idx)]
