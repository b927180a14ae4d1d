"""Host-side mirror of the counter-based dropout used by every fused dropout site
(get_b200/csrc/common.cuh `drop_bits` / `drop_keep`): lets tests and callers materialise the exact keep-mask a
kernel applied, e.g. to run the reference (`nn.Dropout` replaced by a supplied mask) on identical noise.

One 32-bit hash serves an aligned pair of elements, 16 bits each; element i is kept iff its 16-bit field is
>= floor(p * 65536). Pure integer arithmetic (numpy uint64 with explicit 32-bit masking).
"""
import numpy as np
import torch

_M32 = np.uint64(0xFFFFFFFF)


def _mul32(x, c):
    return (x * np.uint64(c)) & _M32


def _mix32(x):
    x = x ^ (x >> np.uint64(15))
    x = _mul32(x, 0x2C1B3C6D)
    x = x ^ (x >> np.uint64(12))
    x = _mul32(x, 0x297A2D39)
    x = x ^ (x >> np.uint64(15))
    return x


def drop_threshold(p: float) -> int:
    return int(np.float32(p) * np.float32(65536.0))


def keep_mask(numel: int, p: float, seed: int, device="cpu", salt: int = 0) -> torch.Tensor:
    """float32 (numel,) tensor: 1/(1-p) where element i is kept, 0 where it is dropped. `salt` = the device salt word
    (get_dropout_salt_*), 0 unless a captured training step advanced it."""
    seed = (seed + salt) & 0xFFFFFFFF
    if p <= 0:
        return torch.ones(numel, dtype=torch.float32, device=device)
    idx = np.arange(numel, dtype=np.uint64)
    word = idx >> np.uint64(1)
    lo, hi = word & _M32, (word >> np.uint64(32)) & _M32
    s = _mul32(np.uint64(seed & 0xFFFFFFFF), 0x85EBCA6B)
    x = (_mul32(lo, 0x9E3779B1) + _mul32(hi, 0x632BE5AB) + s + np.uint64(0x6A09E667)) & _M32
    bits = _mix32(x)
    field = (bits >> (np.uint64(16) * (idx & np.uint64(1)))) & np.uint64(0xFFFF)
    keep = field >= np.uint64(drop_threshold(p))
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return torch.from_numpy(keep.astype(np.float32) * scale).to(device)
