"""Host-side mirror of the counter-based dropout used by every fused dropout site
(get_b200/csrc/common.cuh `drop_keep`): lets tests and callers materialise the exact keep-mask a kernel
applied, e.g. to run the reference (`nn.Dropout` replaced by a supplied mask) on identical noise.
Pure integer arithmetic on int64 tensors; multipliers are < 2^31 so no product overflows.
"""
import torch

_M32 = 0xFFFFFFFF


def _hash32(x: torch.Tensor) -> torch.Tensor:
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x2C1B3C6D) & _M32
    x = x ^ (x >> 16)
    x = (x * 0x297A2D39) & _M32
    x = x ^ (x >> 15)
    return x


def keep_mask(numel: int, p: float, seed: int, device="cpu") -> torch.Tensor:
    """float32 (numel,) tensor: 1/(1-p) where element i is kept, 0 where it is dropped."""
    if p <= 0:
        return torch.ones(numel, dtype=torch.float32, device=device)
    idx = torch.arange(numel, dtype=torch.int64, device=device)
    lo = idx & _M32
    hi = (idx >> 32) & _M32
    # i32 = lo + hi * 0x632be5ab (mod 2^32), computed in two 16-bit halves of the multiplier to stay < 2^63
    i32 = (lo + ((hi * 0x632B) & _M32) * 65536 + hi * 0xE5AB) & _M32
    s = ((seed & _M32) * 0x9E3779B9) & _M32
    h = _hash32((i32 + s) & _M32)
    thr = int(torch.tensor(p, dtype=torch.float32).item() * 16777216.0)
    keep = (h >> 8) >= thr
    scale = (torch.tensor(1.0, dtype=torch.float32) / (torch.tensor(1.0, dtype=torch.float32) - torch.tensor(p, dtype=torch.float32))).item()
    return keep.to(torch.float32) * scale
