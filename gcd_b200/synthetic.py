"""Deterministic, order-independent synthetic weights and inputs (no checkpoints or datasets are available offline).
Used by bench.py / smoke for the product path and re-exported by oracle/weights.py for the parity tests.

The reference zero-initialises 339 tensors (zero_module on every ResBlock out conv, proj_out, final out conv —
SURVEY.md §8(c)), so a default-init forward is identically zero. Every tensor is therefore overwritten from a
per-name seeded generator: conv/linear weights ~ N(0, 1/fan_in), norm scales ~ 1 + 0.1 N, biases ~ 0.05 N,
mix factors ~ 0.5 N.
"""
import zlib

import torch


def seeded_tensor(name, shape, seed=0):
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
    shape = tuple(shape)
    if name.endswith("mix_factor"):
        return torch.randn(shape, generator=g) * 0.5
    if len(shape) == 1:
        if name.endswith(".weight"):
            return 1.0 + 0.1 * torch.randn(shape, generator=g)
        return 0.05 * torch.randn(shape, generator=g)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return torch.randn(shape, generator=g) * fan_in ** -0.5


def seeded_state(shapes, seed=0):
    """shapes: mapping name -> shape. Returns {name: float32 tensor}."""
    return {k: seeded_tensor(k, v, seed) for k, v in shapes.items()}


def seeded_inputs(cfg, B, T, H, W, seed=1234):
    """Synthetic sampler inputs per SURVEY.md §8(d): x0, cond c, uncond uc (crossattn/concat zeroed,
    models/diffusion.py:522-524), image_only_indicator."""
    g = torch.Generator().manual_seed(seed)
    BT = B * T
    x = torch.randn(BT, 4, H, W, generator=g)
    ydim = cfg["adm_in_channels"] + cfg["aux_emb_dim"]
    c = {"crossattn": torch.randn(BT, 1, cfg["context_dim"], generator=g),
         "concat": torch.randn(BT, 4, H, W, generator=g),
         "vector": torch.rand(BT, ydim, generator=g) * 2 - 1}
    uc = {"crossattn": torch.zeros_like(c["crossattn"]), "concat": torch.zeros_like(c["concat"]),
          "vector": c["vector"].clone()}
    ioi = torch.zeros(2 * B, T)
    return x, c, uc, ioi
