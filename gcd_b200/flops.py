"""Analytic algorithmic FLOP counts (2*MAC) of the hot path, used for roofline fractions and CPU-sample scaling.

Counts the *minimal* work (length-1 cross attention folded into a bias) — never work that is
skipped; it reproduces SURVEY.md §8(d): 86.160 TFLOP per CFG forward (28 frames, 72x128), 97.202 TFLOP per VAE
decode, 2251.2 TFLOP per 25-step clip => 160.8 TFLOP per latent frame (tests/test_host_cpu.py)."""
from . import spec


def unet_forward_flops(cfg, n, H, W, T=14):
    inp, mid, out = spec.unet_plan(cfg)
    E = 4 * cfg["model_channels"]
    fl = 0.0
    # embedding head: time_embed + label_emb (+aux) + one emb_layers Linear per ResBlock (x2: spatial + time_stack)
    mc = cfg["model_channels"]
    fl += 2.0 * n * (mc * E + E * E + cfg["adm_in_channels"] * E + E * E)
    if cfg["aux_emb_dim"] > 0:
        fl += 2.0 * n * (cfg["aux_emb_dim"] * E + E * E)
    h, w = H, W
    for layers in inp + [mid] + out:
        for kind, p, cin, cout in layers:
            hw = h * w
            if kind == "conv_in":
                fl += 2.0 * n * hw * cin * cout * 9
            elif kind == "down":
                h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
                fl += 2.0 * n * h * w * cin * cout * 9
            elif kind == "up":
                h, w = 2 * h, 2 * w
                fl += 2.0 * n * h * w * cin * cout * 9
            elif kind == "vrb":
                fl += 2.0 * n * hw * (cin * cout * 9 + cout * cout * 9)          # two 3x3 convs
                if cin != cout:
                    fl += 2.0 * n * hw * cin * cout                              # 1x1 skip
                fl += 2.0 * n * hw * (2 * cout * cout * 3)                       # two (3,1,1) convs
                fl += 2.0 * n * 2 * E * cout                                     # emb_layers x2
            elif kind == "svt":
                c = cout
                rows = n * hw
                lin = lambda k, m: 2.0 * rows * k * m
                fl += lin(c, c) * 2                                              # proj_in, proj_out
                for temporal in (False, True):
                    fl += lin(c, 3 * c) + lin(c, c)                              # self-attn qkv + out
                    seq = T if temporal else hw
                    fl += 4.0 * rows * seq * c                                   # QK^T + PV
                    fl += lin(c, 8 * c) + lin(4 * c, c)                          # GEGLU FF
                    if temporal:
                        fl += lin(c, 8 * c) + lin(4 * c, c)                      # ff_in
                fl += 2.0 * n * (c * 4 * c * 2)                                  # time_pos_embed (as written; tiny)
    fl += 2.0 * n * h * w * cfg["model_channels"] * cfg["out_channels"] * 9
    return fl


def decoder_flops(cfg, n, H, W):
    fl = 0.0
    h, w = H, W
    for kind, p, cin, cout in spec.decoder_plan(cfg):
        hw = h * w
        if kind == "conv_in":
            fl += 2.0 * n * hw * cin * cout * 9
        elif kind == "res":
            fl += 2.0 * n * hw * (cin * cout * 9 + cout * cout * 9 + 2 * cout * cout * 3)
            if cin != cout:
                fl += 2.0 * n * hw * cin * cout
        elif kind == "attn":
            fl += 2.0 * n * hw * cin * cin * 4 + 4.0 * n * hw * hw * cin
        elif kind == "up":
            h, w = 2 * h, 2 * w
            fl += 2.0 * n * h * w * cin * cout * 9
        elif kind == "out":
            fl += 2.0 * n * hw * cin * cout * 9 + 2.0 * n * hw * cout * cout * 3
    return fl


def clip_flops(unet_cfg, vae_cfg, T, H, W, steps, clips=1):
    """One clip: `steps` CFG-doubled UNet forwards (2*T frames) + one VAE decode of T frames."""
    return clips * (steps * unet_forward_flops(unet_cfg, 2 * T, H, W, T) + decoder_flops(vae_cfg, T, H, W))
