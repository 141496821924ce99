"""Architecture description of the two networks on the hot path: block plan and the parameter names/shapes of the
reference checkpoints, so that the drop-in modules expose *exactly* the reference `state_dict` keys
(`model.diffusion_model.*`, `first_stage_model.decoder.*`; models/diffusion.py:191-219 loads with strict=False, which
would silently skip renamed keys). tests/test_spec_cpu.py checks these tables against key lists dumped from the
reference modules (tests/golden/*_keys.json).
"""
from collections import OrderedDict

UNET_KUBRIC = dict(  # gcd-model/configs/infer_kubric.yaml:18-40
    in_channels=8, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
    channel_mult=[1, 2, 4, 4], num_head_channels=64, transformer_depth=1, context_dim=1024, adm_in_channels=768,
    aux_emb_dim=128)
UNET_PARDOM = dict(UNET_KUBRIC, aux_emb_dim=0)                       # configs/infer_pardom.yaml
UNET_TINY = dict(UNET_KUBRIC, model_channels=64)                     # reduced width for fast tests
VAE_DECODER = dict(ch=128, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, z_channels=4)  # infer_kubric.yaml:151-164
VAE_TINY = dict(VAE_DECODER, ch=64)
# conditioning-frame encoder (AutoencoderKLModeOnly.encoder, infer_kubric.yaml:78-96) — SURVEY.md §8(f) rank 1
VAE_ENCODER = dict(ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2, z_channels=4, in_channels=3, double_z=True)
VAE_ENCODER_TINY = dict(VAE_ENCODER, ch=64)


def unet_plan(cfg):
    """Block layout built by VideoUNet.__init__ (video_model.py:213-459).
    Returns (input_blocks, middle_block, output_blocks); each block is a list of (kind, prefix, cin, cout) with
    kind in {conv_in, vrb (VideoResBlock), svt (SpatialVideoTransformer), down, up}."""
    mc, mult, nrb = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    inp = [[("conv_in", "input_blocks.0.0", cfg["in_channels"], mc)]]
    chans, ch, ds = [mc], mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            i = len(inp)
            layers = [("vrb", f"input_blocks.{i}.0", ch, m * mc)]
            ch = m * mc
            if ds in cfg["attention_resolutions"]:
                layers.append(("svt", f"input_blocks.{i}.1", ch, ch))
            inp.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            ds *= 2
            inp.append([("down", f"input_blocks.{len(inp)}.0", ch, ch)])
            chans.append(ch)
    mid = [("vrb", "middle_block.0", ch, ch), ("svt", "middle_block.1", ch, ch), ("vrb", "middle_block.2", ch, ch)]
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            j = len(out)
            layers = [("vrb", f"output_blocks.{j}.0", ch + ich, m * mc)]
            ch = m * mc
            if ds in cfg["attention_resolutions"]:
                layers.append(("svt", f"output_blocks.{j}.{len(layers)}", ch, ch))
            if level and i == nrb:
                ds //= 2
                layers.append(("up", f"output_blocks.{j}.{len(layers)}", ch, ch))
            out.append(layers)
    return inp, mid, out


def _wb(d, p, *wshape):
    d[p + ".weight"] = tuple(wshape)
    d[p + ".bias"] = (wshape[0],)


def _resblock3(d, p, c, emb_dim):
    """time_stack ResBlock(dims=3, kernel (3,1,1)) (video_model.py:42-55 / temporal_ae.py:32-45)."""
    _wb(d, p + ".in_layers.0", c)
    _wb(d, p + ".in_layers.2", c, c, 3, 1, 1)
    if emb_dim:
        _wb(d, p + ".emb_layers.1", c, emb_dim)
    _wb(d, p + ".out_layers.0", c)
    _wb(d, p + ".out_layers.3", c, c, 3, 1, 1)


def _attn(d, p, c, ctx):
    d[p + ".to_q.weight"] = (c, c)
    d[p + ".to_k.weight"] = (c, ctx)
    d[p + ".to_v.weight"] = (c, ctx)
    _wb(d, p + ".to_out.0", c, c)


def _ff(d, p, c):
    _wb(d, p + ".net.0.proj", 8 * c, c)
    _wb(d, p + ".net.2", c, 4 * c)


def unet_param_shapes(cfg):
    d = OrderedDict()
    mc = cfg["model_channels"]
    E = 4 * mc
    ctx = cfg["context_dim"]
    _wb(d, "time_embed.0", E, mc)
    _wb(d, "time_embed.2", E, E)
    _wb(d, "label_emb.0.0", E, cfg["adm_in_channels"])
    _wb(d, "label_emb.0.2", E, E)
    if cfg["aux_emb_dim"] > 0:
        _wb(d, "aux_label_emb.0", E, cfg["aux_emb_dim"])
        _wb(d, "aux_label_emb.2", E, E)
    inp, mid, out = unet_plan(cfg)
    for layers in inp + [mid] + out:
        for kind, p, cin, cout in layers:
            if kind == "conv_in":
                _wb(d, p, cout, cin, 3, 3)
            elif kind == "down":
                _wb(d, p + ".op", cout, cin, 3, 3)
            elif kind == "up":
                _wb(d, p + ".conv", cout, cin, 3, 3)
            elif kind == "vrb":
                _wb(d, p + ".in_layers.0", cin)
                _wb(d, p + ".in_layers.2", cout, cin, 3, 3)
                _wb(d, p + ".emb_layers.1", cout, E)
                _wb(d, p + ".out_layers.0", cout)
                _wb(d, p + ".out_layers.3", cout, cout, 3, 3)
                if cin != cout:
                    _wb(d, p + ".skip_connection", cout, cin, 1, 1)
                _resblock3(d, p + ".time_stack", cout, E)
                d[p + ".time_mixer.mix_factor"] = (1,)
            elif kind == "svt":
                c = cout
                _wb(d, p + ".norm", c)
                _wb(d, p + ".proj_in", c, c)
                for i in range(cfg["transformer_depth"]):
                    q = f"{p}.transformer_blocks.{i}"
                    _attn(d, q + ".attn1", c, c)
                    _ff(d, q + ".ff", c)
                    _attn(d, q + ".attn2", c, ctx)
                    for n in ("norm1", "norm2", "norm3"):
                        _wb(d, f"{q}.{n}", c)
                _wb(d, p + ".proj_out", c, c)
                for i in range(cfg["transformer_depth"]):
                    q = f"{p}.time_stack.{i}"
                    _wb(d, q + ".norm_in", c)
                    _ff(d, q + ".ff_in", c)
                    _attn(d, q + ".attn1", c, c)
                    _ff(d, q + ".ff", c)
                    _wb(d, q + ".norm2", c)
                    _attn(d, q + ".attn2", c, ctx)
                    _wb(d, q + ".norm1", c)
                    _wb(d, q + ".norm3", c)
                _wb(d, p + ".time_pos_embed.0", 4 * c, c)
                _wb(d, p + ".time_pos_embed.2", c, 4 * c)
                d[p + ".time_mixer.mix_factor"] = (1,)
    _wb(d, "out.0", mc)
    _wb(d, "out.2", cfg["out_channels"], mc, 3, 3)
    return d


def decoder_plan(cfg):
    """Decoder.__init__ (diffusionmodules/model.py:604-713): list of (kind, prefix, cin, cout) in execution order."""
    ch, mult, nrb = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    nres = len(mult)
    block_in = ch * mult[nres - 1]
    plan = [("conv_in", "conv_in", cfg["z_channels"], block_in), ("res", "mid.block_1", block_in, block_in),
            ("attn", "mid.attn_1", block_in, block_in), ("res", "mid.block_2", block_in, block_in)]
    for lvl in reversed(range(nres)):
        block_out = ch * mult[lvl]
        for i in range(nrb + 1):
            plan.append(("res", f"up.{lvl}.block.{i}", block_in, block_out))
            block_in = block_out
        if lvl != 0:
            plan.append(("up", f"up.{lvl}.upsample", block_in, block_in))
    plan.append(("out", "conv_out", block_in, cfg["out_ch"]))
    return plan


def decoder_param_shapes(cfg):
    d = OrderedDict()
    for kind, p, cin, cout in decoder_plan(cfg):
        if kind == "conv_in":
            _wb(d, p, cout, cin, 3, 3)
        elif kind == "res":
            d[p + ".mix_factor"] = (1,)
            _wb(d, p + ".norm1", cin)
            _wb(d, p + ".conv1", cout, cin, 3, 3)
            _wb(d, p + ".norm2", cout)
            _wb(d, p + ".conv2", cout, cout, 3, 3)
            if cin != cout:
                _wb(d, p + ".nin_shortcut", cout, cin, 1, 1)
            _resblock3(d, p + ".time_stack", cout, 0)
        elif kind == "attn":
            _wb(d, p + ".norm", cin)
            for n in ("q", "k", "v", "proj_out"):
                _wb(d, f"{p}.{n}", cin, cin, 1, 1)
        elif kind == "up":
            _wb(d, p + ".conv", cout, cin, 3, 3)
        elif kind == "out":
            _wb(d, "norm_out", cin)
            _wb(d, "conv_out", cout, cin, 3, 3)
            _wb(d, "conv_out.time_mix_conv", cout, cout, 3, 1, 1)
    return d


def encoder_plan(cfg):
    """Encoder.__init__/forward (diffusionmodules/model.py:487-601): (kind, prefix, cin, cout) in execution order."""
    ch, mult, nrb = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    plan = [("conv_in", "conv_in", cfg["in_channels"], ch)]
    block_in = ch
    for lvl in range(len(mult)):
        block_out = ch * mult[lvl]
        for i in range(nrb):
            plan.append(("res", f"down.{lvl}.block.{i}", block_in, block_out))
            block_in = block_out
        if lvl != len(mult) - 1:
            plan.append(("down", f"down.{lvl}.downsample", block_in, block_in))
    plan += [("res", "mid.block_1", block_in, block_in), ("attn", "mid.attn_1", block_in, block_in),
             ("res", "mid.block_2", block_in, block_in),
             ("out", "conv_out", block_in, (2 if cfg.get("double_z", True) else 1) * cfg["z_channels"])]
    return plan


def encoder_param_shapes(cfg):
    d = OrderedDict()
    for kind, p, cin, cout in encoder_plan(cfg):
        if kind == "conv_in":
            _wb(d, p, cout, cin, 3, 3)
        elif kind == "res":
            _wb(d, p + ".norm1", cin)
            _wb(d, p + ".conv1", cout, cin, 3, 3)
            _wb(d, p + ".norm2", cout)
            _wb(d, p + ".conv2", cout, cout, 3, 3)
            if cin != cout:
                _wb(d, p + ".nin_shortcut", cout, cin, 1, 1)
        elif kind == "attn":
            _wb(d, p + ".norm", cin)
            for n in ("q", "k", "v", "proj_out"):
                _wb(d, f"{p}.{n}", cin, cin, 1, 1)
        elif kind == "down":
            _wb(d, p + ".conv", cout, cin, 3, 3)
        elif kind == "out":
            _wb(d, "norm_out", cin)
            _wb(d, "conv_out", cout, cin, 3, 3)
    return d


def encoder_ctor_kwargs(cfg):
    """Constructor kwargs of the VAE Encoder as in configs/infer_kubric.yaml:83-94 (ddconfig)."""
    return dict(attn_type="vanilla", double_z=cfg.get("double_z", True), z_channels=cfg["z_channels"], resolution=256,
                in_channels=cfg["in_channels"], out_ch=3, ch=cfg["ch"], ch_mult=cfg["ch_mult"],
                num_res_blocks=cfg["num_res_blocks"], attn_resolutions=[], dropout=0.0)


def unet_ctor_kwargs(cfg):
    """Constructor kwargs of VideoUNet as written in the GCD YAML (configs/infer_kubric.yaml:21-40)."""
    return dict(adm_in_channels=cfg["adm_in_channels"], num_classes="sequential", use_checkpoint=True,
                in_channels=cfg["in_channels"], out_channels=cfg["out_channels"], model_channels=cfg["model_channels"],
                attention_resolutions=cfg["attention_resolutions"], num_res_blocks=cfg["num_res_blocks"],
                channel_mult=cfg["channel_mult"], num_head_channels=cfg["num_head_channels"],
                use_linear_in_transformer=True, transformer_depth=cfg["transformer_depth"], context_dim=cfg["context_dim"],
                spatial_transformer_attn_type="softmax-xformers", extra_ff_mix_layer=True, use_spatial_context=True,
                merge_strategy="learned_with_images", video_kernel_size=[3, 1, 1], aux_emb_dim=cfg["aux_emb_dim"],
                aux_zero_init=False)


def decoder_ctor_kwargs(cfg):
    """Constructor kwargs of VideoDecoder as in configs/infer_kubric.yaml:152-164."""
    return dict(attn_type="vanilla", double_z=True, z_channels=cfg["z_channels"], resolution=256, in_channels=3,
                out_ch=cfg["out_ch"], ch=cfg["ch"], ch_mult=cfg["ch_mult"], num_res_blocks=cfg["num_res_blocks"],
                attn_resolutions=[], dropout=0.0, video_kernel_size=[3, 1, 1])
