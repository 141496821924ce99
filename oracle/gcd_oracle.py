"""CPU ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the gcd_b200 product path.

A plain-PyTorch, fp32, functional restatement of the reference's denoising hot path (basilevh/gcd @ 761569f,
gcd-model/sgm): VideoUNet forward, Denoiser + VScalingWithEDMcNoise, LinearPredictionGuider, EDMDiscretization,
EulerEDMSampler loop and the temporal VAE VideoDecoder. Each function cites the reference file:line it follows.
It consumes a state dict with the *reference's own parameter names* (SURVEY.md Appendix A).

PINNING: the reference ships no tests or golden vectors for this path. This restatement is pinned
against the reference's own modules imported in the build container (oracle/pin_against_reference.py, which also
writes the golden fixtures under tests/golden/); tests/test_host_cpu.py re-checks the oracle against those committed
fixtures everywhere (the reference itself cannot travel to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.
"""
import math

import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------------------- config
UNET_KUBRIC = dict(  # gcd-model/configs/infer_kubric.yaml:18-40
    in_channels=8, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
    channel_mult=[1, 2, 4, 4], num_head_channels=64, transformer_depth=1, context_dim=1024, adm_in_channels=768,
    aux_emb_dim=128)
UNET_PARDOM = dict(UNET_KUBRIC, aux_emb_dim=0)  # gcd-model/configs/infer_pardom.yaml
VAE_DECODER = dict(ch=128, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, z_channels=4)  # infer_kubric.yaml:151-164
VAE_ENCODER = dict(ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2, z_channels=4, in_channels=3, double_z=True)  # :83-94


def unet_plan(cfg):
    """Block layout of VideoUNet.__init__ (video_model.py:213-459): lists of (kind, prefix, cin, cout)."""
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


# ---------------------------------------------------------------------------------------------------------- pieces
def timestep_embedding(t, dim, max_period=10000):
    """diffusionmodules/util.py:207-231."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _lin(sd, p, x, bias=True):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"] if bias else None)


def _gn(sd, p, x, eps):
    """GroupNorm32 (util.py:274-276, eps 1e-5) / Normalize (attention.py:125-128, model.py:52-55, eps 1e-6)."""
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _mlp(sd, p, x):
    """Linear-SiLU-Linear (video_model.py:161-199 time_embed/label_emb, video_attention.py:216-221 time_pos_embed)."""
    return _lin(sd, p + ".2", F.silu(_lin(sd, p + ".0", x)))


def _alpha(mix_factor, image_only_indicator):
    """AlphaBlender.get_alpha, merge_strategy learned_with_images (util.py:341-356): [b, t]."""
    ind = image_only_indicator.bool()
    return torch.where(ind, torch.ones(1, 1), torch.sigmoid(mix_factor)[..., None])


def resblock(sd, p, x, emb, three_d=False):
    """ResBlock._forward (openaimodel.py:331-357). three_d: the VideoResBlock.time_stack variant
    (dims=3, kernel (3,1,1), exchange_temb_dims=True; x: b c t h w, emb: b t E)."""
    conv = (lambda q, h: F.conv3d(h, sd[q + ".weight"], sd[q + ".bias"], padding=(1, 0, 0))) if three_d else \
           (lambda q, h: F.conv2d(h, sd[q + ".weight"], sd[q + ".bias"], padding=1))
    h = conv(p + ".in_layers.2", F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5)))
    if (p + ".emb_layers.1.weight") in sd:
        emb_out = _lin(sd, p + ".emb_layers.1", F.silu(emb))
        while emb_out.dim() < h.dim():
            emb_out = emb_out[..., None]
        if three_d:
            emb_out = emb_out.transpose(1, 2)  # "b t c ... -> b c t ..."
        h = h + emb_out
    # else: skip_t_emb (temporal_ae.py:41): emb_out = zeros
    h = conv(p + ".out_layers.3", F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5)))
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def video_resblock(sd, p, x, emb, T, ioi):
    """VideoResBlock.forward (video_model.py:62-81)."""
    x = resblock(sd, p, x, emb)
    bt, c, hh, ww = x.shape
    x5 = x.view(bt // T, T, c, hh, ww).permute(0, 2, 1, 3, 4)
    xt = resblock(sd, p + ".time_stack", x5, emb.view(bt // T, T, -1), three_d=True)
    a = _alpha(sd[p + ".time_mixer.mix_factor"], ioi)[:, None, :, None, None]   # "b t -> b 1 t 1 1"
    out = a * x5 + (1.0 - a) * xt
    return out.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)


def attention(sd, p, x, context=None, heads=None):
    """CrossAttention.forward (attention.py:283-344): softmax(q k^T / sqrt(d)) v, heads outer in the channel dim."""
    ctx = x if context is None else context
    q, k, v = (F.linear(x, sd[p + ".to_q.weight"]), F.linear(ctx, sd[p + ".to_k.weight"]),
               F.linear(ctx, sd[p + ".to_v.weight"]))
    b, n, c = q.shape
    d = c // heads
    q, k, v = (t.view(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
    return _lin(sd, p + ".to_out.0", o)


def feedforward(sd, p, x):
    """FeedForward with GEGLU (attention.py:87-113)."""
    val, gate = _lin(sd, p + ".net.0.proj", x).chunk(2, dim=-1)
    return _lin(sd, p + ".net.2", val * F.gelu(gate))


def basic_transformer_block(sd, p, x, context, heads):
    """BasicTransformerBlock._forward (attention.py:551-572)."""
    x = attention(sd, p + ".attn1", _ln(sd, p + ".norm1", x), None, heads) + x
    x = attention(sd, p + ".attn2", _ln(sd, p + ".norm2", x), context, heads) + x
    return feedforward(sd, p + ".ff", _ln(sd, p + ".norm3", x)) + x


def video_transformer_block(sd, p, x, context, T, heads):
    """VideoTransformerBlock._forward (video_attention.py:109-140), ff_in=True, is_res=True."""
    B, S, C = x.shape
    x = x.view(B // T, T, S, C).transpose(1, 2).reshape(-1, T, C)          # (b t) s c -> (b s) t c
    x = feedforward(sd, p + ".ff_in", _ln(sd, p + ".norm_in", x)) + x
    x = attention(sd, p + ".attn1", _ln(sd, p + ".norm1", x), None, heads) + x
    x = attention(sd, p + ".attn2", _ln(sd, p + ".norm2", x), context, heads) + x
    x = feedforward(sd, p + ".ff", _ln(sd, p + ".norm3", x)) + x
    return x.view(B // T, S, T, C).transpose(1, 2).reshape(B, S, C)


def spatial_video_transformer(sd, p, x, context, T, ioi, depth=1):
    """SpatialVideoTransformer.forward (video_attention.py:230-301), use_linear=True, use_spatial_context=True."""
    bt, c, hh, ww = x.shape
    heads = c // 64
    x_in = x
    time_context = context[::T].repeat_interleave(hh * ww, dim=0)         # :249-253
    x = _gn(sd, p + ".norm", x, 1e-6)
    x = x.flatten(2).transpose(1, 2)                                       # b c h w -> b (h w) c
    x = _lin(sd, p + ".proj_in", x)
    frames = torch.arange(T).repeat(bt // T)
    emb = _mlp(sd, p + ".time_pos_embed", timestep_embedding(frames, c))[:, None, :]
    for i in range(depth):
        x = basic_transformer_block(sd, f"{p}.transformer_blocks.{i}", x, context, heads)
        x_mix = video_transformer_block(sd, f"{p}.time_stack.{i}", x + emb, time_context, T, heads)
        a = _alpha(sd[p + ".time_mixer.mix_factor"], ioi).reshape(-1)[:, None, None]   # "b t -> (b t) 1 1"
        x = a * x + (1.0 - a) * x_mix
    x = _lin(sd, p + ".proj_out", x)
    x = x.transpose(1, 2).reshape(bt, c, hh, ww)
    return x + x_in


def unet_forward(sd, cfg, x, timesteps, context, y, num_video_frames, image_only_indicator):
    """VideoUNet.forward (video_model.py:461-540)."""
    T = num_video_frames
    emb = _mlp(sd, "time_embed", timestep_embedding(timesteps, cfg["model_channels"]))
    adm = cfg["adm_in_channels"]
    emb = emb + _mlp(sd, "label_emb.0", y[..., :adm])
    if cfg["aux_emb_dim"] > 0:
        emb = emb + _mlp(sd, "aux_label_emb", y[..., adm:])
    inp, mid, out = unet_plan(cfg)

    def run(layers, h):
        for kind, p, cin, cout in layers:
            if kind == "conv_in":
                h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
            elif kind == "vrb":
                h = video_resblock(sd, p, h, emb, T, image_only_indicator)
            elif kind == "svt":
                h = spatial_video_transformer(sd, p, h, context, T, image_only_indicator, cfg["transformer_depth"])
            elif kind == "down":   # openaimodel.py:163-210
                h = F.conv2d(h, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
            elif kind == "up":     # openaimodel.py:110-160
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
        return h

    hs, h = [], x
    for layers in inp:
        h = run(layers, h)
        hs.append(h)
    h = run(mid, h)
    for layers in out:
        h = run(layers, torch.cat([h, hs.pop()], dim=1))
    h = F.silu(_gn(sd, "out.0", h, 1e-5))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


# ---------------------------------------------------------------------------------------------------------- sampler
def edm_sigmas(n, sigma_min=0.002, sigma_max=700.0, rho=7.0):
    """EDMDiscretization.get_sigmas + Discretization.__call__ (discretizer.py:28-39,17-21): fp32 CPU ops, append 0."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def vscaling_edm_cnoise(sigma):
    """VScalingWithEDMcNoise (denoiser_scaling.py:53-61) -> c_skip, c_out, c_in, c_noise."""
    c_skip = 1.0 / (sigma ** 2 + 1.0)
    c_out = -sigma / (sigma ** 2 + 1.0) ** 0.5
    c_in = 1.0 / (sigma ** 2 + 1.0) ** 0.5
    c_noise = 0.25 * sigma.log()
    return c_skip, c_out, c_in, c_noise


def denoise(network, x, sigma, cond, **extra):
    """Denoiser.forward (denoiser.py:23-49) with OpenAIWrapper.forward (wrappers.py:23-34) inlined:
    network(x_cat, c_noise, context, y, **extra)."""
    sig = sigma.view(-1, *([1] * (x.dim() - 1)))
    c_skip, c_out, c_in, c_noise = vscaling_edm_cnoise(sig)
    c_noise = c_noise.reshape(sigma.shape)
    xin = torch.cat((x * c_in, cond["concat"]), dim=1)
    return network(xin, c_noise, cond["crossattn"], cond["vector"], **extra) * c_out + x * c_skip


def guider_scale(num_frames, max_scale, min_scale=1.0):
    """LinearPredictionGuider.__init__ (guiders.py:60-77)."""
    return torch.linspace(min_scale, max_scale, num_frames).unsqueeze(0)


def euler_edm_sample(network, x, cond, uc, num_steps, num_frames, max_scale, min_scale=1.0, sigma_max=700.0,
                     return_all=False, **extra):
    """EDMSampler.__call__ / sampler_step with s_churn=0 (sampling.py:46-59,101-144), Euler (225-230),
    LinearPredictionGuider.prepare_inputs/__call__ (guiders.py:79-100), to_d (sampling_utils.py:34-35)."""
    sigmas = edm_sigmas(num_steps, sigma_max=sigma_max)
    scale = guider_scale(num_frames, max_scale, min_scale)
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    c_cat = {k: torch.cat((uc[k], cond[k]), 0) for k in ("vector", "crossattn", "concat")}
    traj = []
    for i in range(num_steps):
        sigma_hat = s_in * sigmas[i]          # gamma = 0
        next_sigma = s_in * sigmas[i + 1]
        den = denoise(network, torch.cat([x] * 2), torch.cat([sigma_hat] * 2), c_cat, **extra)
        x_u, x_c = den.chunk(2)
        bt = x_u.shape[0]
        x_u5 = x_u.view(bt // num_frames, num_frames, *x_u.shape[1:])
        x_c5 = x_c.view(bt // num_frames, num_frames, *x_c.shape[1:])
        sc = scale.view(1, num_frames, 1, 1, 1)
        denoised = (x_u5 + sc * (x_c5 - x_u5)).reshape(x_u.shape)
        sh = sigma_hat.view(-1, 1, 1, 1)
        d = (x - denoised) / sh
        dt = next_sigma.view(-1, 1, 1, 1) - sh
        x = x + dt * d
        if return_all:
            traj.append(x.clone())
    return (x, traj) if return_all else x


# ---------------------------------------------------------------------------------------------------------- VAE decoder
def vae_resblock(sd, p, x, T):
    """temporal_ae.VideoResBlock.forward (temporal_ae.py:64-83) over model.ResnetBlock.forward (model.py:127-151)."""
    h = F.conv2d(F.silu(_gn(sd, p + ".norm1", x, 1e-6)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(F.silu(_gn(sd, p + ".norm2", h, 1e-6)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    x = x + h
    bt, c, hh, ww = x.shape
    x5 = x.view(bt // T, T, c, hh, ww).permute(0, 2, 1, 3, 4)
    xt = resblock(sd, p + ".time_stack", x5, None, three_d=True)
    a = torch.sigmoid(sd[p + ".mix_factor"])
    out = a * xt + (1.0 - a) * x5                 # NOTE: alpha weighs the TEMPORAL branch here (temporal_ae.py:79-80)
    return out.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)


def vae_attn(sd, p, x):
    """AttnBlock (model.py:161-201): single head over h*w tokens, d = C."""
    b, c, hh, ww = x.shape
    h = _gn(sd, p + ".norm", x, 1e-6)
    q, k, v = (F.conv2d(h, sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"]).flatten(2).transpose(1, 2)[:, None]
               for n in ("q", "k", "v"))
    o = F.scaled_dot_product_attention(q, k, v)[:, 0].transpose(1, 2).reshape(b, c, hh, ww)
    return x + F.conv2d(o, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])


def decoder_forward(sd, cfg, z, timesteps):
    """VideoDecoder (time_mode conv-only) = Decoder.forward (model.py:715-748) with VideoResBlock / AE3DConv."""
    T = timesteps
    nres = len(cfg["ch_mult"])
    h = F.conv2d(z, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    h = vae_resblock(sd, "mid.block_1", h, T)
    h = vae_attn(sd, "mid.attn_1", h)
    h = vae_resblock(sd, "mid.block_2", h, T)
    for lvl in reversed(range(nres)):
        for i in range(cfg["num_res_blocks"] + 1):
            h = vae_resblock(sd, f"up.{lvl}.block.{i}", h, T)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"up.{lvl}.upsample.conv.weight"], sd[f"up.{lvl}.upsample.conv.bias"], padding=1)
    h = F.silu(_gn(sd, "norm_out", h, 1e-6))
    h = F.conv2d(h, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)          # AE3DConv (temporal_ae.py:86-107)
    bt, c, hh, ww = h.shape
    h5 = h.view(bt // T, T, c, hh, ww).permute(0, 2, 1, 3, 4)
    h5 = F.conv3d(h5, sd["conv_out.time_mix_conv.weight"], sd["conv_out.time_mix_conv.bias"], padding=(1, 0, 0))
    return h5.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)


def decode_first_stage(sd, cfg, z, T, scale_factor=0.18215):
    """DiffusionEngine.decode_first_stage (models/diffusion.py:233-251), one chunk of T frames, fp32."""
    return decoder_forward(sd, cfg, z / scale_factor, T)


# ---------------------------------------------------------------------------------------------------------- VAE encoder
# the conditioning-frame encoder that runs once per sample, immediately before the hot path.
def vae_resblock2d(sd, p, x):
    """model.ResnetBlock.forward with temb=None (diffusionmodules/model.py:127-151)."""
    h = F.conv2d(F.silu(_gn(sd, p + ".norm1", x, 1e-6)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(F.silu(_gn(sd, p + ".norm2", h, 1e-6)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def encoder_forward(sd, cfg, x):
    """Encoder.forward (model.py:576-601); Downsample = pad (0,1,0,1) + 3x3 stride-2 conv without padding (model.py:84-88)."""
    nres = len(cfg["ch_mult"])
    h = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    for lvl in range(nres):
        for i in range(cfg["num_res_blocks"]):
            h = vae_resblock2d(sd, f"down.{lvl}.block.{i}", h)
        if lvl != nres - 1:
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[f"down.{lvl}.downsample.conv.weight"],
                         sd[f"down.{lvl}.downsample.conv.bias"], stride=2)
    h = vae_resblock2d(sd, "mid.block_1", h)
    h = vae_attn(sd, "mid.attn_1", h)
    h = vae_resblock2d(sd, "mid.block_2", h)
    h = F.silu(_gn(sd, "norm_out", h, 1e-6))
    return F.conv2d(h, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)


def encode_cond_frames(sd, cfg, x, quant_weight, quant_bias, scale_factor=0.18215):
    """AutoencodingEngineLegacy.encode (models/autoencoder.py:493-513): quant_conv(encoder(x)), regularised by
    DiagonalGaussianRegularizer(sample=False) = the distribution's mode = the mean half of the channels
    (distributions/distributions.py:25-28,71-72), then VideoPredictionEmbedderWithEncoder's `vid *= scale_factor`
    (encoders/modules.py:1106)."""
    moments = F.conv2d(encoder_forward(sd, cfg, x), quant_weight, quant_bias)
    return moments[:, :cfg["z_channels"]] * scale_factor


# ---------------------------------------------------------------------------------------------------------- small embedders
def concat_timestep_embedder_nd(x, outdim):
    """ConcatTimestepEmbedderND.forward (encoders/modules.py:1008-1016): each scalar of x[b, d] -> timestep_embedding(outdim),
    "(b d) d2 -> b (d d2)"."""
    if x.ndim == 1:
        x = x[:, None]
    b, dims = x.shape
    return timestep_embedding(x.reshape(-1), outdim).reshape(b, dims * outdim)


def spherical_embedder(proj_weight, proj_bias, x):
    """SphericalEmbedder.forward (encoders/modules.py:255-287): cos/sin of 1x, 2x, 4x azimuth, the same for elevation, the raw
    radius (13 features), then Linear(13, embed_dim)."""
    az, el, rad = x[..., 0], x[..., 1], x[..., 2]
    feats = [f(a * m) for a in (az, el) for m in (1.0, 2.0, 4.0) for f in (torch.cos, torch.sin)]
    return F.linear(torch.stack(feats + [rad], dim=-1), proj_weight, proj_bias)
