"""GPU parity tests of the individual C-ABI ops against plain torch fp32 math on the same (16-bit rounded) inputs."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from gcd_b200 import ops as o
    o.lib()
    return o


def rnd(*shape, scale=1.0, dtype=None, seed=[0]):
    seed[0] += 1
    g = torch.Generator().manual_seed(seed[0])
    t = torch.randn(*shape, generator=g) * scale
    return t.cuda() if dtype is None else t.to(dtype).cuda()


def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("M,K,N", [(300, 320, 320), (128, 64, 1280), (1000, 1280, 192), (28, 1280, 640), (513, 960, 2560)])
@pytest.mark.parametrize("out_f32", [True, False])
def test_linear_plain(ops, M, K, N, out_f32):
    AD = ops.act_dtype()
    x, w = rnd(M, K, dtype=AD), rnd(N, K, scale=K ** -0.5, dtype=AD)
    bias = rnd(N)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if out_f32 else AD)
    ops.linear(x, w, ops.make_ep(out, bias=bias))
    ref = x.float() @ w.float().t() + bias
    assert relerr(out, ref) < (2e-5 if out_f32 else 1e-3)


def test_linear_epilogue_full(ops):
    AD = ops.act_dtype()
    M, K, N = 700, 320, 320
    x, w = rnd(M, K, dtype=AD), rnd(N, K, scale=K ** -0.5, dtype=AD)
    bias, rv = rnd(N), rnd(7, N)
    r1, r2 = rnd(M, N), rnd(M, N, dtype=AD)
    out = torch.empty(M, N, device="cuda")
    ops.linear(x, w, ops.make_ep(out, bias=bias, rowvec=rv, rows_per_vec=100, res1=r1, a_res1=0.3, res2=r2, a_res2=0.6,
                                 a_acc=0.7, act=1))
    acc = x.float() @ w.float().t() + bias + rv.repeat_interleave(100, 0)
    ref = 0.7 * F.silu(acc) + 0.3 * r1 + 0.6 * r2.float()
    assert relerr(out, ref) < 2e-5
    # in-place residual update (out aliases res1)
    h = r1.clone()
    ops.linear(x, w, ops.make_ep(h, bias=bias, res1=h))
    assert relerr(h, x.float() @ w.float().t() + bias + r1) < 2e-5


def test_linear_geglu(ops):
    AD = ops.act_dtype()
    M, K, C4 = 384, 320, 1280
    x = rnd(M, K, dtype=AD)
    w = rnd(2 * C4, K, scale=K ** -0.5, dtype=AD)   # rows [0,C4) value, [C4, 2*C4) gate (attention.py:93)
    b = rnd(2 * C4)
    idx = torch.arange(2 * C4).view(2, C4 // 16, 16).permute(1, 0, 2).reshape(-1).cuda()   # 16 value / 16 gate blocks
    out = torch.empty(M, C4, device="cuda", dtype=AD)
    ops.linear(x, w[idx].contiguous(), ops.make_ep(out, bias=b[idx].contiguous(), geglu=True))
    y = x.float() @ w.float().t() + b
    ref = y[:, :C4] * F.gelu(y[:, C4:])
    assert relerr(out, ref) < 1.5e-3


def test_linear_strided_small_n(ops):
    AD = ops.act_dtype()
    M, K = 257, 320
    big = rnd(M, 3 * K, dtype=AD)
    x = big[:, K:2 * K]                       # strided rows
    w = rnd(4, K, scale=K ** -0.5, dtype=AD)
    out = torch.zeros(M, 4, device="cuda")
    ops.linear(x, w, ops.make_ep(out))
    assert relerr(out, x.float() @ w.float().t()) < 2e-5


@pytest.mark.parametrize("n,H,W,C,Co", [(3, 18, 32, 64, 128), (2, 9, 16, 128, 160), (2, 8, 128, 64, 64),
                                        (5, 4, 6, 192, 320), (28, 2, 2, 64, 320), (1, 36, 64, 320, 640)])
@pytest.mark.parametrize("stride", [1, 2])
def test_conv2d_3x3(ops, n, H, W, C, Co, stride):
    AD = ops.act_dtype()
    x = rnd(n, H, W, C, dtype=AD)
    w = rnd(Co, C, 3, 3, scale=(9 * C) ** -0.5, dtype=AD)
    bias = rnd(Co)
    wp = w.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous()
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = torch.empty(n * Ho * Wo, Co, device="cuda")
    ops.conv2d_3x3(x, wp, ops.make_ep(out, bias=bias), stride=stride)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=stride, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(n * Ho * Wo, Co)
    assert relerr(out, ref) < 3e-5


@pytest.mark.parametrize("n,H,W,C,res,f32out", [(3, 24, 64, 448, True, True), (2, 9, 40, 640, False, False), (5, 16, 128, 960, True, True)])
def test_conv2d_3x3_wide_tiles_n320(ops, n, H, W, C, res, f32out):
    """N = 320 with K = 9*C >= 3840: tc_gemm mode 4 (256 x 320 pair tiles, both 160-column halves fed from one activation tile,
    three rotating TMEM regions) incl. ragged M (phantom tiles), several passes per cluster, fp32 residual + fused GroupNorm
    statistics and 16-bit output."""
    AD = ops.act_dtype()
    Co = 320
    x = rnd(n, H, W, C, dtype=AD)
    w = rnd(Co, C, 3, 3, scale=(9 * C) ** -0.5, dtype=AD)
    bias = rnd(Co)
    wp = w.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous()
    rows = n * H * W
    r0 = rnd(rows, Co)
    out = r0.clone() if f32out else torch.empty(rows, Co, device="cuda", dtype=AD)
    st = torch.zeros(max(n, 64) * 64, device="cuda", dtype=torch.float64)
    ok = ops.conv2d_3x3(x, wp, ops.make_ep(out, bias=bias, res1=out if res else None, a_res1=0.5 if res else 1.0,
                                             gn_stats=(st, Co // 32, 32, H * W)))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(rows, Co)
    if res:
        ref = ref + 0.5 * r0
    assert relerr(out, ref) < (3e-5 if f32out else 1e-3)
    if ok:
        v = out.double().view(n, H * W, 32, Co // 32)
        want = torch.stack([v.sum((1, 3)), (v * v).sum((1, 3))], -1).reshape(-1)
        assert torch.allclose(st[: n * 64], want, rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("B,T,HW,C,Co", [(2, 14, 144, 128, 128), (1, 14, 4, 64, 320), (2, 14, 512, 320, 320), (3, 5, 96, 64, 64)])
def test_conv_t3(ops, B, T, HW, C, Co):
    AD = ops.act_dtype()
    x = rnd(B, T, HW, C, dtype=AD)
    w = rnd(Co, C, 3, 1, 1, scale=(3 * C) ** -0.5, dtype=AD)
    bias = rnd(Co)
    wp = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(Co, 3 * C).contiguous()
    out = torch.empty(B * T * HW, Co, device="cuda")
    ops.conv_t3(x, wp, ops.make_ep(out, bias=bias))
    ref = F.conv3d(x.float().permute(0, 3, 1, 2).unsqueeze(-1), w.float(), bias, padding=(1, 0, 0))  # b c t hw 1
    ref = ref.squeeze(-1).permute(0, 2, 3, 1).reshape(B * T * HW, Co)
    assert relerr(out, ref) < 3e-5


def test_bmm_nt(ops):
    AD = ops.act_dtype()
    G, M, K, N = 3, 200, 512, 300
    a, b = rnd(G, M, K, dtype=AD), rnd(G, N, K, scale=K ** -0.5, dtype=AD)
    out = torch.empty(G * M, N, device="cuda")
    ops.bmm_nt(a, b, ops.make_ep(out))
    ref = torch.bmm(a.float(), b.float().transpose(1, 2)).reshape(G * M, N)
    assert relerr(out, ref) < 2e-5


@pytest.mark.parametrize("n_img,rows,C,f32", [(4, 9216, 320, True), (2, 14 * 64, 640, True), (3, 100, 960, False),
                                              (2, 333, 128, True), (1, 50, 2560, True), (2, 77, 512, False)])
@pytest.mark.parametrize("silu", [False, True])
def test_groupnorm(ops, n_img, rows, C, f32, silu):
    AD = ops.act_dtype()
    x = rnd(n_img * rows, C) * 2 + 0.5
    if not f32:
        x = x.to(AD)
    g, b = rnd(C) * 0.1 + 1, rnd(C) * 0.1
    out = torch.empty(n_img * rows, C, device="cuda", dtype=AD)
    stats = torch.empty(n_img * 64, device="cuda", dtype=torch.float64)
    ops.groupnorm(x, n_img, rows, C, g, b, 1e-5, silu, out, stats)
    xr = x.float().view(n_img, rows, C).permute(0, 2, 1)
    ref = F.group_norm(xr, 32, g, b, 1e-5)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(n_img * rows, C)
    assert relerr(out, ref) < 1e-3


@pytest.mark.parametrize("rows,C", [(1000, 320), (77, 640), (4032, 1280)])
def test_layernorm(ops, rows, C):
    AD = ops.act_dtype()
    x = rnd(rows, C) * 3 + 1
    g, b = rnd(C) * 0.1 + 1, rnd(C) * 0.1
    out = torch.empty(rows, C, device="cuda", dtype=AD)
    ops.layernorm(x, g, b, out)
    assert relerr(out, F.layer_norm(x, (C,), g, b, 1e-5)) < 1e-3
    # with frame-indexed additive embedding and sum output
    T, per = 7, 11
    add = rnd(T, C)
    so = torch.empty_like(x)
    ops.layernorm(x, g, b, out, add=add, add_rows_per=per, add_mod=T, sum_out=so)
    idx = (torch.arange(rows, device="cuda") // per) % T
    xs = x + add[idx]
    assert torch.equal(so, xs)
    assert relerr(out, F.layer_norm(xs, (C,), g, b, 1e-5)) < 1e-3


def test_softmax_rows(ops):
    AD = ops.act_dtype()
    x = rnd(300, 9216) * 5
    out = torch.empty(300, 9216, device="cuda", dtype=AD)
    ops.softmax_rows(x, 0.044, out)
    assert relerr(out, torch.softmax(x * 0.044, -1)) < 2e-3


@pytest.mark.parametrize("clips,T,tokens,heads", [(2, 14, 144, 5), (1, 14, 37, 20), (1, 25, 16, 5)])
def test_attention_temporal(ops, clips, T, tokens, heads):
    AD = ops.act_dtype()
    C = heads * 64
    qkv = rnd(clips * T, tokens, 3 * C, dtype=AD)
    out = torch.empty(clips * T, tokens, C, device="cuda", dtype=AD)
    ops.attention_temporal(qkv, clips, T, tokens, heads, out)
    q, k, v = qkv.float().view(clips, T, tokens, 3, heads, 64).permute(3, 0, 2, 4, 1, 5)   # [clips, tokens, heads, T, 64]
    ref = F.scaled_dot_product_attention(q, k, v)                                          # over T
    ref = ref.permute(0, 3, 1, 2, 4).reshape(clips * T, tokens, C)
    assert relerr(out, ref) < 2e-3


def test_elementwise(ops):
    AD = ops.act_dtype()
    x = rnd(2, 3, 5, 64)
    o = torch.empty(2, 6, 10, 64, device="cuda", dtype=AD)
    ops.upsample2x_to_act(x, 2, 3, 5, 64, o)
    ref = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(o, ref.to(AD))
    a, b = rnd(33, 64), rnd(33, 128)
    c = torch.empty(33, 192, device="cuda")
    ops.concat_channels(a, b, c)
    assert torch.equal(c, torch.cat([a, b], 1))
    t = torch.tensor([0.5756, -1.2, 3.0], device="cuda")
    e = torch.empty(3, 320, device="cuda")
    ops.timestep_embedding(t, 320, out_f32=e)
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half).cuda()
    args = t[:, None] * freqs[None]
    assert (e - torch.cat([args.cos(), args.sin()], -1)).abs().max() < 2e-6


def test_sampler_kernels(ops):
    AD = ops.act_dtype()
    BT, T, H, W = 6, 3, 8, 16
    x, cc = rnd(BT, 4, H, W) * 10, rnd(BT, 4, H, W)
    sigma, sigma_next = 7.3, 5.1
    c_in = 1 / (sigma ** 2 + 1) ** 0.5
    buf = torch.empty(2 * BT, H, W, 64, device="cuda", dtype=AD)
    ops.sampler_prep(x, None, cc, BT, H, W, c_in, buf)
    ref = torch.zeros(2 * BT, 64, H, W, device="cuda")
    ref[:BT, :4] = x * c_in
    ref[BT:, :4] = x * c_in
    ref[BT:, 4:8] = cc
    assert torch.equal(buf, ref.permute(0, 2, 3, 1).to(AD))
    net = rnd(2 * BT, H, W, 16)
    scale = torch.linspace(1.0, 1.5, T).cuda()
    c_skip, c_out = 1 / (sigma ** 2 + 1), -sigma / (sigma ** 2 + 1) ** 0.5
    xn = x.clone()
    ops.sampler_update(xn, net, 16, BT, T, H, W, c_out, c_skip, sigma, sigma_next - sigma, scale)
    n4 = net[..., :4].permute(0, 3, 1, 2)
    den = n4 * c_out + torch.cat([x, x]) * c_skip
    du, dc = den[:BT], den[BT:]
    sc = scale.repeat(BT // T).view(BT, 1, 1, 1)
    d = du + sc * (dc - du)
    ref = x + (sigma_next - sigma) * (x - d) / sigma
    assert (xn - ref).abs().max() < 1e-4 * ref.abs().max()


def _attn_ref(qkv, frames, tokens, heads):
    """softmax(q k^T / 8) v in fp64-accumulating fp32 torch math, chunked over queries (attention.py:332-336 semantics)."""
    C = heads * 64
    q, k, v = qkv.float().view(frames, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    outs = []
    for i in range(0, tokens, 2048):
        s = torch.matmul(q[:, :, i:i + 2048], k.transpose(-1, -2)) * 0.125
        outs.append(torch.matmul(torch.softmax(s, -1), v))
    return torch.cat(outs, 2).permute(0, 2, 1, 3).reshape(frames, tokens, C)


# 4096 / 9216 tokens: the level-1 shape bench.py runs at latent 72x128; 4100: ragged last key block; scale 6: logits ~ +-100,
# forces lazy rescales. (GCD_FA_EMU=4 — every 4th exponential pair on the FMA pipe — is an experiment path, off by default.)
@pytest.mark.parametrize("frames,tokens,heads,scale", [
    (2, 256, 5, 1.5), (3, 144, 20, 1.5), (1, 576, 10, 1.5), (2, 100, 5, 1.5), (1, 2304, 5, 1.5), (2, 4, 5, 1.5),
    (2, 4096, 5, 1.5), (2, 9216, 5, 1.5), (1, 4100, 5, 1.5), (1, 9216, 5, 6.0), (1, 2304, 10, 6.0)])
def test_attention_spatial(ops, frames, tokens, heads, scale):
    AD = ops.act_dtype()
    C = heads * 64
    qkv = rnd(frames, tokens, 3 * C, dtype=AD, scale=scale)
    out = torch.empty(frames, tokens, C, device="cuda", dtype=AD)
    ops.attention_spatial(qkv, frames, tokens, heads, out)
    ref = _attn_ref(qkv, frames, tokens, heads)
    assert relerr(out, ref) < 3e-3


@pytest.mark.parametrize("n_img,Co,temporal", [(2, 128, False), (1, 128, True), (1, 512, False)])
def test_fused_groupnorm_stats_accumulated(ops, n_img, Co, temporal):
    """Images spanning >= 4 x num_SMs tiles take the per-CTA accumulation path (smem table, flush on image change) —
    the shape class of the VAE decoder's full-resolution convolutions (one clip or 14 frames over 64 512 tiles)."""
    AD = ops.act_dtype()
    C = 64
    if temporal:                                   # Conv3d (3,1,1) over [B=1, T=4, HW, C]; statistics over the whole clip
        T, HW = 4, 48 * 512
        x = rnd(1, T, HW, C, dtype=AD)
        wp = rnd(Co, 3 * C, scale=(3 * C) ** -0.5, dtype=AD)
        rows, rpi = T * HW, T * HW
        run = lambda ep: ops.conv_t3(x, wp, ep)
    else:
        H, W = (192, 512) if n_img == 2 else (256, 512)
        x = rnd(n_img, H, W, C, dtype=AD)
        wp = rnd(Co, 9 * C, scale=(9 * C) ** -0.5, dtype=AD)
        rows, rpi = n_img * H * W, H * W
        run = lambda ep: ops.conv2d_3x3(x, wp, ep)
    assert rpi // 128 >= 4 * torch.cuda.get_device_properties(0).multi_processor_count
    out = torch.empty(rows, Co, device="cuda", dtype=AD)
    st = torch.zeros(n_img * 64, device="cuda", dtype=torch.float64)
    assert run(ops.make_ep(out, bias=rnd(Co), gn_stats=(st, Co // 32, 32, rpi)))
    v = out.double().view(n_img, rpi, 32, Co // 32)
    ref = torch.stack([v.sum((1, 3)), (v * v).sum((1, 3))], -1).reshape(-1)
    assert torch.allclose(st, ref, rtol=1e-6, atol=1e-2)
    st2 = torch.zeros_like(st)                     # order-insensitive fp64 accumulation: repeats bit-for-bit (to fp64 rounding)
    assert run(ops.make_ep(out, bias=rnd(Co) * 0 , gn_stats=(st2, Co // 32, 32, rpi)))


@pytest.mark.parametrize("out_f32", [True, False])
def test_fused_groupnorm_stats(ops, out_f32):
    """gcd_epilogue.gn_stats: the conv epilogue accumulates (sum, sumsq) per (image, group) of the values it stores."""
    AD = ops.act_dtype()
    n, H, W, C, Co = 3, 8, 128, 64, 320
    x = rnd(n, H, W, C, dtype=AD)
    w = rnd(Co, C, 3, 3, scale=(9 * C) ** -0.5, dtype=AD)
    wp = w.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous()
    bias = rnd(Co)
    out = torch.empty(n * H * W, Co, device="cuda", dtype=torch.float32 if out_f32 else AD)
    st = torch.zeros(n * 32 * 2, device="cuda", dtype=torch.float64)
    ok = ops.conv2d_3x3(x, wp, ops.make_ep(out, bias=bias, gn_stats=(st, Co // 32, 32, H * W)))
    assert ok
    v = out.double().view(n, H * W, 32, Co // 32)
    ref = torch.stack([v.sum((1, 3)), (v * v).sum((1, 3))], -1).reshape(-1)
    assert torch.allclose(st, ref, rtol=1e-5, atol=1e-3)
    # consumer: GroupNorm with the fused statistics == GroupNorm with its own statistics pass
    g, b = rnd(Co) * 0.1 + 1, rnd(Co) * 0.1
    y1 = torch.empty(n * H * W, Co, device="cuda", dtype=AD)
    y2 = torch.empty_like(y1)
    ops.groupnorm(out, n, H * W, Co, g, b, 1e-5, True, y1, st, have_stats=True)
    ops.groupnorm(out, n, H * W, Co, g, b, 1e-5, True, y2, torch.empty(n * 64, device="cuda", dtype=torch.float64))
    assert relerr(y1, y2) < 1e-5
    # a tiling that cannot guarantee one image per tile reports "not produced" instead of wrong statistics
    x2 = rnd(4, 4, 6, 64, dtype=AD)
    out2 = torch.empty(4 * 4 * 6, Co, device="cuda")
    assert not ops.conv2d_3x3(x2, wp, ops.make_ep(out2, bias=bias, gn_stats=(st, Co // 32, 32, 24)))


def test_conditioner_embedders_vs_reference(ops):
    """ConcatTimestepEmbedderND and SphericalEmbedder (one kernel each) vs the reference classes' outputs."""
    from gcd_b200.embedders import ConcatTimestepEmbedderND, SphericalEmbedder
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "embedders.pt"))
    emb = ConcatTimestepEmbedderND(256)
    for k in ("1", "2"):
        y = emb(gold["concat_x" + k].cuda())
        assert y.shape == gold["concat_y" + k].shape and y.dtype == torch.float32
        # arguments reach 127 * 1 = 127 rad: cosf/sinf of the device vs the host libm agree to a few ulp of the argument
        assert (y.cpu() - gold["concat_y" + k]).abs().max() < 2e-5
    sph = SphericalEmbedder(128)
    sph.load_state_dict({"proj.weight": gold["sph_w"], "proj.bias": gold["sph_b"]})
    sph = sph.cuda()
    y = sph(gold["sph_x"].cuda())
    assert y.shape == gold["sph_y"].shape
    assert (y.cpu() - gold["sph_y"]).abs().max() < 1e-5 * max(1.0, float(gold["sph_y"].abs().max()))
    assert sph(gold["sph_x"].cuda().view(2, 14, 3)).shape == (2, 14, 128)


@pytest.mark.parametrize("n,rows,Ca,Cb", [(28, 9216, 640, 320), (28, 9216, 320, 320), (28, 144, 1280, 1280), (3, 100, 64, 192)])
def test_concat_channels_with_fused_groupnorm_stats(ops, n, rows, Ca, Cb):
    """Skip concat of the UNet's output blocks: one pass writes cat(a, b) and the (image, group) sums for the next GroupNorm."""
    a, b = rnd(n * rows, Ca) + 0.3, rnd(n * rows, Cb) * 2.0
    out = torch.empty(n * rows, Ca + Cb, device="cuda")
    st = torch.zeros(max(n, 64) * 64, device="cuda", dtype=torch.float64)
    ops.concat_channels(a, b, out, stats=st, n_img=n)
    ref = torch.cat([a, b], 1)
    assert torch.equal(out, ref)
    v = ref.double().view(n, rows, 32, (Ca + Cb) // 32)
    want = torch.stack([v.sum((1, 3)), (v * v).sum((1, 3))], -1).reshape(-1)
    assert torch.allclose(st[: n * 64], want, rtol=2e-6, atol=1e-3)
    st2 = torch.zeros_like(st)
    ops.concat_channels(a, b, out, stats=st2, n_img=n)
    assert torch.equal(st, st2)                       # fixed summation order inside a block; fp64 atomics across blocks
    # consumer equivalence: GroupNorm with these statistics == GroupNorm with its own statistics pass
    g, be = rnd(Ca + Cb) * 0.1 + 1, rnd(Ca + Cb) * 0.1
    AD = ops.act_dtype()
    y1 = torch.empty(n * rows, Ca + Cb, device="cuda", dtype=AD)
    y2 = torch.empty_like(y1)
    ops.groupnorm(out, n, rows, Ca + Cb, g, be, 1e-5, True, y1, st, have_stats=True)
    ops.groupnorm(out, n, rows, Ca + Cb, g, be, 1e-5, True, y2, torch.empty_like(st))
    assert relerr(y1, y2) < 1e-5
    out2 = torch.empty_like(out)
    ops.concat_channels(a, b, out2)                   # plain variant unchanged
    assert torch.equal(out2, ref)
    # 16-bit-only concat (what the UNet's output ResBlocks consume): the rounded values, statistics of the fp32 inputs
    out16 = torch.empty(n * rows, Ca + Cb, device="cuda", dtype=AD)
    st3 = torch.zeros_like(st)
    ops.concat_channels(a, b, out16, stats=st3, n_img=n)
    assert torch.equal(out16, ref.to(AD)) and torch.equal(st3, st)
