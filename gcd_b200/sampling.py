"""Sampler-side plugin classes: drop-ins for the `target:` strings of the GCD configs
(configs/infer_kubric.yaml:13-16,113-127), same constructor and call contracts as the reference:

    EulerEDMSampler          <- sgm.modules.diffusionmodules.sampling.EulerEDMSampler      (sampling.py:26-144,225-230)
    EDMDiscretization        <- ...discretizer.EDMDiscretization                           (discretizer.py:17-39)
    LinearPredictionGuider   <- ...guiders.LinearPredictionGuider                          (guiders.py:60-100)
    Denoiser                 <- ...denoiser.Denoiser                                       (denoiser.py:11-49)
    VScalingWithEDMcNoise    <- ...denoiser_scaling.VScalingWithEDMcNoise                  (denoiser_scaling.py:53-61)
    OpenAIWrapper            <- ...wrappers.OpenAIWrapper                                  (wrappers.py:23-34)

Two execution paths, identical arithmetic:
  * generic  — the reference control flow (`denoiser(*guider.prepare_inputs(...))`, works with ANY denoiser closure);
    tensor glue is torch, the network call lands in gcd_b200.VideoUNet.forward (CUDA kernels).
  * fused    — taken when the closure is recognised as DiffusionEngine.sample_video's (models/diffusion.py:531-532)
    around a gcd_b200.VideoUNet + VScalingWithEDMcNoise + LinearPredictionGuider: the per-step input build
    (c_in scale, concat, CFG doubling, channels-last fp16) and the update (c_out/c_skip, CFG combine, to_d, Euler)
    each run as ONE kernel (csrc/elem.cu), and the UNet consumes/produces channels-last buffers directly.
The scalar schedule (sigmas, c_*, dt) is always computed with the same torch fp32 ops as the reference.
"""
import importlib
import warnings

import torch
import torch.nn as nn

from . import ops

# reference target string -> local class name, used when `sgm` itself is not importable
_ALIASES = {
    "sgm.modules.diffusionmodules.discretizer.EDMDiscretization": "EDMDiscretization",
    "sgm.modules.diffusionmodules.guiders.LinearPredictionGuider": "LinearPredictionGuider",
    "sgm.modules.diffusionmodules.denoiser_scaling.VScalingWithEDMcNoise": "VScalingWithEDMcNoise",
    "sgm.modules.diffusionmodules.denoiser.Denoiser": "Denoiser",
    "sgm.modules.diffusionmodules.sampling.EulerEDMSampler": "EulerEDMSampler",
}


def instantiate_from_config(config):
    """sgm/util.py:168-185. `sgm.*` targets that have a local equivalent resolve locally if sgm cannot be imported."""
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    target = config["target"]
    params = dict(config.get("params", dict()) or {})
    module, cls = target.rsplit(".", 1)
    try:
        return getattr(importlib.import_module(module), cls)(**params)
    except ImportError:
        if target in _ALIASES:
            return globals()[_ALIASES[target]](**params)
        raise


def append_dims(x, target_dims):
    return x[(...,) + (None,) * (target_dims - x.ndim)]


class EDMDiscretization:
    """discretizer.py:17-39 (Karras rho-schedule, zero appended by __call__)."""

    def __init__(self, sigma_min=0.002, sigma_max=80.0, rho=7.0):
        self.sigma_min, self.sigma_max, self.rho = sigma_min, sigma_max, rho

    def get_sigmas(self, n, device="cpu"):
        ramp = torch.linspace(0, 1, n, device=device)
        min_inv_rho = self.sigma_min ** (1 / self.rho)
        max_inv_rho = self.sigma_max ** (1 / self.rho)
        return (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** self.rho

    def __call__(self, n, do_append_zero=True, device="cpu", flip=False):
        sigmas = self.get_sigmas(n, device=device)
        sigmas = torch.cat([sigmas, sigmas.new_zeros([1])]) if do_append_zero else sigmas
        return sigmas if not flip else torch.flip(sigmas, (0,))


class VScalingWithEDMcNoise:
    """denoiser_scaling.py:53-61."""

    def __call__(self, sigma):
        c_skip = 1.0 / (sigma ** 2 + 1.0)
        c_out = -sigma / (sigma ** 2 + 1.0) ** 0.5
        c_in = 1.0 / (sigma ** 2 + 1.0) ** 0.5
        c_noise = 0.25 * sigma.log()
        return c_skip, c_out, c_in, c_noise


class Denoiser(nn.Module):
    """denoiser.py:11-49."""

    def __init__(self, scaling_config):
        super().__init__()
        self.scaling = instantiate_from_config(scaling_config)

    def forward(self, network, input, sigma, cond, **additional_model_inputs):
        sigma_shape = sigma.shape
        sigma = append_dims(sigma, input.ndim)
        c_skip, c_out, c_in, c_noise = self.scaling(sigma)
        c_noise = c_noise.reshape(sigma_shape)
        return network(input * c_in, c_noise, cond, **additional_model_inputs) * c_out + input * c_skip


class OpenAIWrapper(nn.Module):
    """wrappers.py:10-34 (compile_model is accepted and ignored: there is no tracing compiler on this path)."""

    def __init__(self, diffusion_model, compile_model=False):
        super().__init__()
        self.diffusion_model = diffusion_model

    def forward(self, x, t, c, **kwargs):
        x = torch.cat((x, c.get("concat", torch.Tensor([]).type_as(x))), dim=1)
        return self.diffusion_model(x, timesteps=t, context=c.get("crossattn", None), y=c.get("vector", None), **kwargs)


class LinearPredictionGuider:
    """guiders.py:60-100. NOTE: `scale` is fixed at construction; later writes to max_scale/min_scale/num_frames
    (scripts/eval_utils.py:169-172) do not refresh it — same as the reference."""

    def __init__(self, max_scale, num_frames, min_scale=1.0, additional_cond_keys=None):
        self.min_scale, self.max_scale, self.num_frames = min_scale, max_scale, num_frames
        self.scale = torch.linspace(min_scale, max_scale, num_frames).unsqueeze(0)
        if additional_cond_keys is None:
            additional_cond_keys = []
        if isinstance(additional_cond_keys, str):
            additional_cond_keys = [additional_cond_keys]
        self.additional_cond_keys = additional_cond_keys

    def __call__(self, x, sigma):
        x_u, x_c = x.chunk(2)
        T = self.num_frames
        x_u = x_u.reshape(-1, T, *x_u.shape[1:])
        x_c = x_c.reshape(-1, T, *x_c.shape[1:])
        scale = append_dims(self.scale.repeat(x_u.shape[0], 1), x_u.ndim).to(x_u.device)
        out = x_u + scale * (x_c - x_u)
        return out.reshape(-1, *out.shape[2:])

    def prepare_inputs(self, x, s, c, uc):
        c_out = dict()
        for k in c:
            if k in ["vector", "crossattn", "concat"] + self.additional_cond_keys:
                c_out[k] = torch.cat((uc[k], c[k]), 0)
            elif "hijack" not in k:
                assert c[k] == uc[k]
                c_out[k] = c[k]
        return torch.cat([x] * 2), torch.cat([s] * 2), c_out


class FusedDenoiser:
    """Explicit handle for the fused path: what DiffusionEngine.sample_video's closure captures, made visible.
    Calling it behaves exactly like that closure (generic path)."""

    def __init__(self, denoiser, model, **additional_model_inputs):
        self.denoiser, self.model, self.extra = denoiser, model, additional_model_inputs

    def __call__(self, input, sigma, c):
        return self.denoiser(self.model, input, sigma, c, **self.extra)


def _unwrap_closure(fn):
    """Recognise `lambda input, sigma, c: self.denoiser(self.model, input, sigma, c, **additional_model_inputs)`
    (models/diffusion.py:531-532). Returns (denoiser_module, wrapped_model, extra) or None."""
    if isinstance(fn, FusedDenoiser):
        return fn.denoiser, fn.model, fn.extra
    cells = getattr(fn, "__closure__", None)
    names = getattr(getattr(fn, "__code__", None), "co_freevars", ())
    if not cells:
        return None
    env = {}
    for n, c in zip(names, cells):
        try:
            env[n] = c.cell_contents
        except ValueError:
            return None
    eng, extra = env.get("self"), env.get("additional_model_inputs")
    if eng is None or not isinstance(extra, dict) or not hasattr(eng, "denoiser") or not hasattr(eng, "model"):
        return None
    return eng.denoiser, eng.model, extra


class EulerEDMSampler:
    """sampling.py:26-144,225-230 (BaseDiffusionSampler + EDMSampler + EulerEDMSampler)."""
    _warned_generic = False

    def __init__(self, discretization_config, num_steps=None, guider_config=None, verbose=False, device="cuda",
                 s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0):
        self.num_steps = num_steps
        self.discretization = instantiate_from_config(discretization_config)
        if guider_config is None:
            raise NotImplementedError("gcd_b200.EulerEDMSampler needs a guider_config (GCD uses LinearPredictionGuider)")
        self.guider = instantiate_from_config(guider_config)
        self.verbose, self.device = verbose, device
        self.s_churn, self.s_tmin, self.s_tmax, self.s_noise = s_churn, s_tmin, s_tmax, s_noise
        self.last_path = None
        self.trace_steps, self.trace = (), {}   # tests: keep a copy of x after these step counts (fused path), e.g. {1, 5, 10}
        self.cfg_group = None            # set_cfg_parallel(): 2-rank process group splitting the CFG pair of ONE clip

    def set_cfg_parallel(self, group):
        """Single-clip latency mode: the two ranks of `group` hold the same clip; rank 0 evaluates the
        unconditional half of the classifier-free-guidance batch (guiders.py:89-100 puts `uc` first), rank 1 the conditional
        half, and one all_gather of the network output per step (8 MB at 14x72x128) lets both apply the identical guided
        Euler update. `None` switches back. Only the fused CUDA path honours it."""
        if group is not None:
            import torch.distributed as dist
            if dist.get_world_size(group) != 2:
                raise ValueError("cfg-parallel needs a process group of exactly 2 ranks (uc | c)")
        self.cfg_group = group

    # ---- reference control flow -------------------------------------------------------------------------------
    def prepare_sampling_loop(self, x, cond, uc=None, num_steps=None):
        # The schedule is evaluated on the HOST with the reference's fp32 expressions and shipped to the device: torch's CUDA
        # `pow`/`linspace` need not round like the CPU ones, and the reference goldens (tests/golden/closed_forms.pt) are CPU
        # values — this keeps the scheduler bit-exact whatever `self.device` is (tests/test_pipeline_gpu.py asserts it).
        sigmas = self.discretization(self.num_steps if num_steps is None else num_steps, device="cpu").to(self.device)
        uc = cond if uc is None else uc
        x *= torch.sqrt(1.0 + sigmas[0] ** 2.0)
        return x, x.new_ones([x.shape[0]]), sigmas, len(sigmas), cond, uc

    def denoise(self, x, denoiser, sigma, cond, uc):
        denoised = denoiser(*self.guider.prepare_inputs(x, sigma, cond, uc))
        return self.guider(denoised, sigma)

    def sampler_step(self, sigma, next_sigma, denoiser, x, cond, uc=None, gamma=0.0):
        sigma_hat = sigma * (gamma + 1.0)
        if gamma > 0:
            eps = torch.randn_like(x) * self.s_noise
            x = x + eps * append_dims(sigma_hat ** 2 - sigma ** 2, x.ndim) ** 0.5
        denoised = self.denoise(x, denoiser, sigma_hat, cond, uc)
        d = (x - denoised) / append_dims(sigma_hat, x.ndim)          # sampling_utils.py:34-35
        dt = append_dims(next_sigma - sigma_hat, x.ndim)
        return x + dt * d                                            # euler_step; no correction step for Euler

    def _gamma(self, sigmas, i, num_sigmas):
        return min(self.s_churn / (num_sigmas - 1), 2 ** 0.5 - 1) if self.s_tmin <= sigmas[i] <= self.s_tmax else 0.0

    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        fused = self._fused_plan(denoiser, x, cond, uc)
        x, s_in, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        if fused is not None and self.s_churn == 0.0:
            self.last_path = "fused"
            return self._run_fused(fused, x, sigmas, cond, uc)
        self.last_path = "generic"
        if not EulerEDMSampler._warned_generic and x.is_cuda:
            EulerEDMSampler._warned_generic = True
            warnings.warn("gcd_b200.EulerEDMSampler: denoiser closure / guider / conditioning not recognised as the GCD "
                          "sample_video configuration -> generic path (reference control flow with torch glue per step; the "
                          "UNet still runs the CUDA kernels). Pass a gcd_b200.sampling.FusedDenoiser for the fused path.",
                          RuntimeWarning, stacklevel=2)
        for i in range(num_sigmas - 1):
            x = self.sampler_step(s_in * sigmas[i], s_in * sigmas[i + 1], denoiser, x, cond, uc,
                                  self._gamma(sigmas, i, num_sigmas))
        return x

    # ---- fused path ---------------------------------------------------------------------------------------------
    def _fused_plan(self, denoiser, x, cond, uc):
        from .unet import VideoUNet
        got = _unwrap_closure(denoiser)
        if got is None or not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 4:
            return None
        den, model, extra = got
        net = getattr(model, "diffusion_model", model)
        if not isinstance(net, VideoUNet) or type(getattr(den, "scaling", None)).__name__ != "VScalingWithEDMcNoise":
            return None
        if type(self.guider).__name__ != "LinearPredictionGuider" or getattr(self.guider, "additional_cond_keys", []):
            return None
        uc_ = cond if uc is None else uc
        if set(cond.keys()) != {"vector", "crossattn", "concat"} or "num_video_frames" not in extra:
            return None
        T = int(extra["num_video_frames"])
        ioi = extra.get("image_only_indicator")
        if T != self.guider.num_frames or x.shape[0] % T != 0 or (ioi is not None and bool((ioi != 0).any())):
            return None
        if self.guider.scale.numel() != T:      # num_frames mutated after construction (scripts/eval_utils.py:170): `scale` is
            return None                         # stale — the generic path then fails with the reference's own shape error
        if cond["concat"].shape != x.shape or uc_["concat"].shape != x.shape:
            return None
        return dict(net=net, scaling=den.scaling, T=T)

    @torch.no_grad()
    def _run_fused(self, plan, x, sigmas, cond, uc):
        net, T = plan["net"], plan["T"]
        BT, _, H, W = x.shape
        dev = x.device
        eng = net.engine(dev)
        x = x.contiguous()
        cc = cond["concat"].to(torch.float32).contiguous()
        ucc = uc["concat"].to(torch.float32).contiguous()
        # conditioning goes into pool buffers: fixed addresses let the captured CUDA graph of the forward serve every sample
        ctx_new = torch.cat((uc["crossattn"], cond["crossattn"]), 0)               # guiders.py:89-100
        y_new = torch.cat((uc["vector"], cond["vector"]), 0)
        ctx = eng.pool.get("ctx_static", tuple(ctx_new.shape), ctx_new.dtype)
        y = eng.pool.get("y_static", tuple(y_new.shape), y_new.dtype)
        ctx.copy_(ctx_new)
        y.copy_(y_new)
        scale = self.guider.scale.reshape(-1).to(dev, torch.float32).contiguous()
        # scalar schedule with the reference's own tensor ops (on the sigmas' device), read back once
        sig = sigmas.to(torch.float32)
        c_skip, c_out, c_in, c_noise = plan["scaling"](sig[:-1])
        dts = sig[1:] - sig[:-1]
        host = torch.stack([sig[:-1], c_skip, c_out, c_in, dts]).cpu().tolist()
        c_noise_dev = c_noise.to(dev)
        x_cl = eng.pool.get("x_cl", (2 * BT, H, W, 64), eng.AD)
        t_in = eng.pool.get("t_in", (2 * BT,), torch.float32)
        if self.cfg_group is not None:
            return self._run_fused_cfg_parallel(eng, x, ucc, cc, ctx, y, scale, host, c_noise_dev, x_cl, t_in, T)
        ca = eng.cross_attn_vectors(ctx, T, static=True)   # constant over the steps: computed once per sample
        for i in range(sig.numel() - 1):
            ops.sampler_prep(x, ucc, cc, BT, H, W, host[3][i], x_cl)
            t_in.copy_(c_noise_dev[i].expand(2 * BT))
            res = eng.forward_graphed(x_cl, 2 * BT, H, W, t_in, ctx, y, T, ca)
            ops.sampler_update(x, res, res.stride(0), BT, T, H, W, host[2][i], host[1][i], host[0][i], host[4][i], scale)
            if (i + 1) in self.trace_steps:
                self.trace[i + 1] = x.clone()
        return x

    def _run_fused_cfg_parallel(self, eng, x, ucc, cc, ctx, y, scale, host, c_noise_dev, x_cl, t_in, T):
        """This rank's half of the CFG batch ([uc | c] on the batch axis) through the UNet, all_gather of the two halves."""
        import torch.distributed as dist
        BT, _, H, W = x.shape
        r = dist.get_rank(self.cfg_group)
        half = slice(r * BT, (r + 1) * BT)
        ctx_h, y_h = ctx[half], y[half]              # contiguous views of the static pool buffers
        ca = eng.cross_attn_vectors(ctx_h, T, static=True)
        full = eng.pool.get("net_out_cfg", (2 * BT * H * W, 16), torch.float32)
        parts = list(full.chunk(2, 0))
        for i in range(len(host[0])):
            ops.sampler_prep(x, ucc, cc, BT, H, W, host[3][i], x_cl)
            t_in.copy_(c_noise_dev[i].expand(2 * BT))
            res = eng.forward_graphed(x_cl[half], BT, H, W, t_in[:BT], ctx_h, y_h, T, ca)
            dist.all_gather(parts, res, group=self.cfg_group)
            ops.sampler_update(x, full, full.stride(0), BT, T, H, W, host[2][i], host[1][i], host[0][i], host[4][i], scale)
        return x
