"""Find the first layer whose output for clip 0 depends on whether clip 1 is in the batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcd_b200 import spec, synthetic
from gcd_b200.unet import VideoUNet
name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
cfg = spec.UNET_TINY if name == "tiny" else spec.UNET_KUBRIC
T, H, W = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (14, 72, 128)))
net = VideoUNet(**spec.unet_ctor_kwargs(cfg)); net.load_state_dict(synthetic.seeded_state(spec.unet_param_shapes(cfg))); net = net.cuda()
xa, ca, _, _ = synthetic.seeded_inputs(cfg, 1, T, H, W, seed=11)
xb, cb, _, _ = synthetic.seeded_inputs(cfg, 1, T, H, W, seed=12)
mk = lambda x, c: (torch.cat((x * 0.3, c["concat"]), 1).cuda(), c["crossattn"].cuda(), c["vector"].cuda())
(xa_, ctxa, ya), (xb_, ctxb, yb) = mk(xa, ca), mk(xb, cb)
eng = net.engine(torch.device("cuda", 0))
def run(x, ctx, y, B):
    eng.debug = []
    net(x, torch.full((B * T,), 0.57).cuda(), context=ctx, y=y, num_video_frames=T, image_only_indicator=torch.zeros(B, T).cuda())
    torch.cuda.synchronize(); d = eng.debug; eng.debug = None; return d
d2 = run(torch.cat((xa_, xb_)), torch.cat((ctxa, ctxb)), torch.cat((ya, yb)), 2)
d1 = run(xa_, ctxa, ya, 1)
for (p, kind, h2), (_, _, h1) in zip(d2, d1):
    a = h2[: h1.shape[0]].float(); b = h1.float()
    print(f"{((a-b).norm()/b.norm()).item():.3e}  {kind:8s} {p}  rows {tuple(h1.shape)}")
