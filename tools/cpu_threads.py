import time, torch, os, sys
sys.path.insert(0,'.')
from gcd_b200 import spec, synthetic, flops
from oracle import gcd_oracle as O
cfg=spec.UNET_KUBRIC
sd=synthetic.seeded_state(spec.unet_param_shapes(cfg))
h,w=8,16
x,c,uc,ioi=synthetic.seeded_inputs(cfg,1,14,h,w)
xin=torch.cat((torch.cat([x,x]),torch.cat((uc["concat"],c["concat"]))),1)
ctx=torch.cat((uc["crossattn"],c["crossattn"])); y=torch.cat((uc["vector"],c["vector"])); t=torch.full((28,),0.57)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "OMP", os.environ.get("OMP_NUM_THREADS"))
for nt in (16, 32, 64, 128):
    torch.set_num_threads(nt)
    with torch.no_grad():
        t0=time.time(); O.unet_forward(sd,cfg,xin,t,ctx,y,14,ioi); dt=time.time()-t0
    print(nt,"threads:",round(dt,2),"s", flush=True)
    if dt > 40: break
