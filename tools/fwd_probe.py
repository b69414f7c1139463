import sys, contextlib, io, time
sys.path.insert(0, '/root/repo')
import torch
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
from mri_image_generation_b200 import ops
dev='cuda'
torch.manual_seed(0)
model = UNet3DModelWithAttention(3, base_channels=128, channel_mults=(1,2,4), time_emb_dim=256).to(dev).train()
with contextlib.redirect_stdout(io.StringIO()):
    diff = GaussianDiffusionLatent3D(model, 3, timesteps=1000).to(dev)
B=8
z = torch.randn(B,3,40,48,40, device=dev)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t = torch.randint(1,1000,(B,),device=dev)
    for p in model.parameters(): p.grad=None
    t0=sync(); noise = torch.randn_like(z); xn = diff.q_sample(z, t, noise); t1=sync()
    prog = model.program(B,(40,48,40),training=True)
    ch = prog.params_changed(); t2=sync()
    prog.x_in.copy_(xn); prog.t_in.copy_(t); t3=sync()
    prog.run(); t4=sync()
    S=40*48*40
    ops.nhwc_to_nchw(prog.eps_nhwc, prog.out, B, S, prog.cout, prog.cout_pad); t5=sync()
    pred = model(xn, t); t6=sync()
    loss = diff._loss(pred, noise, t, 5.0); t7=sync()
    loss.backward(); t8=sync()
    print(it, 'q_sample %.2f changed=%s %.2f copy %.2f run %.2f nhwc %.2f | model() %.2f loss %.2f bwd %.2f' % ((t1-t0)*1e3, ch, (t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3,(t5-t4)*1e3,(t6-t5)*1e3,(t7-t6)*1e3,(t8-t7)*1e3))
