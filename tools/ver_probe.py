import sys, contextlib, io
sys.path.insert(0, '/root/repo')
import torch
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
dev='cuda'
model = UNet3DModelWithAttention(3, base_channels=64, channel_mults=(1,2,4), time_emb_dim=64).to(dev).train()
with contextlib.redirect_stdout(io.StringIO()):
    diff = GaussianDiffusionLatent3D(model, 3, timesteps=1000).to(dev)
z = torch.randn(2,3,8,8,8, device=dev)
ps = list(model.parameters())
def vers(): return [p._version for p in ps]
for it in range(3):
    v0 = vers()
    t = torch.randint(1,1000,(2,),device=dev)
    for p in ps: p.grad=None
    loss = diff.p_losses(z,t,cond=None,min_snr_gamma=5.0)
    v1 = vers()
    loss.backward()
    v2 = vers()
    print(it, 'fwd bumped', sum(a!=b for a,b in zip(v0,v1)), 'bwd bumped', sum(a!=b for a,b in zip(v1,v2)))
prog = model.program(2,(8,8,8),training=True)
print('tracked', len(prog._params), 'changed?', prog.params_changed())
