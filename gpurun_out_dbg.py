import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import synthetic_state_dict, shapes_of
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
m=UNet3DModelWithAttention(in_channels=3, base_channels=64, time_emb_dim=64)
m.load_state_dict(synthetic_state_dict(shapes_of(m),7)); m=m.cuda().eval()
g=torch.Generator().manual_seed(3)
x=torch.randn(2,3,8,8,8,generator=g).cuda(); t=torch.randint(0,1000,(2,),generator=g).cuda()
prog=m.program(2,(8,8,8))
prog.x_in.copy_(x); prog.t_in.copy_(t)
tr=[prog.run_traced() for _ in range(3)]
for i,(name,outs) in enumerate(tr[0]):
    for r in (1,2):
        for a,b in zip(outs,tr[r][i][1]):
            d=(a.float()-b.float()).abs().max().item()
            if d>0:
                rel=((a.float()-b.float()).norm()/a.float().norm()).item()
                print(f"op {i:3d} {name:32s} run0 vs run{r}: maxabs {d:.3e} rel {rel:.3e} shape {tuple(a.shape)} nan {torch.isnan(a.float()).any().item()}")
print("boxes:", [(pl.name, pl.box, pl.tiles, pl.block_n, pl._args.stages) for pl in prog.plans][:12])
