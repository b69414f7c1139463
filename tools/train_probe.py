import sys, os, contextlib, io, time
sys.path.insert(0, '/root/repo')
import torch
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
dev='cuda'
torch.manual_seed(0)
model = UNet3DModelWithAttention(3, base_channels=128, channel_mults=(1,2,4), time_emb_dim=256, groups=8, num_heads=4).to(dev).train()
with contextlib.redirect_stdout(io.StringIO()):
    diff = GaussianDiffusionLatent3D(model, 3, timesteps=1000).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=2e-4)
B=8
z = torch.randn(B,3,40,48,40, device=dev)
def step(do_opt=True):
    t = torch.randint(1,1000,(B,),device=dev)
    opt.zero_grad(set_to_none=True)
    loss = diff.p_losses(z,t,cond=None,min_snr_gamma=5.0)
    loss.backward()
    if do_opt: opt.step()
def timeit(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n, (time.perf_counter()-t0)*1e3/n
print('full step', timeit(step))
print('no optimizer step (no param change -> no refresh)', timeit(lambda: step(False)))
prog = model.program(B,(40,48,40),training=True)
names = prog.bwd_names
import collections
cnt=collections.Counter(n.split(':')[0] for n in names)
print(cnt)
# time backward op classes with events
def time_ops(pred):
    evs=[]
    prog.forward(z, torch.randint(1,1000,(B,),device=dev))
    torch.cuda.synchronize()
    import torch as T
    T._foreach_zero_(prog._zero_each_bwd)
    tot=0.0
    for n,fn in zip(prog.bwd_names, prog.bwd_ops):
        if pred(n):
            a,b=T.cuda.Event(enable_timing=True),T.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); evs.append((a,b))
        else: fn()
    T.cuda.synchronize()
    return sum(a.elapsed_time(b) for a,b in evs)
prog.dout_in.normal_()
from mri_image_generation_b200 import ops
for cls in ['unpack','wgrad','gemm','gn_bwd','colsum','add']:
    print(cls, time_ops(lambda n: n.startswith(cls)))
# refresh cost
torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
t0=time.perf_counter(); e0.record()
for fn in prog.refresh: fn()
e1.record(); torch.cuda.synchronize(); print('refresh: gpu ms', e0.elapsed_time(e1), 'wall ms', (time.perf_counter()-t0)*1e3, 'n', len(prog.refresh))
e0.record(); opt.step(); e1.record(); torch.cuda.synchronize(); print('adam ms', e0.elapsed_time(e1))
print('---- phase timing (synchronised between phases) ----')
def phases(do_opt):
    res = collections.defaultdict(float)
    for it in range(6):
        t = torch.randint(1,1000,(B,),device=dev)
        opt.zero_grad(set_to_none=True)
        torch.cuda.synchronize(); t0=time.perf_counter()
        loss = diff.p_losses(z,t,cond=None,min_snr_gamma=5.0)
        torch.cuda.synchronize(); t1=time.perf_counter()
        loss.backward()
        torch.cuda.synchronize(); t2=time.perf_counter()
        if do_opt: opt.step()
        torch.cuda.synchronize(); t3=time.perf_counter()
        if it >= 2:
            res['fwd']+= (t1-t0)*1e3/4; res['bwd'] += (t2-t1)*1e3/4; res['opt'] += (t3-t2)*1e3/4
    return dict(res)
print('with opt', phases(True))
print('no opt  ', phases(False))
print('---- forward op classes (training program) ----')
tt = torch.randint(1,1000,(B,),device=dev)
prog.x_in.copy_(z); prog.t_in.copy_(tt)
def time_fwd():
    prog._arena[:max(prog._arena_used,4)].zero_()
    evs=[]
    for n,fn in zip(prog.op_names, prog.ops):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((n,a,b))
    torch.cuda.synchronize()
    agg=collections.defaultdict(float)
    for n,a,b in evs: agg[n.split(':')[0].split('.')[-1] if not n.startswith('gemm') else 'gemm'] += a.elapsed_time(b)
    return dict(agg)
time_fwd(); r=time_fwd(); print({k: round(v,2) for k,v in sorted(r.items(), key=lambda kv:-kv[1])[:12]}, 'total', round(sum(r.values()),2))
torch.cuda.synchronize(); t0=time.perf_counter(); prog.run(); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
print('prog.run(): cpu launch ms', (t1-t0)*1e3, 'until done ms', (t2-t0)*1e3)
t0=time.perf_counter(); out = model(z, tt); torch.cuda.synchronize(); t1=time.perf_counter(); print('model(z,t) training-mode call ms', (t1-t0)*1e3)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); out = model(z, tt); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
