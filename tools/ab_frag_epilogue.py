import os, sys, contextlib, io, torch
sys.path.insert(0, "/root/repo")
from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet
with contextlib.redirect_stdout(io.StringIO()):
    m = UNet(img_channels=1).cuda().eval()
prog = m.program(64, (240, 240), 1, 0)
def run():
    prog._arena[:max(prog._arena_used, 4)].zero_()
    for fn in prog.ops: fn()
for _ in range(3): run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
print("FRAG", os.environ.get("MRI_GEMM_FRAG_EPI", "1"), "cfg2 forward %.2f ms" % (a.elapsed_time(b) / 10))
