#!/usr/bin/env python
"""Per-op CUDA-event timing of one training step's backward launch list (GPU only).
  CFG=3d|2d|25d python tools/perop_train.py"""
import collections, contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_image_generation_b200 import ops  # noqa: E402
os.environ["MRI_NO_GRAPH"] = "1"
cfg = os.environ.get("CFG", "25d")
with contextlib.redirect_stdout(io.StringIO()):
    if cfg == "3d":
        from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention as U
        m = U(3, base_channels=128, channel_mults=(1, 2, 4), time_emb_dim=256).cuda().train()
        B = int(os.environ.get("B", "8")); prog = m.program(B, (40, 48, 40), training=True)
    elif cfg == "2d":
        from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet as U
        m = U(img_channels=1, base_channels=64, channel_mults=(1, 2, 4, 8), time_emb_dim=256).cuda().train()
        B = int(os.environ.get("B", "64")); prog = m.program(B, (240, 240), 1, 0, training=True)
    else:
        from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.unet import UNet as U
        m = U(in_channels=20, out_channels=4, base_channels=64, channel_mults=(1, 2, 4, 8), time_emb_dim=256).cuda().train()
        B = int(os.environ.get("B", "32")); prog = m.program(B, (192, 192), 4, 16, training=True)
prog.x_in.normal_(); prog.t_in.fill_(500)
def fwd():
    prog._arena[:max(prog._arena_used, 4)].zero_()
    ev = []
    for n, fn in zip(prog.op_names, prog.ops):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((n, a, b))
    torch.cuda.synchronize()
    return [(n, a.elapsed_time(b)) for n, a, b in ev]
def bwd():
    prog.dout_in.normal_()
    for chunk, used in prog._zero_each_bwd:
        ops.memset_zero(chunk, used)
    ev = []
    for n, fn in zip(prog.bwd_names, prog.bwd_ops):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((n, a, b))
    if prog._final:   # deferred parameter-gradient finalisers: one launch for the whole list
        from mri_image_generation_b200 import _lib
        tab = prog._final_table(0, len(prog.bwd_ops))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(_lib.load().mri_grad_finalize(tab[0].data_ptr(), tab[1], tab[2], _lib.current_stream_ptr()), "fin")
        b.record(); ev.append(("grad_finalize", a, b))
    torch.cuda.synchronize()
    return [(n, a.elapsed_time(b)) for n, a, b in ev]
for _ in range(2): fwd(); bwd()
f, r = fwd(), bwd()
for nm, lst in (("forward", f), ("backward", r)):
    agg = collections.defaultdict(float)
    for n, t in lst: agg[n.split(":")[0]] += t
    print(nm, "total %.2f ms" % sum(t for _, t in lst), {k: round(v, 2) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]})
print("slowest backward ops:")
for n, t in sorted(r, key=lambda kv: -kv[1])[:14]: print("  %-44s %8.3f ms" % (n, t))
print("slowest forward ops:")
for n, t in sorted(f, key=lambda kv: -kv[1])[:8]: print("  %-44s %8.3f ms" % (n, t))
