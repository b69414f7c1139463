"""Host logic of the sharded paths on CPU with the gloo backend, world_size 2."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT
from mri_image_generation_b200.parallel import shard_bounds


def test_shard_bounds_cover_everything_once():
    for total in (0, 1, 7, 8, 155, 1000):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(total, world, r)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi))
            assert seen == list(range(total))
            sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


class _StubDiffusion:
    """sample() = seeded noise + per-sample z (no GPU): checks seeding, slicing and gathering."""

    def sample(self, batch_size, size, z_pos=None):
        x = torch.randn(batch_size, 1, size, size)
        if z_pos is not None:
            x = x + z_pos.view(-1, 1, 1, 1)
        return x


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mri_image_generation_b200.parallel import sample_sharded
    z = torch.arange(5, dtype=torch.float32)
    got = sample_sharded(_StubDiffusion(), 5, 4, base_seed=77, per_sample_kwargs={"z_pos": z})
    local = sample_sharded(_StubDiffusion(), 5, 4, base_seed=77, gather=False,
                           per_sample_kwargs={"z_pos": z})
    torch.save({"gathered": got, "local": local}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sample_sharded_gloo_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    assert r1["gathered"] is None
    # expected: rank r draws its share under seed 77 + r
    z = torch.arange(5, dtype=torch.float32)
    parts = []
    for r, (lo, hi) in enumerate([(0, 3), (3, 5)]):
        torch.manual_seed(77 + r)
        parts.append(torch.randn(hi - lo, 1, 4, 4) + z[lo:hi].view(-1, 1, 1, 1))
    assert torch.equal(r0["local"], parts[0]) and torch.equal(r1["local"], parts[1])
    assert torch.equal(r0["gathered"], torch.cat(parts, 0))


# ------------------------------------------------------------------ overlapped gradient all-reduce
class _FakeProgram:
    """The bookkeeping half of backward.BackwardMixin over CPU tensors: 'records' whose ops write
    rank-dependent values into parameter gradients, one of them touching a parameter that an
    earlier record already touched (as the three q/k/v slices of mid_attn.qkv.weight do)."""

    def __new__(cls, rank):
        from mri_image_generation_b200.backward import BackwardMixin

        class Prog(BackwardMixin):
            def _replay(self, key, body):
                self.replayed.append(key)
                body()

            def _bwd_head(self, S):
                self.head_calls += 1

        pr = Prog()
        pr._binit()
        pr.device = "cpu"
        pr.replayed, pr.head_calls = [], 0
        pr.dout_in = torch.zeros(1)
        g = torch.Generator().manual_seed(5)
        pr.shapes = [(3, 5), (70,), (128, 9), (1,), (64, 64), (200,), (33, 3), (512,)]
        pr._params = [torch.zeros(s) for s in pr.shapes]
        pr.values = {}

        def record(idxs, scale):
            for i in idxs:
                p = pr._params[i]
                gbuf = pr.pg(p)
                val = torch.randn(p.shape, generator=g)
                pr.values[i] = val * scale

                def op(gbuf=gbuf, val=val, scale=scale):
                    gbuf.copy_(val * scale * (rank + 1))   # rank r contributes (r + 1) * value
                pr.badd(f"w{i}", op)
            pr.badd("filler", lambda: None)
            pr._close_record()

        record([0, 1], 1.0)
        record([2], 1.0)
        record([3, 4], 1.0)
        record([2], 2.0)          # parameter 2 is written again: its first value must not be reduced
        record([5], 1.0)
        record([6, 7], 1.0)
        blk = pr.pg_block([torch.zeros(4, 6), torch.zeros(2, 6)], 6)   # untracked -> arena slack
        pr.values["blk"] = torch.arange(36, dtype=torch.float32).view(6, 6)
        pr.badd("blk", lambda: blk.copy_(pr.values["blk"] * (rank + 1)))
        pr._close_record()
        return pr


def test_plan_segments_prefix_property():
    pr = _FakeProgram(0)
    segs = pr.plan_segments(bucket_bytes=256)
    assert segs[0][0] == 0 and segs[-1][1] == len(pr.bwd_ops)
    assert segs[0][2] == 0 and segs[-1][3] == pr._garena_used
    for (lo, hi, a, b), (lo2, hi2, a2, b2) in zip(segs, segs[1:]):
        assert hi == lo2 and b == a2 and lo < hi and a < b
    assert len(segs) >= 3
    # every parameter inside a segment's arena slice is final once the segment's ops have run
    for (lo, hi, a, b) in segs:
        for pid, (off, n) in pr._g_off.items():
            if off < b:
                assert pr._g_last_op[pid] <= hi
    # parameter 2 (re-written by a later record) delays the cut that contains it
    off2 = pr._g_off[id(pr._params[2])][0]
    seg2 = next(s for s in segs if s[2] <= off2 < s[3])
    assert seg2[1] >= pr._g_last_op[id(pr._params[2])]
    # one huge bucket -> a single segment
    assert len(pr.plan_segments(bucket_bytes=1 << 30)) == 1


def _sync_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mri_image_generation_b200.parallel import GradSync
    pr = _FakeProgram(rank)
    sync = GradSync(bucket_cap_mb=256 / (1 << 20))
    out = {}
    for step in range(2):
        pr._backward(torch.zeros(1), 1, sync)
        out[f"arena{step}"] = pr.garena[:pr._garena_used].clone()
    out["buckets"] = list(sync.buckets_last_step)
    out["replayed"] = list(pr.replayed)
    out["off"] = {i: pr._g_off[id(p)] for i, p in enumerate(pr._params)}
    out["values"] = {k: v for k, v in pr.values.items()}
    out["blk_off"] = min(o for pid, (o, n) in pr._g_off.items() if pid not in {id(p) for p in pr._params})
    sync.enabled = False      # no_sync(): plain local backward, one un-segmented launch list
    pr._backward(torch.zeros(1), 1, sync)
    out["local"] = pr.garena[:pr._garena_used].clone()
    # gradient accumulation: what the no_sync() step left in .grad is reduced by the next
    # synchronised backward (torch DDP semantics)
    assert sync._pending_local
    ps = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))]
    ps[0].grad = torch.full((3, 2), float(rank + 1))
    ps[1].grad = torch.arange(5, dtype=torch.float32) * (rank + 1)
    sync.enabled = True
    sync.reduce_pending(ps)
    out["accum"] = [p.grad.clone() for p in ps]
    assert not sync._pending_local
    torch.save(out, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_sync_gloo_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_sync_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["arena0"], r1["arena0"]) and torch.equal(r0["arena1"], r0["arena0"])
    # mean over ranks of (r + 1) * value = 1.5 * value, for the LAST value written to each parameter
    for i, (off, n) in r0["off"].items():
        want = (1.5 * r0["values"][i]).reshape(-1)
        assert torch.allclose(r0["arena0"][off:off + n], want, rtol=0, atol=1e-6), i
    bo = r0["blk_off"]
    assert torch.allclose(r0["arena0"][bo:bo + 36], 1.5 * r0["values"]["blk"].reshape(-1))
    # buckets tile the arena in order and there are several of them
    b = r0["buckets"]
    assert len(b) >= 3 and b[0][0] == 0 and all(x[1] == y[0] for x, y in zip(b, b[1:]))
    assert b[-1][1] == r0["arena0"].numel()
    assert any(k.startswith("bwd_seg") for k in r0["replayed"])
    assert torch.equal(r0["accum"][0], torch.full((3, 2), 1.5))
    assert torch.equal(r1["accum"][1], torch.arange(5, dtype=torch.float32) * 1.5)
    # no_sync: rank-local gradients, whole-list launch
    for i, (off, n) in r1["off"].items():
        assert torch.allclose(r1["local"][off:off + n], (2.0 * r1["values"][i]).reshape(-1))


def test_plan_segments_prefix_property_random_programs():
    """Random backward programs (random parameter sizes, records that touch several parameters,
    parameters re-touched by later records): the buckets always tile the arena in order, and every
    gradient inside a bucket is final when the bucket's last op has run."""
    import random
    from mri_image_generation_b200.backward import BackwardMixin

    class Prog(BackwardMixin):
        pass

    rng = random.Random(7)
    for trial in range(40):
        pr = Prog()
        pr._binit()
        pr.device = "cpu"
        n_params = rng.randint(1, 14)
        pr._params = [torch.zeros(rng.choice([1, 3, 64, 65, 700, 4096])) for _ in range(n_params)]
        touched = set()
        for _rec in range(rng.randint(1, 20)):
            for i in rng.sample(range(n_params), rng.randint(1, min(3, n_params))):
                pr.pg(pr._params[i])
                touched.add(i)
                for _ in range(rng.randint(1, 3)):
                    pr.badd("op", lambda: None)
            pr._close_record()
        bucket = rng.choice([1, 256, 4096, 1 << 20])
        segs = pr.plan_segments(bucket)
        assert segs[0][0] == 0 and segs[0][2] == 0
        assert segs[-1][1] == len(pr.bwd_ops) and segs[-1][3] == pr._garena_used
        for (lo, hi, a, b), (lo2, hi2, a2, b2) in zip(segs, segs[1:]):
            assert hi == lo2 and b == a2 and lo < hi and a < b
        for (lo, hi, a, b) in segs:
            for pid, (off, n) in pr._g_off.items():
                if off < b:
                    assert pr._g_last_op[pid] <= hi, (trial, off, b, hi)
        # arena slices are disjoint, 256-byte aligned and cover exactly the touched parameters
        spans = sorted(pr._g_off.values())
        assert len(spans) == len(touched)
        for (o1, n1), (o2, _) in zip(spans, spans[1:]):
            assert o1 % 64 == 0 and o1 + n1 <= o2


def _ddp_factory_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torch.nn.parallel import DistributedDataParallel as TorchDDP

    from mri_image_generation_b200 import overlay, parallel
    # what `overlay --overlap-ddp` binds the scripts' `DDP(...)` to (ddpm_3d_ldm/train.py:232-233):
    # modules the overlapped wrapper does not know (the reference's own VAE, any plain nn.Module)
    # go to torch's wrapper with the script's call signature
    DDP = parallel.ddp_for_scripts(TorchDDP)
    torch.manual_seed(rank)
    lin = torch.nn.Linear(4, 3)
    wrapped = DDP(lin, device_ids=None, output_device=None, find_unused_parameters=False)
    ok_type = isinstance(wrapped, TorchDDP)
    x = torch.full((2, 4), float(rank + 1))
    wrapped(x).sum().backward()
    g = lin.weight.grad.clone()
    # install / uninstall round trip of the module-level patch
    overlay.install(packages=["ddpm_3d_ldm"], overlap_ddp=True, fused_adam=False)
    import torch.nn.parallel as tnp
    patched = tnp.DistributedDataParallel is not TorchDDP
    overlay.uninstall()
    restored = tnp.DistributedDataParallel is TorchDDP
    torch.save({"ok_type": ok_type, "grad": g, "patched": patched, "restored": restored},
               os.path.join(out_dir, f"d{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_factory_for_scripts_hands_plain_modules_to_torch_ddp(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_ddp_factory_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "d0.pt")
    r1 = torch.load(tmp_path / "d1.pt")
    assert r0["ok_type"] and r1["ok_type"] and r0["patched"] and r0["restored"]
    # torch DDP averaged the two ranks' gradients: d/dW of sum(W x + b) = x summed over the batch
    want = torch.full((3, 4), 2.0 * (1 + 2) / 2)
    assert torch.allclose(r0["grad"], want) and torch.allclose(r1["grad"], want)
