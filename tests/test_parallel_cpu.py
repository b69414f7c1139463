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
