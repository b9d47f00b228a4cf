"""CPU, world_size 2, gloo: the batch-sharding + single all_gather of the sampler (dimsum_b200/sampler.py) with a stand-in
model (the host-side logic has no CUDA dependency; the kernels themselves are covered by the -m gpu tests)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Toy(torch.nn.Module):
    num_classes = 10
    in_channels = 4

    def forward_with_cfg(self, x, t, y, cfg_scale=1.0):
        half = x[: len(x) // 2]
        cond = torch.tanh(half) * (1 + y[: len(half)].view(-1, 1, 1, 1).float()) * t[: len(half)].view(-1, 1, 1, 1)
        uncond = torch.tanh(half) * 0.5
        g = uncond + cfg_scale * (cond - uncond)
        return torch.cat([g, g], dim=0)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dimsum_b200.sampler import sample_cfg, sample_cfg_sharded, shard_batch
    g = torch.Generator().manual_seed(0)
    z = torch.randn(8, 4, 8, 8, generator=g)
    y = torch.randint(0, 10, (8,), generator=g)
    full = sample_cfg(_Toy(), z, y, cfg_scale=4.0, num_steps=12)
    got = sample_cfg_sharded(_Toy(), z, y, cfg_scale=4.0, num_steps=12)
    lo, hi = shard_batch(8, rank, world)
    ok = torch.allclose(got, full, atol=1e-6) and (hi - lo) == 4
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_sampling_matches_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_euler_grid_matches_closed_form():
    from dimsum_b200.sampler import euler_velocity_ode
    x = euler_velocity_ode(lambda xx, tt: -xx, torch.ones(3, 2), num_steps=250)
    assert torch.allclose(x, torch.full((3, 2), (1 - 1 / 249) ** 249), atol=1e-6)
