"""Class-conditional CFG sampling of DiMSUM on the velocity ODE, sharded by batch across ranks.

Reference: dimsum/sample_ddp.py:52-191 (per-rank seed, CFG batching :168-173, `sample_fn(z, model.forward_with_cfg, ...)`
:178), dimsum/transport/transport.py:181-183 (`velocity_ode`: the drift is the model output) and
dimsum/transport/integrators.py:76-111 (`torchdiffeq.odeint` on `linspace(t0, t1, num_steps)`).  torchdiffeq is not
available offline; its fixed-grid `euler` method is restated: x_{i+1} = x_i + (t_{i+1} - t_i) f(t_i, x_i).  The adaptive
`dopri5` of the released scripts cannot be reproduced without the package and is not offered.

Multi-GPU: every latent is independent, so ranks take contiguous slices of the batch (CFG pairs stay together), run
with no communication, and exchange only the final latents with ONE all_gather (the reference writes PNGs per rank and
only barriers, sample_ddp.py:187-191).
"""
import torch
import torch.distributed as dist


def euler_velocity_ode(drift, x, num_steps=250, t0=0.0, t1=1.0):
    """drift(x, t_vec) -> dx/dt.  Fixed grid linspace(t0, t1, num_steps): num_steps - 1 evaluations."""
    ts = torch.linspace(t0, t1, num_steps, device=x.device)
    ones = torch.ones(x.shape[0], device=x.device)
    for i in range(num_steps - 1):
        x = x + (ts[i + 1] - ts[i]) * drift(x, ones * ts[i])
    return x


@torch.no_grad()
def sample_cfg(model, z, y, cfg_scale=4.0, num_steps=250, null_class=None, use_graph=False):
    """z (n, C, H, W) noise, y (n,) labels -> (n, C, H, W) latents.  CFG doubles the rows like sample_ddp.py:168-173."""
    null_class = model.num_classes if null_class is None else null_class
    x = torch.cat([z, z], dim=0)
    yy = torch.cat([y, torch.full_like(y, null_class)], dim=0)
    if use_graph and x.is_cuda:
        # one graph replay per grid point: model forward + fused CFG combine + Euler update (x stays in the graph's buffer)
        step = GraphedCfgStep(model, x, yy, cfg_scale)
        ts = torch.linspace(0.0, 1.0, num_steps, device=x.device)
        ones = torch.ones(x.shape[0], device=x.device)
        for i in range(num_steps - 1):
            x = step.euler(x, ones * ts[i], ts[i + 1] - ts[i])
        return x[: len(z)].clone()
    drift = lambda xx, tt: model.forward_with_cfg(xx, tt, yy, cfg_scale=cfg_scale)
    x = euler_velocity_ode(drift, x, num_steps)
    return x[: len(z)]


class GraphedCfgStep:
    """One CFG denoising evaluation captured as a CUDA graph and replayed: the model forward on the doubled batch plus, for fp32
    latents, ONE kernel of this repo for the guidance combine and the Euler update (`euler()`); `__call__` returns the drift
    v = model.forward_with_cfg(x, t, y).

    The ~3000 kernel launches of a DiM-L/2 forward cost more host time than device time once the per-rank batch is
    small (8-GPU sharding), so the sampler replays a captured graph: static input buffers, one `cudaGraphLaunch` per
    evaluation.  The C-ABI launches go to torch's current stream, which is the capturing stream inside
    `torch.cuda.graph`, so this repo's kernels are captured like any other."""

    def __init__(self, model, x_example, y, cfg_scale, autocast_dtype=None):
        self.x = x_example.clone()
        self.t = torch.zeros(x_example.shape[0], device=x_example.device)
        self.y = y.clone()
        amp = lambda: torch.autocast("cuda", dtype=autocast_dtype or torch.bfloat16, enabled=autocast_dtype is not None)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad(), amp():
            for _ in range(2):
                model.forward_with_cfg(self.x, self.t, self.y, cfg_scale=cfg_scale)
        torch.cuda.current_stream().wait_stream(side)
        from . import _lib
        before = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        self.dt = torch.zeros((), device=x_example.device)
        self.x_new = torch.empty_like(self.x)
        fused_ok = (self.x.dtype == torch.float32 and self.x.dim() == 4 and not getattr(model, "learn_sigma", False)
                    and self.x[0].numel() % 4 == 0)
        with torch.cuda.graph(self.graph), torch.no_grad(), amp():
            if fused_ok:
                # forward_with_cfg (models_dim.py:1886-1902) = forward on [half, half] + the guidance combine; the combine and
                # the Euler update run as ONE kernel of this repo inside the graph (dimsum_cfg_euler_step)
                from . import fused
                half = self.x[: len(self.x) // 2]
                out = model.forward(torch.cat([half, half], dim=0), self.t, self.y)
                self.v = torch.empty_like(self.x)
                fused.cfg_euler_step(self.x, out, cfg_scale, self.dt, out=self.x_new, v_out=self.v)
            else:
                self.v = model.forward_with_cfg(self.x, self.t, self.y, cfg_scale=cfg_scale)
        self.fused = fused_ok
        self.cfg_scale = cfg_scale
        self.model = model
        self.launches_per_replay = _lib.launch_count() - before

    def __call__(self, x, t):
        """-> v = forward_with_cfg(x, t, y) (the drift)."""
        self.x.copy_(x, non_blocking=True)
        self.t.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.v

    def euler(self, x, t, dt):
        """-> x + dt * forward_with_cfg(x, t, y): one graph replay (the returned tensor is the graph's buffer, valid until the
        next replay)."""
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        self.t.copy_(t, non_blocking=True)
        if not self.fused:
            self.graph.replay()
            return x + dt * self.v.float()
        if torch.is_tensor(dt):
            self.dt.copy_(dt, non_blocking=True)
        else:
            self.dt.fill_(float(dt))
        self.graph.replay()
        return self.x_new                            # the graph reads self.x and writes self.x_new (static buffers)


def shard_batch(n_total, rank, world):
    """Contiguous slice [lo, hi) of the batch owned by `rank`."""
    per = (n_total + world - 1) // world
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


@torch.no_grad()
def sample_cfg_sharded(model, z_all, y_all, cfg_scale=4.0, num_steps=250):
    """Every rank passes the same (n_total, ...) noise / labels, computes its slice, and receives all latents."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sample_cfg(model, z_all, y_all, cfg_scale, num_steps)
    rank, world = dist.get_rank(), dist.get_world_size()
    n = z_all.shape[0]
    if n % world:
        raise RuntimeError("sample_cfg_sharded: batch must be divisible by the world size")
    lo, hi = shard_batch(n, rank, world)
    mine = sample_cfg(model, z_all[lo:hi], y_all[lo:hi], cfg_scale, num_steps).contiguous()
    out = torch.empty((world,) + tuple(mine.shape), device=mine.device, dtype=mine.dtype)
    dist.all_gather_into_tensor(out.view(-1), mine.view(-1))
    return out.view((n,) + tuple(mine.shape[1:]))
