"""One launch of each glue kernel at its DiM-L/2 shape, for an ncu capture:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,\\
sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum \\
        --clock-control none -k regex:"colsum|norm_kernel|rmsnorm_bwd|rowwise" --csv python tools/microbench/glue_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dimsum_b200 import fused  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s, dt=torch.float32: torch.randn(*s, generator=g, device="cuda").to(dt)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

# training shapes (32 latents): bias gradient of w12, adaLN reductions, RMSNorm backward
gr = rn(32, 256, 8192, dt=torch.bfloat16)
flush.zero_(); fused.token_colsum(gr, out_dtype=torch.float32)
gr, x = rn(32, 256, 1024, dt=torch.bfloat16), rn(32, 256, 1024)
flush.zero_(); fused.token_colsum(gr, x, out_dtype=torch.float32)
xb = rn(32, 256, 1024, dt=torch.bfloat16).requires_grad_(True)
res = rn(32, 256, 1024).requires_grad_(True)
w = torch.ones(1024, device="cuda", requires_grad=True)
y, h = fused.add_rmsnorm_fn(xb, res, w, 1e-5, out_dtype=torch.float32)
flush.zero_(); torch.autograd.grad((y, h), (xb, res, w), (torch.randn_like(y), torch.randn_like(h)))
# sampling shapes (512 CFG rows, fp32)
xs, rs = rn(512, 256, 1024), rn(512, 256, 1024)
sh = rn(512, 2048)
flush.zero_(); fused.add_rmsnorm(xs, rs, w.detach(), 1e-5)
flush.zero_(); fused.norm_modulate(xs, rs, w.detach(), 1e-5, sh[:, :1024], sh[:, 1024:], want_residual=True)
xh = rn(512, 256, 512, dt=torch.bfloat16)
shb = rn(512, 1024, dt=torch.bfloat16)
idx = torch.randperm(256, device="cuda").int()
flush.zero_(); fused.modulate(xh, shb[:, :512], shb[:, 512:], idx)
torch.cuda.synchronize()
print("done")
