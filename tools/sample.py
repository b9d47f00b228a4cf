"""Class-conditional sampling to disk, the ODE mode of the reference's dimsum/sample_ddp.py on this repo's kernels.

    python tools/sample.py --model DiM-L/2 --ckpt pytorch_model.bin --num-fid-samples 50000 --per-proc-batch-size 256
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sample.py ...

Same bookkeeping as sample_ddp.py:52-191: per-rank seed `global_seed * world + rank` (:64), `total_samples` rounded up to a
multiple of the global batch (:140-149), labels uniform in [0, num_classes - 1) with the null class appended for CFG
(:159-173), sample `index = i * world + rank + total` (:184), one `<folder>.npz` with `arr_0` of the first --num-fid-samples
samples written by rank 0 after a barrier (:187-191), folder name `<model>-<ckpt>-cfg-<s>-<n>-ODE-<steps>-euler` (:113-118).

What differs: the integrator is the fixed-grid Euler restatement of `torchdiffeq.odeint` (dimsum_b200/sampler.py; the adaptive
dopri5 needs the package), each grid point is one CUDA-graph replay, and the SD-VAE decode (:176, a `diffusers` download) is
optional: without `--decoder` the LATENTS are written (`<index>.npy`, `arr_0` float32 (N, 4, H/8, W/8)); `--decoder` names a
TorchScript module mapping latents / 0.18215 to images in [-1, 1], whose output is stored as uint8 PNGs like the reference.
"""
import argparse
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def plan(num_fid_samples, per_proc_batch_size, world):
    """-> (total_samples, iterations per rank), sample_ddp.py:140-149."""
    global_batch = per_proc_batch_size * world
    total = int(math.ceil(num_fid_samples / global_batch) * global_batch)
    return total, total // world // per_proc_batch_size


def sample_index(i, rank, world, total_so_far):
    return i * world + rank + total_so_far


def folder_name(args):
    ckpt = os.path.basename(args.ckpt).replace(".pt", "") if args.ckpt else "pretrained"
    return f"{args.model.replace('/', '-')}-{ckpt}-cfg-{args.cfg_scale}-{args.per_proc_batch_size}-ODE-{args.num_sampling_steps}-euler"


def build_npz(sample_dir, num, as_images):
    """One .npz from the per-sample files (sample_ddp.py:36-50)."""
    if as_images:
        from PIL import Image
        arr = np.stack([np.asarray(Image.open(f"{sample_dir}/{i:06d}.png")).astype(np.uint8) for i in range(num)])
    else:
        arr = np.stack([np.load(f"{sample_dir}/{i:06d}.npy") for i in range(num)])
    path = f"{sample_dir}.npz"
    np.savez(path, arr_0=arr)
    return path, arr.shape


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="DiM-L/2")
    ap.add_argument("--ckpt", default=None, help="reference checkpoint (state dict, or a train.py checkpoint: its EMA weights)")
    ap.add_argument("--image-size", type=int, default=256)
    ap.add_argument("--num-classes", type=int, default=1001, help="classes INCLUDING the null class (sample_ddp.py:151-155)")
    ap.add_argument("--cfg-scale", type=float, default=4.0)
    ap.add_argument("--num-sampling-steps", type=int, default=250)
    ap.add_argument("--per-proc-batch-size", type=int, default=32)
    ap.add_argument("--num-fid-samples", type=int, default=50_000)
    ap.add_argument("--global-seed", type=int, default=0)
    ap.add_argument("--sample-dir", default="samples")
    ap.add_argument("--scan-type", default="none")
    ap.add_argument("--tf32", action=argparse.BooleanOptionalAction, default=True)
    ap.add_argument("--decoder", default=None, help="TorchScript latent decoder; without it latents are written")
    args = ap.parse_args()

    from dimsum_b200 import checkpoint
    from dimsum_b200.models_dim import DiM_models
    from dimsum_b200.sampler import sample_cfg, euler_velocity_ode
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = args.tf32
    torch.set_grad_enabled(False)
    seed = args.global_seed * world + rank
    torch.manual_seed(seed)
    print(f"Starting rank={rank}, seed={seed}, world_size={world}.")

    latent = args.image_size // 8
    real_classes = args.num_classes - 1 if args.num_classes > 1 else args.num_classes
    model = DiM_models[args.model](img_resolution=latent, in_channels=4, num_classes=real_classes, label_dropout=0.1,
                                   scan_type=args.scan_type).to(dev).eval()
    if args.ckpt:
        checkpoint.load_model(model, args.ckpt, strict=True)
    elif rank == 0:
        print("no --ckpt: sampling from randomly initialised weights")
    decoder = torch.jit.load(args.decoder, map_location=dev).eval() if args.decoder else None

    out_dir = os.path.join(args.sample_dir, folder_name(args))
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)
    if world > 1:
        dist.barrier()
    n = args.per_proc_batch_size
    total_samples, iterations = plan(args.num_fid_samples, n, world)
    if rank == 0:
        print(f"Total number of samples that will be drawn: {total_samples}")
    total = 0
    for _ in range(iterations):
        z = torch.randn(n, model.in_channels, latent, latent, device=dev)
        y = torch.randint(0, real_classes, (n,), device=dev)
        if args.cfg_scale > 1.0:
            samples = sample_cfg(model, z, y, cfg_scale=args.cfg_scale, num_steps=args.num_sampling_steps,
                                 null_class=real_classes, use_graph=True)
        else:
            samples = euler_velocity_ode(lambda xx, tt: model(xx, tt, y), z, args.num_sampling_steps)
        if decoder is not None:
            from PIL import Image
            img = decoder(samples / 0.18215)
            img = torch.clamp(127.5 * img + 128.0, 0, 255).permute(0, 2, 3, 1).to("cpu", dtype=torch.uint8).numpy()
            for i, im in enumerate(img):
                Image.fromarray(im).save(f"{out_dir}/{sample_index(i, rank, world, total):06d}.png")
        else:
            host = samples.float().cpu().numpy()
            for i, lat in enumerate(host):
                np.save(f"{out_dir}/{sample_index(i, rank, world, total):06d}.npy", lat)
        total += n * world
        if world > 1:
            dist.barrier()
    if world > 1:
        dist.barrier()
    if rank == 0:
        path, shape = build_npz(out_dir, args.num_fid_samples, decoder is not None)
        print(f"Saved .npz file to {path} [shape={shape}].")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
