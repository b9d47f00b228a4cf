"""Where do the framework's copy / cast / cat / add passes of one eager training step come from?

    python tools/trace_copies.py [--depth 4] [--min-elems 1000000]

Runs one forward + backward + optimizer step of tools/train_step.py under a TorchDispatchMode that records every
aten copy / cast / cat / elementwise add-mul call on tensors of at least --min-elems elements, with dtypes, contiguity and the
innermost frames of THIS repo on the Python stack, and prints them grouped.  A development aid for removing glue passes
around the kernels (profiles/r2_train_step_glue.md)."""
import argparse
import collections
import os
import sys
import traceback

import torch
from torch.utils._python_dispatch import TorchDispatchMode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

WATCH = ("copy_", "_to_copy", "clone", "cat", "add", "mul", "sum", "contiguous", "fill_", "zero_", "stack", "add_", "mul_")


class Tracer(TorchDispatchMode):
    def __init__(self, min_elems):
        super().__init__()
        self.min_elems, self.seen = min_elems, collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        name = func.__name__.split(".")[0]
        if name in WATCH:
            tensors = [a for a in args if isinstance(a, torch.Tensor)]
            if args and isinstance(args[0], (list, tuple)):
                tensors = [a for a in args[0] if isinstance(a, torch.Tensor)]
            big = [t for t in tensors if t.numel() >= self.min_elems]
            if big:
                desc = ", ".join(f"{tuple(t.shape)}:{str(t.dtype)[6:]}:{'c' if t.is_contiguous() else 's'}" for t in tensors[:3])
                o = out if isinstance(out, torch.Tensor) else None
                if o is not None:
                    desc += f" -> {str(o.dtype)[6:]}"
                frames = [f for f in traceback.extract_stack() if ROOT in f.filename and "trace_copies" not in f.filename]
                where = " < ".join(f"{os.path.basename(f.filename)}:{f.lineno}" for f in frames[-3:][::-1])
                self.seen[(name, desc, where)] += 1
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--min-elems", type=int, default=1_000_000)
    args = ap.parse_args()
    from train_step import TrainStep
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = True
    ts = TrainStep(dev, 0, 1, batch=32, dtype="bf16", depth=args.depth, use_graph=False)
    ts.step()
    tr = Tracer(args.min_elems)
    with tr:
        ts.step()
    torch.cuda.synchronize()
    for (name, desc, where), n in sorted(tr.seen.items(), key=lambda kv: (-kv[1], kv[0])):
        print(f"{n:4d}  {name:10s} {desc}\n        {where}")


if __name__ == "__main__":
    main()
