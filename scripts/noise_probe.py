"""How reproducible is one train step on ONE GPU?  (context for the data-parallel parity floors)

The same iteration is run from the same snapshot (a) twice back to back on an otherwise idle GPU and (b) once more while an
unrelated memory-bound kernel runs on a second stream (what the data-parallel step looks like: the gradient exchange and the
decoder's Adam update run beside the encoder backward).  Floating-point atomics (mean-pool partial sums, layer-1 weight
gradient, critic) and TMA reduce-adds (split-K weight gradients) then complete in another order; a last-bit difference can flip
the bf16 rounding of a stored activation or a Chamfer nearest neighbour, and sums with cancellation (BatchNorm d beta = sum of
dz over all points) show it most.  Prints flat and worst-tensor relative deviations.

    python scripts/noise_probe.py [B] [N]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opensetgaitrecognition_pcaa_b200 import dp, synth  # noqa: E402
from opensetgaitrecognition_pcaa_b200.train import build_variant4  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda", 0)
tr = build_variant4(4, N, seed=0, device=dev)
pcs, gt = synth.synth_batch(B, N, 4, seed=11)
rng = np.random.default_rng(3)
inp = (pcs.to(dev), gt.to(dev), torch.from_numpy(rng.normal(0, 1, (B, 32))).float().to(dev),
       torch.from_numpy(rng.uniform(0, 1, (B, 1)).astype(np.float32)).to(dev))
tr.step(*inp)
snap = tr.snapshot()


def run(disturb: bool):
    tr.restore(snap)
    torch.cuda.synchronize()
    if disturb:
        side = torch.cuda.Stream()
        junk = torch.empty(1 << 28, device=dev)
        with torch.cuda.stream(side):
            for _ in range(40):
                junk.mul_(1.0001)
    tr.step(*inp)
    torch.cuda.synchronize()
    return tr.G.g.clone()


g0 = run(False)
for name, g in (("repeat, idle GPU", run(False)), ("repeat, idle GPU", run(False)), ("with a concurrent memory-bound stream", run(True)),
                ("with a concurrent memory-bound stream", run(True))):
    flat = float((g - g0).norm() / g0.norm())
    worst, where = dp.per_tensor_relnorm(tr.G, g, g0)
    print(f"B={B} N={N} {name:40s}: ||g - g0|| / ||g0|| = {flat:.2e}, worst tensor {worst:.2e} ({where})")
