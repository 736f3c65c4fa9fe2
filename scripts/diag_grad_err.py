"""Diagnostic: per-tensor relative gradient error of the fused step vs the oracle (GPU box)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pcaa_oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_step import build, CFG, relnorm, bn_cancelled_bias
from opensetgaitrecognition_pcaa_b200.train import PCAATrainer
B, nmax, C, seed = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (4, 50, 2, 0)
p = O.det_params(C, nmax, seed); po = {k: v.clone() for k, v in p.items()}
enc, dec, dis, gph = build(p, C, nmax)
means = O.sample_distant_points(32, C, 10, 10).float()
tr = PCAATrainer(enc, dec, dis, gph, means, CFG); ost = {}
rng = np.random.default_rng(999 + seed)
for s in range(2):
    pcs, gt = O.synth_batch(B, nmax, C, seed=4321 + 10 * seed + s)
    z0 = torch.from_numpy(rng.normal(0, 1, (B, 32))).float(); alphas = torch.from_numpy(rng.uniform(0, 1, (B, 1)).astype(np.float32))
    ref = O.train_step_variant4(po, ost, pcs, gt, z0, alphas, means, dict(CFG, NMAX=nmax))
    out = tr.step(pcs.cuda(), gt.cuda(), z0.cuda(), alphas.cuda())
    print("step", s, {k: (float(out[k]), float(ref[k])) for k in ("rec_loss", "d_loss", "sup_loss", "loss_g")})
    for kind, flat in (("g_grads", tr.G), ("d_grads", tr.D)):
        for n, g_ref in ref[kind].items():
            if g_ref is None or n not in flat.slices or bn_cancelled_bias(n): continue
            print(f"  {n:50s} relnorm {relnorm(flat.view(flat.g, n), g_ref):.4f}  |g|={float(g_ref.norm()):.3e}")
