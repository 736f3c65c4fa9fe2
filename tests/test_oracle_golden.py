"""CPU: the oracle (oracle/pcaa_oracle.py) re-checked against the golden vectors that oracle/gen_golden.py produced
by running the REFERENCE itself (its nn.Modules, its unmodified train_variant4 trainer, scipy / sklearn scoring) in
the build container.  /root/reference is not needed (and not read) here.

Tolerances are the pins recorded when the vectors were generated: fp32 CPU arithmetic of two different
formulations (oracle = explicit restatement, reference = torch.nn layers) agrees to a few 1e-4 at worst.
"""
import os

import numpy as np
import pytest
import torch

from oracle import pcaa_oracle as O

CFG = dict(LR=1e-4, B1=0.9, B2=0.99, GP_WEIGHT=15, ADV_WEIGHT=1)
TOL = 5e-4


def digest(t):
    """Same fingerprint as oracle/gen_golden.py:digest."""
    f = t.detach().double().flatten()
    w = torch.cos(torch.arange(f.numel(), dtype=torch.float64) * 0.37)
    head = torch.zeros(8, dtype=torch.float64)
    head[: min(8, f.numel())] = f[:8]
    return torch.cat([torch.stack([f.sum(), f.abs().sum(), (f * w).sum()]), head]).numpy()


def maxdiff(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


@pytest.fixture(scope="module", autouse=True)
def _threads():
    torch.set_num_threads(min(8, os.cpu_count() or 1))


@pytest.mark.parametrize("name", ["n50_c2_b4", "n70_c4_b3"])
def test_modules_oracle_vs_reference_golden(golden_dir, name):
    gd = np.load(os.path.join(golden_dir, f"modules_{name}.npz"))
    B, nmax, C, seed = int(gd["B"]), int(gd["nmax"]), int(gd["C"]), int(gd["seed"])
    p = O.det_params(C, nmax, seed)
    pcs, gt = O.synth_batch(B, nmax, C, seed=1234 + seed)
    assert pcs.shape == (B, 4, 30, nmax) and pcs.dtype == torch.float32 and gt.dtype == torch.int64
    with torch.no_grad():
        upd = {}
        lg, fv = O.encoder_forward(p, pcs, True, True, upd)
        assert maxdiff(lg, gd["enc_train_logits"]) < TOL and maxdiff(fv, gd["enc_train_fv"]) < TOL
        for k in gd.files:
            if k.startswith("run:"):
                assert maxdiff(upd[k[4:]], gd[k]) < TOL, k
        lg, fv = O.encoder_forward(p, pcs, False, True)
        assert maxdiff(lg, gd["enc_eval_logits"]) < TOL and maxdiff(fv, gd["enc_eval_fv"]) < TOL
        # decoder + Chamfer evaluated on the reference's own embedding
        fv_ref = torch.from_numpy(gd["enc_eval_fv"])
        rec = O.decoder_forward(p, O.proj_head_forward(p, fv_ref), nmax)
        assert rec.shape == (B, 4, 30, nmax)
        assert maxdiff(rec[:, :, :2, :8], gd["rec_head"]) < TOL
        d = digest(rec)
        assert abs(d[1] - gd["rec_digest"][1]) / gd["rec_digest"][1] < 1e-5
        loss, i1, i2 = O.chamfer(rec, pcs)
        per, _, _ = O.chamfer(rec, pcs, avg_out=False)
        assert abs(float(loss) - float(gd["chamfer"])) < TOL * max(1.0, float(gd["chamfer"]))
        assert maxdiff(per, gd["chamfer_per_sample"]) < TOL * max(1.0, float(np.max(gd["chamfer_per_sample"])))
        # nearest-neighbour indices: bit-exact (integer output of the path, lowest index on ties)
        assert np.array_equal(i1.numpy(), gd["idx_gt_for_pred"].astype(np.int64))
        assert np.array_equal(i2.numpy(), gd["idx_pred_for_gt"].astype(np.int64))
        # independent clouds
        rng = np.random.default_rng(77 + seed)
        pr = torch.from_numpy(rng.normal(0, 0.6, pcs.shape).astype(np.float32))
        l2, k1, k2 = O.chamfer(pr, pcs)
        assert abs(float(l2) - float(gd["chamfer2"])) < 1e-4 * float(gd["chamfer2"])
        assert np.array_equal(k1.numpy(), gd["chamfer2_idx_gt_for_pred"].astype(np.int64))
        assert np.array_equal(k2.numpy(), gd["chamfer2_idx_pred_for_gt"].astype(np.int64))
        # critic
        oh = torch.nn.functional.one_hot(gt, C).float()
        assert maxdiff(O.disc_forward(p, fv_ref, oh), gd["disc_out"]) < 1e-5
    # the pins stored at generation time are within the stated tolerance
    for k in gd.files:
        if k.startswith("pin_") and "idx" not in k:
            assert float(gd[k]) < TOL, (k, float(gd[k]))
        if k.startswith("pin_") and "idx" in k:
            assert int(gd[k]) == 0, k


@pytest.mark.parametrize("name", ["n50_c2_b4", "v1_n50_c4_b8", "v2_n50_c2_b4", "v3_n50_c4_b4"])
def test_train_step_oracle_vs_reference_golden(golden_dir, name):
    """Two iterations of the variant-4 (PCAA_ablation.py:882-1021), variant-2 (train_AAE.py:126-290) and variant-3
    (PCAA_ablation.py:500-660) loops at B=4, N=50: losses, embeddings, gradient digests."""
    gd = np.load(os.path.join(golden_dir, f"step_{name}.npz"))
    variant = int(gd["variant"]) if "variant" in gd.files else 4
    B, nmax, C, seed, nsteps = (int(gd[k]) for k in ("B", "nmax", "C", "seed", "nsteps"))
    p = {4: lambda: O.det_params(C, nmax, seed), 1: lambda: O.det_params(C, nmax, seed, mean_learner=True)}.get(
        variant, lambda: O.det_params(C, nmax, seed, use_projection_head=False, dec_in=32))()
    means = torch.from_numpy(gd["means"])
    assert maxdiff(O.sample_distant_points(32, C, 10, 10).float(), means) == 0.0
    ost = {}
    rng = np.random.default_rng(999 + seed)
    for s in range(nsteps):
        pcs, gt = O.synth_batch(B, nmax, C, seed=4321 + 10 * seed + s)
        z0 = torch.from_numpy(rng.normal(0, 1, (B, 32))).float()
        alphas = torch.from_numpy(rng.uniform(0, 1, (B, 1)).astype(np.float32))
        o = O.train_step(p, ost, pcs, gt, z0, alphas, means, dict(CFG, NMAX=nmax), variant)
        for k in ("d_loss", "gp", "rec_loss", "loss_g", "sup_loss", "tot_loss"):
            want = float(gd[f"s{s}:{k}"])
            assert abs(float(o[k]) - want) < 2e-3 * max(1.0, abs(want)), (s, k, float(o[k]), want)
        assert maxdiff(o["fv"], gd[f"s{s}:fv"]) < 2e-3
        assert maxdiff(o["logits"], gd[f"s{s}:logits"]) < 2e-3
        if s == 0:
            for kind in ("d_grads", "g_grads"):
                for k, v in o[kind].items():
                    key = f"s{s}:grad:{k}"
                    if v is None:
                        assert key not in gd.files, k          # decoder bn1-4: grad is None in the reference too
                        continue
                    if k.endswith("module.0.bias") or k.endswith("conv1d.bias"):
                        continue                                # mathematically zero (feeds a train-mode BatchNorm)
                    want = gd[key]
                    got = digest(v)
                    assert abs(got[1] - want[1]) <= 2e-3 * want[1] + 1e-12, (k, got[1], want[1])
    for k in gd.files:
        if k.startswith("pin_") and ("loss" in k or k.endswith(":gp") or k.endswith(":fv")):
            assert float(gd[k]) < 2e-3, k
        if k.startswith("pin_") and "grad_rel" in k:
            assert float(gd[k]) < 2e-3, k
        if k.startswith("pin_") and "param_abs" in k:
            assert float(gd[k]) <= 2 * nsteps * CFG["LR"] * 1.01, k


def test_trainer_pin_recorded(golden_dir):
    """gen_golden.py replays the UNMODIFIED PCAA_ablation.train_variant4 with the oracle step; the outcome is recorded."""
    txt = open(os.path.join(golden_dir, "PIN.txt")).read()
    assert "unmodified PCAA_ablation.train_variant4" in txt
    w = float(txt.split("max |param diff| vs oracle replay = ")[1].split()[0])
    assert w <= 2 * 1e-4 * 2 * 1.01           # 2*lr*steps


def test_scoring_oracle_vs_scipy_sklearn_golden(golden_dir):
    gd = np.load(os.path.join(golden_dir, "scoring.npz"))
    emb, means, lik, labels, preds = gd["emb"], gd["means"], gd["lik"], gd["labels"], gd["preds"]
    C = means.shape[0]
    lo = O.joint_likelihood(emb, means)
    nz = lik > 0
    assert nz.sum() > 50
    assert np.max(np.abs(lo[nz] - lik[nz]) / lik[nz]) < 1e-9
    ll = O.joint_log_likelihood(emb, means)
    assert np.max(np.abs(ll[nz] - np.log(lik[nz]))) < 1e-9
    thr = float(gd["threshold"])
    assert O.roc_youden_threshold(labels, lik) == thr
    for k in (1, 2, 4, 6):
        n = (len(lik) // k) * k
        want = gd[f"votes_k{k}"]
        assert np.array_equal(O.openset_vote(lik[:n], preds[:n], thr, k, C), want)
        # the log-domain decision (what the CUDA kernel evaluates) gives the same labels wherever the float64 pdf did
        # not underflow to zero (SURVEY.md 8a exception class (i))
        ok = np.array([np.all(nz[w * k:(w + 1) * k]) for w in range(n // k)])
        votes_log = O.openset_vote(ll[:n], preds[:n], np.log(thr), k, C)
        assert np.array_equal(votes_log[ok], want[ok])


def test_vote_edge_cases():
    """strict majority (n_above > k/2), ties of bincount -> lowest class, 'unknown' = n_labels (inference_PCAA.py:263-271)."""
    lik = np.array([1.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0, 0.0])
    preds = np.array([2, 1, 0, 0, 3, 3, 1, 1])
    assert O.openset_vote(lik, preds, 0.5, 4, 4).tolist() == [4, 1]      # 2 of 4 is not > 2 ; {3,3,1,1} ties -> 1
    assert O.openset_vote(lik, preds, 0.5, 2, 4).tolist() == [1, 4, 3, 4]
    assert O.openset_vote(lik, preds, 1.0, 1, 4).tolist() == [4] * 8      # strict '>' against the threshold
    assert O.openset_vote(lik[:0], preds[:0], 0.5, 2, 4).tolist() == []


def test_causal_conv_matches_torch_conv1d():
    """models.py:59-76: Conv1d(k=3, dilation d, padding 2d) then drop the last 2d frames."""
    torch.manual_seed(3)
    for dil in (1, 2, 4):
        conv = torch.nn.Conv1d(5, 7, 3, padding=2 * dil, dilation=dil)
        x = torch.randn(2, 5, 30)
        want = conv(x)[:, :, :-2 * dil]
        got = O.causal_dilated_conv(x.permute(0, 2, 1).contiguous(), conv.weight, conv.bias, dil).permute(0, 2, 1)
        assert maxdiff(got.detach(), want.detach()) < 1e-5


def test_chamfer_matches_expanded_form_and_ties():
    """utils.py:98-132 on a cloud with duplicated gt points (padding of datasets.py:127-134): lowest index wins."""
    rng = np.random.default_rng(0)
    gts = torch.from_numpy(rng.normal(0, 1, (1, 4, 2, 6)).astype(np.float32))
    gts[..., 3:] = gts[..., :3]                                  # duplicates -> exact ties over the gt index
    preds = torch.from_numpy(rng.normal(0, 1, (1, 4, 2, 6)).astype(np.float32))
    loss, i1, i2 = O.chamfer(preds, gts)
    assert int(i1.max()) < 3                                     # first of each duplicate pair
    x = gts.permute(0, 2, 3, 1).double()
    y = preds.permute(0, 2, 3, 1).double()
    P = ((x[:, :, :, None, :] - y[:, :, None, :, :]) ** 2).sum(-1)
    want = (P.min(2).values.sum(-1) + P.min(3).values.sum(-1)).mean()
    assert abs(float(loss) - float(want)) < 1e-4


@pytest.mark.parametrize("k", [1, 2, 4, 6])
def test_procedure_oracle_vs_reference_golden(golden_dir, k):
    """oracle.naive_sequential_procedure == the reference's unmodified inference_PCAA.naive_sequential_procedure run on a
    stand-in dataset / encoder (oracle/gen_golden.py::procedure_case): window skip rules, validation-subject draw,
    ROC threshold, vote, label conventions and metrics."""
    g = np.load(os.path.join(golden_dir, "procedure.npz"))
    o = O.naive_sequential_procedure(k, g["t_emb"], g["t_pred"], g["t_lab"], g["u_emb"], g["u_pred"], g["u_lab"], g["means"], 0, 0.2)
    assert np.array_equal(o["preds"], g[f"preds_k{k}"]) and np.array_equal(o["labels"], g[f"labels_k{k}"])
    m = o["metrics"]
    assert np.allclose([m["accuracy"], m["f1_micro"], m["f1_macro"], m["f1_weighted"]], g[f"metrics_k{k}"], atol=1e-12)
    assert o["threshold"] == float(g["threshold"])
