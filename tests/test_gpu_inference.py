"""Open-set inference on the B200 path against the oracle and the reference-generated golden vectors (B200 only).

Integer outputs (window labels, class predictions away from logit ties) are bit-exact; embeddings within the bf16
tolerance stated in tests/test_gpu_step.py (3e-2 of max |ref|)."""
import os

import numpy as np
import pytest
import torch

from oracle import pcaa_oracle as O

pytestmark = pytest.mark.gpu


def _encoder(C, nmax, seed=0):
    from opensetgaitrecognition_pcaa_b200 import models
    p = O.det_params(C, nmax, seed)
    enc = models.CGEncoder(n_out_labels=C, use_projection_head=True, nmax_points=nmax)
    enc.load_state_dict({k[2:]: v.clone() for k, v in p.items() if k.startswith("E.")})
    return enc.cuda().float().eval(), p


@pytest.mark.parametrize("k", [1, 2, 4, 6])
def test_procedure_from_embeddings_matches_reference_golden(golden_dir, k):
    """Window skip rules, validation-subject draw, ROC threshold (host), fused log-likelihood + vote kernels: identical
    labels / predictions / metrics to the reference's own procedure (tests/golden/procedure.npz)."""
    from opensetgaitrecognition_pcaa_b200 import inference as I
    g = np.load(os.path.join(golden_dir, "procedure.npz"))
    emb = {"test": (torch.from_numpy(g["t_emb"]).cuda(), torch.from_numpy(g["t_pred"].astype(np.int32)).cuda()),
           "unseen": (torch.from_numpy(g["u_emb"]).cuda(), torch.from_numpy(g["u_pred"].astype(np.int32)).cuda())}
    out = I.naive_sequential_procedure(k, None, torch.from_numpy(g["means"]).cuda(), None, g["t_lab"], None, g["u_lab"],
                                       seed=0, unseen_valid_ratio=0.2, embeddings=emb)
    assert np.array_equal(out["preds"], g[f"preds_k{k}"]) and np.array_equal(out["labels"], g[f"labels_k{k}"])
    m = out["metrics"]
    assert np.allclose([m["accuracy"], m["f1_micro"], m["f1_macro"], m["f1_weighted"]], g[f"metrics_k{k}"], atol=1e-12)
    assert abs(out["threshold"] - float(g["threshold"])) <= 1e-9 * float(g["threshold"])


def test_encode_matches_oracle_eval_forward_and_is_batch_independent():
    from opensetgaitrecognition_pcaa_b200 import inference as I
    C, nmax, M = 4, 50, 13
    enc, p = _encoder(C, nmax)
    pcs, _ = O.synth_batch(M, nmax, C, seed=77)
    logits_ref, fv_ref = O.encoder_forward(p, pcs, False, True)
    fv, pred = I.encode(enc, pcs.cuda(), batch=5)                 # ragged last chunk
    err = float((fv.cpu() - fv_ref).abs().max() / fv_ref.abs().max())
    assert err < 3e-2, err
    top2 = logits_ref.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 6e-2 * float(logits_ref.abs().max())
    assert torch.equal(pred.cpu().long()[decided], logits_ref.argmax(1)[decided])
    # the reference encodes a crop alone (phase 1) and inside a batch of k (phase 2): same embedding in eval mode.  Here the
    # position of a crop inside the batch changes the summation order of the mean pool (fp32 rounding), and a last-bit
    # difference can flip the bf16 rounding of a TCN operand: equal to within a few bf16 ulps of the largest entry.
    fv1, pred1 = I.encode(enc, pcs.cuda(), batch=1)
    assert float((fv1 - fv).abs().max()) <= 4e-3 * float(fv.abs().max())
    assert torch.equal(pred1.cpu().long()[decided], pred.cpu().long()[decided])
    # the drop-in module's forward gives the same embedding
    with torch.no_grad():
        lg_m, fv_m = enc(pcs.cuda())
    assert float((fv_m - fv).abs().max()) <= 4e-3 * float(fv.abs().max())


def test_full_procedure_from_point_clouds_and_sharded_stream():
    """End to end from crops: the device procedure equals the oracle procedure evaluated on the SAME embeddings (exact
    integer parity), and a 2-way sharded stream gives the same window labels as the unsharded one."""
    from opensetgaitrecognition_pcaa_b200 import inference as I
    C, nmax, k = 4, 50, 2
    enc, p = _encoder(C, nmax, seed=3)
    rng = np.random.default_rng(0)
    t_pcs, t_lab = O.synth_batch(36, nmax, C, seed=5)
    t_lab = np.sort(t_lab.numpy())
    u_pcs, _ = O.synth_batch(30, nmax, C, seed=6)
    u_pcs = u_pcs * 2.5
    u_lab = np.repeat(np.array([21, 22, 23, 24, 25]), 6)
    means = O.sample_distant_points(32, C, 10, 10).float()
    out = I.naive_sequential_procedure(k, enc, means.cuda(), t_pcs.cuda(), t_lab, u_pcs.cuda(), u_lab, seed=0)
    (fv_t, pr_t), (fv_u, pr_u) = out["embeddings"]["test"], out["embeddings"]["unseen"]
    ref = O.naive_sequential_procedure(k, fv_t.cpu().numpy(), pr_t.cpu().numpy(), t_lab, fv_u.cpu().numpy(), pr_u.cpu().numpy(),
                                       u_lab, means.numpy(), 0, 0.2)
    assert np.array_equal(out["preds"], ref["preds"]) and np.array_equal(out["labels"], ref["labels"])
    assert np.array_equal(out["val_subjects"], ref["val_subjects"])
    assert abs(out["threshold"] - ref["threshold"]) <= 1e-9 * abs(ref["threshold"])
    # sharded stream (config 5): two shards of whole windows == the unsharded stream
    lthr = out["log_threshold"]
    ll, votes, _ = I.sharded_stream_inference(enc, means, t_pcs.cuda(), k, lthr, C)
    ll_a, v_a, _ = I.sharded_stream_inference(enc, means, t_pcs[:20].cuda(), k, lthr, C)
    ll_b, v_b, _ = I.sharded_stream_inference(enc, means, t_pcs[20:].cuda(), k, lthr, C)
    assert torch.equal(torch.cat([v_a, v_b]), votes)
    assert float((torch.cat([ll_a, ll_b]) - ll).abs().max()) < 5e-2 * float(ll.abs().max())   # scores move with the embeddings' rounding
