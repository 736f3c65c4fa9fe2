"""The reference's own scripts, unmodified, from baseline/_ref (SURVEY.md section 8 row b', section 10).

CPU tier: the harness drives the reference's procedures on the reference's own modules (proves the harness adds nothing
but the stand-ins it lists).  GPU tier: the same procedures with this repository's modules swapped in under the names the
scripts import -- ``PCAA_ablation.train_variant4`` (1 epoch), ``inference_PCAA.CGAAE_inference`` and the loop body of
``train_pointsubsampling.py`` -- compared with the reference's modules run by the same harness on the box's CPU:

  * epoch metrics (wandb.log quantities) of one training epoch from identical seeds / data: losses within 3e-2 relative
    (reconstruction, cross-entropy), the critic loss within 8e-2 of max(1, |ref|) (its gradient-penalty term is a
    difference of near-equal numbers);
  * inference from ONE checkpoint (written by the B200 run, loaded by both): the true open-set labels are equal; on each
    side the script's predicted labels are EXACTLY what the oracle procedure derives from the embeddings / logits that side's
    encoder returned (captured by a forward hook); across sides the embeddings agree within 3e-2 of max |ref| and the class
    predictions wherever the reference's top-2 logit gap exceeds the logit tolerance.  (The predicted open-set labels of the
    two sides are printed, not asserted: after 2 training iterations all likelihoods of this model are near-ties around the
    ROC threshold, so a 1e-2 embedding difference legitimately flips windows; tests/test_gpu_baseline_sizes.py bounds
    that effect on a well-separated 4 096-crop stream.)
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "scripts", "run_reference_scripts.py")
sys.path.insert(0, ROOT)

from baseline import refenv, install_ref  # noqa: E402


def run_harness(tmp, name, *args, timeout=1500):
    out = os.path.join(str(tmp), f"{name}.json")
    env = dict(os.environ, PYTHONHASHSEED="0", WANDB_MODE="disabled")
    r = subprocess.run([sys.executable, SCRIPT, "--out", out, *args], capture_output=True, text=True, env=env, timeout=timeout,
                       cwd=str(tmp))
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    with open(out) as f:
        return json.load(f)


needs_ref = pytest.mark.skipif(not refenv.available(), reason="baseline/_ref not installed (python baseline/install_ref.py)")

EXPECTED_FILES = {"config.pkl", "discriminator_means.pt"} | {f"m_{s}.pt" for s in ("E", "G", "D", "GPH", "DPH")}
EPOCH_KEYS = {"Reconstruction Loss Train", "Reconstruction Loss Valid", "Cross Entropy Loss Train", "Cross Entropy Loss Valid",
              "Discriminator Loss", "Total Loss Train", "Train Accuracy", "Valid Accuracy"}


@needs_ref
def test_installed_reference_is_unmodified():
    assert install_ref.verify()
    if os.path.isdir("/root/reference"):
        for name in os.listdir(refenv.REF_DIR):
            if name.endswith(".py"):
                with open(os.path.join(refenv.REF_DIR, name), "rb") as a, open(os.path.join("/root/reference", name), "rb") as b:
                    assert a.read() == b.read(), name


@needs_ref
def test_harness_runs_the_reference_scripts_on_cpu(tmp_path):
    res = run_harness(tmp_path, "ref", "--impl", "reference", "--device", "cpu", "--workdir", str(tmp_path / "w"),
                      "--model-name", "m", "--ks", "2")
    run = res["runs"][0]
    assert set(run["files"]) == EXPECTED_FILES
    assert len(run["epochs"]) == 1 and set(run["epochs"][0]) == EPOCH_KEYS
    inf = run["inference"]["2"]
    assert len(inf["preds"]) == len(inf["labels"]) > 0
    assert set(inf["labels"]) <= {0, 1, 2} and set(inf["preds"]) <= {0, 1, 2}
    # the reference's labels == the oracle procedure applied to what the reference's encoder returned (pins the replay helper
    # the GPU test uses, and once more the oracle's procedure against the unmodified inference_PCAA.naive_sequential_procedure)
    rep = _replay_with_oracle(run, (2,))[2][0]
    assert inf["preds"] == rep["preds"].tolist() and inf["labels"] == rep["labels"].tolist()


def _replay_with_oracle(run, ks):
    """Re-derive the open-set labels of one harness run from what its encoder returned, with the oracle procedure
    (threshold from the batch-1 pass, votes from the batch-k pass: inference_PCAA.py:195-231, 239-314).  Returns
    {k: (oracle result, first-pass embeddings [n_t + n_u, 32], first-pass logits)}."""
    from oracle import pcaa_oracle as O
    t_lab, u_lab = np.array(run["test_labels"]), np.array(run["unseen_labels"])
    means = np.array(run["means"], dtype=np.float32)
    calls = run["encoder_calls"]
    n_t, n_u = len(t_lab), len(u_lab)
    rng = np.random.default_rng(0)
    subj = np.unique(u_lab)
    val = rng.choice(subj, size=np.ceil(0.2 * len(subj)).astype(int), replace=False)
    pos, out = 0, {}
    for k in ks:
        p1 = calls[pos:pos + n_t + n_u]
        pos += n_t + n_u
        assert all(len(c[1]) == 1 for c in p1), "first pass: one crop per encoder call (inference_PCAA.py:196-208)"
        emb1 = np.array([c[1][0] for c in p1], dtype=np.float32)
        lg1 = np.array([c[0][0] for c in p1], dtype=np.float32)
        emb2, lg2 = emb1.copy(), lg1.copy()
        for base, lab, is_unseen in ((0, t_lab, False), (n_t, u_lab, True)):
            for w in range(len(lab) // k):
                sl = lab[w * k:(w + 1) * k]
                if len(np.unique(sl)) != 1 or (is_unseen and sl[0] in val):
                    continue                                              # skipped before the encoder is called (:243-244, :279-284)
                lgk, fvk = calls[pos]
                pos += 1
                assert len(fvk) == k
                emb2[base + w * k: base + (w + 1) * k] = np.array(fvk, dtype=np.float32)
                lg2[base + w * k: base + (w + 1) * k] = np.array(lgk, dtype=np.float32)
        o = O.naive_sequential_procedure(k, emb1[:n_t], lg2[:n_t].argmax(1), t_lab, emb1[n_t:], lg2[n_t:].argmax(1), u_lab, means, 0, 0.2,
                                         vote_test_emb=emb2[:n_t], vote_unseen_emb=emb2[n_t:])
        out[k] = (o, emb1, lg1)
    assert pos == len(calls), (pos, len(calls))
    return out


@needs_ref
@pytest.mark.gpu
def test_reference_scripts_run_unchanged_on_the_b200_modules(tmp_path):
    w = str(tmp_path / "w_b200")
    b200 = run_harness(tmp_path, "b200", "--impl", "b200", "--workdir", w, "--model-name", "m", "--ks", "6,2")
    run = b200["runs"][0]
    assert b200["device"] == "cuda" and b200["c_abi_calls"] > 500            # the CUDA library did the work
    assert set(run["files"]) == EXPECTED_FILES                                # config.pkl + the five state_dicts + means
    assert len(run["epochs"]) == 1 and set(run["epochs"][0]) == EPOCH_KEYS
    # (1) one training epoch: B200 modules vs the reference's modules (CPU, fp32), same seeds and data
    ref = run_harness(tmp_path, "ref_train", "--impl", "reference", "--device", "cpu", "--workdir", str(tmp_path / "w_ref"),
                      "--model-name", "m", "--skip-infer")
    eb, er = run["epochs"][0], ref["runs"][0]["epochs"][0]
    print("epoch metrics  b200:", eb, "\n               ref :", er)
    for k in ("Reconstruction Loss Train", "Reconstruction Loss Valid", "Cross Entropy Loss Train", "Cross Entropy Loss Valid",
              "Total Loss Train"):
        assert abs(eb[k] - er[k]) <= 3e-2 * abs(er[k]), (k, eb[k], er[k])
    assert abs(eb["Discriminator Loss"] - er["Discriminator Loss"]) <= 8e-2 * max(1.0, abs(er["Discriminator Loss"]))
    # (2) inference from the SAME checkpoint (the one the B200 run wrote): reference modules on the CPU vs B200 modules
    refi = run_harness(tmp_path, "ref_infer", "--impl", "reference", "--device", "cpu", "--workdir", w, "--model-name", "m",
                       "--ks", "6,2", "--skip-train")
    rb, rr = _replay_with_oracle(run, (6, 2)), _replay_with_oracle(refi["runs"][0], (6, 2))
    for k in (6, 2):
        a, b = run["inference"][str(k)], refi["runs"][0]["inference"][str(k)]
        assert a["labels"] == b["labels"], k                                  # skip rules / validation subjects: exact
        # the script's labels are exactly what the oracle procedure makes of the embeddings the script's encoder returned:
        # on the B200 modules as on the reference's (no tolerance: integer outputs)
        for name, res, rep in (("b200", a, rb[k][0]), ("reference", b, rr[k][0])):
            assert res["preds"] == rep["preds"].tolist() and res["labels"] == rep["labels"].tolist(), (name, k)
        # same checkpoint, same crops: embeddings within the bf16 tolerance, class predictions equal where decided
        emb_b, lg_b, emb_r, lg_r = rb[k][1], rb[k][2], rr[k][1], rr[k][2]
        e = float(np.abs(emb_b - emb_r).max() / np.abs(emb_r).max())
        top2 = np.sort(lg_r, axis=1)[:, -2:]
        decided = (top2[:, 1] - top2[:, 0]) > 3e-2 * np.abs(lg_r).max()
        pm = np.nonzero(lg_b.argmax(1) != lg_r.argmax(1))[0]
        pa, pb = np.array(a["preds"]), np.array(b["preds"])
        diff = np.nonzero(pa != pb)[0]
        print(f"k={k}: {len(emb_r)} crops, embeddings relmax {e:.4f}; class predictions differ at {pm.tolist()} "
              f"(undecided: {np.nonzero(~decided)[0].tolist()}); thresholds b200 {rb[k][0]['threshold']:.4e} / ref {rr[k][0]['threshold']:.4e}; "
              f"{len(pa)} windows, open-set labels differ at {diff.tolist()}")
        assert e < 3e-2 and not decided[pm].any()


@needs_ref
@pytest.mark.gpu
def test_pointsubsampling_loop_on_the_b200_modules(tmp_path):
    """train_pointsubsampling.py:52-71 for n_points = 50 and 70: generate_splits(nmax_points), train_variant4, then
    CGAAE_inference(ks=[1, 2, 4, 6], variation=V4) -- every (N, k) cell produces its log and label files."""
    res = run_harness(tmp_path, "npts", "--impl", "b200", "--workdir", str(tmp_path / "w"), "--pointsubsampling", "50,70")
    assert [r["nmax"] for r in res["runs"]] == [50, 70]
    for r in res["runs"]:
        assert r["model_name"] == f"PCAA_npts_V4_{r['nmax']}.2.1"
        assert set(r["inference"]) == {"1", "2", "4", "6"}
        for k, inf in r["inference"].items():
            assert inf["log"]["n_steps"] == int(k) and len(inf["preds"]) == len(inf["labels"]) > 0
        assert np.isfinite(list(r["epochs"][0].values())).all()
