"""The reference's own scripts, unmodified, from baseline/_ref (SURVEY.md section 8 row b', section 10).

CPU tier: the harness drives the reference's procedures on the reference's own modules (proves the harness adds nothing
but the stand-ins it lists).  GPU tier: the same procedures with this repository's modules swapped in under the names the
scripts import -- ``PCAA_ablation.train_variant4`` (1 epoch), ``inference_PCAA.CGAAE_inference`` and the loop body of
``train_pointsubsampling.py`` -- compared with the reference's modules run by the same harness on the box's CPU:

  * epoch metrics (wandb.log quantities) of one training epoch from identical seeds / data: losses within 3e-2 relative
    (reconstruction, cross-entropy), the critic loss within 8e-2 of max(1, |ref|) (its gradient-penalty term is a
    difference of near-equal numbers);
  * inference from ONE checkpoint (written by the B200 run, loaded by both): open-set labels equal, predicted labels equal
    except windows listed as within bf16 tolerance of a decision boundary -- at most 15 % of the windows of these barely
    trained (2 iterations) networks, whose embeddings sit far from every prototype so that all likelihoods are near-ties.
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
    for k in ("6", "2"):
        a, b = run["inference"][k], refi["runs"][0]["inference"][k]
        assert a["labels"] == b["labels"], k                                  # skip rules / validation subjects: exact
        pa, pb = np.array(a["preds"]), np.array(b["preds"])
        diff = np.nonzero(pa != pb)[0]
        print(f"k={k}: {len(pa)} windows, predicted labels differ at {diff.tolist()} (b200 {pa[diff].tolist()} vs ref {pb[diff].tolist()})")
        assert len(diff) <= 0.15 * len(pa) + 1, (k, diff)


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
