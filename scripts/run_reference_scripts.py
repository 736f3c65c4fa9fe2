#!/usr/bin/env python
"""Run the reference's OWN, UNMODIFIED procedures -- ``PCAA_ablation.train_variant4``, ``inference_PCAA.CGAAE_inference``
and the loop body of ``train_pointsubsampling.py:52-71`` -- from ``baseline/_ref`` on a synthetic on-disk dataset,
either on the reference's own modules (``--impl reference``, CPU or CUDA eager) or with this repository's drop-in
modules swapped in under the names the scripts import (``--impl b200``).

No reference file is edited.  What the harness supplies from outside, exactly as SURVEY section 10 lists:
  * stand-in modules for the plotting-only imports matplotlib / umap, WANDB_MODE=disabled;
  * ``MSRadarDataset.generate_splits`` replaced by a writer of synthetic crops in the reference's file convention
    (the real one needs the raw mmGait10 pickles, which are not available offline);
  * ``inference_PCAA.plot_confusion_matrix_cgaae`` replaced by its two data lines (the rest is matplotlib + usetex);
  * ``constants.BATCH_SIZE = config["BATCH_SIZE"]`` (the gradient-penalty alphas read the module constant);
  * ``wandb.log`` wrapped to capture the per-epoch metrics the trainer reports.

Writes one JSON (``--out``) with the captured epoch metrics, the inference logs and the predicted / true open-set labels.
Run in a fresh process (it registers modules under the reference's names) with PYTHONHASHSEED fixed: the
``sequential=True`` crop order depends on Python set iteration (datasets.py:398-411).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--device", default=None, help="constants.DEVICE (default: cuda for b200, cpu for reference)")
    ap.add_argument("--workdir", default=None, help="scratch directory (cwd of the reference's relative paths)")
    ap.add_argument("--model-name", default="PCAA_b200_V4")
    ap.add_argument("--nmax", type=int, default=50)
    ap.add_argument("--classes", type=int, default=2)
    ap.add_argument("--unseen", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--crops-per-track", type=int, default=8)
    ap.add_argument("--tracks", type=int, default=2)
    ap.add_argument("--ks", default="6,2")
    ap.add_argument("--skip-train", action="store_true", help="inference only, from the checkpoint already in --workdir")
    ap.add_argument("--skip-infer", action="store_true")
    ap.add_argument("--pointsubsampling", default="", help="comma list of n_points: run the loop body of "
                    "train_pointsubsampling.py:52-71 (generate_splits, train_variant4, CGAAE_inference ks=1,2,4,6) for each")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    import numpy as np
    import torch
    from baseline import refenv
    from opensetgaitrecognition_pcaa_b200 import synth            # synthetic crop writer only (numpy)

    device = args.device or ("cuda" if args.impl == "b200" else "cpu")
    constants = refenv.activate(device)
    if args.impl == "b200":
        refenv.swap_in_b200()
    workdir = args.workdir or tempfile.mkdtemp(prefix="pcaa_scripts_")
    os.makedirs(workdir, exist_ok=True)
    os.chdir(workdir)

    import wandb
    import datasets
    import PCAA_ablation
    import inference_PCAA

    train_subjects = list(range(args.classes))
    unseen_subjects = list(range(args.classes, args.classes + args.unseen))

    def generate_splits(train_classes=None, seed=0, safe_mode=False, force_pc_subsampling=0, nmax_points=None, **kw):
        """Stand-in for MSRadarDataset.generate_splits (datasets.py:183-379): same effect on disk (split directories
        deleted and re-created with (30, nmax, 4) float64 crops in the reference's file naming), synthetic content."""
        n = int(nmax_points or constants.NMAX)
        marker = os.path.join(os.path.dirname(constants.GEN_DATA_PATH), "generated_nmax.txt")
        if os.path.exists(marker) and open(marker).read().strip() == str(n) and os.path.isdir(os.path.join(constants.GEN_DATA_PATH, "train")):
            return
        if os.path.isdir(constants.GEN_DATA_PATH):
            shutil.rmtree(constants.GEN_DATA_PATH)
        synth.write_dataset(constants.GEN_DATA_PATH, n, list(train_classes), unseen_subjects,
                            crops_per_track=args.crops_per_track, tracks_per_subject=args.tracks, seed=args.seed)
        with open(marker, "w") as f:
            f.write(str(n))

    datasets.MSRadarDataset.generate_splits = staticmethod(generate_splits)
    inference_PCAA.plot_confusion_matrix_cgaae = lambda k, figures_folder, n_labels, preds, labels, title: (preds, labels.astype(int))
    epoch_logs = []
    orig_init = wandb.init

    def init_and_tap(*a, **k):
        # wandb.init() rebinds the module-level wandb.log to the new run's: wrap it afterwards
        run = orig_init(*a, **k)
        run_log = wandb.log

        def log(d, *aa, **kk):
            epoch_logs.append({key: float(v) for key, v in d.items() if isinstance(v, (int, float)) or hasattr(v, "item")})
            return run_log(d, *aa, **kk)

        wandb.log = log
        return run

    wandb.init = init_and_tap

    # every eval-mode encoder call of the inference procedure, in order: (logits, embedding) -- what the labels are made of
    captured = []

    def tap_encoder(mod, inputs, output):
        if type(mod).__name__ == "CGEncoder" and not mod.training:
            captured.append((output[0].detach().float().cpu().numpy().tolist(), output[1].detach().float().cpu().numpy().tolist()))

    torch.nn.modules.module.register_module_forward_hook(tap_encoder)

    cfg = constants.CONFIG
    ks = [int(x) for x in args.ks.split(",") if x]
    result = {"impl": args.impl, "device": device, "workdir": workdir, "runs": []}
    if args.impl == "b200":
        from opensetgaitrecognition_pcaa_b200 import _lib
        result["lib"] = _lib.version()

    def one_run(model_name, nmax, ks_run, regenerate):
        cfg.update(MODEL_NAME=model_name, TRAIN_CLASSES=train_subjects, EPOCHS=args.epochs, BATCH_SIZE=args.batch, NMAX=nmax,
                   CHECKPOINT_FREQUENCY=1, NOTES="synthetic drop-in run")
        constants.BATCH_SIZE = args.batch
        if regenerate:
            datasets.MSRadarDataset.generate_splits(train_classes=cfg["TRAIN_CLASSES"], seed=0, safe_mode=False, nmax_points=nmax)
        rec = {"model_name": model_name, "nmax": nmax}
        if not args.skip_train:
            torch.manual_seed(args.seed)
            np.random.seed(args.seed)
            del epoch_logs[:]
            t0 = time.time()
            PCAA_ablation.train_variant4(cfg, wandb_mode="disabled", proj_head_on_discriminator=False)
            if device == "cuda":
                torch.cuda.synchronize()
            rec["train_s"] = time.time() - t0
            rec["epochs"] = list(epoch_logs)
            rec["files"] = sorted(os.listdir(os.path.join("models", model_name)))
        if not args.skip_infer:
            t0 = time.time()
            del captured[:]
            inference_PCAA.CGAAE_inference(model_names=[model_name], ks=ks_run, variation=inference_PCAA.VARIATION.V4)
            rec["infer_s"] = time.time() - t0
            rec["inference"] = {}
            # the crops the procedure walked (sequential=True order) and everything its encoder returned
            rec["test_labels"] = [int(v) for v in datasets.MSRadarDataset(constants.SPLIT.TEST, sequential=True).labels]
            rec["unseen_labels"] = [int(v) for v in datasets.MSRadarDataset(constants.SPLIT.UNSEEN, sequential=True).labels]
            rec["means"] = torch.load(os.path.join("models", model_name, "discriminator_means.pt")).float().cpu().numpy().tolist()
            rec["encoder_calls"] = list(captured)
            for k in ks_run:
                with open(os.path.join("models", model_name, f"naive_seq_log_{k}.json")) as f:
                    lg = json.load(f)
                rec["inference"][str(k)] = {
                    "log": lg,
                    "preds": np.load(os.path.join("models", model_name, f"final_preds_{k}.npy")).astype(int).tolist(),
                    "labels": np.load(os.path.join("models", model_name, f"final_labels_{k}.npy")).astype(int).tolist()}
        result["runs"].append(rec)

    if args.pointsubsampling:
        # train_pointsubsampling.py:52-71, one (n_tr, i) cell: for n_points in n_points_subs: NMAX, generate_splits,
        # train_variant4(wandb_mode="disabled"), CGAAE_inference(ks=[1, 2, 4, 6], variation=V4)
        for n_points in [int(x) for x in args.pointsubsampling.split(",")]:
            one_run(f"PCAA_npts_V4_{n_points}.{args.classes}.1", n_points, [1, 2, 4, 6], regenerate=True)
    else:
        one_run(args.model_name, args.nmax, ks, regenerate=not args.skip_train or not os.path.isdir(constants.GEN_DATA_PATH))
    if args.impl == "b200":
        from opensetgaitrecognition_pcaa_b200 import _lib
        result["c_abi_calls"] = _lib.CALLS
    out = json.dumps(result)
    if args.out:
        with open(args.out, "w") as f:
            f.write(out)
    print("RESULT " + out[:1500])


if __name__ == "__main__":
    main()
