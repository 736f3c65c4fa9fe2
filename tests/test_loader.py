"""Crop store / batch feeding against the reference's MSRadarDataset contract (datasets.py:62-76, 381-479)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from opensetgaitrecognition_pcaa_b200 import loader, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def dataset(tmp_path):
    root = str(tmp_path / "generated_dataset")
    synth.write_dataset(root, 50, train_subjects=[1, 4, 7], unseen_subjects=[2, 9], crops_per_track=5, tracks_per_subject=3, seed=3)
    return root


def test_filename_parsers_follow_the_reference_convention():
    f = "crop12_subj7_hands_in_pockets_track07011.npy"
    assert loader.filename2crop(f) == 12 and loader.filename2subj(f) == 7
    assert loader.filename2track(f) == "07011" and loader.filename2scenario(f) == "hands_in_pockets"


@pytest.mark.parametrize("sequential", [False, True])
def test_packed_crops_are_the_getitem_outputs(dataset, sequential):
    d = os.path.join(dataset, "train")
    pk = loader.PackedCrops.from_directory(d, sequential=sequential, pin=False)
    assert len(pk) == 3 * 3 * 5 and pk.pcs.shape == (45, 4, 30, 50) and pk.pcs.dtype == torch.float32
    assert sorted(pk.filenames) == sorted(os.listdir(d))
    # label mapping: enumerate(list(set(subjects)))  (datasets.py:425-462)
    subj = [loader.filename2subj(f) for f in pk.filenames]
    lab = {c: i for i, c in enumerate(list(set(subj)))}
    assert pk.labels.tolist() == [lab[s] for s in subj]
    for i in (0, 7, 44):
        want = torch.permute(torch.from_numpy(np.load(os.path.join(d, pk.filenames[i]), allow_pickle=True)).to(torch.float), (2, 0, 1))
        x, y = pk[i]
        assert torch.equal(x, want) and int(y) == lab[subj[i]]
    if sequential:
        # every (subject, track) is one consecutive run in increasing crop order
        keys = [(loader.filename2subj(f), loader.filename2track(f)) for f in pk.filenames]
        runs = [k for i, k in enumerate(keys) if i == 0 or keys[i - 1] != k]
        assert len(runs) == len(set(keys)) == 9
        for k in set(keys):
            ids = [loader.filename2crop(f) for f, kk in zip(pk.filenames, keys) if kk == k]
            assert ids == sorted(ids) == list(range(5))


def test_scenario_filter(dataset):
    d = os.path.join(dataset, "test")
    pk = loader.PackedCrops.from_directory(d, scenarios=("smartphone",), pin=False)
    assert len(pk) == 3 * 5 and all(loader.filename2scenario(f) == "smartphone" for f in pk.filenames)
    with pytest.raises(FileNotFoundError):
        loader.PackedCrops.from_directory(d, scenarios=("running",), pin=False)


def test_batches_follow_dataloader_order(dataset):
    pk = loader.PackedCrops.from_directory(os.path.join(dataset, "train"), pin=False)
    got = list(pk.batches(16, drop_last=True))
    assert len(got) == 2 and torch.equal(got[1][0], pk.pcs[16:32]) and torch.equal(got[1][1], pk.labels[16:32])
    assert len(list(pk.batches(16))) == 3 and list(pk.batches(16))[-1][0].shape[0] == 13
    g = torch.Generator().manual_seed(11)
    perm = torch.randperm(len(pk), generator=torch.Generator().manual_seed(11))
    for b, (x, y) in enumerate(pk.batches(16, shuffle=True, drop_last=True, generator=g)):
        ix = perm[b * 16:(b + 1) * 16]
        assert torch.equal(x, pk.pcs[ix]) and torch.equal(y, pk.labels[ix])


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_against_the_reference_dataset_class(dataset, tmp_path):
    """The unmodified MSRadarDataset (run in a subprocess, cwd = the synthetic data root) lists the same files in the
    same order with the same labels and returns the same tensors."""
    out = str(tmp_path / "ref.npz")
    code = textwrap.dedent(f"""
        import sys, os, numpy as np, torch
        sys.path.insert(0, {ROOT!r})
        from oracle.refload import load_reference
        constants, _, _ = load_reference()
        constants.GEN_DATA_PATH = {dataset!r}
        import datasets
        res = {{}}
        for seq in (False, True):
            ds = datasets.MSRadarDataset(datasets.SPLIT.TRAIN, sequential=seq)
            res[f"names{{int(seq)}}"] = np.array(ds.filenames)
            res[f"labels{{int(seq)}}"] = np.asarray(ds.labels)
            res[f"x{{int(seq)}}"] = torch.stack([ds[i][0] for i in range(len(ds))]).numpy()
        np.savez({out!r}, **res)
    """)
    env = dict(os.environ, PYTHONHASHSEED="0", WANDB_MODE="disabled")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = np.load(out)
    code2 = textwrap.dedent(f"""
        import sys, numpy as np
        sys.path.insert(0, {ROOT!r})
        from opensetgaitrecognition_pcaa_b200 import loader
        res = {{}}
        for seq in (False, True):
            pk = loader.PackedCrops.from_directory({os.path.join(dataset, "train")!r}, sequential=seq, pin=False)
            res[f"names{{int(seq)}}"] = np.array(pk.filenames)
            res[f"labels{{int(seq)}}"] = pk.labels.numpy()
            res[f"x{{int(seq)}}"] = pk.pcs.numpy()
        np.savez({out + '.ours.npz'!r}, **res)
    """)
    r = subprocess.run([sys.executable, "-c", code2], capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    ours = np.load(out + ".ours.npz")
    for seq in (0, 1):
        assert ours[f"names{seq}"].tolist() == ref[f"names{seq}"].tolist()
        assert np.array_equal(ours[f"labels{seq}"], ref[f"labels{seq}"])
        assert np.array_equal(ours[f"x{seq}"], ref[f"x{seq}"])


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_device_store_gather_and_prefetcher(dataset):
    pk = loader.PackedCrops.from_directory(os.path.join(dataset, "train"))
    assert pk.pcs.is_pinned()
    dv = pk.to_device()
    g = torch.Generator().manual_seed(5)
    idx = torch.randint(0, len(pk), (37,), generator=g)
    x, y = dv.batch(idx.cuda())
    torch.cuda.synchronize()
    assert torch.equal(x.cpu(), pk.pcs[idx]) and torch.equal(y.cpu(), pk.labels[idx])
    # out-of-range indices give zero rows (the kernel cannot raise)
    from opensetgaitrecognition_pcaa_b200 import ops
    bad = ops.gather_rows(dv.pcs, torch.tensor([0, -1, len(pk), 3], device="cuda"))
    assert torch.equal(bad[0], dv.pcs[0]) and float(bad[1].abs().max()) == 0 and float(bad[2].abs().max()) == 0 and torch.equal(bad[3], dv.pcs[3])
    # gather straight into caller-owned buffers
    ox, oy = torch.empty(8, 4, 30, 50, device="cuda"), torch.empty(8, dtype=torch.int64, device="cuda")
    dv.batch(torch.arange(8, 16, device="cuda"), out=(ox, oy))
    assert torch.equal(ox.cpu(), pk.pcs[8:16]) and torch.equal(oy.cpu(), pk.labels[8:16])
    n = sum(xb.shape[0] for xb, _ in dv.epoch(16, shuffle=True, drop_last=True, generator=g))
    assert n == 32
    # prefetcher: same batches, same order, consumer-stream safe
    for depth in (2, 3):
        pf = loader.DevicePrefetcher(pk.batches(8, shuffle=True, generator=torch.Generator().manual_seed(2)), "cuda", depth=depth)
        want = list(pk.batches(8, shuffle=True, generator=torch.Generator().manual_seed(2)))
        perm = torch.randperm(len(pk), generator=torch.Generator().manual_seed(2))
        k = 0
        for xb, yb in pf:
            ix = perm[k * 8:(k + 1) * 8]
            s = float(xb.sum())                           # consumer work on the current stream
            assert torch.equal(xb.cpu(), pk.pcs[ix]) and torch.equal(yb.cpu(), pk.labels[ix]) and np.isfinite(s)
            k += 1
        assert k == len(want) == 6 and pf.h2d_bytes == sum(x.numel() * 4 + y.numel() * 8 for x, y in want)
