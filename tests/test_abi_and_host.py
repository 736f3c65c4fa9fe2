"""CPU: the C-ABI library loads and exports every symbol include/pcaa.h declares; host-side logic of the package
(ctypes signature table, shape contract, module surface / state_dict keys, flat-parameter layout, synthetic data,
no-CPU-fallback behaviour).  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import pcaa_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcaa.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcaa_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from opensetgaitrecognition_pcaa_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def test_header_declares_expected_entry_points():
    names = declared_symbols()
    for want in ("pcaa_version", "pcaa_last_error", "pcaa_gemm_tc", "pcaa_chamfer_fwd", "pcaa_chamfer_bwd",
                 "pcaa_wgangp_dstep", "pcaa_adam_flat", "pcaa_openset_score", "pcaa_openset_vote",
                 "pcaa_bn_elu_meanpool", "pcaa_pointnet_l1_fwd"):
        assert want in names
    assert len(names) >= 35


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/pcaa.h but not exported by libpcaa_sm100.so"


def test_ctypes_table_covers_the_header(lib):
    from opensetgaitrecognition_pcaa_b200 import _lib
    declared = set(declared_symbols()) - {"pcaa_version", "pcaa_last_error", "pcaa_sm_count"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    # argument counts of the ctypes table agree with the header's parameter lists
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, args in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        assert len([a for a in m.group(1).split(",") if a.strip()]) == len(args), name


def test_version_and_argument_errors_without_gpu(lib):
    assert b"sm_100a" in lib.pcaa_version()
    # argument validation happens before any CUDA call: status + message, no exception, no exit
    rc = lib.pcaa_gemm_simt(None, 0, 1, 1, None, 0, 1, 1, None, 0, 1, 1, -1, 4, 4, None, 0, 0, None)
    assert rc == 1                                             # PCAA_ERR_SHAPE
    assert b"gemm_simt" in lib.pcaa_last_error()
    rc = lib.pcaa_gemm_tc(None, 8, 0, None, 8, 0, None, 8, 1, 4, 4, 4, 99, None, None, None, 0, None, None, None, None, None)
    assert rc == 3                                             # PCAA_ERR_UNSUPPORTED
    from opensetgaitrecognition_pcaa_b200 import _lib
    with pytest.raises(RuntimeError, match="gemm_simt"):
        _lib.call("pcaa_gemm_simt", None, 0, 1, 1, None, 0, 1, 1, None, 0, 1, 1, -1, 4, 4, None, 0, 0, None)


def test_no_cpu_fallback():
    """north_star: no CPU fallback -- CPU tensors are rejected loudly at every module entry."""
    from opensetgaitrecognition_pcaa_b200 import models, ops, utils
    enc = models.CGEncoder(2, nmax_points=50, use_projection_head=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(torch.zeros(1, 4, 30, 50))
    with pytest.raises(RuntimeError, match="CUDA"):
        models.CGDecoder(input_dim=64, nmax_points=50)(torch.zeros(1, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        models.CGDiscriminator(2)(torch.zeros(1, 32), torch.zeros(1, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        utils.SeqChamferLoss()(torch.zeros(1, 4, 30, 50), torch.zeros(1, 4, 30, 50))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gemm(torch.zeros(2, 2), torch.zeros(2, 2))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "opensetgaitrecognition_pcaa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "pcaa_oracle" not in txt, f


@pytest.mark.parametrize("C,nmax,head", [(4, 150, True), (2, 50, True), (6, 110, False)])
def test_module_surface_matches_reference_state_dict(C, nmax, head):
    """Same keys / shapes as the reference's CGEncoder / CGDecoder / CGDiscriminator (models.py:232-421)."""
    from opensetgaitrecognition_pcaa_b200 import models
    enc = models.CGEncoder(C, nmax_points=nmax, use_projection_head=head)
    dec = models.CGDecoder(input_dim=64 if head else 32, nmax_points=nmax)
    dis = models.CGDiscriminator(C)
    shapes = O.param_shapes(C, nmax, use_projection_head=head, dec_in=64 if head else 32)
    for pre, m in (("E.", enc), ("G.", dec), ("D.", dis)):
        want = {k[len(pre):]: tuple(v) for k, v in shapes.items() if k.startswith(pre)}
        got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert got == want, pre
    S = 120 * nmax
    assert [dec.dense1.out_features, dec.dense2.out_features, dec.dense3.out_features, dec.dense4.out_features,
            dec.dense5.out_features] == [S // 16, S // 8, S // 4, S // 2, S]
    # decoder bn1-4 exist (state_dict / optimizer contract) although forward never applies them
    assert all(hasattr(dec, f"bn{i}") for i in (1, 2, 3, 4))
    assert isinstance(enc.glob_avg_pool1, torch.nn.AvgPool2d) and enc.glob_avg_pool1.kernel_size == (1, nmax)
    for dead in (models.Encoder, models.Decoder, models.Discriminator):
        with pytest.raises(AttributeError):
            dead(4)


def test_reference_checkpoint_keys_round_trip():
    """A reference-format state_dict (oracle.det_params uses the reference's keys) loads strictly and comes back."""
    from opensetgaitrecognition_pcaa_b200 import models
    p = O.det_params(4, 70, 1)
    enc = models.CGEncoder(4, nmax_points=70, use_projection_head=True)
    sd = {k[2:]: v.clone() for k, v in p.items() if k.startswith("E.")}
    enc.load_state_dict(sd, strict=True)
    back = enc.state_dict()
    assert set(back) == set(sd)
    for k in sd:
        assert torch.equal(back[k], sd[k]), k


def test_flat_parameter_layout_aliases_module_parameters():
    """train._Flat: module parameters become views of one flat buffer (16-byte aligned slots, registration order)."""
    from opensetgaitrecognition_pcaa_b200.train import _Flat
    lin1, lin2 = torch.nn.Linear(5, 3), torch.nn.Linear(3, 7)
    named = [("a." + k, p) for k, p in lin1.named_parameters()] + [("b." + k, p) for k, p in lin2.named_parameters()]
    before = {n: p.detach().clone() for n, p in named}
    fl = _Flat(named, "cpu")
    off = 0
    for n, p in named:
        o, cnt, shp, ld = fl.slices[n]
        assert o == off and o % 8 == 0 and cnt == p.numel() and shp == tuple(p.shape) and ld is None
        off += (cnt + 7) // 8 * 8
        assert torch.equal(p.detach(), before[n])
        assert p.data_ptr() == fl.p.data_ptr() + 4 * o            # aliasing, not a copy
    assert fl.size == off
    fl.p.add_(1.0)
    assert torch.allclose(lin1.weight.detach(), before["a.weight"] + 1.0)
    lo, hi = fl.span(["b.weight", "b.bias"])
    assert lo == fl.slices["b.weight"][0] and hi == fl.size
    assert fl.view(fl.g, "b.weight").shape == (7, 3)
    # row-padded matrices (the decoder weights with an odd input width): leading dimension rounded up to 8 elements, the
    # parameter is the strided [rows, cols] view, pad columns are zero
    lin3 = torch.nn.Linear(13, 6)
    w0 = lin3.weight.detach().clone()
    fp = _Flat([("w", lin3.weight), ("b", lin3.bias)], "cpu", pad_rows=lambda n, t: n == "w")
    o, cnt, shp, ld = fp.slices["w"]
    assert (o, cnt, shp, ld) == (0, 6 * 16, (6, 13), 16) and fp.slices["b"][0] == 96
    assert torch.equal(lin3.weight.detach(), w0) and lin3.weight.stride() == (16, 1) and lin3.weight.data_ptr() == fp.p.data_ptr()
    assert float(fp.p[:96].view(6, 16)[:, 13:].abs().max()) == 0.0
    assert fp.view(fp.g, "w").shape == (6, 13) and fp.view(fp.g, "w").stride() == (16, 1)
    sd = lin3.state_dict()
    lin3.load_state_dict({k: v.clone() + 1 for k, v in sd.items()})              # checkpoints load into the strided views
    assert torch.equal(lin3.weight.detach(), w0 + 1) and float(fp.p[:96].view(6, 16)[:, 13:].abs().max()) == 0.0


def test_synthetic_crops_follow_the_dataset_contract(tmp_path):
    """datasets.py:98-161, 466-479: float64 (30, N, 4) on disk, float32 (4, 30, N) served, per-frame mean removed,
    padded frames repeat real points (exact duplicates)."""
    from opensetgaitrecognition_pcaa_b200 import synth
    crops = synth.synth_crops(3, 50, seed=5)
    assert crops.shape == (3, 30, 50, 4) and crops.dtype == np.float64
    assert np.max(np.abs(crops.mean(axis=2))) < 1e-12
    assert np.array_equal(crops, synth.synth_crops(3, 50, seed=5))
    dup = 0
    for t in range(30):
        fr = crops[0, t]
        dup += fr.shape[0] - np.unique(fr, axis=0).shape[0]
    assert dup > 0
    pcs, gt = synth.synth_batch(3, 50, 4, seed=5)
    assert pcs.shape == (3, 4, 30, 50) and pcs.dtype == torch.float32 and pcs.is_contiguous()
    assert gt.dtype == torch.int64 and int(gt.max()) < 4
    synth.write_dataset(str(tmp_path), 50, [0, 1], [2, 3], crops_per_track=2, tracks_per_subject=1, seed=1)
    for split in ("train", "valid", "test", "unseen"):
        files = sorted(os.listdir(tmp_path / split))
        assert len(files) == 4
        for f in files:
            assert re.fullmatch(r"crop\d+_subj\d+_(free_walk|hands_in_pockets|smartphone)_track[0-9]+\.npy", f)
        a = np.load(tmp_path / split / files[0])
        assert a.shape == (30, 50, 4) and a.dtype == np.float64


def test_sample_distant_points_matches_oracle():
    from opensetgaitrecognition_pcaa_b200 import utils
    for C in (2, 4, 8):
        a = utils.sample_distant_points(32, C, 10, 10)
        b = O.sample_distant_points(32, C, 10, 10)
        assert a.dtype == torch.float64 and torch.equal(a, b)
        assert torch.allclose(a.norm(dim=1), torch.full((C,), 10.0, dtype=torch.float64))


def test_inference_host_logic_matches_sklearn():
    """The host-side pieces of the open-set procedure (ROC / Youden threshold, F1 metrics, log-domain threshold) against
    sklearn, incl. tied scores and scores that underflowed to exactly 0.0."""
    from sklearn.metrics import f1_score, roc_curve
    from opensetgaitrecognition_pcaa_b200 import inference as I
    rng = np.random.default_rng(5)
    for trial in range(20):
        n = int(rng.integers(5, 200))
        labels = (rng.random(n) < 0.6).astype(np.float64)
        labels[0], labels[1] = 0, 1
        scores = np.exp(rng.normal(-40, 25, n))
        if trial % 3 == 0:
            scores[rng.random(n) < 0.3] = 0.0                  # underflowed likelihoods: all tied at 0
        if trial % 4 == 0:
            scores = np.round(scores, 3)                       # more ties
        fpr, tpr, thr = roc_curve(labels, scores)
        assert I.roc_youden_threshold(labels, scores) == thr[np.argmax(tpr - fpr)]
        y = rng.integers(0, 5, n)
        p = np.where(rng.random(n) < 0.7, y, rng.integers(0, 6, n))
        m = I.f1_scores(y, p)
        for avg in ("micro", "macro", "weighted"):
            assert abs(m[f"f1_{avg}"] - f1_score(y, p, average=avg)) < 1e-12
    # `pdf > thr` in float64 <=> `log pdf > log_threshold(thr)`
    assert I.log_threshold(0.0) == I.LOG_MIN_POSITIVE and np.exp(I.LOG_MIN_POSITIVE - 1e-9) == 0.0 < np.exp(I.LOG_MIN_POSITIVE + 0.7)
    assert I.log_threshold(np.inf) == np.inf and abs(I.log_threshold(1e-30) - np.log(1e-30)) < 1e-12


def test_peer_exchange_chunking_and_fallback():
    """Host logic of the copy-engine gradient exchange (dp.PeerExchange): chunks tile the span, are 8-element aligned
    and identical on every rank; without CUDA / a process group the gradient buffer is a plain tensor (NCCL / gloo path)."""
    from opensetgaitrecognition_pcaa_b200 import dp

    class H:                                    # stands in for the symmetric-memory handle
        def __init__(self, rank, world):
            self.rank, self.world_size = rank, world

    class FakeFlat:
        device = torch.device("cpu")

    for world in (2, 3, 8):
        for lo, hi in ((0, 217_000_008), (24, 1_000_000), (8, 16), (0, 8 * world), (16, 16 + 8 * (world - 1))):
            per_rank = []
            for rank in range(world):
                px = dp.PeerExchange.__new__(dp.PeerExchange)
                px.rank, px.world = rank, world
                per_rank.append(px.chunks(lo, hi))
            assert all(c == per_rank[0] for c in per_rank)
            ch = per_rank[0]
            assert len(ch) == world and ch[0][0] == lo and ch[-1][1] == hi
            assert all(a[1] == b[0] for a, b in zip(ch, ch[1:])) and all(b <= e for b, e in ch)
            assert all((b - lo) % 8 == 0 for b, _ in ch)
    assert dp.exchange_mode() in ("peer", "nccl")
    t, peer = dp.alloc_exchange_buffer(64, "cpu")
    assert peer is None and t.shape == (64,) and float(t.abs().sum()) == 0
