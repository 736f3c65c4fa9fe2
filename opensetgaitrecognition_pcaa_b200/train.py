"""Fused PCAA (ablation variant 4 = the paper's model) training step on one B200, optionally data-parallel.

Mirrors one iteration of the reference loop ``PCAA_ablation.py:882-1021`` (encoder forward, WGAN-GP critic step,
decoder + Chamfer + adversarial + cross-entropy generator step, two Adam updates) as a fixed sequence of C-ABI
kernels with no host synchronisation: losses stay on the device until the caller reads them.

All trainable tensors of the generator side (encoder, decoder projection head, decoder dense layers) are re-pointed
into ONE flat fp32 buffer (plus flat gradient / Adam-moment buffers), so the optimizer is a single fused kernel and
the data-parallel gradient exchange is a bucketed NCCL all-reduce over contiguous slices (decoder slice first, while
the PointNet backward is still running).  ``state_dict()`` of the modules keeps the reference's keys.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch

from . import dp, engine, ops
from ._lib import ACT_ELU, EW_ADD


class _Flat:
    """Flat fp32 storage for a list of (name, tensor): parameters become views; same layout for grad / m / v.

    `pad_rows(name, tensor)` (optional) marks 2-D tensors whose rows are stored with a leading dimension rounded up to 8
    elements: the decoder weights [out, in] with odd `in` (1125, 2250, 4500 ...).  Their bf16 shadow is then directly a
    valid TMA / tensor-core operand (16-byte row stride) and their gradient a TMA-storable fp32 matrix -- no per-step
    re-packing.  The parameter is the [out, in] strided view; the pad columns are zero and stay zero (their gradient is
    never written, so Adam leaves them alone)."""

    def __init__(self, named: List, device, group=None, exchanged: bool = False, pad_rows=None, shared_params: bool = False):
        self.names, self.slices = [], {}
        self._group, self._shared = group, shared_params
        off = 0
        for name, t in named:
            ld = None
            if pad_rows is not None and t.dim() == 2 and t.shape[1] % 8 and pad_rows(name, t):
                ld = (t.shape[1] + 7) // 8 * 8
            n = t.numel() if ld is None else t.shape[0] * ld
            self.names.append(name)
            self.slices[name] = (off, n, tuple(t.shape), ld)
            off += (n + 7) // 8 * 8                     # 16-byte alignment of every tensor, also in the bf16 shadow
        self.size = off
        # shared_params: parameters (and their bf16 copy, make_shadow) in symmetric memory too, so that the rank that owns a
        # chunk of the optimizer state can update it and the others pull the result (dp.GradExchange.start_sharded)
        self.p_hdl = self.shadow_hdl = None
        if shared_params:
            self.p, self.p_hdl = dp.alloc_symmetric(off, torch.float32, device, group)
        else:
            self.p = torch.zeros(off, device=device, dtype=torch.float32)
        # the gradient buffer of a data-parallel run may live in symmetric (peer-mapped) memory: dp.PeerExchange
        self.g, self.peer = dp.alloc_exchange_buffer(off, device, group) if exchanged else (torch.zeros(off, device=device, dtype=torch.float32), None)
        self.m = torch.zeros(off, device=device, dtype=torch.float32)
        self.v = torch.zeros(off, device=device, dtype=torch.float32)
        for name, t in named:
            dst = self.view(self.p, name)
            dst.copy_(t.detach())
            t.data = dst                                # the module parameter now aliases the flat buffer
        # Adam step count: authoritative copy on the device (advanced by pcaa_adam_advance inside the step, so a replayed
        # CUDA graph keeps counting), host mirror for bookkeeping; assigning .step sets both
        self.step_dev = torch.zeros(1, device=device, dtype=torch.int32)
        self.coef_dev = torch.zeros(2, device=device, dtype=torch.float32)
        self._step = 0
        self.shadow = None

    @property
    def step(self) -> int:
        return self._step

    @step.setter
    def step(self, n: int) -> None:
        self._step = int(n)
        self.step_dev.fill_(int(n))

    def make_shadow(self):
        """bf16 copy of the whole flat parameter buffer (same indexing); kept current by the fused Adam kernel.  Allocated
        once (collectively in symmetric memory when the parameters are), refreshed in place afterwards."""
        if self.shadow is None:
            if self._shared and self.p_hdl is not None:
                self.shadow, self.shadow_hdl = dp.alloc_symmetric(self.size, torch.bfloat16, self.p.device, self._group)
            else:
                self.shadow = torch.empty(self.size, device=self.p.device, dtype=torch.bfloat16)
        ops.convert_into(self.p, self.shadow)
        return self.shadow

    def view(self, buf, name):
        """The tensor `name` inside a buffer of this layout (a strided [rows, cols] view for row-padded matrices)."""
        o, n, shp, ld = self.slices[name]
        if ld is None:
            return buf[o:o + n].view(shp)
        return buf[o:o + n].view(shp[0], ld)[:, :shp[1]]

    def span(self, names):
        """[start, end) of the flat range covered by `names` (must be contiguous in registration order)."""
        o0 = min(self.slices[n][0] for n in names)
        o1 = max(self.slices[n][0] + (self.slices[n][1] + 7) // 8 * 8 for n in names)
        return o0, o1


class PCAATrainer:
    """encoder/decoder/discriminator: the drop-in modules of models.py (CUDA, fp32).

    config keys (reference constants.CONFIG): LR, B1, B2, GP_WEIGHT, ADV_WEIGHT; optional B2_G = second Adam beta of
    optimizer_G (train_variant3 passes betas=(B1, B1), PCAA_ablation.py:452-456).

    The same step serves three of the reference's loops (fixed prototypes `means`, SeqChamfer + WGAN-GP + CE):
      * variant 4, the paper's PCAA (PCAA_ablation.py:746-1100): encoder with projection head, decoder behind the
        32->64 `decoder_projection_head`;
      * variant 2 = train_CGAAE (train_AAE.py:25-364): `decoder_projection_head=None`, the decoder reads sup_fv;
      * variant 3 (PCAA_ablation.py:392-743): `decoder=None`, no reconstruction term.
      * variant 1 (PCAA_ablation.py:28-378): variant 4 plus `mean_learner` = models.GaussianMeanLearner, whose output
        for the batch's labels replaces the fixed prototypes (forward only: the reference's Variable() detaches it).
    """

    def __init__(self, encoder, decoder, discriminator, decoder_projection_head, means: torch.Tensor, config: dict,
                 process_group=None, mean_learner=None, discriminator_projection_head=None, sync_bn: bool = False):
        if decoder is None and decoder_projection_head is not None:
            raise ValueError("PCAATrainer: a decoder projection head needs a decoder")
        self.enc, self.dec, self.dis, self.gph = encoder, decoder, discriminator, decoder_projection_head
        # train_variant4 builds, "optimises" and saves a 64->32 discriminator_projection_head that its default flag never
        # applies (PCAA_ablation.py:783-786, 933-937): kept only so that checkpoints carry _DPH.pt
        self.dph = discriminator_projection_head
        # variant 1 (PCAA_ablation.py:28-378): the class prototypes are the GaussianMeanLearner's output for the batch's
        # one-hot labels (train-mode BatchNorm1d, so they depend on the batch composition).  In the reference
        # `z = Variable(z0 + mus)` (:186) detaches, so the learner never receives a gradient although optimizer_D lists
        # its parameters: it is a forward-only module here, its BatchNorm running statistics still move
        self.ml = mean_learner
        dev = next(encoder.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("PCAATrainer needs CUDA modules (no CPU fallback)")
        self.dev = dev
        self.cfg = dict(config)
        self.means = means.to(dev).float().contiguous()
        self.C = self.means.shape[0]
        self.pg = process_group
        self.rank, self.world = dp.world_info(process_group)
        # ---- flat generator-side parameters: optimizer_G chain order (PCAA_ablation.py:821-826); the decoder's bn1-4
        # never receive a gradient (models.py:373-385), torch.optim.Adam skips them -> they stay outside the flat range
        named = [("E." + k, p) for k, p in encoder.named_parameters()]
        if decoder_projection_head is not None:
            named += [("GPH." + k, p) for k, p in decoder_projection_head.named_parameters()]
        if decoder is not None:
            named += [("G." + k, p) for k, p in decoder.named_parameters() if k.startswith("dense")]
        # PCAA_DP_SHARD_ADAM=1 (default) with the peer exchange: the decoder span's Adam update is sharded over the ranks
        # (each rank updates 1/world of it and the others pull the updated weights; `exchange_decoder_span` below)
        want_shard = (self.world > 1 and decoder is not None and dp.exchange_mode() == "peer"
                      and os.environ.get("PCAA_DP_SHARD_ADAM", "1") == "1")
        self.G = _Flat(named, dev, process_group, exchanged=True, pad_rows=lambda n, t: n.startswith("G.dense"),
                       shared_params=want_shard)
        self.D = _Flat([("D." + k, p) for k, p in discriminator.named_parameters()], dev, process_group, exchanged=True)
        dec_names = [n for n in self.G.names if n.startswith("G.") or n.startswith("GPH.")]
        self._dec_span = self.G.span(dec_names) if dec_names else None
        self._enc_span = self.G.span([n for n in self.G.names if n.startswith("E.")])
        # the classifier layers (MLP_head, MLP_sup2: the last encoder parameters) only receive a gradient on supervised
        # iterations; torch.optim.Adam skips parameters without one, moments and per-parameter step count included
        # (PCAA_ablation.py:1005-1021 with SUPERVISION_FREQUENCY > 1), so they carry their own Adam step counter
        # [upper, end of the encoder span) = PointNet layers 3, 4, the TCN and the heads: final (and W3 / W4 no longer read) once
        # the data gradient of layer 3 is enqueued -- the data-parallel step exchanges and updates that part while the
        # backward of layers 2 and 1 still runs; only pointnet1 + pointnet2 (1 MB) are exchanged at the very end
        self._enc_upper = self.G.slices["E.pc_block.pointnet3.module.0.weight"][0]
        self.enc_buckets = int(os.environ.get("PCAA_DP_ENC_BUCKETS", "2"))       # 1: one exchange of the whole encoder span
        if self.enc_buckets not in (1, 2):
            raise ValueError("PCAA_DP_ENC_BUCKETS must be 1 or 2")
        self._cls_span = self.G.span([n for n in self.G.names if n.startswith("E.MLP_head.") or n.startswith("E.MLP_sup2.")])
        assert self._cls_span[1] == self._enc_span[1], "classifier layers must close the encoder span"
        self._cls_step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self._cls_coef_dev = torch.zeros(2, device=dev, dtype=torch.float32)
        self._cls_step = 0
        self._cls_diverged = False          # set by the first unsupervised iteration; until then one Adam call covers both
        self.G.make_shadow()
        self._refresh_views()
        self.shard_adam = False
        if want_shard:
            # every rank must take the same path: sharded only if all three symmetric allocations worked everywhere
            ok = torch.tensor([int(self.G.peer is not None and self.G.p_hdl is not None and self.G.shadow_hdl is not None)], device=dev)
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN,
                                         group=None if isinstance(process_group, str) else process_group)
            self.shard_adam = bool(int(ok))
        # gradient exchange (dp.py): decoder-side span first (overlaps the encoder backward), then the encoder span
        self.xG = dp.GradExchange(self.G.g, process_group, side_stream=True, peer=self.G.peer)
        self.xD = dp.GradExchange(self.D.g, process_group, peer=self.D.peer)
        # sync_bn: every encoder BatchNorm normalises with the statistics of the GLOBAL batch (all-reduced sums), so
        # N ranks compute what one process would on the concatenated batch; default is per-rank statistics (torch-DDP
        # semantics).  Eager `step` only: the reductions are NCCL calls between kernels.
        self.bn_sync = None
        if sync_bn and self.world > 1:
            if mean_learner is not None:
                raise NotImplementedError("sync_bn: the variant-1 mean learner's BatchNorm is not synchronised")
            self.bn_sync = engine.BnSync(None if isinstance(process_group, str) else process_group)
        self._one = torch.ones((), device=dev, dtype=torch.float32)
        self._zero_means = torch.zeros_like(self.means)
        # PCAA_WGRAD_OVERLAP=1 runs the PointNet weight-gradient GEMMs on their own stream, concurrently with the
        # BatchNorm-backward passes (engine.pointnet_backward).  Off by default: measured on B200 the step is power
        # capped (sw_power_cap, ~1.6-1.7 GHz under load) and the overlap buys nothing (21.99 vs 22.07 ms at B=256).
        self._wgrad_stream = torch.cuda.Stream(device=dev) if os.environ.get("PCAA_WGRAD_OVERLAP", "0") == "1" else None
        self._graphs: Dict = {}
        self._warm = set()
        # phase_timing = True: step / step_graphed (split graphs) bracket every phase of `_phases` with CUDA events on the
        # main stream; phase_ms() reads them (data-parallel diagnosis: where an N-rank step spends its time)
        self.phase_timing = False
        self._phase_events: List = []
        # data-parallel: never capture a collective (kernel phases -> graphs, exchanges eager in between);
        # PCAA_SPLIT_GRAPHS=1 forces that program structure on one rank (tests)
        # With the copy-engine exchange on BOTH gradient buffers the data-parallel step contains no NCCL call: it can be
        # captured as ONE graph like the single-rank step (cross-rank ordering = the symmetric-memory barriers inside it).
        # PCAA_DP_ONE_GRAPH=0 (default) keeps the split-graph structure; any NCCL exchange forces it.
        all_peer = self.G.peer is not None and self.D.peer is not None
        one_graph = all_peer and os.environ.get("PCAA_DP_ONE_GRAPH", "0") == "1"
        self.split_graphs = (self.world > 1 and not one_graph) or os.environ.get("PCAA_SPLIT_GRAPHS", "0") == "1"

    def _refresh_views(self):
        enc_t = {k: v for k, v in self.enc.named_parameters()}
        enc_t.update({k: v for k, v in self.enc.named_buffers()})
        self.P_E = enc_t
        self.P_G = {k: v for k, v in self.dec.named_parameters()} if self.dec is not None else {}
        self.P_GPH = {k: v for k, v in self.gph.named_parameters()} if self.gph is not None else {}
        self.gb_E = {k: self.G.view(self.G.g, "E." + k) for k, _ in self.enc.named_parameters()}
        self.gb_E["__zeroed__"] = True      # the step zero-fills the whole encoder span of G.g once (engine._zeros_like_param)
        self.gb_G = {k: self.G.view(self.G.g, "G." + k) for k in self.P_G if k.startswith("dense")}
        self.gb_GPH = {k: self.G.view(self.G.g, "GPH." + k) for k in self.P_GPH}
        self.Dw = [self.D.view(self.D.p, f"D.model.{i}.{s}") for i in (0, 2, 4) for s in ("weight", "bias")]
        self.Dg = [self.D.view(self.D.g, f"D.model.{i}.{s}") for i in (0, 2, 4) for s in ("weight", "bias")]
        self._nbt = [v for k, v in self.enc.named_buffers() if k.endswith("num_batches_tracked")]
        # tensor-core operand copies of the decoder weights: a view of the Adam-maintained bf16 shadow when the row
        # length is a multiple of 8 elements (TMA needs 16-byte row strides), else re-packed (padded) every step
        self._enc_wb16 = {}
        for l in (2, 3, 4):
            W = self.P_E[f"pc_block.pointnet{l}.module.0.weight"]
            self._enc_wb16[l] = self.G.view(self.G.shadow, f"E.pc_block.pointnet{l}.module.0.weight").view(W.shape[0], W.shape[1])
        self._tcn_wb16 = {}
        for l in range(1, 7):
            W = self.P_E[f"tc_block.dtc{l}.conv1d.weight"]
            self._tcn_wb16[l] = self.G.view(self.G.shadow, f"E.tc_block.dtc{l}.conv1d.weight").view(W.shape[0], W.shape[1] * 3)
        # every decoder weight's slice of the bf16 shadow is a tensor-core operand as it is (rows padded to 16 bytes: _Flat)
        self._dec_shadow = {l: self.G.view(self.G.shadow, f"G.dense{l}.weight") for l in (range(1, 6) if self.dec is not None else ())}

    def _decoder_weights_bf16(self):
        return self._dec_shadow

    # ------------------------------------------------------------------------------------------------------------
    def _phases(self, pcs: torch.Tensor, gt: torch.Tensor, z0: torch.Tensor, alphas: torch.Tensor, supervised: bool = True):
        """One variant-4 iteration as an ordered list of (kind, fn): "kernels" phases only enqueue C-ABI kernels (and
        torch fills) on the current stream, "exchange" phases are the data-parallel gradient exchanges (NCCL all-reduce,
        on the side stream for the generator spans, each followed by the Adam update of its span).  Running them in
        order IS the step; the split exists so that step_graphed can capture the kernel phases into CUDA graphs and
        keep the collectives outside of them.  Returns (phases, st) -- st["out"] holds the result dict afterwards."""
        cfg = self.cfg
        B = pcs.shape[0]
        S = pcs.shape[1] * pcs.shape[2] * pcs.shape[3]
        gscale = 1.0 / self.world
        b2_g = cfg.get("B2_G", cfg["B2"])
        st: Dict = {}
        if not supervised and not self._cls_diverged:
            self._cls_diverged = True
            self.set_cls_step(self.G._step)                      # from here on the classifier layers count their own steps
        cls_split = self._cls_diverged

        def adam_span(lo, hi, coef=None):
            G = self.G
            ops.adam_flat_dev(G.p[lo:hi], G.g[lo:hi], G.m[lo:hi], G.v[lo:hi], cfg["B1"], b2_g, 1e-8,
                              G.coef_dev if coef is None else coef, gscale, G.shadow[lo:hi])

        def encoder_and_critic():
            self.enc.train(), self.dis.train()
            if self.dec is not None:
                self.dec.train()
            # encoder forward (train-mode BatchNorm; running statistics updated in place)
            st["logits"], st["fv"], st["saved"] = engine.encoder_forward(pcs, self.P_E, True, self.enc.use_projection_head,
                                                                         self._enc_wb16, self._tcn_wb16, bn=self.bn_sync)
            torch._foreach_add_(self._nbt, 1)
            # critic step (PCAA_ablation.py:900-980): one fused kernel forms d_loss and its parameter gradients
            self.D.g.zero_()
            z0_eff, means_eff = z0, self.means
            if self.ml is not None:
                with torch.no_grad():
                    self.ml.train()
                    mus = self.ml(torch.nn.functional.one_hot(gt, self.C).float())
                z0_eff, means_eff = z0 + mus, self._zero_means        # z = z0 + mus enters the kernel as its noise input
            st["d_losses"] = ops.wgangp_dstep(st["fv"], z0_eff, means_eff, gt, alphas.reshape(-1), *self.Dw, cfg["GP_WEIGHT"], self.Dg)

        def exchange_critic_start():
            # the critic's 4.5 k gradients are reduced on the exchange stream WHILE the decoder forward / Chamfer / decoder
            # backward below run: none of them needs the updated critic (only the adversarial term does), so the latency of this
            # small exchange -- two cross-rank barriers, i.e. also the skew between the ranks' encoder forwards -- is hidden
            self.xD.start(0, self.D.size)

        def decoder_forward_and_backward():
            fv = st["fv"]
            if self.dec is None:                                 # variant 3: tot = loss_g + sup (PCAA_ablation.py:640)
                st["rec_loss"] = torch.zeros((), device=self.dev, dtype=torch.float32)
                return
            st["h0"] = engine.linear_forward(fv, self.P_GPH["0.weight"], self.P_GPH["0.bias"], ACT_ELU) if self.gph is not None else fv
            wb = self._decoder_weights_bf16()
            rec, acts = engine.decoder_forward_tc(st["h0"], self.P_G, wb)
            rec4 = rec.view(pcs.shape)
            frame_loss, i1, i2 = ops.chamfer_fwd(rec4, pcs)
            st["rec_loss"] = ops.chamfer_reduce(frame_loss, True)
            drec = ops.chamfer_bwd(rec4, pcs, i1, i2, self._one, True)
            # backward: Chamfer -> decoder (the projection head follows once the adversarial gradient exists)
            st["dh0"], _ = engine.decoder_backward_tc(drec.view(B, S), acts, self.P_G, wb, self.gb_G)

        def exchange_critic_finish():
            self.xD.finish()

        def critic_update_and_generator_losses():
            fv, logits = st["fv"], st["logits"]
            ops.adam_advance(self.D.step_dev, self.D.coef_dev, cfg["LR"], cfg["B1"], cfg["B2"])
            ops.adam_flat_dev(self.D.p, self.D.g, self.D.m, self.D.v, cfg["B1"], cfg["B2"], 1e-8, self.D.coef_dev, gscale)
            # generator step (PCAA_ablation.py:985-1021)
            adv = -float(cfg["ADV_WEIGHT"]) / B
            _, st["dfv"], st["loss_g"] = ops.disc_fwd(fv, gt, *self.Dw, self.C, want_out=False, want_dx=True, dx_scale=adv,
                                                      want_sum=True, out_scale=adv)           # critic already updated (:996)
            # cross-entropy term only on iterations with i % SUPERVISION_FREQUENCY == 0 (PCAA_ablation.py:1005-1018): on the
            # others tot = rec + loss_g, the classifier head receives no gradient (sup_loss is still reported)
            st["sup_loss"], st["dlogits"], st["pred"] = ops.softmax_ce(logits, gt, want_grad=supervised)
            if self.dec is not None:
                dh0 = st.pop("dh0")
                if self.gph is not None:                         # projection head: its gradient + the adversarial one into d fv
                    G_unused: Dict[str, torch.Tensor] = {}
                    engine.linear_backward(dh0, fv, st.pop("h0"), self.P_GPH["0.weight"], "0.weight", "0.bias", G_unused,
                                           self.gb_GPH, dx_out=st["dfv"], dx_acc=True)
                else:                                            # train_CGAAE: the decoder reads sup_fv (train_AAE.py:243)
                    st["dfv"] = ops.ew(EW_ADD, st["dfv"], dh0)
                    st.pop("h0", None)
            ops.adam_advance(self.G.step_dev, self.G.coef_dev, cfg["LR"], cfg["B1"], b2_g)
            if supervised and cls_split:
                ops.adam_advance(self._cls_step_dev, self._cls_coef_dev, cfg["LR"], cfg["B1"], b2_g)

        def exchange_decoder_span():
            # decoder-side gradients (99 % of the bytes) are final: reduce them AND apply their Adam update (HBM bound)
            # on the side stream while the encoder backward (tensor bound) runs on the main one
            if self._dec_span is None:
                return
            if self.shard_adam:
                # ZeRO-1 over NVLink peer memory: reduce-scatter, Adam on this rank's 1/world of the span, all-gather of the
                # updated fp32 weights and their bf16 copies -- 1/world of the optimizer's HBM traffic per GPU and no
                # all-gather of gradients
                self.xG.start_sharded(*self._dec_span, update=adam_span,
                                      bufs=[(self.G.p, self.G.p_hdl), (self.G.shadow, self.G.shadow_hdl)])
            else:
                self.xG.start(*self._dec_span, then=lambda: adam_span(*self._dec_span))

        def encoder_backward_upper():
            self.G.g[self._enc_span[0]:self._enc_span[1]].zero_()     # one fill instead of one per accumulated gradient
            _, st["resume"] = engine.encoder_backward(st["dlogits"], st["dfv"], st["saved"], self.P_E, self.gb_E,
                                                      side=self._wgrad_stream, bn=self.bn_sync, pause_after=3)

        def encoder_backward():
            encoder_backward_upper()
            st.pop("resume")()

        def exchange_encoder_span():
            self.xG.start(*self._enc_span)
            self.xG.finish()

        def encoder_update_whole():
            adam_upper()
            encoder_update()

        def adam_upper():
            if not cls_split:
                adam_span(self._enc_upper, self._enc_span[1])
            else:
                adam_span(self._enc_upper, self._cls_span[0])
                if supervised:
                    adam_span(*self._cls_span, coef=self._cls_coef_dev)

        def exchange_encoder_upper():
            # heads, TCN, PointNet layers 4 and 3 (8.6 of the encoder's 9.6 MB): exchanged and updated on the side stream
            # while the backward of layers 2 and 1 (~2 ms) runs on the main one
            self.xG.start(self._enc_upper, self._enc_span[1], then=adam_upper)

        def encoder_backward_lower():
            st.pop("resume")()

        def exchange_encoder_lower():
            self.xG.start(self._enc_span[0], self._enc_upper)
            self.xG.finish()

        def encoder_update():
            adam_span(self._enc_span[0], self._enc_upper)
            st["out"] = {"rec_loss": st["rec_loss"], "d_loss": st["d_losses"][0], "gp": st["d_losses"][1],
                         "loss_g": st["loss_g"], "sup_loss": st["sup_loss"], "pred": st["pred"], "logits": st["logits"],
                         "fv": st["fv"]}
            st.pop("saved", None)

        head = [("kernels", encoder_and_critic), ("exchange", exchange_critic_start),
                ("kernels", decoder_forward_and_backward), ("exchange", exchange_critic_finish),
                ("kernels", critic_update_and_generator_losses), ("exchange", exchange_decoder_span)]
        if self.enc_buckets == 1:
            return head + [("kernels", encoder_backward), ("exchange", exchange_encoder_span),
                           ("kernels", encoder_update_whole)], st
        return head + [("kernels", encoder_backward_upper), ("exchange", exchange_encoder_upper),
                       ("kernels", encoder_backward_lower), ("exchange", exchange_encoder_lower),
                       ("kernels", encoder_update)], st

    def step(self, pcs: torch.Tensor, gt: torch.Tensor, z0: torch.Tensor, alphas: torch.Tensor,
             supervised: bool = True) -> Dict[str, torch.Tensor]:
        """One variant-4 iteration.  pcs (B,4,30,N) fp32, gt (B,) int64 in [0, C), z0 (B,32) ~ N(0,1) and alphas (B,1) ~ U(0,1)
        are the host RNG draws of PCAA_ablation.py:915-931, 944-948 (already on the device).  `supervised=False` leaves the
        cross-entropy term out of the generator loss (iterations with i % SUPERVISION_FREQUENCY != 0).  Returns device
        scalars.  A label outside [0, C) turns d_loss / sup_loss (and every gradient) NaN -- the reference raises there."""
        phases, st = self._phases(pcs, gt, z0, alphas, supervised)
        for _, fn in phases:
            self._timed(fn.__name__, fn)
        self.D._step += 1
        self.G._step += 1
        self._cls_step += int(supervised and self._cls_diverged)
        return st["out"]

    # ------------------------------------------------------------------------------------------------------------
    def step_graphed(self, pcs: torch.Tensor, gt: torch.Tensor, z0: torch.Tensor, alphas: torch.Tensor,
                     supervised: bool = True) -> Dict[str, torch.Tensor]:
        """`step` replayed from CUDA graphs (captured once per input shape).  The ~200 launches of an iteration become
        one graph launch: no per-kernel host cost, back-to-back kernel scheduling on the device -- what the launch-bound
        small-batch configurations need.  With one rank the whole iteration is ONE graph (the side-stream Adam update
        is a fork inside it); data-parallel, the kernel phases between the gradient exchanges are six graphs that
        share a memory pool and the NCCL all-reduces (+ the decoder span's Adam update on the side stream) are issued
        eagerly between their replays, so no collective is ever captured.
        The first call with a new shape runs eagerly (it also initialises the library's per-kernel attributes), the
        second one captures; every call performs exactly one training iteration.  Inputs are copied into the graphs'
        static buffers unless they already are those buffers (`static_inputs`).  The returned tensors are graph-owned:
        read them before the next call."""
        if self.bn_sync is not None:
            return self.step(pcs, gt, z0, alphas, supervised)      # SyncBN's reductions are NCCL calls: not capturable
        if not supervised and not self._cls_diverged:
            self._cls_diverged = True
            self.set_cls_step(self.G._step)
        key = (tuple(pcs.shape), tuple(z0.shape)) + ((("cls-split",) if self._cls_diverged else ()) if supervised else ("unsupervised",))
        gs = self._graphs.get(key)
        if gs is None:
            if key not in self._warm:
                self._warm.add(key)
                return self.step(pcs, gt, z0, alphas, supervised)
            gs = self._capture(key, (pcs, gt, z0, alphas), supervised)
        for dst, src in zip(gs["in"], (pcs, gt, z0, alphas)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        for (kind, item), name in zip(gs["program"], gs["names"]):
            self._timed(name, item.replay if kind == "graph" else item)
        self.G._step += 1
        self.D._step += 1
        self._cls_step += int(supervised and self._cls_diverged)
        return gs["out"]

    def _timed(self, name, fn):
        if not self.phase_timing:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        self._phase_events.append((name, e0, e1))
        return r

    def phase_ms(self, reset: bool = True) -> Dict[str, float]:
        """Mean milliseconds per phase over the iterations run since the last reset with phase_timing on (synchronises).
        Measured between events on the MAIN stream: a phase that forks work to the side stream (the decoder-span exchange +
        its Adam update) shows only its enqueue cost; the wait for that work shows up in the phase that joins it
        (exchange_encoder_lower / the end of the step)."""
        torch.cuda.synchronize(self.dev)
        acc: Dict[str, List[float]] = {}
        for name, e0, e1 in self._phase_events:
            acc.setdefault(name, []).append(e0.elapsed_time(e1))
        if reset:
            self._phase_events = []
        return {k: sum(v) / len(v) for k, v in acc.items()}

    def static_inputs(self, pcs_shape, z0_shape=None):
        """The captured graph's input buffers (pcs, gt, z0, alphas) for this shape, or None before capture: a loader can
        copy host batches straight into them (no device-to-device staging copy)."""
        B = pcs_shape[0]
        gs = self._graphs.get((tuple(pcs_shape), tuple(z0_shape or (B, self.means.shape[1]))))
        return None if gs is None else gs["in"]

    def _capture(self, key, example, supervised: bool = True):
        from . import _lib
        static_in = tuple(torch.empty_like(t) for t in example)
        for d, s_ in zip(static_in, example):
            d.copy_(s_)
        torch.cuda.synchronize(self.dev)
        calls0 = _lib.CALLS
        phases, st = self._phases(*static_in, supervised)
        program = []
        names = [fn.__name__ for _, fn in phases] if self.split_graphs else ["whole_step_graph"]
        if not self.split_graphs:
            # the exchanges are no collectives here, only the fork / join of the side-stream Adam update: one graph
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _, fn in phases:
                    fn()
            program.append(("graph", g))
        else:
            pool = torch.cuda.graph_pool_handle()
            for kind, fn in phases:
                if kind == "kernels":
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pool):      # capture enqueues the kernels, it does not run them
                        fn()
                    program.append(("graph", g))
                else:
                    program.append(("exchange", fn))          # issued eagerly between the replays
        gs = {"program": program, "names": names, "in": static_in, "out": st["out"], "launches": _lib.CALLS - calls0, "state": st}
        self._graphs[key] = gs
        return gs

    def graph_launches(self, pcs_shape) -> int:
        """C-ABI kernel launches inside the captured graph of this input shape (0 before capture)."""
        for k, gs in self._graphs.items():
            if k[0] == tuple(pcs_shape):
                return gs["launches"]
        return 0

    def set_cls_step(self, n: int) -> None:
        self._cls_step = int(n)
        self._cls_step_dev.fill_(int(n))

    # ------------------------------------------------------------------------------------------------------------
    def _dp_src(self, r: int) -> int:
        g = None if isinstance(self.pg, str) else self.pg
        return r if g is None else torch.distributed.get_global_rank(g, r)

    def _owner_chunks(self):
        """[(owner rank, lo, hi)] of the decoder span under the sharded optimizer; empty otherwise."""
        if not self.shard_adam or self._dec_span is None:
            return []
        return [(r, b, e) for r, (b, e) in enumerate(self.G.peer.chunks(*self._dec_span)) if e > b]

    def reduced_gradient(self) -> torch.Tensor:
        """The generator-side flat gradient of the last iteration summed over the ranks, on every rank (a copy).  With the
        sharded decoder update the reduced gradient exists only chunk-wise on the chunks' owners: assembled here by
        broadcasts (collective; diagnostics and parity checks only, not part of the step)."""
        g = self.G.g.clone()
        grp = None if isinstance(self.pg, str) else self.pg
        for r, b, e in self._owner_chunks():
            torch.distributed.broadcast(g[b:e], src=self._dp_src(r), group=grp)
        return g

    def sync_optimizer_state(self) -> None:
        """Make the Adam moments of the decoder span complete on every rank (sharded optimizer: each rank only advances the
        moments of the chunk it owns).  Collective; called before a checkpoint is written or a snapshot is compared."""
        grp = None if isinstance(self.pg, str) else self.pg
        for r, b, e in self._owner_chunks():
            for buf in (self.G.m, self.G.v):
                torch.distributed.broadcast(buf[b:e], src=self._dp_src(r), group=grp)

    def snapshot(self) -> Dict:
        """Clone of everything one iteration changes: both optimizers' weights / moments / step counts and the encoder's
        BatchNorm running statistics (+ the mean learner's; data-parallel with the sharded decoder update the moments of the
        decoder span are complete only after `sync_optimizer_state()`).  `restore` puts it back (also into another trainer of the
        same architecture: data-parallel parity checks, resumable probes)."""
        snap = {"G": (self.G.p.clone(), self.G.m.clone(), self.G.v.clone(), self.G.step),
                "D": (self.D.p.clone(), self.D.m.clone(), self.D.v.clone(), self.D.step),
                "enc_buffers": [b.clone() for b in self.enc.buffers()],
                "cls_step": self._cls_step if self._cls_diverged else None}
        if self.ml is not None:
            snap["ml"] = {k: v.clone() for k, v in self.ml.state_dict().items()}
        return snap

    def restore(self, snap: Dict) -> None:
        for flat, key in ((self.G, "G"), (self.D, "D")):
            p, m, v, step = snap[key]
            flat.p.copy_(p), flat.m.copy_(m), flat.v.copy_(v)
            flat.step = step
        for b, b0 in zip(self.enc.buffers(), snap["enc_buffers"]):
            b.copy_(b0)
        self._cls_diverged = snap.get("cls_step") is not None
        self.set_cls_step(snap["cls_step"] if self._cls_diverged else 0)
        if self.ml is not None and "ml" in snap:
            self.ml.load_state_dict(snap["ml"])
        ops.convert_into(self.G.p, self.G.shadow)              # in place: captured graphs keep reading this buffer

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def evaluate(self, pcs: torch.Tensor, gt: torch.Tensor):
        """Validation pass of PCAA_ablation.py:1046-1064: eval-mode encoder, decoder, Chamfer, CE, argmax."""
        self.enc.eval()
        logits, fv, _ = engine.encoder_forward(pcs, self.P_E, False, self.enc.use_projection_head, self._enc_wb16,
                                              self._tcn_wb16)
        ce, _, pred = ops.softmax_ce(logits, gt, want_grad=False)
        if self.dec is None:
            return torch.zeros((), device=self.dev), ce, pred
        self.dec.eval()
        h0 = engine.linear_forward(fv, self.P_GPH["0.weight"], self.P_GPH["0.bias"], ACT_ELU) if self.gph is not None else fv
        rec, _ = engine.decoder_forward_tc(h0, self.P_G, self._decoder_weights_bf16())
        fl, _, _ = ops.chamfer_fwd(rec.view(pcs.shape), pcs, want_idx=False)
        return ops.chamfer_reduce(fl, True), ce, pred


def _ckpt_paths(root: str, name: str):
    d = os.path.join(root, "models", name)
    return d, {k: os.path.join(d, f"{name}_{k}.pt") for k in ("E", "G", "D", "GPH", "DPH", "OPT")}


def write_config(trainer: PCAATrainer, config: Optional[dict], model_name: str, root: str = ".") -> str:
    """models/<name>/config.pkl, the first thing every reference trainer writes (train_AAE.py:27-30, PCAA_ablation.py:754-757)
    and the first thing inference_PCAA.CGAAE_inference_setup reads (:63-65: NMAX, TRAIN_CLASSES, MODEL_NAME).  The caller's
    config is pickled as it is, with those three keys filled in from the trainer when absent."""
    import pickle
    cfg = dict(config or {})
    cfg["MODEL_NAME"] = model_name
    cfg.setdefault("NMAX", int(trainer.enc.nmax_points))
    if not cfg.get("TRAIN_CLASSES"):
        cfg["TRAIN_CLASSES"] = list(range(trainer.C))
    if len(cfg["TRAIN_CLASSES"]) != trainer.C:
        raise ValueError(f"config['TRAIN_CLASSES'] has {len(cfg['TRAIN_CLASSES'])} classes, the networks were built for {trainer.C}")
    d = os.path.join(root, "models", model_name)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "config.pkl")
    with open(path, "wb") as f:
        pickle.dump(cfg, f)
    return path


def save_checkpoint(trainer: PCAATrainer, model_name: str, root: str = ".", config: Optional[dict] = None) -> str:
    """Write the files train_variant4 writes (PCAA_ablation.py:754-757, 859-863, 1088-1112): models/<name>/config.pkl,
    <name>_E.pt, _G.pt, _D.pt, _GPH.pt, _DPH.pt (state_dicts with the reference's keys) and discriminator_means.pt, so that
    inference_PCAA.CGAAE_inference_setup loads the folder as it is; plus <name>_OPT.pt with both Adam states (the reference
    does not checkpoint its optimizers; this is what makes a run resumable).  config.pkl is (re)written when `config` is
    given or the file does not exist yet (the trainer's own hyper-parameters then).  Data-parallel with the sharded decoder
    update: call `trainer.sync_optimizer_state()` on ALL ranks first (fit does), the moments are complete only then."""
    from . import utils
    d, paths = _ckpt_paths(root, model_name)
    os.makedirs(d, exist_ok=True)
    if config is not None or not os.path.exists(os.path.join(d, "config.pkl")):
        write_config(trainer, config if config is not None else trainer.cfg, model_name, root)
    utils.save_model(trainer.enc, paths["E"])
    utils.save_model(trainer.dis, paths["D"])
    if trainer.dec is not None:
        utils.save_model(trainer.dec, paths["G"])
    if trainer.gph is not None:
        utils.save_model(trainer.gph, paths["GPH"])
    if trainer.dph is not None:
        utils.save_model(trainer.dph, paths["DPH"])
    torch.save(trainer.means.detach().cpu(), os.path.join(d, "discriminator_means.pt"))
    torch.save({"G": {"m": trainer.G.m.cpu(), "v": trainer.G.v.cpu(), "step": trainer.G.step, "names": trainer.G.names,
                      "cls_step": trainer._cls_step if trainer._cls_diverged else None},
                "D": {"m": trainer.D.m.cpu(), "v": trainer.D.v.cpu(), "step": trainer.D.step, "names": trainer.D.names}},
               paths["OPT"])
    return d


def load_checkpoint(trainer: PCAATrainer, model_name: str, root: str = ".", optimizer: bool = True) -> None:
    """Load what save_checkpoint (or the reference's trainer) wrote into an already built trainer: weights into the flat
    buffers (the parameters are views of them), bf16 operand copies refreshed, Adam states when present."""
    d, paths = _ckpt_paths(root, model_name)
    for key, mod in (("E", trainer.enc), ("D", trainer.dis), ("G", trainer.dec), ("GPH", trainer.gph), ("DPH", trainer.dph)):
        if mod is not None and (key != "DPH" or os.path.exists(paths[key])):
            mod.load_state_dict(torch.load(paths[key], map_location=trainer.dev))
    trainer.G.make_shadow()
    trainer._refresh_views()
    trainer._graphs.clear()                                   # captured graphs hold views of the old shadow
    trainer._warm.clear()
    if optimizer and os.path.exists(paths["OPT"]):
        st = torch.load(paths["OPT"], map_location=trainer.dev)
        for flat, key in ((trainer.G, "G"), (trainer.D, "D")):
            if st[key]["names"] != flat.names:
                raise ValueError(f"load_checkpoint: optimizer state of a different network layout ({key})")
            flat.m.copy_(st[key]["m"])
            flat.v.copy_(st[key]["v"])
            flat.step = int(st[key]["step"])
        cls = st["G"].get("cls_step")
        trainer._cls_diverged = cls is not None
        trainer.set_cls_step(cls if cls is not None else 0)


def fit(trainer: PCAATrainer, train, valid, config: dict, model_name: str, root: str = ".", np_rng=None, torch_gen=None,
        shuffle_gen: Optional[torch.Generator] = None, log=None, graphed: bool = True) -> List[dict]:
    """The epoch loop of the reference trainers (PCAA_ablation.py:866-1112, train_AAE.py:126-364) on the fused path.

    `train` / `valid` are loader.PackedCrops (the reference's MSRadarDataset(SPLIT.TRAIN / VALID)).  Per epoch: the train
    split in shuffled batches of config["BATCH_SIZE"] with drop_last (the DataLoader settings of PCAA_ablation.py:793-799),
    one fused iteration per batch; the validation split unshuffled with drop_last in eval mode (reconstruction loss,
    cross-entropy, accuracy); every CHECKPOINT_FREQUENCY epochs the model is saved when the validation accuracy improved
    (strictly, starting from 0: PCAA_ablation.py:1087-1091).  Returns one dict per epoch with the quantities the reference
    sends to wandb (:1073-1084); `log(epoch_dict)` is called with each.

    Differences by construction: batches are prefetched to the device one ahead; losses and predictions stay on the device
    during an epoch and are read once at its end (the reference reads five .item()s per iteration); data-parallel, every
    rank takes its shard of each global batch and of the global RNG draws (dp.shard_range / dp.global_draws), metrics
    are those of the local shard."""
    from . import loader
    B = int(config["BATCH_SIZE"])
    rank, world = trainer.rank, trainer.world
    sup_freq = int(config.get("SUPERVISION_FREQUENCY", 1))
    if sup_freq < 1:
        raise ValueError("SUPERVISION_FREQUENCY must be >= 1")
    if B % world:
        # the exchanged gradient is the plain mean of the ranks' shard means (1/world inside Adam): unequal shards would
        # weight samples unequally
        raise ValueError(f"BATCH_SIZE {B} is not a multiple of the {world} data-parallel ranks")
    for name, split in (("train", train), ("valid", valid)):
        lab = split.labels
        if len(lab) and (int(lab.min()) < 0 or int(lab.max()) >= trainer.C):
            raise ValueError(f"{name} split: labels must lie in [0, {trainer.C}) (found {int(lab.min())}..{int(lab.max())})")
    if world > 1:
        # every rank must walk the same permutation (each takes its slice of the same global batch) and draw the same
        # global z0 / alphas: take rank 0's generator states
        if shuffle_gen is None:
            shuffle_gen = torch.Generator()
            shuffle_gen.manual_seed(int(config.get("SHUFFLE_SEED", 0)))
        on_dev = trainer.dev if torch.distributed.get_backend(trainer.pg) == "nccl" else "cpu"
        state = shuffle_gen.get_state().to(on_dev)
        torch.distributed.broadcast(state, 0, group=trainer.pg)
        shuffle_gen.set_state(state.cpu())
        if np_rng is None or torch_gen is None:
            # the reference draws z0 / alphas from the GLOBAL numpy / torch generators; N processes cannot share those, so
            # rank 0 picks a seed and every rank builds the same private generators from it
            seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64).to(on_dev)
            torch.distributed.broadcast(seed, 0, group=trainer.pg)
            if np_rng is None:
                import numpy as np
                np_rng = np.random.default_rng(int(seed))
            if torch_gen is None:
                torch_gen = torch.Generator()
                torch_gen.manual_seed(int(seed))
    os.makedirs(os.path.join(root, "models", model_name), exist_ok=True)
    if rank == 0:
        write_config(trainer, config, model_name, root)                                                            # :754-757
    torch.save(trainer.means.detach().cpu(), os.path.join(root, "models", model_name, "discriminator_means.pt"))   # :859-863
    stepfn = trainer.step_graphed if graphed else trainer.step
    lo, hi = dp.shard_range(B, rank, world)
    history: List[dict] = []
    best_valid_accuracy = 0.0
    latent = trainer.means.shape[1]
    for epoch in range(int(config["EPOCHS"])):
        sums = torch.zeros(4, device=trainer.dev, dtype=torch.float64)
        correct = torch.zeros((), device=trainer.dev, dtype=torch.int64)
        n_it = 0
        batches = ((x[lo:hi], y[lo:hi]) for x, y in train.batches(B, shuffle=True, drop_last=True, generator=shuffle_gen))
        for pcs, gt in loader.DevicePrefetcher(batches, trainer.dev):
            z0, alphas = dp.global_draws(B, latent, rank, world, np_rng, torch_gen)
            out = stepfn(pcs, gt, z0.to(trainer.dev, non_blocking=True), alphas.to(trainer.dev, non_blocking=True),
                         supervised=(n_it % sup_freq == 0))                                                        # :1005
            sums += torch.stack([out["rec_loss"], out["sup_loss"], out["d_loss"],
                                 out["rec_loss"] + out["loss_g"] + out["sup_loss"]]).double()
            correct += (out["pred"].long() == gt).sum()
            n_it += 1
        vs = torch.zeros(2, device=trainer.dev, dtype=torch.float64)
        vcorrect = torch.zeros((), device=trainer.dev, dtype=torch.int64)
        n_v = 0
        vbatches = ((x[lo:hi], y[lo:hi]) for x, y in valid.batches(B, shuffle=False, drop_last=True))
        for pcs, gt in loader.DevicePrefetcher(vbatches, trainer.dev):
            rec, ce, pred = trainer.evaluate(pcs, gt)
            vs += torch.stack([rec, ce]).double()
            vcorrect += (pred.long() == gt).sum()
            n_v += 1
        s, v = (sums / max(n_it, 1)).tolist(), (vs / max(n_v, 1)).tolist()          # the epoch's only host reads
        rec_e = {"epoch": epoch, "iterations": n_it,
                 "Reconstruction Loss Train": s[0], "Reconstruction Loss Valid": v[0], "Cross Entropy Loss Train": s[1],
                 "Cross Entropy Loss Valid": v[1], "Discriminator Loss": s[2], "Total Loss Train": s[3],
                 "Train Accuracy": int(correct) / max(n_it * (hi - lo), 1), "Valid Accuracy": int(vcorrect) / max(n_v * (hi - lo), 1),
                 "saved": False}
        if epoch % int(config.get("CHECKPOINT_FREQUENCY", 1)) == 0:
            trainer.sync_optimizer_state()              # collective, so outside the (per-rank) accuracy condition
        if epoch % int(config.get("CHECKPOINT_FREQUENCY", 1)) == 0 and rec_e["Valid Accuracy"] > best_valid_accuracy:
            best_valid_accuracy = rec_e["Valid Accuracy"]
            if rank == 0:
                save_checkpoint(trainer, model_name, root, config)
            rec_e["saved"] = True
        history.append(rec_e)
        if log is not None:
            log(rec_e)
    return history


def build_variant(variant: int, n_classes: int, nmax: int, config: Optional[dict] = None, device="cuda",
                  seed: Optional[int] = None, process_group=None):
    """Construct the networks of ablation variant 1 (PCAA_ablation.py:40-64), 2 (= train_CGAAE, train_AAE.py:36-46),
    3 (PCAA_ablation.py:407-419) or 4 (PCAA_ablation.py:764-786) as the reference does and wrap them in a PCAATrainer."""
    from . import models, utils
    if variant == 4:
        return build_variant4(n_classes, nmax, config, device, seed, process_group)
    if variant == 1:        # PCAA_ablation.py:40-64: variant 4's networks + the mean learner
        tr = build_variant4(n_classes, nmax, config, device, seed, process_group)
        tr.ml = models.GaussianMeanLearner(n_classes).to(device).float()
        return tr
    if variant not in (2, 3):
        raise ValueError("build_variant: ablation variants are 1, 2, 3 and 4")
    if seed is not None:
        torch.manual_seed(seed)
    cfg = dict(LR=1e-4, B1=0.9, B2=0.99, GP_WEIGHT=15, ADV_WEIGHT=1, SUP_LATENT_DIM=32)
    if config:
        cfg.update(config)
    if variant == 3:
        cfg.setdefault("B2_G", cfg["B1"])                       # optimizer_G betas=(B1, B1), PCAA_ablation.py:452-456
    enc = models.CGEncoder(n_out_labels=n_classes, use_projection_head=False, nmax_points=nmax).to(device).float()
    dec = models.CGDecoder(input_dim=cfg["SUP_LATENT_DIM"], nmax_points=nmax).to(device).float() if variant == 2 else None
    dis = models.CGDiscriminator(n_classes).to(device).float()
    means = utils.sample_distant_points(cfg["SUP_LATENT_DIM"], n_classes, 10, 10).float()
    return PCAATrainer(enc, dec, dis, None, means, cfg, process_group)


def build_variant4(n_classes: int, nmax: int, config: Optional[dict] = None, device="cuda", seed: Optional[int] = None,
                   process_group=None, sync_bn: bool = False):
    """Construct the variant-4 networks exactly as PCAA_ablation.py:764-786 does and wrap them in a PCAATrainer."""
    from . import models, utils
    if seed is not None:
        torch.manual_seed(seed)
    cfg = dict(LR=1e-4, B1=0.9, B2=0.99, GP_WEIGHT=15, ADV_WEIGHT=1, SUP_LATENT_DIM=32)
    if config:
        cfg.update(config)
    enc = models.CGEncoder(n_out_labels=n_classes, use_projection_head=True, nmax_points=nmax).to(device).float()
    dec = models.CGDecoder(input_dim=cfg["SUP_LATENT_DIM"] * 2, nmax_points=nmax).to(device).float()
    dis = models.CGDiscriminator(n_classes).to(device).float()
    gph = torch.nn.Sequential(torch.nn.Linear(cfg["SUP_LATENT_DIM"], cfg["SUP_LATENT_DIM"] * 2), torch.nn.ELU()).to(device).float()
    dph = torch.nn.Sequential(torch.nn.Linear(cfg["SUP_LATENT_DIM"] * 2, cfg["SUP_LATENT_DIM"]), torch.nn.ELU()).to(device).float()
    means = utils.sample_distant_points(cfg["SUP_LATENT_DIM"], n_classes, 10, 10).float()
    return PCAATrainer(enc, dec, dis, gph, means, cfg, process_group, discriminator_projection_head=dph, sync_bn=sync_bn)
