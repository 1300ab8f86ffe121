"""Device-side execution engine of the attention-LSTM decoder: packed bf16 weights, HBM workspaces,
the per-step kernel plans for the two cores on the hot path (Att2in2Core, TopDownCore) and the
fully on-device greedy / beam decode loops (captured into CUDA graphs).

Data layout in HBM (R = decoder rows = images x beams, Kx = width of the activation matrix X):

  att2in2  X = [ xt (E) | h (H) ]                                     models/AttModel.py:581-601
           S = X @ [[W_i2h W_h2h],[0 W_h2att]]^T + b   (R, 5H + A)    one GEMM gives the gate sums AND att_h
  topdown  X = [ h_att_prev | xt (E) | fc | h_lang_prev | h_att | ctx ]   models/AttModel.py:430-446
           G1 = X[:, :E+3H] @ W1^T, att_h = X[:, h_att] @ W_h2att^T, G2 = X[:, h_lang_prev:] @ W2^T
  The weight columns are permuted once at pack time so that every GEMM reads a CONTIGUOUS column
  range of X; torch.cat copies of the reference (:432,:438) disappear.
  att (B, L, H) and p_att (B, L, A) are bf16, produced once per batch by the prologue GEMMs and
  shared by all beams of an image.  Recurrent c stays fp32; h is stored as bf16 GEMM operand.
"""
from __future__ import annotations

import collections
import contextlib
import gc
import os
import weakref

import torch

from . import _lib
from ._lib import check, gemm, ptr, stream

BF16 = torch.bfloat16


def _bf16(t):
    """fp32 weight -> bf16 operand through the library's cast kernel."""
    t2 = t.detach().reshape(t.shape[0], -1) if t.dim() > 1 else t.detach().reshape(1, -1)
    return _lib.cast_bf16(t2.contiguous().float())


class PackedWeights:
    """bf16 operand copies of the module parameters in the layouts the step plans need.

    Rebuilt when any parameter's version counter or storage changes (optimizer steps,
    load_state_dict)."""

    def __init__(self, model):
        self.kind = model.kind
        self.all_gates = bool(getattr(model, "all_gates", False))   # att2all2: the context feeds all five gate sums
        sd = {k: v.detach() for k, v in model.state_dict().items()}
        self.signature = PackedWeights.signature_of(model)
        H, E, A = model.rnn_size, model.input_encoding_size, model.att_hid_size
        self.H, self.E, self.A, self.V = H, E, A, model.vocab_size + 1
        f32 = lambda k: sd[k].float().contiguous()
        # Embedding + ReLU folded into the table copy (models/AttModel.py:73-75)
        self.emb_relu = _lib.cast_bf16(f32("embed.0.weight"), relu=True)
        self.use_bn = int(getattr(model, "use_bn", 0))
        if self.use_bn:   # BatchNorm1d in front of the Linear: folded into the operand per batch (DecoderEngine._fold_bn)
            self.w_att_f32, self.b_att_f32 = f32("att_embed.1.weight"), f32("att_embed.1.bias")
            self.w_att_embed = self.b_att_embed = None
        else:
            self.w_att_embed, self.b_att_embed = _bf16(sd["att_embed.0.weight"]), f32("att_embed.0.bias")
        self.w_ctx2att, self.b_ctx2att = _bf16(sd["ctx2att.weight"]), f32("ctx2att.bias")
        # self.logit (models/AttModel.py:86-91): logit_layers - 1 hidden Linear + ReLU (+ Dropout(0.5), inactive in eval) blocks
        n_logit = int(getattr(model, "logit_layers", 1))
        last = "logit" if n_logit == 1 else f"logit.{3 * (n_logit - 1)}"
        self.logit_hidden = [(_bf16(sd[f"logit.{3 * i}.weight"]), f32(f"logit.{3 * i}.bias")) for i in range(n_logit - 1)]
        self.w_logit, self.b_logit = _bf16(sd[last + ".weight"]), f32(last + ".bias")
        self.bn2 = None
        if self.use_bn == 2:   # eval-mode BatchNorm1d(rnn_size) behind att_embed's Linear + ReLU: y * s + t per column
            bn = model.att_embed[4]
            s2 = f32("att_embed.4.weight") * torch.rsqrt(f32("att_embed.4.running_var") + bn.eps)
            self.bn2 = (s2.contiguous(), (f32("att_embed.4.bias") - f32("att_embed.4.running_mean") * s2).contiguous())
        if self.kind in ("att2in2", "topdown"):
            self.w_alpha = f32("core.attention.alpha_net.weight").reshape(-1).contiguous()
            w_h2att, b_h2att = f32("core.attention.h2att.weight"), f32("core.attention.h2att.bias")
        if self.kind == "att2in2":
            self.Kx = E + H
            w1 = torch.zeros(5 * H + A, E + H, device=w_h2att.device)
            w1[:5 * H, :E] = f32("core.i2h.weight")
            w1[:5 * H, E:] = f32("core.h2h.weight")
            w1[5 * H:, E:] = w_h2att
            self.w1 = _bf16(w1)
            # decode loops without dropout: the input-word term i2h(relu(embed(it))) comes from a (V, 5H) table (built on
            # first use, gate_table()), so the step GEMM contracts over the recurrent columns only: X[:, E:] @ w1h^T
            self.w1h = self.w1[:, E:]
            self._gate_table = None
            if self.all_gates:   # Att2all2Core (models/AttModel.py:618-654): a2h (5H, H); its bias joins the gate bias
                self.b1 = torch.cat([f32("core.i2h.bias") + f32("core.h2h.bias") + f32("core.a2h.bias"), b_h2att]).contiguous()
                self.w_a2c, self.b_a2c = _bf16(sd["core.a2h.weight"]), None
            else:
                self.b1 = torch.cat([f32("core.i2h.bias") + f32("core.h2h.bias"), b_h2att]).contiguous()
                self.w_a2c, self.b_a2c = _bf16(sd["core.a2c.weight"]), f32("core.a2c.bias")
        elif self.kind == "topdown":
            self.Kx = E + 5 * H
            self.w_fc, self.b_fc = _bf16(sd["fc_embed.0.weight"]), f32("fc_embed.0.bias")
            w_ih, w_hh = f32("core.att_lstm.weight_ih"), f32("core.att_lstm.weight_hh")
            # reference input order is [h_lang_prev, fc, xt] (:432); X order is [h_att_prev, xt, fc, h_lang_prev]
            self.w1 = _bf16(torch.cat([w_hh, w_ih[:, 2 * H:], w_ih[:, H:2 * H], w_ih[:, :H]], 1))
            self.b1 = (f32("core.att_lstm.bias_ih") + f32("core.att_lstm.bias_hh")).contiguous()
            w_ih2, w_hh2 = f32("core.lang_lstm.weight_ih"), f32("core.lang_lstm.weight_hh")
            # reference input order is [ctx, h_att] (:438); X order is [h_lang_prev, h_att, ctx]
            self.w2 = _bf16(torch.cat([w_hh2, w_ih2[:, H:], w_ih2[:, :H]], 1))
            self.b2 = (f32("core.lang_lstm.bias_ih") + f32("core.lang_lstm.bias_hh")).contiguous()
            self.w_h2att, self.b_h2att = _bf16(w_h2att), b_h2att
        elif self.kind in ("stackatt", "denseatt"):   # models/AttModel.py:458-526
            dense = self.kind == "denseatt"
            self.Kx = E + (7 if dense else 6) * H
            self.w_fc, self.b_fc = _bf16(sd["fc_embed.0.weight"]), f32("fc_embed.0.bias")
            cell = lambda n: (_bf16(torch.cat([f32(f"core.{n}.i2h.weight"), f32(f"core.{n}.h2h.weight")], 1)),
                              (f32(f"core.{n}.i2h.bias") + f32(f"core.{n}.h2h.bias")).contiguous())
            # i2h columns already come in slot order: lstm0 [xt | fc], lstm1 [h0 | ctx1], lstm2 [h1 or f1 | ctx2]; h2h follows
            (self.wl0, self.bl0), (self.wl1, self.bl1), (self.wl2, self.bl2) = cell("lstm0"), cell("lstm1"), cell("lstm2")
            self.w_h2att, self.b_h2att = _bf16(f32("core.att1.h2att.weight")), f32("core.att1.h2att.bias")
            self.w_alpha = f32("core.att1.alpha_net.weight").reshape(-1).contiguous()
            self.w_alpha2 = f32("core.att2.alpha_net.weight").reshape(-1).contiguous()
            # att2's query is h2att2(h1 + emb2(ctx1)) (:481): one GEMM over [ctx1 | h1] with [W2 We | W2], bias W2 be + b2
            w2, we = f32("core.att2.h2att.weight"), f32("core.emb2.weight")
            self.w_q2 = _bf16(torch.cat([w2 @ we, w2], 1))
            self.b_q2 = (torch.mv(w2, f32("core.emb2.bias")) + f32("core.att2.h2att.bias")).contiguous()
            if dense:
                self.w_f1, self.b_f1 = _bf16(sd["core.fusion1.0.weight"]), f32("core.fusion1.0.bias")
                self.w_f2, self.b_f2 = _bf16(sd["core.fusion2.0.weight"]), f32("core.fusion2.0.bias")
        else:
            raise ValueError(self.kind)

    def gate_table(self):
        """att2in2 / att2all2: T[v] = W_i2h relu(Emb[v]) + b_i2h + b_h2h (+ b_a2h), fp32 (V, 5H) -- the whole contribution of
        the input word to the gate sums (models/AttModel.py:160,584).  One 26 GF GEMM per weight version; the decode
        step then adds row it[r] in the cell kernel and its GEMM drops the E input columns (K = E + H -> H)."""
        if self._gate_table is None:
            H, E = self.H, self.E
            self._gate_table = torch.empty(self.V, 5 * H, dtype=torch.float32, device=self.w1.device)
            gemm(self.emb_relu, self.w1[:5 * H, :E], self.b1[:5 * H].contiguous(), out_f32=self._gate_table)
            self.b1h = torch.cat([torch.zeros(5 * H, device=self.b1.device), self.b1[5 * H:]]).contiguous()
        return self._gate_table

    @staticmethod
    def signature_of(model):
        return tuple((p.data_ptr(), p._version) for p in model.parameters())


class Slots:
    """Column ranges of the activation matrix X."""

    def __init__(self, kind, E, H):
        self.fc = None
        if kind == "att2in2":
            self.xt, self.h_out = (0, E), (E, E + H)
            self.gather = [(E, H), (0, 0)]           # recurrent columns re-ordered by beam parent
            self.n_state = 1
            self.h_load = self.h_read = [self.h_out]  # state layer -> slot its previous value is read from / its new value lands in
        elif kind == "topdown":
            self.h_att_prev, self.xt, self.fc = (0, H), (H, H + E), (H + E, 2 * H + E)
            self.h_lang, self.h_att, self.ctx = (2 * H + E, 3 * H + E), (3 * H + E, 4 * H + E), (4 * H + E, 5 * H + E)
            self.h_out = self.h_lang
            self.gather = [(0, H), (2 * H + E, H)]
            self.n_state = 2
            self.h_load, self.h_read = [self.h_att_prev, self.h_lang], [self.h_att, self.h_lang]
        else:
            # stackatt: [xt | fc | h0 | ctx1 | h1 | ctx2 | h2]; denseatt adds f1 = fusion1(h0, h1) in front of ctx2.  Every cell
            # overwrites its own h slot in place (its GEMM has read the previous value by then), so each GEMM operand is one
            # contiguous column range: lstm0 [xt|fc|h0], lstm1 [h0|ctx1|h1], att2 query [ctx1|h1], lstm2 [h1 or f1|ctx2|h2].
            dense = kind == "denseatt"
            self.xt, self.fc, self.h0, self.ctx1, self.h1 = (0, E), (E, E + H), (E + H, E + 2 * H), (E + 2 * H, E + 3 * H), (E + 3 * H, E + 4 * H)
            o = E + 4 * H
            self.f1 = (o, o + H) if dense else None
            o += H if dense else 0
            self.ctx2, self.h2 = (o, o + H), (o + H, o + 2 * H)
            self.h_out = self.h2
            self.gather = [(E + H, 3 * H), (self.h2[0], H)]      # [h0|ctx1|h1] (ctx1 rides along) and h2
            self.n_state = 3
            self.h_load = self.h_read = [self.h0, self.h1, self.h2]


_NVTX = os.environ.get("UIC_NVTX", "0") != "0"


@contextlib.contextmanager
def _nvtx(name):
    """NVTX range around a phase of the path (UIC_NVTX=1; a no-op otherwise): feature prologue, decode-graph replay / first
    eager pass, training forward / backward -- what a timeline (nsys, ncu --nvtx) needs to tell the phases apart."""
    if _NVTX:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if _NVTX:
            torch.cuda.nvtx.range_pop()


@contextlib.contextmanager
def _no_gc():
    """Python's cyclic garbage collector switched off for the duration of a CUDA-graph capture (after one explicit
    collection): an unreachable object graph that holds a torch.cuda.CUDAGraph (an old model's decode graphs, say) would
    otherwise be finalised whenever the collector happens to run, and its cudaGraphExecDestroy is not permitted while a
    stream is capturing -- the capture then fails with cudaErrorStreamCaptureInvalidated."""
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


class ZeroArena:
    """Tables of a decode loop that start every call at zero, carved out of ONE allocation: a single memset per call instead
    of a fill kernel per table (the beam loop had 14 of them, ~3 % of its time)."""

    def __init__(self, device):
        self.device, self.specs, self.buf = device, [], None

    def take(self, name, shape, dtype):
        self.specs.append((name, tuple(shape), dtype))

    def build(self):
        off, views = 0, {}
        layout = []
        for name, shape, dtype in self.specs:
            n = 1
            for d in shape:
                n *= d
            nbytes = n * torch.empty((), dtype=dtype).element_size()
            layout.append((name, shape, dtype, off, nbytes))
            off = (off + nbytes + 255) // 256 * 256
        self.buf = torch.zeros(max(off, 256), dtype=torch.uint8, device=self.device)
        for name, shape, dtype, o, nbytes in layout:
            views[name] = self.buf[o:o + nbytes].view(dtype).view(shape)
        return views

    def zero(self):
        self.buf.zero_()


class Features:
    """Output of the prologue (models/AttModel.py:107-117): bf16 att / p_att tiles (+ fc for topdown)."""

    def __init__(self, att, p_att, fc, masks, B, L):
        self.att, self.p_att, self.fc, self.masks, self.B, self.L = att, p_att, fc, masks, B, L

    @property
    def device(self):
        return self.att.device


class LazyFeatures:
    """Clipped raw inputs of the prologue; `materialize(out)` runs it into the given Features buffers."""

    def __init__(self, engine, fc_feats, att_feats, masks, B, L, drop=None):
        self.engine, self.fc_feats, self.att_feats, self.masks, self.B, self.L = engine, fc_feats, att_feats, masks, B, L
        self.drop = drop

    @property
    def device(self):
        return self.att_feats.device

    def materialize(self, out):
        return self.engine.prepare(self.fc_feats, self.att_feats, self.masks, out=out, clip=False, drop=self.drop)


class DecoderEngine:
    def __init__(self, model):
        _lib.require_device()
        self._model_ref = weakref.ref(model)   # the model owns the engine: no reference cycle, both die by refcount
        self.kind = model.kind
        self._packed = None
        self._graphs = collections.OrderedDict()   # least-recently-used first
        self.max_graphs = 6     # each entry owns static tile buffers (~200 MB at B=256, L=196): keep a handful of shapes
        self.use_graphs = True
        self.fused_vocab = True   # sampling: logit GEMM with fused LSE/top-k statistics (False: write logits + row kernels)
        self.compact_first_step = True   # beam search: step 0 on one row per image (every beam forks from beam 0 there)
        self.use_gate_table = True       # decode loops: input-word gate term from a (V, 5H) table instead of the E columns of the step GEMM
        self._capture_launches = 0
        self._replayed_launches = 0
        self.lib = _lib.load()

    # ---- weights ---------------------------------------------------------------------------------
    def invalidate(self):
        """Drop the packed bf16 operand copies and the decode graphs built from them.  Staleness is normally detected through
        the parameters' (data_ptr, version) signature, which a CUDA-graph replay that updates parameters in place does NOT
        bump: whoever replays such a graph (train_bench.GraphedTrainStep) calls this."""
        self._packed = None
        self._graphs.clear()

    @property
    def w(self):
        if self._packed is None or self._packed.signature != PackedWeights.signature_of(self.model):
            self._packed = PackedWeights(self.model)
            self._graphs.clear()  # captured graphs hold pointers into the old operand copies
        return self._packed

    # ---- prologue ---------------------------------------------------------------------------------
    def prepare(self, fc_feats, att_feats, att_masks=None, keep_inputs=False, lazy=False, out=None, clip=True, drop=None):
        """clip_att + fc_embed + att_embed + ctx2att (models/AttModel.py:99-117).

        lazy=True only clips and returns a LazyFeatures: the decode loops then run the prologue straight into the
        static buffers of their CUDA graph (`out`), instead of producing the tiles here and copying 200 MB across."""
        w = self.w
        if att_masks is not None and clip:  # clip to the longest valid length (:99-105)
            keep = int(att_masks.long().sum(1).max())
            att_feats, att_masks = att_feats[:, :keep], att_masks[:, :keep].contiguous().float()
        B, L, D = att_feats.shape
        H, A = w.H, w.A
        if lazy:
            return LazyFeatures(self, fc_feats, att_feats, att_masks, B, L, drop)
        if att_feats.dtype == BF16:   # a bf16 feature cache is consumed as is (no staging pass, half the H2D bytes)
            x = att_feats.reshape(B * L, D)
            x = x if x.is_contiguous() else x.contiguous()
        else:
            x = _lib.cast_bf16(att_feats.reshape(B * L, D).float() if att_feats.dtype != torch.float32 else att_feats.reshape(B * L, D))
        att = torch.empty(B * L, H, dtype=BF16, device=x.device) if out is None else out.att.view(B * L, H)
        w_ae, b_ae, bn = w.w_att_embed, w.b_att_embed, None
        if w.use_bn:
            w_ae, b_ae, bn = self._fold_bn(x, att_masks, B, L)
        if w.bn2 is not None:
            if self.model.training:
                raise NotImplementedError("use_bn = 2: the second BatchNorm runs with its running statistics only (eval mode)")
            gemm(x, w_ae, b_ae, out_bf16=att, relu=True, a_stream=True, post_scale=w.bn2[0], post_shift=w.bn2[1])
        else:
            gemm(x, w_ae, b_ae, out_bf16=att, relu=True, a_stream=True)   # the raw features are read once
        if att_masks is not None:
            check(self.lib.uic_zero_padded_rows(ptr(att), ptr(att_masks), B, L, H, stream()))
        if drop is not None:   # att_embed's nn.Dropout (training mode, AttModel.py:79-84): ctx2att sees the dropped tile
            _lib.dropout(att, drop, _lib.DROP_ATT)
        # p_att is stored in the exponential operand form E = exp(2 p_att) (bf16, see _lib.ATT_E_SCALE): the step kernel then
        # gets tanh(p_att + att_h) = 1 - 2/(E F + 1) from an FMA and a shared reciprocal (MUFU.TANH is quarter rate)
        p_att = torch.empty(B * L, A, dtype=_lib.TILE_DTYPE, device=x.device) if out is None else out.p_att.view(B * L, A)
        gemm(att, w.w_ctx2att, w.b_ctx2att, out_bf16=p_att, exp_col0=0, exp_scale=_lib.ATT_E_SCALE)
        fc = None
        if self.kind != "att2in2":   # fc_embed (identity for att2in2 / att2all2, models/AttModel.py:674-675)
            fc = torch.empty(B, H, dtype=BF16, device=x.device) if out is None else out.fc
            fc_in = _lib.cast_bf16(fc_feats.float().contiguous())
            gemm(fc_in, w.w_fc, w.b_fc, out_bf16=fc, relu=True)
            if drop is not None:
                _lib.dropout(fc, drop, _lib.DROP_FC)
        if out is not None:
            if att_masks is not None:
                out.masks.copy_(att_masks)
            return out
        feats = Features(att.view(B, L, H), p_att.view(B, L, A), fc, att_masks, B, L)
        if keep_inputs:  # bf16 operand copies of the raw features, needed by the prologue wgrads
            feats.x_in, feats.fc_in = x, (fc_in if self.kind != "att2in2" else None)
            feats.bn = bn
        return feats

    def _fold_bn(self, x, att_masks, B, L):
        """nn.BatchNorm1d(att_feat_size) in front of att_embed's Linear (use_bn = 1, models/AttModel.py:79-80), applied by
        pack_wrapper (:44-53) to the packed valid regions.  y = (x - mean) s + beta with s = gamma / sqrt(var + eps) is
        affine per column, so it folds into the GEMM operand: W' = W diag(s), b' = b + W (beta - mean s).  train(): batch
        statistics from one uic_col_moments pass over the bf16 operand (and the running statistics are updated like
        torch does); eval(): running statistics.  No host synchronisation."""
        w, bn = self.w, self.model.att_embed[0]
        D = x.shape[1]
        if att_masks is None:   # the reference hands the 3-D batch to BatchNorm1d, which reads the region axis as channels
            raise RuntimeError(f"running_mean should contain {L} elements not {D}")
        gamma, beta = bn.weight.detach().float(), bn.bias.detach().float()
        if bn.training and B * L == 1:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {[1, D]}")   # as torch does
        if bn.training or bn.running_mean is None:
            lens = att_masks.sum(1).to(torch.int32)
            mom = torch.zeros(2, D, dtype=torch.float64, device=x.device)
            check(self.lib.uic_col_moments(ptr(x), 1, x.stride(0), ptr(lens), B, L, D, ptr(mom[0]), ptr(mom[1]), stream()))
            n = lens.sum().double()
            mean64 = mom[0] / n
            var64 = (mom[1] / n - mean64 * mean64).clamp_min_(0.0)
            mean, var = mean64.float(), var64.float()
            if bn.training and bn.running_mean is not None:
                with torch.no_grad():
                    bn.num_batches_tracked += 1
                    m = bn.momentum if bn.momentum is not None else 1.0 / bn.num_batches_tracked.double()
                    bn.running_mean.mul_(1.0 - m).add_((mean64 * m).to(bn.running_mean.dtype))
                    bn.running_var.mul_(1.0 - m).add_((var64 * (n / (n - 1.0)) * m).to(bn.running_var.dtype))
        else:
            mean, var = bn.running_mean.float(), bn.running_var.float()
        inv = torch.rsqrt(var + bn.eps)
        s = gamma * inv
        w_bf = _bf16(w.w_att_f32 * s[None, :])
        # b' = b + W beta - W' mean with the ROUNDED operand W': the GEMM then computes W' (x - mean) + ... and the rounding of
        # W' cancels between the two terms even for low-variance columns (large s)
        b_f = (torch.addmv(w.b_att_f32, w.w_att_f32, beta) - torch.mv(w_bf.float(), mean)).contiguous()
        return w_bf, b_f, {"s": s, "inv": inv, "mean": mean, "beta": beta}

    def _feature_buffers(self, feats):
        """Static per-graph copies of the feature tiles (shapes of `feats`, a Features or a LazyFeatures)."""
        w, dev, B, L = self.w, feats.device, feats.B, feats.L
        return {"att": torch.empty(B, L, w.H, dtype=BF16, device=dev), "p_att": torch.empty(B, L, w.A, dtype=_lib.TILE_DTYPE, device=dev),
                "fc": torch.empty(B, w.H, dtype=BF16, device=dev) if self.kind != "att2in2" else None,
                "masks": None if feats.masks is None else torch.empty(B, L, dtype=torch.float32, device=dev)}

    # ---- one decoder step: X, c -> logits -------------------------------------------------------------
    @property
    def model(self):
        return self._model_ref()

    def _workspace(self, R, dev):
        w = self.w
        ws = {}
        if self.kind == "att2in2":
            ws["S"] = torch.empty(R, 5 * w.H + w.A, dtype=torch.float32, device=dev)
            ws["ctx"] = torch.empty(R, w.H, dtype=BF16, device=dev)
            ws["a2c"] = torch.empty(R, 2 * w.H, dtype=torch.float32, device=dev)
        elif self.kind in ("stackatt", "denseatt"):
            ws["S"] = torch.empty(R, 5 * w.H, dtype=torch.float32, device=dev)
            ws["att_h"] = torch.empty(R, w.A, dtype=torch.float32, device=dev)
            if self.kind == "denseatt":
                ws["cat"] = torch.empty(R, 3 * w.H, dtype=BF16, device=dev)      # [h0 | h1 | h2]: operand of the two fusion layers
                ws["out"] = torch.empty(R, w.H, dtype=BF16, device=dev)
        else:
            ws["G"] = torch.empty(R, 4 * w.H, dtype=torch.float32, device=dev)
            ws["att_h"] = torch.empty(R, w.A, dtype=torch.float32, device=dev)
        ws["logits"] = torch.empty(R, w.V, dtype=torch.float32, device=dev)
        return ws

    def core_step(self, X, c, feats, ws, beams=1, X_next=None, c_out=None, h_all=None, alpha=None, tok=None):
        """Runs the recurrent core for one step, in place on X / c unless X_next / c_out are given
        (teacher-forced runs keep every step's operands for backward).  h_all: optional extra bf16
        destination (rows, H) for the step output (time-batched logit operand).  tok: (rows,) int64 input tokens --
        when given (decode loops, no dropout) the input-word term is gathered from PackedWeights.gate_table() and the
        xt columns of X are not read."""
        w, lib, st = self.w, self.lib, stream()
        H, E, A, R = w.H, w.E, w.A, X.shape[0]
        sl = Slots(self.kind, E, H)
        c_out = c if c_out is None else c_out
        Xn = X if X_next is None else X_next
        ldx = X.stride(0)

        def cols(t, rng):
            return t[:, rng[0]:rng[1]]

        if self.kind == "att2in2":
            S = ws["S"]
            table = w.gate_table() if (tok is not None and self.use_gate_table) else None
            if table is not None:
                gemm(X[:, E:], w.w1h, w.b1h, out_f32=S, exp_col0=5 * H, exp_scale=_lib.ATT_F_SCALE)
            else:
                gemm(X, w.w1, w.b1, out_f32=S, exp_col0=5 * H, exp_scale=_lib.ATT_F_SCALE)   # S[:, 5H:] = F = exp(2 att_h)
            _lib.att_step(S[:, 5 * H:], S.stride(0), feats.p_att, feats.att, w.w_alpha, feats.masks, ws["ctx"], H, None, 0, alpha,
                          feats.B, beams, feats.L, A, H)
            if w.all_gates:   # S[:, :5H] += a2h(ctx): the saved sums already hold everything the cell (and its backward) needs
                gemm(ws["ctx"], w.w_a2c, None, out_f32=S[:, :5 * H], accumulate=True)
                a2c = None
            else:
                a2c = ws["a2c"]
                gemm(ws["ctx"], w.w_a2c, w.b_a2c, out_f32=a2c)
            h_dst = cols(Xn, sl.h_out)
            if table is not None:
                check(lib.uic_lstm_maxout_fwd_add(ptr(S), S.stride(0), ptr(a2c), 2 * H, ptr(c[0]), ptr(c_out[0]), None,
                                                  ptr(h_dst), Xn.stride(0), ptr(h_all), h_all.stride(0) if h_all is not None else 0,
                                                  R, H, ptr(table), table.stride(0), ptr(tok), w.V, None, 0, 1, st))
            else:
                check(lib.uic_lstm_maxout_fwd(ptr(S), S.stride(0), ptr(a2c), 2 * H, ptr(c[0]), ptr(c_out[0]), None,
                                              ptr(h_dst), Xn.stride(0), ptr(h_all), h_all.stride(0) if h_all is not None else 0,
                                              R, H, st))
        elif self.kind in ("stackatt", "denseatt"):
            if X_next is not None or alpha is not None:
                raise NotImplementedError(f"{self.kind}: only the in-place (inference) step is built")
            dense = self.kind == "denseatt"
            S, ah, cat = ws["S"], ws["att_h"], ws.get("cat")

            def cell(k0, k1, wl, bl, layer, h_slot, extra):   # 5H maxout cell (FCModel.LSTMCore): GEMM over X[:, k0:k1], h in place
                gemm(X[:, k0:k1], wl, bl, out_f32=S)
                check(lib.uic_lstm_maxout_fwd(ptr(S), 5 * H, None, 0, ptr(c[layer]), ptr(c_out[layer]), None, ptr(cols(X, h_slot)), ldx,
                                              ptr(extra), extra.stride(0) if extra is not None else 0, R, H, st))

            def attend(a0, a1, wq, bq, w_alpha, ctx_slot):     # Attention (AttModel.py:538-558): query GEMM over X[:, a0:a1]
                gemm(X[:, a0:a1], wq, bq, out_f32=ah, exp_col0=0, exp_scale=_lib.ATT_F_SCALE)
                _lib.att_step(ah, A, feats.p_att, feats.att, w_alpha, feats.masks, cols(X, ctx_slot), ldx, None, 0, None,
                              feats.B, beams, feats.L, A, H)

            cell(0, sl.h0[1], w.wl0, w.bl0, 0, sl.h0, cat[:, :H] if dense else None)                                  # :478 / :518
            attend(sl.h0[0], sl.h0[1], w.w_h2att, w.b_h2att, w.w_alpha, sl.ctx1)                                        # :479 / :519
            cell(sl.h0[0], sl.h1[1], w.wl1, w.bl1, 1, sl.h1, cat[:, H:2 * H] if dense else None)                       # :480 / :520
            attend(sl.ctx1[0], sl.h1[1], w.w_q2, w.b_q2, w.w_alpha2, sl.ctx2)                                           # :481 / :521
            if dense:
                gemm(cat[:, :2 * H], w.w_f1, w.b_f1, out_bf16=cols(X, sl.f1), relu=True)                                # fusion1 :522
            last_extra = cat[:, 2 * H:] if dense else h_all
            cell((sl.f1 if dense else sl.h1)[0], sl.h2[1], w.wl2, w.bl2, 2, sl.h2, last_extra)                          # :482 / :522
            if dense:
                out = ws["out"] if h_all is None else h_all
                gemm(cat, w.w_f2, w.b_f2, out_bf16=out, relu=True)                                                       # fusion2 :524
                return out
            return cols(X, sl.h2)
        else:
            G = ws["G"]
            gemm(X[:, :E + 3 * H], w.w1, w.b1, out_f32=G)
            # h_att goes to this step's h_att slot (operand of h2att and of the language LSTM) and to
            # the next step's h_att_prev slot
            check(lib.uic_lstm_cell_fwd(ptr(G), 4 * H, ptr(c[0]), ptr(c_out[0]), None, ptr(cols(X, sl.h_att)), ldx,
                                        ptr(cols(Xn, sl.h_att_prev)), Xn.stride(0), R, H, st))
            gemm(cols(X, sl.h_att), w.w_h2att, w.b_h2att, out_f32=ws["att_h"], exp_col0=0, exp_scale=_lib.ATT_F_SCALE)
            ctx = cols(X, sl.ctx)
            _lib.att_step(ws["att_h"], A, feats.p_att, feats.att, w.w_alpha, feats.masks, ctx, ldx, None, 0, alpha,
                          feats.B, beams, feats.L, A, H)
            G2 = ws.get("G2", G)  # teacher-forced runs keep both gate tensors for backward
            gemm(X[:, E + 2 * H:], w.w2, w.b2, out_f32=G2)
            check(lib.uic_lstm_cell_fwd(ptr(G2), 4 * H, ptr(c[1]), ptr(c_out[1]), None, ptr(cols(Xn, sl.h_lang)), Xn.stride(0),
                                        ptr(h_all), h_all.stride(0) if h_all is not None else 0, R, H, st))
        return cols(Xn, sl.h_out)

    def logit_input(self, h):
        """logit_layers > 1 (models/AttModel.py:89-91): the hidden Linear + ReLU blocks in front of the vocabulary projection
        (eval mode: their Dropout(0.5) is inactive).  h: (rows, H) bf16 view -> (rows, H) bf16."""
        for w_i, b_i in self.w.logit_hidden:
            nxt = torch.empty(h.shape[0], self.w.H, dtype=BF16, device=h.device)
            gemm(h, w_i, b_i, out_bf16=nxt, relu=True)
            h = nxt
        return h

    def logits_of(self, h, out):
        """Vocabulary projection of logit_input(core output)."""
        gemm(h, self.w.w_logit, self.w.b_logit, out_f32=out)

    def _new_state(self, R, dev, feats, beams, X=None, c=None):
        w = self.w
        sl = Slots(self.kind, w.E, w.H)
        X = torch.zeros(R, w.Kx, dtype=BF16, device=dev) if X is None else X
        c = torch.zeros(sl.n_state, R, w.H, dtype=torch.float32, device=dev) if c is None else c
        if sl.fc is not None:  # every beam row of image i carries fc[i]
            idx = torch.arange(R, device=dev, dtype=torch.int64) // beams
            check(self.lib.uic_embed_rows(ptr(feats.fc), w.H, ptr(idx), ptr(X[:, sl.fc[0]:]), X.stride(0), R, w.H, feats.B, stream()))
        return X, c, sl

    def _embed(self, tok, X, sl):
        w = self.w
        check(self.lib.uic_embed_rows(ptr(w.emb_relu), w.E, ptr(tok), ptr(X[:, sl.xt[0]:]), X.stride(0), X.shape[0], w.E, w.V, stream()))

    # ---- greedy (models/AttModel.py:198-253, sample_max = 1) -------------------------------------------
    def greedy(self, feats, seq_length, decoding_constraint=0, temperature=0.0, seed=None, drop=None):
        """sample_max = 1 (temperature 0) or multinomial sampling with `temperature` (models/AttModel.py:231-239):
        the arg-max of the logits perturbed by Gumbel noise, which is a function of (seed, step, row, column).
        `seed`: an int, or None to draw one from torch's CUDA generator (so torch.manual_seed controls it).
        `drop` = (p, device seed tensor): the roll-out runs with the training-mode dropout masks of the teacher-forced
        path (same counter-based draws: embeddings rows t * B + b, core output rows b * (T + 1) + t), as the reference does
        when it samples in train() mode for self-critical training; `feats` must have been prepared with the same drop."""
        w, lib = self.w, self.lib
        B, dev, T = feats.B, feats.device, seq_length
        flags = _lib.SAMPLE_DECODING_CONSTRAINT if decoding_constraint else 0
        temperature = float(temperature)
        if temperature > 0.0 and not self.fused_vocab:
            raise NotImplementedError("multinomial sampling runs in the fused statistics epilogue (engine.fused_vocab)")
        if drop is not None and not self.fused_vocab:
            raise NotImplementedError("roll-outs with dropout run in the fused statistics path (engine.fused_vocab)")
        key = ("greedy", B, feats.L, T, flags, feats.masks is not None, self.fused_vocab, self.use_gate_table, temperature,
               None if drop is None else (float(drop[0]), drop[1].data_ptr()))

        def alloc():
            sl0 = Slots(self.kind, w.E, w.H)
            arena = ZeroArena(dev)
            arena.take("X", (B, w.Kx), BF16)
            arena.take("c", (sl0.n_state, B, w.H), torch.float32)
            arena.take("seq", (B, T), torch.int64)
            arena.take("lp", (B, T), torch.float32)
            arena.take("nunf", (T,), torch.int32)
            arena.take("tok", (B,), torch.int64)
            z = arena.build()
            s = {**self._feature_buffers(feats), **z, "arena": arena,
                 "ws": self._workspace(B, dev),
                 "unf": torch.zeros(B, dtype=torch.uint8, device=dev), "seed": torch.zeros(1, dtype=torch.int64, device=dev),
                 "parts": int(lib.uic_logit_stats_parts(B, w.V))}
            s["stats"] = torch.empty(B, s["parts"], 4, dtype=torch.float32, device=dev)
            s["feats"] = Features(s["att"], s["p_att"], s["fc"], s["masks"], B, feats.L)
            s["X"], s["c"], s["sl"] = self._new_state(B, dev, s["feats"], 1, X=s["X"], c=s["c"])
            s["img_idx"] = torch.arange(B, device=dev, dtype=torch.int64)
            s["h_drop"] = torch.empty(B, w.H, dtype=BF16, device=dev) if drop is not None else None
            return s

        def run(s):
            f = s["feats"]
            X, c, sl, ws = s["X"], s["c"], s["sl"], s["ws"]
            s["arena"].zero()      # X, c, seq, lp, nunf, tok: one memset
            if sl.fc is not None:
                check(lib.uic_embed_rows(ptr(f.fc), w.H, ptr(s["img_idx"]), ptr(X[:, sl.fc[0]:]), X.stride(0), B, w.H, B, stream()))
            self._embed(s["tok"], X, sl)
            xt_view = X[:, sl.xt[0]:sl.xt[0] + w.E]
            if drop is not None:
                _lib.dropout(xt_view, drop, _lib.DROP_XT, row0=0)
            for t in range(T):
                h = self.core_step(X, c, f, ws, tok=s["tok"] if drop is None else None)
                if drop is not None:   # the logit layer sees the dropped output; the recurrent state does not (AttModel.py:431,599)
                    s["h_drop"].copy_(h)
                    h = s["h_drop"]
                    _lib.dropout(h, drop, _lib.DROP_OUT, row0=t, row_stride=T + 1)
                h = self.logit_input(h)
                if self.fused_vocab:
                    # logit GEMM with the statistics epilogue (max / sum-exp / arg-max per column part) + merge:
                    # the (B, V) logits are never written
                    banned = s["seq"][:, t - 1:] if (flags and t > 0) else None
                    check(lib.uic_logit_stats(ptr(h), h.stride(0), ptr(w.w_logit), w.H, ptr(w.b_logit), ptr(banned), T,
                                              ptr(s["stats"]), B, w.V, w.H, 1, 0, temperature, ptr(s["seed"]), t, stream()))
                    # merge of the parts + the next step's embedding rows in one launch
                    xt = X[:, sl.xt[0]:] if t + 1 < T else None
                    check(lib.uic_greedy_advance(ptr(s["stats"]), s["parts"], ptr(s["seq"]), ptr(s["lp"]), ptr(s["unf"]), ptr(s["tok"]),
                                                 ptr(s["nunf"]), t, T, B, ptr(w.emb_relu), w.E, ptr(xt), X.stride(0), w.E, w.V, temperature, ptr(s["seed"]),
                                                 stream()))
                    if drop is not None and t + 1 < T:
                        _lib.dropout(xt_view, drop, _lib.DROP_XT, row0=(t + 1) * B)
                    continue
                else:
                    self.logits_of(h, ws["logits"])
                    check(lib.uic_greedy_step(ptr(ws["logits"]), w.V, ptr(s["seq"]), ptr(s["lp"]), ptr(s["unf"]), ptr(s["tok"]),
                                              ptr(s["nunf"]), t, T, B, w.V, flags, stream()))
                if t + 1 < T:
                    self._embed(s["tok"], X, sl)
            return s["seq"], s["lp"]

        def pre(s):
            if temperature > 0.0:
                if seed is None:
                    s["seed"].random_()
                else:
                    s["seed"].fill_(int(seed))

        return self._decode(key, alloc, run, feats, pre)

    # ---- beam search (models/AttModel.py:167-196 + models/CaptionModel.py:33-177) -------------------------
    def beam(self, feats, seq_length, beam_size, decoding_constraint=0, max_ppl=0):
        w, lib = self.w, self.lib
        B, dev, T, b = feats.B, feats.device, seq_length, beam_size
        R = B * b
        tk_flags = _lib.SAMPLE_DECODING_CONSTRAINT if decoding_constraint else 0
        bs_flags = _lib.BEAM_MAX_PPL if max_ppl else 0
        key = ("beam", B, b, feats.L, T, tk_flags, bs_flags, feats.masks is not None, self.fused_vocab, self.compact_first_step,
               self.use_gate_table)

        def alloc():
            sl0 = Slots(self.kind, w.E, w.H)
            compact = self.fused_vocab and 1 < b <= 8 and self.compact_first_step
            arena = ZeroArena(dev)
            for name in ("X", "X2"):
                arena.take(name, (R, w.Kx), BF16)
            for name in ("c", "c2"):
                arena.take(name, (sl0.n_state, R, w.H), torch.float32)
            for name, shape, dt in (("beam_seq", (B, b, T), torch.int32), ("beam_lp", (B, b, T), torch.float32), ("beam_sum", (B, b), torch.float32),
                                    ("done_seq", (B, b, T), torch.int32), ("done_lp", (B, b, T), torch.float32), ("done_p", (B, b), torch.float64),
                                    ("done_unaug", (B, b), torch.float32), ("done_cnt", (B,), torch.int32), ("tok", (R,), torch.int64)):
                arena.take(name, shape, dt)
            if compact:
                arena.take("X0", (B, w.Kx), BF16)
                arena.take("c0", (sl0.n_state, B, w.H), torch.float32)
            z = arena.build()
            s = {**self._feature_buffers(feats), **z, "arena": arena,
                 "ws": self._workspace(R, dev),
                 "tk_val": torch.empty(R, b, device=dev), "tk_idx": torch.empty(R, b, dtype=torch.int32, device=dev),
                 "parent": torch.zeros(R, dtype=torch.int32, device=dev),
                 "parts": int(lib.uic_logit_stats_parts(R, w.V)), "parts0": int(lib.uic_logit_stats_parts(B, w.V)),
                 "kslots": 1 if b == 1 else 3 if b <= 3 else 5 if b <= 5 else 8}
            s["stats"] = torch.empty(R, s["parts"], int(lib.uic_logit_stats_entry_floats(s["kslots"])), dtype=torch.float32, device=dev)
            s["feats"] = Features(s["att"], s["p_att"], s["fc"], s["masks"], B, feats.L)
            s["X"], s["c"], s["sl"] = self._new_state(R, dev, s["feats"], b, X=s["X"], c=s["c"])
            s["img_idx"] = torch.arange(R, device=dev, dtype=torch.int64) // b
            if compact:
                # the first step reads beam 0 of every image only (rows = 1, CaptionModel.py:56): run it on B rows
                s["X0"], s["c0"], _ = self._new_state(B, dev, s["feats"], 1, X=s["X0"], c=s["c0"])
                s["ws0"] = self._workspace(B, dev)
                s["stats0"] = torch.empty(B, s["parts0"], s["stats"].shape[2], dtype=torch.float32, device=dev)
                s["tok0"] = torch.zeros(B, dtype=torch.int64, device=dev)
                s["img_idx0"] = torch.arange(B, device=dev, dtype=torch.int64)
            return s

        def run(s):
            f, ws, sl = s["feats"], s["ws"], s["sl"]
            s["arena"].zero()      # both state buffers and every beam / done table (and X0, c0): one memset
            bufs = [(s["X"], s["c"]), (s["X2"], s["c2"])]
            if sl.fc is not None:
                for X, _ in bufs:
                    check(lib.uic_embed_rows(ptr(f.fc), w.H, ptr(s["img_idx"]), ptr(X[:, sl.fc[0]:]), X.stride(0), R, w.H, B, stream()))
            (ga, na), (gb, nb) = sl.gather
            t_first = 0
            if "X0" in s:
                X0, c0 = s["X0"], s["c0"]
                if sl.fc is not None:
                    check(lib.uic_embed_rows(ptr(f.fc), w.H, ptr(s["img_idx0"]), ptr(X0[:, sl.fc[0]:]), X0.stride(0), B, w.H, B, stream()))
                self._embed(s["tok0"], X0, sl)   # BOS (AttModel.py:186-190)
                h = self.logit_input(self.core_step(X0, c0, f, s["ws0"], beams=1, tok=s["tok0"]))
                check(lib.uic_logit_stats(ptr(h), h.stride(0), ptr(w.w_logit), w.H, ptr(w.b_logit), None, 1,
                                          ptr(s["stats0"]), B, w.V, w.H, s["kslots"], 1, 0.0, None, 0, stream()))
                Xn, cn = bufs[1]
                check(lib.uic_beam_advance(ptr(s["stats0"]), s["parts0"], s["kslots"], ptr(s["beam_seq"]), ptr(s["beam_lp"]),
                                           ptr(s["beam_sum"]), ptr(s["done_seq"]), ptr(s["done_lp"]), ptr(s["done_p"]),
                                           ptr(s["done_unaug"]), ptr(s["done_cnt"]), ptr(s["parent"]), ptr(s["tok"]), 0, T, B, b,
                                           bs_flags, int(1 < T), ptr(X0), ptr(Xn), Xn.stride(0), ga, na, gb, nb, ptr(c0), ptr(cn),
                                           sl.n_state, w.H, ptr(w.emb_relu), w.E, sl.xt[0], w.E, w.V, 1, stream()))
                t_first = 1
            else:
                self._embed(s["tok"], bufs[0][0], sl)  # BOS for every beam row (AttModel.py:186-190)
            for t in range(t_first, T):
                X, c = bufs[t % 2]
                Xn, cn = bufs[(t + 1) % 2]
                h = self.logit_input(self.core_step(X, c, f, ws, beams=b, tok=s["tok"]))
                if self.fused_vocab and b <= 8:
                    banned = s["tok"] if (tk_flags and t > 0) else None
                    check(lib.uic_logit_stats(ptr(h), h.stride(0), ptr(w.w_logit), w.H, ptr(w.b_logit), ptr(banned), 1,
                                              ptr(s["stats"]), R, w.V, w.H, s["kslots"], 1, 0.0, None, 0, stream()))
                    # merge + beam bookkeeping + state re-ordering + next embeddings in one launch
                    check(lib.uic_beam_advance(ptr(s["stats"]), s["parts"], s["kslots"], ptr(s["beam_seq"]), ptr(s["beam_lp"]),
                                               ptr(s["beam_sum"]), ptr(s["done_seq"]), ptr(s["done_lp"]), ptr(s["done_p"]),
                                               ptr(s["done_unaug"]), ptr(s["done_cnt"]), ptr(s["parent"]), ptr(s["tok"]), t, T, B, b,
                                               bs_flags, int(t + 1 < T), ptr(X), ptr(Xn), X.stride(0), ga, na, gb, nb, ptr(c), ptr(cn),
                                               sl.n_state, w.H, ptr(w.emb_relu), w.E, sl.xt[0], w.E, w.V, b, stream()))
                    continue
                else:
                    self.logits_of(h, ws["logits"])
                    check(lib.uic_row_topk(ptr(ws["logits"]), w.V, ptr(s["tok"]) if (tk_flags and t > 0) else None, ptr(s["tk_val"]),
                                           ptr(s["tk_idx"]), R, w.V, b, tk_flags if t > 0 else 0, stream()))
                check(lib.uic_beam_step(ptr(s["tk_val"]), ptr(s["tk_idx"]), None, ptr(s["beam_seq"]), ptr(s["beam_lp"]), ptr(s["beam_sum"]),
                                        ptr(s["done_seq"]), ptr(s["done_lp"]), ptr(s["done_p"]), ptr(s["done_unaug"]),
                                        ptr(s["done_cnt"]), ptr(s["parent"]), ptr(s["tok"]), t, T, B, b, bs_flags, stream()))
                if t + 1 < T:  # the reference's last get_logprobs_state (CaptionModel.py:171-172) has no observable effect
                    check(lib.uic_beam_gather(ptr(s["parent"]), ptr(X), ptr(Xn), X.stride(0), ga, na, gb, nb, ptr(c), ptr(cn),
                                              sl.n_state, R, w.H, stream()))
                    self._embed(s["tok"], Xn, sl)
            return s["done_seq"], s["done_lp"], s["done_p"], s["done_unaug"], s["done_cnt"]

        return self._decode(key, alloc, run, feats)

    # ---- diverse beam search (group_size > 1, models/CaptionModel.py:33-177 with add_diversity :36-45) ----------
    def beam_diverse(self, feats, seq_length, beam_size, group_size, diversity_lambda=0.5, decoding_constraint=0, max_ppl=0):
        """`group_size` groups of beam_size // group_size beams; group g runs one step behind group g - 1 and its
        candidates lose diversity_lambda per occurrence of their token among the earlier groups' choices at the same
        local step.  Every table has a leading group axis; returns (done_seq, done_lp, done_p, done_unaug, done_cnt)
        shaped (G, B, b', ...)."""
        w, lib = self.w, self.lib
        B, dev, T, G = feats.B, feats.device, seq_length, group_size
        if G < 1 or beam_size % G != 0:
            raise ValueError(f"beam_size={beam_size} must be a multiple of group_size={G}")
        b = beam_size // G
        R = B * b
        if b * G > 16:   # group g needs its b (g + 1) best entries per row to survive the penalties; uic_row_topk keeps 16
            raise NotImplementedError(f"diverse beam search with beam_size={beam_size} > 16 is not built (row top-k keeps 16 candidates)")
        tk_flags = _lib.SAMPLE_DECODING_CONSTRAINT if decoding_constraint else 0
        bs_flags = _lib.BEAM_MAX_PPL if max_ppl else 0
        key = ("beam_diverse", B, b, G, float(diversity_lambda), feats.L, T, tk_flags, bs_flags, feats.masks is not None,
               self.use_gate_table)

        def alloc():
            s = {**self._feature_buffers(feats),
                 "ws": self._workspace(R, dev),
                 "cand_val": torch.empty(R, beam_size, device=dev), "cand_idx": torch.empty(R, beam_size, dtype=torch.int32, device=dev),
                 "tk_val": torch.empty(R, b, device=dev), "tk_un": torch.empty(R, b, device=dev),
                 "tk_idx": torch.empty(R, b, dtype=torch.int32, device=dev),
                 "beam_seq": torch.zeros(G, B, b, T, dtype=torch.int32, device=dev), "beam_lp": torch.zeros(G, B, b, T, device=dev),
                 "beam_sum": torch.zeros(G, B, b, device=dev),
                 "done_seq": torch.zeros(G, B, b, T, dtype=torch.int32, device=dev), "done_lp": torch.zeros(G, B, b, T, device=dev),
                 "done_p": torch.zeros(G, B, b, dtype=torch.float64, device=dev), "done_unaug": torch.zeros(G, B, b, device=dev),
                 "done_cnt": torch.zeros(G, B, dtype=torch.int32, device=dev),
                 "parent": torch.zeros(R, dtype=torch.int32, device=dev), "tok": torch.zeros(G, R, dtype=torch.int64, device=dev)}
            s["feats"] = Features(s["att"], s["p_att"], s["fc"], s["masks"], B, feats.L)
            s["state"] = []
            for _ in range(G):
                X, c, sl = self._new_state(R, dev, s["feats"], b)
                s["state"].append([(X, c), (torch.zeros_like(X), torch.zeros_like(c))])
                s["sl"] = sl
            s["img_idx"] = torch.arange(R, device=dev, dtype=torch.int64) // b
            return s

        def run(s):
            f, ws, sl = s["feats"], s["ws"], s["sl"]
            for k in ("beam_seq", "beam_lp", "beam_sum", "done_seq", "done_lp", "done_p", "done_unaug", "done_cnt", "tok"):
                s[k].zero_()
            for g in range(G):
                for X, c in s["state"][g]:
                    X.zero_()
                    c.zero_()
                    if sl.fc is not None:
                        check(lib.uic_embed_rows(ptr(f.fc), w.H, ptr(s["img_idx"]), ptr(X[:, sl.fc[0]:]), X.stride(0), R, w.H, B, stream()))
                self._embed(s["tok"][g], s["state"][g][0][0], sl)   # BOS (AttModel.py:186-190)
            (ga, na), (gb, nb) = sl.gather
            for t in range(T + G - 1):
                for g in range(G):
                    lt = t - g                                      # the group's own clock (CaptionModel.py:122-124)
                    if lt < 0 or lt >= T:
                        continue
                    X, c = s["state"][g][lt % 2]
                    Xn, cn = s["state"][g][(lt + 1) % 2]
                    tok = s["tok"][g]
                    h = self.logit_input(self.core_step(X, c, f, ws, beams=b, tok=tok))
                    self.logits_of(h, ws["logits"])
                    kp = b * (g + 1)                                # enough to survive the penalties on <= g * b tokens
                    check(lib.uic_row_topk(ptr(ws["logits"]), w.V, ptr(tok) if (tk_flags and lt > 0) else None, ptr(s["cand_val"]),
                                           ptr(s["cand_idx"]), R, w.V, kp, tk_flags if lt > 0 else 0, stream()))
                    check(lib.uic_diverse_select(ptr(s["cand_val"]), ptr(s["cand_idx"]), kp, ptr(s["beam_seq"]), g, B, b, T, lt,
                                                 float(diversity_lambda), ptr(s["tk_val"]), ptr(s["tk_un"]), ptr(s["tk_idx"]), stream()))
                    check(lib.uic_beam_step(ptr(s["tk_val"]), ptr(s["tk_idx"]), ptr(s["tk_un"]), ptr(s["beam_seq"][g]),
                                            ptr(s["beam_lp"][g]), ptr(s["beam_sum"][g]), ptr(s["done_seq"][g]), ptr(s["done_lp"][g]),
                                            ptr(s["done_p"][g]), ptr(s["done_unaug"][g]), ptr(s["done_cnt"][g]), ptr(s["parent"]),
                                            ptr(tok), lt, T, B, b, bs_flags, stream()))
                    if lt + 1 < T:
                        check(lib.uic_beam_gather(ptr(s["parent"]), ptr(X), ptr(Xn), X.stride(0), ga, na, gb, nb, ptr(c), ptr(cn),
                                                  sl.n_state, R, w.H, stream()))
                        self._embed(tok, Xn, sl)
            return s["done_seq"], s["done_lp"], s["done_p"], s["done_unaug"], s["done_cnt"]

        return self._decode(key, alloc, run, feats)

    def _decode(self, key, alloc, run, feats, pre=None):
        """First call per shape: allocate the static buffers, run the loop eagerly once (warms every
        kernel), then capture it into a CUDA graph; later calls only refresh the feature tiles and
        replay.  No host synchronisation happens inside the loop either way."""
        if not self.use_graphs:
            s = alloc()
            self._load_feats(s, feats)
            if pre is not None:
                pre(s)
            return run(s)
        entry = self._graphs.get(key)
        if entry is not None:
            self._graphs.move_to_end(key)
        if entry is None:
            s = alloc()
            self._load_feats(s, feats)
            if pre is not None:
                pre(s)
            n0 = _lib.launch_count()
            run(s)
            n_kernels = _lib.launch_count() - n0      # library launches inside one decode loop
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with _no_gc():          # a CUDAGraph / tensor finaliser run by the cyclic GC mid-capture invalidates the capture
                with torch.cuda.graph(g):
                    out = run(s)
            self._capture_launches += n_kernels        # issued into the graph, not onto the device
            entry = (g, s, out, n_kernels)
            self._graphs[key] = entry
            while len(self._graphs) > self.max_graphs:   # e.g. att_masks clip every batch to a different length
                self._graphs.popitem(last=False)
        g, s, out, n_kernels = entry
        with _nvtx("uic.prologue"):
            self._load_feats(s, feats)
        if pre is not None:
            pre(s)      # per-call device-side inputs of the captured loop (the sampling seed)
        with _nvtx(f"uic.decode_graph[{key[0] if isinstance(key, tuple) else key}]"):
            g.replay()
        self._replayed_launches += n_kernels
        return out

    def launches(self):
        """Kernels of this library executed on the device so far (eager launches + graph replays)."""
        return _lib.launch_count() - self._capture_launches + self._replayed_launches

    @staticmethod
    def _load_feats(s, feats):
        if isinstance(feats, LazyFeatures):
            feats.materialize(s["feats"])
            return
        s["att"].copy_(feats.att)
        s["p_att"].copy_(feats.p_att)
        if feats.fc is not None:
            s["fc"].copy_(feats.fc)
        if feats.masks is not None:
            s["masks"].copy_(feats.masks)

    # ---- teacher-forced forward (models/AttModel.py:119-156), no autograd ---------------------------------
    @torch.no_grad()
    def teacher_forced_logits(self, feats, seq, n_steps):
        """Runs `n_steps` teacher-forced steps and returns the fp32 logits (B, T_total, V) view of the
        time-batched logit GEMM output (rows b*T_total + t); steps >= n_steps are left untouched."""
        w, lib = self.w, self.lib
        B, dev = feats.B, feats.att.device
        T_total = seq.size(1) - 1
        X, c, sl = self._new_state(B, dev, feats, 1)
        ws = self._workspace_tf(B, dev)
        h_all = torch.zeros(B, T_total, w.H, dtype=BF16, device=dev)
        seq = seq.contiguous()
        for t in range(n_steps):
            self._embed(seq[:, t].contiguous(), X, sl)
            self.core_step(X, c, feats, ws, h_all=h_all[:, t])
        logits = torch.empty(B * T_total, w.V, dtype=torch.float32, device=dev)
        gemm(self.logit_input(h_all.view(B * T_total, w.H)), w.w_logit, w.b_logit, out_f32=logits)
        return logits.view(B, T_total, w.V)

    def _workspace_tf(self, R, dev):
        ws = self._workspace(R, dev)
        del ws["logits"]
        return ws
