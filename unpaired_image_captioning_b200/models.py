"""Host-side mirror of the reference's decoder interface (same names, arguments and error behaviour)
with the arithmetic running in the sm_100a kernels of libuic_b200.so.

Reference surface reproduced here:
  models/__init__.py:22-59      setup(opt)
  models/CaptionModel.py:27-31  CaptionModel.forward(*args, mode=...)
  models/AttModel.py:55-253     AttModel (_prepare_feature, _forward, get_logprobs_state, _sample, _sample_beam)
  models/AttModel.py:529-601    Attention, Att2in2Core        :421-446  TopDownCore
  models/AttModel.py:670-690    Att2in2Model, TopDownModel
Parameter names and shapes are identical to the reference, so `state_dict()`s are interchangeable
(`model_i2t-best.pth` loads, trainer.py:102).

There is no PyTorch/CPU fallback: modules must live on a CUDA (sm_100) device to be called.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr, stream
from .engine import BF16, DecoderEngine, Features, Slots

HOT_PATH_MODELS = ("att2in2", "att2all2", "topdown", "stackatt", "denseatt")


class CaptionModel(nn.Module):
    """models/CaptionModel.py:19-31 -- `mode` keyword dispatch to `_forward` / `_sample`."""

    def forward(self, *args, **kwargs):
        mode = kwargs.pop("mode", "forward")
        return getattr(self, "_" + mode)(*args, **kwargs)


class Attention(nn.Module):
    """models/AttModel.py:529-558."""

    def __init__(self, opt):
        super().__init__()
        self.rnn_size = opt.rnn_size
        self.att_hid_size = opt.att_hid_size
        self.h2att = nn.Linear(self.rnn_size, self.att_hid_size)
        self.alpha_net = nn.Linear(self.att_hid_size, 1)

    def forward(self, h, att_feats, p_att_feats, att_masks=None):
        _lib.require_device()
        lib = _lib.load()
        R, H, A = h.size(0), att_feats.size(-1), self.att_hid_size
        att = att_feats.reshape(-1, att_feats.numel() // att_feats.size(0) // H, H)
        n_img, L = att.size(0), att.size(1)
        if R % n_img:
            raise ValueError(f"Attention: {R} rows do not divide over {n_img} images")
        att_b = att if att.dtype == BF16 else _lib.cast_bf16(att.reshape(-1, H).float())
        # p_att_feats is the ctx2att output like in the reference; the kernel takes it in the exponential operand form
        p3 = p_att_feats.reshape(n_img, L, A)
        p_b = _lib.exp_tile(p3).contiguous()
        att_h = torch.empty(R, A, device=h.device)   # F = 16 exp(2 (h2att(h) + b))
        _lib.gemm(_lib.cast_bf16(h.float().contiguous()), _lib.cast_bf16(self.h2att.weight.detach()),
                  self.h2att.bias.detach().float().contiguous(), out_f32=att_h, exp_col0=0, exp_scale=_lib.ATT_F_SCALE)
        ctx = torch.empty(R, H, device=h.device)
        masks = None if att_masks is None else att_masks.reshape(n_img, L).float().contiguous()
        w = self.alpha_net.weight.detach().float().reshape(-1).contiguous()
        _lib.att_step(att_h, A, p_b, att_b.reshape(n_img, L, H).contiguous(), w, masks, None, 0, ctx, H, None, n_img, R // n_img, L, A, H)
        return ctx


class _CoreBase(nn.Module):
    """Shared plumbing: a core step through the engine from fp32 API tensors."""

    def forward(self, xt, fc_feats, att_feats, p_att_feats, state, att_masks=None):
        return self._owner()._core_api(xt, fc_feats, att_feats, p_att_feats, state, att_masks)


class Att2in2Core(_CoreBase):
    """models/AttModel.py:561-601 (parameters a2c, i2h, h2h, attention.*)."""

    def __init__(self, opt):
        super().__init__()
        self.rnn_size = opt.rnn_size
        self.a2c = nn.Linear(opt.rnn_size, 2 * opt.rnn_size)
        self.i2h = nn.Linear(opt.input_encoding_size, 5 * opt.rnn_size)
        self.h2h = nn.Linear(opt.rnn_size, 5 * opt.rnn_size)
        self.dropout = nn.Dropout(opt.drop_prob_lm)
        self.attention = Attention(opt)


class Att2all2Core(_CoreBase):
    """models/AttModel.py:618-654 (parameters a2h, i2h, h2h, attention.*): the attended context feeds all five gate sums."""

    def __init__(self, opt):
        super().__init__()
        self.rnn_size = opt.rnn_size
        self.a2h = nn.Linear(opt.rnn_size, 5 * opt.rnn_size)
        self.i2h = nn.Linear(opt.input_encoding_size, 5 * opt.rnn_size)
        self.h2h = nn.Linear(opt.rnn_size, 5 * opt.rnn_size)
        self.dropout = nn.Dropout(opt.drop_prob_lm)
        self.attention = Attention(opt)


class TopDownCore(_CoreBase):
    """models/AttModel.py:421-446 (parameters att_lstm.*, lang_lstm.*, attention.*)."""

    def __init__(self, opt, use_maxout=False):
        super().__init__()
        self.drop_prob_lm = opt.drop_prob_lm
        self.att_lstm = nn.LSTMCell(opt.input_encoding_size + opt.rnn_size * 2, opt.rnn_size)
        self.lang_lstm = nn.LSTMCell(opt.rnn_size * 2, opt.rnn_size)
        self.attention = Attention(opt)


class LSTMCore(nn.Module):
    """models/FCModel.py:14-42: the 5H maxout cell (parameters i2h, h2h)."""

    def __init__(self, input_size, rnn_size, drop_prob_lm):
        super().__init__()
        self.i2h = nn.Linear(input_size, 5 * rnn_size)
        self.h2h = nn.Linear(rnn_size, 5 * rnn_size)
        self.dropout = nn.Dropout(drop_prob_lm)


class StackAttCore(_CoreBase):
    """models/AttModel.py:458-486 (parameters att1.*, att2.*, lstm0..2.*, emb2.*)."""

    def __init__(self, opt):
        super().__init__()
        E, H = opt.input_encoding_size, opt.rnn_size
        self.att1, self.att2 = Attention(opt), Attention(opt)
        self.lstm0 = LSTMCore(E + H, H, opt.drop_prob_lm)
        self.lstm1, self.lstm2 = LSTMCore(2 * H, H, opt.drop_prob_lm), LSTMCore(2 * H, H, opt.drop_prob_lm)
        self.emb2 = nn.Linear(H, H)


class DenseAttCore(StackAttCore):
    """models/AttModel.py:489-526: StackAttCore + fusion1 (h_0, h_1 -> third cell) and fusion2 (h_0, h_1, h_2 -> output)."""

    def __init__(self, opt):
        super().__init__(opt)
        H = opt.rnn_size
        self.fusion1 = nn.Sequential(nn.Linear(2 * H, H), nn.ReLU(), nn.Dropout(opt.drop_prob_lm))
        self.fusion2 = nn.Sequential(nn.Linear(3 * H, H), nn.ReLU(), nn.Dropout(opt.drop_prob_lm))


class AttModel(CaptionModel):
    """models/AttModel.py:55-253."""

    kind = None

    def __init__(self, opt):
        super().__init__()
        self.vocab_size = opt.vocab_size
        self.input_encoding_size = opt.input_encoding_size
        self.rnn_size = opt.rnn_size
        self.num_layers = opt.num_layers
        self.drop_prob_lm = opt.drop_prob_lm
        self.seq_length = opt.seq_length
        self.fc_feat_size = opt.fc_feat_size
        self.att_feat_size = opt.att_feat_size
        self.att_hid_size = opt.att_hid_size
        self.use_bn = getattr(opt, "use_bn", 0)
        if self.use_bn not in (0, 1, 2):
            raise ValueError(f"use_bn={self.use_bn}")
        self.logit_layers = int(getattr(opt, "logit_layers", 1))
        if self.logit_layers < 1:
            raise ValueError(f"logit_layers={self.logit_layers}")
        # use_bn = 2 (second BatchNorm behind att_embed) and logit_layers > 1 run in eval() mode (inference); their backward
        # passes are not built, so training calls raise (see _require_training_path)
        self.inference_only = bool(getattr(self, "inference_only", False) or self.use_bn == 2 or self.logit_layers > 1)
        for name, v in (("rnn_size", self.rnn_size), ("input_encoding_size", self.input_encoding_size),
                        ("att_hid_size", self.att_hid_size), ("att_feat_size", self.att_feat_size),
                        ("fc_feat_size", self.fc_feat_size)):
            if v % 8:
                raise ValueError(f"{name}={v} must be a multiple of 8 (16-byte bf16 rows for TMA)")
        self.ss_prob = 0.0
        self.embed = nn.Sequential(nn.Embedding(self.vocab_size + 1, self.input_encoding_size), nn.ReLU(),
                                   nn.Dropout(self.drop_prob_lm))
        self.fc_embed = nn.Sequential(nn.Linear(self.fc_feat_size, self.rnn_size), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        # use_bn = 1 (opts.py:52): BatchNorm1d over the packed valid regions first; the engine folds it into the Linear
        self.att_embed = nn.Sequential(*(((nn.BatchNorm1d(self.att_feat_size),) if self.use_bn else ()) +
                                         (nn.Linear(self.att_feat_size, self.rnn_size), nn.ReLU(), nn.Dropout(self.drop_prob_lm)) +
                                         ((nn.BatchNorm1d(self.rnn_size),) if self.use_bn == 2 else ())))
        if self.logit_layers == 1:
            self.logit = nn.Linear(self.rnn_size, self.vocab_size + 1)
        else:   # models/AttModel.py:89-91
            blocks = [m for _ in range(self.logit_layers - 1) for m in (nn.Linear(self.rnn_size, self.rnn_size), nn.ReLU(), nn.Dropout(0.5))]
            self.logit = nn.Sequential(*(blocks + [nn.Linear(self.rnn_size, self.vocab_size + 1)]))
        self.ctx2att = nn.Linear(self.rnn_size, self.att_hid_size)
        self.done_beams = []
        self._engine = None

    # the engine is created lazily (needs the CUDA device) and is not part of the state_dict
    @property
    def engine(self):
        if self._engine is None:
            self._engine = DecoderEngine(self)
        return self._engine

    def _bind_core(self):
        owner = self
        self.core._owner = lambda: owner  # plain closure: no module cycle in the registry

    def init_hidden(self, bsz):
        weight = next(self.parameters())
        return (weight.new_zeros(self.num_layers, bsz, self.rnn_size), weight.new_zeros(self.num_layers, bsz, self.rnn_size))

    def clip_att(self, att_feats, att_masks):
        if att_masks is not None:
            max_len = att_masks.long().sum(1).max()
            att_feats = att_feats[:, :max_len].contiguous()
            att_masks = att_masks[:, :max_len].contiguous()
        return att_feats, att_masks

    def _dropout(self, device):
        """(p, device seed) when the reference's nn.Dropout layers are active (training mode, drop_prob_lm > 0), else None.
        The seed is drawn from torch's CUDA generator per forward call; set `self.dropout_seed` to an int to pin it."""
        if not (self.training and self.drop_prob_lm > 0):
            return None
        seed = getattr(self, "_drop_seed_dev", None)
        if seed is None or seed.device != device:
            seed = self._drop_seed_dev = torch.zeros(1, dtype=torch.int64, device=device)
        if getattr(self, "dropout_seed", None) is None:
            seed.random_()
        else:
            seed.fill_(int(self.dropout_seed))
        return (float(self.drop_prob_lm), seed)

    def _prepare_feature(self, fc_feats, att_feats, att_masks):
        """Returns (fc, att, p_att, masks) like the reference.  att is the bf16 operand tile; p_att is ctx2att(att) in fp32,
        decoded from the engine's exponential operand tile (engine.prepare), so the single-step API below
        (get_logprobs_state / core / core.attention) takes exactly the reference's tensors."""
        f = self.engine.prepare(fc_feats, att_feats, att_masks)
        fc = fc_feats if self.kind == "att2in2" else f.fc        # (att2in2 / att2all2: fc_embed is the identity)
        return fc, f.att, _lib.tile_value(f.p_att), f.masks

    # ---- teacher-forced forward -----------------------------------------------------------------------
    def _scheduled_sampling(self, device):
        """(ss_prob, device seed) when scheduled sampling applies (training mode and ss_prob > 0, AttModel.py:130), else None.
        The seed is drawn from torch's CUDA generator per forward call; set `self.ss_seed` to an int to pin it."""
        if not (self.training and self.ss_prob > 0.0):
            return None
        seed = getattr(self, "_ss_seed_dev", None)
        if seed is None or seed.device != device:
            seed = self._ss_seed_dev = torch.zeros(1, dtype=torch.int64, device=device)
        if getattr(self, "ss_seed", None) is None:
            seed.random_()
        else:
            seed.fill_(int(self.ss_seed))
        return (float(self.ss_prob), seed)

    def _active_steps(self, seq):
        """Number of steps the reference executes before its all-zero-column break (AttModel.py:148-151)."""
        T = seq.size(1) - 1
        if T <= 1:
            return T
        empty = (seq[:, 1:T].sum(0) == 0).nonzero()
        return T if empty.numel() == 0 else int(empty[0]) + 1

    def _require_training_path(self):
        if getattr(self, "inference_only", False):
            raise NotImplementedError(f"{type(self).__name__} (use_bn={self.use_bn}, logit_layers={self.logit_layers}): the backward pass "
                                      "(and training-mode dropout / batch statistics) of this configuration is not built on the B200 "
                                      "path; call it under torch.no_grad() in eval() mode")

    def _check_tokens(self, seq):
        """Token ids outside [0, vocab_size] raise like the reference's nn.Embedding / gather would (IndexError) instead of
        being clamped by the kernels.  One tiny reduction + host read; skipped while a CUDA graph is being captured."""
        if seq.numel() == 0 or (seq.is_cuda and torch.cuda.is_current_stream_capturing()):
            return
        lo, hi = int(seq.min()), int(seq.max())
        if lo < 0 or hi > self.vocab_size:
            raise IndexError(f"token id out of range: [{lo}, {hi}] not in [0, {self.vocab_size}]")

    def _forward(self, fc_feats, attri_feats, att_feats, seq, att_masks=None):
        self._check_tokens(seq)
        if att_feats.size(0) == 0:   # empty batch: the reference's torch ops return an empty (0, T, V) tensor
            return att_feats.new_zeros((0, seq.size(1) - 1, self.vocab_size + 1), dtype=torch.float32)
        ss, drop = self._scheduled_sampling(att_feats.device), self._dropout(att_feats.device)
        if ss is not None or drop is not None or (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            self._require_training_path()
            from .autograd import decoder_logprobs
            return decoder_logprobs(self, fc_feats, att_feats, seq, att_masks, ss, drop)
        eng, lib = self.engine, _lib.load()
        feats = eng.prepare(fc_feats, att_feats, att_masks)
        n_steps = self._active_steps(seq)
        logits = eng.teacher_forced_logits(feats, seq, n_steps)
        B, T, V = logits.shape
        out = torch.zeros(B, T, V, device=logits.device)
        if n_steps == T:
            check(lib.uic_log_softmax_rows(ptr(logits), V, ptr(out), V, B * T, V, stream()))
        else:  # steps after the break stay zero (:123)
            for t in range(n_steps):
                check(lib.uic_log_softmax_rows(ptr(logits[:, t]), T * V, ptr(out[:, t]), T * V, B, V, stream()))
        return out

    def _forward_loss(self, fc_feats, attri_feats, att_feats, labels, masks, att_masks=None, global_mask_sum=None):
        """Additive fast path: fused teacher-forced forward + masked XE without the
        (B, T, V) log-prob tensor.  Equals crit(model(fc, attri, att, labels, att_masks), labels[:,1:], masks[:,1:]).
        `global_mask_sum` (data parallel): the loss normaliser summed over all ranks (dp.global_mask_sum)."""
        from .autograd import decoder_loss
        self._require_training_path()
        self._check_tokens(labels)
        return decoder_loss(self, fc_feats, att_feats, labels, masks, att_masks, global_mask_sum,
                            self._scheduled_sampling(att_feats.device), self._dropout(att_feats.device))

    # ---- single step API --------------------------------------------------------------------------------
    def _feats_from_api(self, fc, att, p_att, att_masks, rows):
        H, A = self.rnn_size, self.att_hid_size
        n_img = att.size(0)
        att3 = att.reshape(n_img, -1, H)
        L = att3.size(1)
        # the reference's beam path hands in per-beam expanded copies (AttModel.py:181-184): rows == n_img
        att_b = att3 if att3.dtype == BF16 else _lib.cast_bf16(att3.reshape(-1, H).float()).view(n_img, L, H)
        p3 = p_att.reshape(n_img, L, A)
        p_b = _lib.exp_tile(p3)                                        # ctx2att output -> operand tile E = exp(2 p_att)
        masks = None if att_masks is None else att_masks.reshape(n_img, L).float().contiguous()
        fc_b = None
        if self.kind != "att2in2":   # every other core consumes the embedded fc vector
            fc_b = fc if fc.dtype == BF16 else _lib.cast_bf16(fc.float().contiguous())
        if rows % n_img:
            raise ValueError(f"{rows} rows do not divide over {n_img} images")
        return Features(att_b.contiguous(), p_b.contiguous(), fc_b, masks, n_img, L), rows // n_img

    def _step_api(self, xt_or_it, is_token, fc, att, p_att, state, att_masks, want_logprobs):
        eng, lib = self.engine, _lib.load()
        w = eng.w
        R = xt_or_it.size(0)
        dev = att.device
        feats, beams = self._feats_from_api(fc, att, p_att, att_masks, R)
        H, E = w.H, w.E
        sl = Slots(self.kind, E, H)
        X = torch.zeros(R, w.Kx, dtype=BF16, device=dev)
        if feats.fc is not None:
            fc_rows = feats.fc if feats.fc.size(0) == R else feats.fc.repeat_interleave(beams, 0)
            X[:, sl.fc[0]:sl.fc[1]] = fc_rows
        if is_token:
            eng._embed(xt_or_it.contiguous().long(), X, sl)
        else:
            _lib.cast_bf16(xt_or_it.float().contiguous(), X[:, sl.xt[0]:sl.xt[1]])
        h0, c0 = state[0].float(), state[1].float().contiguous().clone()
        for layer, slot in enumerate(sl.h_load):                       # previous hidden states -> their bf16 operand slots
            _lib.cast_bf16(h0[layer].contiguous(), X[:, slot[0]:slot[1]])
        ws = eng._workspace(R, dev)
        h_out = eng.core_step(X, c0, feats, ws, beams=beams)
        h_state = torch.stack([X[:, slot[0]:slot[1]].float() for slot in sl.h_read])
        new_state = (h_state, c0)
        if not want_logprobs:
            return h_out.float(), new_state
        eng.logits_of(eng.logit_input(h_out), ws["logits"])
        out = torch.empty(R, w.V, device=dev)
        check(lib.uic_log_softmax_rows(ptr(ws["logits"]), w.V, ptr(out), w.V, R, w.V, stream()))
        return out, new_state

    @torch.no_grad()
    def get_logprobs_state(self, it, fc_feats, att_feats, p_att_feats, att_masks, state):
        """models/AttModel.py:158-165."""
        return self._step_api(it, True, fc_feats, att_feats, p_att_feats, state, att_masks, True)

    @torch.no_grad()
    def _core_api(self, xt, fc_feats, att_feats, p_att_feats, state, att_masks):
        return self._step_api(xt, False, fc_feats, att_feats, p_att_feats, state, att_masks, False)

    # ---- sampling ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def _sample_beam(self, fc_feats, att_feats, att_masks=None, opt={}):
        """models/AttModel.py:167-196; every image of the batch is searched at once on the device."""
        beam_size = opt.get("beam_size", 10)
        group_size = opt.get("group_size", 1)
        assert beam_size <= self.vocab_size + 1, "lets assume this for now, otherwise this corner case causes a few headaches down the road. can be dealt with in future if needed"
        eng = self.engine
        feats = eng.prepare(fc_feats, att_feats, att_masks, lazy=True)
        if group_size > 1:   # diverse beam search: tables carry a leading group axis (CaptionModel.py:100-177)
            tables = eng.beam_diverse(feats, self.seq_length, beam_size, group_size, opt.get("diversity_lambda", 0.5),
                                      opt.get("decoding_constraint", 0), opt.get("max_ppl", 0))
            done_seq_c, done_lp_c = tables[0].long().cpu(), tables[1].cpu()
            self._done_tables = (done_seq_c, done_lp_c, tables[2].cpu(), tables[3].cpu(), tables[4].cpu())
            self.done_beams = _LazyDoneBeams(self._done_tables)
            return done_seq_c[0, :, 0].contiguous(), done_lp_c[0, :, 0].contiguous()   # best of the first group (:194-195)
        done_seq, done_lp, done_p, done_unaug, done_cnt = eng.beam(
            feats, self.seq_length, beam_size, opt.get("decoding_constraint", 0), opt.get("max_ppl", 0))
        # one D2H for everything the reference keeps on the CPU (seq, seqLogprobs, done_beams)
        done_seq_c, done_lp_c = done_seq.long().cpu(), done_lp.cpu()
        done_p_c, done_unaug_c, done_cnt_c = done_p.cpu(), done_unaug.cpu(), done_cnt.cpu()
        self._done_tables = (done_seq_c, done_lp_c, done_p_c, done_unaug_c, done_cnt_c)
        self.done_beams = _LazyDoneBeams(self._done_tables)
        return done_seq_c[:, 0].contiguous(), done_lp_c[:, 0].contiguous()

    def _sample(self, fc_feats, attri_feats, att_feats, att_masks=None, opt={}):
        """models/AttModel.py:198-253.  sample_max = 0 draws from softmax(logits / temperature) on the device (Gumbel-max
        inside the logit GEMM's epilogue, seeded from torch's CUDA generator).  With gradients enabled the returned
        sample log-probs are differentiable (self-critical training, trainer.py:166-173)."""
        sample_max = opt.get("sample_max", 1)
        beam_size = opt.get("beam_size", 1)
        temperature = opt.get("temperature", 1.0)
        decoding_constraint = opt.get("decoding_constraint", 0)
        if att_feats.size(0) == 0:   # empty batch, like the reference: empty (0, seq_length) results
            return (att_feats.new_zeros((0, self.seq_length), dtype=torch.long), att_feats.new_zeros((0, self.seq_length), dtype=torch.float32))
        drop = self._dropout(att_feats.device)   # train() mode: the reference samples with its dropout layers active
        if drop is not None or (torch.is_grad_enabled() and not sample_max and any(p.requires_grad for p in self.parameters())):
            self._require_training_path() if getattr(self, "inference_only", False) else None
        with torch.no_grad():
            if beam_size > 1:
                if drop is not None:
                    raise NotImplementedError("beam search with active dropout (training mode, drop_prob_lm > 0) is not built: "
                                              "call model.eval()")
                return self._sample_beam(fc_feats, att_feats, att_masks, opt)
            eng = self.engine
            feats = eng.prepare(fc_feats, att_feats, att_masks, lazy=True, drop=drop)
            if sample_max:
                seq, lp = eng.greedy(feats, self.seq_length, decoding_constraint, drop=drop)
                return seq.clone(), lp.clone()
            if not temperature > 0.0:
                raise ValueError(f"sample_max=0 needs temperature > 0 (got {temperature})")
            seq, lp = eng.greedy(feats, self.seq_length, decoding_constraint, temperature=temperature, seed=opt.get("seed"), drop=drop)
            seq, lp = seq.clone(), lp.clone()
        if not (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            return seq, lp
        # differentiable log-probs of the sampled tokens; the steps the reference never reaches (after every row has
        # finished, AttModel.py:250-251) stay zero
        from . import autograd as AG
        B = seq.size(0)
        labels = torch.cat([seq.new_zeros(B, 1), seq, seq.new_zeros(B, 1)], 1)
        # (decoding_constraint only adds -inf to the banned token AFTER the log-softmax, AttModel.py:220-223: the log-prob
        # of the sampled token -- never the banned one -- is the plain teacher-forced one)
        lp_g = AG.decoder_token_logprobs(self, fc_feats, att_feats, labels, att_masks, drop)[:, :self.seq_length]
        all_fin = ((seq == 0).cumsum(1) > 0).all(0)                     # (T,) every row has emitted its end token by step t
        unwritten = (all_fin.cumsum(0) - all_fin.long()) > 0            # the loop broke before step t
        return seq, lp_g * (~unwritten).to(lp_g.dtype)


class _LazyDoneBeams:
    """`model.done_beams[k]` -> list of dicts {'seq','logps','unaug_p','p'} like the reference
    (models/CaptionModel.py:157-162), materialised from the device tables on first access."""

    def __init__(self, tables):
        self._t = tables

    def __len__(self):
        return self._t[0].size(-3)

    def __getitem__(self, k):
        seq, lp, p, unaug, cnt = self._t
        if seq.dim() == 4:   # diverse beam search: the groups' lists one after the other (CaptionModel.py:176)
            return [e for g in range(seq.size(0)) for e in _LazyDoneBeams((seq[g], lp[g], p[g], unaug[g], cnt[g]))[k]]
        return [{"seq": seq[k, j].clone(), "logps": lp[k, j].clone(), "unaug_p": float(unaug[k, j]), "p": float(p[k, j])}
                for j in range(int(cnt[k]))]

    def __iter__(self):
        return (self[k] for k in range(len(self)))


class Att2in2Model(AttModel):
    """models/AttModel.py:670-675."""
    kind = "att2in2"

    def __init__(self, opt):
        super().__init__(opt)
        self.core = Att2in2Core(opt)
        delattr(self, "fc_embed")
        self.fc_embed = lambda x: x
        self._bind_core()


class Att2all2Model(AttModel):
    """models/AttModel.py:678-683.  Same step plan as att2in2 (`kind`); `all_gates` selects the a2h accumulation."""
    kind = "att2in2"
    all_gates = True

    def __init__(self, opt):
        super().__init__(opt)
        self.core = Att2all2Core(opt)
        delattr(self, "fc_embed")
        self.fc_embed = lambda x: x
        self._bind_core()


class TopDownModel(AttModel):
    """models/AttModel.py:686-690."""
    kind = "topdown"

    def __init__(self, opt):
        super().__init__(opt)
        self.num_layers = 2
        self.core = TopDownCore(opt)
        self._bind_core()


class StackAttModel(AttModel):
    """models/AttModel.py:693-697.  Inference (teacher-forced log-probs, greedy / multinomial / beam sampling) runs on the
    B200 path; its backward pass is not built yet, so training calls raise."""
    kind = "stackatt"
    inference_only = True

    def __init__(self, opt):
        super().__init__(opt)
        self.num_layers = 3
        self.core = StackAttCore(opt)
        self._bind_core()


class DenseAttModel(AttModel):
    """models/AttModel.py:700-704 (the reference's best model, train.sh case 0).  Inference only, like StackAttModel."""
    kind = "denseatt"
    inference_only = True

    def __init__(self, opt):
        super().__init__(opt)
        self.num_layers = 3
        self.core = DenseAttCore(opt)
        self._bind_core()


def setup(opt):
    """models/__init__.py:22-59 for the models on the hot path."""
    if opt.caption_model == "att2in2":
        return Att2in2Model(opt)
    if opt.caption_model == "att2all2":
        return Att2all2Model(opt)
    if opt.caption_model == "topdown":
        return TopDownModel(opt)
    if opt.caption_model == "stackatt":
        return StackAttModel(opt)
    if opt.caption_model == "denseatt":
        return DenseAttModel(opt)
    raise Exception("Caption model not supported: {}".format(opt.caption_model))
