"""Training leg of bench.py: TopDown (configs[2]: 36x2048 region features, rnn 512, vocab 10k, seq 16)
teacher-forced forward + masked XE + BPTT + gradient all-reduce + grad-norm clip + Adam."""
from __future__ import annotations

import torch

from . import dp, synth

CFG = "cfg3"
_state = {}


def _setup(local, rank, cfg_name=CFG):
    key = (local, cfg_name)
    if key in _state:
        return _state[key]
    import unpaired_image_captioning_b200 as uic
    opt, cfg = synth.opt_for(cfg_name)
    model = uic.setup(opt)
    model.load_state_dict(synth.init_state_dict(opt, seed=1234))
    model = model.cuda().train()
    B = cfg["batch"]
    fc, att = synth.make_features(B, cfg["att_size"], opt.att_feat_size, seed=4321 + rank)
    labels, masks = synth.make_captions(B, opt.seq_length, opt.vocab_size, seed=4321 + rank)
    bucket = dp.GradBucket(model)
    optim = torch.optim.Adam(model.parameters(), lr=4e-4, betas=(0.9, 0.999), eps=1e-8, fused=True, capturable=True)
    st = dict(model=model, opt=opt, cfg=cfg, fc=fc.cuda(), att=att.cuda(), labels=labels.cuda(), masks=masks.cuda(),
              bucket=bucket, optim=optim)
    _state[key] = st
    return st


def one_train_step(model=None, opt=None, cfg=None, fc=None, att=None, st=None, local=0, rank=0):
    st = st or _setup(local, rank)
    model, bucket = st["model"], st["bucket"]
    bucket.zero()
    norm = dp.global_mask_sum(st["masks"][:, 1:])
    loss = model(st["fc"], None, st["att"], st["labels"], st["masks"], None, mode="forward_loss", global_mask_sum=norm)
    loss.backward()
    bucket.allreduce()
    bucket.clip_(5.0)
    st["optim"].step()
    return loss


class GraphedTrainStep:
    """The whole training step (zero grads, global mask sum, fused fwd+loss, BPTT, gradient all-reduce, clip,
    Adam) captured into ONE CUDA graph: ~330 kernel launches per step would otherwise be issued from Python."""

    def __init__(self, st):
        self.st = st
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up on a side stream, as torch.cuda.graph requires
            for _ in range(3):
                one_train_step(st=st)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        from .engine import _no_gc
        with _no_gc():
            with torch.cuda.graph(self.graph):
                self.loss = one_train_step(st=st).detach()

    def __call__(self):
        self.graph.replay()
        # The replay updated the parameters in place without bumping their version counters, and the operand copies the
        # captured step packs for itself are one update behind afterwards: eval / sampling calls on this model must not
        # reuse them (or decode graphs built from them).
        self.st["model"].engine.invalidate()
        return self.loss


def train_samples_per_s(args, world, rank, local):
    from bench import _timed
    st = _setup(local, rank)
    try:
        step = GraphedTrainStep(st)
        mode = "cuda_graph"
    except Exception as e:  # keep the measurement alive if capture is not possible on this stack
        torch.cuda.synchronize()
        step, mode = (lambda: one_train_step(st=st)), f"eager ({type(e).__name__}: {str(e)[:80]})"
    ms = _timed(step, args.steps, args.warmup, world)
    B = st["cfg"]["batch"]
    opt = st["opt"]
    return {"metric": "train_samples_per_s", "value": world * B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
            "config": {"workload": "configs[2]: TopDown XE training (fwd + loss + bwd + allreduce + clip + Adam)",
                       "caption_model": opt.caption_model, "rows_per_gpu": B, "att_regions": st["cfg"]["att_size"],
                       "rnn_size": opt.rnn_size, "vocab": opt.vocab_size + 1, "seq_length": opt.seq_length},
            "scaling": "weak", "launch_mode": mode, "loss_rank0_share": float(step().detach() if mode != "cuda_graph" else step())}
