"""Training leg of bench.py: TopDown (configs[2]: 36x2048 region features, rnn 512, vocab 10k, seq 16)
teacher-forced forward + masked XE + BPTT + bucketed gradient all-reduce + grad-norm clip + Adam, the whole step in
one CUDA graph.  Weak scaling: 512 rows per GPU; strong scaling: a global batch of 512 rows split over the ranks."""
from __future__ import annotations

import torch

from . import dp, synth

CFG = "cfg3"
GLOBAL_ROWS_STRONG = 512


def make_state(local, rank, world, strong=False, cfg_name=CFG, rows=None):
    import unpaired_image_captioning_b200 as uic
    opt, cfg = synth.opt_for(cfg_name)
    model = uic.setup(opt)
    model.load_state_dict(synth.init_state_dict(opt, seed=1234))
    model = model.cuda().train()
    if rows is None:
        rows = cfg["batch"]
    if strong:   # one global batch (same seed on every rank), this rank's contiguous shard of it
        fc, att = synth.make_features(GLOBAL_ROWS_STRONG, cfg["att_size"], opt.att_feat_size, seed=4321)
        labels, masks = synth.make_captions(GLOBAL_ROWS_STRONG, opt.seq_length, opt.vocab_size, seed=4321)
        lo, hi = uic.shard_bounds(GLOBAL_ROWS_STRONG, rank, world)
        fc, att, labels, masks = fc[lo:hi], att[lo:hi], labels[lo:hi], masks[lo:hi]
    else:
        fc, att = synth.make_features(rows, cfg["att_size"], opt.att_feat_size, seed=4321 + rank)
        labels, masks = synth.make_captions(rows, opt.seq_length, opt.vocab_size, seed=4321 + rank)
    import os
    step = dp.DataParallelStep(model, clip=5.0, overlap=os.environ.get("UIC_DP_OVERLAP", "1") != "0")   # (0: A/B experiments)
    optim = torch.optim.Adam(model.parameters(), lr=4e-4, betas=(0.9, 0.999), eps=1e-8, fused=True, capturable=True)
    host = dict(fc=fc.pin_memory(), att=att.pin_memory(), labels=labels.pin_memory(), masks=masks.pin_memory())
    return dict(model=model, opt=opt, cfg=cfg, rows=fc.size(0), fc=fc.cuda(), att=att.cuda(), labels=labels.cuda(), masks=masks.cuda(),
                host=host, step=step, bucket=step.buckets, optim=optim)


def one_train_step(model=None, opt=None, cfg=None, fc=None, att=None, st=None, local=0, rank=0):
    st = st or make_state(local, rank, dp.world_size())
    loss = st["step"](st["fc"], st["att"], st["labels"], st["masks"])
    st["optim"].step()
    return loss


class GraphedTrainStep:
    """The whole training step (fused fwd + loss, BPTT, bucketed gradient all-reduce, normalise, clip, Adam) captured into
    ONE CUDA graph: ~330 kernel launches per step would otherwise be issued from Python."""

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


def grad_parity_check(local, rank, world, strong):
    """N > 1: the all-reduced, normalised gradient of the sharded step against rank 0's single-GPU gradient of the GATHERED
    batch (relative Frobenius error); the multi-GPU analogue of the parity tests."""
    st = make_state(local, rank, world, strong)
    st["step"].clip = 0.0
    st["step"](st["fc"], st["att"], st["labels"], st["masks"])
    got = st["bucket"].grads.clone()
    err = None
    if rank == 0:
        parts = [make_state(local, r, world, strong) for r in range(world)]       # every rank's data, regenerated from its seed
        ref = parts[0]
        cat = lambda k: torch.cat([p[k] for p in parts], 0)
        ref["step"].clip = 0.0
        world_saved = dp.world_size
        dp.world_size = lambda: 1                                                  # the reference run must not communicate
        try:
            ref["step"](cat("fc"), cat("att"), cat("labels"), cat("masks"))
        finally:
            dp.world_size = world_saved
        want = ref["bucket"].grads
        err = float((got - want).norm() / want.norm())
    return err


def train_leg(args, world, rank, local, strong=False, e2e=False, profile=False):
    from bench import _timed, _timed_wall
    from . import _lib
    parity = grad_parity_check(local, rank, world, strong) if world > 1 else None
    st = make_state(local, rank, world, strong)
    l0 = _lib.launch_count()
    one_train_step(st=st)
    launches = _lib.launch_count() - l0
    try:
        step = GraphedTrainStep(st)
        mode = "cuda_graph"
    except Exception as e:  # keep the measurement alive if capture is not possible on this stack
        torch.cuda.synchronize()
        step, mode = (lambda: one_train_step(st=st)), f"eager ({type(e).__name__}: {str(e)[:80]})"
    ms = _timed(step, args.steps, args.warmup, world)
    rows = st["rows"]
    opt = st["opt"]
    total_rows = GLOBAL_ROWS_STRONG if strong else world * rows
    out = {"metric": "train_samples_per_s", "value": total_rows / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
           "config": {"workload": "configs[2]: TopDown XE training (fwd + loss + bwd + bucketed all-reduce + clip + Adam)",
                      "caption_model": opt.caption_model, "rows_per_gpu": rows, "global_rows": total_rows,
                      "att_regions": st["cfg"]["att_size"], "rnn_size": opt.rnn_size, "vocab": opt.vocab_size + 1,
                      "seq_length": opt.seq_length},
           "scaling": "strong" if strong else "weak", "launch_mode": mode, "launches_per_step": launches,
           "exchange": {"buckets_mb": [round((hi - lo) * 4 / 2 ** 20, 1) for lo, hi in st["bucket"].bounds],
                        "order": "logit | core + fc_embed | embed | features + normaliser", "overlapped": st["step"].overlap,
                        "grad_parity_rel_err_vs_single_gpu": parity},
           "loss_rank0_share": float(step())}
    if e2e:   # the same step fed from pinned host memory: H2D of the batch and D2H of the loss inside the timing
        h = st["host"]

        def e2e_step():
            for k in ("fc", "att", "labels", "masks"):
                st[k].copy_(h[k], non_blocking=True)
            return float(step())

        ms_e = _timed_wall(e2e_step, args.steps, args.warmup, world)
        out["e2e"] = {"value": total_rows / (ms_e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e,
                      "h2d_bytes_per_step": sum(h[k].numel() * h[k].element_size() for k in ("fc", "att", "labels", "masks")),
                      "d2h_bytes_per_step": 4}
    if profile:   # live per-kernel timing of eager steps (rank 0)
        one_train_step(st=st)
        torch.cuda.synchronize()
        _lib.profile(True)
        for _ in range(2):
            one_train_step(st=st)
        prof = _lib.profile_dump()
        _lib.profile(False)
        out["profile"] = {k: (n / 2.0, msk / 2.0) for k, (n, msk) in prof.items()}
    return out
