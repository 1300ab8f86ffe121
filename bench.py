#!/usr/bin/env python
"""Benchmark of the decoder hot path (BASELINE.json metric: train samples/s + beam-3 captions/s; % roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg3|cfg5] [--strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input.  Rank 0 prints ONE JSON line.

  --workload cfg2 (default, BASELINE.json configs[1]): beam-3 sampling of a 256-image batch with the att2in2 decoder
      (rnn 512, 14x14x2048 features, vocab 10k, seq 16), prologue GEMMs included.
        value    : captions/s with the fp32 features already resident in HBM (device-timed, CUDA events)
        e2e      : captions/s through `model(fc, attri, att, masks, opt, mode='sample')` fed by the package's FeatureStream
                   from a pinned HOST FeatureCache (bf16, the operand precision): H2D of the features and D2H of the
                   sequences inside the timing; `e2e.fp32_host_features` is the same loop with the reference's fp32 arrays
        roofline : the fused attention-step kernel (HBM-bound), timed live with CUDA events per launch; `roofline.gemms`
                   lists the tcgen05 GEMMs of the step against the measured bf16 peak
        config.legs : greedy, the TopDown training step (configs[2], weak and -- N > 1 -- strong scaling), configs[4]
                   (beam-5, vocab 30k, rnn 1024) and the stock-PyTorch eager decoder on the same GPU
  --workload cfg3: the TopDown XE training step is the headline (`--strong`: global batch 512 split over the ranks)
  --workload cfg5: beam-5 sampling of 500-image chunks of configs[4]
  --workload cfg4: configs[3], the "unpaired" joint step: TopDown decoder step (batch 256 per GPU) + pivot-translator step
      (PyTorch arithmetic, the whole step replayed from one CUDA graph -- pivot.py); joint samples/s
  cpu_baseline / --impl reference : the oracle port of the reference's CPU path on this box's cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json; bf16 = sustained cuBLAS figure)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1500.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        # "median under load": the upper half of the samples (idle gaps between legs pull the clock down)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def _dist():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def _barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(ms, world):
    if world == 1:
        return ms
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def _timed(fn, steps, warmup, world):
    """W untimed + exactly K timed calls, barrier + synchronize on both sides, device time by CUDA
    events on the launching stream, max over ranks.  Returns ms per step."""
    for _ in range(warmup):
        fn()
    _barrier(world)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    _barrier(world)
    return _max_over_ranks(a.elapsed_time(b), world) / steps


def _timed_wall(fn, steps, warmup, world):
    """Same bracket with the host clock: for legs whose step ends with a device-to-host read (the result is on the host
    when the step returns), so the wall clock between two synchronisation points is the device + copy time."""
    for _ in range(warmup):
        fn()
    _barrier(world)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    _barrier(world)
    return _max_over_ranks(ms, world) / steps


def _bind_near_gpu(local):
    """N > 1: pin this rank's threads to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host cache is
    allocated (first touch).  Returns the node (or None when the topology is not exposed, e.g. a single-node VM)."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def _workload_config(name, cfg, opt, world, extra=None):
    label = {"cfg4": "configs[3]: joint step = TopDown decoder training step + pivot-translator (NMT) training step",
             "cfg2": "configs[1]: att2in2 decoder beam-%d sampling" % cfg["beam_size"],
             "cfg3": "configs[2]: TopDown XE training step (fwd + loss + bwd + gradient all-reduce + clip + Adam)",
             "cfg5": "configs[4]: att2in2 rnn 1024 / vocab 30k beam-%d sampling, one 500-image chunk per step" % cfg["beam_size"]}[name]
    c = {"workload": label, "caption_model": opt.caption_model, "rows_per_gpu": cfg["batch"], "beam_size": cfg["beam_size"],
         "att_regions": cfg["att_size"], "att_feat_size": opt.att_feat_size, "rnn_size": opt.rnn_size,
         "vocab": opt.vocab_size + 1, "seq_length": opt.seq_length, "parallelism": f"dp{world}",
         "l2": "inputs larger than L2 (%d MB of fp32 features per step, 126 MB L2); no flush needed"
               % (cfg["batch"] * cfg["att_size"] * opt.att_feat_size * 4 // 2 ** 20)}
    c.update(extra or {})
    return c


# ----------------------------------------------------------------------------------------------------
# CPU baselines (the oracle port of the reference's CPU path; bounded samples)
# ----------------------------------------------------------------------------------------------------
def cpu_beam_baseline(cfg, opt, sd, fc, att, repeats, threads, gpu_seq=None):
    """Oracle port of the reference's CPU beam search on a bounded sample of the same workload: `fc`, `att` are the first
    images of the very batch the GPU leg decoded, and `gpu_seq` its captions for them -- they must equal the oracle's
    (except at decision margins inside the north-star tolerance), so the timed GPU result is tied to the oracle."""
    from oracle import decoder_oracle as O
    from oracle.compare import compare_beam
    torch.set_num_threads(threads)
    best = float("inf")
    with torch.no_grad():
        for i in range(repeats + 1):
            t0 = time.perf_counter()
            O.sample_beam(sd, opt.caption_model, fc, att, opt.seq_length, cfg["beam_size"])
            dt = time.perf_counter() - t0
            if i > 0:                      # first pass is warm-up
                best = min(best, dt)
        check = None
        if gpu_seq is not None:            # (untimed pass that also records the decision margins)
            ref_seq, _, _, margins = O.sample_beam(sd, opt.caption_model, fc, att, opt.seq_length, cfg["beam_size"], return_margins=True)
            exact, exempt, failures = compare_beam(gpu_seq, ref_seq, margins, tol=1e-3)
            if failures:
                raise AssertionError(f"bench: the timed GPU captions differ from the oracle away from near-ties: {failures}")
            check = {"rows": int(ref_seq.size(0)), "exact": exact, "exempt_near_ties": exempt, "tol_rel": 1e-3}
    return fc.size(0) / best, check


def cpu_train_baseline(cfg_name, rows, repeats, threads):
    """Oracle port of the reference's CPU training computation (teacher-forced forward + masked XE + backward, fp32) on
    `rows` rows of the named workload; samples/s."""
    from oracle import decoder_oracle as O
    from unpaired_image_captioning_b200 import synth
    torch.set_num_threads(threads)
    opt, cfg = synth.opt_for(cfg_name)
    sd = synth.init_state_dict(opt, seed=1234)
    fc, att = synth.make_features(rows, cfg["att_size"], opt.att_feat_size, seed=4321)
    labels, masks = synth.make_captions(rows, opt.seq_length, opt.vocab_size, seed=4321)
    best = float("inf")
    for i in range(repeats + 1):
        t0 = time.perf_counter()
        O.loss_and_grads(sd, opt.caption_model, fc, att, labels, masks)
        dt = time.perf_counter() - t0
        if i > 0:
            best = min(best, dt)
    return {"value": rows / best, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{cfg_name} shapes, {rows} rows: oracle teacher-forced forward + XE + backward, best of {repeats} after warm-up, torch CPU fp32"}


def run_reference(args, world, rank):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port: the reference is Python and
    /root/reference does not exist on the GPU box), all host threads, same metric / unit as the b200 arm's workload."""
    if rank != 0:
        return
    from oracle import decoder_oracle as O
    from unpaired_image_captioning_b200 import synth
    opt, cfg = synth.opt_for(args.workload)
    sd = synth.init_state_dict(opt, seed=1234)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    if args.workload == "cfg4":
        _emit(_cfg4_cpu_line(args, world, threads, impl_reference=True))
        return
    if args.workload == "cfg3":
        sample = 16
        fc, att = synth.make_features(sample, cfg["att_size"], opt.att_feat_size, seed=4321)
        labels, masks = synth.make_captions(sample, opt.seq_length, opt.vocab_size, seed=4321)
        step = lambda: O.loss_and_grads(sd, opt.caption_model, fc, att, labels, masks)
        metric, unit, what = "train_samples_per_s", "samples/s", "teacher-forced forward + XE + backward"
    else:
        sample = 8 if args.workload == "cfg2" else 2
        fc, att = synth.make_features(sample, cfg["att_size"], opt.att_feat_size, seed=1234)

        def step():
            with torch.no_grad():
                O.sample_beam(sd, opt.caption_model, fc, att, opt.seq_length, cfg["beam_size"])
        metric, unit, what = "beam%d_captions_per_s" % cfg["beam_size"], "captions/s", "beam-%d sampling" % cfg["beam_size"]
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    note = f"each step = {what} over a bounded sample of {sample} rows of the workload"
    _emit({"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": _workload_config(args.workload, cfg, opt, world, {"sample": note}),
           "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port",
                            "sample": f"{sample} rows x {args.steps} steps, oracle/decoder_oracle.py ({what}), torch CPU fp32"},
           "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


# ----------------------------------------------------------------------------------------------------
# legs
# ----------------------------------------------------------------------------------------------------
def _decode_legs(args, world, rank, local, cfg_name, steps, with_roofline):
    """Beam (and greedy) decode of one workload: device-resident, end to end from the host cache, kernel profile."""
    import unpaired_image_captioning_b200 as uic
    from unpaired_image_captioning_b200 import _lib, synth
    opt, cfg = synth.opt_for(cfg_name)
    B, beam, T = cfg["batch"], cfg["beam_size"], opt.seq_length
    sd = synth.init_state_dict(opt, seed=1234)                     # identical weights on every rank
    model = uic.setup(opt)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    eng = model.engine
    fc_h, att_h = synth.make_features(B, cfg["att_size"], opt.att_feat_size, seed=1234 + rank)   # per-rank images
    fc_d, att_d = fc_h.cuda(), att_h.cuda()
    out = {"opt": opt, "cfg": cfg, "sd": sd, "fc_h": fc_h, "att_h": att_h}

    def beam_step():
        eng.beam(eng.prepare(fc_d, att_d, lazy=True), T, beam)

    def greedy_step():
        eng.greedy(eng.prepare(fc_d, att_d, lazy=True), T)

    out["ms_beam"] = _timed(beam_step, steps, args.warmup, world)
    l0 = eng.launches()
    beam_step()                                   # one more (untimed) step just to count its kernels
    out["launches"] = eng.launches() - l0
    # the captions of the timed plan (same graph, same inputs), kept for the oracle check beside cpu_baseline
    out["timed_seq"] = eng.beam(eng.prepare(fc_d, att_d, lazy=True), T, beam)[0][:, 0].long().cpu()
    out["ms_greedy"] = _timed(greedy_step, steps, args.warmup, world)

    # ---- end to end through the public API: host FeatureCache -> FeatureStream -> model(..., mode='sample') -> host ----
    sample_opt = {"beam_size": beam}
    d2h = B * beam * T * (8 + 4) + B * beam * (8 + 4) + B * 4      # done tables: seq (int64 after cast) + logps, p, unaug, cnt

    def e2e_leg(dtype):
        cache = uic.FeatureCache(fc_h, att_h, dtype=dtype)
        it = iter(uic.FeatureStream(cache, B, torch.device("cuda", local), loop=True))

        def step():
            fc, att, masks, _ = next(it)
            seq, lp = model(fc, None, att, masks, opt=sample_opt, mode="sample")   # returns CPU tensors (D2H inside)
            return seq

        ms = _timed_wall(step, steps, args.warmup, world)
        return {"value": world * B / (ms * 1e-3), "unit": "captions/s", "ms_per_step": ms, "h2d_bytes_per_step": cache.nbytes(B),
                "d2h_bytes_per_step": d2h, "host_gbs_per_rank": cache.nbytes(B) / (ms * 1e-3) / 1e9}

    out["e2e_bf16"] = e2e_leg(torch.bfloat16)
    out["e2e_fp32"] = e2e_leg(torch.float32)

    # ---- live per-kernel timing of one eager (non-graph) step ---------------------------------------------------------
    if with_roofline and rank == 0:
        eng.use_graphs = False
        beam_step()
        torch.cuda.synchronize()
        _lib.profile(True)
        for _ in range(3):
            beam_step()
        prof = _lib.profile_dump()
        _lib.profile(False)
        eng.use_graphs = True
        out["profile"] = {k: (n / 3.0, ms / 3.0) for k, (n, ms) in prof.items()}     # per step: launches, ms
    del fc_d, att_d
    model._engine = None
    return out


def _att_step_in_graph_us(n_img, beams, L, A, H, launches=32, replays=5):
    """Steady-state duration of the attention-step kernel at the workload's shape: `launches` back-to-back launches captured
    in one CUDA graph, the replays timed with CUDA events on the launching stream.  No host in the loop -- an event pair around
    an eagerly issued launch also counts the gap until the host gets the launch out (that reading moved between 42 and 77 us
    from run to run); this is the kernel's own duration, as it runs inside the captured decode loop."""
    from unpaired_image_captioning_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(0)
    R = n_img * beams
    e_tile = _lib.exp_tile((torch.randn(n_img, L, A, generator=g) * 0.5).cuda())
    att = torch.randn(n_img, L, H, generator=g).cuda().to(torch.bfloat16)
    f = (torch.exp(2.0 * torch.randn(R, A, generator=g).cuda()) * _lib.ATT_F_SCALE).contiguous()
    w = (torch.randn(A, generator=g) * 0.2).cuda()
    ctx = torch.empty(R, H, device="cuda", dtype=torch.bfloat16)
    run = lambda: _lib.att_step(f, A, e_tile, att, w, None, ctx, H, None, 0, None, n_img, beams, L, A, H)
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        for _ in range(3):
            run()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for _ in range(launches):
                run()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(replays):
            graph.replay()
        b.record()
        torch.cuda.synchronize()
    torch.cuda.current_stream().wait_stream(stream)
    return a.elapsed_time(b) * 1e3 / (launches * replays)


def _gemm_rooflines(profile, peak_tflops, top=4):
    """Tensor-pipe view of the GEMMs of the profiled step: algorithmic 2MNK flops / live event time vs the bf16 peak."""
    rows = []
    for name, (n, ms) in profile.items():
        if not name.startswith("gemm_") or name == "gemm_bf16_simt":
            continue
        m, nn, k = (int(v) for v in name[5:].split("x"))
        us = ms / n * 1e3
        tf = 2.0 * m * nn * k / (us * 1e-6) / 1e12
        rows.append({"kernel": name, "launches_per_step": n, "us_per_launch": us, "achieved": tf, "unit": "TFLOP/s", "peak": peak_tflops,
                     "frac": tf / peak_tflops, "ms_per_step": ms})
    rows.sort(key=lambda r: -r["ms_per_step"])
    return rows[:top]


def _eager_baseline():
    """Stock torch.nn decoder on this GPU (baseline/eager_decoder.py): cfg2 greedy and cfg3 training step, fp32 and bf16
    autocast.  Rank 0 only, a handful of steps."""
    from baseline.eager_decoder import EagerDecoder, xe_loss
    from unpaired_image_captioning_b200 import synth
    res = {"what": "the reference's architecture in stock torch.nn, eager (cuBLAS / ATen, one host sync per step) on this GPU",
           "torch": torch.__version__}
    opt, cfg = synth.opt_for("cfg2")
    m = EagerDecoder(opt)
    m.load_state_dict(synth.init_state_dict(opt, seed=1234))
    m = m.cuda().eval()
    fc, att = synth.make_features(cfg["batch"], cfg["att_size"], opt.att_feat_size, seed=1234)
    fc, att = fc.cuda(), att.cuda()
    for tag, ac in (("fp32", False), ("bf16_autocast", True)):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            ms = _timed(lambda: m.sample_greedy(fc, att), 5, 2, 1)
        res[f"cfg2_greedy_{tag}"] = {"value": cfg["batch"] / (ms * 1e-3), "unit": "captions/s", "ms_per_step": ms}
    del m, fc, att
    opt, cfg = synth.opt_for("cfg3")
    B = cfg["batch"]
    fc, att = synth.make_features(B, cfg["att_size"], opt.att_feat_size, seed=4321)
    labels, masks = synth.make_captions(B, opt.seq_length, opt.vocab_size, seed=4321)
    fc, att, labels, masks = fc.cuda(), att.cuda(), labels.cuda(), masks.cuda()
    for tag, ac in (("fp32", False), ("bf16_autocast", True)):
        m = EagerDecoder(opt)
        m.load_state_dict(synth.init_state_dict(opt, seed=1234))
        m = m.cuda().train()
        optim = torch.optim.Adam(m.parameters(), lr=4e-4, fused=True)

        def step():
            optim.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                out = m(fc, att, labels)
            loss = xe_loss(out.float(), labels[:, 1:], masks[:, 1:])
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
            optim.step()

        ms = _timed(step, 5, 2, 1)
        res[f"cfg3_train_{tag}"] = {"value": B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms}
        del m, optim
    return res


def _cfg4_cpu(rows, repeats, threads):
    """CPU joint step on a bounded sample: oracle decoder training computation (TopDown, `rows` rows) + the translator's
    forward + NLL + backward in PyTorch on the host cores (`rows` sentence pairs); samples/s."""
    from oracle import decoder_oracle as O
    from unpaired_image_captioning_b200 import pivot, synth
    torch.set_num_threads(threads)
    opt, cfg = synth.opt_for("cfg4")
    sd = synth.init_state_dict(opt, seed=1234)
    fc, att = synth.make_features(rows, cfg["att_size"], opt.att_feat_size, seed=4321)
    labels, masks = synth.make_captions(rows, opt.seq_length, opt.vocab_size, seed=4321)
    gen = torch.Generator().manual_seed(1234)
    torch.manual_seed(1234)
    nmt = pivot.PivotNMT().train()
    src, n = pivot.sentences(rows, 12000, gen)
    tgt, _ = pivot.sentences(rows, 8600, gen, bos=pivot.BOS)
    best = float("inf")
    for i in range(repeats + 1):
        t0 = time.perf_counter()
        O.loss_and_grads(sd, opt.caption_model, fc, att, labels, masks)
        nmt.zero_grad(set_to_none=True)
        nll, cnt = nmt(src, n, tgt)
        (nll / cnt).backward()
        dt = time.perf_counter() - t0
        if i > 0:
            best = min(best, dt)
    return rows / best


def _cfg4_cpu_line(args, world, threads, impl_reference=False):
    from unpaired_image_captioning_b200 import synth
    opt, cfg = synth.opt_for("cfg4")
    rows = 16
    v = _cfg4_cpu(rows, max(1, args.steps if impl_reference else 1), threads)
    base = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{rows} rows: oracle TopDown forward + XE + backward and the PyTorch translator forward + NLL + backward on the host cores, fp32"}
    if not impl_reference:
        return base
    return {"impl": "reference", "metric": "joint_train_samples_per_s", "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": rows / v * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": _workload_config("cfg4", cfg, opt, world, {"sample": base["sample"]}),
            "cpu_baseline": base, "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def _run_cfg4(args, world, rank, local, peaks, sampler, threads):
    """configs[3]: decoder step (this library's kernels, CUDA graph) + translator step (PyTorch arithmetic, CUDA graph)."""
    from unpaired_image_captioning_b200 import _lib, pivot, synth, train_bench
    st = train_bench.make_state(local, rank, world, cfg_name="cfg4")
    B = st["rows"]
    l0 = _lib.launch_count()
    train_bench.one_train_step(st=st)
    launches = _lib.launch_count() - l0
    dec = train_bench.GraphedTrainStep(st)
    torch.manual_seed(1234)                                    # identical translator weights on every rank
    gen = torch.Generator().manual_seed(1234 + rank)           # per-rank sentences
    nmt = pivot.PivotNMT().cuda().train()
    src, n = pivot.sentences(B, 12000, gen, max_len=30)
    tgt, _ = pivot.sentences(B, 8600, gen, bos=pivot.BOS, max_len=30)
    host = [t.pin_memory() for t in (src, n, tgt)]
    batch = [t.cuda() for t in host]
    eager = pivot.PivotTrainStep(nmt, src.size(0), tgt.size(0), B, graph=False, batch=batch, amp=False)
    ms_eager = _timed(eager.step, max(3, args.steps // 4), 2, world)
    graphed32 = pivot.PivotTrainStep(nmt, src.size(0), tgt.size(0), B, graph=True, batch=batch, amp=False)
    ms_graph32 = _timed(graphed32.step, max(3, args.steps // 4), 2, world)
    del graphed32
    graphed = pivot.PivotTrainStep(nmt, src.size(0), tgt.size(0), B, graph=True, batch=batch, amp=True)
    ms_dec = _timed(dec, args.steps, args.warmup, world)
    ms_nmt = _timed(graphed.step, args.steps, args.warmup, world)

    def joint():
        dec()
        graphed.step()

    ms_joint = _timed(joint, args.steps, args.warmup, world)
    hd = st["host"]

    def e2e_step():          # both halves fed from pinned host memory, both losses read back
        for k in ("fc", "att", "labels", "masks"):
            st[k].copy_(hd[k], non_blocking=True)
        graphed.load(*host)
        a = dec()
        b = graphed.step()
        return float(a) + float(b)

    ms_e2e = _timed_wall(e2e_step, args.steps, args.warmup, world)
    clocks = sampler.stop()
    # decoder half: tensor roofline of its GEMMs from a profiled eager step (every rank: the step contains the all-reduces)
    train_bench.one_train_step(st=st)
    torch.cuda.synchronize()
    _lib.profile(True)
    train_bench.one_train_step(st=st)
    prof = {k: (float(nn_), ms) for k, (nn_, ms) in _lib.profile_dump().items()}
    _lib.profile(False)
    if rank != 0:
        return None
    opt, cfg = st["opt"], st["cfg"]
    h2d = sum(hd[k].numel() * hd[k].element_size() for k in ("fc", "att", "labels", "masks")) + sum(t.numel() * 8 for t in host)
    gemms = _gemm_rooflines(prof, peaks["bf16_tflops"])
    g0 = gemms[0] if gemms else None
    cpu = None if (world > 1 or args.no_cpu_baseline) else _cfg4_cpu_line(args, world, threads)
    return {"metric": "joint_train_samples_per_s", "value": world * B / (ms_joint * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_joint, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 (decoder kernels) + fp32 (translator, PyTorch)", "data": "synthetic",
            "config": _workload_config("cfg4", cfg, opt, world, {
                "decoder_ms": ms_dec, "translator_graph_bf16_ms": ms_nmt, "translator_graph_fp32_ms": ms_graph32, "translator_eager_fp32_ms": ms_eager,
                "translator_share_of_joint_step": ms_nmt / (ms_nmt + ms_dec), "translator": "PyTorch restatement (pivot.py), whole step "
                "in one CUDA graph, bf16 autocast; parity unpinned (reference translator not importable, SURVEY F2/F3)",
                "sentence_length": "U{5..30}, padded to 30 / 32", "src_vocab": 12000, "tgt_vocab": 8600,
                "translator_params_M": round(sum(p.numel() for p in nmt.parameters()) / 1e6, 1)}),
            "clocks": clocks,
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8},
            "gpu_launches": int(launches * args.steps),
            "roofline": None if g0 is None else {"kernel": g0["kernel"] + " (largest GEMM of the decoder half)", "bound": "tensor", "achieved": g0["achieved"],
                                                 "peak": g0["peak"], "unit": "TFLOP/s", "frac": g0["frac"], "traffic": None,
                                                 "peak_source": peaks["source"], "gemms": gemms},
            "cpu_baseline": cpu}


# ----------------------------------------------------------------------------------------------------
def run_b200(args, world, rank, local):
    from unpaired_image_captioning_b200 import _lib, train_bench

    torch.cuda.set_device(local)
    if world > 1:
        pg_opts = None
        if os.environ.get("UIC_NCCL_HIGH_PRIORITY", "1") != "0":
            # the collectives' stream gets priority over the compute stream: the gradient buckets are exchanged while BPTT
            # still runs, and their CTAs must win the SM slots that compute kernels free (dp.DataParallelStep)
            pg_opts = torch.distributed.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=pg_opts)
    _lib.require_device()
    numa_node = _bind_near_gpu(local) if world > 1 else None
    peaks = _peaks()
    sampler = ClockSampler(local)
    sampler.start()
    threads = os.cpu_count() or 1
    wl = args.workload
    line = None

    if wl == "cfg4":
        line = _run_cfg4(args, world, rank, local, peaks, sampler, threads)
    elif wl in ("cfg2", "cfg5"):
        d = _decode_legs(args, world, rank, local, wl, args.steps, with_roofline=True)
        opt, cfg = d["opt"], d["cfg"]
        B, beam = cfg["batch"], cfg["beam_size"]
        legs = {"greedy": {"value": world * B / (d["ms_greedy"] * 1e-3), "unit": "captions/s", "ms_per_step": d["ms_greedy"]}}
        if wl == "cfg2" and not args.no_legs:
            legs["train_cfg3_weak"] = train_bench.train_leg(args, world, rank, local, strong=False)
            if world > 1:
                legs["train_cfg3_strong"] = train_bench.train_leg(args, world, rank, local, strong=True)
            d5 = _decode_legs(args, world, rank, local, "cfg5", max(3, args.steps // 4), with_roofline=False)
            legs["cfg5_beam5"] = {"value": world * d5["cfg"]["batch"] / (d5["ms_beam"] * 1e-3), "unit": "captions/s",
                                  "ms_per_step": d5["ms_beam"], "images_per_gpu_per_step": d5["cfg"]["batch"],
                                  "e2e_bf16_cache": d5["e2e_bf16"], "config": _workload_config("cfg5", d5["cfg"], d5["opt"], world)}
            del d5
            if rank == 0:
                legs["eager_b200_baseline"] = _eager_baseline()
        clocks = sampler.stop()
        if rank == 0:
            prof = d["profile"]
            total = sum(ms for _, ms in prof.values())
            shares = {k: {"launches": n, "ms_per_step": ms, "share": ms / total}
                      for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}
            n_att, ms_att = prof["att_step_fwd"]
            gemms = _gemm_rooflines(prof, peaks["bf16_tflops"])
            if wl == "cfg2":      # the attention step dominates: HBM roofline
                alg_bytes = B * cfg["att_size"] * (opt.att_hid_size + opt.rnn_size) * 2      # p_att + att tiles, bf16, once per image
                us_att = _att_step_in_graph_us(B, beam, cfg["att_size"], opt.att_hid_size, opt.rnn_size)
                achieved = alg_bytes / (us_att * 1e-6) / 1e9
                us_att1 = _att_step_in_graph_us(B, 1, cfg["att_size"], opt.att_hid_size, opt.rnn_size)   # the greedy leg's launch: one row per image
                traffic, traffic_src = _ncu_traffic()
                roofline = {"kernel": "att_step_fwd", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks["source"],
                            "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": us_att,
                            "us_per_launch_eager_event_pairs": ms_att / n_att * 1e3,
                            "beam1": {"us_per_launch": us_att1, "achieved": alg_bytes / (us_att1 * 1e-6) / 1e9,
                                      "frac": alg_bytes / (us_att1 * 1e-6) / 1e9 / peaks["hbm_gbs"],
                                      "what": "the same kernel with one row per image (greedy decode, first beam step): same bytes, a third of the arithmetic"},
                            "traffic_source": traffic_src, "gemms": gemms, "kernel_shares": shares,
                            "note": "duration: CUDA events around replays of a captured graph of 32 back-to-back launches of the kernel at "
                                    "this workload's shape (steady state, as inside the captured decode loop); us_per_launch_eager_event_pairs "
                                    "and kernel_shares: event pairs around every launch of an eagerly issued decode (include host launch "
                                    "gaps); traffic: dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture"}
            else:                  # vocab 30k x rnn 1024: the logit statistics GEMM dominates: tensor roofline
                g0 = gemms[0]
                roofline = {"kernel": g0["kernel"] + " (logit statistics GEMM)", "bound": "tensor", "achieved": g0["achieved"],
                            "peak": g0["peak"], "unit": "TFLOP/s", "frac": g0["frac"], "traffic": None, "peak_source": peaks["source"],
                            "us_per_launch": g0["us_per_launch"], "gemms": gemms, "kernel_shares": shares,
                            "note": "algorithmic 2 M N K flops / live CUDA-event time per launch of an eager decode in this run"}
            cpu = None
            if world == 1 and not args.no_cpu_baseline:
                sample = 8 if wl == "cfg2" else 2
                v, check = cpu_beam_baseline(cfg, opt, d["sd"], d["fc_h"][:sample].clone(), d["att_h"][:sample].clone(),
                                             repeats=2 if wl == "cfg2" else 1, threads=threads, gpu_seq=d["timed_seq"][:sample])
                cpu = {"value": v, "unit": "captions/s", "cores": threads, "kind": "port",
                       "sample": f"beam-{beam} over the first {sample} images of the GPU leg's own batch, best of 2 after warm-up, torch CPU fp32",
                       "parity_check": check}
                if wl == "cfg2" and not args.no_legs:
                    cpu["train"] = cpu_train_baseline("cfg1", 16, repeats=2, threads=threads)
            e2e = dict(d["e2e_bf16"])
            e2e.update({"host_feature_cache": "bf16 (FeatureCache default: the operand precision of the att_embed GEMM; captions identical "
                                              "to fp32 inputs, tests/test_gpu_loader.py)", "host_numa_node_rank0": numa_node,
                        "fp32_host_features": d["e2e_fp32"]})
            line = {"metric": "beam%d_captions_per_s" % beam, "value": world * B / (d["ms_beam"] * 1e-3), "unit": "captions/s",
                    "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": d["ms_beam"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": _workload_config(wl, cfg, opt, world, {"legs": legs}), "clocks": clocks, "e2e": e2e,
                    "gpu_launches": int(d["launches"] * args.steps), "roofline": roofline, "cpu_baseline": cpu}
    else:   # cfg3: the training step is the headline
        t = train_bench.train_leg(args, world, rank, local, strong=args.strong, e2e=True, profile=True)   # (every rank: the profiled eager steps contain the all-reduces)
        clocks = sampler.stop()
        if rank == 0:
            cpu = None
            if world == 1 and not args.no_cpu_baseline:
                cpu = cpu_train_baseline("cfg3", 16, repeats=2, threads=threads)
            prof = t.pop("profile")
            dims = lambda k: [int(v) for v in k[5:].split("x")]
            gemm_keys = [k for k in prof if k.startswith("gemm_") and k != "gemm_bf16_simt"]
            gemm_ms = sum(prof[k][1] for k in gemm_keys)
            gemm_flop = sum(2.0 * dims(k)[0] * dims(k)[1] * dims(k)[2] * prof[k][0] for k in gemm_keys)
            tf = gemm_flop / (gemm_ms * 1e-3) / 1e12
            total = sum(ms for _, ms in prof.values())
            shares = {k: {"launches": n, "ms_per_step": ms, "share": ms / total}
                      for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}
            roofline = {"kernel": "gemm_bf16_tcgen05_kernel (all GEMM launches of the step)", "bound": "tensor", "achieved": tf,
                        "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"], "traffic": None,
                        "peak_source": peaks["source"], "share_of_step": gemm_ms / total, "gemms": _gemm_rooflines(prof, peaks["bf16_tflops"]),
                        "kernel_shares": shares,
                        "note": "algorithmic 2 M N K flops of every GEMM launch / their live CUDA-event time in an eager step of this run"}
            line = {"metric": "train_samples_per_s", "value": t["value"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": t["ms_per_step"], "higher_is_better": True,
                    "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": dict(t["config"], parallelism=f"dp{world}", launch_mode=t["launch_mode"], exchange=t.get("exchange")),
                    "clocks": clocks, "e2e": t["e2e"], "gpu_launches": int(t["launches_per_step"] * args.steps), "roofline": roofline,
                    "cpu_baseline": cpu}
    if rank == 0:
        _emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


def _ncu_traffic():
    """DRAM bytes per launch of the attention kernel from the newest committed `ncu --set full` capture of this workload
    (profiles/*att_step_fwd_ncu_raw.csv, written by scripts/gpu_profiles.sh); (None, reason) when there is none."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*att_step_fwd_ncu_raw.csv")))
    files.sort(key=lambda f: (not os.path.basename(f).startswith("r2"), f))   # newest round first
    if not files:
        return None, "no ncu capture under profiles/"
    pick = files[0] if os.path.basename(files[0]).startswith("r2") else files[-1]
    try:
        rows = list(csv.reader(open(pick)))
        head, units, vals = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = head.index(key)
            total += float(vals[i].replace(",", "")) * scale[units[i]]
        return total, os.path.relpath(pick, ROOT)
    except (ValueError, KeyError, IndexError, OSError) as exc:
        return None, f"unreadable capture {os.path.basename(pick)}: {exc}"


_RESULT_OUT = None


def _emit(line):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--strong", action="store_true", help="cfg3: global batch 512 split over the ranks (strong scaling)")
    ap.add_argument("--no-legs", action="store_true", help="cfg2: skip the train / cfg5 / eager-baseline legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # stdout carries exactly ONE line, the JSON result: everything libraries print on file descriptor 1 (NCCL's
    # version banner, for one) goes to stderr instead
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    world, rank, local = _dist()
    if args.impl == "reference":
        run_reference(args, world, rank)
    else:
        run_b200(args, world, rank, local)


if __name__ == "__main__":
    main()
