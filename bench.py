#!/usr/bin/env python
"""Benchmark of the decoder hot path (BASELINE.json metric: beam-3 captions/s + train samples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input: beam-3 sampling of a
256-image batch with the att2in2 decoder (BASELINE.json configs[1]: rnn 512, 14x14x2048 features,
vocab 10k, seq 16), prologue GEMMs included.  Rank 0 prints ONE JSON line.

  value    : captions/s with the fp32 features already resident in HBM (device-timed, CUDA events)
  e2e      : captions/s through the reference-facing API `model(fc, attri, att, masks, opt, mode='sample')`
             from pinned HOST buffers: H2D of the features and D2H of the sequences inside the timing
  roofline : the fused attention-step kernel (HBM-bound), timed live with CUDA events per launch
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

WORKLOAD = "cfg2"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


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


# ----------------------------------------------------------------------------------------------------
def cpu_baseline(cfg, opt, sd, fc, att, repeats, threads, gpu_seq=None):
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


def run_reference(args, world, rank):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port: the
    reference is Python and /root/reference does not exist on the GPU box), all host threads."""
    if rank != 0:
        return
    from oracle import decoder_oracle as O
    from unpaired_image_captioning_b200 import synth
    opt, cfg = synth.opt_for(WORKLOAD)
    sd = synth.init_state_dict(opt, seed=1234)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample = 8
    fc, att = synth.make_features(sample, cfg["att_size"], opt.att_feat_size, seed=99)

    def step():
        with torch.no_grad():
            O.sample_beam(sd, opt.caption_model, fc, att, opt.seq_length, cfg["beam_size"])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": "beam3_captions_per_s", "value": value, "unit": "captions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(cfg, opt, sample_note=f"each step = beam-{cfg['beam_size']} over a bounded sample of {sample} images"),
            "cpu_baseline": {"value": value, "unit": "captions/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} images x {args.steps} steps, oracle/decoder_oracle.py sample_beam, torch CPU fp32"},
            "e2e": {"value": value, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def _config(cfg, opt, sample_note=None):
    c = {"workload": "configs[1]: att2in2 decoder beam-%d sampling (greedy reported beside it)" % cfg["beam_size"],
         "caption_model": opt.caption_model, "images_per_gpu": cfg["batch"], "beam_size": cfg["beam_size"],
         "att_regions": cfg["att_size"], "att_feat_size": opt.att_feat_size, "rnn_size": opt.rnn_size,
         "vocab": opt.vocab_size + 1, "seq_length": opt.seq_length,
         "l2": "inputs larger than L2 (411 MB of fp32 features per step, 126 MB L2); no flush needed"}
    if sample_note:
        c["sample"] = sample_note
    return c


def _bind_near_gpu(local):
    """N > 1: pin this rank's threads to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned staging buffers are
    allocated (first touch), so that eight ranks do not stream their 413 MB per step through one socket's memory
    controllers and the inter-socket link.  Returns the node (or None when the topology is not exposed)."""
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


# ----------------------------------------------------------------------------------------------------
def run_b200(args, world, rank, local):
    import unpaired_image_captioning_b200 as uic
    from unpaired_image_captioning_b200 import _lib, synth

    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.require_device()
    numa_node = _bind_near_gpu(local) if world > 1 else None
    opt, cfg = synth.opt_for(WORKLOAD)
    B, beam, T = cfg["batch"], cfg["beam_size"], opt.seq_length
    sd = synth.init_state_dict(opt, seed=1234)                     # identical weights on every rank
    model = uic.setup(opt)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    eng = model.engine
    fc_h, att_h = synth.make_features(B, cfg["att_size"], opt.att_feat_size, seed=1234 + rank)   # per-rank images
    fc_h, att_h = fc_h.pin_memory(), att_h.pin_memory()
    fc_d, att_d = fc_h.cuda(), att_h.cuda()
    sampler = ClockSampler(local)
    sampler.start()

    # ---- device-resident legs ------------------------------------------------------------------
    def beam_step():
        feats = eng.prepare(fc_d, att_d, lazy=True)
        eng.beam(feats, T, beam)

    def greedy_step():
        feats = eng.prepare(fc_d, att_d, lazy=True)
        eng.greedy(feats, T)

    ms_beam = _timed(beam_step, args.steps, args.warmup, world)
    l0 = eng.launches()
    beam_step()                                   # one more (untimed) step just to count its kernels
    launches = eng.launches() - l0
    # the captions of the timed plan (same graph, same inputs), kept for the oracle check beside cpu_baseline
    timed_seq = eng.beam(eng.prepare(fc_d, att_d, lazy=True), T, beam)[0][:, 0].long().cpu()
    ms_greedy = _timed(greedy_step, args.steps, args.warmup, world)

    # ---- end-to-end leg through the public API, host buffers -------------------------------------
    sample_opt = {"beam_size": beam}
    # Input pipeline of the e2e leg: every step copies its own batch host->device (pinned memory) and reads
    # its sequences back; the copy of step i+1 is issued on a side stream while step i decodes (two device
    # buffer sets), as a loader would.  Each timed step still contains exactly one H2D and one D2H.
    copy_stream = torch.cuda.Stream()

    def make_e2e(att_host):
        dev_bufs = [(torch.empty_like(fc_d), torch.empty(att_host.shape, dtype=att_host.dtype, device=fc_d.device)) for _ in range(2)]
        pipe = {"i": 0, "ready": None}

        def _prefetch(slot):
            with torch.cuda.stream(copy_stream):
                dev_bufs[slot][0].copy_(fc_h, non_blocking=True)
                dev_bufs[slot][1].copy_(att_host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return ev

        def e2e_step():
            slot = pipe["i"] % 2
            ev = pipe["ready"] if pipe["ready"] is not None else _prefetch(slot)
            torch.cuda.current_stream().wait_event(ev)
            pipe["ready"] = _prefetch(1 - slot)      # next step's batch, overlapped with this step's decode
            pipe["i"] += 1
            fc, att = dev_bufs[slot]
            seq, lp = model(fc, None, att, None, opt=sample_opt, mode="sample")   # returns CPU tensors (D2H inside)
            return seq

        return e2e_step

    ms_e2e = _timed_wall(make_e2e(att_h), args.steps, args.warmup, world)
    h2d = fc_h.numel() * 4 + att_h.numel() * 4
    d2h = B * beam * T * (8 + 4) + B * beam * (8 + 4) + B * 4      # done tables: seq(int64 after cast)+logps, p, unaug, cnt
    # the same loop fed from a bf16 feature cache on the host: half the PCIe bytes, no staging cast.
    # Reported beside the contract's e2e number, which keeps the reference's fp32 inputs.
    att_h16 = att_h.to(torch.bfloat16).pin_memory()
    ms_e2e16 = _timed_wall(make_e2e(att_h16), args.steps, args.warmup, world)
    e2e_bf16 = {"value": world * B / (ms_e2e16 * 1e-3), "unit": "captions/s", "ms_per_step": ms_e2e16,
                "h2d_bytes_per_step": fc_h.numel() * 4 + att_h16.numel() * 2, "d2h_bytes_per_step": d2h,
                "note": "host features cached in bf16 (the precision the kernels consume); not the contract's e2e"}
    del att_h16

    # ---- training leg (teacher-forced fwd + XE + bwd + Adam), if the autograd path is present -------
    train = None
    if not args.no_train:
        try:
            from unpaired_image_captioning_b200.train_bench import train_samples_per_s
            train = train_samples_per_s(args, world, rank, local)
        except ImportError:
            train = None
    clocks = sampler.stop()

    # ---- live per-kernel timing of one eager (non-graph) step: roofline of the dominant kernel -----
    roofline, shares = None, None
    if rank == 0:
        eng.use_graphs = False
        beam_step()
        torch.cuda.synchronize()
        _lib.profile(True)
        for _ in range(3):
            beam_step()
        prof = _lib.profile_dump()
        _lib.profile(False)
        eng.use_graphs = True
        total = sum(ms for _, ms in prof.values())
        shares = {k: {"launches": n // 3, "ms_per_step": ms / 3, "share": ms / total}
                  for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}
        n_att, ms_att = prof["att_step_fwd"]
        peak, how = _peaks()
        alg_bytes = B * cfg["att_size"] * (opt.att_hid_size + opt.rnn_size) * 2      # p_att + att tiles, bf16, once per image
        achieved = alg_bytes / (ms_att / n_att * 1e-3) / 1e9
        traffic, traffic_src = _ncu_traffic()
        roofline = {"kernel": "att_step_fwd", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": how,
                    "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": ms_att / n_att * 1e3,
                    "traffic_source": traffic_src,
                    "note": "duration: live CUDA-event pairs around every launch of an eager decode in this run; traffic: "
                            "dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture of the same launch"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = 8
        v, check = cpu_baseline(cfg, opt, sd, fc_h[:sample].clone(), att_h[:sample].clone(), repeats=2, threads=threads,
                                gpu_seq=timed_seq[:sample])
        cpu = {"value": v, "unit": "captions/s", "cores": threads, "kind": "port",
               "sample": f"beam-{beam} over the first {sample} images of the GPU leg's own batch, best of 2 after warm-up, torch CPU fp32",
               "parity_check": check}

    if rank == 0:
        line = {"metric": "beam3_captions_per_s", "value": world * B / (ms_beam * 1e-3), "unit": "captions/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_beam,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": _config(cfg, opt), "clocks": clocks,
                "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "captions/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "host_numa_node_rank0": numa_node},
                "e2e_bf16_feature_cache": e2e_bf16,
                "gpu_launches": int(launches * args.steps),
                "greedy_captions_per_s": world * B / (ms_greedy * 1e-3), "greedy_ms_per_step": ms_greedy,
                "train": train, "roofline": roofline, "kernel_shares": shares, "cpu_baseline": cpu}
        _emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


def _ncu_traffic():
    """DRAM bytes per launch of the attention kernel from the newest committed `ncu --set full` capture of this workload
    (profiles/*att_step_fwd_ncu_raw.csv, written by scripts/gpu_round5.sh); (None, reason) when there is none."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*att_step_fwd_ncu_raw.csv")))
    if not files:
        return None, "no ncu capture under profiles/"
    try:
        rows = list(csv.reader(open(files[-1])))
        head, units, vals = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = head.index(key)
            total += float(vals[i].replace(",", "")) * scale[units[i]]
        return total, os.path.relpath(files[-1], ROOT)
    except (ValueError, KeyError, IndexError, OSError) as exc:
        return None, f"unreadable capture {os.path.basename(files[-1])}: {exc}"


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
    ap.add_argument("--no-train", action="store_true")
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
