#!/usr/bin/env python
"""bench.py -- headline benchmark of the NSC hot path on B200 (contract in the task statement, section 4).

Workload (BASELINE.json configs[1], "cq2"): collaborative-quantisation encode+decode of synthetic 16 kHz audio,
per frame: LPC analysis of the 1024-sample window -> 16 LSFs -> 256-bin LSF codebook -> lsf2poly -> sub-framed
LPC residual -> 2 cascaded bottleneck codecs ('9 9 100 20 1 2', stride 2, 32 bins, hard codes) -> sum -> LPC
synthesis.  Metric: seconds of audio coded per wall second (x real-time), 30 ms of new audio per frame (hop 480).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames B] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); frames shard by rank, no data-path collective ("weak").
"""
import argparse
import datetime
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

SEC_PER_FRAME = 480.0 / 16000.0            # hop-based: utilities.py:26
FLOP_PER_FRAME_CODEC = 2.0 * 150369280     # one bottleneck codec, stride [2]
METRIC = "seconds of 16 kHz audio coded per second (x real-time), CQ 2-codec encode+decode"


def synth_audio(n_frames, seed):
    """AR(2)-coloured unit-variance noise: frames (B,512) and their LPC windows (B,1024)."""
    from util import ar_frames
    win = ar_frames(n_frames, 1024, seed=seed)
    return np.ascontiguousarray(win[:, 256:768]), win


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), d.get('bf16_tflops_sustained', 1400.0), 'measured'
    return 6650.0, 1590.0, 1400.0, 'fallback'



ALG_BYTES_PER_FRAME = {"cq": 4624, "cq_with_lpc_window": 8720, "codec1": 4352}   # compulsory HBM bytes per frame


def ncu_facts():
    """Numbers that only a profiler gives, from the committed ncu summary of THIS round (profiles/r02_ncu_facts.json, written by
    tools/ncu_facts.py from `ncu --set full` / `--metrics dram__bytes...` captures of the same bench command): per kernel
    dram bytes per launch and sm__pipe_tensor_cycles_active, and the whole step's DRAM bytes per frame."""
    p = os.path.join(ROOT, 'profiles', 'r02_ncu_facts.json')
    return json.load(open(p)) if os.path.exists(p) else {}


def roofline_record(agg, tot_ms, args, frames):
    """the conv kernels are dense contractions (1.3e5 FLOP per compulsory byte), so the roof that binds them is the
    TENSOR PIPE: frac = algorithmic FLOP/s (2 x real-channel MACs, no padding, ONE product per MAC) / measured sustained bf16
    rate.  The fp32-class mode issues 3 MMAs per product (fp16 hi/lo split), so `issued_frac` = 3 x frac is what the pipe
    executes; ncu's own sm__pipe_tensor_cycles_active sits beside it.  HBM stays in the record as `traffic` (measured DRAM bytes
    of one launch) against the kernel's algorithmic bytes, and for the whole step against SURVEY's compulsory bytes per frame."""
    hbm, bf16, bf16_sus, how = peaks()
    is_conv = lambda k: (k.startswith(('conv_', 'pT', 'pX', 'pG', 'pB')) or (k.startswith('tc') and k != 'tc_pack_weights'))
    conv = [(k, v) for k, v in agg.items() if is_conv(k)]
    top_name, top = max(conv, key=lambda kv: kv[1][0])
    tensor_path = not top_name.startswith('conv_')
    mma_per_product = 3 if (args.precision == 'tc_f16x3' and tensor_path) else 1
    secs = top[0] * 1e-3
    ach_tf = top[1] / secs / 1e12
    ach_gbs = top[2] / secs / 1e9
    conv_ms = sum(v[0] for _, v in conv)
    conv_fl = sum(v[1] for _, v in conv)
    all_fl = sum(v[1] for v in agg.values())
    facts = ncu_facts()
    kf = facts.get('kernels', {}).get(top_name, {})
    alg_bytes_launch = top[2] / top[3]
    # ncu measured DRAM bytes of ONE launch of this layer at the pass size of ITS run; per launch of THIS run = the same bytes per
    # algorithmic byte (the launch list's pass size may differ from the live one: NSC_PLANE_CHUNK, --frames)
    traffic = kf.get('dram_bytes_per_launch')
    if traffic and kf.get('algorithmic_bytes_per_launch'):
        traffic = traffic * alg_bytes_launch / kf['algorithmic_bytes_per_launch']
    step = facts.get('whole_step', {})
    dram_pf = step.get('dram_bytes_per_frame')
    peak = bf16_sus if tensor_path else None
    rec = {
        "bound": "tensor", "achieved": ach_tf, "peak": bf16_sus, "unit": "TFLOP/s", "frac": ach_tf / bf16_sus,
        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({how}); the kernel is timed inside a long step",
        "kernel": top_name, "launch_ms": top[0] / top[3], "launches_per_step": top[3], "share_of_step": top[0] / tot_ms,
        "pipe": ("fp32 FFMA (CUDA cores)" if not tensor_path else "tcgen05 kind::f16, fp32 accumulate in TMEM"),
        "mma_per_product": mma_per_product,
        "issued_frac": ach_tf * mma_per_product / bf16_sus,
        "tensor_pipe_active_ncu": kf.get('sm__pipe_tensor_cycles_active_pct'),
        "traffic": traffic,
        "traffic_detail": ({"ncu_dram_bytes_of_the_captured_launch": kf.get('dram_bytes_per_launch'),
                            "algorithmic_bytes_of_that_launch": kf.get('algorithmic_bytes_per_launch'),
                            "traffic_vs_algorithmic": (kf['dram_bytes_per_launch'] / kf['algorithmic_bytes_per_launch']) if kf.get('dram_bytes_per_launch') and kf.get('algorithmic_bytes_per_launch') else None,
                            "frames_of_that_launch": kf.get('frames_per_launch'), "source": kf.get('source')} if kf else None),
        "hbm": {"achieved_gbs_algorithmic": ach_gbs, "frac_of_measured": ach_gbs / hbm, "peak_gbs": hbm,
                "algorithmic_bytes_per_launch": alg_bytes_launch,
                "note": "real channels only (2 B per channel and fp16 plane), input + output (+ residual) moved once"},
        "all_conv": {"tflops_algorithmic": conv_fl / (conv_ms * 1e-3) / 1e12, "frac": conv_fl / (conv_ms * 1e-3) / 1e12 / bf16_sus,
                     "issued_frac": conv_fl * mma_per_product / (conv_ms * 1e-3) / 1e12 / bf16_sus, "share_of_step": conv_ms / tot_ms},
        "whole_step": {"tflops_algorithmic": all_fl / (tot_ms * 1e-3) / 1e12, "frac": all_fl / (tot_ms * 1e-3) / 1e12 / bf16_sus,
                       "dram_bytes_per_frame": dram_pf, "algorithmic_bytes_per_frame": ALG_BYTES_PER_FRAME["cq"] if args.codecs == 2 else None,
                       "traffic_vs_algorithmic": (dram_pf / ALG_BYTES_PER_FRAME["cq"]) if (dram_pf and args.codecs == 2) else None,
                       "source": step.get('source')},
    }
    return rec


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(',')]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_throughput(n_frames, chunk=128, threads=None):
    """The restated reference (oracle/) on the host cores: same workload, same seeded weights, hard path."""
    import torch
    from oracle import ref_codec, ref_lpc
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = ref_codec.OracleCodecCfg()
    codecs = [ref_codec.OracleCodec(cfg, seed=5), ref_codec.OracleCodec(cfg, seed=6)]
    bins = np.load(os.path.join(ROOT, 'tests', 'golden', 'lsf_bins_f64.npy')).astype(np.float32)
    x, win = synth_audio(n_frames, seed=1234)
    t0 = time.perf_counter()
    with torch.no_grad():
        for b0 in range(0, n_frames, chunk):
            xs, ws = x[b0:b0 + chunk], win[b0:b0 + chunk]
            lsf = ref_lpc.lpc_analysis_windows(ws, 16).astype(np.float32)
            ref_codec.cq_feedforward(codecs, -300.0, bins, torch.from_numpy(xs)[:, :, None],
                                     torch.from_numpy(lsf)[:, :, None], False, 1.0)
    dt = time.perf_counter() - t0
    return n_frames * SEC_PER_FRAME / dt, dt, threads


def workload_name(codecs, bins):
    """config.workload of both arms (ours and --impl reference)"""
    return (f"cq{codecs}: LPC analysis + 256-bin LSF codebook + {codecs} cascaded bottleneck codecs "
            f"('9 9 100 20 1 2', stride 2, {bins} bins, hard codes) + LPC synthesis")


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = args.ref_frames
    cpu_reference_throughput(min(n, 128))   # warm-up (thread pools, filter design caches)
    vals = []
    for _ in range(args.steps):
        v, dt, thr = cpu_reference_throughput(n)
        vals.append((v, dt))
    v = float(np.median([a for a, _ in vals]))
    ms = float(np.median([b for _, b in vals])) * 1e3
    sample = f"{n} frames per step in chunks of 128, restated reference (numpy/scipy/torch-CPU oracle), hard path"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "x real-time", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (LPC parts f64)", "data": "synthetic",
        "config": {"workload": workload_name(2, 32), "frames_per_step": n,
                   "sample_of": "the headline arm's workload, bounded to what the host cores finish in seconds per step"},
        "cpu_baseline": {"value": v, "unit": "x real-time", "cores": thr, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "x real-time", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "TensorFlow/audiolazy/spectrum are not installable here; this is the CPU restatement, not TensorFlow",
    }))


# ------------------------------------------------------------------------------------------------ our arm
def kernel_breakdown(lib, fn, cap=8192):
    """Runs fn() once with per-launch CUDA-event recording on (nsc_profile_begin/end) and aggregates by kernel name:
    name -> [ms, flops, bytes, launches]."""
    lib.nsc_profile_begin(cap)
    fn()
    n = C.c_int32(0)
    names = C.create_string_buffer(cap * 32)
    ms = (C.c_float * cap)(); fl = (C.c_double * cap)(); by = (C.c_double * cap)()
    lib.nsc_profile_end(C.byref(n), names, ms, fl, by, cap)
    agg = {}
    for i in range(n.value):
        nm = names.raw[i * 32:(i + 1) * 32].split(b'\0')[0].decode()
        a = agg.setdefault(nm, [0.0, 0.0, 0.0, 0])
        a[0] += ms[i]; a[1] += fl[i]; a[2] += by[i]; a[3] += 1
    return agg


def breakdown_table(agg):
    tot = sum(a[0] for a in agg.values()) or 1.0
    return {k: {"ms": round(v[0], 3), "share": round(v[0] / tot, 4), "launches": v[3],
                "tflops": round(v[1] / (v[0] * 1e-3) / 1e12, 2) if v[0] > 0 else None,
                "gbs": round(v[2] / (v[0] * 1e-3) / 1e9, 1) if v[0] > 0 else None}
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from nsc_b200 import _lib, codec, lpc_utilities as lu
    from nsc_b200.sharding import max_over_ranks

    world, rank, local, dev = init_dist()
    lib = _lib.load()

    B = args.frames                      # frames per GPU per step (weak scaling)
    cfg = codec.CodecConfig(resnet_type='bottleneck', precision=args.precision, num_bins=args.bins)
    gcs = [codec.NeuralCodec(cfg, device=dev, seed=5 + i) for i in range(args.codecs)]
    cm = codec.CMRL(gcs, res_scalar=1.0)
    x_np, win_np = synth_audio(min(B, 4096), seed=1234 + rank)
    reps = -(-B // x_np.shape[0])
    x_host = torch.from_numpy(np.tile(x_np, (reps, 1))[:B]).pin_memory()
    win_host = torch.from_numpy(np.tile(win_np, (reps, 1))[:B]).pin_memory()
    x_dev, win_dev = x_host.to(dev), win_host.to(dev)

    def step_device():
        lsf = lu.lpc_analysis_windows(win_dev, 16, dtype=torch.float32)
        return cm.feedforward_lpc(x_dev, lsf, False, 1.0)

    # once per WEIGHTS, outside the timed region, like the reference's variable initialisation: packed operand slabs of every layer
    # and the zero rows of the activation images (nsc_prepare).  Every call with a batch of at least one engine pass then skips the
    # 3.6 GB border memset and the ~75 packing launches -- 1.7 ms per call, which the end-to-end path paid eight times per step.
    # The same step without it is reported as `unprepared`.
    cm.prepare(B)

    out_host = {}
    # End-to-end step through the public API with HOST buffers: the batch goes through in sub-batches so that the pinned-memory
    # H2D copy of sub-batch k+1 and the D2H copy of sub-batch k-1 (copy stream) run under the compute of sub-batch k.
    # Sub-batches are at least one pass of the engine (nsc_pass_frames) so the prepared workspace serves every call; the exposed part of the copies is the
    # first sub-batch's H2D and the last one's D2H.  H2D and D2H use separate streams (both copy engines).
    sub = cm.pass_frames()
    n_sub = max(1, min(8, B // sub))               # (the last sub-batch takes the remainder: every call is at least one pass)
    bounds = [(i * sub, (i + 1) * sub if i + 1 < n_sub else B) for i in range(n_sub)]
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    dev_in = [None] * n_sub

    def step_e2e():
        main = torch.cuda.current_stream()
        copy_stream.wait_stream(main)
        d2h_stream.wait_stream(main)
        h2d_done = []
        for i, (lo, hi) in enumerate(bounds):
            with torch.cuda.stream(copy_stream):
                dev_in[i] = (x_host[lo:hi].to(dev, non_blocking=True), win_host[lo:hi].to(dev, non_blocking=True))
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                h2d_done.append(ev)
        r = None
        for i, (lo, hi) in enumerate(bounds):
            main.wait_event(h2d_done[i])
            xd, wd = dev_in[i]
            lsf = lu.lpc_analysis_windows(wd, 16, dtype=torch.float32)
            r = cm.feedforward_lpc(xd, lsf, False, 1.0)
            done = torch.cuda.Event()
            done.record(main)
            outs = [('lsf_idx', r['lsf_idx']), ('syn', r['synthesized'])] + [('idx%d' % k, t) for k, t in enumerate(r['idx'])]
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                for k, t in outs:
                    if k not in out_host:
                        out_host[k] = torch.empty((B,) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory()
                    out_host[k][lo:hi].copy_(t, non_blocking=True)
                    t.record_stream(d2h_stream)
            xd.record_stream(main)
            wd.record_stream(main)
        main.wait_stream(copy_stream)
        main.wait_stream(d2h_stream)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)

    for _ in range(args.warmup):
        step_device()
    if args.profile_one_step:
        # for `ncu --profile-from-start off`: exactly ONE device-resident step inside the profiler range, with the launch names in
        # order (nsc_profile records) written beside it so that tools/ncu_facts.py can attribute ncu's rows to layers
        torch.cuda.synchronize()
        agg_names = []
        lib.nsc_profile_begin(8192)
        torch.cuda.cudart().cudaProfilerStart()
        step_device()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        n = C.c_int32(0)
        names = C.create_string_buffer(8192 * 32)
        ms = (C.c_float * 8192)(); fl = (C.c_double * 8192)(); by = (C.c_double * 8192)()
        lib.nsc_profile_end(C.byref(n), names, ms, fl, by, 8192)
        recs = [{"name": names.raw[i * 32:(i + 1) * 32].split(b'\0')[0].decode(), "ms": ms[i], "flops": fl[i], "bytes": by[i]} for i in range(n.value)]
        json.dump({"frames": B, "records": recs}, open(args.profile_one_step, 'w'))
        return
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = lib.nsc_launch_count()
    t_dev = timed(step_device, args.steps)
    launches = lib.nsc_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    t_e2e = timed(step_e2e, args.steps)
    cm.release()                                    # the same device-resident step when every call packs the weights and clears the borders
    step_device()
    t_cold = timed(step_device, 2)
    cm.prepare(B)

    # ---- per-kernel breakdown of ONE step with CUDA events on the launching stream (roofline line)
    roof = None
    breakdown = None
    if rank == 0:
        agg = kernel_breakdown(lib, step_device)
        tot = sum(a[0] for a in agg.values())
        breakdown = breakdown_table(agg)
        roof = roofline_record(agg, tot, args, B)

    # ---- the other BASELINE.json configurations as bounded sub-records of the same run (every rank takes part: the training
    # step contains the gradient all-reduce, the corpus and the sweeps shard by rank)
    subrec = None
    if args.sub_records:
        del x_dev, win_dev
        torch.cuda.empty_cache()
        subrec = {}
        for name, fn in (("codec1_b128", lambda: measure_codec1(args, world, rank, dev, lib, cpu=not args.no_cpu_baseline)),
                         ("cq_scaled", lambda: measure_cq_sweep(args, world, rank, dev, lib)),
                         ("cq2_gln", lambda: measure_variant(args, world, rank, dev, lib, 'gln', (2,))),
                         ("cq2_stride4", lambda: measure_variant(args, world, rank, dev, lib, 'bottleneck', (2, 2))),
                         ("cq2_gln_stride4", lambda: measure_variant(args, world, rank, dev, lib, 'gln', (2, 2))),
                         ("train", lambda: measure_train(args, world, rank, dev, lib, 20, 5, breakdown=True)),
                         ("train_gln", lambda: measure_train(args, world, rank, dev, lib, 10, 3, breakdown=False, resnet_type='gln')),
                         ("corpus_1h_per_gpu", lambda: measure_corpus(args, world, rank, dev, lib, 2, 1, n_utt=360))):
            try:
                subrec[name] = fn()
            except Exception as e:      # a sub-record must never cost the headline line
                subrec[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                if world > 1:
                    raise
        if world > 1 and isinstance(subrec.get("train"), dict):
            subrec["train"].pop("kernel_breakdown", None)

    if rank == 0:
        frames_total = B * world
        value = frames_total * args.steps * SEC_PER_FRAME / t_dev
        e2e_v = frames_total * args.steps * SEC_PER_FRAME / t_e2e
        h2d = int(x_host.numel() * 4 + win_host.numel() * 4)
        d2h = int(sum(t.numel() * t.element_size() for t in out_host.values()))
        cpu_v, cpu_dt, cpu_thr = (None, None, None)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_v, cpu_dt, cpu_thr = cpu_reference_throughput(args.cpu_frames)
            cpu = {"value": cpu_v, "unit": "x real-time", "cores": cpu_thr, "kind": "port",
                   "sample": f"{args.cpu_frames} frames of the same workload in chunks of 128 ({cpu_dt:.1f} s), restated "
                             "reference (numpy/scipy/torch-CPU oracle); TensorFlow itself is not installable here"}
        line = {
            "metric": METRIC, "value": value, "unit": "x real-time", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": {"fp32": "f32", "tc_f16x3": "f32-equivalent (fp16 hi/lo split on tensor cores, fp32 accumulate)",
                      "tc_f16": "f16 inputs / f32 accumulate (REDUCED precision)"}[args.precision] +
                     "; LPC analysis/residual/synthesis f64", "data": "synthetic",
            "config": {"workload": workload_name(args.codecs, args.bins),
                       "frames_per_gpu_per_step": B, "frames_per_step": frames_total, "conv_precision": args.precision,
                       "l2_policy": f"inputs larger than L2: {h2d / 1e6:.0f} MB of frames+windows per GPU per step, plus a "
                                    "multi-GB activation workspace cycled per ~2k-frame chunk (L2 is 126 MB); no explicit flush",
                       "parallelism": f"dp{world} (frames sharded by rank, no collective)",
                       "weights": "prepared once per weights (nsc_prepare: packed fp16 hi/lo operand slabs + zero rows of the activation "
                                  "images), outside the timed region, as the reference keeps its variables resident"},
            "unprepared": {"ms_per_step": t_cold / 2 * 1e3, "value": frames_total * 2 * SEC_PER_FRAME / t_cold,
                           "note": "the same device-resident step when every call packs the weights and clears the image borders"},
            "frames_per_s": frames_total * args.steps / t_dev,
            "e2e": {"value": e2e_v, "unit": "x real-time", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": t_e2e / args.steps * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "kernel_breakdown": breakdown,
            "sub_records": subrec,
            "compute_tflops_whole_step": frames_total * args.steps * args.codecs * FLOP_PER_FRAME_CODEC / t_dev / 1e12,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def init_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1 and not dist.is_initialized():
        # the contract is ONE JSON line on stdout: keep NCCL's banner ("NCCL version ...") off it
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    return world, rank, local, dev


def _barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _timed(fn, steps, world, dev):
    """steps calls of fn between a barrier + synchronize on both sides, CUDA events, max over ranks -> seconds"""
    import torch
    from nsc_b200.sharding import max_over_ranks
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    _barrier(world)
    return max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)


def measure_corpus(args, world, rank, dev, lib, steps, warmup, n_utt=None):
    """BASELINE.json configs[4]: an hour-scale synthetic corpus, wav -> hard codes (packed records) -> wav, through
    nsc_b200.pipeline (cmrl.py:666-737 batched): host signals in, packed records + synthesized signals out, every step
    (normalisation, filters, framing, LPC analysis, CQ pass, overlap-add, de-emphasis, bit packing, H2D/D2H) inside the timed
    region.  Utterances are sharded across ranks (weak scaling: n_utt per GPU)."""
    import torch
    from nsc_b200 import codec, pipeline
    cfg = codec.CodecConfig(resnet_type='bottleneck', precision=args.precision)
    cm = codec.CMRL([codec.NeuralCodec(cfg, device=dev, seed=5), codec.NeuralCodec(cfg, device=dev, seed=6)], res_scalar=1.0)
    n_utt = n_utt or args.utterances
    T = int(args.utt_seconds * 16000)
    cm.prepare(max(cm.pass_frames(), n_utt))        # once per weights (nsc_prepare): any call of at least one engine pass uses it
    x_np, _ = synth_audio(64, seed=4321 + rank)
    base = np.tile(x_np.reshape(-1), -(-T * 8 // x_np.size))        # a few distinct utterances, tiled
    host = [torch.from_numpy(np.ascontiguousarray(base[(i % 8) * 4000:(i % 8) * 4000 + T])).pin_memory() for i in range(n_utt)]
    out_host = {}
    frames = [0]

    def step():
        sigs = [h.to(dev, non_blocking=True) for h in host]
        res = pipeline.code_utterances(cm, sigs, the_share=False, pack=True)
        rec = torch.cat([r['records'] for r in res])
        syn = torch.cat([r['synthesized'] for r in res])
        for k, t in (('rec', rec), ('syn', syn)):
            if k not in out_host:
                out_host[k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            out_host[k].copy_(t, non_blocking=True)
        frames[0] = sum(r['n_frames'] for r in res)

    for _ in range(warmup):
        step()
    l0 = lib.nsc_launch_count()
    t = _timed(step, steps, world, dev)
    launches = lib.nsc_launch_count() - l0
    audio_s = n_utt * world * args.utt_seconds
    v = audio_s * steps / t
    return {
        "metric": "seconds of 16 kHz audio coded per second (x real-time), corpus wav -> packed hard codes -> wav, end to end",
        "value": v, "unit": "x real-time", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": t / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32-equivalent convs (fp16 hi/lo on tensor cores); filters / LPC f64", "data": "synthetic",
        "config": {"workload": f"corpus: {n_utt} utterances x {args.utt_seconds:g} s per GPU ({audio_s / 3600:.2f} h in total), "
                               "cq2 codec, hard codes packed to 336-byte frame records, utterance filters + framing + overlap-add on the GPU",
                   "frames_per_gpu_per_step": int(frames[0]), "conv_precision": args.precision,
                   "l2_policy": "inputs larger than L2", "parallelism": f"dp{world} (utterances sharded by rank, no collective)"},
        "e2e": {"value": v, "unit": "x real-time", "h2d_bytes_per_step": int(n_utt * T * 4 * world),
                "d2h_bytes_per_step": int(sum(o.numel() * o.element_size() for o in out_host.values()) * world)},
        "gpu_launches": int(launches), "record_kbps": 336 * 8 * 16000 / 480 / 1000.0}


def run_corpus(args):
    import torch.distributed as dist
    from nsc_b200 import _lib
    world, rank, local, dev = init_dist()
    rec = measure_corpus(args, world, rank, dev, _lib.load(), args.steps, args.warmup)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


def measure_train(args, world, rank, dev, lib, steps, warmup, breakdown=True, resnet_type='bottleneck'):
    """BASELINE.json configs[3]: full CQ training step -- forward keeping activations, backward of every kernel, ONE flat
    all-reduce (NCCL) of gradients + soft histograms, TF1 Adam -- `_finetuning_lpc`-shaped loss, train_batch frames per GPU."""
    import torch
    from nsc_b200 import codec, lpc_utilities as lu
    from nsc_b200.training import CQTrainer
    B = args.train_batch
    cfg = codec.CodecConfig(resnet_type=resnet_type, precision=args.precision)
    cm = codec.CMRL([codec.NeuralCodec(cfg, device=dev, seed=5), codec.NeuralCodec(cfg, device=dev, seed=6)], res_scalar=1.0)
    tr = CQTrainer.finetuning_lpc(cm, (60.0, 10.0, 10.0, 0.0), lr=2e-6)
    x_np, win_np = synth_audio(B, seed=4321 + rank)
    x = torch.from_numpy(x_np).to(dev) * 0.3
    lsf = lu.lpc_analysis_windows(torch.from_numpy(win_np).to(dev), 16, dtype=torch.float32)
    out = {}

    def step():
        out['r'] = tr.step(x, lsf)

    for _ in range(warmup):
        step()
    l0 = lib.nsc_launch_count()
    t = _timed(step, steps, world, dev)
    launches = lib.nsc_launch_count() - l0
    # the per-kernel breakdown is one more step: EVERY rank runs it (it contains the all-reduce), rank 0 reports
    agg = kernel_breakdown(lib, step) if breakdown else None
    _barrier(world)
    fps = B * world * steps / t
    rec = {"metric": "CQ training step throughput (frames/s; seconds of audio per second = x0.030)", "value": fps,
           "unit": "frames/s", "x_real_time": fps * SEC_PER_FRAME, "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": t / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32 weight gradients, epilogues and Adam; forward and data-gradient convs " + args.precision,
           "data": "synthetic",
           "config": {"workload": f"train: 2-codec CQ cascade ('{resnet_type}' blocks), finetuning_lpc loss (60/10/10), soft path, TF1 Adam, "
                                  "one flat all-reduce of gradients + soft histograms", "frames_per_gpu": B, "parallelism": f"dp{world}"},
           "gpu_launches": int(launches), "launches_per_step": int(launches) // max(1, steps),
           "collectives_per_step": getattr(tr, 'collectives_per_step', None),
           "loss_first_frame": float(out['r']['loss_vector'][0])}
    if agg is not None:
        bt = breakdown_table(agg)
        rec["kernel_breakdown"] = bt
        ar = [v for k, v in bt.items() if 'allreduce' in k or 'all_reduce' in k]
        rec["allreduce_share"] = round(sum(v['share'] for v in ar), 4) if ar else (0.0 if world == 1 else None)
    return rec


def run_train(args):
    import torch.distributed as dist
    from nsc_b200 import _lib
    world, rank, local, dev = init_dist()
    rec = measure_train(args, world, rank, dev, _lib.load(), args.steps, args.warmup)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


def measure_codec1(args, world, rank, dev, lib, cpu=True):
    """BASELINE.json configs[0]: ONE neural codec (num_resnets = 1), no LPC, eval forward (hard codes) on synthetic 512-sample
    frames, batch 128 -- the shape of the reference's only timing probe (cmrl.py:513-543, :584-611).  Device-resident and
    end-to-end (pinned host frames in, codes + decoded audio back) x real-time, plus the oracle on the host cores."""
    import torch
    from nsc_b200 import codec
    B = 128
    cfg = codec.CodecConfig(resnet_type='bottleneck', precision=args.precision)
    gc = codec.NeuralCodec(cfg, device=dev, seed=5)
    x_np, _ = synth_audio(B, seed=99 + rank)
    x_np = (x_np / 33.461480140686035).astype(np.float32)          # pure-time-domain normalisation (constants.py:16)
    xh = torch.from_numpy(x_np).pin_memory()
    xd = xh.to(dev)
    outs = {}

    def step_dev():
        outs['r'] = gc.computational_graph_end2end_quan_on(xd, False, 1.0)

    def step_e2e():
        r = gc.computational_graph_end2end_quan_on(xh.to(dev, non_blocking=True), False, 1.0)
        for k in ('idx', 'out'):
            if k not in outs:
                outs[k] = torch.empty(r[k].shape, dtype=r[k].dtype).pin_memory()
            outs[k].copy_(r[k], non_blocking=True)

    n = 50
    for _ in range(5):
        step_dev(); step_e2e()
    l0 = lib.nsc_launch_count()
    t_cold = _timed(step_dev, n, world, dev)               # every call clears the images' zero rows and packs the weights
    launches_cold = (lib.nsc_launch_count() - l0) // n
    gc.prepare(B)                                          # a serving loop does that once per weights (nsc_prepare)
    for _ in range(5):
        step_dev(); step_e2e()
    l0 = lib.nsc_launch_count()
    t_dev = _timed(step_dev, n, world, dev)
    launches = (lib.nsc_launch_count() - l0) // n
    t_e2e = _timed(step_e2e, n, world, dev)
    # the same call captured once in a CUDA graph (codec.GraphedCall) and replayed
    t_graph = None
    try:
        g = gc.graphed_forward(B)
        for _ in range(5):
            g.run(xd)
        t_graph = _timed(lambda: g.run(xd), n, world, dev)
        del g
    except Exception:
        t_graph = None
    gc.release()
    rec = {"workload": "codec1: one bottleneck codec ('9 9 100 20 1 2', stride 2, 32 bins), no LPC, hard codes, batch 128 per GPU "
                       "(BASELINE.json configs[0]); weights prepared once (nsc_prepare), as TensorFlow keeps its variables resident",
           "value": B * world * n * SEC_PER_FRAME / t_dev, "unit": "x real-time", "ms_per_call": t_dev / n * 1e3,
           "e2e": {"value": B * world * n * SEC_PER_FRAME / t_e2e, "unit": "x real-time", "ms_per_call": t_e2e / n * 1e3,
                   "h2d_bytes_per_step": B * 512 * 4 * world, "d2h_bytes_per_step": (B * 256 + B * 512 * 4) * world},
           "launches_per_call": int(launches), "steps": n,
           "cuda_graph_replay": None if t_graph is None else {"value": B * world * n * SEC_PER_FRAME / t_graph, "ms_per_call": t_graph / n * 1e3},
           "unprepared": {"value": B * world * n * SEC_PER_FRAME / t_cold, "ms_per_call": t_cold / n * 1e3, "launches_per_call": int(launches_cold),
                          "note": "the same call when every call packs the weights and clears the image borders"}}
    if cpu and rank == 0 and world == 1:
        from oracle import ref_codec
        torch.set_num_threads(os.cpu_count())
        oc = ref_codec.OracleCodec(ref_codec.OracleCodecCfg(), seed=5)
        xt = torch.from_numpy(x_np)[:, :, None]
        with torch.no_grad():
            oc.forward(xt, False, 1.0)
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                oc.forward(xt, False, 1.0)
            dt = (time.perf_counter() - t0) / reps
        rec["cpu_baseline"] = {"value": B * SEC_PER_FRAME / dt, "unit": "x real-time", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"the same 128-frame batch, {reps} calls of the restated reference (torch-CPU oracle), {dt * 1e3:.0f} ms per call"}
    return rec


def measure_variant(args, world, rank, dev, lib, resnet_type, strides, frames=4144):
    """The reference's SHIPPED switches (constants.py:14 resnet_type = 'gln'; the_strides '4' -> [2, 2], cmrl.py:804) on the same
    CQ 2-codec workload.  All of them run on the plane engine in the fp32-class mode (fused k15 gate pair, depthwise + pointwise
    up-conv, three resolution levels); `engine` says which engine ran."""
    import torch
    from nsc_b200 import codec, lpc_utilities as lu
    cfg = codec.CodecConfig(resnet_type=resnet_type, the_strides=strides, precision=args.precision)
    cm = codec.CMRL([codec.NeuralCodec(cfg, device=dev, seed=5 + i) for i in range(2)], res_scalar=1.0)
    B = frames
    x_np, win_np = synth_audio(min(B, 2072), seed=55 + rank)
    reps = -(-B // x_np.shape[0])
    xd = torch.from_numpy(np.tile(x_np, (reps, 1))[:B]).to(dev)
    wd = torch.from_numpy(np.tile(win_np, (reps, 1))[:B]).to(dev)

    def step():
        cm.feedforward_lpc(xd, lu.lpc_analysis_windows(wd, 16, dtype=torch.float32), False, 1.0)

    cm.prepare(B)                                   # once per weights (nsc_prepare), as in the headline
    for _ in range(3):
        step()
    n = 5
    t = _timed(step, n, world, dev)
    agg = kernel_breakdown(lib, step)
    fl = sum(v[1] for v in agg.values())
    ms = sum(v[0] for v in agg.values())
    hbm, bf16, bf16_sus, how = peaks()
    return {"workload": f"cq2 with resnet_type '{resnet_type}', strides {list(strides)}: LPC + LSF codebook + 2 codecs + synthesis, hard codes, "
                        f"{B} frames per GPU; weights prepared once (nsc_prepare)", "value": B * world * n * SEC_PER_FRAME / t, "unit": "x real-time", "ms_per_call": t / n * 1e3,
            "engine": ("plane engine (tcgen05, fp16 hi/lo plane images between layers)" if lib.nsc_codec_on_plane_engine(C.byref(cfg.to_struct())) == 1
                       else "layer-by-layer (first tensor engine tcgen05 fp16 hi/lo + CUDA-core stem/heads/depthwise)"),
            "tflops_algorithmic": fl / (ms * 1e-3) / 1e12, "tensor_frac_of_measured_bf16_sustained": fl / (ms * 1e-3) / 1e12 / bf16_sus,
            "top_kernels": {k: {"ms": round(v[0], 3), "launches": v[3]} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:6]}}


def measure_cq_sweep(args, world, rank, dev, lib):
    """BASELINE.json configs[2]: CQ scaled towards 24 kbps -- more cascaded codecs and bins -- batch-size sweep, encode+decode,
    device-resident and end-to-end."""
    import torch
    from nsc_b200 import codec, lpc_utilities as lu
    rows = []
    for n_codecs, bins in ((3, 32), (4, 64)):
        cfg = codec.CodecConfig(resnet_type='bottleneck', precision=args.precision, num_bins=bins)
        cm = codec.CMRL([codec.NeuralCodec(cfg, device=dev, seed=5 + i) for i in range(n_codecs)], res_scalar=1.0)
        for B in (128, 2072, 8288):
            x_np, win_np = synth_audio(min(B, 2072), seed=77 + rank)
            reps = -(-B // x_np.shape[0])
            xh = torch.from_numpy(np.tile(x_np, (reps, 1))[:B]).pin_memory()
            wh = torch.from_numpy(np.tile(win_np, (reps, 1))[:B]).pin_memory()
            xd, wd = xh.to(dev), wh.to(dev)
            keep = {}

            def step_dev():
                keep['r'] = cm.feedforward_lpc(xd, lu.lpc_analysis_windows(wd, 16, dtype=torch.float32), False, 1.0)

            def step_e2e():
                a, b = xh.to(dev, non_blocking=True), wh.to(dev, non_blocking=True)
                r = cm.feedforward_lpc(a, lu.lpc_analysis_windows(b, 16, dtype=torch.float32), False, 1.0)
                outs = [('lsf_idx', r['lsf_idx']), ('syn', r['synthesized'])] + [('idx%d' % k, t) for k, t in enumerate(r['idx'])]
                for k, t in outs:
                    kk = (k, B)
                    if kk not in keep:
                        keep[kk] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                    keep[kk].copy_(t, non_blocking=True)

            n = 20 if B <= 2072 else 5
            cm.prepare(B)                                  # streaming loop: zero rows + packed weights once per weights (nsc_prepare)
            for _ in range(3):
                step_dev(); step_e2e()
            t_dev = _timed(step_dev, n, world, dev)
            t_e2e = _timed(step_e2e, n, world, dev)
            cm.release()
            rows.append({"codecs": n_codecs, "bins": bins, "frames_per_gpu": B, "value": B * world * n * SEC_PER_FRAME / t_dev,
                         "e2e": B * world * n * SEC_PER_FRAME / t_e2e, "unit": "x real-time", "ms_per_call": t_dev / n * 1e3})
    return {"workload": "cq3 (3 codecs x 32 bins) and cq4x64 (4 codecs x 64 bins): LPC + LSF codebook + cascade + synthesis, hard codes, "
                        "batch sweep (BASELINE.json configs[2]); weights prepared once per batch size (nsc_prepare)", "rows": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--frames', type=int, default=32768, help='frames per GPU per step')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cpu-frames', type=int, default=8192, help='bounded CPU-baseline sample (frames)')
    ap.add_argument('--ref-frames', type=int, default=512, help='frames per step of the reference arm')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='cq2', choices=['cq2', 'train', 'corpus'],
                    help="cq2 = headline encode+decode; train = training step; corpus = wav -> packed codes -> wav")
    ap.add_argument('--codecs', type=int, default=2, help='cascaded codecs (BASELINE.json configs[2] scales this and --bins)')
    ap.add_argument('--bins', type=int, default=32, help='code bins per codec')
    ap.add_argument('--utterances', type=int, default=360, help='corpus workload: utterances per GPU')
    ap.add_argument('--utt-seconds', type=float, default=10.0, help='corpus workload: seconds per utterance')
    ap.add_argument('--train-batch', type=int, default=128, help='frames per GPU per training step')
    ap.add_argument('--profile-one-step', default=None, metavar='JSON',
                    help="profiling aid: warm up, run ONE step inside cudaProfilerStart/Stop, write the launch names to JSON, exit")
    ap.add_argument('--no-sub-records', dest='sub_records', action='store_false',
                    help="skip the bounded sub-records (codec1 batch 128, scaled CQ sweep, training step, 1-hour corpus) of the default run")
    ap.add_argument('--precision', default='tc_f16x3', choices=['fp32', 'tc_f16x3', 'tc_f16'],
                    help="conv arithmetic: fp32 FFMA, tcgen05 fp16 hi/lo split (fp32-class, default), tcgen05 fp16 (reduced)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'train':
        run_train(args)
    elif args.workload == 'corpus':
        run_corpus(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
