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

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # the contract is ONE JSON line on stdout: keep NCCL's banner ("NCCL version ...") off it
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    lib = _lib.load()

    B = args.frames                      # frames per GPU per step (weak scaling)
    cfg = codec.CodecConfig(precision=args.precision, num_bins=args.bins)
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

    out_host = {}
    # End-to-end step through the public API with HOST buffers: the batch goes through in sub-batches so that the pinned-memory
    # H2D copy of sub-batch k+1 and the D2H copy of sub-batch k-1 (copy stream) run under the compute of sub-batch k.
    # Sub-batches are whole passes of the engine (2 x 2,072 frames) so no pass is split; the exposed part of the copies is the
    # first sub-batch's H2D and the last one's D2H.  H2D and D2H use separate streams (both copy engines).
    sub = 4144
    n_sub = max(1, min(8, -(-B // sub)))
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

    # ---- per-kernel breakdown of ONE step with CUDA events on the launching stream (roofline line)
    roof = None
    breakdown = None
    if rank == 0:
        agg = kernel_breakdown(lib, step_device)
        tot = sum(a[0] for a in agg.values())
        breakdown = breakdown_table(agg)
        is_conv = lambda k: (k.startswith(('conv_', 'pT', 'pX', 'pG')) or (k.startswith('tc') and k != 'tc_pack_weights'))
        conv = [(k, v) for k, v in agg.items() if is_conv(k)]
        top_name, top = max(conv, key=lambda kv: kv[1][0])
        hbm, bf16, bf16_sus, how = peaks()
        tensor_path = not top_name.startswith('conv_')
        mma_per_product = 3 if (args.precision == 'tc_f16x3' and tensor_path) else 1
        secs = top[0] * 1e-3
        ach_tf = top[1] / secs / 1e12
        ach_gbs = top[2] / secs / 1e9
        # which roof binds this kernel: time its algorithmic bytes need at the measured HBM rate vs the time its ISSUED
        # tensor flops need at the measured bf16 rate
        t_hbm = top[2] / (hbm * 1e9)
        t_tc = top[1] * mma_per_product / (bf16_sus * 1e12)
        conv_ms = sum(v[0] for _, v in conv)
        conv_fl = sum(v[1] for _, v in conv)
        conv_by = sum(v[2] for _, v in conv)
        traffic = traffic_detail = None
        tpath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
        if os.path.exists(tpath):
            t = json.load(open(tpath)).get(top_name)
            if t:   # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed ncu --set full capture
                traffic = t["dram_bytes_per_launch"]
                traffic_detail = {"frames_of_that_launch": t["frames_per_launch"],
                                  "algorithmic_bytes_of_that_launch": 589824 * t["frames_per_launch"], "source": t["source"]}
        common = {"kernel": top_name, "launch_ms": top[0] / top[3], "share_of_step": top[0] / tot, "traffic": traffic, "traffic_detail": traffic_detail,
                  "pipe": ("fp32 FFMA (CUDA cores)" if not tensor_path else
                           "tcgen05 kind::f16, fp32 accumulate in TMEM" + (" -- 3 MMAs per product (fp16 hi/lo split): "
                           "issued tensor flops are 3x the algorithmic flops" if mma_per_product == 3 else "")),
                  "achieved_tflops_algorithmic": ach_tf, "issued_tflops": ach_tf * mma_per_product,
                  "tensor_frac_of_measured_bf16_sustained": ach_tf * mma_per_product / bf16_sus,
                  "achieved_gbs_algorithmic": ach_gbs, "hbm_frac_of_measured": ach_gbs / hbm,
                  "all_conv": {"tflops": conv_fl / (conv_ms * 1e-3) / 1e12, "gbs": conv_by / (conv_ms * 1e-3) / 1e9,
                               "hbm_frac_of_measured": conv_by / (conv_ms * 1e-3) / 1e9 / hbm,
                               "share_of_step": conv_ms / tot}}
        if t_hbm >= t_tc:
            roof = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm, "unit": "GB/s", "frac": ach_gbs / hbm,
                    "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({how}); algorithmic bytes = the layer's input, residual and "
                                   "output plane images, each moved once", **common}
        else:
            roof = {"bound": "tensor", "achieved": ach_tf, "peak": bf16_sus, "unit": "TFLOP/s", "frac": ach_tf / bf16_sus,
                    "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({how}); kernel timed inside a long step", **common}

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
                       "parallelism": f"dp{world} (frames sharded by rank, no collective)"},
            "frames_per_s": frames_total * args.steps / t_dev,
            "e2e": {"value": e2e_v, "unit": "x real-time", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": t_e2e / args.steps * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "kernel_breakdown": breakdown,
            "compute_tflops_whole_step": frames_total * args.steps * args.codecs * FLOP_PER_FRAME_CODEC / t_dev / 1e12,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_corpus(args):
    """Secondary workload (BASELINE.json configs[4]): an hour-scale synthetic corpus, wav -> hard codes (packed records) -> wav,
    through nsc_b200.pipeline (cmrl.py:666-737 batched): host signals in, packed records + synthesized signals out, every step
    (normalisation, filters, framing, LPC analysis, CQ pass, overlap-add, de-emphasis, bit packing, H2D/D2H) inside the timed
    region.  Utterances are sharded across ranks."""
    import torch
    import torch.distributed as dist
    from nsc_b200 import _lib, codec, pipeline
    from nsc_b200.sharding import max_over_ranks

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    lib = _lib.load()
    cfg = codec.CodecConfig(precision=args.precision)
    cm = codec.CMRL([codec.NeuralCodec(cfg, device=dev, seed=5), codec.NeuralCodec(cfg, device=dev, seed=6)], res_scalar=1.0)
    n_utt, T = args.utterances, int(args.utt_seconds * 16000)
    x_np, _ = synth_audio(64, seed=4321 + rank)
    base = np.tile(x_np.reshape(-1), -(-T * 8 // x_np.size))        # a few distinct utterances, tiled
    host = [torch.from_numpy(np.ascontiguousarray(base[(i % 8) * 4000:(i % 8) * 4000 + T])).pin_memory() for i in range(n_utt)]
    out_host = {}

    def step():
        sigs = [h.to(dev, non_blocking=True) for h in host]
        res = pipeline.code_utterances(cm, sigs, the_share=False, pack=True)
        rec = torch.cat([r['records'] for r in res])
        syn = torch.cat([r['synthesized'] for r in res])
        for k, t in (('rec', rec), ('syn', syn)):
            if k not in out_host:
                out_host[k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            out_host[k].copy_(t, non_blocking=True)
        return sum(r['n_frames'] for r in res)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        frames = step()
    barrier()
    l0 = lib.nsc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t = max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)
    launches = lib.nsc_launch_count() - l0
    if rank == 0:
        audio_s = n_utt * world * args.utt_seconds
        v = audio_s * args.steps / t
        print(json.dumps({
            "metric": "seconds of 16 kHz audio coded per second (x real-time), corpus wav -> packed hard codes -> wav, end to end",
            "value": v, "unit": "x real-time", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32-equivalent convs (fp16 hi/lo on tensor cores); filters / LPC f64", "data": "synthetic",
            "config": {"workload": f"corpus: {n_utt} utterances x {args.utt_seconds:g} s per GPU ({audio_s / 3600:.2f} h in total), "
                                   "cq2 codec, hard codes packed to 336-byte frame records, utterance filters + framing + overlap-add on the GPU",
                       "frames_per_gpu_per_step": int(frames), "conv_precision": args.precision,
                       "l2_policy": "inputs larger than L2", "parallelism": f"dp{world} (utterances sharded by rank, no collective)"},
            "e2e": {"value": v, "unit": "x real-time", "h2d_bytes_per_step": int(n_utt * T * 4 * world),
                    "d2h_bytes_per_step": int(sum(o.numel() * o.element_size() for o in out_host.values()) * world)},
            "gpu_launches": int(launches), "record_kbps": 336 * 8 * 16000 / 480 / 1000.0}))
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """Secondary workload (BASELINE.json configs[3]): full CQ training step -- forward keeping activations, backward of
    every kernel, histogram + gradient all-reduce (NCCL), TF1 Adam -- `_finetuning_lpc`-shaped loss, 128 frames per GPU."""
    import torch
    import torch.distributed as dist
    from nsc_b200 import _lib, codec, lpc_utilities as lu
    from nsc_b200.sharding import max_over_ranks
    from nsc_b200.training import CQTrainer
    world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    lib = _lib.load()
    B = args.train_batch
    cfg = codec.CodecConfig(precision=args.precision)
    cm = codec.CMRL([codec.NeuralCodec(cfg, device=dev, seed=5), codec.NeuralCodec(cfg, device=dev, seed=6)], res_scalar=1.0)
    tr = CQTrainer.finetuning_lpc(cm, (60.0, 10.0, 10.0, 0.0), lr=2e-6)
    x_np, win_np = synth_audio(B, seed=4321 + rank)
    x = torch.from_numpy(x_np).to(dev) * 0.3
    lsf = lu.lpc_analysis_windows(torch.from_numpy(win_np).to(dev), 16, dtype=torch.float32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        tr.step(x, lsf)
    barrier()
    l0 = lib.nsc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = tr.step(x, lsf)
    e1.record()
    barrier()
    t = max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)
    # the per-kernel breakdown is one more step: EVERY rank runs it (it contains the all-reduces), rank 0 reports
    agg = kernel_breakdown(lib, lambda: tr.step(x, lsf))
    barrier()
    if rank == 0:
        fps = B * world * args.steps / t
        print(json.dumps({"metric": "CQ training step throughput (frames/s; seconds of audio per second = x0.030)", "value": fps,
                          "unit": "frames/s", "x_real_time": fps * SEC_PER_FRAME, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32 weight gradients, epilogues and Adam; forward and data-gradient convs " + args.precision,
                          "data": "synthetic",
                          "config": {"workload": "train: 2-codec CQ cascade, finetuning_lpc loss (60/10/10), soft path, TF1 Adam, "
                                                 "hist + flat-gradient all-reduce", "frames_per_gpu": B, "parallelism": f"dp{world}"},
                          "gpu_launches": int(lib.nsc_launch_count() - l0), "loss_first_frame": float(out['loss_vector'][0]),
                          "kernel_breakdown": breakdown_table(agg)}))
    if world > 1:
        dist.destroy_process_group()


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
