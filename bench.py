#!/usr/bin/env python
"""bench.py - headline benchmark of the nspeech spectrogram / Griffin-Lim hot path on B200.

Metric (BASELINE.json): audio-seconds synthesised per second by Griffin-Lim ``inv_spectrogram``.
Workload at N=1 (BASELINE config 3, "batched Griffin-Lim: 64 utterances at 12.5 s on 1 B200"): 64 synthetic
normalised linear spectrograms [1000 frames x 1025 bins] (U(0,1), what a random-init Tacotron emits), default
hparams (n_fft 2048, hop 250, win 1000, power 1.5, griffin_lim_iters 60).  One step = the whole
``inv_spectrogram`` path over that batch (denormalise -> dB->amp -> **power -> initial iSTFT -> 60 x (STFT,
phase renormalise, iSTFT) -> de-emphasis).  With N GPUs every rank runs its own such batch (utterance sharding,
no collective on the data path; weak scaling) and the value is the whole-job aggregate.

  value : device-resident inputs/outputs, CUDA events on the launching stream, max over ranks
  e2e   : the public Python API on HOST buffers, H2D and D2H of every step inside the timed region:
          nspeech_b200.batch.inv_spectrogram_stream (nsb_griffin_lim_submit / nsb_wait, three batches in flight) on page-locked
          input arrays; beside it `e2e_sync` (one synchronous inv_spectrogram_batch call per step, round 1's e2e) and
          `e2e_pageable` (the same stream fed from plain numpy arrays)
  roofline : the Griffin-Lim iteration kernel (k_gl_stream; one launch runs all 60 iterations) timed alone with CUDA events;
          algorithmic bytes = 6,100 B per frame per iteration against MEASURED_PEAKS.json's HBM copy bandwidth;
          the FP32-side numbers are reported beside it because the fused iteration is FP32-bound
  cpu_baseline : the numpy oracle (a port of the reference's librosa path; the reference itself cannot be
          imported here) on this box's host cores, on a bounded sample of the same workload
  features : the second BASELINE metric, mel frames/s of spectrogram + melspectrogram over BASELINE config 2 (13,100
          LJSpeech-shaped clips), with its own roofline, e2e and cpu_baseline
  configs  : BASELINE configs 1 (single-utterance latency), 4 (the synthesis stage at batch 32) and 5 (batch x iterations sweep)

``--impl reference`` times that CPU path as the arm of its own (all host cores, one utterance per process, the
parallelism of the reference's datasets/process.py:11-18).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_UTT, N_FRAMES, N_BINS = 64, 1000, 1025
HOP, SR, ITERS = 250, 20000, 60
BYTES_PER_FRAME_ITER = 4 * N_BINS + 8 * HOP            # magnitude read + y write + y read
FLOPS_PER_FRAME_ITER = 2 * 56320 + 12 * N_BINS         # 2 real 2048-FFTs (2.5 N log2 N) + per-bin work
BYTES_PER_FRAME_FULL = (ITERS + 1) * 4 * N_BINS + (2 * ITERS + 1) * 4 * HOP + 8 * HOP + 4 * N_BINS
FP32_PEAK_TFLOPS_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12
# dram__bytes_read.sum + dram__bytes_write.sum of one k_gl_iter launch, from the committed ncu --set full capture
NCU_TRAFFIC_BYTES_PER_LAUNCH = 20.584e9 + 3.885e9
NCU_TRAFFIC_SOURCE = "profiles/r2/ncu_full_k_gl_stream_fused60.txt (dram__bytes_read.sum + dram__bytes_write.sum of one 60-iteration launch)"
# the feature kernel: DRAM bytes per frame of one k_analysis<FEATURES> launch (268,734 frames) from the committed ncu --set full capture
NCU_FEATURES_TRAFFIC_BYTES_PER_FRAME = (270.52e6 + 1130.54e6) / 268734
NCU_FEATURES_TRAFFIC_SOURCE = "profiles/r2/ncu_full_k_analysis_features.txt (dram__bytes_read.sum + dram__bytes_write.sum per frame of a 268,734-frame launch, scaled to this launch's frames)"


def oracle_hp():
    import yaml
    with open(os.path.join(ROOT, "nspeech_b200", "hparams", "audio.yaml")) as f:
        return types.SimpleNamespace(**yaml.safe_load(f))


def _cpu_one(args):
    """One utterance through the oracle's inv_spectrogram (worker process)."""
    seed, n_frames, iters = args
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"
    from oracle import audio_oracle as ao
    hp = oracle_hp()
    rs = np.random.RandomState(seed)
    S = rs.rand(N_BINS, n_frames).astype(np.float32)
    ang = np.exp(2j * np.pi * rs.rand(N_BINS, n_frames))
    t0 = time.perf_counter()
    y = ao.inv_spectrogram(S, hp, angles=ang, iters=iters)
    return time.perf_counter() - t0, len(y)


def cpu_baseline_sample(n_frames=N_FRAMES, iters=ITERS, cores=None):
    """All host cores, one utterance per process; returns (audio-s/s, cores, description)."""
    from concurrent.futures import ProcessPoolExecutor
    cores = cores or os.cpu_count() or 1
    jobs = [(1000 + i, n_frames, iters) for i in range(cores)]
    t0 = time.perf_counter()
    import multiprocessing as mp
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn")) as ex:
        res = list(ex.map(_cpu_one, jobs))
    wall = time.perf_counter() - t0
    busy = max(t for t, _ in res)            # the processes run side by side; interpreter start-up is not counted
    audio_s = sum(n for _, n in res) / float(SR)
    sample = "%d utterances x %d frames (%.2f s each), %d iters, one process per core, numpy oracle; slowest worker %.1f s, wall incl. spawn %.1f s" % (
        cores, n_frames, HOP * (n_frames - 1) / SR, iters, busy, wall)
    return audio_s / busy, cores, sample


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.windows = index, None, [], []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=5.0):
        """nvidia-smi needs a few hundred ms to print its first line: do not start a short timed region before it"""
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        self.windows.append(time.perf_counter())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # keep the samples taken inside the marked (start, end) windows = the timed regions
        wins = list(zip(self.windows[0::2], self.windows[1::2]))
        inside = [ln for t, ln in self.lines if any(a <= t <= b + 0.1 for a, b in wins)] if wins else []
        for ln in (inside or [ln for _, ln in self.lines]):
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        # "under load": samples in the upper half of the observed power range
        if sm:
            thr = (max(power) + min(power)) / 2 if power else 0
            loaded = [s for s, p in zip(sm, power) if p >= thr] or sm
            return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                    "samples": len(sm), "power_w_max": max(power) if power else None}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(reasons), "samples": 0}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    vals, sample = [], ""
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_baseline_sample()
        if i >= args.warmup:
            vals.append(v)
    value = statistics.mean(vals)
    audio_per_step = cores * HOP * (N_FRAMES - 1) / SR
    line = {
        "impl": "reference", "metric": "griffin_lim_audio_sec_per_sec", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * audio_per_step / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config():
    return {"workload": "BASELINE config 3: inv_spectrogram (Griffin-Lim, %d iters) over %d utterances x %d frames "
                        "(12.5 s each) per GPU, default hparams" % (ITERS, N_UTT, N_FRAMES),
            "batch_per_gpu": N_UTT, "frames": N_FRAMES, "num_freq": N_BINS, "n_fft": 2048, "hop": HOP, "win": 1000,
            "griffin_lim_iters": ITERS, "sharding": "by utterance, no collective",
            "reference_arm_sample": "the CPU arm times one 12.5 s utterance per host core (a bounded sample of the same workload); the rate is per audio second either way",
            "l2": "working set per step (270 MB magnitudes + 128 MB waveforms) exceeds the 126 MB L2"}


def _features_cpu_one(args):
    """spectrogram + melspectrogram of one clip through the oracle, as datasets/process.py:30,33 does (worker process)"""
    seed, n = args
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"
    from oracle import audio_oracle as ao
    hp = oracle_hp()
    rs = np.random.RandomState(seed)
    wav = (0.3 * rs.standard_normal(n)).astype(np.float32)
    t0 = time.perf_counter()
    lin = ao.spectrogram(wav, hp)
    ao.melspectrogram(wav, hp)
    return time.perf_counter() - t0, lin.shape[1]


def features_cpu_baseline(clips_per_core=3, cores=None):
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    rs = np.random.default_rng(1234)
    durs = np.clip(rs.normal(6.57, 2.19, size=cores * clips_per_core), 1.0, 10.0)
    jobs = [(2000 + i, int(d * SR)) for i, d in enumerate(durs)]
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn")) as ex:
        res = list(ex.map(_features_cpu_one, jobs, chunksize=clips_per_core))
    wall = time.perf_counter() - t0
    busy = max(sum(t for t, _ in res[c * clips_per_core:(c + 1) * clips_per_core]) for c in range(cores))
    frames = sum(n for _, n in res)
    return frames / busy, cores, "%d clips (%d per core, durations clip(N(6.57,2.19),1,10) s), spectrogram + melspectrogram per clip as datasets/process.py:30,33, numpy oracle; busiest worker %.1f s, wall incl. spawn %.1f s" % (
        len(jobs), clips_per_core, busy, wall)


def speechlike_corpus(torch, dev, n_clips, seed=1234):
    """BASELINE config 2: n_clips at 20 kHz, durations clip(N(6.57, 2.19), 1, 10) s from default_rng(seed); content = harmonic
    stack (f0 100-250 Hz, 30 harmonics 1/k, 3 Hz amplitude modulation) + white noise at -50 dB, peak 0.9.  Synthesised on the GPU
    (there is no dataset offline); returns (device float32 samples packed back to back, lengths)."""
    rs = np.random.default_rng(seed)
    durs = np.clip(rs.normal(6.57, 2.19, size=n_clips), 1.0, 10.0)
    ns = [int(d * SR) for d in durs]
    f0 = torch.from_numpy(rs.uniform(100.0, 250.0, size=n_clips)).to(dev, torch.float32)
    lens = torch.tensor(ns, device=dev)
    clip_of = torch.repeat_interleave(torch.arange(n_clips, device=dev), lens)
    starts = torch.cumsum(lens, 0) - lens
    t = (torch.arange(int(lens.sum()), device=dev) - starts[clip_of]).to(torch.float32) / SR
    w = 2 * np.pi * f0[clip_of] * t
    x = torch.zeros_like(t)
    for k in range(1, 31):
        x += torch.sin(k * w + 0.7 * k) / k
    x *= 0.6 + 0.4 * torch.sin(2 * np.pi * 3 * t)
    g = torch.Generator(device=dev).manual_seed(seed)
    x += 10 ** (-50 / 20) * torch.randn(x.shape, device=dev, generator=g)
    peak = torch.zeros(n_clips, device=dev).scatter_reduce(0, clip_of, x.abs(), reduce="amax")
    x *= 0.9 / peak[clip_of]
    return x.contiguous(), ns


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-features", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries ONE JSON line: NCCL prints its version banner to stdout while the communicator comes up (NCCL_DEBUG=VERSION
        # and above), so file descriptor 1 points at stderr until the first collective is through
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    import __graft_entry__ as ge
    ge.build()
    from nspeech_b200 import _lib, audio, batch, hparams
    audio.DEVICE = local_rank
    # every rank stages through page-locked memory on its GPU's socket (matters from 4 ranks on: two sockets share 8 GPUs)
    host_cpus = batch.bind_host_to_gpu(local_rank) if world > 1 else None
    hparams.load()
    h = audio._handle()
    assert (h.n_fft, h.hop, h.win) == (2048, HOP, 1000)
    dev = torch.device("cuda", local_rank)
    st = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def ev_ms(fn, reps, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # ---- synthetic inputs: pinned host copies for the e2e legs, device copy for the kernel-only leg ----
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    pin_in = [_lib.PinnedArray((N_UTT, N_FRAMES, N_BINS), np.float32) for _ in range(3)]      # three batches in rotation (three in flight)
    for p_ in pin_in:
        p_.array[...] = torch.rand((N_UTT, N_FRAMES, N_BINS), generator=gen, dtype=torch.float32).numpy()
    n_samp = HOP * (N_FRAMES - 1)
    pin_out = _lib.PinnedArray((N_UTT * n_samp,), np.float64)
    d_spec = torch.from_numpy(pin_in[0].array).to(dev)
    d_out = torch.empty(N_UTT * n_samp, dtype=torch.float64, device=dev)
    Ts = [N_FRAMES] * N_UTT
    flags = _lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS
    audio_s_per_step = N_UTT * n_samp / float(SR)

    def step_device(seed):
        h.griffin_lim(d_spec, _lib.FRAME_MAJOR, Ts, d_out, init_phase=None, seed=seed, iters=ITERS, flags=flags,
                      out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)

    # ---- kernel-only (device-resident) ----
    for i in range(args.warmup):
        step_device(i)
    h.check_status(st)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    sampler.mark()
    l0 = h.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_device(100 + i)
    e1.record()
    torch.cuda.synchronize()
    launches = h.kernel_launches() - l0
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    barrier()
    h.check_status(st)

    # ---- the dominant kernel alone: one launch of k_gl_stream = all ITERS Griffin-Lim iterations over the batch ----
    n_launch = 3
    h.griffin_lim_iterate(ITERS, st)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(n_launch):
        h.griffin_lim_iterate(ITERS, st)
    k1.record()
    torch.cuda.synchronize()
    sampler.mark()
    ms_launch = k0.elapsed_time(k1) / n_launch
    ms_iter = ms_launch / ITERS
    frames = N_UTT * N_FRAMES
    peaks, peak_kind = measured_peaks()
    achieved_gbs = frames * BYTES_PER_FRAME_ITER / (ms_iter * 1e-3) / 1e9
    achieved_tflops = frames * FLOPS_PER_FRAME_ITER / (ms_iter * 1e-3) / 1e12

    # ---- end to end through the public API on host buffers: every step copies its batch in and its waveforms out ----
    # headline: batch.inv_spectrogram_stream (nsb_griffin_lim_submit / nsb_wait, three batches in flight, results in pooled page-locked
    # memory); beside it the synchronous call (round 1's e2e) and the same stream fed from PAGEABLE numpy arrays
    def e2e_stream(n_steps, inputs, seed0, dtype=None):
        got = 0
        for outs in batch.inv_spectrogram_stream((inputs[i % len(inputs)] for i in range(n_steps)), seed=seed0, iters=ITERS, dtype=dtype):
            got += len(outs)
            last = outs[-1]
        assert got == n_steps * N_UTT and np.isfinite(last[:1000]).all()

    def wall_ms(fn):
        barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return max_over_ranks((time.perf_counter() - t0) * 1e3)

    pinned_inputs = [p_.array for p_ in pin_in]
    e2e_stream(max(5, args.warmup), pinned_inputs, 0)      # (also fills the pinned result pool: four blocks are alive in steady state)
    sampler.mark()
    ms_e2e = wall_ms(lambda: e2e_stream(args.steps, pinned_inputs, 200))
    sampler.mark()
    clocks = sampler.stop()

    def sync_steps(n):
        for i in range(n):
            batch.inv_spectrogram_batch(pin_in[0].array, seed=300 + i, iters=ITERS, out=pin_out.array)
    sync_steps(1)
    ms_e2e_sync = wall_ms(lambda: sync_steps(args.steps))
    # the same stream with float32 waveforms back (NOT the reference's dtype, so not the headline): half the device-to-host bytes,
    # which is what the end-to-end rate of several ranks on one host is bound by
    e2e_stream(3, pinned_inputs, 0, np.float32)
    ms_e2e_f32 = wall_ms(lambda: e2e_stream(args.steps, pinned_inputs, 300, np.float32))
    pageable_inputs = [np.array(p_.array) for p_ in pin_in]          # plain numpy memory, what a drop-in caller holds
    e2e_stream(3, pageable_inputs, 0)
    ms_e2e_pageable = wall_ms(lambda: e2e_stream(args.steps, pageable_inputs, 400))
    barrier()

    extra = {}
    # ---- second BASELINE metric: mel frames/s of spectrogram + melspectrogram, BASELINE config 2 (13,100 LJSpeech-shaped clips) ----
    if not args.no_features:
        d_wav, ns = speechlike_corpus(torch, dev, 13100)
        Tn = [h.num_frames(n) for n in ns]
        n_fr = sum(Tn)
        d_lin = torch.empty((n_fr, N_BINS), dtype=torch.float32, device=dev)
        d_mel = torch.empty((n_fr, 80), dtype=torch.float32, device=dev)
        ms_feat = max_over_ranks(ev_ms(lambda: h.features(d_wav, ns, d_lin, d_mel, space=_lib.DEVICE, stream=st), 5, warm=2))
        h.check_status(st)
        assert 0.0 < float(d_mel[:1000].mean()) <= 1.0
        del d_lin, d_mel
        # end to end from host memory on a bounded sample (every 32nd clip: the full corpus would need 30 GB of page-locked results)
        sub = list(range(0, 13100, 32))
        offs = np.concatenate([[0], np.cumsum(ns)])
        sub_wavs = [d_wav[offs[i]:offs[i + 1]].cpu().numpy() for i in sub]
        sub_frames = sum(Tn[i] for i in sub)
        del batch.features_batch(sub_wavs)[:]                # warm-up of the same size: the result blocks return to the pinned pool
        del batch.features_batch(sub_wavs)[:]
        t0 = time.perf_counter()
        feats = None
        for _ in range(3):
            del feats                                          # its page-locked result blocks return to the pool before the next call takes them
            feats = batch.features_batch(sub_wavs)
        ms_feat_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3 / 3)
        in_bytes, out_bytes = 4 * sum(len(w) for w in sub_wavs), 4 * sub_frames * (N_BINS + 80)
        del feats, d_wav
        feat_bytes = 4 * HOP + 4 * N_BINS + 4 * 80
        extra["features"] = {
            "metric": "mel_frames_per_sec", "unit": "mel frames/s", "value": world * n_fr / (ms_feat * 1e-3), "ms_per_pass": ms_feat,
            "workload": "BASELINE config 2: spectrogram + melspectrogram (one pass) over 13,100 synthetic LJSpeech-shaped clips per GPU "
                        "(durations clip(N(6.57,2.19),1,10) s from default_rng(1234), speech-like harmonic stack + noise), device-resident",
            "frames": n_fr, "audio_hours": sum(ns) / SR / 3600.0,
            "roofline": {"bound": "hbm", "kernel": "k_analysis<FEATURES>", "achieved": n_fr * feat_bytes / (ms_feat * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": n_fr * feat_bytes / (ms_feat * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "algorithmic_bytes_per_frame": feat_bytes, "traffic": NCU_FEATURES_TRAFFIC_BYTES_PER_FRAME * n_fr,
                         "traffic_source": NCU_FEATURES_TRAFFIC_SOURCE},
            "e2e": {"value": world * sub_frames / (ms_feat_e2e * 1e-3), "unit": "mel frames/s", "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(out_bytes),
                    "sample": "every 32nd clip of the corpus (%d clips, %d frames) through batch.features_batch from numpy arrays (pipeline of clip groups: nsb_features_submit / nsb_wait), results in pooled page-locked memory" % (len(sub), sub_frames)}}
        if rank == 0 and not args.no_cpu_baseline and world == 1:
            v, cores, sample = features_cpu_baseline()
            extra["features"]["cpu_baseline"] = {"value": v, "unit": "mel frames/s", "cores": cores, "kind": "port", "sample": sample}

    # ---- the other BASELINE configurations, driver-run: config 1 (latency), config 4 (synthesis stage at batch 32), config 5 (sweep) ----
    if not args.no_configs:
        cfgs = {}
        S1 = np.random.default_rng(0).random((N_BINS, 401)).astype(np.float32)
        ang = np.exp(2j * np.pi * np.random.default_rng(0).random((N_BINS, 401))).astype(np.complex64)
        audio.inv_spectrogram(S1, init_phase=ang)
        t0 = time.perf_counter()
        for _ in range(5):
            audio.inv_spectrogram(S1, init_phase=ang)
        lat = (time.perf_counter() - t0) / 5 * 1e3
        dS, dA = torch.from_numpy(S1).to(dev), torch.from_numpy(ang).to(dev)
        lat_dev = ev_ms(lambda: audio.inv_spectrogram(dS, init_phase=dA), 5)
        cfgs["config1_single_5s_utterance"] = {"latency_ms_host_numpy": lat, "latency_ms_device_arrays": lat_dev, "audio_s_per_s_host_numpy": 5.0 / (lat * 1e-3),
                                               "what": "audio.inv_spectrogram on one [1025,401] spectrogram, 60 iterations, supplied initial phase"}
        lin4 = _lib.PinnedArray((32, 1500, N_BINS), np.float32)
        lin4.array[...] = torch.rand((32, 1500, N_BINS), generator=gen, dtype=torch.float32).numpy()
        for _ in range(3):          # the same call as the timed one: workspaces, the int16 staging buffer and the pooled result block exist afterwards
            audio.synthesize_waveforms(lin4.array, peak_normalize=True, dtype=np.int16)
        t0 = time.perf_counter()
        for _ in range(5):
            w4 = audio.synthesize_waveforms(lin4.array, peak_normalize=True, dtype=np.int16)
        ms4 = max_over_ranks((time.perf_counter() - t0) / 5 * 1e3)
        cfgs["config4_synthesis_stage_batch32"] = {"ms_per_batch": ms4, "audio_s_per_s": world * 32 * (HOP * 1499 + 1000) / SR / (ms4 * 1e-3),
                                                   "what": "audio.synthesize_waveforms on [32,1500,1025] from host memory: TF-twin Griffin-Lim (60 iterations) + de-emphasis + "
                                                           "find_endpoint + save_wav scaling to int16, the stage after the network in synthesizer.py:51-53 / eval.py:43"}
        del w4, lin4
        sweep = []
        for nb in (8, 64, 256, 1024):
            spec5 = torch.rand((nb * N_FRAMES, N_BINS), device=dev)
            out5 = torch.empty(nb * n_samp, dtype=torch.float64, device=dev)
            for it in (60, 100):
                ms5 = max_over_ranks(ev_ms(lambda: h.griffin_lim(spec5, _lib.FRAME_MAJOR, [N_FRAMES] * nb, out5, seed=1, iters=it, flags=flags, out_dtype=_lib.F64,
                                                                 space=_lib.DEVICE, stream=st), 2 if nb >= 256 else 5))
                sweep.append({"batch_per_gpu": nb, "iters": it, "ms": ms5, "audio_s_per_s": world * nb * n_samp / SR / (ms5 * 1e-3)})
            h.check_status(st)
            del spec5, out5
        cfgs["config5_sweep_device_resident"] = {"gpus": world, "frames_per_utterance": N_FRAMES, "points": sweep}
        extra["configs"] = cfgs

    value = world * audio_s_per_step * args.steps / (ms_dev * 1e-3)
    e2e_value = world * audio_s_per_step * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": "griffin_lim_audio_sec_per_sec", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "host_binding": ("every rank pinned to its GPU's %d NUMA-local CPUs" % len(host_cpus)) if host_cpus else "none",
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(pin_in[0].array.nbytes),
                "d2h_bytes_per_step": int(pin_out.array.nbytes), "ms_per_step": ms_e2e / args.steps,
                "api": "nspeech_b200.batch.inv_spectrogram_stream (nsb_griffin_lim_submit / nsb_wait, three batches in flight), page-locked input arrays, results in pooled page-locked memory"},
        "e2e_sync": {"value": world * audio_s_per_step * args.steps / (ms_e2e_sync * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_e2e_sync / args.steps,
                     "api": "nspeech_b200.batch.inv_spectrogram_batch, one synchronous call per step (round 1's e2e)"},
        "e2e_float32_out": {"value": world * audio_s_per_step * args.steps / (ms_e2e_f32 * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_e2e_f32 / args.steps,
                            "d2h_bytes_per_step": int(pin_out.array.nbytes) // 2,
                            "api": "the headline stream with dtype=np.float32: waveforms come back as float32 (the reference returns float64 - not the headline)"},
        "e2e_pageable": {"value": world * audio_s_per_step * args.steps / (ms_e2e_pageable * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_e2e_pageable / args.steps,
                         "api": "the same stream fed from pageable numpy arrays (what a drop-in caller holds)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_gl_stream (one launch = all %d Griffin-Lim iterations over the batch)" % ITERS,
                     "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / peaks["hbm_gbs"],
                     "peak_source": peak_kind, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH, "traffic_source": NCU_TRAFFIC_SOURCE,
                     "ms_per_launch": ms_launch, "iterations_per_launch": ITERS, "ms_per_iteration": ms_iter,
                     "algorithmic_bytes_per_launch": frames * BYTES_PER_FRAME_ITER * ITERS,
                     "fp32": {"achieved_tflops_conventional_count": achieved_tflops, "peak_tflops_nominal": FP32_PEAK_TFLOPS_NOMINAL,
                              "frac": achieved_tflops / FP32_PEAK_TFLOPS_NOMINAL},
                     "whole_path_hbm_frac": frames * BYTES_PER_FRAME_FULL / (ms_dev / args.steps * 1e-3) / 1e9 / peaks["hbm_gbs"]},
        "clocks": clocks,
    }
    line.update(extra)
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample = cpu_baseline_sample()
            line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
