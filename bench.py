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
  e2e   : the public Python API (nspeech_b200.batch.inv_spectrogram_batch) on pinned HOST buffers, H2D and
          D2H inside the timed region
  roofline : the Griffin-Lim iteration kernel (k_gl_stream; one launch runs all 60 iterations) timed alone with CUDA events;
          algorithmic bytes = 6,100 B per frame per iteration against MEASURED_PEAKS.json's HBM copy bandwidth;
          the FP32-side numbers are reported beside it because the fused iteration is FP32-bound
  cpu_baseline : the numpy oracle (a port of the reference's librosa path; the reference itself cannot be
          imported here) on this box's host cores, on a bounded sample of the same workload

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
NCU_TRAFFIC_BYTES_PER_LAUNCH = 20.580e9 + 3.885e9
NCU_TRAFFIC_SOURCE = "profiles/r1/ncu_full_k_gl_stream_fused60.txt (dram__bytes_read.sum + dram__bytes_write.sum of one 60-iteration launch)"


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
            "l2": "working set per step (270 MB magnitudes + 128 MB waveforms) exceeds the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-features", action="store_true")
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
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

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

    # ---- synthetic inputs: pinned host copy for the e2e leg, device copy for the kernel-only leg ----
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    pin_in = _lib.PinnedArray((N_UTT, N_FRAMES, N_BINS), np.float32)
    pin_in.array[...] = torch.rand((N_UTT, N_FRAMES, N_BINS), generator=gen, dtype=torch.float32).numpy()
    n_samp = HOP * (N_FRAMES - 1)
    pin_out = _lib.PinnedArray((N_UTT * n_samp,), np.float64)
    d_spec = torch.from_numpy(pin_in.array).to(dev)
    d_out = torch.empty(N_UTT * n_samp, dtype=torch.float64, device=dev)
    Ts = [N_FRAMES] * N_UTT
    flags = _lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS
    audio_s_per_step = N_UTT * n_samp / float(SR)

    def step_device(seed):
        h.griffin_lim(d_spec, _lib.FRAME_MAJOR, Ts, d_out, init_phase=None, seed=seed, iters=ITERS, flags=flags,
                      out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)

    def step_host(seed):
        batch.inv_spectrogram_batch(pin_in.array, seed=seed, iters=ITERS, out=pin_out.array)

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

    # ---- the dominant kernel alone: one launch of k_gl_iter = all ITERS Griffin-Lim iterations over the batch ----
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

    # ---- end to end through the public API on pinned host buffers ----
    for i in range(max(1, args.warmup // 2)):
        step_host(i)
    barrier()
    sampler.mark()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host(200 + i)
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
    sampler.mark()
    clocks = sampler.stop()
    barrier()
    assert np.isfinite(pin_out.array[:1000]).all()

    # ---- secondary metric: feature extraction (mel frames/s), BASELINE config 2 sample ----
    extra = {}
    if not args.no_features:
        rs = np.random.RandomState(1234)
        durs = np.clip(rs.normal(6.57, 2.19, size=512), 1.0, 10.0)
        ns = [int(d * SR) for d in durs]
        wav_host = _lib.PinnedArray((sum(ns),), np.float32)
        wav_host.array[...] = (0.3 * rs.standard_normal(sum(ns))).astype(np.float32)
        Tn = [h.num_frames(n) for n in ns]
        d_wav = torch.from_numpy(wav_host.array).to(dev)
        d_lin = torch.empty((sum(Tn), N_BINS), dtype=torch.float32, device=dev)
        d_mel = torch.empty((sum(Tn), 80), dtype=torch.float32, device=dev)
        for _ in range(3):
            h.features(d_wav, ns, d_lin, d_mel, space=_lib.DEVICE, stream=st)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(10):
            h.features(d_wav, ns, d_lin, d_mel, space=_lib.DEVICE, stream=st)
        f1.record()
        torch.cuda.synchronize()
        ms_feat = f0.elapsed_time(f1) / 10
        lin_h = _lib.PinnedArray((sum(Tn), N_BINS), np.float32)
        mel_h = _lib.PinnedArray((sum(Tn), 80), np.float32)
        h.features(wav_host.array, ns, lin_h.array, mel_h.array)
        t0 = time.perf_counter()
        for _ in range(3):
            h.features(wav_host.array, ns, lin_h.array, mel_h.array)
        ms_feat_e2e = (time.perf_counter() - t0) * 1e3 / 3
        extra["features"] = {
            "workload": "BASELINE config 2 sample: 512 clips, durations clip(N(6.57,2.19),1,10) s, spectrogram+melspectrogram in one pass",
            "mel_frames_per_s_device": sum(Tn) / (ms_feat * 1e-3), "mel_frames_per_s_e2e": sum(Tn) / (ms_feat_e2e * 1e-3),
            "frames": sum(Tn), "hbm_gbs_algorithmic": sum(Tn) * (4 * HOP + 4 * N_BINS + 4 * 80) / (ms_feat * 1e-3) / 1e9,
            "hbm_frac": sum(Tn) * (4 * HOP + 4 * N_BINS + 4 * 80) / (ms_feat * 1e-3) / 1e9 / peaks["hbm_gbs"]}

    value = world * audio_s_per_step * args.steps / (ms_dev * 1e-3)
    e2e_value = world * audio_s_per_step * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": "griffin_lim_audio_sec_per_sec", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(), host_binding=("rank 0 pinned to its GPU's %d NUMA-local CPUs" % len(host_cpus)) if host_cpus else "none"),
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(pin_in.array.nbytes),
                "d2h_bytes_per_step": int(pin_out.array.nbytes), "ms_per_step": ms_e2e / args.steps},
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
