"""The five BASELINE.json configurations on one B200 (configs 1-4) and the throughput sweep (config 5, this GPU's
share).  Parity for these shapes is in tests/test_gpu_parity.py; this script only times them (CUDA events for
device-resident runs, wall clock for host-API runs) and prints a table.  Usage: python profiles/configs.py [quick]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, batch, hparams  # noqa: E402

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
hparams.load()
h = audio._handle()
st = torch.cuda.current_stream().cuda_stream
HOP, SR = h.hop, 20000


def ev_time(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gl_device(n, T, iters, reps=3):
    spec = torch.rand((n, T, 1025), device="cuda")
    out = torch.empty(n * h.num_samples(T), dtype=torch.float64, device="cuda")
    ms = ev_time(lambda: h.griffin_lim(spec, _lib.FRAME_MAJOR, [T] * n, out, seed=1, iters=iters, flags=3, out_dtype=_lib.F64,
                                       space=_lib.DEVICE, stream=st), reps)
    h.check_status(st)
    del spec, out
    return ms


print("config 1: one 5 s spectrogram [1025,401], 60 iters, host numpy API (inv_spectrogram)")
S = np.random.default_rng(0).random((1025, 401)).astype(np.float32)
ang = np.exp(2j * np.pi * np.random.default_rng(0).random((1025, 401)))
audio.inv_spectrogram(S, init_phase=ang)
t0 = time.perf_counter()
for _ in range(10):
    y = audio.inv_spectrogram(S, init_phase=ang)
dt = (time.perf_counter() - t0) / 10
print("  latency %.2f ms  -> %.0f audio-s/s (single utterance, incl. H2D/D2H)" % (dt * 1e3, 5.0 / dt))
ms = gl_device(1, 401, 60, reps=10)
print("  device-resident: %.2f ms -> %.0f audio-s/s" % (ms, 5.0 / (ms * 1e-3)))

print("config 2: spectrogram + melspectrogram over an LJSpeech-shaped corpus (13,100 clips, clip(N(6.57,2.19),1,10) s)")
rs = np.random.default_rng(1234)
n_clips = 1310 if quick else 13100
durs = np.clip(rs.normal(6.57, 2.19, size=n_clips), 1.0, 10.0)
ns = [int(d * SR) for d in durs]
Ts = [h.num_frames(n) for n in ns]
chunk = 1024
tot_frames, t_dev, t_host = 0, 0.0, 0.0
for c0 in range(0, n_clips, chunk):
    cn, cT = ns[c0:c0 + chunk], Ts[c0:c0 + chunk]
    wav_h = _lib.PinnedArray((sum(cn),), np.float32)
    wav_h.array[...] = (0.3 * rs.standard_normal(sum(cn))).astype(np.float32)
    lin_h = _lib.PinnedArray((sum(cT), 1025), np.float32)
    mel_h = _lib.PinnedArray((sum(cT), 80), np.float32)
    d_wav = torch.from_numpy(wav_h.array).cuda()
    d_lin = torch.empty((sum(cT), 1025), device="cuda")
    d_mel = torch.empty((sum(cT), 80), device="cuda")
    t_dev += ev_time(lambda: h.features(d_wav, cn, d_lin, d_mel, space=_lib.DEVICE, stream=st), 3) * 1e-3
    h.features(wav_h.array, cn, lin_h.array, mel_h.array)
    t0 = time.perf_counter()
    h.features(wav_h.array, cn, lin_h.array, mel_h.array)
    t_host += time.perf_counter() - t0
    tot_frames += sum(cT)
    wav_h.free(); lin_h.free(); mel_h.free()
    del d_wav, d_lin, d_mel
print("  %d clips, %d frames (%.1f h of audio): device-resident %.3f s = %.1f M mel frames/s; from pinned host %.3f s = %.2f M mel frames/s"
      % (n_clips, tot_frames, sum(durs) / 3600, t_dev, tot_frames / t_dev / 1e6, t_host, tot_frames / t_host / 1e6))

print("config 3: batched Griffin-Lim, 64 utterances, 60 iters, device-resident")
for T in (1000, 1500):
    ms = gl_device(64, T, 60)
    print("  T=%d: %.2f ms -> %.0f audio-s/s" % (T, ms, 64 * HOP * (T - 1) / SR / (ms * 1e-3)))

print("config 4: eval.py spectrogram->waveform stage, batch 32 x 1500 frames (random-init-Tacotron-like), host API")
specs = np.random.default_rng(3).random((32, 1500, 1025)).astype(np.float32)
pin = _lib.PinnedArray(specs.shape, np.float32)
pin.array[...] = specs
w1 = batch.inv_spectrogram_batch(pin.array, seed=1)
w2 = batch.inv_spectrogram_batch(pin.array, seed=1)       # two result blocks in the pinned pool: `outs` below holds one while the next
del w1, w2                                                # call fills the other (a fresh 96 MB cudaHostAlloc costs ~30 ms, once)
t0 = time.perf_counter()
for _ in range(6):
    outs = batch.inv_spectrogram_batch(pin.array, seed=1)
dt = (time.perf_counter() - t0) / 6
print("  %.1f ms per batch -> %.0f audio-s/s end to end" % (dt * 1e3, 32 * HOP * 1499 / SR / dt))
pin.free()

print("config 5: throughput sweep on this GPU (T=1000, device-resident)")
for iters in (60, 100):
    for n in ((8, 64, 512) if quick else (8, 16, 32, 64, 128, 256, 512, 1024)):
        ms = gl_device(n, 1000, iters, reps=2)
        print("  iters=%3d batch=%4d: %8.2f ms -> %7.0f audio-s/s" % (iters, n, ms, n * HOP * 999 / SR / (ms * 1e-3)), flush=True)
