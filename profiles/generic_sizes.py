"""Throughput of the generic-size kernels (csrc/gen_kernels.cuh) against the fused n_fft = 2048 path: Griffin-Lim (60 iterations,
device-resident, 16 x 1000 frames) and feature extraction for several num_freq at their natural hop / window (12.5 / 50 ms)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

st = torch.cuda.current_stream().cuda_stream
N, T, ITERS = 16, 1000, 60
for over in ("num_freq=1025", "num_freq=513,sample_rate=16000,frame_length_ms=50", "num_freq=2049", "num_freq=2048,sample_rate=24000", "num_freq=401,sample_rate=16000"):
    hp = hparams.load()
    hp.parse(over)
    h = audio._handle()
    F = h.num_freq
    spec = torch.rand((N * T, F), device="cuda")
    out = torch.empty(N * h.num_samples(T), dtype=torch.float64, device="cuda")
    flags = _lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS

    def gl():
        h.griffin_lim(spec, _lib.FRAME_MAJOR, [T] * N, out, seed=1, iters=ITERS, flags=flags, out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)
    gl()
    h.check_status(st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        gl()
    e1.record()
    torch.cuda.synchronize()
    ms_gl = e0.elapsed_time(e1) / 3
    n = h.hop * 4000
    wav = torch.rand(N * n, device="cuda") - 0.5
    Tn = h.num_frames(n)
    lin = torch.empty((N * Tn, F), device="cuda")
    mel = torch.empty((N * Tn, 80), device="cuda")

    def feat():
        h.features(wav, [n] * N, lin, mel, space=_lib.DEVICE, stream=st)
    feat()
    h.check_status(st)
    e0.record()
    for _ in range(5):
        feat()
    e1.record()
    torch.cuda.synchronize()
    ms_f = e0.elapsed_time(e1) / 5
    print("%-48s n_fft %5d hop %4d win %5d | Griffin-Lim %d x %d frames x %d it: %8.2f ms = %6.1f M frame-iterations/s | features %6.1f M frames/s" % (
        over, h.n_fft, h.hop, h.win, N, T, ITERS, ms_gl, N * T * ITERS / ms_gl / 1e3, N * Tn / ms_f / 1e3), flush=True)
    del spec, out, wav, lin, mel
hparams.load()
