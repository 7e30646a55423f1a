"""End-to-end (pinned host in, pinned host out) Griffin-Lim step time against the number of pipeline chunks."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, batch, hparams  # noqa: E402

hparams.load()
h = audio._handle()
N, T = 64, 1000
pin_in = _lib.PinnedArray((N, T, 1025), np.float32)
pin_in.array[...] = np.random.RandomState(0).rand(N, T, 1025).astype(np.float32)
pin_out = _lib.PinnedArray((N * h.num_samples(T),), np.float64)
for chunks in [int(a) for a in sys.argv[1:]] or [0, 2, 4, 6, 8, 12, 16]:
    h.set_host_chunks(chunks)
    for _ in range(2):
        batch.inv_spectrogram_batch(pin_in.array, seed=1, iters=60, out=pin_out.array)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 8
    for i in range(n):
        batch.inv_spectrogram_batch(pin_in.array, seed=2 + i, iters=60, out=pin_out.array)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / n
    print("host_chunks %2d: %.2f ms/step  %.0f audio-s/s" % (chunks, ms, N * h.num_samples(T) / 20000.0 / (ms * 1e-3)), flush=True)
h.set_host_chunks(0)
