"""Where the end-to-end step goes: the host-API call with 0 / 1 / 60 Griffin-Lim iterations, chunked and unchunked."""
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
for chunks in (0, 1):
    h.set_host_chunks(chunks)
    for iters in (0, 1, 60):
        for _ in range(2):
            batch.inv_spectrogram_batch(pin_in.array, seed=1, iters=iters, out=pin_out.array)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 8
        for i in range(n):
            batch.inv_spectrogram_batch(pin_in.array, seed=2 + i, iters=iters, out=pin_out.array)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / n
        print("host_chunks %d (0 = automatic)  iters %2d: %.2f ms/step" % (chunks, iters, ms), flush=True)
h.set_host_chunks(0)
