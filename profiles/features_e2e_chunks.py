"""End-to-end feature extraction (pinned host wav in, pinned host spectrogram + mel out, BASELINE config 2 sample) against
the number of pipeline chunks of the host call (0 = automatic, ~24 MB of results per chunk)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

hparams.load()
h = audio._handle()
rs = np.random.RandomState(1234)
durs = np.clip(rs.normal(6.57, 2.19, size=512), 1.0, 10.0)
ns = [int(d * 20000) for d in durs]
wav = _lib.PinnedArray((sum(ns),), np.float32)
wav.array[...] = (0.3 * rs.standard_normal(sum(ns))).astype(np.float32)
Tn = [h.num_frames(n) for n in ns]
lin = _lib.PinnedArray((sum(Tn), 1025), np.float32)
mel = _lib.PinnedArray((sum(Tn), 80), np.float32)
for chunks in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64, 0]:
    h.set_host_chunks(chunks)
    h.features(wav.array, ns, lin.array, mel.array)
    t0 = time.perf_counter()
    for _ in range(5):
        h.features(wav.array, ns, lin.array, mel.array)
    ms = (time.perf_counter() - t0) * 1e3 / 5
    out_gb = sum(Tn) * (1025 + 80) * 4 / 1e9
    print("host_chunks %2d: %.2f ms for %d frames -> %.2f M mel frames/s, %.1f GB/s of results" % (
        chunks, ms, sum(Tn), sum(Tn) / ms / 1e3, out_gb / (ms * 1e-3)), flush=True)
h.set_host_chunks(0)
