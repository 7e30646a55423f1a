"""End-to-end (pinned host in, pinned host out) Griffin-Lim step time of the bench workload: one compute stream against
two overlapping ones, for several cut lists (NSB_CHUNK_CUTS, frame positions)."""
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
cut_lists = sys.argv[1:] or ["", "8000,52000", "8000,32000,52000", "8000,24000,40000,54000", "6000,20000,34000,48000,58000",
                             "4000,16000,28000,40000,52000,60000", "16000,32000,48000", "12000,26000,40000,54000"]
for cuts in cut_lists:
    if cuts.startswith("wave"):                      # wave[:K[:first[:growth]]] - the wave schedule and its tuning hooks
        parts = cuts.split(":")
        os.environ["NSB_CHUNK_CUTS"] = ""
        for key, val in zip(("NSB_WAVES", "NSB_WAVE_FIRST", "NSB_WAVE_GROWTH"), parts[1:]):
            os.environ[key] = val
        h.set_option(_lib.OPT_WAVE_SCHEDULE, 1)
    else:
        h.set_option(_lib.OPT_WAVE_SCHEDULE, 0)
        os.environ["NSB_CHUNK_CUTS"] = cuts
    for overlap in (0, 1):
        h.set_option(_lib.OPT_OVERLAP_CHUNKS, overlap)
        for _ in range(2):
            batch.inv_spectrogram_batch(pin_in.array, seed=1, iters=60, out=pin_out.array)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 8
        for i in range(n):
            batch.inv_spectrogram_batch(pin_in.array, seed=2 + i, iters=60, out=pin_out.array)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / n
        print("cuts %-40s overlap %d: %.2f ms/step  %.0f audio-s/s" % (cuts or "(automatic 16k|rest|12k)", overlap, ms,
                                                                      N * h.num_samples(T) / 20000.0 / (ms * 1e-3)), flush=True)
