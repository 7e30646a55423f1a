"""End-to-end feature extraction from a list of numpy clips (what datasets/process.py holds): batch.features_batch as one
synchronous call against the pipeline of clip groups (nsb_features_submit / nsb_wait), on the bench's bounded sample of BASELINE
config 2 (every 32nd clip of the 13,100).  Usage: python profiles/features_e2e.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from nspeech_b200 import audio, batch, hparams  # noqa: E402

hparams.load()
h = audio._handle()
dev = torch.device("cuda:0")
d_wav, ns = bench.speechlike_corpus(torch, dev, 13100)
offs = np.concatenate([[0], np.cumsum(ns)])
sub = list(range(0, 13100, 32))
wavs = [d_wav[offs[i]:offs[i + 1]].cpu().numpy() for i in sub]
frames = sum(h.num_frames(len(w)) for w in wavs)
del d_wav
print("%d clips, %d frames, %.0f MB in, %.0f MB out" % (len(wavs), frames, 4e-6 * sum(len(w) for w in wavs), 4e-6 * frames * (1025 + 80)))


def run(**kw):
    del batch.features_batch(wavs, **kw)[:]
    del batch.features_batch(wavs, **kw)[:]
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        f = batch.features_batch(wavs, **kw)
        ts.append((time.perf_counter() - t0) * 1e3)
        del f
    return min(ts), sum(ts) / len(ts)


t0 = time.perf_counter()
packed = np.concatenate(wavs)
print("np.concatenate of the clips alone (pageable destination): %.1f ms" % ((time.perf_counter() - t0) * 1e3))
del packed
best, mean = run(in_flight=0)
print("one synchronous call: best %.1f ms, mean %.1f ms -> %.2f M mel frames/s" % (best, mean, frames / mean / 1e3))
for gb in (48, 96, 192):
    for fl in (2, 3):
        best, mean = run(in_flight=fl, group_bytes=gb << 20)
        print("groups of %3d MB of results, %d in flight: best %.1f ms, mean %.1f ms -> %.2f M mel frames/s" % (gb, fl, best, mean, frames / mean / 1e3))
