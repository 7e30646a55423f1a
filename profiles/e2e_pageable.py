"""The asynchronous Griffin-Lim stream (bench workload: 64 x 1000 frames, 60 iterations) fed from PAGEABLE numpy arrays, results in
pooled page-locked memory, against the number of batches in flight; page-locked inputs beside it.  Usage: python profiles/e2e_pageable.py IN_FLIGHT  (one process per setting: the worker slots are fixed at the first submit)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nspeech_b200 import _lib, audio, batch, hparams  # noqa: E402

hparams.load()
N, T, F = 64, 1000, 1025
in_flight = int(sys.argv[1]) if len(sys.argv) > 1 else 3
rs = np.random.default_rng(0)
h = audio._handle()
pageable = [rs.random((N, T, F), dtype=np.float32) for _ in range(3)]
pinned = [_lib.PinnedArray((N, T, F), np.float32) for _ in range(3)]
for p, q in zip(pinned, pageable):
    p.array[...] = q


def run(inputs, in_flight, n):
    """wall time of n steps from the first submit to the last result (ramp and drain included, as bench.py does), per step"""
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for wavs in batch.inv_spectrogram_stream((inputs[i % len(inputs)] for i in range(n)), seed=1, in_flight=in_flight):
        del wavs
    return (time.perf_counter() - t0) * 1e3 / n


h.set_async_slots(in_flight)
run([p.array for p in pinned], in_flight, 6)
run(pageable, in_flight, 6)
for n in (10, 40):
    a = run([p.array for p in pinned], in_flight, n)
    b = run(pageable, in_flight, n)
    print("%d batches in flight, %2d steps: page-locked inputs %.2f ms per step, pageable inputs %.2f ms per step" % (in_flight, n, a, b), flush=True)
