"""End-to-end step time of the bench workload (64 x 1000 frames, pinned host in / out) through the synchronous call and
through the asynchronous submit / wait API with 1, 2 and 3 batches in flight; then the same from PAGEABLE numpy arrays."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nspeech_b200 import _lib, audio, batch, hparams  # noqa: E402

hparams.load()
N, T, F, ITERS = 64, 1000, 1025, 60
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rs = np.random.RandomState(0)
n_samp = 250 * (T - 1)
flags = _lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS


def bufs(k, pinned):
    ins, outs = [], []
    for i in range(k):
        if pinned:
            a = _lib.PinnedArray((N, T, F), np.float32); b = _lib.PinnedArray((N * n_samp,), np.float64)
            a.array[...] = rs.rand(N, T, F).astype(np.float32)
            ins.append(a); outs.append(b)
        else:
            ins.append(type("P", (), {"array": rs.rand(N, T, F).astype(np.float32)})()); outs.append(type("P", (), {"array": np.empty(N * n_samp, np.float64)})())
    return ins, outs


for pinned, one_chunk in ((True, 0), (True, 1), (False, 0), (False, 1)):
    for slots in ((0, 2, 3) if not one_chunk else (2, 3)):
        h = _lib.Handle(hparams.get_hparams(), 0)
        h.set_host_chunks(1 if one_chunk else 0)
        k = max(1, slots)
        ins, outs = bufs(k, pinned)
        if slots:
            h.set_async_slots(slots)

        def run(n):
            if not slots:
                for i in range(n):
                    h.griffin_lim(ins[0].array, _lib.FRAME_MAJOR, [T] * N, outs[0].array, seed=i, iters=ITERS, flags=flags, out_dtype=_lib.F64)
                return
            pend = []
            for i in range(n):
                if len(pend) == k:
                    h.wait(pend.pop(0))
                pend.append(h.griffin_lim_submit(ins[i % k].array, _lib.FRAME_MAJOR, [T] * N, outs[i % k].array, seed=i, iters=ITERS, flags=flags, out_dtype=_lib.F64))
            for t in pend:
                h.wait(t)
        run(3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(steps)
        ms = (time.perf_counter() - t0) * 1e3 / steps
        print("%-8s host buffers, %s, %s: %.2f ms per step, %.0f audio-s/s" % ("pinned" if pinned else "pageable", "one chunk per call" if one_chunk else "wave schedule",
              "synchronous call" if not slots else "submit/wait, %d in flight" % slots, ms, N * n_samp / 20000.0 / (ms * 1e-3)), flush=True)
        h.close()
        del ins, outs
