"""A/B of the round-2 switches of the iteration kernel at the bench shape (64 x 1000 frames, 60 iterations, device-resident):
NSB_OPT_STREAM_BULK (bulk asynchronous copies for the magnitude rows / staged samples) off and on, alternating; CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

hparams.load()
h = audio._handle()
st = torch.cuda.current_stream().cuda_stream
N, T, ITERS = 64, 1000, 60
g = torch.Generator(device="cuda").manual_seed(1)
spec = torch.rand((N * T, 1025), device="cuda", generator=g)
out = torch.empty(N * 250 * (T - 1), dtype=torch.float64, device="cuda")
flags = _lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS
ref = None
modes = [int(v) for v in sys.argv[1:]] or [0, 1]
for rep in range(2):
    for bulk in modes:
        h.set_option(_lib.OPT_STREAM_BULK, bulk)
        h.griffin_lim(spec, _lib.FRAME_MAJOR, [T] * N, out, seed=3, iters=ITERS, flags=flags, out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)
        h.check_status(st)
        if ref is None:
            ref = out.clone()
        same = bool(torch.equal(out, ref))
        h.griffin_lim_iterate(ITERS, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            h.griffin_lim_iterate(ITERS, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("bulk %d: %.3f ms per launch of %d iterations (%.4f ms per iteration), waveform identical to the first run: %s" % (bulk, ms, ITERS, ms / ITERS, same), flush=True)
h.set_option(_lib.OPT_STREAM_BULK, 1)
