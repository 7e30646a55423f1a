"""Timing sweep of the Griffin-Lim iteration kernel (device-resident, CUDA events): tile size x batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

hparams.load()
h = audio._handle()
T = 1000
st = torch.cuda.current_stream().cuda_stream
for batch in (64, 8, 1, 256):
    spec = torch.rand((batch, T, 1025), device="cuda")
    out = torch.empty(batch * h.num_samples(T), dtype=torch.float64, device="cuda")
    for tile in (0, 29, 21, 13, 5):
        h.set_tile_hops(tile)
        h.griffin_lim(spec, _lib.FRAME_MAJOR, [T] * batch, out, seed=1, iters=2, flags=3, out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)
        h.griffin_lim_iterate(20, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 100
        e0.record()
        h.griffin_lim_iterate(n, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("batch %4d tile %2d: %.4f ms/iter  %.2f ns/frame-iter  -> %.0f audio-s/s at 61 passes" % (
            batch, tile, ms, ms * 1e6 / (batch * T), batch * T * 0.0125 / (ms * 1e-3 * 61)), flush=True)
    del spec, out
h.set_tile_hops(0)
