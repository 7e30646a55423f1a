"""Timing sweep of the Griffin-Lim iteration kernel (device-resident, CUDA events): kernel variant x tile x batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

hparams.load()
h = audio._handle()
T = 1000
st = torch.cuda.current_stream().cuda_stream
cases = [(64, 0, 0), (64, 0, 2), (64, 0, 1), (8, 0, 0), (8, 0, 2), (1, 0, 0), (1, 0, 2), (256, 0, 0), (256, 0, 2)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in c.split(",")) for c in sys.argv[1:]]
for batch, tile, generic in cases:
  for sm in (2,):
      h.set_option(_lib.OPT_STREAM_SYNC_MODE, sm)
      spec = torch.rand((batch, T, 1025), device="cuda")
      out = torch.empty(batch * h.num_samples(T), dtype=torch.float64, device="cuda")
      if tile > 28:      # k_gl_stream chunks of tile/4 groups
          h.set_tile_hops(0); h.set_stream_grid(-(-(batch * (T // 4 + 1)) // (tile // 4)))
      else:
          h.set_tile_hops(tile); h.set_stream_grid(0)
      h.set_generic_iteration(generic)
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
      print("sync %d batch %4d tile %2d %s: %.4f ms/iter  %.2f ns/frame-iter  -> %.0f audio-s/s at 61 passes" % (
          sm, batch, tile, {0: "k_gl_stream", 1: "k_synth<Y> ", 2: "k_gl_iter  "}[generic], ms, ms * 1e6 / (batch * T), batch * T * 0.0125 / (ms * 1e-3 * 61)), flush=True)
      del spec, out
h.set_tile_hops(0)
h.set_generic_iteration(-1)
h.set_stream_grid(0)
h.set_option(_lib.OPT_STREAM_SYNC_MODE, 2)
