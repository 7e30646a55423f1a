"""k_gl_iter on small batches: wide mode (tiles of C hops, one frame per warp and step) forced on / off, device-resident."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

hparams.load()
h = audio._handle()
st = torch.cuda.current_stream().cuda_stream
for T in (1000, 401):
    for batch in (1, 2, 3, 4, 6, 8, 12, 16):
        res = []
        for wide in (0, 1, -1):
            h.set_option(_lib.OPT_WIDE_MODE, wide)
            h.set_generic_iteration(2 if wide >= 0 else -1)
            spec = torch.rand((batch, T, 1025), device="cuda")
            out = torch.empty(batch * h.num_samples(T), dtype=torch.float64, device="cuda")
            h.griffin_lim(spec, _lib.FRAME_MAJOR, [T] * batch, out, seed=1, iters=2, flags=3, out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)
            h.griffin_lim_iterate(20, st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h.griffin_lim_iterate(100, st)
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 100)
        print("T %4d batch %2d: tile kernel %.4f ms/iter, wide %.4f, automatic choice %.4f" % (T, batch, res[0], res[1], res[2]), flush=True)
h.set_option(_lib.OPT_WIDE_MODE, -1)
h.set_generic_iteration(-1)
