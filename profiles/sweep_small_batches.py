"""Griffin-Lim iteration on small and medium batches (BASELINE config 5's low end), device-resident, T = 1000: the automatic
kernel choice against the tile kernel at several tile lengths, its wide mode, and the streaming kernel at several chunkings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

hparams.load()
h = audio._handle()
st = torch.cuda.current_stream().cuda_stream
T = 1000


def ms_per_iter(batch):
    spec = torch.rand((batch, T, 1025), device="cuda")
    out = torch.empty(batch * h.num_samples(T), dtype=torch.float64, device="cuda")
    h.griffin_lim(spec, _lib.FRAME_MAJOR, [T] * batch, out, seed=1, iters=2, flags=3, out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)
    h.griffin_lim_iterate(20, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h.griffin_lim_iterate(120, st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 120


def reset():
    h.set_option(_lib.OPT_WIDE_MODE, -1); h.set_generic_iteration(-1); h.set_tile_hops(0); h.set_stream_grid(0)


for batch in (4, 8, 12, 16, 24, 32, 48):
    reset()
    res = ["auto %.4f" % ms_per_iter(batch)]
    for hops in (8, 12, 16, 20, 28):
        reset(); h.set_generic_iteration(2); h.set_option(_lib.OPT_WIDE_MODE, 0); h.set_tile_hops(hops)
        res.append("tile%d %.4f" % (hops, ms_per_iter(batch)))
    reset(); h.set_generic_iteration(2); h.set_option(_lib.OPT_WIDE_MODE, 1)
    res.append("wide %.4f" % ms_per_iter(batch))
    for chunks in (296, 592, 1184):     # streaming kernel: that many chunks per iteration (2 CTAs per SM = 296 in flight)
        reset(); h.set_generic_iteration(0); h.set_stream_grid(chunks)
        res.append("stream/%d %.4f" % (chunks, ms_per_iter(batch)))
    reset(); h.set_generic_iteration(0)
    res.append("stream-auto %.4f" % ms_per_iter(batch))
    print("batch %2d (%5d frames), ms per iteration: %s" % (batch, batch * T, "  ".join(res)), flush=True)
reset()
