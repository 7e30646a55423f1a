"""k_gl_stream, barrier after every colour step (sync mode 2, production) against the split-phase form (mode 10: a warp
publishes its finished overlap-add and waits for the others only before its next one).  BASELINE config 3 shape,
device-resident, 60 iterations.  Usage: python profiles/sync_ab.py [modes...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

modes = [int(a) for a in sys.argv[1:]] or [2, 10, 2, 10]
hparams.load()
h = audio._handle()
st = torch.cuda.current_stream().cuda_stream
N, T = 64, 1000
spec = torch.rand((N * T, 1025), device="cuda")
out = torch.empty(N * h.num_samples(T), dtype=torch.float64, device="cuda")
ref = None
for mode in modes:
    h.set_option(_lib.OPT_STREAM_SYNC_MODE, mode)
    fn = lambda: h.griffin_lim(spec, _lib.FRAME_MAJOR, [T] * N, out, seed=1, iters=60, flags=3, out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)  # noqa: E731
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    h.check_status(st)
    same = "" if ref is None else ("  identical to the first mode's waveform: %s" % bool(torch.equal(out, ref)))
    if ref is None:
        ref = out.clone()
    print("sync mode %2d: %.3f ms per step (60 iterations, %d frames)%s" % (mode, e0.elapsed_time(e1) / 5, N * T, same))
h.set_option(_lib.OPT_STREAM_SYNC_MODE, 2)
