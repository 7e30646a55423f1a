"""Small fixed target for ncu: the bench workload's shapes (64 x 1000 frames, default hparams), device-resident,
a few Griffin-Lim iterations.  Run plainly first, then under ncu (see profiles/README.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = 1000
hparams.load()
h = audio._handle()
if os.environ.get('NSB_SYNC_MODE'):
    h.set_option(_lib.OPT_STREAM_SYNC_MODE, int(os.environ['NSB_SYNC_MODE']))
if os.environ.get('NSB_GL_KERNEL'):
    h.set_generic_iteration(int(os.environ['NSB_GL_KERNEL']))      # 0 k_gl_stream, 2 k_gl_iter
spec = torch.rand((batch, T, 1025), device="cuda")
out = torch.empty(batch * h.num_samples(T), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    h.griffin_lim(spec, _lib.FRAME_MAJOR, [T] * batch, out, seed=1, iters=iters,
                  flags=_lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS, out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)
wav = torch.rand(batch * 125000, device="cuda") - 0.5
lin = torch.empty((batch * 501, 1025), device="cuda")
mel = torch.empty((batch * 501, 80), device="cuda")
h.features(wav, [125000] * batch, lin, mel, space=_lib.DEVICE, stream=st)
h.check_status(st)
torch.cuda.synchronize()
print("ok", float(out.abs().max()), h.kernel_launches())
