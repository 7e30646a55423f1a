"""ncu target for the generic-size kernels: one Griffin-Lim call (ITERS iterations, 16 x 1000 frames, device-resident) and one
feature pass at the given hparams.  Usage: python profiles/generic_profile_target.py "num_freq=513,sample_rate=16000,frame_length_ms=50" [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

over = sys.argv[1] if len(sys.argv) > 1 else "num_freq=513,sample_rate=16000,frame_length_ms=50"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
hp = hparams.load()
hp.parse(over)
h = audio._handle()
st = torch.cuda.current_stream().cuda_stream
N, T, F = 16, 1000, h.num_freq
spec = torch.rand((N * T, F), device="cuda")
out = torch.empty(N * h.num_samples(T), dtype=torch.float64, device="cuda")
for _ in range(2):
    h.griffin_lim(spec, _lib.FRAME_MAJOR, [T] * N, out, seed=1, iters=iters, flags=3, out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)
n = h.hop * 4000
wav = torch.rand(N * n, device="cuda") - 0.5
Tn = h.num_frames(n)
lin = torch.empty((N * Tn, F), device="cuda")
mel = torch.empty((N * Tn, 80), device="cuda")
for _ in range(2):
    h.features(wav, [n] * N, lin, mel, space=_lib.DEVICE, stream=st)
torch.cuda.synchronize()
h.check_status(st)
