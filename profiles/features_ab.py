"""Device-resident feature extraction (spectrogram + melspectrogram, BASELINE config 2 sample) with the mel projection as
sparse rows (NSB_OPT_MEL_LINES 0) against line segments / two moments per band on the plain (1) and the skewed (2) magnitude row and over balanced pieces of the segments (3, production), plus the differences."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

hp = hparams.load()
hp.parse("min_level_db=-100")          # the yaml's +100 saturates every feature at 1.0: a vacuous comparison
h = audio._handle()
st = torch.cuda.current_stream().cuda_stream
rs = np.random.RandomState(1234)
durs = np.clip(rs.normal(6.57, 2.19, size=512), 1.0, 10.0)
ns = [int(d * 20000) for d in durs]
t = np.arange(sum(ns)) / 20000.0
wav = (0.3 * np.sin(2 * np.pi * 180 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 0.003 * rs.standard_normal(sum(ns))).astype(np.float32)
Tn = [h.num_frames(n) for n in ns]
d_wav = torch.from_numpy(wav).cuda()
outs = {}
for mode in (3, 2, 1, 0, 3, 2, 3, 2):
    h.set_option(_lib.OPT_MEL_LINES, mode)
    d_lin = torch.empty((sum(Tn), 1025), dtype=torch.float32, device="cuda")
    d_mel = torch.empty((sum(Tn), 80), dtype=torch.float32, device="cuda")
    for _ in range(3):
        h.features(d_wav, ns, d_lin, d_mel, space=_lib.DEVICE, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        h.features(d_wav, ns, d_lin, d_mel, space=_lib.DEVICE, stream=st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    outs[mode] = d_mel.cpu().numpy()
    print("mel_lines %d: %.3f ms for %d frames -> %.1f M mel frames/s" % (mode, ms, sum(Tn), sum(Tn) / ms / 1e3), flush=True)
a, b = outs[0].astype(np.float64), outs[1].astype(np.float64)
assert np.array_equal(outs[1], outs[2])        # the pad words change addresses, not the order of the adds
print("rows vs lines: rel-L2 %.3g, max abs %.3g (normalised dB scale, min_level_db=-100; fraction of values strictly inside (0,1): %.2f)" % (
    np.linalg.norm(a - b) / np.linalg.norm(a), np.abs(a - b).max(), float(((a > 0) & (a < 1)).mean())))
c = outs[3].astype(np.float64)
print("balanced pieces (3) vs whole segments (2): rel-L2 %.3g, max abs %.3g" % (np.linalg.norm(c - b) / np.linalg.norm(b), np.abs(c - b).max()))
h.set_option(_lib.OPT_MEL_LINES, 3)
