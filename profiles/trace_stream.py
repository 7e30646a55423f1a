"""Per-CTA timeline of one k_gl_stream launch (SM id, start, end): which SMs finish late, and is it systematic?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

hparams.load()
h = audio._handle()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sync = int(sys.argv[2]) if len(sys.argv) > 2 else 1
T = 1000
st = torch.cuda.current_stream().cuda_stream
spec = torch.rand((batch, T, 1025), device="cuda")
out = torch.empty(batch * h.num_samples(T), dtype=torch.float64, device="cuda")
h.set_option(_lib.OPT_STREAM_SYNC_MODE, sync)
h.griffin_lim(spec, _lib.FRAME_MAJOR, [T] * batch, out, seed=1, iters=2, flags=3, out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)
h.griffin_lim_iterate(10, st)
h.stream_trace(True)
runs = []
for rep in range(4):
    h.griffin_lim_iterate(1, st)
    tr = h.stream_trace(True).astype(np.int64)
    t0 = tr[:, 1].min()
    dur = (tr[:, 2] - tr[:, 1]) / 1e3
    end = (tr[:, 2] - t0) / 1e3
    runs.append((tr[:, 0].copy(), dur, end))
    print("rep %d: ctas %d  kernel span %.1f us  cta duration us: min %.1f  mean %.1f  max %.1f   start skew max %.1f us" % (
        rep, len(tr), end.max(), dur.min(), dur.mean(), dur.max(), ((tr[:, 1] - t0) / 1e3).max()))
sm, dur, end = runs[-1]
per_sm = {}
for s_, d_ in zip(sm, dur):
    per_sm.setdefault(int(s_), []).append(d_)
print("CTAs per SM histogram:", np.bincount([len(v) for v in per_sm.values()]))
order = sorted(per_sm, key=lambda k: -max(per_sm[k]))
print("slowest SMs:", [(k, round(max(per_sm[k]), 1)) for k in order[:16]])
print("fastest SMs:", [(k, round(max(per_sm[k]), 1)) for k in order[-16:]])
# is it systematic?  correlation of per-SM duration between two launches
a = {}
for s_, d_ in zip(runs[-2][0], runs[-2][1]):
    a.setdefault(int(s_), []).append(d_)
common = sorted(set(a) & set(per_sm))
x = np.array([max(a[k]) for k in common]); y = np.array([max(per_sm[k]) for k in common])
print("correlation of per-SM duration between consecutive launches: %.3f" % np.corrcoef(x, y)[0, 1])
print("per-SM duration by SM id (us):", " ".join("%d:%.0f" % (k, max(per_sm[k])) for k in sorted(per_sm)))
h.stream_trace(False)
h.set_option(_lib.OPT_STREAM_SYNC_MODE, 2)
sm, dur, end = runs[-1]
print("duration by blockIdx (us):", " ".join("%d:%.0f/%d" % (i, dur[i], sm[i]) for i in range(len(dur))))
