"""Where the time of BASELINE config 4 goes: the synthesis stage `audio.synthesize_waveforms` on [32, 1500, 1025]
(TF-twin Griffin-Lim, 60 iterations, + de-emphasis + find_endpoint + save_wav scaling), device-resident against
host calls, the iteration kernels with the default hparams' geometry as immediates against the general
instantiations (NSB_OPT_SPECIALIZE), and the librosa-geometry Griffin-Lim on the same shape for scale.
Usage: python profiles/config4_breakdown.py [batch] [frames]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from nspeech_b200 import _lib, audio, hparams  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
hparams.load()
h = audio._handle()
st = torch.cuda.current_stream().cuda_stream
gen = torch.Generator().manual_seed(4)


def ev_ms(fn, reps):
    fn()
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def wall_ms(fn, reps):
    fn()
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


lin = _lib.PinnedArray((N, T, 1025), np.float32)
lin.array[...] = torch.rand((N, T, 1025), generator=gen, dtype=torch.float32).numpy()
d_lin = torch.from_numpy(lin.array).cuda()
pageable = np.array(lin.array)
frame_its = N * T * 60

for spec in (1, 0):
    h.set_option(_lib.OPT_SPECIALIZE, spec)
    tag = "immediates" if spec else "general   "
    # the TF-twin iterations alone (device arrays in, float32 waveform out, no de-emphasis / endpoint)
    d_out = torch.empty(N * h.num_samples_tf(T), dtype=torch.float32, device="cuda")
    ms = ev_ms(lambda: h.griffin_lim(d_lin.view(N * T, 1025), _lib.FRAME_MAJOR, [T] * N, d_out, iters=60,
                                     flags=_lib.GL_DENORMALIZE | _lib.GL_TF_TWIN, out_dtype=_lib.F32, space=_lib.DEVICE, stream=st), 5)
    print("[%s] TF-twin Griffin-Lim, device-resident: %.2f ms (%.0f M frame-iterations/s)" % (tag, ms, frame_its / ms / 1e3))
    d_wav = torch.empty(N * h.num_samples_tf(T), dtype=torch.int16, device="cuda")
    d_end = torch.empty(N, dtype=torch.int64, device="cuda")
    ms = ev_ms(lambda: h.synthesize(d_lin.view(N * T, 1025), [T] * N, d_wav, d_end, flags=_lib.SYNTH_PEAK_NORMALIZE, out_dtype=_lib.I16,
                                    space=_lib.DEVICE, stream=st), 5)
    print("[%s] nsb_synthesize_ex, device arrays (int16 out): %.2f ms" % (tag, ms))
    ms = wall_ms(lambda: audio.synthesize_waveforms(lin.array, peak_normalize=True, dtype=np.int16), 8)
    print("[%s] synthesize_waveforms, page-locked host array (int16 out): %.2f ms" % (tag, ms))
    ms = wall_ms(lambda: audio.synthesize_waveforms(lin.array), 8)
    print("[%s] synthesize_waveforms, page-locked host array (float64 out): %.2f ms" % (tag, ms))
    ms = wall_ms(lambda: audio.synthesize_waveforms(pageable, peak_normalize=True, dtype=np.int16), 5)
    print("[%s] synthesize_waveforms, pageable numpy array (int16 out): %.2f ms" % (tag, ms))
    # the librosa-geometry iterations on the same shape
    d_out64 = torch.empty(N * h.num_samples(T), dtype=torch.float64, device="cuda")
    ms = ev_ms(lambda: h.griffin_lim(d_lin.view(N * T, 1025), _lib.FRAME_MAJOR, [T] * N, d_out64, seed=1, iters=60,
                                     flags=3, out_dtype=_lib.F64, space=_lib.DEVICE, stream=st), 5)
    print("[%s] librosa-geometry Griffin-Lim, device-resident: %.2f ms (%.0f M frame-iterations/s)" % (tag, ms, frame_its / ms / 1e3))
    h.check_status(st)
h.set_option(_lib.OPT_SPECIALIZE, 1)
# host calls with the chunk pipeline instead of the wave schedule
for waves in (1, 0):
    h.set_option(_lib.OPT_WAVE_SCHEDULE, waves)
    ms = wall_ms(lambda: audio.synthesize_waveforms(lin.array, peak_normalize=True, dtype=np.int16), 8)
    print("wave schedule %d: synthesize_waveforms, page-locked host array (int16 out): %.2f ms" % (waves, ms))
h.set_option(_lib.OPT_WAVE_SCHEDULE, 1)
for n in (1, 2, 3, 4):
    h.set_host_chunks(n)
    ms = wall_ms(lambda: audio.synthesize_waveforms(lin.array, peak_normalize=True, dtype=np.int16), 8)
    print("host_chunks %d: synthesize_waveforms, page-locked host array (int16 out): %.2f ms" % (n, ms))
h.set_host_chunks(0)
