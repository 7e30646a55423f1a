"""Parity tests proper: the sm_100a kernels, called through the C ABI (ctypes) on a real B200, against the
CPU oracle on the same seeded inputs, the committed golden fixtures, and - at BASELINE.json's full sizes -
through size-independent properties.  Tolerances are BASELINE.json's: single STFT / iSTFT / mel call within
1e-5 relative L2; Griffin-Lim waveform SNR >= 40 dB vs the oracle and spectral convergence within 1 %."""
import numpy as np
import pytest

import parity_checks as pc
from conftest import make_hp, speechlike
from nspeech_b200 import _lib, audio, batch, hparams
from oracle import audio_oracle as ao

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _native():
    lib = _lib.default_lib()                       # raises if libnspeech_b200.so is missing: no fallback
    assert lib.path == _lib.DEFAULT_LIB            # the sm_100a build, not the CPU-emulated test library
    assert lib.device_count() >= 1
    yield
    hparams.load()


@pytest.mark.parametrize("over", pc.CONFIGS)
def test_single_ops_vs_oracle(over):
    pc.check_single_ops_vs_oracle(over)


@pytest.mark.parametrize("over", pc.CONFIGS)
def test_griffin_lim_vs_oracle(over):
    pc.check_griffin_lim_vs_oracle(over)


@pytest.mark.parametrize("over", pc.GENERIC_CONFIGS)
def test_generic_num_freq_vs_oracle(over):
    pc.check_single_ops_vs_oracle(over)
    pc.check_griffin_lim_vs_oracle(over)


def test_generic_num_freq_tf_twin_and_stages():
    pc.check_generic_tf_twin_and_stages()


def test_golden_fixtures_through_kernels(golden):
    pc.check_golden_fixtures_through_kernels(golden)


def test_ragged_batch_and_tiles():
    pc.check_ragged_batch_and_tiles()


@pytest.mark.parametrize("over", pc.CONFIGS[:2])
def test_tf_twin_vs_oracle(over):
    pc.check_tf_twin_vs_oracle(over)


@pytest.mark.parametrize("tf", [False, True])
def test_streaming_rounds(tf):
    pc.check_streaming_rounds(T=700, iters=6, tf=tf)


def test_find_endpoint_and_synthesis_stage(golden):
    pc.check_find_endpoint_and_synthesis_stage(golden)


def test_trimming(golden_process):
    pc.check_trimming(golden_process)


def test_feeder_targets():
    pc.check_feeder_targets()


def test_errors_and_edge_cases():
    pc.check_errors_and_edge_cases()


def test_save_wav_scaling_and_int16():
    pc.check_save_wav_scaling_and_int16()


def test_feeder_groups():
    pc.check_feeder_groups()


def test_async_submit_wait():
    pc.check_async_submit_wait()


def test_api_guards():
    pc.check_api_guards()


def test_features_pipeline():
    pc.check_features_pipeline()


def test_mel_projection_modes():
    pc.check_mel_projection_modes()


def test_stale_griffin_lim_state():
    torch = pytest.importorskip("torch")
    pc.check_stale_griffin_lim_state(lambda a: torch.from_numpy(a).cuda(), torch.cuda.current_stream().cuda_stream)


def test_device_random_phase_is_deterministic_per_seed():
    pc.check_device_random_phase_is_deterministic_per_seed()


@pytest.mark.parametrize("mn", [100, -100])
def test_config1_full_griffin_lim(mn):
    """BASELINE config 1: one 5 s spectrogram [1025,401], 60 iterations, identical supplied initial phase."""
    ohp = pc._load(min_level_db=mn)
    wav = speechlike(100000, 11)
    S = ao.spectrogram(wav, ohp)
    assert S.shape == (1025, 401)
    ang = np.exp(2j * np.pi * np.random.default_rng(0).random((1025, 401)))
    y = audio.inv_spectrogram(S, init_phase=ang)            # griffin_lim_iters = 60 from hparams
    yref = ao.inv_spectrogram(S, ohp, angles=ang)
    assert y.shape == yref.shape == (100000,)
    assert ao.snr_db(y, yref) >= 40.0
    mag = ao._db_to_amp(ao._denormalize(S, ohp) + ohp.ref_level_db) ** ohp.power
    sc = ao.spectral_convergence(ao.preemphasis(y, ohp), mag, ohp)
    sc_ref = ao.spectral_convergence(ao.preemphasis(yref, ohp), mag, ohp)
    assert abs(sc - sc_ref) <= 0.01 * max(sc_ref, 1e-12) + 1e-6


def test_random_init_tacotron_like_input():
    """U(0,1) 'spectrogram' (what a random-init Tacotron emits, BASELINE config 3 variant ii), 60 iterations."""
    ohp = pc._load()
    rs = np.random.RandomState(5)
    S = rs.rand(1025, 120).astype(np.float32)
    ang = np.exp(2j * np.pi * rs.rand(1025, 120))
    y = audio.inv_spectrogram(S, init_phase=ang)
    assert ao.snr_db(y, ao.inv_spectrogram(S, ohp, angles=ang)) >= 40.0


def test_full_size_properties():
    """BASELINE config 3 size (64 x 1000 frames): properties that need no CPU oracle run."""
    pc._load(min_level_db=-100)
    h = audio._handle()
    rs = np.random.RandomState(9)
    N, T = 64, 1000
    wavs = [speechlike(h.hop * (T - 1), 100 + i) for i in range(4)]
    # (1) STFT -> iSTFT round trip at full length
    for w in wavs[:2]:
        D = audio._stft(w)
        assert D.shape == (1025, T)
        assert ao.rel_l2(audio._istft(D), w) < 1e-5
    # (2) linearity of the STFT
    a, b = wavs[0], wavs[1]
    assert ao.rel_l2(audio._stft(a + 2 * b), audio._stft(a) + 2 * audio._stft(b)) < 1e-5
    # (3) Parseval on the iSTFT/STFT pair: energy of the re-analysed resynthesis equals the original
    assert abs(np.sum(audio._istft(audio._stft(a)).astype(np.float64) ** 2) / np.sum(a.astype(np.float64) ** 2) - 1) < 1e-5
    # (4) a batch gives bit-identical waveforms to one-at-a-time calls (same supplied phase), any tile size
    feats = [audio.spectrogram(w) for w in wavs]
    specs = np.stack([feats[i % 4].T for i in range(N)])              # [N,T,F], Tacotron layout
    phase = np.exp(2j * np.pi * rs.rand(N, T, 1025)).astype(np.complex64)
    outs = batch.inv_spectrogram_batch(specs, init_phase=phase, iters=8)
    assert len(outs) == N and all(o.shape == (h.hop * (T - 1),) and o.dtype == np.float64 for o in outs)
    for i in (0, 5, 63):
        single = audio.inv_spectrogram(specs[i].T, init_phase=phase[i].T, iters=8)
        np.testing.assert_array_equal(outs[i], single)
    # (5) Griffin-Lim reduces the spectral inconsistency monotonically-ish: more iterations -> lower error
    ohp = make_hp(min_level_db=-100)
    mag = ao._db_to_amp(ao._denormalize(feats[0], ohp) + ohp.ref_level_db) ** ohp.power
    errs = []
    for it in (0, 10, 60):
        y = audio.inv_spectrogram(feats[0], init_phase=phase[0].T, iters=it)
        D = np.abs(audio._stft(audio.preemphasis(y).astype(np.float32)))
        errs.append(np.linalg.norm(D - mag) / np.linalg.norm(mag))
    assert errs[2] < errs[1] < errs[0]


def test_features_full_size_properties():
    """BASELINE config 2 at its full size (13,100 LJSpeech-shaped clips, 6.8 M frames, 37 GB of device buffers) through ONE
    feature pass on device arrays: sampled clips against the oracle and against one-at-a-time calls (identical bits), every
    value finite and inside [0, 1], and the whole result identical under another launch of the same pass."""
    torch = pytest.importorskip("torch")
    import bench
    ohp = pc._load(min_level_db=-100)
    h = audio._handle()
    dev = torch.device("cuda:0")
    free, _ = torch.cuda.mem_get_info()
    n_clips = 13100 if free > (60 << 30) else 1310
    d_wav, ns = bench.speechlike_corpus(torch, dev, n_clips)
    Ts = [h.num_frames(n) for n in ns]
    tot = sum(Ts)
    d_lin = torch.empty((tot, 1025), dtype=torch.float32, device=dev)
    d_mel = torch.empty((tot, 80), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    h.features(d_wav, ns, d_lin, d_mel, space=_lib.DEVICE, stream=st)
    h.check_status(st)
    assert tot == sum(1 + n // h.hop for n in ns)
    assert bool(torch.isfinite(d_mel).all()) and float(d_mel.min()) >= 0.0 and float(d_mel.max()) <= 1.0
    for lo in range(0, tot, 1 << 20):                                   # (the linear feature in slices: no second 28 GB temporary)
        blk = d_lin[lo:lo + (1 << 20)]
        assert bool(torch.isfinite(blk).all()) and float(blk.min()) >= 0.0 and float(blk.max()) <= 1.0
    assert 0.0 < float(d_mel.mean()) < 1.0                             # min_level_db = -100: not saturated
    soff = np.concatenate([[0], np.cumsum(ns)])
    foff = np.concatenate([[0], np.cumsum(Ts)])
    for i in (0, 1, n_clips // 3, n_clips // 2 + 1, n_clips - 1):
        w = d_wav[soff[i]:soff[i + 1]].cpu().numpy()
        lin = d_lin[foff[i]:foff[i + 1]].cpu().numpy().T
        mel = d_mel[foff[i]:foff[i + 1]].cpu().numpy().T
        assert ao.rel_l2(lin, ao.spectrogram(w, ohp)) < 1e-5 and ao.rel_l2(mel, ao.melspectrogram(w, ohp)) < 1e-5
        l1, m1 = audio.spectrogram_and_mel(w)                          # the clip alone, from host memory
        np.testing.assert_array_equal(lin, l1)
        np.testing.assert_array_equal(mel, m1)
    ck = (d_mel.double().sum().item(), d_lin[: 1 << 22].double().sum().item())
    h.features(d_wav, ns, d_lin, d_mel, space=_lib.DEVICE, stream=st)
    h.check_status(st)
    assert ck == (d_mel.double().sum().item(), d_lin[: 1 << 22].double().sum().item())


def test_largest_sweep_batch_matches_single_calls():
    """BASELINE config 5's largest point (1024 utterances x 1000 frames = 1.02 M frames in one call, device-resident, supplied
    phase): sampled utterances equal one-at-a-time calls bit for bit - the streaming kernel's chunking, its item counters and
    the 32-bit frame / 64-bit sample offsets at the largest size the benchmark uses."""
    torch = pytest.importorskip("torch")
    pc._load(min_level_db=-100)
    h = audio._handle()
    free, _ = torch.cuda.mem_get_info()
    N, T = (1024 if free > (60 << 30) else 128), 1000
    g = torch.Generator(device="cuda").manual_seed(5)
    spec = torch.rand((N * T, 1025), device="cuda", generator=g)
    ang = torch.rand((N * T, 1025), device="cuda", generator=g) * (2 * np.pi)
    phase = torch.polar(torch.ones_like(ang), ang)                      # complex64 [frames, 1025]
    del ang
    ns = h.num_samples(T)
    out = torch.empty(N * ns, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    h.griffin_lim(spec, _lib.FRAME_MAJOR, [T] * N, out, init_phase=phase, iters=3, flags=3, out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)
    h.check_status(st)
    assert bool(torch.isfinite(out).all())
    for i in (0, 1, N // 2 + 3, N - 1):
        one = torch.empty(ns, dtype=torch.float64, device="cuda")
        h.griffin_lim(spec[i * T:(i + 1) * T], _lib.FRAME_MAJOR, [T], one, init_phase=phase[i * T:(i + 1) * T], iters=3, flags=3,
                      out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)
        h.check_status(st)
        assert torch.equal(one, out[i * ns:(i + 1) * ns]), i


def test_device_pointer_api_with_torch():
    """NSB_DEVICE entry points on torch-owned memory and torch's current stream."""
    torch = pytest.importorskip("torch")
    ohp = pc._load(min_level_db=-100)
    h = audio._handle()
    rs = np.random.RandomState(2)
    Ts = [50, 31]
    specs = [rs.rand(T, 1025).astype(np.float32) for T in Ts]
    phases = [np.exp(2j * np.pi * rs.rand(T, 1025)).astype(np.complex64) for T in Ts]
    d_spec = torch.from_numpy(np.concatenate(specs)).cuda()
    d_ph = torch.view_as_real(torch.from_numpy(np.concatenate(phases))).contiguous().cuda()
    d_out = torch.empty(sum(h.num_samples(T) for T in Ts), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    h.griffin_lim(d_spec, _lib.FRAME_MAJOR, Ts, d_out, init_phase=d_ph, iters=5,
                  flags=_lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS, out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)
    h.check_status(st)
    out = d_out.cpu().numpy()
    off = 0
    for s, p, T in zip(specs, phases, Ts):
        n = h.num_samples(T)
        assert ao.snr_db(out[off:off + n], ao.inv_spectrogram(s.T, ohp, angles=p.T, iters=5)) > 60
        off += n
    # device-resident features
    wav = speechlike(20000, 4)
    d_wav = torch.from_numpy(wav).cuda()
    T = h.num_frames(wav.size)
    d_lin = torch.empty((T, 1025), dtype=torch.float32, device="cuda")
    d_mel = torch.empty((T, 80), dtype=torch.float32, device="cuda")
    h.features(d_wav, [wav.size], d_lin, d_mel, space=_lib.DEVICE, stream=st)
    h.check_status(st)
    assert ao.rel_l2(d_lin.cpu().numpy().T, ao.spectrogram(wav, ohp)) < 1e-5
    assert ao.rel_l2(d_mel.cpu().numpy().T, ao.melspectrogram(wav, ohp)) < 1e-5


def test_threads_share_nothing():
    """Feeder-style concurrency (datasets/datafeeder.py:110-116): N python threads call the module API at once."""
    import threading
    ohp = pc._load(min_level_db=-100)
    wavs = [speechlike(8000 + 500 * i, i) for i in range(6)]
    refs = [ao.melspectrogram(w, ohp) for w in wavs]
    errs = []

    def work(i):
        try:
            for _ in range(3):
                m = audio.melspectrogram(wavs[i])
                assert ao.rel_l2(m, refs[i]) < 1e-5
        except Exception as e:
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(wavs))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


def test_fused_iterations_match_one_launch_per_iteration():
    """All Griffin-Lim iterations of a call run in ONE launch: (iteration, tile | chunk) items from a global counter, an
    item starts when its three neighbours of the previous iteration are stored.  The result must not differ by a bit
    from one launch per iteration, nor between the tile kernel (k_gl_iter) and the streaming kernel (k_gl_stream) - at
    config-3 size, where 296 CTAs race through the batch."""
    torch = pytest.importorskip("torch")
    pc._load()
    h = audio._handle()
    st = torch.cuda.current_stream().cuda_stream
    try:
        for Ts, iters in (([1000] * 64, 60), ([2, 3, 700, 41, 1500, 29, 5] * 9, 25)):
            g = torch.Generator(device="cuda").manual_seed(5)
            spec = torch.rand((sum(Ts), 1025), device="cuda", generator=g)
            outs = []
            for kernel, mode in ((2, 1), (2, 0), (0, 1), (0, 0), (0, 1)):   # 1: all iterations in one launch, 0: one launch per iteration
                h.set_generic_iteration(kernel)
                h.set_option(_lib.OPT_FUSE_ITERATIONS, mode)
                out = torch.empty(sum(h.num_samples(t) for t in Ts), dtype=torch.float64, device="cuda")
                h.griffin_lim(spec, _lib.FRAME_MAJOR, Ts, out, seed=3, iters=iters, flags=_lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS,
                              out_dtype=_lib.F64, space=_lib.DEVICE, stream=st)
                h.check_status(st)
                outs.append(out.cpu().numpy())
            assert np.isfinite(outs[0]).all()
            for o in outs[1:]:
                np.testing.assert_array_equal(outs[0], o)
    finally:
        h.set_option(_lib.OPT_FUSE_ITERATIONS, 1)
        h.set_generic_iteration(-1)


def test_host_pipelines_match_one_launch():
    """NSB_HOST Griffin-Lim never runs the batch as one launch: long batches follow the wave schedule (chunk g joins at wave g,
    every launch runs all chunks in flight for a few iterations), others a chunk pipeline over two compute streams (a chunk's
    tail overlaps the next chunk's start, two persistent launches share the SMs).  Whatever the cuts, the waves and the
    streams, the waveforms must not differ by a bit from the unchunked call - ragged batch, even and odd iteration counts,
    both iteration kernels in the mix."""
    pc._load()
    h = audio._handle()
    rs = np.random.RandomState(11)
    Ts = [300, 17, 1000, 2, 640, 1500, 77, 256] * 12              # 45.5k frames: above the wave schedule's threshold
    specs = [rs.rand(T, 1025).astype(np.float32) for T in Ts]
    try:
        for iters in (12, 13):
            outs = []
            for chunks, overlap, wave in ((1, 0, 0), (0, 1, 1), (0, 1, 0), (5, 1, 1), (5, 0, 1), (13, 1, 0), (2, 1, 0)):
                h.set_host_chunks(chunks)
                h.set_option(_lib.OPT_OVERLAP_CHUNKS, overlap)
                h.set_option(_lib.OPT_WAVE_SCHEDULE, wave)
                for rep in range(2):                              # twice: the second call reuses every workspace
                    outs.append(np.concatenate(batch.inv_spectrogram_batch(specs, seed=7, iters=iters)).copy())
            assert np.isfinite(outs[0]).all() and np.abs(outs[0]).max() > 0
            for o in outs[1:]:
                np.testing.assert_array_equal(outs[0], o)
        # the synthesis stage (TF-twin Griffin-Lim + de-emphasis + endpoint search, synthesizer.py:30, 51-53) at eval.py's
        # batch size rides the same schedules
        lin = rs.rand(32, 1500, 1025).astype(np.float32)
        lin[:, 900:, :] = 1.0                                      # quiet tail (the yaml's +100 dB floor inverts the scale): endpoints differ from the length
        res = []
        for chunks, wave in ((1, 0), (0, 1), (0, 0), (3, 1)):
            h.set_host_chunks(chunks)
            h.set_option(_lib.OPT_WAVE_SCHEDULE, wave)
            res.append(audio.synthesize_waveforms(lin, iters=20))
        for r in res[1:]:
            assert [len(w) for w in r] == [len(w) for w in res[0]]
            for a, b in zip(res[0], r):
                np.testing.assert_array_equal(a, b)
    finally:
        h.set_host_chunks(0)
        h.set_option(_lib.OPT_OVERLAP_CHUNKS, 0)
        h.set_option(_lib.OPT_WAVE_SCHEDULE, 1)


def test_feature_host_pipeline_matches_one_chunk():
    """Host-side feature calls are cut into chunks of whole utterances (copy-in, transform and copy-out of consecutive chunks
    overlap).  Every cut must give the bits of the single-chunk call: packed features, the STFT, and the feeder's padded
    batch tensors - ragged clip lengths."""
    pc._load(min_level_db=-100)
    h = audio._handle()
    rs = np.random.RandomState(21)
    lens = [int(n) for n in rs.randint(20000, 200000, size=48)] + [250, 999, 300]
    wavs = [speechlike(n, 300 + i) for i, n in enumerate(lens)]
    try:
        ref = None
        for chunks in (1, 0, 7, 51):
            h.set_host_chunks(chunks)
            feats = batch.features_batch(wavs)
            mel_t, lin_t, Ts = batch.feeder_targets(wavs, 5)
            D = audio._stft(np.concatenate(wavs[:6]))
            cur = ([f[0].copy() for f in feats], [f[1].copy() for f in feats], mel_t.copy(), lin_t.copy(), list(Ts), D.copy())
            if ref is None:
                ref = cur
                assert all(np.isfinite(a).all() for a in ref[0]) and 0 < float(ref[0][0].mean()) < 1
                continue
            for a, b in zip(ref[0] + ref[1] + [ref[2], ref[3], ref[5]], cur[0] + cur[1] + [cur[2], cur[3], cur[5]]):
                np.testing.assert_array_equal(a, b)
            assert ref[4] == cur[4]
    finally:
        h.set_host_chunks(0)


class _CaiOnly(object):
    """a device array of 'some other framework': exposes nothing but the CUDA array interface (keeps the torch tensor alive)"""

    def __init__(self, t):
        self._t = t
        self.__cuda_array_interface__ = t.__cuda_array_interface__


class _DlpackOnly(object):
    def __init__(self, t):
        self._t = t

    def __dlpack__(self, stream=None):
        return self._t.__dlpack__()

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


def test_device_arrays_at_the_python_boundary():
    """torch CUDA tensors, CUDA-array-interface and DLPack producers go in without a host round trip and device arrays come
    back (north-star: 'numpy/DLPack buffers'); the results equal the host path bit for bit."""
    torch = pytest.importorskip("torch")
    from nspeech_b200 import _buffers
    pc._load(min_level_db=-100)
    h = audio._handle()
    rs = np.random.RandomState(8)
    wav = speechlike(9000, 3)
    lin, mel = audio.spectrogram_and_mel(wav)
    D = audio._stft(wav)
    y = audio._istft(D)
    S = rs.rand(1025, 23).astype(np.float32)
    ang = np.exp(2j * np.pi * rs.rand(1025, 23)).astype(np.complex64)
    g = audio.inv_spectrogram(S, init_phase=ang, iters=4)
    twin = audio.inv_spectrogram_tensorflow(S.T.copy(), iters=3)
    tw = torch.from_numpy(wav).cuda()
    # torch in -> torch out
    tl, tm = audio.spectrogram_and_mel(tw)
    assert isinstance(tl, torch.Tensor) and tl.is_cuda and tl.shape == lin.shape and tl.dtype == torch.float32
    np.testing.assert_array_equal(tl.cpu().numpy(), lin)
    np.testing.assert_array_equal(tm.cpu().numpy(), mel)
    np.testing.assert_array_equal(audio.spectrogram(tw).cpu().numpy(), lin)
    np.testing.assert_array_equal(audio.melspectrogram(tw).cpu().numpy(), mel)
    tD = audio._stft(tw)
    assert tD.dtype == torch.complex64
    np.testing.assert_array_equal(tD.cpu().numpy(), D)
    np.testing.assert_array_equal(audio._istft(tD).cpu().numpy(), y)                     # Fortran-ordered [F,T] view: frame-major
    np.testing.assert_array_equal(audio._istft(tD.contiguous()).cpu().numpy(), y)        # C-ordered: bin-major
    tS, ta = torch.from_numpy(S).cuda(), torch.from_numpy(ang).cuda()
    tg = audio.inv_spectrogram(tS, init_phase=ta, iters=4)
    assert tg.dtype == torch.float64 and tg.is_cuda
    np.testing.assert_array_equal(tg.cpu().numpy(), g)
    np.testing.assert_array_equal(audio.inv_spectrogram(tS.T.contiguous().T, init_phase=ta.T.contiguous().T, iters=4).cpu().numpy(), g)
    np.testing.assert_array_equal(audio.inv_spectrogram_tensorflow(tS.T.contiguous(), iters=3).cpu().numpy(), twin)
    np.testing.assert_array_equal(audio.inv_preemphasis(tw).cpu().numpy(), audio.inv_preemphasis(wav))
    # float64 torch input is converted on the device
    np.testing.assert_array_equal(audio.spectrogram(tw.double()).cpu().numpy(), lin)
    # another framework's arrays (CUDA array interface / DLPack only) -> DeviceArray, adoptable without a copy
    for wrap in (_CaiOnly, _DlpackOnly):
        r = audio.inv_spectrogram(wrap(tS.T.contiguous().T), init_phase=wrap(ta.T.contiguous().T), iters=4)
        assert isinstance(r, _buffers.DeviceArray) and r.shape == g.shape and r.dtype == np.float64
        np.testing.assert_array_equal(r.copy_to_host(), g)
        adopted = torch.from_dlpack(r)
        assert adopted.is_cuda and adopted.data_ptr() == r.ptr
        np.testing.assert_array_equal(adopted.cpu().numpy(), g)
        np.testing.assert_array_equal(torch.as_tensor(r, device="cuda").cpu().numpy(), g)          # __cuda_array_interface__
        rl, rm = audio.spectrogram_and_mel(wrap(tw))
        assert rl.shape == lin.shape and rm.shape == mel.shape
        np.testing.assert_array_equal(np.asarray(rl), lin)
        np.testing.assert_array_equal(torch.from_dlpack(rm).cpu().numpy(), mel)
        del adopted, r, rl, rm
    # Tacotron's [N, T, F] batch on the GPU (models/tacotron.py:98, 107)
    B = rs.rand(3, 31, 1025).astype(np.float32)
    ref = batch.inv_spectrogram_batch(B, seed=9, iters=3)
    got = batch.inv_spectrogram_batch(torch.from_numpy(B).cuda(), seed=9, iters=3)
    assert got.shape == (3, h.num_samples(31)) and got.is_cuda
    for i in range(3):
        np.testing.assert_array_equal(got[i].cpu().numpy(), ref[i])
    with pytest.raises(audio.ParameterError):
        audio.spectrogram(torch.full((3000,), float("nan"), device="cuda"))
    with pytest.raises(TypeError):
        audio.inv_spectrogram(tS, init_phase=ang, iters=1)              # phase on the host, spectrogram on the device


def test_two_streams_on_one_handle_are_ordered():
    """NSB_DEVICE calls on different streams share the handle's descriptors and counters - the second call must
    wait for the first on the device.  Results equal the single-stream run; nothing hangs."""
    torch = pytest.importorskip("torch")
    pc._load()
    h = audio._handle()
    g = torch.Generator(device="cuda").manual_seed(2)
    Ts = [400] * 40
    spec = torch.rand((sum(Ts), 1025), device="cuda", generator=g)
    flags = _lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS
    n = sum(h.num_samples(t) for t in Ts)
    ref = torch.empty(n, dtype=torch.float64, device="cuda")
    s0 = torch.cuda.current_stream().cuda_stream
    h.griffin_lim(spec, _lib.FRAME_MAJOR, Ts, ref, seed=1, iters=10, flags=flags, out_dtype=_lib.F64, space=_lib.DEVICE, stream=s0)
    h.check_status(s0)
    a, b = torch.cuda.Stream(), torch.cuda.Stream()
    o1, o2 = torch.empty_like(ref), torch.empty_like(ref)
    torch.cuda.synchronize()
    h.griffin_lim(spec, _lib.FRAME_MAJOR, Ts, o1, seed=1, iters=10, flags=flags, out_dtype=_lib.F64, space=_lib.DEVICE, stream=a.cuda_stream)
    h.griffin_lim(spec, _lib.FRAME_MAJOR, Ts, o2, seed=1, iters=10, flags=flags, out_dtype=_lib.F64, space=_lib.DEVICE, stream=b.cuda_stream)
    host = np.empty(h.num_samples(50), np.float64)                      # and a synchronous HOST call right behind them
    h.griffin_lim(np.random.RandomState(1).rand(50, 1025).astype(np.float32), _lib.FRAME_MAJOR, [50], host, seed=1, iters=2, flags=flags, out_dtype=_lib.F64)
    torch.cuda.synchronize()
    assert torch.equal(o1, ref) and torch.equal(o2, ref) and np.isfinite(host).all()


def test_tf_twin_golden(golden_tf):
    pc.check_tf_twin_golden(golden_tf)


def test_length_and_hparams_property_sweep():
    """hypothesis over clip length x sample rate x window length x num_freq"""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import strategies as hs

    @hyp.settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(n=hs.one_of(hs.integers(1, 6144), hs.sampled_from([1, 2, 249, 250, 251, 1023, 1024, 1025, 2047, 2048, 2049, 4096, 6144])),
               sr=hs.sampled_from([16000, 20000, 22050, 24000]), fl=hs.sampled_from([25, 40, 50, 64]), nf=hs.sampled_from([1025, 1025, 513, 2049]))
    def run(n, sr, fl, nf):
        pc.check_length_and_hparams_sweep([(n, sr, fl, nf)])
    run()


def test_config3_shape_through_the_streaming_kernel_vs_oracle():
    """Two 12.5 s utterances (1000 frames, U(0,1) spectrograms, the yaml's +100 dB floor), all 60 iterations through the
    production kernel k_gl_stream itself (forced: the automatic choice would take the tile kernel for two utterances),
    against the oracle: SNR >= 40 dB and spectral convergence within 1 % (BASELINE.json) - no inference through the
    bit-equality of the kernels."""
    ohp = pc._load()
    h = audio._handle()
    rs = np.random.RandomState(77)
    specs = [rs.rand(1025, 1000).astype(np.float32) for _ in range(2)]
    phases = [np.exp(2j * np.pi * rs.rand(1025, 1000)) for _ in range(2)]
    h.set_generic_iteration(0)
    try:
        launches = h.kernel_launches()
        outs = batch.inv_spectrogram_batch(specs, init_phase=phases, layout="FT")
        assert h.kernel_launches() - launches <= 6          # prepare, initial iSTFT, ONE iteration launch, de-emphasis
    finally:
        h.set_generic_iteration(-1)
    for S, ang, y in zip(specs, phases, outs):
        yref = ao.inv_spectrogram(S, ohp, angles=ang)
        assert y.shape == yref.shape == (250 * 999,)
        assert ao.snr_db(y, yref) >= 40.0, ao.snr_db(y, yref)
        mag = ao._db_to_amp(ao._denormalize(S, ohp) + ohp.ref_level_db) ** ohp.power
        sc, sc_ref = (ao.spectral_convergence(ao.preemphasis(v, ohp), mag, ohp) for v in (y, yref))
        assert abs(sc - sc_ref) <= 0.01 * sc_ref
