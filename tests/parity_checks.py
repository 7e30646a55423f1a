"""Parity checks shared by the CPU-emulated run (tests/test_emulated_kernels.py) and the real-GPU run
(tests/test_gpu_parity.py).  Each function drives the product's Python mirror (nspeech_b200.audio) - which
binds whichever native library the caller selected - and compares with the oracle on the same inputs."""
import numpy as np
import pytest

from conftest import make_hp, speechlike, trim_signals
from nspeech_b200 import _lib, audio, hparams
from oracle import audio_oracle as ao
from oracle import tf_signal17 as tfo

CONFIGS = [{"min_level_db": -100}, {}, {"min_level_db": -100, "sample_rate": 22050}]
# hparams-driven num_freq (audio.py:126-130): every n_fft != 2048 runs the generic-size kernels (csrc/gen_kernels.cuh) - powers of
# two, the yaml's commented alternative num_freq 2048 (n_fft 4094 = 2 * 23 * 89, audio.yaml:7) with its 24 kHz (audio.yaml:8),
# a 2^5 * 5^2 size, and a window as long as the transform
GENERIC_CONFIGS = [{"min_level_db": -100, "num_freq": 513}, {"min_level_db": -100, "num_freq": 2049},
                   {"num_freq": 2048, "sample_rate": 24000}, {"min_level_db": -100, "num_freq": 401, "sample_rate": 16000},
                   {"min_level_db": -100, "num_freq": 501, "frame_length_ms": 50, "sample_rate": 20000}]


def _load(**over):
    hp = hparams.load()
    if over:
        hp.parse(",".join("%s=%s" % kv for kv in over.items()))
    return make_hp(**over)


def check_single_ops_vs_oracle(over):
    ohp = _load(**over)
    assert audio._stft_parameters() == ao._stft_parameters(ohp)
    wav = speechlike(4100, 1)
    D = audio._stft(wav)
    Dref = ao._stft(wav.astype(np.float64), ohp)
    assert D.shape == Dref.shape and D.dtype == np.complex64 and D.flags.f_contiguous
    assert ao.rel_l2(D, Dref) < 1e-5            # BASELINE.json: single STFT within 1e-5 rel L2
    y = audio._istft(Dref)
    yref = ao._istft(Dref, ohp)
    assert y.dtype == np.float32 and y.shape == yref.shape
    assert ao.rel_l2(y, yref) < 1e-5
    assert ao.rel_l2(audio._istft(np.ascontiguousarray(Dref)), yref) < 1e-5   # bin-major input
    lin, mel = audio.spectrogram_and_mel(wav)
    assert lin.shape == (ohp.num_freq, D.shape[1]) and mel.shape == (80, D.shape[1]) and lin.dtype == np.float32
    assert ao.rel_l2(lin, ao.spectrogram(wav, ohp)) < 1e-5
    assert ao.rel_l2(mel, ao.melspectrogram(wav, ohp)) < 1e-5
    np.testing.assert_array_equal(audio.spectrogram(wav), lin)
    np.testing.assert_array_equal(audio.melspectrogram(wav), mel)
    M = np.abs(Dref)
    assert ao.rel_l2(audio._linear_to_mel(M), ao._linear_to_mel(M, ohp)) < 1e-6
    assert np.abs(audio._build_mel_basis() - ao._build_mel_basis(ohp)).max() < 1e-12
    x = speechlike(7001, 3)
    pe, de = audio.preemphasis(x), audio.inv_preemphasis(x)
    assert pe.dtype == np.float64 and de.dtype == np.float64
    assert ao.rel_l2(pe, ao.preemphasis(x, ohp)) < 1e-12
    assert ao.rel_l2(de, ao.inv_preemphasis(x, ohp)) < 1e-12
    # long signal: k_deemphasis cuts it into segments of 32768 samples, one CTA each, with a warm-up instead of a carry
    xl = speechlike(90001, 4)
    assert ao.rel_l2(audio.inv_preemphasis(xl), ao.inv_preemphasis(xl, ohp)) < 1e-12
    v = (np.random.RandomState(5).randn(777) * 50).astype(np.float32)
    assert ao.rel_l2(audio._amp_to_db(np.abs(v)), ao._amp_to_db(np.abs(v))) < 1e-6
    assert ao.rel_l2(audio._db_to_amp(v), ao._db_to_amp(v)) < 1e-6
    assert ao.rel_l2(audio._normalize(v), ao._normalize(v, ohp)) < 1e-6
    assert ao.rel_l2(audio._denormalize(v / 50), ao._denormalize(v / 50, ohp)) < 1e-6


def check_griffin_lim_vs_oracle(over):
    ohp = _load(**over)
    wav = speechlike(6000, 2)
    S = ao.spectrogram(wav, ohp)
    F, T = S.shape
    rs = np.random.RandomState(0)
    ang = np.exp(2j * np.pi * rs.rand(F, T))
    y = audio.inv_spectrogram(S, init_phase=ang, iters=5)
    yref = ao.inv_spectrogram(S, ohp, angles=ang, iters=5)
    assert y.dtype == np.float64 and y.shape == yref.shape == (audio._stft_parameters()[1] * (T - 1),)
    assert ao.snr_db(y, yref) > 60          # bar is 40 dB (BASELINE.json); fp32 pipeline sits far above
    y2 = audio.inv_spectrogram(np.ascontiguousarray(S), init_phase=np.ascontiguousarray(ang), iters=5)
    np.testing.assert_array_equal(y, y2)    # layout must not change a single bit (also a determinism check)
    S2 = rs.rand(F, 9).astype(np.float32)   # inconsistent (random) spectrogram, Tacotron-at-init like
    a2 = np.exp(2j * np.pi * rs.rand(F, 9))
    assert ao.snr_db(audio.inv_spectrogram(S2, init_phase=a2, iters=4), ao.inv_spectrogram(S2, ohp, angles=a2, iters=4)) > 60
    g = audio._griffin_lim(S2 * 3, init_phase=a2, iters=3)
    assert g.dtype == np.float32
    assert ao.snr_db(g, ao._griffin_lim(S2.astype(np.float64) * 3, ohp, angles=a2, iters=3)) > 60


def check_golden_fixtures_through_kernels(golden):
    """The committed fixtures came from the reference's own audio.py (tests/golden/make_golden.py)."""
    for tag, mn in (("yaml", 100), ("neg", -100)):
        _load(min_level_db=mn)
        wav = golden[tag + "_wav"]
        assert ao.rel_l2(audio._stft(audio.preemphasis(wav).astype(np.float32)), golden[tag + "_stft"]) < 1e-5
        lin, mel = audio.spectrogram_and_mel(wav)
        # rel-L2 is the stated bar; bins ~80 dB below the frame peak carry fp32-FFT noise of ~1e-2 dB, so the
        # element-wise bound is looser
        assert ao.rel_l2(lin, golden[tag + "_spec"]) < 1e-5 and np.abs(lin - golden[tag + "_spec"]).max() < 1e-3
        assert ao.rel_l2(mel, golden[tag + "_mel"]) < 1e-5 and np.abs(mel - golden[tag + "_mel"]).max() < 1e-3
        assert ao.rel_l2(audio._istft(golden[tag + "_stft"]), golden[tag + "_istft"]) < 1e-5
        F, T = golden[tag + "_gl_in"].shape
        np.random.seed(1234)       # the reference drew its phase from the global numpy RNG (audio.py:81)
        y = audio.inv_spectrogram(golden[tag + "_gl_in"], iters=4)
        assert ao.snr_db(y, golden[tag + "_gl_wav"]) > 60


def check_ragged_batch_and_tiles():
    """Ragged batch through the raw handle, several tile sizes: every tiling must give the same waveform."""
    ohp = _load(min_level_db=-100)
    h = audio._handle()
    rs = np.random.RandomState(3)
    Ts = [2, 3, 9, 41, 5]
    specs = [rs.rand(T, 1025).astype(np.float32) for T in Ts]
    phases = [np.exp(2j * np.pi * rs.rand(T, 1025)).astype(np.complex64) for T in Ts]
    packed, pph = np.concatenate(specs), np.concatenate(phases)
    refs = [ao.inv_spectrogram(s.T, ohp, angles=p.T, iters=3) for s, p in zip(specs, phases)]
    outs = []
    for tile in (0, 4, 8, 16, 28):
        h.set_tile_hops(tile)
        out = np.empty(sum(h.num_samples(T) for T in Ts), dtype=np.float64)
        h.griffin_lim(packed, _lib.FRAME_MAJOR, Ts, out, init_phase=pph, iters=3,
                      flags=_lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS, out_dtype=_lib.F64)
        outs.append(out)
        off = 0
        for T, ref in zip(Ts, refs):
            n = h.num_samples(T)
            assert ao.snr_db(out[off:off + n], ref) > 60, (tile, T)
            off += n
    h.set_tile_hops(0)
    # host-side pipelining: cutting the batch into chunks (H2D / compute / D2H overlap) must not change a bit either
    for chunks in (2, 3, 5):
        h.set_host_chunks(chunks)
        out = np.empty_like(outs[0])
        h.griffin_lim(packed, _lib.FRAME_MAJOR, Ts, out, init_phase=pph, iters=3,
                      flags=_lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS, out_dtype=_lib.F64)
        outs.append(out)
    h.set_host_chunks(0)
    for o in outs[1:]:
        np.testing.assert_array_equal(o, outs[0])      # the tiling must not change a single bit
    # ragged analysis batch
    wavs = [speechlike(n, i) for i, n in enumerate((300, 5118, 1, 2049))]
    ns = [len(w) for w in wavs]
    Tn = [h.num_frames(n) for n in ns]
    lin = np.empty((sum(Tn), 1025), np.float32)
    mel = np.empty((sum(Tn), 80), np.float32)
    h.features(np.concatenate(wavs), ns, lin, mel)
    off = 0
    for w, T in zip(wavs, Tn):
        # the 1-sample "utterance" is a constant after reflect padding: its spectrum spans > 130 dB, so bins
        # at 1e-4 of the peak sit inside the fp32 FFT's noise and only a loose dB bound is meaningful there
        tol = 1e-5 if len(w) > 100 else 5e-3
        assert ao.rel_l2(lin[off:off + T].T, ao.spectrogram(w, ohp)) < tol
        assert ao.rel_l2(mel[off:off + T].T, ao.melspectrogram(w, ohp)) < tol
        off += T


def check_errors_and_edge_cases():
    _load(min_level_db=-100)
    wav = speechlike(3000, 1)
    bad = wav.copy()
    bad[100] = np.nan
    with pytest.raises(audio.ParameterError):
        audio.spectrogram(bad)
    with pytest.raises(audio.ParameterError):
        audio._stft(np.zeros((2, 100), np.float32))
    with pytest.raises(ValueError):
        audio.spectrogram(np.zeros(0, np.float32))                         # librosa cannot reflect-pad an empty signal either
    # scipy.signal.lfilter (audio.py:32, 36) takes an empty array and integer samples
    for f, of in ((audio.preemphasis, ao.preemphasis), (audio.inv_preemphasis, ao.inv_preemphasis)):
        e = f(np.zeros(0, np.float32))
        assert e.shape == (0,) and e.dtype == np.float64
        ints = np.arange(-50, 50, dtype=np.int16)
        assert np.abs(f(ints) - of(ints.astype(np.float32), make_hp(min_level_db=-100))).max() < 1e-9
    from nspeech_b200 import batch as _batch
    with pytest.raises(ValueError):
        _batch.features_batch([])
    assert audio.find_endpoint(np.zeros(0)) == 0 and audio.peak_normalize(np.zeros(0)).shape == (0,)
    with pytest.raises(ValueError):
        audio.inv_spectrogram(np.zeros((1025, 1), np.float32))
    with pytest.raises(ValueError):
        audio.inv_spectrogram(np.zeros((513, 10), np.float32))
    S = np.full((1025, 4), np.inf, np.float32)
    with pytest.raises(audio.ParameterError):
        audio.inv_spectrogram(S, iters=1)
    # silence: phase of an all-zero STFT is 0 (np.angle(0) == 0) -> finite output, no NaN from 0/0
    y = audio._griffin_lim(np.zeros((1025, 6), np.float32), init_phase=np.ones((1025, 6), np.complex64), iters=2)
    assert np.all(y == 0)
    for bad_hp in ("num_freq=2", "num_freq=257", "num_freq=9000"):     # n_fft 2; window (1000) longer than n_fft 512; n_fft > 16384
        hp = hparams.load()
        hp.parse(bad_hp)
        with pytest.raises(ValueError):
            audio._stft(wav)
    hparams.load()


def check_device_random_phase_is_deterministic_per_seed():
    _load(min_level_db=-100)
    S = np.random.RandomState(1).rand(1025, 7).astype(np.float32)
    a = audio.inv_spectrogram(S, seed=7, iters=2)
    b = audio.inv_spectrogram(S, seed=7, iters=2)
    c = audio.inv_spectrogram(S, seed=8, iters=2)
    np.testing.assert_array_equal(a, b)
    assert np.isfinite(a).all() and not np.array_equal(a, c)


def check_tf_twin_vs_oracle(over):
    """The TensorFlow twin (reference audio.py:51-58, 90-103, 116-123) that synthesizer.py:30 and
    models/tacotron.py:107 call: tf.contrib.signal framing, zero initial phase, no window-sum normalisation."""
    ohp = _load(**over)
    h = audio._handle()
    wav = speechlike(5300, 1)
    D = audio._stft_tensorflow(wav)
    Dref = tfo._stft_tensorflow(wav, ohp)
    assert D.shape == Dref.shape == (h.num_frames_tf(wav.size), 1025) and D.dtype == np.complex64
    assert ao.rel_l2(D, Dref) < 1e-5
    Db = audio._stft_tensorflow(np.stack([wav, wav[::-1]]))                  # batched [N, n] -> [N, T, F]
    assert Db.shape == (2,) + Dref.shape
    np.testing.assert_array_equal(Db[0], D)
    assert ao.rel_l2(Db[1], tfo._stft_tensorflow(wav[::-1].copy(), ohp)) < 1e-5
    y = audio._istft_tensorflow(Dref)
    yref = tfo._istft_tensorflow(Dref, ohp)
    assert y.dtype == np.float32 and y.shape == yref.shape == (h.num_samples_tf(Dref.shape[0]),)
    assert ao.rel_l2(y, yref) < 1e-5
    # Griffin-Lim twin on a random (Tacotron-at-init like) normalised spectrogram, single and batched
    rs = np.random.RandomState(0)
    S = rs.rand(12, 1025).astype(np.float32)
    g = audio.inv_spectrogram_tensorflow(S, iters=4)
    gref = tfo.inv_spectrogram_tensorflow(S, ohp, iters=4)
    assert g.dtype == np.float32 and g.shape == gref.shape
    assert ao.snr_db(g, gref) > 60
    S3 = rs.rand(3, 9, 1025).astype(np.float32)
    g3 = audio.inv_spectrogram_tensorflow(S3, iters=3)
    assert g3.shape == (3, h.num_samples_tf(9))
    for i in range(3):
        assert ao.snr_db(g3[i], tfo.inv_spectrogram_tensorflow(S3[i], ohp, iters=3)) > 60
        np.testing.assert_array_equal(g3[i], audio.inv_spectrogram_tensorflow(S3[i], iters=3))   # batch == single, bitwise
    # consistent magnitudes (|STFT| of a real signal), raw _griffin_lim_tensorflow
    Sc = np.abs(Dref)
    assert ao.snr_db(audio._griffin_lim_tensorflow(Sc, iters=5), tfo._griffin_lim_tensorflow(Sc, ohp, iters=5)) > 60
    # silence: est = 0 -> angles = 0 / 1e-8 = 0 -> all-zero waveform, no NaN
    assert np.all(audio._griffin_lim_tensorflow(np.zeros((5, 1025), np.float32), iters=2) == 0)
    with pytest.raises(ValueError):
        audio._stft_tensorflow(wav[:10])
    with pytest.raises(ValueError):
        audio.inv_spectrogram_tensorflow(np.zeros((4, 513), np.float32))
    assert ao.rel_l2(audio._db_to_amp_tensorflow(S[0] * 40 - 20), tfo._db_to_amp_tensorflow(S[0] * 40 - 20)) < 1e-6
    assert ao.rel_l2(audio._denormalize_tensorflow(S[0] * 1.2 - 0.1), tfo._denormalize_tensorflow(S[0] * 1.2 - 0.1, ohp)) < 1e-6


def check_streaming_rounds(T=150, iters=2, tf=False, emulated=False):
    """k_gl_stream on long pieces: one CTA walks many rounds of its 8-group ring (slot reuse, wrap-around of the
    accumulate, the warp-7 -> warp-0 dependency across rounds), against the oracle and bit-for-bit against every
    other partition of the same batch and against the older kernels."""
    ohp = _load(min_level_db=-100)
    h = audio._handle()
    rs = np.random.RandomState(11)
    Ts = [T, 2, 37, T // 2 + 3]
    specs = [rs.rand(t, 1025).astype(np.float32) for t in Ts]
    packed = np.concatenate(specs)
    flags = _lib.GL_DENORMALIZE | (_lib.GL_TF_TWIN if tf else _lib.GL_DEEMPHASIS)
    nsamp = [h.num_samples_tf(t) if tf else h.num_samples(t) for t in Ts]
    phases = None if tf else np.concatenate([np.exp(2j * np.pi * rs.rand(t, 1025)).astype(np.complex64) for t in Ts])

    def run():
        out = np.empty(sum(nsamp), dtype=np.float32 if tf else np.float64)
        h.griffin_lim(packed, _lib.FRAME_MAJOR, Ts, out, init_phase=phases, iters=iters, flags=flags,
                      out_dtype=_lib.F32 if tf else _lib.F64)
        return out
    outs = {}
    try:
        h.set_generic_iteration(0)           # k_gl_stream
        for grid in ((1, 2, 0) if emulated else (1, 2, 3, 0)):       # (the CPU emulation runs a reduced set: every run is seconds there)
            h.set_stream_grid(grid)
            outs["grid%d" % grid] = run()
        if not emulated:
            h.set_stream_grid(1)
            h.set_tile_hops(12)
            outs["piece3"] = run()
        h.set_tile_hops(0)
        h.set_generic_iteration(2)           # the tile kernel adds in the same colour order: identical bits
        h.set_option(_lib.OPT_WIDE_MODE, 0)             # ... with C consecutive frames per warp
        outs["tile"] = run()
        if not emulated:
            h.set_option(_lib.OPT_WIDE_MODE, 1)         # ... and in its wide (low-latency) mode: one frame per warp and step
            outs["tile_wide"] = run()
        # the instantiations with run-time geometry (what other hop / window lengths use) against the ones with the
        # default hparams folded into immediates
        h.set_option(_lib.OPT_SPECIALIZE, 0)
        h.set_option(_lib.OPT_WIDE_MODE, 0)
        outs["tile_general"] = run()
        h.set_generic_iteration(0)
        h.set_stream_grid(2)
        outs["stream_general"] = run()
        h.set_option(_lib.OPT_SPECIALIZE, 1)
        h.set_option(_lib.OPT_STREAM_SYNC_MODE, 10)     # split-phase barrier: publish after the overlap-add, wait before the next one
        if not emulated:                                # (its spin-waits are what OS threads are worst at: one run is enough on the CPU)
            outs["stream_split_phase"] = run()
        h.set_stream_grid(0)
        outs["stream_split_phase_grid0"] = run()
    finally:
        h.set_option(_lib.OPT_STREAM_SYNC_MODE, 2)
        h.set_stream_grid(0); h.set_tile_hops(0); h.set_generic_iteration(-1); h.set_option(_lib.OPT_WIDE_MODE, -1)
        h.set_option(_lib.OPT_SPECIALIZE, 1)
    ref = outs.pop("grid1")
    for name, o in outs.items():
        np.testing.assert_array_equal(o, ref, err_msg=name)
    off = 0
    for i, (t, n) in enumerate(zip(Ts, nsamp)):
        if tf:
            want = tfo.inv_spectrogram_tensorflow(specs[i], ohp, iters=iters)
        else:
            ph = phases[sum(Ts[:i]):sum(Ts[:i + 1])]
            want = ao.inv_spectrogram(specs[i].T, ohp, angles=ph.T, iters=iters)
        assert ao.snr_db(ref[off:off + n], want) > 60, (i, t)
        off += n


def check_find_endpoint_and_synthesis_stage(golden):
    """find_endpoint (reference audio.py:67-74) and the fused spectrogram -> waveform stage of Synthesizer.synthesize
    (synthesizer.py:30, 51-53: inv_spectrogram_tensorflow -> inv_preemphasis -> find_endpoint)."""
    ohp = _load(min_level_db=-100)
    # the reference's own find_endpoint on this input (tests/golden/make_golden.py)
    ep_in = np.concatenate([speechlike(30000, 3), np.zeros(30000, np.float32)])
    assert len(ep_in) == int(golden["endpoint_in_len"])
    assert audio.find_endpoint(ep_in) == int(golden["endpoint"]) == ao.find_endpoint(ep_in, ohp)
    rs = np.random.RandomState(4)
    cases = []
    w = speechlike(70000, 5).astype(np.float64)
    w[21000:52000] *= 1e-4                                   # a silent stretch in the middle
    cases.append((w, {}))
    cases.append((w.astype(np.float32), {}))
    cases.append((speechlike(50000, 6), {}))                 # never silent -> len(wav)
    cases.append((speechlike(9000, 7), {}))                  # shorter than one window -> len(wav)
    cases.append((np.zeros(40000, np.float32), {}))          # silent from the start -> 2 * hop
    cases.append((w, {"threshold_db": -25, "min_silence_sec": 0.35}))
    cases.append((w, {"min_silence_sec": 0.80005}))          # window % 4 != 0: the window is 4 hops + 1 sample
    neg = -np.abs(speechlike(60000, 8)).astype(np.float64)   # np.max, not max|.|: an all-negative loud signal counts as silent
    cases.append((neg, {}))
    cases.append((rs.randn(33000) * 0.004, {}))              # noise whose peaks straddle the threshold
    for wav, kw in cases:
        assert audio.find_endpoint(wav, **kw) == ao.find_endpoint(wav, ohp, **kw), (wav.dtype, len(wav), kw)
    with pytest.raises(ValueError):
        audio.find_endpoint(np.zeros((2, 10)))
    # the fused stage: loud frames followed by frames at the floor (normalised 0 -> 1e-6 amplitude), batched and single
    T = 140
    specs = rs.rand(2, T, 1025).astype(np.float32)
    specs[0, 50:] = 0.0
    outs = audio.synthesize_waveforms(specs, iters=3)
    assert len(outs) == 2
    for i in range(2):
        ref = ao.inv_preemphasis(tfo.inv_spectrogram_tensorflow(specs[i], ohp, iters=3), ohp)
        end = ao.find_endpoint(ref, ohp)
        assert outs[i].dtype == np.float64 and outs[i].shape == (end,), (i, outs[i].shape, end)
        assert ao.snr_db(outs[i], ref[:end]) > 60
    assert outs[0].size < outs[1].size == audio._handle().num_samples_tf(T)
    single = audio.synthesize_waveforms(specs[0], iters=3)
    np.testing.assert_array_equal(single, outs[0])


def _span(sub, wav):
    if sub.size == 0:
        return (0, 0)
    return ((sub.__array_interface__["data"][0] - wav.__array_interface__["data"][0]) // wav.itemsize, sub.size)


def check_trimming(golden_process):
    """trim_wav / trim_silence (reference datasets/process.py:39-54): frame energies from the GPU, the reference's interval
    logic on the host; sample indices must equal the oracle's and the fixtures made by the reference's own process.py."""
    from nspeech_b200 import process
    from oracle import process_oracle as po
    _load(min_level_db=-100)
    for name, wav in trim_signals().items():
        for fl, hop in ((1024, 512), (2048, 512)):
            if wav.size >= fl:
                assert ao.rel_l2(process.frame_energy(wav, fl, hop), (po.rmse(wav.astype(np.float64), fl, hop) ** 2)[0]) < 1e-12
        t = process.trim_wav(wav)
        assert _span(t, wav) == _span(po.trim_wav(wav), wav) == tuple(golden_process[name + "_trim_wav"]), name
        for thr in (0.01, 0.1):
            s = process.trim_silence(wav, thr)
            assert _span(s, wav) == _span(po.trim_silence(wav, thr), wav) == tuple(golden_process["%s_trim_silence_%g" % (name, thr)]), (name, thr)
    wav = trim_signals()["quiet_ends"]
    w2, lin_t, mel_t, n_frames = process.process_utterance_arrays(wav)
    assert _span(w2, wav) == tuple(golden_process["quiet_ends_trim_wav"])
    assert lin_t.shape == (n_frames, 1025) and mel_t.shape == (n_frames, 80) and n_frames == 1 + w2.size // 250


def check_feeder_targets():
    """batch.feeder_targets = the reference feeder's _prepare_targets over spectrogram(w).T / melspectrogram(w).T
    (datasets/datafeeder.py:190-216): padded, time-major, stacked - written by the feature kernel itself."""
    from nspeech_b200 import batch
    ohp = _load(min_level_db=-100)
    wavs = [speechlike(n, i) for i, n in enumerate((5200, 900, 12345))]
    r = 5
    mel, lin, Ts = batch.feeder_targets(wavs, r)
    max_len = max(Ts) + 1
    rows = max_len if max_len % r == 0 else max_len + r - max_len % r
    assert lin.shape == (3, rows, 1025) and mel.shape == (3, rows, 80) and lin.dtype == mel.dtype == np.float32
    for i, w in enumerate(wavs):
        assert Ts[i] == 1 + w.size // 250
        assert ao.rel_l2(lin[i, :Ts[i]], ao.spectrogram(w, ohp).T) < 1e-5
        assert ao.rel_l2(mel[i, :Ts[i]], ao.melspectrogram(w, ohp).T) < 1e-5
        assert not lin[i, Ts[i]:].any() and not mel[i, Ts[i]:].any()          # _pad = 0


def check_save_wav_scaling_and_int16(golden=None):
    """save_wav's scaling (reference audio.py:17-19) on the device: bit-exact float64, bit-exact int16 (integer work), alone
    and fused into the synthesis stage (synthesizer.py:51-53 -> eval.py:43); ragged lists of spectrograms."""
    ohp = _load(min_level_db=-100)
    rs = np.random.RandomState(12)
    for wav in (speechlike(33000, 21).astype(np.float64) * 0.37, speechlike(5000, 22), rs.randn(70001) * 3.0,
                np.zeros(300), np.full(10, 1e-5), -np.abs(speechlike(2000, 23)).astype(np.float64)):
        ref = ao.save_wav_scaling(wav.astype(np.float64))
        got = audio.peak_normalize(wav)
        assert got.dtype == np.float64
        np.testing.assert_array_equal(got, ref)                       # one IEEE multiply per sample by the same factor
        np.testing.assert_array_equal(audio.peak_normalize(wav, dtype=np.int16), ref.astype(np.int16))
    w = speechlike(4000, 24).astype(np.float64)
    w2 = w.copy()
    import os, tempfile
    from scipy.io import wavfile
    with tempfile.TemporaryDirectory() as d:
        audio.save_wav(w2, os.path.join(d, "a.wav"))                  # mutates its argument like the reference
        np.testing.assert_array_equal(w2, ao.save_wav_scaling(w))
        sr, back = wavfile.read(os.path.join(d, "a.wav"))
        assert sr == 20000
        np.testing.assert_array_equal(back, w2)
        lin = audio.spectrogram(speechlike(3000, 25))
        audio.save_spectrogram(lin, os.path.join(d, "s.npy"))
        spec, n = audio.load_spectrogram(os.path.join(d, "s.npy"))
        assert n == lin.shape[1] and spec.flags.f_contiguous
        np.testing.assert_array_equal(spec, lin)
    # fused: ragged list, trimmed at the endpoint, scaled by the peak of the TRIMMED waveform
    Ts = [140, 33, 60]
    specs = [rs.rand(T, 1025).astype(np.float32) for T in Ts]
    specs[0][50:] = 0.0
    plain = audio.synthesize_waveforms(specs, iters=3)
    scaled = audio.synthesize_waveforms(specs, iters=3, peak_normalize=True)
    ints = audio.synthesize_waveforms(specs, iters=3, peak_normalize=True, dtype=np.int16)
    assert len(plain) == len(scaled) == len(ints) == 3
    for i in range(3):
        ref = ao.inv_preemphasis(tfo.inv_spectrogram_tensorflow(specs[i], ohp, iters=3), ohp)
        ref = ref[:ao.find_endpoint(ref, ohp)]
        assert plain[i].shape == ref.shape and ao.snr_db(plain[i], ref) > 60
        np.testing.assert_array_equal(scaled[i], ao.save_wav_scaling(plain[i]))         # the same device result, scaled like numpy does
        assert ints[i].dtype == np.int16
        np.testing.assert_array_equal(ints[i], scaled[i].astype(np.int16))
        assert np.abs(ints[i].astype(np.int64) - ao.save_wav_scaling(ref).astype(np.int16)).max() <= 1   # vs the oracle's own pipeline
        np.testing.assert_array_equal(audio.synthesize_waveforms(specs[i], iters=3), plain[i])            # batch == single, bitwise
    assert plain[0].size < audio._handle().num_samples_tf(Ts[0])
    with pytest.raises(ValueError):
        audio.synthesize_waveforms(specs, iters=1, dtype=np.int16)


def check_feeder_groups():
    """batch.feeder_groups = one group of the reference feeder (datasets/datafeeder.py:130-158, 190-216): bucket by length,
    pad every batch on its own - the feature kernel writes the padded batch tensors directly (nsb_features_rows)."""
    from nspeech_b200 import batch
    from oracle import feeder_oracle as fo
    ohp = _load(min_level_db=-100)
    lens = (5200, 900, 12345, 2500, 2500, 777, 9000)
    wavs = [speechlike(n, 40 + i) for i, n in enumerate(lens)]
    n, r = 3, 5
    groups = batch.feeder_groups(wavs, n, r)
    examples = [(i, w, ao.melspectrogram(w, ohp).T, ao.spectrogram(w, ohp).T, 1 + w.size // 250) for i, w in enumerate(wavs)]
    want = fo.bucket(examples, n)
    assert [g["indices"] for g in groups] == [[e[0] for e in b] for b in want]
    for g, b in zip(groups, want):
        mel_ref = fo._prepare_targets([e[2] for e in b], r)
        lin_ref = fo._prepare_targets([e[3] for e in b], r)
        assert g["mel_targets"].shape == mel_ref.shape and g["linear_targets"].shape == lin_ref.shape
        assert ao.rel_l2(g["mel_targets"], mel_ref) < 1e-5 and ao.rel_l2(g["linear_targets"], lin_ref) < 1e-5
        np.testing.assert_array_equal(g["mel_targets"] == 0, mel_ref == 0)           # the padding (and only it) is exactly _pad = 0
        np.testing.assert_array_equal(g["audios"], fo._prepare_inputs([e[1] for e in b]))
    import random
    shuffled = batch.bucket_by_length([e[4] for e in examples], n, rng=random.Random(3))
    assert sorted(map(tuple, shuffled)) == sorted(tuple(e[0] for e in b) for b in want)


def check_async_submit_wait():
    """nsb_*_submit / nsb_wait: two batches in flight on private worker slots; results must equal the synchronous calls bit
    for bit, in any wait order; a ticket is good for one wait."""
    from nspeech_b200 import batch
    _load(min_level_db=-100)
    h = audio._handle()
    rs = np.random.RandomState(31)
    jobs = []
    for i, Ts in enumerate(([9, 41], [17], [5, 6, 7])):
        spec = np.concatenate([rs.rand(T, 1025).astype(np.float32) for T in Ts])
        sync = np.empty(sum(h.num_samples(T) for T in Ts), np.float64)
        h.griffin_lim(spec, _lib.FRAME_MAJOR, Ts, sync, seed=5 + i, iters=3, flags=_lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS, out_dtype=_lib.F64)
        out = np.full_like(sync, np.nan)
        jobs.append((Ts, spec, sync, out))
    tickets = [h.griffin_lim_submit(spec, _lib.FRAME_MAJOR, Ts, out, seed=5 + i, iters=3, flags=_lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS,
                                    out_dtype=_lib.F64) for i, (Ts, spec, sync, out) in enumerate(jobs)]
    assert len(set(tickets)) == 3
    for t in (tickets[2], tickets[0], tickets[1]):
        h.wait(t)
    for Ts, spec, sync, out in jobs:
        np.testing.assert_array_equal(out, sync)
    with pytest.raises(ValueError):
        h.wait(tickets[0])                       # collected already
    with pytest.raises(ValueError):
        h.wait(10 ** 9)
    # a failing job reports through its ticket
    bad, bad_out = np.full((4, 1025), np.inf, np.float32), np.empty(h.num_samples(4))     # (buffers must outlive the call: keep the names)
    t = h.griffin_lim_submit(bad, _lib.FRAME_MAJOR, [4], bad_out, iters=1, flags=_lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS,
                             out_dtype=_lib.F64)
    with pytest.raises(audio.ParameterError):
        h.wait(t)
    # features, and the generator front end
    wav = speechlike(6000, 33)
    T = h.num_frames(wav.size)
    lin, mel = np.empty((T, 1025), np.float32), np.empty((T, 80), np.float32)
    h.wait(h.features_submit(wav, [wav.size], lin, mel))
    l2, m2 = audio.spectrogram_and_mel(wav)
    np.testing.assert_array_equal(lin.T, l2)
    np.testing.assert_array_equal(mel.T, m2)
    batches = [[rs.rand(1025, T).astype(np.float32) for T in Ts] for Ts in ([4, 9], [12], [3, 3, 3], [8])]
    got = list(batch.inv_spectrogram_stream(batches, seed=11, iters=2, layout="FT"))
    assert len(got) == 4
    for i, (b, outs) in enumerate(zip(batches, got)):
        ref = batch.inv_spectrogram_batch(b, seed=11 + i, iters=2, layout="FT")
        for a, c in zip(outs, ref):
            np.testing.assert_array_equal(a, c)
    # float32 results (half the bytes back): the de-emphasised float64 waveform rounded once
    got32 = list(batch.inv_spectrogram_stream(batches[:2], seed=11, iters=2, layout="FT", dtype=np.float32))
    for outs32, outs64 in zip(got32, got):
        for a, c in zip(outs32, outs64):
            assert a.dtype == np.float32
            np.testing.assert_array_equal(a, c.astype(np.float32))
    with pytest.raises(ValueError):
        list(batch.inv_spectrogram_stream(batches[:1], dtype=np.int16))


def check_api_guards():
    """stale Griffin-Lim state, ambiguous layouts, undersized buffers."""
    from nspeech_b200 import batch
    _load(min_level_db=-100)
    h = audio._handle()
    rs = np.random.RandomState(2)
    with pytest.raises(ValueError):
        batch.inv_spectrogram_batch([rs.rand(1025, 1025).astype(np.float32)], iters=1)          # square: which axis is time?
    S = rs.rand(1025, 7).astype(np.float32)
    a = batch.inv_spectrogram_batch([S], seed=1, iters=2, layout="FT")[0]
    b = batch.inv_spectrogram_batch([np.ascontiguousarray(S.T)], seed=1, iters=2, layout="TF")[0]
    np.testing.assert_array_equal(a, b)
    with pytest.raises(ValueError):
        batch.inv_spectrogram_batch([S], iters=1, layout="TF")
    with pytest.raises(ValueError):
        batch.inv_spectrogram_batch([S], iters=1, out=np.empty(10, np.float64))                 # too small
    with pytest.raises(ValueError):
        batch.inv_spectrogram_batch([S], iters=1, out=np.empty(h.num_samples(7), np.float32))   # wrong dtype
    with pytest.raises(ValueError):
        batch.inv_spectrogram_batch([S], iters=1, init_phase=[np.ones((1025, 6), np.complex64)])
    with pytest.raises(ValueError):
        h.griffin_lim(S.T.copy(), _lib.FRAME_MAJOR, [7], np.empty(5, np.float64), iters=1, flags=_lib.GL_DEEMPHASIS, out_dtype=_lib.F64)
    with pytest.raises(ValueError):
        h.features(speechlike(1000, 1), [1000], np.empty((2, 1025), np.float32), None)


def check_stale_griffin_lim_state(to_dev, stream=None):
    """nsb_griffin_lim_iterate runs on the state of the last NSB_DEVICE Griffin-Lim call; any later call that rewrites the
    handle's descriptors must invalidate it.  ``to_dev`` puts a numpy array where NSB_DEVICE pointers live."""
    _load(min_level_db=-100)
    h = audio._handle()
    rs = np.random.RandomState(4)
    spec = to_dev(rs.rand(9, 1025).astype(np.float32))
    out = to_dev(np.zeros(h.num_samples(9), np.float64))
    flags = _lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS
    h.griffin_lim(spec, _lib.FRAME_MAJOR, [9], out, seed=1, iters=1, flags=flags, out_dtype=_lib.F64, space=_lib.DEVICE, stream=stream)
    h.griffin_lim_iterate(1, stream)
    wav = to_dev(speechlike(2000, 1))
    lin = to_dev(np.zeros((h.num_frames(2000), 1025), np.float32))
    h.features(wav, [2000], lin, None, space=_lib.DEVICE, stream=stream)
    with pytest.raises(ValueError):
        h.griffin_lim_iterate(1, stream)
    h.synchronize(stream)
    # a HOST call leaves no device-resident state either
    host_out = np.empty(h.num_samples(9), np.float64)
    h.griffin_lim(rs.rand(9, 1025).astype(np.float32), _lib.FRAME_MAJOR, [9], host_out, seed=1, iters=1, flags=flags, out_dtype=_lib.F64)
    with pytest.raises(ValueError):
        h.griffin_lim_iterate(1, stream)


def check_generic_tf_twin_and_stages():
    """The generic-size kernels behind the rest of the interface at num_freq = 513 (n_fft 1024): the TensorFlow twin
    (audio.py:51-58, 90-123), the fused synthesis stage (synthesizer.py:51-53), ragged batches in both layouts, the device RNG."""
    from nspeech_b200 import batch
    ohp = _load(min_level_db=-100, num_freq=513)
    h = audio._handle()
    F = 513
    assert (h.n_fft, h.num_freq) == (1024, F)
    wav = speechlike(5300, 1)
    D = audio._stft_tensorflow(wav)
    Dref = tfo._stft_tensorflow(wav, ohp)
    assert D.shape == Dref.shape == (h.num_frames_tf(wav.size), F)
    assert ao.rel_l2(D, Dref) < 1e-5
    assert ao.rel_l2(audio._istft_tensorflow(Dref), tfo._istft_tensorflow(Dref, ohp)) < 1e-5
    rs = np.random.RandomState(0)
    S3 = rs.rand(3, 9, F).astype(np.float32)
    g3 = audio.inv_spectrogram_tensorflow(S3, iters=3)
    for i in range(3):
        assert ao.snr_db(g3[i], tfo.inv_spectrogram_tensorflow(S3[i], ohp, iters=3)) > 60
    assert np.all(audio._griffin_lim_tensorflow(np.zeros((5, F), np.float32), iters=2) == 0)
    specs = rs.rand(2, 140, F).astype(np.float32)
    specs[0, 50:] = 0.0
    outs = audio.synthesize_waveforms(specs, iters=3, peak_normalize=True)
    for i in range(2):
        ref = ao.inv_preemphasis(tfo.inv_spectrogram_tensorflow(specs[i], ohp, iters=3), ohp)
        ref = ao.save_wav_scaling(ref[:ao.find_endpoint(ref, ohp)])
        assert outs[i].shape == ref.shape and ao.snr_db(outs[i], ref) > 60
    # ragged Griffin-Lim batch, [F,T] and [T,F] members, against one-at-a-time calls (bitwise) and the oracle
    Ts = [2, 17, 5]
    mats = [rs.rand(F, T).astype(np.float32) for T in Ts]
    phs = [np.exp(2j * np.pi * rs.rand(F, T)).astype(np.complex64) for T in Ts]
    outs = batch.inv_spectrogram_batch(mats, init_phase=phs, iters=3, layout="FT")
    outs_t = batch.inv_spectrogram_batch([np.ascontiguousarray(m.T) for m in mats], init_phase=[np.ascontiguousarray(p.T) for p in phs], iters=3, layout="TF")
    for m, p, o, ot in zip(mats, phs, outs, outs_t):
        assert ao.snr_db(o, ao.inv_spectrogram(m, ohp, angles=p, iters=3)) > 60
        np.testing.assert_array_equal(o, ot)
        np.testing.assert_array_equal(o, audio.inv_spectrogram(m, init_phase=p, iters=3))
        np.testing.assert_array_equal(o, audio.inv_spectrogram(np.ascontiguousarray(m), init_phase=np.ascontiguousarray(p), iters=3))   # bin-major
    a = audio.inv_spectrogram(mats[1], seed=7, iters=2)
    np.testing.assert_array_equal(a, audio.inv_spectrogram(mats[1], seed=7, iters=2))
    assert np.isfinite(a).all() and not np.array_equal(a, audio.inv_spectrogram(mats[1], seed=8, iters=2))
    # ragged features + the feeder's padded tensors
    wavs = [speechlike(n, i) for i, n in enumerate((3000, 5118, 700))]
    feats = batch.features_batch(wavs)
    mel_t, lin_t, Tn = batch.feeder_targets(wavs, 5)
    for i, (w, (lin, mel)) in enumerate(zip(wavs, feats)):
        assert lin.shape == (F, Tn[i]) and ao.rel_l2(lin, ao.spectrogram(w, ohp)) < 1e-5 and ao.rel_l2(mel, ao.melspectrogram(w, ohp)) < 1e-5
        np.testing.assert_array_equal(lin_t[i, :Tn[i]], lin.T)
        assert not lin_t[i, Tn[i]:].any()
    with pytest.raises(audio.ParameterError):
        audio.inv_spectrogram(np.full((F, 4), np.inf, np.float32), iters=1)


def check_mel_projection_modes():
    """The mel projection's four forms (NSB_OPT_MEL_LINES: sparse rows, line segments on the plain / the skewed magnitude row, the
    segments cut into balanced pieces) against the oracle; the two row layouts add in the same order (identical bits)."""
    ohp = _load(min_level_db=-100)
    h = audio._handle()
    wavs = [speechlike(9000, 3), (0.4 * np.random.RandomState(5).randn(5000)).astype(np.float32), speechlike(777, 9)]
    want = [(ao.spectrogram(w, ohp), ao.melspectrogram(w, ohp)) for w in wavs]
    got = {}
    try:
        for mode in (0, 1, 2, 3):
            h.set_option(_lib.OPT_MEL_LINES, mode)
            got[mode] = [audio.spectrogram_and_mel(w) for w in wavs]
            for (lin, mel), (wl, wm) in zip(got[mode], want):
                assert ao.rel_l2(lin, wl) < 1e-5 and ao.rel_l2(mel, wm) < 1e-5, mode
    finally:
        h.set_option(_lib.OPT_MEL_LINES, 3)
    for i in range(len(wavs)):
        np.testing.assert_array_equal(got[1][i][1], got[2][i][1])
        np.testing.assert_array_equal(got[0][i][0], got[3][i][0])          # the linear feature does not depend on the mode
        assert ao.rel_l2(got[3][i][1], got[2][i][1]) < 1e-6


def check_features_pipeline():
    """batch.features_batch as a pipeline of clip groups (nsb_features_submit / nsb_wait, packing on host threads) against
    the single synchronous call: identical bits whatever the grouping."""
    from nspeech_b200 import batch
    ohp = _load(min_level_db=-100)
    many = [speechlike(600 + 37 * i, i) for i in range(34)]
    one = batch.features_batch(many, in_flight=0)
    for kw in ({"group_bytes": 1 << 14}, {"group_bytes": 1 << 16, "in_flight": 1}, {"group_bytes": 1 << 14, "want_linear": False}):
        piped = batch.features_batch(many, **kw)
        assert len(piped) == len(one)
        for (l0, m0), (l1, m1) in zip(one, piped):
            if kw.get("want_linear", True):
                np.testing.assert_array_equal(l0, l1)
            else:
                assert l1 is None
            np.testing.assert_array_equal(m0, m1)
    assert ao.rel_l2(one[5][0], ao.spectrogram(many[5], ohp)) < 1e-5 and ao.rel_l2(one[33][1], ao.melspectrogram(many[33], ohp)) < 1e-5
    assert batch._group_cuts([5] * 34, 4) == [0, 9, 17, 26, 34] and batch._group_cuts([3, 1], 16) == [0, 1, 2]


def check_tf_twin_golden(golden_tf):
    """Fixtures written by the reference's OWN TF-twin functions (tests/golden/make_golden_tf.py: audio.py:51-58, 90-123 run
    unmodified on an eager numpy stand-in for tf) through the kernels and the oracle."""
    for tag, mn in (("yaml", 100), ("neg", -100)):
        ohp = _load(min_level_db=mn)
        S, mag = golden_tf[tag + "_tw_in"], golden_tf[tag + "_tw_S"]
        assert ao.snr_db(tfo.inv_spectrogram_tensorflow(S, make_hp(min_level_db=mn, griffin_lim_iters=4), iters=4), golden_tf[tag + "_tw_wav"]) > 100   # oracle == reference
        y = audio.inv_spectrogram_tensorflow(S, iters=4)
        assert y.dtype == np.float32 and y.shape == golden_tf[tag + "_tw_wav"].shape
        assert ao.snr_db(y, golden_tf[tag + "_tw_wav"]) > 60
        assert ao.snr_db(audio._griffin_lim_tensorflow(mag, iters=4), golden_tf[tag + "_tw_raw"]) > 60
    _load()
    assert ao.rel_l2(audio._stft_tensorflow(golden_tf["tw_sig"]), golden_tf["tw_stft"]) < 1e-5
    assert ao.rel_l2(audio._istft_tensorflow(golden_tf["tw_stft"]), golden_tf["tw_istft"]) < 1e-5


def check_length_and_hparams_sweep(cases):
    """(n_samples, sample_rate, frame_length_ms, num_freq) cases: the STFT, both features and the iSTFT round trip against
    the oracle - clip lengths from one sample to several transforms (incl. len <= n_fft / 2, where the reflect padding folds
    more than once), the sample rates 16 / 20 / 22.05 / 24 kHz (truncating int() of audio.py:128-129), window lengths."""
    for n, sr, fl, nf in cases:
        over = {"min_level_db": -100, "sample_rate": sr, "frame_length_ms": fl, "num_freq": nf}
        ohp = _load(**over)
        n_fft, hop, win = ao._stft_parameters(ohp)
        if win > n_fft:
            with pytest.raises(ValueError):
                audio._stft(speechlike(max(n, 2), 1))
            continue
        wav = speechlike(n, n % 97, sr=sr)
        D = audio._stft(wav)
        Dref = ao._stft(wav.astype(np.float64), ohp)
        assert D.shape == Dref.shape == (nf, 1 + n // hop), (n, sr, fl, nf)
        # short clips are (near-)constant after the reflection: the spectrum spans > 100 dB and only its large bins carry a relative bound
        tol = 1e-5 if n > 100 else 1e-4
        assert ao.rel_l2(D, Dref) < tol, (n, sr, fl, nf, ao.rel_l2(D, Dref))
        lin, mel = audio.spectrogram_and_mel(wav)
        ftol = 1e-5 if n > 100 else 5e-3
        assert ao.rel_l2(lin, ao.spectrogram(wav, ohp)) < ftol, (n, sr, fl, nf)
        assert ao.rel_l2(mel, ao.melspectrogram(wav, ohp)) < ftol, (n, sr, fl, nf)
        if D.shape[1] >= 2:
            assert ao.rel_l2(audio._istft(Dref), ao._istft(Dref, ohp)) < 1e-5, (n, sr, fl, nf)
    hparams.load()
