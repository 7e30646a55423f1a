"""Pins the CPU oracle: (1) against the fixtures the reference's own audio.py produced
(tests/golden/make_golden.py), (2) against independent implementations (torch / torchaudio /
scipy), (3) against analytic known answers.  CPU only."""
import numpy as np
import pytest
import torch

from conftest import make_hp, speechlike
from oracle import audio_oracle as ao
from oracle import librosa060


@pytest.mark.parametrize("tag,mn", [("yaml", 100), ("neg", -100)])
def test_oracle_matches_reference_composition(golden, tag, mn):
    hp = make_hp(min_level_db=mn)
    wav = golden[tag + "_wav"]
    np.testing.assert_array_equal(ao.preemphasis(wav, hp), golden[tag + "_pre"])
    D = ao._stft(ao.preemphasis(wav, hp), hp)
    np.testing.assert_array_equal(D, golden[tag + "_stft"])
    np.testing.assert_array_equal(ao.spectrogram(wav, hp), golden[tag + "_spec"])
    np.testing.assert_array_equal(ao.melspectrogram(wav, hp), golden[tag + "_mel"])
    np.testing.assert_array_equal(ao._istft(D, hp), golden[tag + "_istft"])
    F, T = golden[tag + "_gl_in"].shape
    rs = np.random.RandomState(1234)
    angles = np.exp(2j * np.pi * rs.rand(F, T))
    y = ao.inv_spectrogram(golden[tag + "_gl_in"], hp, angles=angles, iters=4)
    np.testing.assert_array_equal(y, golden[tag + "_gl_wav"])
    assert y.dtype == np.float64 and len(y) == 250 * (T - 1)
    y = ao._griffin_lim(golden[tag + "_gl_S"], hp, angles=angles, iters=4)
    np.testing.assert_array_equal(y, golden[tag + "_gl_raw"])
    assert y.dtype == np.float32


def test_neg_spec_is_not_saturated(golden):
    # with the yaml's +100 the normalised features saturate; the -100 variant is the
    # non-vacuous one
    s = golden["neg_spec"]
    assert 0.05 < np.mean((s > 0) & (s < 1))


def test_misc_goldens(golden):
    hp = make_hp()
    assert tuple(golden["stft_parameters"]) == ao._stft_parameters(hp) == (2048, 250, 1000)
    basis = ao._build_mel_basis(hp)
    assert basis.shape == (80, 1025) and basis.dtype == np.float64
    nz = np.nonzero(basis)
    np.testing.assert_array_equal(nz[0], golden["mel_nz_rows"])
    np.testing.assert_array_equal(nz[1], golden["mel_nz_cols"])
    np.testing.assert_array_equal(basis[nz], golden["mel_nz_vals"])
    assert len(nz[0]) == 2002 and (np.count_nonzero(basis, axis=0) <= 2).all()
    ep = np.concatenate([speechlike(30000, 3), np.zeros(30000, np.float32)])
    assert len(ep) == int(golden["endpoint_in_len"])
    assert ao.find_endpoint(ep, hp) == int(golden["endpoint"])
    np.testing.assert_array_equal(ao._amp_to_db(np.array([0.0, 1e-6, 1e-5, 0.5, 1.0, 123.0], np.float32)), golden["db"])
    np.testing.assert_array_equal(ao._db_to_amp(np.array([-100.0, -20.0, 0.0, 20.0, 180.0])), golden["amp"])


def test_truncating_stft_parameters():
    hp = make_hp(sample_rate=22050)
    assert ao._stft_parameters(hp) == (2048, 275, 1102)


# ---------- independent implementations ----------

def _torch_stft(y, n_fft=2048, hop=250, win=1000):
    w = torch.hann_window(win, periodic=True, dtype=torch.float64)
    return torch.stft(torch.from_numpy(np.asarray(y, np.float64)), n_fft, hop, win, w, center=True,
                      pad_mode="reflect", return_complex=True).numpy()


def test_stft_vs_torch():
    y = speechlike(20000, 1).astype(np.float64)
    D = librosa060.stft(y, 2048, 250, 1000)
    assert D.shape == (1025, 1 + 20000 // 250) and D.dtype == np.complex64 and D.flags.f_contiguous
    assert ao.rel_l2(D, _torch_stft(y)) < 1e-6


def test_istft_vs_torch_and_roundtrip():
    y = speechlike(20000, 2).astype(np.float64)
    D = _torch_stft(y)
    yi = librosa060.istft(D.astype(np.complex64), 250, 1000)
    assert yi.dtype == np.float32 and len(yi) == 250 * (D.shape[1] - 1)
    w = torch.hann_window(1000, periodic=True, dtype=torch.float64)
    yt = torch.istft(torch.from_numpy(D), 2048, 250, 1000, w, center=True).numpy()
    assert ao.rel_l2(yi, yt[:len(yi)]) < 1e-6
    assert ao.rel_l2(yi, y[:len(yi)]) < 1e-6


def test_mel_vs_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    fb = torchaudio.functional.melscale_fbanks(1025, 0.0, 10000.0, 80, 20000, norm="slaney", mel_scale="slaney")
    basis = librosa060.mel(20000, 2048, n_mels=80)
    assert ao.rel_l2(basis, fb.numpy().T.astype(np.float64)) < 1e-5


def test_lfilter_restated():
    hp = make_hp()
    x = speechlike(5000, 4)
    y = ao.preemphasis(x, hp)
    ref = x.astype(np.float64).copy()
    ref[1:] -= 0.97 * x[:-1].astype(np.float64)
    np.testing.assert_allclose(y, ref, rtol=0, atol=1e-15)
    z = ao.inv_preemphasis(y, hp)
    np.testing.assert_allclose(z, x.astype(np.float64), atol=1e-9)
    assert y.dtype == np.float64 and z.dtype == np.float64


# ---------- analytic known answers ----------

def test_window_sum_is_1p5_in_interior():
    w = librosa060.pad_center(librosa060.get_window("hann", 1000), 2048)
    acc = np.zeros(2048 + 250 * 40)
    for k in range(41):
        acc[k * 250:k * 250 + 2048] += w * w
    assert np.allclose(acc[2048:-2048], 1.5, atol=1e-12)


def test_bin_centred_sinusoid_peak():
    n = np.arange(20000)
    k0 = 100
    y = 0.7 * np.cos(2 * np.pi * k0 * n / 2048)
    D = np.abs(librosa060.stft(y, 2048, 250, 1000))
    mid = D[:, 20:60]
    assert (np.argmax(mid, axis=0) == k0).all()
    assert np.allclose(mid[k0], 0.7 * 500.0 / 2, rtol=1e-5)  # A * sum(hann)/2, sum(hann_1000) = 500


def test_impulse_flat_spectrum():
    y = np.zeros(6000)
    y[3000] = 1.0
    D = np.abs(librosa060.stft(y, 2048, 250, 1000))
    w = librosa060.get_window("hann", 1000)
    # frame 12 is centred on sample 3000 -> window value at its centre
    assert np.allclose(D[:, 12], w[500], atol=1e-6)


def test_invalid_audio_raises():
    y = np.zeros(4000)
    y[10] = np.nan
    with pytest.raises(librosa060.ParameterError):
        librosa060.stft(y, 2048, 250, 1000)
    with pytest.raises(librosa060.ParameterError):
        librosa060.stft(np.zeros((2, 4000)), 2048, 250, 1000)


def test_griffin_lim_fp32_vs_fp64_margin():
    """GL is benign to precision: a float32-pipeline rerun stays far above the 40 dB bar."""
    hp = make_hp(min_level_db=-100)
    wav = speechlike(5000, 5)
    S = ao._db_to_amp(ao._denormalize(ao.spectrogram(wav, hp), hp) + hp.ref_level_db) ** hp.power
    rs = np.random.RandomState(0)
    angles = np.exp(2j * np.pi * rs.rand(*S.shape))
    y1 = ao._griffin_lim(S, hp, angles=angles, iters=10)
    y2 = ao._griffin_lim(S.astype(np.float32), hp, angles=angles.astype(np.complex64), iters=10)
    assert ao.snr_db(y2, y1) > 60


def test_tf_signal_restatement_vs_torch():
    """tf.contrib.signal.stft / inverse_stft (TF 1.7) restated in oracle/tf_signal17.py, against torch.stft(center=False)
    with the window placed on the first win samples of the n_fft frame (tf zero-pads at the END)."""
    from oracle import tf_signal17 as tfo
    hp = make_hp()
    y = speechlike(6100, 3)
    D = tfo._stft_tensorflow(y, hp)
    assert D.shape == (1 + (6100 - 1000) // 250, 1025) and D.dtype == np.complex64
    w = torch.zeros(2048, dtype=torch.float64)
    w[:1000] = torch.hann_window(1000, periodic=True, dtype=torch.float64)
    yp = torch.cat([torch.from_numpy(y.astype(np.float64)), torch.zeros(2048 - 1000, dtype=torch.float64)])
    Dt = torch.stft(yp, 2048, hop_length=250, win_length=2048, window=w, center=False, return_complex=True).numpy().T
    assert ao.rel_l2(D, Dt[:D.shape[0]]) < 1e-6
    # inverse: un-normalised windowed overlap-add; with hop = win/4 the Hann^2 sum is 1.5 in the interior
    x = tfo._istft_tensorflow(D, hp)
    assert x.shape == (1000 + 250 * (D.shape[0] - 1),) and x.dtype == np.float32
    assert ao.rel_l2(x[750:-750], 1.5 * y[750:x.size - 750]) < 1e-5
    # zero phase start + est/max(1e-8,|est|): an all-zero spectrogram stays silent
    assert np.all(tfo._griffin_lim_tensorflow(np.zeros((4, 1025), np.float32), hp, iters=2) == 0)


def test_trimming_oracle_matches_reference_process(golden_process):
    """oracle/process_oracle.py against the fixtures written by the reference's own datasets/process.py
    (tests/golden/make_golden_process.py)."""
    from conftest import trim_signals
    from oracle import process_oracle as po
    for name, wav in trim_signals().items():
        t = po.trim_wav(wav)
        start = (t.__array_interface__["data"][0] - wav.__array_interface__["data"][0]) // wav.itemsize if t.size else 0
        assert (start, t.size) == tuple(golden_process[name + "_trim_wav"])
        for thr in (0.01, 0.1):
            s_ = po.trim_silence(wav, thr)
            start = (s_.__array_interface__["data"][0] - wav.__array_interface__["data"][0]) // wav.itemsize if s_.size else 0
            assert (start, s_.size) == tuple(golden_process["%s_trim_silence_%g" % (name, thr)])
    # rmse against a direct definition
    y = trim_signals()["loud"].astype(np.float64)
    e = po.rmse(y, 1024, 512)[0]
    yp = np.pad(y, 512, mode="reflect")
    assert abs(e[3] - np.sqrt(np.mean(yp[3 * 512:3 * 512 + 1024] ** 2))) < 1e-12
