"""numpy restatement of the reference's ``neural_speech/utils/audio.py`` hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py) - the product never imports this.  Every function
cites the reference lines it follows; the third-party arithmetic is in ``oracle/librosa060.py``.
All functions take the hparams object explicitly (the reference reads a module global on every
call, ``audio.py:14-167``) so tests can run several configurations side by side.

Extra, clearly-marked hooks that the reference does not have:
* ``_griffin_lim(..., angles=...)``: supply the initial phase instead of drawing it from the global
  numpy RNG (``audio.py:81``) so the oracle and the GPU path start from identical phase;
* ``inv_spectrogram(..., angles=...)`` forwards it.
"""
import numpy as np
from scipy import signal

from . import librosa060 as librosa


def _stft_parameters(hp):
    # audio.py:126-130 (truncating int())
    n_fft = (hp.num_freq - 1) * 2
    hop_length = int(hp.frame_shift_ms / 1000 * hp.sample_rate)
    win_length = int(hp.frame_length_ms / 1000 * hp.sample_rate)
    return n_fft, hop_length, win_length


def preemphasis(x, hp):
    # audio.py:31-32
    return signal.lfilter([1, -hp.preemphasis], [1], x)


def inv_preemphasis(x, hp):
    # audio.py:35-36
    return signal.lfilter([1], [1, -hp.preemphasis], x)


def _stft(y, hp):
    # audio.py:106-108
    n_fft, hop_length, win_length = _stft_parameters(hp)
    return librosa.stft(y=y, n_fft=n_fft, hop_length=hop_length, win_length=win_length)


def _istft(y, hp):
    # audio.py:111-113
    _, hop_length, win_length = _stft_parameters(hp)
    return librosa.istft(y, hop_length=hop_length, win_length=win_length)


def _build_mel_basis(hp):
    # audio.py:145-147
    n_fft = (hp.num_freq - 1) * 2
    return librosa.mel(hp.sample_rate, n_fft, n_mels=hp.num_mels)


_mel_cache = {}


def _linear_to_mel(spectrogram, hp):
    # audio.py:138-142 (the reference caches one basis in a module global; keyed here)
    key = (hp.sample_rate, hp.num_freq, hp.num_mels)
    if key not in _mel_cache:
        _mel_cache[key] = _build_mel_basis(hp)
    return np.dot(_mel_cache[key], spectrogram)


def _amp_to_db(x):
    # audio.py:150-151
    return 20 * np.log10(np.maximum(1e-5, x))


def _db_to_amp(x):
    # audio.py:154-155
    return np.power(10.0, x * 0.05)


def _normalize(S, hp):
    # audio.py:162-163
    return np.clip((S - hp.min_level_db) / -hp.min_level_db, 0, 1)


def _denormalize(S, hp):
    # audio.py:166-167
    return (np.clip(S, 0, 1) * -hp.min_level_db) + hp.min_level_db


def spectrogram(y, hp):
    # audio.py:39-42
    D = _stft(preemphasis(y, hp), hp)
    S = _amp_to_db(np.abs(D)) - hp.ref_level_db
    return _normalize(S, hp).astype(np.float32)


def melspectrogram(y, hp):
    # audio.py:61-64
    D = _stft(preemphasis(y, hp), hp)
    S = _amp_to_db(_linear_to_mel(np.abs(D), hp))
    return _normalize(S, hp).astype(np.float32)


def _griffin_lim(S, hp, angles=None, iters=None):
    # audio.py:77-87.  ``angles`` / ``iters`` are test hooks (see module docstring).
    if angles is None:
        angles = np.exp(2j * np.pi * np.random.rand(*S.shape))
    S_complex = np.abs(S).astype(complex)
    y = _istft(S_complex * angles, hp)
    for i in range(hp.griffin_lim_iters if iters is None else iters):
        angles = np.exp(1j * np.angle(_stft(y, hp)))
        y = _istft(S_complex * angles, hp)
    return y


def inv_spectrogram(spectrogram, hp, angles=None, iters=None):
    # audio.py:45-48
    S = _db_to_amp(_denormalize(spectrogram, hp) + hp.ref_level_db)
    return inv_preemphasis(_griffin_lim(S ** hp.power, hp, angles=angles, iters=iters), hp)


def find_endpoint(wav, hp, threshold_db=-40, min_silence_sec=0.8):
    # audio.py:67-74
    window_length = int(hp.sample_rate * min_silence_sec)
    hop_length = int(window_length / 4)
    threshold = _db_to_amp(threshold_db)
    for x in range(hop_length, len(wav) - window_length, hop_length):
        if np.max(wav[x:x + window_length]) < threshold:
            return x + hop_length
    return len(wav)


def save_wav_scaling(wav):
    # audio.py:17-18: the arithmetic of save_wav (in place there; a copy here); the file itself is written by
    # librosa.output.write_wav -> scipy.io.wavfile.write
    wav = np.array(wav, copy=True)
    wav *= 32767 / max(0.01, np.max(np.abs(wav)))
    return wav


# ---- metrics used by the parity tests and the bench (not in the reference) ----

def rel_l2(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def snr_db(test, ref):
    test = np.asarray(test, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    num = np.sum(ref ** 2); den = np.sum((test - ref) ** 2)
    return float(10 * np.log10(num / max(den, 1e-300)))


def spectral_convergence(y, S, hp):
    """|| |STFT(y)| - S ||_F / ||S||_F for a magnitude target S [F,T]."""
    D = np.abs(_stft(np.asarray(y, dtype=np.float64), hp))
    return float(np.linalg.norm(D - S) / np.linalg.norm(S))
