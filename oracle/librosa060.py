"""Restatement of the librosa==0.6.0 routines the reference's audio path calls.

TEST INFRASTRUCTURE (see oracle/__init__.py). librosa is pinned in the reference's
``requirements.txt:6`` and is not vendored under /root/reference nor installed here, so this file
restates its *published* algorithms (librosa 0.6.0: ``core/spectrum.py`` stft/istft,
``filters.py`` mel, ``core/time_frequency.py`` hz_to_mel/mel_to_hz/fft_frequencies/mel_frequencies,
``util/utils.py`` pad_center/frame/tiny/valid_audio).  Call sites in the reference:
``neural_speech/utils/audio.py:108`` (stft), ``:113`` (istft), ``:147`` (filters.mel).

Known residual uncertainty (stated in SURVEY.md section 8c): releases before 0.6 carried a
``.conj()`` on the STFT ("to match phase from DPWE code").  ``CONJ_DPWE`` selects that convention; the
default is the standard e^{-i w n} sign.  |D|, mel, dB features are invariant to it.
"""
import numpy as np
import scipy.fftpack as fft
import scipy.signal

MAX_MEM_BLOCK = 2 ** 8 * 2 ** 10  # librosa.util.MAX_MEM_BLOCK
CONJ_DPWE = False


class ParameterError(Exception):
    """librosa.util.exceptions.ParameterError"""


def valid_audio(y, mono=True):
    if not isinstance(y, np.ndarray):
        raise ParameterError('data must be of type numpy.ndarray')
    if not np.issubdtype(y.dtype, np.floating):
        raise ParameterError('data must be floating-point')
    if mono and y.ndim != 1:
        raise ParameterError('Invalid shape for monophonic audio: ndim={:d}, shape={}'.format(y.ndim, y.shape))
    if not np.isfinite(y).all():
        raise ParameterError('Audio buffer is not finite everywhere')
    return True


def pad_center(data, size):
    n = data.shape[-1]
    lpad = int((size - n) // 2)
    if lpad < 0:
        raise ParameterError('Target size ({:d}) must be at least input size ({:d})'.format(size, n))
    return np.pad(data, [(lpad, int(size - n - lpad))], mode='constant')


def frame(y, frame_length=2048, hop_length=512):
    if len(y) < frame_length:
        raise ParameterError('Buffer is too short (n={:d}) for frame_length={:d}'.format(len(y), frame_length))
    if hop_length < 1:
        raise ParameterError('Invalid hop_length: {:d}'.format(hop_length))
    n_frames = 1 + int((len(y) - frame_length) / hop_length)
    y = np.ascontiguousarray(y)
    return np.lib.stride_tricks.as_strided(y, shape=(frame_length, n_frames),
                                           strides=(y.itemsize, hop_length * y.itemsize))


def tiny(x):
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.floating) or np.issubdtype(x.dtype, np.complexfloating):
        dtype = x.dtype
    else:
        dtype = np.float32
    return np.finfo(dtype).tiny


def get_window(window, Nx, fftbins=True):
    return scipy.signal.get_window(window, Nx, fftbins=fftbins)


def stft(y, n_fft=2048, hop_length=None, win_length=None, window='hann', center=True,
         dtype=np.complex64, pad_mode='reflect'):
    """librosa.core.stft (0.6.0): periodic window padded centrally to n_fft, reflect-padded signal,
    frames every hop, complex FFT of each (real) frame in double, first 1+n_fft/2 bins, stored as
    ``dtype`` in a Fortran-ordered [1+n_fft/2, n_frames] matrix, computed in MAX_MEM_BLOCK column blocks."""
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    fft_window = get_window(window, win_length, fftbins=True)
    fft_window = pad_center(fft_window, n_fft)
    fft_window = fft_window.reshape((-1, 1))
    valid_audio(y)
    if center:
        if pad_mode == 'reflect' and len(y) <= n_fft // 2:
            # numpy 1.14's np.pad would silently wrap the reflection; modern numpy does the same.
            pass
        y = np.pad(y, int(n_fft // 2), mode=pad_mode)
    y_frames = frame(y, frame_length=n_fft, hop_length=hop_length)
    stft_matrix = np.empty((int(1 + n_fft // 2), y_frames.shape[1]), dtype=dtype, order='F')
    n_columns = int(MAX_MEM_BLOCK / (stft_matrix.shape[0] * stft_matrix.itemsize))
    for bl_s in range(0, stft_matrix.shape[1], n_columns):
        bl_t = min(bl_s + n_columns, stft_matrix.shape[1])
        blk = fft.fft(fft_window * y_frames[:, bl_s:bl_t], axis=0)[:stft_matrix.shape[0]]
        stft_matrix[:, bl_s:bl_t] = blk.conj() if CONJ_DPWE else blk
    return stft_matrix


def istft(stft_matrix, hop_length=None, win_length=None, window='hann', center=True,
          dtype=np.float32, length=None):
    """librosa.core.istft (0.6.0): per frame Hermitian-extend, ifft (double), keep the real part,
    multiply by the padded window, overlap-add into a ``dtype`` (float32) buffer, divide by the summed
    squared window where it exceeds tiny(), trim n_fft//2 from both ends."""
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    ifft_window = get_window(window, win_length, fftbins=True)
    ifft_window = pad_center(ifft_window, n_fft)
    n_frames = stft_matrix.shape[1]
    expected_signal_len = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(expected_signal_len, dtype=dtype)
    ifft_window_sum = np.zeros(expected_signal_len, dtype=dtype)
    ifft_window_square = ifft_window * ifft_window
    for i in range(n_frames):
        sample = i * hop_length
        spec = stft_matrix[:, i].flatten()
        if CONJ_DPWE:
            spec = np.concatenate((spec.conj(), spec[-2:0:-1]), 0)
        else:
            spec = np.concatenate((spec, spec[-2:0:-1].conj()), 0)
        ytmp = ifft_window * fft.ifft(spec).real
        y[sample:(sample + n_fft)] = y[sample:(sample + n_fft)] + ytmp
        ifft_window_sum[sample:(sample + n_fft)] += ifft_window_square
    approx_nonzero_indices = ifft_window_sum > tiny(ifft_window_sum)
    y[approx_nonzero_indices] /= ifft_window_sum[approx_nonzero_indices]
    if length is None:
        if center:
            y = y[int(n_fft // 2):-int(n_fft // 2)]
    else:
        start = int(n_fft // 2) if center else 0
        y = y[start:start + length]
    return y


def fft_frequencies(sr=22050, n_fft=2048):
    return np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)


def hz_to_mel(frequencies, htk=False):
    frequencies = np.asanyarray(frequencies, dtype=float)
    if htk:
        return 2595.0 * np.log10(1.0 + frequencies / 700.0)
    f_min = 0.0
    f_sp = 200.0 / 3
    mels = (frequencies - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if frequencies.ndim:
        log_t = (frequencies >= min_log_hz)
        mels[log_t] = min_log_mel + np.log(frequencies[log_t] / min_log_hz) / logstep
    elif frequencies >= min_log_hz:
        mels = min_log_mel + np.log(frequencies / min_log_hz) / logstep
    return mels


def mel_to_hz(mels, htk=False):
    mels = np.asanyarray(mels, dtype=float)
    if htk:
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_min = 0.0
    f_sp = 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = (mels >= min_log_mel)
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_frequencies(n_mels=128, fmin=0.0, fmax=11025.0, htk=False):
    min_mel = hz_to_mel(fmin, htk=htk)
    max_mel = hz_to_mel(fmax, htk=htk)
    mels = np.linspace(min_mel, max_mel, n_mels)
    return mel_to_hz(mels, htk=htk)


def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm=1):
    """librosa.filters.mel (0.6.0): Slaney-scale triangular filters, area-normalised, float64 [n_mels, 1+n_fft/2]."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)))
    fftfreqs = fft_frequencies(sr=sr, n_fft=n_fft)
    mel_f = mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == 1:
        enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    return weights
