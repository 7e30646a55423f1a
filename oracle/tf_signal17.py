"""numpy restatement of the TensorFlow-1.7 ``tf.contrib.signal`` ops behind the reference's Griffin-Lim TWIN
(``neural_speech/utils/audio.py:51-58,90-103,116-123,158-159,170-171``), and of that twin itself.

TEST INFRASTRUCTURE (see oracle/__init__.py).  tensorflow-gpu==1.7.0 is pinned in the reference's
``requirements.txt:14`` and absent here, so this restates the published behaviour of
``tf.contrib.signal.stft`` / ``inverse_stft`` / ``frame`` / ``overlap_and_add`` / ``hann_window`` (TF 1.7
``tensorflow/contrib/signal/python/ops/spectral_ops.py``):

* stft(signals, frame_length, frame_step, fft_length, pad_end=False): frames start at k*frame_step (NOT centred,
  no padding), T = 1 + (n - frame_length) // frame_step, each frame times the PERIODIC Hann window of
  frame_length, zero-padded at the END to fft_length, rfft -> complex64 [..., T, fft_length/2+1];
* inverse_stft(stfts, frame_length, frame_step, fft_length) with its TF-1.7 default
  ``window_fn=hann_window(periodic=True)``: irfft to fft_length, keep the first frame_length samples, times the
  window, overlap_and_add -> length frame_length + frame_step*(T-1); NO division by the summed squared window;
* everything in float32 / complex64.

"Parity unpinned" against a real TF 1.7 run; cross-checked against torch.stft(center=False) in tests/test_oracle.py.
"""
import numpy as np


def hann_window(n, periodic=True, dtype=np.float32):
    m = n if periodic else n - 1
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / m)).astype(dtype)


def frame(signals, frame_length, frame_step):
    n = signals.shape[-1]
    T = 1 + (n - frame_length) // frame_step if n >= frame_length else 0
    idx = np.arange(frame_length)[None, :] + frame_step * np.arange(T)[:, None]
    return signals[..., idx]


def stft(signals, frame_length, frame_step, fft_length):
    signals = np.asarray(signals, dtype=np.float32)
    framed = frame(signals, frame_length, frame_step) * hann_window(frame_length)
    return np.fft.rfft(framed.astype(np.float32), n=fft_length, axis=-1).astype(np.complex64)


def overlap_and_add(frames, frame_step):
    T, frame_length = frames.shape[-2:]
    out = np.zeros(frames.shape[:-2] + (frame_length + frame_step * (T - 1),), dtype=frames.dtype)
    for k in range(T):
        out[..., k * frame_step:k * frame_step + frame_length] += frames[..., k, :]
    return out


def inverse_stft(stfts, frame_length, frame_step, fft_length):
    real_frames = np.fft.irfft(np.asarray(stfts, dtype=np.complex64), n=fft_length, axis=-1)[..., :frame_length].astype(np.float32)
    real_frames = real_frames * hann_window(frame_length)
    return overlap_and_add(real_frames, frame_step)


# ---- the reference's TF twin (audio.py), eager: tensors are numpy arrays, [..., T, F] time-major ----

def _stft_parameters(hp):
    n_fft = (hp.num_freq - 1) * 2
    return n_fft, int(hp.frame_shift_ms / 1000 * hp.sample_rate), int(hp.frame_length_ms / 1000 * hp.sample_rate)


def _stft_tensorflow(signals, hp):
    # audio.py:116-118
    n_fft, hop_length, win_length = _stft_parameters(hp)
    return stft(signals, win_length, hop_length, n_fft)


def _istft_tensorflow(stfts, hp):
    # audio.py:121-123
    n_fft, hop_length, win_length = _stft_parameters(hp)
    return inverse_stft(stfts, win_length, hop_length, n_fft)


def _denormalize_tensorflow(S, hp):
    # audio.py:170-171
    return (np.clip(np.asarray(S, np.float32), 0, 1) * np.float32(-hp.min_level_db)) + np.float32(hp.min_level_db)


def _db_to_amp_tensorflow(x):
    # audio.py:158-159 (float32 pow)
    return np.power(np.float32(10.0), np.asarray(x, np.float32) * np.float32(0.05)).astype(np.float32)


def _griffin_lim_tensorflow(S, hp, iters=None):
    # audio.py:90-103
    S = np.asarray(S, np.float32)[None]
    S_complex = S.astype(np.complex64)
    y = _istft_tensorflow(S_complex, hp)
    for _ in range(hp.griffin_lim_iters if iters is None else iters):
        est = _stft_tensorflow(y, hp)
        angles = est / np.maximum(np.float32(1e-8), np.abs(est)).astype(np.complex64)
        y = _istft_tensorflow(S_complex * angles, hp)
    return y[0]


def inv_spectrogram_tensorflow(spectrogram, hp, iters=None):
    # audio.py:51-58 (no inverse pre-emphasis: the caller applies it, synthesizer.py:52)
    S = _db_to_amp_tensorflow(_denormalize_tensorflow(spectrogram, hp) + np.float32(hp.ref_level_db))
    return _griffin_lim_tensorflow(np.power(S, np.float32(hp.power)), hp, iters=iters)
