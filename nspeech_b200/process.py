"""Silence trimming of the preprocessing path - the mirror of ``neural_speech/datasets/process.py:39-68`` (``trim_wav``,
``trim_silence`` and its two interval searches) and of the two librosa 0.6.0 calls behind them
(``librosa.effects.split``, ``librosa.feature.rmse``).  The per-frame energies (the only pass over the samples) come from
the GPU (``nsb_frame_energy``); the interval logic on the handful of frame values runs on the host."""
import numpy as np

from . import audio


def frame_energy(wav, frame_length, hop_length=512):
    """mean(|x|^2) per centred frame = librosa.feature.rmse(wav, frame_length, hop_length) ** 2, float64 [1 + n // hop]"""
    w = audio._as_wav(wav)
    out = np.empty(1 + w.size // hop_length, dtype=np.float64)
    audio._handle().frame_energy(w, [w.size], frame_length, hop_length, out)
    return out


def _split(wav, top_db, frame_length, hop_length):
    # librosa.effects.split(y, top_db, ref=np.max, frame_length, hop_length) of librosa 0.6.0
    mse = frame_energy(wav, frame_length, hop_length)
    amin = 1e-10
    db = 10.0 * np.log10(np.maximum(amin, mse)) - 10.0 * np.log10(np.maximum(amin, np.max(mse)))
    non_silent = db > -top_db
    edges = [np.flatnonzero(np.diff(non_silent.astype(int))) + 1]
    if non_silent[0]:
        edges.insert(0, [0])
    if non_silent[-1]:
        edges.append([len(non_silent)])
    edges = np.concatenate(edges).astype(np.int64) * hop_length
    edges = np.minimum(edges, len(wav))
    return edges.reshape((-1, 2))


def _trim_bounds(splits, num_samples, min_samples=2000):
    """[start, end) that ``trim_wav`` keeps: from the first to the last non-silent interval longer than ``min_samples``, widened by
    ``min_samples`` on either side and clipped to the clip (what the reference's two interval searches, datasets/process.py:57-68,
    return); the whole clip when no interval is that long."""
    splits = np.asarray(splits, dtype=np.int64).reshape(-1, 2)
    keep = np.flatnonzero(splits[:, 1] - splits[:, 0] > min_samples)
    if keep.size == 0:
        return 0, num_samples
    return max(0, int(splits[keep[0], 0]) - min_samples), min(num_samples, int(splits[keep[-1], 1]) + min_samples)


def trim_wav(wav, threshold_db=25):
    """Silence off both ends of the clip (reference datasets/process.py:39-42): energies of the centred 1024-sample frames every 512
    samples from the GPU, the interval logic of ``librosa.effects.split`` and of the reference on those few values."""
    start, end = _trim_bounds(_split(wav, threshold_db, frame_length=1024, hop_length=512), len(wav))
    return wav[start:end]


def trim_silence(wav, threshold, frame_length=2048):
    """Everything before the first and from the last frame whose RMS exceeds ``threshold`` goes (reference
    datasets/process.py:45-54: ``librosa.feature.rmse`` at hop 512, ``frames_to_samples``); an all-silent clip becomes empty."""
    frame_length = min(frame_length, wav.size)
    loud = np.flatnonzero(np.sqrt(frame_energy(wav, frame_length)) > threshold)
    if loud.size == 0:
        return wav[:0]
    return wav[512 * int(loud[0]):512 * int(loud[-1])]


def process_utterance_arrays(wav):
    """The array part of ``process_utterance`` (reference datasets/process.py:22-36): trim, then both features from one
    STFT.  Returns (wav, spectrogram.T, mel_spectrogram.T, n_frames) - the features time-major as the reference stores them."""
    wav = trim_wav(wav)
    lin, mel = audio.spectrogram_and_mel(wav)
    return wav, lin.T, mel.T, lin.shape[1]
