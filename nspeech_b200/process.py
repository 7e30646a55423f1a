"""Silence trimming of the preprocessing path - the mirror of ``neural_speech/datasets/process.py:39-68`` (``trim_wav``,
``trim_silence``, ``_find_start``, ``_find_end``) and of the two librosa 0.6.0 calls behind them
(``librosa.effects.split``, ``librosa.feature.rmse``).  The per-frame energies (the only pass over the samples) come from
the GPU (``nsb_frame_energy``); the interval logic on the handful of frame values is the reference's own, on the host."""
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


def _find_start(splits, min_samples=2000):
    # reference datasets/process.py:57-61
    for split_start, split_end in splits:
        if split_end - split_start > min_samples:
            return max(0, split_start - min_samples)
    return 0


def _find_end(splits, num_samples, min_samples=2000):
    # reference datasets/process.py:64-68
    for split_start, split_end in reversed(splits):
        if split_end - split_start > min_samples:
            return min(num_samples, split_end + min_samples)
    return num_samples


def trim_wav(wav, threshold_db=25):
    '''Trims silence from the ends of the wav (reference datasets/process.py:39-42)'''
    splits = _split(wav, threshold_db, frame_length=1024, hop_length=512)
    return wav[_find_start(splits):_find_end(splits, len(wav))]


def trim_silence(wav, threshold, frame_length=2048):
    '''Removes silence at the beginning and end of a sample (reference datasets/process.py:45-54).'''
    if wav.size < frame_length:
        frame_length = wav.size
    energy = np.sqrt(frame_energy(wav, frame_length))
    indices = np.nonzero(energy > threshold)[0] * 512
    # Note: indices can be an empty array, if the whole audio was silence.
    return wav[indices[0]:indices[-1]] if indices.size else wav[:0]


def process_utterance_arrays(wav):
    """The array part of ``process_utterance`` (reference datasets/process.py:22-36): trim, then both features from one
    STFT.  Returns (wav, spectrogram.T, mel_spectrogram.T, n_frames) - the features time-major as the reference stores them."""
    wav = trim_wav(wav)
    lin, mel = audio.spectrogram_and_mel(wav)
    return wav, lin.T, mel.T, lin.shape[1]
