"""CPU restatement of the silence trimming in ``neural_speech/datasets/process.py:39-68`` and of the librosa 0.6.0 calls it
makes (``librosa.effects.split``, ``librosa.feature.rmse``, ``core.power_to_db``, ``core.frames_to_samples``).
TEST INFRASTRUCTURE (see oracle/__init__.py); librosa is absent here: "parity unpinned" against a real librosa 0.6.0 run."""
import numpy as np


def frame(y, frame_length, hop_length):
    # librosa.util.frame: [frame_length, n_frames] strided view
    n_frames = 1 + (len(y) - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n_frames)[None, :]
    return y[idx]


def rmse(y, frame_length=2048, hop_length=512, center=True, pad_mode="reflect"):
    # librosa.feature.rmse (0.6.0), time-domain branch
    y = np.asarray(y)
    if center:
        y = np.pad(y, int(frame_length // 2), mode=pad_mode)
    x = frame(y, frame_length, hop_length)
    return np.sqrt(np.mean(np.abs(x) ** 2, axis=0, keepdims=True))


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    magnitude = np.abs(np.asarray(S))
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def frames_to_samples(frames, hop_length=512):
    return (np.asanyarray(frames) * hop_length).astype(int)


def split(y, top_db=60, ref=np.max, frame_length=2048, hop_length=512):
    # librosa.effects.split (0.6.0)
    mse = rmse(y, frame_length=frame_length, hop_length=hop_length) ** 2
    non_silent = power_to_db(mse.squeeze(), ref=ref, top_db=None) > -top_db
    edges = np.flatnonzero(np.diff(non_silent.astype(int)))
    edges = [edges + 1]
    if non_silent[0]:
        edges.insert(0, [0])
    if non_silent[-1]:
        edges.append([len(non_silent)])
    edges = frames_to_samples(np.concatenate(edges), hop_length=hop_length)
    edges = np.minimum(edges, y.shape[-1])
    return edges.reshape((-1, 2))


def _find_start(splits, min_samples=2000):
    # datasets/process.py:57-61
    for split_start, split_end in splits:
        if split_end - split_start > min_samples:
            return max(0, split_start - min_samples)
    return 0


def _find_end(splits, num_samples, min_samples=2000):
    # datasets/process.py:64-68
    for split_start, split_end in reversed(splits):
        if split_end - split_start > min_samples:
            return min(num_samples, split_end + min_samples)
    return num_samples


def trim_wav(wav, threshold_db=25):
    # datasets/process.py:39-42
    splits = split(wav, threshold_db, frame_length=1024, hop_length=512)
    return wav[_find_start(splits):_find_end(splits, len(wav))]


def trim_silence(wav, threshold, frame_length=2048):
    # datasets/process.py:45-54
    if wav.size < frame_length:
        frame_length = wav.size
    energy = rmse(wav, frame_length=frame_length)
    frames = np.nonzero(energy > threshold)
    indices = frames_to_samples(frames)[1]
    return wav[indices[0]:indices[-1]] if indices.size else wav[:0]
