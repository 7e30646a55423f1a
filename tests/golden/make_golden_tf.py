"""Generate tests/golden/reference_tf_twin.npz by running the REFERENCE's own TensorFlow-twin functions
(``neural_speech/utils/audio.py:51-58, 90-103, 116-123, 158-159, 170-171``), unmodified, in the build container
(``python tests/golden/make_golden_tf.py``; /root/reference does not exist on the GPU box).

tensorflow-gpu==1.7.0 (``requirements.txt:14``) is absent here, so the ``tf`` module those functions call is shimmed by an
EAGER numpy stand-in: tensors are numpy arrays, ``tf.contrib.signal.stft`` / ``inverse_stft`` are ``oracle/tf_signal17.py``
(a restatement of the TF 1.7 ops), the element-wise ops are numpy's in float32 / complex64 like TF's.  What the fixtures
pin is therefore the reference's COMPOSITION (zero initial phase, est / max(1e-8, |est|), no de-emphasis, the
denormalise -> dB -> power chain in float32); the signal ops remain "parity unpinned" against a real TF run, exactly like
librosa in make_golden.py.
"""
import contextlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402
from oracle import tf_signal17  # noqa: E402


def eager_tf():
    tf = types.ModuleType("tensorflow")
    tf.complex64, tf.float32 = np.complex64, np.float32
    tf.variable_scope = lambda name: contextlib.nullcontext()
    tf.expand_dims = lambda x, axis: np.expand_dims(np.asarray(x), axis)
    tf.squeeze = lambda x, axis=None: np.squeeze(x, axis)
    tf.identity = lambda x: np.array(x, copy=True)
    tf.cast = lambda x, dtype: np.asarray(x).astype(dtype)
    tf.maximum = lambda a, b: np.maximum(a, b).astype(np.float32)
    tf.abs = lambda x: np.abs(x).astype(np.float32)
    tf.pow = lambda a, b: np.power(np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)).astype(np.float32)
    tf.ones = lambda shape: np.ones(shape, dtype=np.float32)
    tf.shape = lambda x: np.shape(x)
    tf.clip_by_value = lambda x, lo, hi: np.clip(np.asarray(x, dtype=np.float32), lo, hi)
    sig = types.SimpleNamespace(
        stft=lambda signals, frame_length, frame_step, fft_length, pad_end=False: tf_signal17.stft(signals, frame_length, frame_step, fft_length),
        inverse_stft=lambda stfts, frame_length, frame_step, fft_length: tf_signal17.inverse_stft(stfts, frame_length, frame_step, fft_length))
    tf.contrib = types.SimpleNamespace(signal=sig, training=types.SimpleNamespace(HParams=lambda **kw: types.SimpleNamespace(**kw)))
    return tf


def main():
    ref_hparams, ref_audio = mg.import_reference_audio()
    ref_audio.tf = eager_tf()                      # the module-level name the twin's functions resolve
    out = {}
    rng = np.random.RandomState(21)
    for tag, overrides in (("yaml", {}), ("neg", {"min_level_db": -100})):
        hp = mg.ref_hparams_namespace(**overrides)
        hp.griffin_lim_iters = 4
        ref_hparams._hparams = hp
        S = rng.rand(14, hp.num_freq).astype(np.float32)                 # [T, F] time-major, as synthesizer.py:30 passes linear_outputs[0]
        out[tag + "_tw_in"] = S
        out[tag + "_tw_wav"] = np.asarray(ref_audio.inv_spectrogram_tensorflow(S), dtype=np.float32)
        mag = (rng.rand(9, hp.num_freq) * 2.0).astype(np.float32)
        out[tag + "_tw_S"] = mag
        out[tag + "_tw_raw"] = np.asarray(ref_audio._griffin_lim_tensorflow(mag), dtype=np.float32)
    wav = mg.test_signal(5300, seed=5)
    D = ref_audio._stft_tensorflow(wav[None, :])
    out["tw_sig"] = wav
    out["tw_stft"] = np.asarray(D[0], dtype=np.complex64)
    out["tw_istft"] = np.asarray(ref_audio._istft_tensorflow(D)[0], dtype=np.float32)
    path = os.path.join(HERE, "reference_tf_twin.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
