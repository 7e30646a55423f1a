"""Generate tests/golden/*.npz by running the REFERENCE's own audio.py, unmodified.

Run in the build container only (``python tests/golden/make_golden.py``); /root/reference does not
exist on the GPU box, which is why the outputs are committed as fixtures.

What is and is not the reference here:
* ``neural_speech/utils/audio.py`` is imported from /root/reference and executed as is - its
  composition (preemphasis -> _stft -> abs -> dB -> normalise; _denormalize -> _db_to_amp -> **power
  -> _griffin_lim -> inv_preemphasis) is what these fixtures pin;
* its third-party imports are absent in this image, so they are shimmed: ``librosa`` by
  ``oracle/librosa060.py`` (a restatement of librosa 0.6.0), ``tensorflow`` and ``matplotlib`` by
  empty stubs (the TF twin is never called), ``np.complex`` by ``complex`` (removed in NumPy >= 1.24),
  and ``get_hparams()`` by a namespace filled from the reference's own ``hparams/audio.yaml`` through
  ``yaml.safe_load`` (the reference's ``yaml.load(f)`` raises on PyYAML 6).
"""
import os
import sys
import types

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import librosa060  # noqa: E402


def import_reference_audio():
    if not hasattr(np, "complex"):
        np.complex = complex  # audio.py:82
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    sys.modules["matplotlib"] = mpl
    tf = types.ModuleType("tensorflow")
    tf.contrib = types.SimpleNamespace(training=types.SimpleNamespace(HParams=lambda **kw: types.SimpleNamespace(**kw)))
    sys.modules["tensorflow"] = tf
    lr = types.ModuleType("librosa")
    lr.stft = librosa060.stft
    lr.istft = librosa060.istft
    lr.core = types.SimpleNamespace(load=None, stft=librosa060.stft, istft=librosa060.istft)
    lr.output = types.SimpleNamespace(write_wav=None)
    flt = types.ModuleType("librosa.filters")
    flt.mel = librosa060.mel
    lr.filters = flt
    sys.modules["librosa"] = lr
    sys.modules["librosa.filters"] = flt
    sys.path.insert(0, REF)
    import neural_speech.hparams as ref_hparams
    from neural_speech.utils import audio as ref_audio
    return ref_hparams, ref_audio


def ref_hparams_namespace(**overrides):
    with open(os.path.join(REF, "neural_speech/hparams/audio.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg.update(overrides)
    return types.SimpleNamespace(**cfg)


def test_signal(n, seed, sr=20000):
    """speech-like harmonic stack + noise, peak 0.9 (SURVEY.md section 8d config 2 content)."""
    rng = np.random.RandomState(seed)
    t = np.arange(n) / sr
    f0 = rng.uniform(100, 250)
    x = np.zeros(n)
    for k in range(1, 31):
        x += np.sin(2 * np.pi * f0 * k * t + rng.uniform(0, 2 * np.pi)) / k
    x *= 0.6 + 0.4 * np.sin(2 * np.pi * 3 * t)
    x += 10 ** (-50 / 20) * rng.randn(n)
    x *= 0.9 / np.max(np.abs(x))
    return x.astype(np.float32)


def main():
    ref_hparams, ref_audio = import_reference_audio()
    out = {}
    for tag, overrides in (("yaml", {}), ("neg", {"min_level_db": -100})):
        hp = ref_hparams_namespace(**overrides)
        ref_hparams._hparams = hp
        ref_audio._mel_basis = None
        wav = test_signal(9000, seed=7)
        out[tag + "_wav"] = wav
        out[tag + "_pre"] = ref_audio.preemphasis(wav)
        D = ref_audio._stft(ref_audio.preemphasis(wav))
        out[tag + "_stft"] = np.ascontiguousarray(D)
        out[tag + "_spec"] = np.ascontiguousarray(ref_audio.spectrogram(wav))
        out[tag + "_mel"] = np.ascontiguousarray(ref_audio.melspectrogram(wav))
        out[tag + "_istft"] = ref_audio._istft(D)
        # Griffin-Lim with the reference's own RNG draw (audio.py:81), 4 iterations to keep it small
        hp.griffin_lim_iters = 4
        rng = np.random.RandomState(11)
        S_in = rng.rand(hp.num_freq, 12).astype(np.float32)
        out[tag + "_gl_in"] = S_in
        np.random.seed(1234)
        out[tag + "_gl_wav"] = ref_audio.inv_spectrogram(S_in)
        np.random.seed(1234)
        S_lin = (rng.rand(hp.num_freq, 12) * 3.0)
        out[tag + "_gl_S"] = S_lin
        out[tag + "_gl_raw"] = ref_audio._griffin_lim(S_lin)
    hp = ref_hparams_namespace()
    ref_hparams._hparams = hp
    out["stft_parameters"] = np.array(ref_audio._stft_parameters())
    ref_audio._mel_basis = None
    basis = ref_audio._build_mel_basis()
    nz = np.nonzero(basis)
    out["mel_nz_rows"] = nz[0].astype(np.int16)
    out["mel_nz_cols"] = nz[1].astype(np.int16)
    out["mel_nz_vals"] = basis[nz]
    ep = np.concatenate([test_signal(30000, seed=3), np.zeros(30000, np.float32)])
    out["endpoint_in_len"] = np.array(len(ep))
    out["endpoint"] = np.array(ref_audio.find_endpoint(ep))
    out["db"] = ref_audio._amp_to_db(np.array([0.0, 1e-6, 1e-5, 0.5, 1.0, 123.0], dtype=np.float32))
    out["amp"] = ref_audio._db_to_amp(np.array([-100.0, -20.0, 0.0, 20.0, 180.0]))
    path = os.path.join(HERE, "reference_audio.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
