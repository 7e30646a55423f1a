"""Generate tests/golden/reference_process.npz by running the REFERENCE's own datasets/process.py (trim_wav, trim_silence),
unmodified, with librosa shimmed by the restatement in oracle/ (librosa is absent here; see make_golden.py for the shims).
Run in the build container only: python tests/golden/make_golden_process.py"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden  # noqa: E402
from oracle import process_oracle as po  # noqa: E402


def signals():
    """speech-like clips with quiet ends / pauses (seeded; the tests rebuild them from the same code)"""
    sr = 20000
    out = {}
    a = np.concatenate([0.002 * make_golden.test_signal(9000, 1), make_golden.test_signal(30000, 2), 0.001 * make_golden.test_signal(12000, 3)])
    out["quiet_ends"] = a.astype(np.float32)
    b = np.concatenate([np.zeros(5000, np.float32), make_golden.test_signal(15000, 4), np.zeros(7000, np.float32),
                        make_golden.test_signal(800, 5), np.zeros(6000, np.float32), make_golden.test_signal(14000, 6), np.zeros(3000, np.float32)])
    out["pauses"] = b.astype(np.float32)
    out["loud"] = make_golden.test_signal(20000, 7)
    out["short"] = make_golden.test_signal(1500, 8)
    out["silence"] = np.zeros(8000, np.float32)
    return out


def main():
    ref_hparams, ref_audio = make_golden.import_reference_audio()
    import librosa  # the shim module registered by import_reference_audio
    librosa.effects = types.SimpleNamespace(split=po.split)
    librosa.feature = types.SimpleNamespace(rmse=po.rmse)
    librosa.core.frames_to_samples = po.frames_to_samples
    from neural_speech.datasets import process as ref_process
    out = {}
    for name, wav in signals().items():
        t = ref_process.trim_wav(wav)
        # the trimmed clip is a view: record where it starts and how long it is
        start = (t.__array_interface__["data"][0] - wav.__array_interface__["data"][0]) // wav.itemsize if t.size else 0
        out[name + "_trim_wav"] = np.array([start, t.size])
        for thr in (0.01, 0.1):
            s = ref_process.trim_silence(wav, thr)
            start = (s.__array_interface__["data"][0] - wav.__array_interface__["data"][0]) // wav.itemsize if s.size else 0
            out["%s_trim_silence_%g" % (name, thr)] = np.array([start, s.size])
    path = os.path.join(HERE, "reference_process.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")
    for k in sorted(out):
        print(" ", k, out[k])


if __name__ == "__main__":
    main()
