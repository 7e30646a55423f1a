"""hparams shim: the global, yaml-driven config the audio hot path reads on every call.

Mirrors the reference's ``neural_speech/hparams/__init__.py:1-26`` (``load``, ``get_hparams``,
``debug_string``, module-global ``yaml_path``) without TensorFlow: the reference wraps the merged
yaml dict in ``tf.contrib.training.HParams``; here :class:`HParams` offers the subset of that class
the repo uses (attribute access, ``values()``, ``parse("k=v,...")`` as used by ``train.py:162-163`` /
``eval.py:73-74``).

Differences that are deliberate:
* ``yaml.safe_load`` (the reference's bare ``yaml.load(f)`` raises on PyYAML >= 6);
* ``yaml_path`` defaults to this package's directory instead of a cwd-relative string, and
  ``train.yaml`` / ``<model>.yaml`` are optional (the hot path only needs the audio keys);
* ``get_hparams()`` raises a clear ``RuntimeError`` when ``load()`` was never called, where the
  reference returns ``None`` and the caller dies with ``AttributeError``.
"""
import os
import threading

import yaml

yaml_path = os.path.dirname(os.path.abspath(__file__)) + os.sep
_hparams = None
_lock = threading.Lock()

# the keys the hot path reads
AUDIO_KEYS = ("num_freq", "num_mels", "sample_rate", "frame_shift_ms", "frame_length_ms",
              "preemphasis", "ref_level_db", "min_level_db", "power", "griffin_lim_iters")


class HParams(object):
    """Minimal stand-in for ``tf.contrib.training.HParams``."""

    def __init__(self, **kwargs):
        object.__setattr__(self, "_values", {})
        object.__setattr__(self, "_version", 0)
        for k, v in kwargs.items():
            self.add_hparam(k, v)

    def add_hparam(self, name, value):
        if name in self._values:
            raise ValueError("Hyperparameter name is reserved: %s" % name)
        self._values[name] = value
        object.__setattr__(self, "_version", self._version + 1)

    def set_hparam(self, name, value):
        if name not in self._values:
            raise KeyError(name)
        old = self._values[name]
        self._values[name] = _coerce(value, old, name)
        object.__setattr__(self, "_version", self._version + 1)

    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, "_values")[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in self._values:
            self.set_hparam(name, value)
        else:
            self.add_hparam(name, value)

    def __contains__(self, name):
        return name in self._values

    def values(self):
        return dict(self._values)

    def parse(self, values):
        """Override from a ``"name=value,name2=value2"`` string; unknown names raise ValueError."""
        if not values:
            return self
        for item in _split_top_level(values):
            if not item.strip():
                continue
            if "=" not in item:
                raise ValueError("Could not parse hparam override %r" % item)
            name, value = item.split("=", 1)
            name = name.strip()
            if name not in self._values:
                raise ValueError("Unknown hyperparameter: %s" % name)
            self.set_hparam(name, value.strip())
        return self

    def audio_key(self):
        """Tuple of the audio values; the native handle cache is keyed on it so a changed hparam
        can never hit a stale plan (the reference's ``_mel_basis`` cache is never invalidated,
        ``utils/audio.py:135-142``)."""
        return tuple(self._values.get(k) for k in AUDIO_KEYS)

    def __repr__(self):
        return "HParams(%s)" % ", ".join("%s=%r" % kv for kv in sorted(self._values.items()))


def _split_top_level(s):
    out, depth, cur = [], 0, []
    for ch in s:
        if ch == "[":
            depth += 1
        elif ch == "]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur))
    return out


def _coerce(value, old, name):
    if not isinstance(value, str):
        return value
    if isinstance(old, bool):
        if value.lower() in ("true", "1"):
            return True
        if value.lower() in ("false", "0"):
            return False
        raise ValueError("Could not parse bool for %s: %r" % (name, value))
    if isinstance(old, int):
        try:
            return int(value)
        except ValueError:
            f = float(value)
            if f != int(f):
                raise ValueError("Could not parse int for %s: %r" % (name, value))
            return int(f)
    if isinstance(old, float):
        return float(value)
    if isinstance(old, (list, tuple)):
        inner = value.strip()
        if inner.startswith("[") and inner.endswith("]"):
            inner = inner[1:-1]
        elem = old[0] if len(old) else ""
        return [_coerce(v.strip(), elem, name) for v in inner.split(",") if v.strip()]
    return value


def debug_string(hp):
    values = hp.values()
    hp = ['  %s: %s' % (name, values[name]) for name in sorted(values)]
    return 'Hyperparameters:\n' + '\n'.join(hp)


def _read(path, required):
    if not os.path.exists(path):
        if required:
            raise FileNotFoundError(path)
        return {}
    with open(path) as f:
        return yaml.safe_load(f) or {}


def load(model_type=None, path=None):
    """Merge ``audio.yaml`` + ``train.yaml`` + ``<model_type>.yaml`` (later files win, as in the
    reference ``hparams/__init__.py:14-22``) and install the result as the global hparams."""
    global _hparams
    base = yaml_path if path is None else path
    config = _read(os.path.join(base, "audio.yaml"), required=True)
    config.update(_read(os.path.join(base, "train.yaml"), required=False))
    if model_type is not None:
        config.update(_read(os.path.join(base, model_type + ".yaml"), required=False))
    with _lock:
        _hparams = HParams(**config)
    return _hparams


def set_hparams(hp):
    """Install an already-built HParams (tests, embedding applications)."""
    global _hparams
    with _lock:
        _hparams = hp
    return hp


def get_hparams():
    if _hparams is None:
        raise RuntimeError("hparams not loaded: call nspeech_b200.hparams.load() first")
    return _hparams
