"""CPU restatement of the batch assembly in ``neural_speech/datasets/datafeeder.py`` (bucketing 130-147, ``_prepare_batch``
190-199, ``_prepare_inputs`` / ``_prepare_targets`` / ``_pad_input`` / ``_pad_target`` / ``_round_up`` 202-221).
TEST INFRASTRUCTURE (see oracle/__init__.py).  The random shuffles of the reference (``random.shuffle(batches)``,
``random.shuffle(batch)``) are left to the caller: the functions return the deterministic part."""
import numpy as np

_pad = 0          # datafeeder.py:17


def _round_up(x, multiple):
    # datafeeder.py:219-221
    remainder = x % multiple
    return x if remainder == 0 else x + multiple - remainder


def _pad_input(x, length):
    # datafeeder.py:211-212
    return np.pad(x, (0, length - x.shape[0]), mode='constant', constant_values=_pad)


def _pad_target(t, length):
    # datafeeder.py:215-216
    return np.pad(t, [(0, length - t.shape[0]), (0, 0)], mode='constant', constant_values=_pad)


def _prepare_inputs(inputs):
    # datafeeder.py:202-204
    max_len = max((len(x) for x in inputs))
    return np.stack([_pad_input(x, max_len) for x in inputs])


def _prepare_targets(targets, alignment):
    # datafeeder.py:206-208
    max_len = max((len(t) for t in targets)) + 1
    return np.stack([_pad_target(t, _round_up(max_len, alignment)) for t in targets])


def bucket(examples, n):
    # datafeeder.py:143-146: examples.sort(key=lambda x: x[-1]); batches = [examples[i:i + n] ...]   (before random.shuffle(batches))
    examples = list(examples)
    examples.sort(key=lambda x: x[-1])
    return [examples[i:i + n] for i in range(0, len(examples), n)]
