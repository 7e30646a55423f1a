"""Batched / multi-GPU front end of the hot path (not in the reference, which is one-utterance-at-a-time
numpy; these are the calls the BASELINE configs 2-5 time).

* ``features_batch(wavs)``            - spectrogram + melspectrogram for a ragged list of clips, one launch
                                        (the per-utterance work of ``datasets/process.py:23-36``)
* ``inv_spectrogram_batch(specs)``    - Griffin-Lim inversion of a ragged list (or a uniform [N,T,F] array
                                        as produced by ``models/tacotron.py:98``), one pipeline
* ``shard_by_frames`` / ``run_sharded`` - utterance sharding across the GPUs of one box: independent units,
                                        no collective; one host thread per GPU
"""
import threading

import numpy as np

from . import _lib, audio


def _frames_major(spec, num_freq):
    """[F,T] (any order) or [T,F] C-contiguous -> frame-major [T,F] float32 view/copy."""
    spec = np.asarray(spec)
    if spec.ndim != 2:
        raise ValueError("expected 2-D spectrogram, got %r" % (spec.shape,))
    if spec.shape[0] == num_freq and spec.shape[1] != num_freq:
        spec = spec.T
    elif spec.shape[1] != num_freq:
        raise ValueError("no axis of %r has num_freq=%d bins" % (spec.shape, num_freq))
    return np.ascontiguousarray(spec, dtype=np.float32)


def features_batch(wavs, device=None, want_linear=True, want_mel=True):
    """-> list of (linear [F,T] float32, mel [M,T] float32) per clip (None for a feature not requested)."""
    h = audio._handle(device)
    wavs = [audio._as_wav(w) for w in wavs]
    ns = [w.size for w in wavs]
    Ts = [h.num_frames(n) for n in ns]
    packed = np.concatenate(wavs) if len(wavs) > 1 else wavs[0]
    pool = h.lib.pinned_pool()             # results in pooled page-locked memory: the copy out runs at PCIe speed
    lin = pool.empty((sum(Ts), h.num_freq), np.float32) if want_linear else None
    mel = pool.empty((sum(Ts), h.num_mels), np.float32) if want_mel else None
    h.features(packed, ns, lin, mel)
    out, off = [], 0
    for T in Ts:
        out.append((lin[off:off + T].T if want_linear else None, mel[off:off + T].T if want_mel else None))
        off += T
    return out


def _round_up(x, multiple):
    # reference datasets/datafeeder.py:219-221
    remainder = x % multiple
    return x if remainder == 0 else x + multiple - remainder


def feeder_targets(wavs, outputs_per_step, device=None):
    """The feeder's target tensors straight from the waveforms (reference datasets/datafeeder.py:190-216): what
    ``_prepare_targets([spectrogram(w).T ...], r)`` and ``_prepare_targets([melspectrogram(w).T ...], r)`` build with one
    numpy pad per utterance - time-major, zero-padded to ``round_up(max frames + 1, outputs_per_step)`` rows, stacked.
    Returns (mel_targets [N, Tpad, num_mels], linear_targets [N, Tpad, num_freq], n_frames list)."""
    h = audio._handle(device)
    wavs = [audio._as_wav(w) for w in wavs]
    ns = [w.size for w in wavs]
    Ts = [h.num_frames(n) for n in ns]
    rows = _round_up(max(Ts) + 1, outputs_per_step)
    packed = np.concatenate(wavs) if len(wavs) > 1 else wavs[0]
    pool = h.lib.pinned_pool()
    lin = pool.empty((len(wavs), rows, h.num_freq), np.float32)
    mel = pool.empty((len(wavs), rows, h.num_mels), np.float32)
    h.features_padded(packed, ns, rows, lin, mel)
    return mel, lin, Ts


def inv_spectrogram_batch(specs, init_phase=None, seed=0, iters=None, device=None, denormalize=True, deemphasis=True,
                          out=None):
    """Griffin-Lim inversion of a batch.

    specs: list of [F,T_b] / [T_b,F] arrays, or one [N,T,F] array.  init_phase: same structure (complex) or
    None -> device Philox keyed by ``seed``.  Returns a list of float64 waveforms (views into one buffer),
    or float32 when ``deemphasis`` is False (what ``_griffin_lim`` returns).
    """
    h = audio._handle(device)
    F = h.num_freq
    if isinstance(specs, np.ndarray) and specs.ndim == 3:
        if specs.shape[2] != F:
            raise ValueError("uniform batch must be [N,T,%d]" % F)
        packed = np.ascontiguousarray(specs, dtype=np.float32)
        Ts = [specs.shape[1]] * specs.shape[0]
        if init_phase is not None:
            init_phase = np.ascontiguousarray(init_phase, dtype=np.complex64)
    else:
        mats = [_frames_major(s, F) for s in specs]
        Ts = [m.shape[0] for m in mats]
        packed = np.concatenate(mats) if len(mats) > 1 else mats[0]
        if init_phase is not None:
            ph = []
            for p, T in zip(init_phase, Ts):
                p = np.asarray(p)
                if p.shape == (F, T) and T != F:
                    p = p.T
                ph.append(np.ascontiguousarray(p, dtype=np.complex64))
            init_phase = np.concatenate(ph) if len(ph) > 1 else ph[0]
    ns = [h.num_samples(T) for T in Ts]
    dt = np.float64 if deemphasis else np.float32
    if out is None:
        # results land in pooled page-locked memory (PinnedPool): the copy out runs at PCIe speed
        nbytes = sum(ns) * np.dtype(dt).itemsize
        out = h.lib.pinned_pool().empty((sum(ns),), dt) if nbytes >= (1 << 20) else np.empty(sum(ns), dtype=dt)
    flags = (_lib.GL_DENORMALIZE if denormalize else 0) | (_lib.GL_DEEMPHASIS if deemphasis else 0)
    h.griffin_lim(packed, _lib.FRAME_MAJOR, Ts, out, init_phase=init_phase, seed=seed, iters=-1 if iters is None else iters,
                  flags=flags, out_dtype=_lib.F64 if deemphasis else _lib.F32)
    res, off = [], 0
    for n in ns:
        res.append(out[off:off + n])
        off += n
    return res


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_host_to_gpu(device=0):
    """Pin the calling host thread to the CPUs next to ``device`` (its PCIe root's NUMA node, from sysfs), so that the
    page-locked buffers it allocates afterwards and the copies it drives stay on the GPU's socket.  With 8 GPUs on two sockets
    every rank / shard thread otherwise stages through node 0.  Returns the CPU list, or None when the topology is unknown
    (single node, container without sysfs) - then nothing is changed."""
    import os
    try:
        bus = _lib.default_lib().device_pci_bus_id(device).lower()
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bus) as f:
            cpus = _parse_cpulist(f.read()) & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except (OSError, ValueError, RuntimeError, AttributeError):
        return None


def shard_by_frames(n_frames, world_size):
    """Greedy longest-first assignment of utterances to ``world_size`` shards, balanced by frame count.
    Deterministic; returns a list (per shard) of utterance indices in ascending order."""
    order = sorted(range(len(n_frames)), key=lambda i: (-int(n_frames[i]), i))
    loads = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (loads[j], j))
        shards[r].append(i)
        loads[r] += int(n_frames[i])
    return [sorted(s) for s in shards]


def run_sharded(fn, items, lengths, devices):
    """Run ``fn(sub_items, device)`` on each device's shard in its own host thread (ctypes releases the GIL)
    and put the per-item results back in input order.  No cross-device communication."""
    shards = shard_by_frames(lengths, len(devices))
    results = [None] * len(items)
    errors = []

    def work(rank):
        try:
            idx = shards[rank]
            if not idx:
                return
            if len(devices) > 1 and isinstance(devices[rank], int):
                bind_host_to_gpu(devices[rank])        # this thread's staging buffers and copies on the GPU's socket
            outs = fn([items[i] for i in idx], devices[rank])
            for i, o in zip(idx, outs):
                results[i] = o
        except Exception as e:  # surface in the caller's thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(r,)) for r in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def inv_spectrogram_multi_gpu(specs, devices, **kw):
    h = audio._handle(devices[0])
    F = h.num_freq
    lengths = [s.shape[1] if s.shape[0] == F and s.shape[1] != F else s.shape[0] for s in specs]
    return run_sharded(lambda sub, dev: inv_spectrogram_batch(sub, device=dev, **kw), list(specs), lengths, devices)
