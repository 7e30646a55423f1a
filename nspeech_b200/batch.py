"""Batched / multi-GPU front end of the hot path (not in the reference, which is one-utterance-at-a-time
numpy; these are the calls the BASELINE configs 2-5 time).

* ``features_batch(wavs)``            - spectrogram + melspectrogram for a ragged list of clips, one launch
                                        (the per-utterance work of ``datasets/process.py:23-36``)
* ``inv_spectrogram_batch(specs)``    - Griffin-Lim inversion of a ragged list (or a uniform [N,T,F] array
                                        as produced by ``models/tacotron.py:98``), one pipeline
* ``shard_by_frames`` / ``run_sharded`` - utterance sharding across the GPUs of one box: independent units,
                                        no collective; one host thread per GPU
"""
import collections
import threading

import numpy as np

from . import _buffers, _lib, audio


def _is_time_major(shape, num_freq, layout):
    """Which axis of a 2-D spectrogram is time.  ``layout``: "FT" = the reference's [num_freq, T] (what ``audio.spectrogram``
    returns and ``audio.inv_spectrogram`` takes), "TF" = time-major [T, num_freq] (Tacotron's outputs), None = decide from the
    shape - refused when both axes have num_freq entries (a 1025-frame clip would otherwise be inverted transposed, silently)."""
    if len(shape) != 2:
        raise ValueError("expected 2-D spectrogram, got %r" % (shape,))
    if layout is not None:
        if layout not in ("FT", "TF"):
            raise ValueError("layout must be 'FT', 'TF' or None, got %r" % (layout,))
        if shape[0 if layout == "FT" else 1] != num_freq:
            raise ValueError("layout %r: shape %r has no %d bins on that axis" % (layout, shape, num_freq))
        return layout == "TF"
    if shape[0] == num_freq and shape[1] == num_freq:
        raise ValueError("shape %r is ambiguous (T == num_freq): pass layout='FT' or 'TF'" % (shape,))
    if shape[0] == num_freq:
        return False
    if shape[1] == num_freq:
        return True
    raise ValueError("no axis of %r has num_freq=%d bins" % (shape, num_freq))


def _frames_major(spec, num_freq, layout=None, dtype=np.float32):
    """[F,T] (any order) or [T,F] -> frame-major [T,F] C-contiguous view/copy."""
    spec = np.asarray(spec)
    if not _is_time_major(spec.shape, num_freq, layout):
        spec = spec.T
    return np.ascontiguousarray(spec, dtype=dtype)


def _pack(h, arrays, dtype):
    """list of arrays -> one packed array; large batches are concatenated straight into pooled page-locked memory (the one host
    copy a list input needs anyway), so that the copy to the device runs at PCIe speed"""
    if len(arrays) == 1:
        return arrays[0]
    total = sum(a.shape[0] for a in arrays)
    shape = (total,) + tuple(arrays[0].shape[1:])
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if nbytes < (4 << 20):
        return np.concatenate(arrays)
    out = h.lib.pinned_pool().empty(shape, dtype)
    np.concatenate(arrays, out=out)
    return out


_copy_pool = None


def _copy_many(out, arrays, offsets):
    """out[offsets[i] : offsets[i] + len(arrays[i])] = arrays[i] on a few host threads (numpy's copies release the GIL; one
    thread's memcpy moves ~10 GB/s, a PCIe 5 link takes 55)."""
    global _copy_pool
    if len(arrays) < 8:
        for a, o in zip(arrays, offsets):
            out[o:o + a.shape[0]] = a
        return
    if _copy_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _copy_pool = ThreadPoolExecutor(max_workers=4, thread_name_prefix="nsb-pack")
    n = len(arrays)
    cuts = [n * i // 4 for i in range(5)]

    def part(i):
        for a, o in zip(arrays[cuts[i]:cuts[i + 1]], offsets[cuts[i]:cuts[i + 1]]):
            out[o:o + a.shape[0]] = a
    for f in [_copy_pool.submit(part, i) for i in range(4)]:
        f.result()


def _group_cuts(sizes, n_groups):
    """cut a list into at most n_groups contiguous runs of about equal total size -> boundaries [0, ..., len(sizes)]"""
    total = float(sum(sizes))
    cuts, acc = [0], 0
    for i, sz in enumerate(sizes):
        acc += sz
        if len(cuts) < n_groups and acc >= total * len(cuts) / n_groups and i + 1 < len(sizes):
            cuts.append(i + 1)
    cuts.append(len(sizes))
    return cuts


def features_batch(wavs, device=None, want_linear=True, want_mel=True, in_flight=3, group_bytes=96 << 20):
    """-> list of (linear [F,T] float32, mel [M,T] float32) per clip (None for a feature not requested).
    Large batches run as a pipeline of clip groups through ``nsb_features_submit`` / ``nsb_wait``: while group g is being
    packed into page-locked memory on the host, group g-1 is copied in and transformed and the results of group g-2 are
    copied out (groups of about ``group_bytes`` of results, at most 16, ``in_flight`` of them submitted at a time;
    ``in_flight=0``: one synchronous call)."""
    h = audio._handle(device)
    wavs = [audio._as_wav(w) for w in wavs]
    if not wavs:
        raise ValueError("empty batch")
    ns = [w.size for w in wavs]
    Ts = [h.num_frames(n) for n in ns]
    pool = h.lib.pinned_pool()             # results in pooled page-locked memory: the copy out runs at PCIe speed
    lin = pool.empty((sum(Ts), h.num_freq), np.float32) if want_linear else None
    mel = pool.empty((sum(Ts), h.num_mels), np.float32) if want_mel else None
    out_bytes = 4 * sum(Ts) * ((h.num_freq if want_linear else 0) + (h.num_mels if want_mel else 0))
    n_groups = min(len(wavs) // 8, int(out_bytes // max(1, group_bytes)), 16)
    if n_groups < 2 or in_flight < 1:
        h.features(_pack(h, wavs, np.float32), ns, lin, mel)
    else:
        soff = np.concatenate([[0], np.cumsum(ns)]).tolist()
        foff = np.concatenate([[0], np.cumsum(Ts)]).tolist()
        packed = pool.empty((soff[-1],), np.float32)
        cuts = _group_cuts(ns, n_groups)
        pending = collections.deque()
        try:
            for g0, g1 in zip(cuts[:-1], cuts[1:]):
                _copy_many(packed, wavs[g0:g1], soff[g0:g1])
                pending.append(h.features_submit(packed[soff[g0]:soff[g1]], ns[g0:g1],
                                                 lin[foff[g0]:foff[g1]] if want_linear else None,
                                                 mel[foff[g0]:foff[g1]] if want_mel else None))
                if len(pending) >= in_flight:
                    h.wait(pending.popleft())
            while pending:
                h.wait(pending.popleft())
        finally:
            while pending:                 # an error above: the calls still in flight hold pointers into these buffers
                try:
                    h.wait(pending.popleft())
                except (_lib.NativeError, _lib.ParameterError):
                    pass
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
    packed = _pack(h, wavs, np.float32)
    pool = h.lib.pinned_pool()
    lin = pool.empty((len(wavs), rows, h.num_freq), np.float32)
    mel = pool.empty((len(wavs), rows, h.num_mels), np.float32)
    h.features_padded(packed, ns, rows, lin, mel)
    return mel, lin, Ts


def _gl_batch_inputs(h, specs, init_phase, layout):
    """-> (packed [sum T, F] float32, n_frames, packed phase or None)"""
    F = h.num_freq
    if isinstance(specs, np.ndarray) and specs.ndim == 3:
        if specs.shape[2] != F:
            raise ValueError("uniform batch must be [N,T,%d]" % F)
        packed = np.ascontiguousarray(specs, dtype=np.float32)
        Ts = [specs.shape[1]] * specs.shape[0]
        if init_phase is not None:
            init_phase = np.ascontiguousarray(init_phase, dtype=np.complex64)
            if init_phase.shape != packed.shape:
                raise ValueError("init_phase shape %r != batch shape %r" % (init_phase.shape, packed.shape))
        return packed, Ts, init_phase
    mats = [_frames_major(s, F, layout) for s in specs]
    if not mats:
        raise ValueError("empty batch")
    Ts = [m.shape[0] for m in mats]
    packed = _pack(h, mats, np.float32)
    if init_phase is not None:
        if len(init_phase) != len(mats):
            raise ValueError("init_phase has %d entries for %d spectrograms" % (len(init_phase), len(mats)))
        ph = []
        for p, s, T in zip(init_phase, specs, Ts):
            if np.shape(p) != np.shape(s):
                raise ValueError("init_phase shape %r != spectrogram shape %r" % (np.shape(p), np.shape(s)))
            ph.append(_frames_major(p, F, layout, dtype=np.complex64))
        init_phase = _pack(h, ph, np.complex64)
    return packed, Ts, init_phase


def _gl_out_buffer(h, ns, dt, out):
    total = sum(ns)
    if out is None:
        # results land in pooled page-locked memory (PinnedPool): the copy out runs at PCIe speed
        nbytes = total * np.dtype(dt).itemsize
        return h.lib.pinned_pool().empty((total,), dt) if nbytes >= (1 << 20) else np.empty(total, dtype=dt)
    if not isinstance(out, np.ndarray) or out.dtype != np.dtype(dt) or not out.flags.c_contiguous or out.size < total:
        raise ValueError("out must be a C-contiguous numpy array of %s with at least %d elements" % (np.dtype(dt), total))
    return out


def _split(out, ns):
    res, off = [], 0
    for n in ns:
        res.append(out[off:off + n])
        off += n
    return res


def _inv_spectrogram_batch_device(specs, init_phase, seed, iters, denormalize, deemphasis):
    """uniform [N, T, F] batch that already lives on the GPU (models/tacotron.py:98): NSB_DEVICE path, device result [N, n]"""
    specs, b, _ = audio._dev_tf_batch(specs, np.float32, "specs")
    if b.ndim != 3:
        raise ValueError("device batches must be one [N, T, F] array")
    h = audio._handle(b.device)
    if b.shape[2] != h.num_freq:
        raise ValueError("uniform batch must be [N,T,%d]" % h.num_freq)
    N, T = b.shape[0], b.shape[1]
    if init_phase is not None:
        init_phase, pb, _ = audio._dev_tf_batch(init_phase, np.complex64, "init_phase")
        if pb.shape != b.shape:
            raise ValueError("init_phase shape %r != batch shape %r" % (pb.shape, b.shape))
    st = _buffers.stream_of(specs, init_phase)
    dt = np.float64 if deemphasis else np.float32
    out = _buffers.empty_like_source(h.lib, (N, h.num_samples(T)), dt, b.device, (specs,))
    flags = (_lib.GL_DENORMALIZE if denormalize else 0) | (_lib.GL_DEEMPHASIS if deemphasis else 0)
    h.griffin_lim(specs, _lib.FRAME_MAJOR, [T] * N, out, init_phase=init_phase, seed=seed, iters=-1 if iters is None else iters,
                  flags=flags, out_dtype=_lib.F64 if deemphasis else _lib.F32, space=_lib.DEVICE, stream=st)
    h.check_status(st)
    return out


def inv_spectrogram_batch(specs, init_phase=None, seed=0, iters=None, device=None, denormalize=True, deemphasis=True,
                          out=None, layout=None):
    """Griffin-Lim inversion of a batch.

    specs: list of [F,T_b] / [T_b,F] arrays (``layout`` "FT" / "TF" says which; None decides per array from the shape and
    refuses the ambiguous T == num_freq), or one [N,T,F] array.  init_phase: same structure (complex) or
    None -> device Philox keyed by ``seed``.  Returns a list of float64 waveforms (views into one buffer),
    or float32 when ``deemphasis`` is False (what ``_griffin_lim`` returns).  A [N,T,F] array that lives on the GPU
    (torch / ``__cuda_array_interface__`` / DLPack) is inverted in place there and a device array [N, n] comes back.
    """
    if _buffers.is_device_array(specs):
        return _inv_spectrogram_batch_device(specs, init_phase, seed, iters, denormalize, deemphasis)
    h = audio._handle(device)
    packed, Ts, init_phase = _gl_batch_inputs(h, specs, init_phase, layout)
    ns = [h.num_samples(T) for T in Ts]
    dt = np.float64 if deemphasis else np.float32
    out = _gl_out_buffer(h, ns, dt, out)
    flags = (_lib.GL_DENORMALIZE if denormalize else 0) | (_lib.GL_DEEMPHASIS if deemphasis else 0)
    h.griffin_lim(packed, _lib.FRAME_MAJOR, Ts, out, init_phase=init_phase, seed=seed, iters=-1 if iters is None else iters,
                  flags=flags, out_dtype=_lib.F64 if deemphasis else _lib.F32)
    return _split(out, ns)


def inv_spectrogram_stream(batches, seed=0, iters=None, device=None, denormalize=True, deemphasis=True, layout=None, in_flight=3, dtype=None):
    """Generator over an iterable of batches (each what ``inv_spectrogram_batch`` takes as ``specs``): yields each batch's
    list of waveforms, in order, with ``in_flight`` batches inside the library at any time (``nsb_griffin_lim_submit`` /
    ``nsb_wait``) - batch i+1 is copied in and starts while batch i finishes and is copied out.  The caller pattern is the
    reference's synthesis loops (eval.py:36-59: sentence after sentence) and feeder threads (datasets/datafeeder.py:110-152).
    The phase is drawn on the device (Philox, ``seed`` + the batch's index).  ``dtype``: the waveforms' type - None = the
    reference's (float64 after ``inv_preemphasis``, float32 without it); ``np.float32`` halves the bytes that come back, which
    is what bounds the end-to-end rate when several GPUs share one host."""
    h = audio._handle(device)
    flags = (_lib.GL_DENORMALIZE if denormalize else 0) | (_lib.GL_DEEMPHASIS if deemphasis else 0)
    dt = np.dtype(np.float64 if deemphasis else np.float32) if dtype is None else np.dtype(dtype)
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("dtype must be float32 or float64")
    pending = collections.deque()

    def collect():
        ticket, out, ns, _keep = pending.popleft()
        h.wait(ticket)
        return _split(out, ns)

    try:
        for i, specs in enumerate(batches):
            packed, Ts, _ = _gl_batch_inputs(h, specs, None, layout)
            ns = [h.num_samples(T) for T in Ts]
            out = _gl_out_buffer(h, ns, dt, None)
            t = h.griffin_lim_submit(packed, _lib.FRAME_MAJOR, Ts, out, seed=seed + i, iters=-1 if iters is None else iters,
                                     flags=flags, out_dtype=_lib.F64 if dt == np.dtype(np.float64) else _lib.F32)
            pending.append((t, out, ns, packed))
            if len(pending) >= max(1, in_flight):
                yield collect()
        while pending:
            yield collect()
    finally:
        while pending:                       # the consumer stopped early: the buffers must outlive the submitted calls
            try:
                collect()
            except Exception:
                pass


def bucket_by_length(n_frames, batch_size, rng=None):
    """The feeder's bucketing (reference datasets/datafeeder.py:143-147): sort the group's examples by output length
    (stable, like ``list.sort``), cut into batches of ``batch_size``, shuffle the batches with ``rng`` (an object with
    ``shuffle``, e.g. the ``random`` module as in the reference; None: keep the sorted order).  Returns lists of indices."""
    order = sorted(range(len(n_frames)), key=lambda i: int(n_frames[i]))
    batches = [order[i:i + batch_size] for i in range(0, len(order), batch_size)]
    if rng is not None:
        rng.shuffle(batches)
    return batches


def feeder_groups(wavs, batch_size, outputs_per_step, rng=None, device=None):
    """One group of the feeder (reference datasets/datafeeder.py:130-158 with ``_prepare_batch`` 190-216) from the waveforms:
    features of ALL ``batch_size * batch_group_size`` clips in one device pass, written straight into the bucketed batches'
    padded, time-major target tensors (``nsb_features_rows``) - no per-utterance numpy pad, no second pass over the frames.
    Returns a list of batches (dicts): ``indices`` (into ``wavs``), ``mel_targets`` [n, Tpad, num_mels], ``linear_targets``
    [n, Tpad, num_freq] with Tpad = round_up(max frames of the batch + 1, outputs_per_step), ``audios`` [n, max len] (padded
    with 0 like ``_prepare_inputs``), ``n_frames``."""
    h = audio._handle(device)
    wavs = [audio._as_wav(w) for w in wavs]
    ns = [w.size for w in wavs]
    Ts = [h.num_frames(n) for n in ns]
    buckets = bucket_by_length(Ts, batch_size, rng)
    row_off, rows_of, base, total = [0] * len(wavs), [], [], 0
    for idx in buckets:
        rows = _round_up(max(Ts[i] for i in idx) + 1, outputs_per_step)
        rows_of.append(rows)
        base.append(total)
        for j, i in enumerate(idx):
            row_off[i] = total + j * rows
        total += rows * len(idx)
    packed = _pack(h, wavs, np.float32)
    pool = h.lib.pinned_pool()
    lin = pool.empty((total, h.num_freq), np.float32)
    mel = pool.empty((total, h.num_mels), np.float32)
    h.features_rows(packed, ns, row_off, total, lin, mel)
    out = []
    for idx, rows, r0 in zip(buckets, rows_of, base):
        n = len(idx)
        max_len = max(ns[i] for i in idx)
        audios = np.zeros((n, max_len), dtype=np.float32)
        for j, i in enumerate(idx):
            audios[j, :ns[i]] = wavs[i]
        out.append({"indices": list(idx), "n_frames": [Ts[i] for i in idx], "audios": audios,
                    "mel_targets": mel[r0:r0 + n * rows].reshape(n, rows, h.num_mels),
                    "linear_targets": lin[r0:r0 + n * rows].reshape(n, rows, h.num_freq)})
    return out


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
    lengths = [np.shape(s)[0] if _is_time_major(np.shape(s), F, kw.get("layout")) else np.shape(s)[1] for s in specs]
    return run_sharded(lambda sub, dev: inv_spectrogram_batch(sub, device=dev, **kw), list(specs), lengths, devices)
