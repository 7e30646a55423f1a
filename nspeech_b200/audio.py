"""Drop-in mirror of the reference's ``neural_speech/utils/audio.py`` hot path on B200.

Same names, signatures, dtypes, shapes and hparams-driven behaviour as the reference module
(``/root/reference/neural_speech/utils/audio.py``, cited per function), but every function body marshals
its numpy buffers through ctypes into ``libnspeech_b200.so`` (hand-written sm_100a kernels).  No librosa,
scipy, cuFFT or CPU arithmetic on the path; if the native library or a B200 is missing the call raises.

Like the reference, every call reads the global hparams (``nspeech_b200.hparams.get_hparams()``).  Native
handles are cached per (audio hparams, device, thread), so changing an hparam can never hit a stale plan
(the reference's ``_mel_basis`` global is never invalidated, ``audio.py:135-142``) and feeder threads
(``datasets/datafeeder.py:110-116``) do not serialise on one handle.

Additions that the reference does not have (all optional keyword arguments or new names):
``inv_spectrogram(..., init_phase=, seed=, iters=)``, ``spectrogram_and_mel``, ``peak_normalize``, the ``*_batch`` functions in
``nspeech_b200.batch``.

Device arrays.  Every function that takes a waveform or a spectrogram also takes an array that already lives on the GPU -
a torch CUDA tensor, or anything exposing ``__cuda_array_interface__`` / ``__dlpack__`` (cupy, numba, jax) - and then runs
the NSB_DEVICE path on the caller's stream (torch's current stream; the stream the array interface names; else the NULL
stream) and returns a device array: a torch tensor for torch input, otherwise a ``nspeech_b200._buffers.DeviceArray``
(adopt it with ``torch.from_dlpack`` / ``cupy.asarray``; ``.copy_to_host()`` gives numpy).  No host round trip.  The call
checks the device-side error flag before it returns, so non-finite input raises exactly as on the host path.
"""
import threading

import numpy as np

from . import _buffers, _lib
from ._lib import ParameterError  # noqa: F401  (re-exported)
from .hparams import get_hparams

# device used by the module-level functions (one process per GPU: set from LOCAL_RANK by the caller)
DEVICE = 0
# how inv_spectrogram/_griffin_lim draw the initial phase when none is supplied:
#   "numpy"  - np.random.rand(*S.shape) from the global numpy RNG, exactly as the reference (audio.py:81)
#   "device" - Philox-4x32 on the GPU keyed by `seed` (fast path; no host RNG, no phase upload)
RANDOM_PHASE = "numpy"

_tls = threading.local()


def _handle(device=None):
    hp = get_hparams()
    dev = DEVICE if device is None else device
    key = (tuple(getattr(hp, k) for k in ("num_freq", "num_mels", "sample_rate", "griffin_lim_iters", "frame_shift_ms",
                                            "frame_length_ms", "preemphasis", "ref_level_db", "min_level_db", "power")),
           dev, id(_lib.default_lib()))
    cache = getattr(_tls, "handles", None)
    if cache is None:
        cache = _tls.handles = {}
    h = cache.get(key)
    if h is None:
        if len(cache) > 8:
            for old in cache.values():
                old.close()
            cache.clear()
        h = cache[key] = _lib.Handle(hp, dev)
    return h


def _as_wav(y):
    y = np.asarray(y)
    if y.ndim != 1:
        raise ParameterError('Invalid shape for monophonic audio: ndim={:d}, shape={}'.format(y.ndim, y.shape))
    if not np.issubdtype(y.dtype, np.floating):
        raise ParameterError('data must be floating-point')
    if y.size == 0:
        raise ParameterError('empty audio buffer')
    return np.ascontiguousarray(y, dtype=np.float32)


def _lfilter_input(x):
    """what scipy.signal.lfilter (audio.py:32, 36) accepts where librosa's valid_audio does not: any real dtype and an empty
    array (-> an empty float64 result)"""
    x = np.asarray(x)
    if x.ndim != 1:
        raise ParameterError('Invalid shape for monophonic audio: ndim={:d}, shape={}'.format(x.ndim, x.shape))
    if np.iscomplexobj(x) or not (np.issubdtype(x.dtype, np.number) or x.dtype == np.bool_):
        raise ParameterError('data must be real-valued')
    return np.ascontiguousarray(x, dtype=np.float32)


def _spec_layout(S, dtype):
    """Return (buffer, layout) for a [F, T] array without copying when it is either Fortran- or C-ordered."""
    S = np.asarray(S)
    if S.ndim != 2:
        raise ValueError("expected a [num_freq, T] array, got shape %r" % (S.shape,))
    if S.dtype != dtype:
        S = S.astype(dtype)
    if S.T.flags.c_contiguous:
        return S, _lib.FRAME_MAJOR
    if S.flags.c_contiguous:
        return S, _lib.BIN_MAJOR
    return np.asfortranarray(S), _lib.FRAME_MAJOR


def _dev_array(a, dtype, ndim, what):
    """a device array of the right dtype and rank, dense -> (object to pass, Buf).  torch tensors are converted with torch
    ops when dtype or strides do not fit; other producers must hand in what the kernels read."""
    if _buffers._torch_of(a) is not None:
        import torch
        tdt = {np.float32: torch.float32, np.complex64: torch.complex64, np.float64: torch.float64}[dtype]
        if a.dtype != tdt:
            a = a.to(tdt)
        if not (a.is_contiguous() or (a.dim() == 2 and a.T.is_contiguous())):
            a = a.contiguous()
        it = np.dtype(dtype).itemsize
        buf = _buffers.Buf(a.data_ptr(), tuple(a.shape), dtype, tuple(st * it for st in a.stride()), True, a.device.index or 0, a)
    else:
        buf = _buffers.as_buffer(a)
    if not buf.on_device:
        raise TypeError("%s: expected a device array" % what)
    if buf.dtype != np.dtype(dtype):
        raise TypeError("%s: device array must be %s, got %s" % (what, np.dtype(dtype), buf.dtype))
    if buf.ndim != ndim:
        raise ValueError("%s: expected %d dimensions, got shape %r" % (what, ndim, buf.shape))
    if not (buf.c_contiguous or buf.f_contiguous):
        raise ValueError("%s: device array must be dense (C- or Fortran-ordered)" % what)
    return a, buf


def _dev_finish(h, stream):
    """deferred error flag of the NSB_DEVICE calls (non-finite data): synchronises the stream, raises like the host path"""
    h.check_status(stream)


def _stft_parameters():
    # reference audio.py:126-130
    h = _handle()
    return h.n_fft, h.hop, h.win


def _dev_emph(x, inverse):
    x, b = _dev_array(x, np.float32, 1, "x")
    h = _handle(b.device)
    st = _buffers.stream_of(x)
    out = _buffers.empty_like_source(h.lib, b.shape, np.float64, b.device, (x,))
    h.preemphasis(x, [b.size], out, _lib.F64, space=_lib.DEVICE, stream=st, inverse=inverse)
    _dev_finish(h, st)
    return out


def preemphasis(x):
    # reference audio.py:31-32 -> float64, like scipy.signal.lfilter
    if _buffers.is_device_array(x):
        return _dev_emph(x, False)
    x = _lfilter_input(x)
    out = np.empty(x.shape, dtype=np.float64)
    if x.size:
        _handle().preemphasis(x, [x.size], out, _lib.F64)
    return out


def inv_preemphasis(x):
    # reference audio.py:35-36
    if _buffers.is_device_array(x):
        return _dev_emph(x, True)
    x = _lfilter_input(x)
    out = np.empty(x.shape, dtype=np.float64)
    if x.size:
        _handle().preemphasis(x, [x.size], out, _lib.F64, inverse=True)
    return out


def _dev_features(y, want_stft=False, want_lin=True, want_mel=True):
    y, b = _dev_array(y, np.float32, 1, "y")
    if b.size == 0:
        raise ParameterError('empty audio buffer')
    h = _handle(b.device)
    st = _buffers.stream_of(y)
    T = h.num_frames(b.size)
    if want_stft:
        out = _buffers.empty_like_source(h.lib, (T, h.num_freq), np.complex64, b.device, (y,))
        h.stft(y, [b.size], out, preemphasis=False, space=_lib.DEVICE, stream=st)
        _dev_finish(h, st)
        return out.T
    lin = _buffers.empty_like_source(h.lib, (T, h.num_freq), np.float32, b.device, (y,)) if want_lin else None
    mel = _buffers.empty_like_source(h.lib, (T, h.num_mels), np.float32, b.device, (y,)) if want_mel else None
    h.features(y, [b.size], lin, mel, space=_lib.DEVICE, stream=st)
    _dev_finish(h, st)
    return (lin.T if want_lin else None), (mel.T if want_mel else None)


def _stft(y):
    # reference audio.py:106-108 -> complex64 [F, T], Fortran-ordered like librosa.stft
    if _buffers.is_device_array(y):
        return _dev_features(y, want_stft=True)
    y = _as_wav(y)
    h = _handle()
    T = h.num_frames(y.size)
    out = np.empty((T, h.num_freq), dtype=np.complex64)
    h.stft(y, [y.size], out, preemphasis=False)
    return out.T


def _dev_spec(S, dtype, what):
    """[F, T] device spectrogram -> (object, Buf, layout)"""
    S, b = _dev_array(S, dtype, 2, what)
    return S, b, (_lib.FRAME_MAJOR if b.f_contiguous else _lib.BIN_MAJOR)


def _istft(D):
    # reference audio.py:111-113 -> float32, hop*(T-1) samples
    if _buffers.is_device_array(D):
        D, b, layout = _dev_spec(D, np.complex64, "D")
        h = _handle(b.device)
        if b.shape[0] != h.num_freq:
            raise ValueError("expected %d frequency bins, got %d" % (h.num_freq, b.shape[0]))
        st = _buffers.stream_of(D)
        out = _buffers.empty_like_source(h.lib, (h.num_samples(b.shape[1]),), np.float32, b.device, (D,))
        h.istft(D, layout, [b.shape[1]], out, space=_lib.DEVICE, stream=st)
        _dev_finish(h, st)
        return out
    D, layout = _spec_layout(D, np.complex64)
    h = _handle()
    if D.shape[0] != h.num_freq:
        raise ValueError("expected %d frequency bins, got %d" % (h.num_freq, D.shape[0]))
    T = D.shape[1]
    out = np.empty(h.num_samples(T), dtype=np.float32)
    h.istft(D, layout, [T], out)
    return out


def spectrogram_and_mel(y):
    """``(spectrogram(y), melspectrogram(y))`` from ONE pass (the reference runs the STFT twice:
    ``datasets/process.py:30,33`` -> ``audio.py:40,62``)."""
    if _buffers.is_device_array(y):
        return _dev_features(y)
    y = _as_wav(y)
    h = _handle()
    T = h.num_frames(y.size)
    lin = np.empty((T, h.num_freq), dtype=np.float32)
    mel = np.empty((T, h.num_mels), dtype=np.float32)
    h.features(y, [y.size], lin, mel)
    return lin.T, mel.T


def spectrogram(y):
    # reference audio.py:39-42 -> float32 [F, T]
    if _buffers.is_device_array(y):
        return _dev_features(y, want_mel=False)[0]
    y = _as_wav(y)
    h = _handle()
    lin = np.empty((h.num_frames(y.size), h.num_freq), dtype=np.float32)
    h.features(y, [y.size], lin, None)
    return lin.T


def melspectrogram(y):
    # reference audio.py:61-64 -> float32 [M, T]
    if _buffers.is_device_array(y):
        return _dev_features(y, want_lin=False)[1]
    y = _as_wav(y)
    h = _handle()
    mel = np.empty((h.num_frames(y.size), h.num_mels), dtype=np.float32)
    h.features(y, [y.size], None, mel)
    return mel.T


def _draw_phase(shape, layout):
    # reference audio.py:81: angles = exp(2j*pi*rand(*S.shape)) from the global numpy RNG
    u = np.random.rand(*shape)
    ang = np.exp(2j * np.pi * u).astype(np.complex64)
    return np.asfortranarray(ang) if layout == _lib.FRAME_MAJOR else np.ascontiguousarray(ang)


def _run_gl_device(S, flags, out_dtype, init_phase, seed, iters):
    S, b, layout = _dev_spec(S, np.float32, "spectrogram")
    h = _handle(b.device)
    if b.shape[0] != h.num_freq:
        raise ValueError("expected %d frequency bins, got %d" % (h.num_freq, b.shape[0]))
    T = b.shape[1]
    if init_phase is not None:
        if not _buffers.is_device_array(init_phase):
            raise TypeError("init_phase must live on the device like the spectrogram")
        init_phase, pb, pl = _dev_spec(init_phase, np.complex64, "init_phase")
        if pb.shape != b.shape or pl != layout:
            raise ValueError("init_phase must have the spectrogram's shape %r and memory order" % (b.shape,))
    st = _buffers.stream_of(S, init_phase)
    out = _buffers.empty_like_source(h.lib, (h.num_samples(T),), np.float64 if out_dtype == _lib.F64 else np.float32, b.device, (S,))
    h.griffin_lim(S, layout, [T], out, init_phase=init_phase, seed=0 if seed is None else seed, iters=-1 if iters is None else iters,
                  flags=flags, out_dtype=out_dtype, space=_lib.DEVICE, stream=st)
    _dev_finish(h, st)
    return out


def _run_gl(S, flags, out_dtype, init_phase, seed, iters):
    if _buffers.is_device_array(S):          # the phase is then drawn on the device (Philox keyed by `seed`) unless supplied there
        return _run_gl_device(S, flags, out_dtype, init_phase, seed, iters)
    S, layout = _spec_layout(S, np.float32)
    h = _handle()
    if S.shape[0] != h.num_freq:
        raise ValueError("expected %d frequency bins, got %d" % (h.num_freq, S.shape[0]))
    T = S.shape[1]
    if init_phase is None and RANDOM_PHASE == "numpy" and seed is None:
        init_phase = _draw_phase(S.shape, layout)
    if init_phase is not None:
        init_phase = np.asarray(init_phase)
        if init_phase.shape != S.shape:
            raise ValueError("init_phase shape %r != spectrogram shape %r" % (init_phase.shape, S.shape))
        if layout == _lib.FRAME_MAJOR:
            init_phase = np.asfortranarray(init_phase, dtype=np.complex64)
        else:
            init_phase = np.ascontiguousarray(init_phase, dtype=np.complex64)
    out = np.empty(h.num_samples(T), dtype=np.float64 if out_dtype == _lib.F64 else np.float32)
    h.griffin_lim(S, layout, [T], out, init_phase=init_phase, seed=0 if seed is None else seed,
                  iters=-1 if iters is None else iters, flags=flags, out_dtype=out_dtype)
    return out


def _griffin_lim(S, init_phase=None, seed=None, iters=None):
    # reference audio.py:77-87 -> float32 waveform (no de-emphasis)
    return _run_gl(S, 0, _lib.F32, init_phase, seed, iters)


def inv_spectrogram(spectrogram, init_phase=None, seed=None, iters=None):
    '''Converts spectrogram to waveform (reference audio.py:45-48) -> float64, hop*(T-1) samples'''
    return _run_gl(spectrogram, _lib.GL_DENORMALIZE | _lib.GL_DEEMPHASIS, _lib.F64, init_phase, seed, iters)


def _build_mel_basis():
    # reference audio.py:145-147 -> float64 [num_mels, num_freq]
    return _handle().mel_basis()


def _linear_to_mel(spectrogram):
    # reference audio.py:138-142 -> float64 [num_mels, T] (np.dot with the float64 basis)
    S, layout = _spec_layout(spectrogram, np.float32)
    h = _handle()
    T = S.shape[1]
    out = np.empty((T, h.num_mels), dtype=np.float64)
    h.linear_to_mel(S, layout, [T], out, _lib.F64)
    return out.T


def _elementwise(op, x):
    x = np.asarray(x)
    flat = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    out = np.empty_like(flat)
    if flat.size:
        _handle().elementwise(op, flat, out)
    return out.reshape(x.shape)


def _amp_to_db(x):
    # reference audio.py:150-151
    return _elementwise(_lib.EW_AMP_TO_DB, x)


def _db_to_amp(x):
    # reference audio.py:154-155
    return _elementwise(_lib.EW_DB_TO_AMP, x)


def _normalize(S):
    # reference audio.py:162-163
    return _elementwise(_lib.EW_NORMALIZE, S)


def _denormalize(S):
    # reference audio.py:166-167
    return _elementwise(_lib.EW_DENORMALIZE, S)


# ---------------------------------------------------------------------------------------------------------------
# The TensorFlow twin (reference audio.py:51-58, 90-103, 116-123, 158-159, 170-171) - the variant the reference's live
# callers use (synthesizer.py:30, models/tacotron.py:107).  In the reference these build TF-1 graph ops; here they are
# eager functions on numpy arrays with the same semantics: time-major [T, F] (or batched [N, T, F]) spectrograms,
# tf.contrib.signal framing (no centring / padding, window on the first win samples, zero-padded at the end),
# inverse without window-sum normalisation, zero initial phase, est / max(1e-8, |est|), float32 results of
# win + hop*(T-1) samples, and NO inverse pre-emphasis (the caller applies it, synthesizer.py:52).
# ---------------------------------------------------------------------------------------------------------------

def _tf_batch(S, dtype):
    S = np.asarray(S)
    if S.ndim not in (2, 3):
        raise ValueError("expected [T, F] or [N, T, F], got shape %r" % (S.shape,))
    return np.ascontiguousarray(S, dtype=dtype), S.ndim == 3


def _dev_tf_batch(S, dtype, what):
    """[T, F] or [N, T, F] C-contiguous device array -> (object, Buf, batched)"""
    if _buffers._torch_of(S) is not None and not S.is_contiguous():
        S = S.contiguous()
    b = _buffers.as_buffer(S)
    if b.ndim not in (2, 3):
        raise ValueError("expected [T, F] or [N, T, F], got shape %r" % (b.shape,))
    S, b = _dev_array(S, dtype, b.ndim, what)
    if not b.c_contiguous:
        raise ValueError("%s: time-major device arrays must be C-contiguous" % what)
    return S, b, b.ndim == 3


def _run_gl_tf_device(S, flags, iters):
    S, b, batched = _dev_tf_batch(S, np.float32, "spectrogram")
    h = _handle(b.device)
    if b.shape[-1] != h.num_freq:
        raise ValueError("expected %d frequency bins on the last axis, got %d" % (h.num_freq, b.shape[-1]))
    N, T = (b.shape[0] if batched else 1), b.shape[-2]
    st = _buffers.stream_of(S)
    n = h.num_samples_tf(T)
    out = _buffers.empty_like_source(h.lib, (N, n) if batched else (n,), np.float32, b.device, (S,))
    h.griffin_lim(S, _lib.FRAME_MAJOR, [T] * N, out, iters=-1 if iters is None else iters, flags=flags | _lib.GL_TF_TWIN,
                  out_dtype=_lib.F32, space=_lib.DEVICE, stream=st)
    _dev_finish(h, st)
    return out


def _run_gl_tf(S, flags, iters):
    if _buffers.is_device_array(S):
        return _run_gl_tf_device(S, flags, iters)
    S, batched = _tf_batch(S, np.float32)
    h = _handle()
    if S.shape[-1] != h.num_freq:
        raise ValueError("expected %d frequency bins on the last axis, got %d" % (h.num_freq, S.shape[-1]))
    N = S.shape[0] if batched else 1
    T = S.shape[-2]
    n = h.num_samples_tf(T)
    out = np.empty((N, n), dtype=np.float32)
    h.griffin_lim(S, _lib.FRAME_MAJOR, [T] * N, out, iters=-1 if iters is None else iters, flags=flags | _lib.GL_TF_TWIN,
                  out_dtype=_lib.F32)
    return out if batched else out[0]


def _griffin_lim_tensorflow(S, iters=None):
    # reference audio.py:90-103
    return _run_gl_tf(S, 0, iters)


def inv_spectrogram_tensorflow(spectrogram, iters=None):
    '''reference audio.py:51-58; like there, this does NOT invert the preemphasis'''
    return _run_gl_tf(spectrogram, _lib.GL_DENORMALIZE, iters)


def _stft_tensorflow(signals):
    # reference audio.py:116-118 -> complex64 [T, F] (or [N, T, F])
    x = np.asarray(signals)
    batched = x.ndim == 2
    x = np.ascontiguousarray(x, dtype=np.float32)
    h = _handle()
    N = x.shape[0] if batched else 1
    n = x.shape[-1]
    if n < h.win:
        raise ValueError("signal shorter than one frame (%d < %d)" % (n, h.win))
    T = h.num_frames_tf(n)
    out = np.empty((N, T, h.num_freq), dtype=np.complex64)
    h.stft_tf(x, [n] * N, out)
    return out if batched else out[0]


def _istft_tensorflow(stfts):
    # reference audio.py:121-123 -> float32 [win + hop*(T-1)] (or [N, ...])
    D, batched = _tf_batch(stfts, np.complex64)
    h = _handle()
    N = D.shape[0] if batched else 1
    T = D.shape[-2]
    out = np.empty((N, h.num_samples_tf(T)), dtype=np.float32)
    h.istft_tf(D, _lib.FRAME_MAJOR, [T] * N, out)
    return out if batched else out[0]


def _db_to_amp_tensorflow(x):
    # reference audio.py:158-159
    return _db_to_amp(x)


def _denormalize_tensorflow(S):
    # reference audio.py:170-171
    return _denormalize(S)


# ---------------------------------------------------------------------------------------------------------------
# The steps after the vocoder in Synthesizer.synthesize (reference synthesizer.py:51-53)
# ---------------------------------------------------------------------------------------------------------------

def find_endpoint(wav, threshold_db=-40, min_silence_sec=0.8):
    # reference audio.py:67-74 (np.max of the window, not max |.|; returns len(wav) when no silent window is found)
    w = np.asarray(wav)
    if w.ndim != 1:
        raise ValueError("expected a 1-D waveform, got shape %r" % (w.shape,))
    if w.size == 0:
        return 0
    dt = _lib.F64 if w.dtype == np.float64 else _lib.F32
    w = np.ascontiguousarray(w, dtype=np.float64 if dt == _lib.F64 else np.float32)
    out = np.zeros(1, dtype=np.int64)
    _handle().find_endpoint(w, [w.size], out, dtype=dt, threshold_db=threshold_db, min_silence_sec=min_silence_sec)
    return int(out[0])


def _synth_inputs(linear_outputs, h):
    """[T, F], [N, T, F] or a list of [T_i, F] (the per-utterance calls of synthesizer.py:51-53 differ in length)
    -> (packed [sum T, F] float32, n_frames list, batched?)"""
    if isinstance(linear_outputs, (list, tuple)):
        mats = [np.ascontiguousarray(np.asarray(m), dtype=np.float32) for m in linear_outputs]
        for m in mats:
            if m.ndim != 2 or m.shape[1] != h.num_freq:
                raise ValueError("expected [T, %d] spectrograms, got shape %r" % (h.num_freq, m.shape))
        if not mats:
            raise ValueError("empty batch")
        return (np.concatenate(mats) if len(mats) > 1 else mats[0]), [m.shape[0] for m in mats], True
    S, batched = _tf_batch(linear_outputs, np.float32)
    if S.shape[-1] != h.num_freq:
        raise ValueError("expected %d frequency bins on the last axis, got %d" % (h.num_freq, S.shape[-1]))
    N = S.shape[0] if batched else 1
    return S.reshape(-1, h.num_freq), [S.shape[-2]] * N, batched


def synthesize_waveforms(linear_outputs, iters=None, threshold_db=-40, min_silence_sec=0.8, peak_normalize=False, dtype=np.float64):
    """The spectrogram -> waveform stage of ``Synthesizer.synthesize`` (synthesizer.py:30, 51-53) for ``[T, F]``, a uniform
    batch ``[N, T, F]`` or a ragged list of ``[T_i, F]`` normalised linear spectrograms, in one device pipeline:
    ``wav = inv_spectrogram_tensorflow(lin); wav = inv_preemphasis(wav); wav = wav[:find_endpoint(wav)]``.
    ``peak_normalize=True`` adds what ``save_wav`` does to that result before writing it (audio.py:17-19, called by
    eval.py:43 / train.py:108): ``wav *= 32767 / max(0.01, max |wav|)``; ``dtype=np.int16`` (needs ``peak_normalize``)
    returns the scaled samples as 16-bit integers (``astype(np.int16)``), a quarter of the bytes to bring back.
    Returns a waveform (or a list of them)."""
    h = _handle()
    dtype = np.dtype(dtype)
    if dtype not in (np.dtype(np.float64), np.dtype(np.int16)):
        raise ValueError("dtype must be float64 or int16")
    if dtype == np.int16 and not peak_normalize:
        raise ValueError("int16 output needs peak_normalize=True (save_wav's scaling)")
    packed, Ts, batched = _synth_inputs(linear_outputs, h)
    ns = [h.num_samples_tf(T) for T in Ts]
    total = sum(ns)
    # large results land in pooled page-locked memory (the copy out then runs at PCIe speed, see batch.inv_spectrogram_batch)
    wav = h.lib.pinned_pool().empty((total,), dtype) if total * dtype.itemsize >= (1 << 20) else np.empty((total,), dtype=dtype)
    ends = np.zeros(len(Ts), dtype=np.int64)
    h.synthesize(packed, Ts, wav, ends, iters=-1 if iters is None else iters, threshold_db=threshold_db, min_silence_sec=min_silence_sec,
                 flags=_lib.SYNTH_PEAK_NORMALIZE if peak_normalize else 0, out_dtype=_lib.I16 if dtype == np.int16 else _lib.F64)
    outs, off = [], 0
    for n, e in zip(ns, ends):
        outs.append(wav[off:off + int(e)])
        off += n
    return outs if batched else outs[0]


def peak_normalize(wav, dtype=np.float64):
    """The arithmetic of ``save_wav`` (reference audio.py:17-19): ``wav * (32767 / max(0.01, np.max(np.abs(wav))))`` as a new
    float64 array, or (``dtype=np.int16``) that array through ``astype(np.int16)``."""
    w = np.asarray(wav)
    if w.ndim != 1:
        raise ValueError("expected a 1-D waveform, got shape %r" % (w.shape,))
    dtype = np.dtype(dtype)
    if dtype not in (np.dtype(np.float64), np.dtype(np.int16)):
        raise ValueError("dtype must be float64 or int16")
    out = np.empty(w.shape, dtype=dtype)
    if w.size == 0:
        return out
    dt = _lib.F64 if w.dtype == np.float64 else _lib.F32
    w = np.ascontiguousarray(w, dtype=np.float64 if dt == _lib.F64 else np.float32)
    _handle().peak_normalize(w, [w.size], out, wav_dtype=dt, out_dtype=_lib.I16 if dtype == np.int16 else _lib.F64)
    return out


def save_wav(wav, path):
    # reference audio.py:17-19: scales the caller's array IN PLACE, then librosa.output.write_wav -> scipy.io.wavfile.write
    # (file I/O is host work; the scaling runs on the device)
    from scipy.io import wavfile
    wav[...] = peak_normalize(wav)
    wavfile.write(path, int(get_hparams().sample_rate), wav)


def load_spectrogram(path):
    # reference audio.py:22-24
    spec = np.load(path)
    return spec, spec.shape[1]


def save_spectrogram(spec, path):
    # reference audio.py:27-28: the [F, T] feature as numpy wrote it (Fortran-ordered for what spectrogram() returns, like librosa's)
    np.save(path, spec, allow_pickle=False)
