"""ctypes binding of libnspeech_b200.so (the C ABI in include/nspeech_b200.h).

The product loads exactly one file: ``nspeech_b200/libnspeech_b200.so`` built by ``__graft_entry__.build()``
(nvcc, sm_100a).  If it is missing or cannot be loaded the import of any compute entry point raises - there
is no CPU fallback.  (``NativeLib(path)`` accepts an explicit path only so that tests can bind the
CPU-emulated build of the same sources under tests/emu/.)
"""
import ctypes
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "libnspeech_b200.so")

NSB_OK, NSB_ERR_INVALID, NSB_ERR_CUDA, NSB_ERR_UNSUPPORTED, NSB_ERR_NONFINITE, NSB_ERR_NODEVICE, NSB_ERR_OOM = range(7)
HOST, DEVICE = 0, 1
FRAME_MAJOR, BIN_MAJOR = 0, 1
F32, F64, I16 = 0, 1, 2
SYNTH_PEAK_NORMALIZE = 1
EW_AMP_TO_DB, EW_DB_TO_AMP, EW_NORMALIZE, EW_DENORMALIZE = range(4)
GL_DENORMALIZE, GL_DEEMPHASIS, GL_TF_TWIN = 1, 2, 4
OPT_STREAM_SYNC_MODE, OPT_FUSE_ITERATIONS, OPT_WIDE_MODE, OPT_OVERLAP_CHUNKS, OPT_WAVE_SCHEDULE, OPT_MEL_LINES, OPT_SPECIALIZE = 1, 2, 3, 4, 5, 6, 7


class ParameterError(ValueError):
    """Raised where librosa 0.6.0 raises ``librosa.util.exceptions.ParameterError`` (non-finite audio, bad shapes)."""


class NativeError(RuntimeError):
    pass


class HParamsStruct(ctypes.Structure):
    _fields_ = [("num_freq", ctypes.c_int32), ("num_mels", ctypes.c_int32), ("sample_rate", ctypes.c_int32),
                ("griffin_lim_iters", ctypes.c_int32), ("frame_shift_ms", ctypes.c_double),
                ("frame_length_ms", ctypes.c_double), ("preemphasis", ctypes.c_double),
                ("ref_level_db", ctypes.c_double), ("min_level_db", ctypes.c_double), ("power", ctypes.c_double)]


_vp, _i32, _i64, _u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
_pi32, _pi64 = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes); every symbol declared in include/nspeech_b200.h
SIGNATURES = {
    "nsb_abi_version": (ctypes.c_int, []),
    "nsb_last_error": (ctypes.c_char_p, []),
    "nsb_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "nsb_device_pci_bus_id": (ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, _i32]),
    "nsb_create": (ctypes.c_int, [ctypes.POINTER(HParamsStruct), ctypes.c_int, ctypes.POINTER(_vp)]),
    "nsb_destroy": (ctypes.c_int, [_vp]),
    "nsb_synchronize": (ctypes.c_int, [_vp, _vp]),
    "nsb_stft_parameters": (ctypes.c_int, [_vp, _pi32, _pi32, _pi32]),
    "nsb_num_frames": (_i64, [_vp, _i64]),
    "nsb_num_samples": (_i64, [_vp, _i64]),
    "nsb_stft": (ctypes.c_int, [_vp, _vp, _pi64, _i32, _i32, _vp, _i32, _vp]),
    "nsb_stft_tf": (ctypes.c_int, [_vp, _vp, _pi64, _i32, _vp, _i32, _vp]),
    "nsb_istft_tf": (ctypes.c_int, [_vp, _vp, _i32, _pi32, _i32, _vp, _i32, _vp]),
    "nsb_features": (ctypes.c_int, [_vp, _vp, _pi64, _i32, _vp, _vp, _i32, _vp]),
    "nsb_istft": (ctypes.c_int, [_vp, _vp, _i32, _pi32, _i32, _vp, _i32, _vp]),
    "nsb_griffin_lim": (ctypes.c_int, [_vp, _vp, _i32, _pi32, _i32, _vp, _u64, _i32, _i32, _vp, _i32, _i32, _vp]),
    "nsb_preemphasis": (ctypes.c_int, [_vp, _vp, _pi64, _i32, _vp, _i32, _i32, _vp]),
    "nsb_inv_preemphasis": (ctypes.c_int, [_vp, _vp, _pi64, _i32, _vp, _i32, _i32, _vp]),
    "nsb_linear_to_mel": (ctypes.c_int, [_vp, _vp, _i32, _pi32, _i32, _vp, _i32, _i32, _vp]),
    "nsb_mel_basis": (ctypes.c_int, [_vp, _vp]),
    "nsb_elementwise": (ctypes.c_int, [_vp, _i32, _vp, _i64, _vp, _i32, _vp]),
    "nsb_check_status": (ctypes.c_int, [_vp, _vp]),
    "nsb_set_tile_hops": (ctypes.c_int, [_vp, _i32]),
    "nsb_set_generic_iteration": (ctypes.c_int, [_vp, _i32]),
    "nsb_set_stream_grid": (ctypes.c_int, [_vp, _i32]),
    "nsb_set_option": (ctypes.c_int, [_vp, _i32, _i32]),
    "nsb_set_host_chunks": (ctypes.c_int, [_vp, _i32]),
    "nsb_stream_trace": (ctypes.c_int, [_vp, _i32, _vp, _i32]),
    "nsb_kernel_launches": (_u64, [_vp]),
    "nsb_features_padded": (ctypes.c_int, [_vp, _vp, _pi64, _i32, _i32, _vp, _vp, _i32, _vp]),
    "nsb_frame_energy": (ctypes.c_int, [_vp, _vp, _pi64, _i32, _i32, _i32, _vp, _i32, _vp]),
    "nsb_find_endpoint": (ctypes.c_int, [_vp, _vp, _i32, _pi64, _i32, ctypes.c_double, ctypes.c_double, _vp, _i32, _vp]),
    "nsb_synthesize": (ctypes.c_int, [_vp, _vp, _pi32, _i32, _i32, ctypes.c_double, ctypes.c_double, _vp, _vp, _i32, _vp]),
    "nsb_griffin_lim_iterate": (ctypes.c_int, [_vp, _i32, _vp]),
    "nsb_synthesize_ex": (ctypes.c_int, [_vp, _vp, _pi32, _i32, _i32, ctypes.c_double, ctypes.c_double, _i32, _vp, _i32, _vp, _i32, _vp]),
    "nsb_peak_normalize": (ctypes.c_int, [_vp, _vp, _i32, _pi64, _i32, _vp, _vp, _i32, _i32, _vp]),
    "nsb_features_rows": (ctypes.c_int, [_vp, _vp, _pi64, _i32, _pi64, _i64, _vp, _vp, _i32, _vp]),
    "nsb_griffin_lim_submit": (ctypes.c_int, [_vp, _vp, _i32, _pi32, _i32, _vp, _u64, _i32, _i32, _vp, _i32, ctypes.POINTER(_u64)]),
    "nsb_features_submit": (ctypes.c_int, [_vp, _vp, _pi64, _i32, _vp, _vp, ctypes.POINTER(_u64)]),
    "nsb_synthesize_submit": (ctypes.c_int, [_vp, _vp, _pi32, _i32, _i32, ctypes.c_double, ctypes.c_double, _i32, _vp, _i32, _vp, ctypes.POINTER(_u64)]),
    "nsb_wait": (ctypes.c_int, [_vp, _u64]),
    "nsb_set_async_slots": (ctypes.c_int, [_vp, _i32]),
    "nsb_device_alloc": (ctypes.c_int, [ctypes.c_int, _u64, ctypes.POINTER(_vp)]),
    "nsb_device_free": (ctypes.c_int, [ctypes.c_int, _vp]),
    "nsb_device_copy": (ctypes.c_int, [ctypes.c_int, _vp, _vp, _u64, _i32]),
    "nsb_alloc_pinned": (ctypes.c_int, [_u64, ctypes.POINTER(_vp)]),
    "nsb_free_pinned": (ctypes.c_int, [_vp]),
}


class NativeLib(object):
    _pool = None

    def pinned_pool(self):
        if self._pool is None:
            self._pool = PinnedPool(self)
        return self._pool

    def __init__(self, path=None):
        self.path = DEFAULT_LIB if path is None else path
        if not os.path.exists(self.path):
            raise NativeError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). There is no CPU fallback." % self.path)
        self.dll = ctypes.CDLL(self.path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.dll, name)       # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if self.dll.nsb_abi_version() != 2:
            raise NativeError("ABI version mismatch")

    def check(self, rc):
        if rc == NSB_OK:
            return
        msg = (self.dll.nsb_last_error() or b"").decode("utf-8", "replace")
        if rc == NSB_ERR_NONFINITE:
            raise ParameterError(msg)
        if rc in (NSB_ERR_INVALID, NSB_ERR_UNSUPPORTED):
            raise ValueError(msg)
        if rc == NSB_ERR_OOM:
            raise MemoryError(msg)
        raise NativeError("nspeech_b200 error %d: %s" % (rc, msg))

    def device_count(self):
        n = ctypes.c_int(0)
        rc = self.dll.nsb_device_count(ctypes.byref(n))
        return n.value if rc == NSB_OK else 0

    def device_pci_bus_id(self, device):
        buf = ctypes.create_string_buffer(32)
        self.check(self.dll.nsb_device_pci_bus_id(int(device), buf, 32))
        return buf.value.decode()


class PinnedArray(object):
    """numpy view of a cudaHostAlloc'ed buffer; keep the object alive as long as the view is used."""

    def __init__(self, shape, dtype, lib=None):
        self.lib = default_lib() if lib is None else lib
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = ctypes.c_void_p()
        self.lib.check(self.lib.dll.nsb_alloc_pinned(ctypes.c_uint64(n), ctypes.byref(p)))
        self._p = p
        buf = (ctypes.c_char * max(n, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p is not None and self._p.value:
            self.array = None
            self.lib.dll.nsb_free_pinned(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedPool(object):
    """Result buffers in page-locked memory, recycled.  A device-to-host copy into pageable numpy memory is staged by the
    driver at a fraction of PCIe bandwidth (measured: 37.6 vs 25 ms for BASELINE config 4), and cudaHostAlloc itself costs
    milliseconds, so the batch front end hands out arrays backed by pooled pinned blocks: when the last view of such an
    array is garbage-collected the block returns to the pool.  At most ``max_bytes`` of free blocks are kept; over that the
    blocks that have been free the longest go first (a workload that changes its sizes must not end up allocating and
    freeing on every call because blocks of its old sizes fill the pool).  A request takes the smallest free block that
    holds it without wasting more than half."""

    def __init__(self, lib, max_bytes=2 << 30, granule=1 << 20):
        self.lib, self.max_bytes, self.granule = lib, max_bytes, granule
        self.free, self.kept, self.lock = [], 0, threading.Lock()        # free: [(size, ptr)], oldest first

    def _release(self, ptr, size):
        evict = []
        with self.lock:
            self.free.append((size, ptr))
            self.kept += size
            while self.kept > self.max_bytes and self.free:
                sz, p = self.free.pop(0)
                self.kept -= sz
                evict.append(p)
        for p in evict:
            self.lib.dll.nsb_free_pinned(ctypes.c_void_p(p))

    def empty(self, shape, dtype):
        import weakref
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        size = max(self.granule, -(-count * dtype.itemsize // self.granule) * self.granule)
        ptr = None
        with self.lock:
            best = -1
            for i, (sz, _) in enumerate(self.free):
                if size <= sz <= 2 * size and (best < 0 or sz < self.free[best][0]):
                    best = i
            if best >= 0:
                size, ptr = self.free.pop(best)
                self.kept -= size
        if ptr is None:
            p = ctypes.c_void_p()
            self.lib.check(self.lib.dll.nsb_alloc_pinned(ctypes.c_uint64(size), ctypes.byref(p)))
            ptr = p.value
        buf = (ctypes.c_char * size).from_address(ptr)
        base = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
        weakref.finalize(buf, self._release, ptr, size)          # every view keeps `buf` alive through .base
        return base


_default = None
_default_lock = threading.Lock()


def default_lib():
    global _default
    with _default_lock:
        if _default is None:
            _default = NativeLib()
        return _default


def _ptr(a, need=None, what="buffer"):
    """void* of a numpy array, a torch tensor, anything with __cuda_array_interface__ / __dlpack__, or a raw integer
    address.  ``need`` = bytes the native call will touch: arrays that say how large they are must be dense (C- or
    Fortran-ordered, no gaps) and at least that large - a short or strided buffer would otherwise be read or written
    past its end by cudaMemcpyAsync and the kernels (raw integer addresses cannot be checked: the caller vouches)."""
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    if isinstance(a, np.ndarray):
        if need is not None:
            if not (a.flags.c_contiguous or a.flags.f_contiguous):
                raise ValueError("%s must be contiguous (got strides %r for shape %r)" % (what, a.strides, a.shape))
            if a.nbytes < need:
                raise ValueError("%s holds %d bytes, the call needs %d" % (what, a.nbytes, need))
        return ctypes.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        if need is not None and hasattr(a, "numel"):
            if not (a.is_contiguous() or a.T.is_contiguous()):
                raise ValueError("%s must be contiguous" % what)
            if a.numel() * a.element_size() < need:
                raise ValueError("%s holds %d bytes, the call needs %d" % (what, a.numel() * a.element_size(), need))
        return ctypes.c_void_p(a.data_ptr())
    from . import _buffers
    b = _buffers.as_buffer(a)
    if need is not None:
        if not (b.c_contiguous or b.f_contiguous):
            raise ValueError("%s must be contiguous" % what)
        if b.nbytes < need:
            raise ValueError("%s holds %d bytes, the call needs %d" % (what, b.nbytes, need))
    _keep.refs = (getattr(_keep, "refs", ()) + (b,))[-16:]     # DLPack capsules must outlive the native call
    return ctypes.c_void_p(b.ptr)


_keep = threading.local()


class Handle(object):
    """One native handle = hparams + device + workspaces.  Calls on one handle serialise (internal mutex)."""

    def __init__(self, hp, device=0, lib=None):
        self.lib = default_lib() if lib is None else lib
        s = HParamsStruct(int(hp.num_freq), int(hp.num_mels), int(hp.sample_rate), int(hp.griffin_lim_iters),
                          float(hp.frame_shift_ms), float(hp.frame_length_ms), float(hp.preemphasis),
                          float(hp.ref_level_db), float(hp.min_level_db), float(hp.power))
        h = ctypes.c_void_p()
        self.lib.check(self.lib.dll.nsb_create(ctypes.byref(s), int(device), ctypes.byref(h)))
        self._h = h
        self.device = int(device)
        a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        self.lib.check(self.lib.dll.nsb_stft_parameters(h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        self.n_fft, self.hop, self.win = a.value, b.value, c.value
        self.num_freq = int(hp.num_freq)
        self.num_mels = int(hp.num_mels)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.dll.nsb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ----
    def num_frames(self, n):
        return 1 + int(n) // self.hop

    def num_samples(self, T):
        return self.hop * (int(T) - 1)

    @staticmethod
    def _lens(lengths, ctype):
        arr = (ctype * len(lengths))(*[int(v) for v in lengths])
        return arr

    def _call(self, name, *args):
        self.lib.check(getattr(self.lib.dll, name)(self._h, *args))

    # ---- raw entry points (buffers: numpy arrays for HOST; torch tensors, __cuda_array_interface__ / DLPack objects or
    #      int addresses for DEVICE).  Every buffer that knows its size is checked against what the native call touches. ----
    def _frames_total(self, n_samples):
        return sum(self.num_frames(n) for n in n_samples)

    def stft(self, wav, n_samples, out, preemphasis=False, space=HOST, stream=None):
        F, T = self.num_freq, self._frames_total(n_samples)
        self._call("nsb_stft", _ptr(wav, 4 * sum(n_samples), "wav"), self._lens(n_samples, ctypes.c_int64), len(n_samples), int(bool(preemphasis)),
                   _ptr(out, 8 * F * T, "out"), space, _ptr(stream))

    def stft_tf(self, wav, n_samples, out, space=HOST, stream=None):
        T = sum(max(0, self.num_frames_tf(n)) for n in n_samples)
        self._call("nsb_stft_tf", _ptr(wav, 4 * sum(n_samples), "wav"), self._lens(n_samples, ctypes.c_int64), len(n_samples),
                   _ptr(out, 8 * self.num_freq * T, "out"), space, _ptr(stream))

    def istft_tf(self, spec, layout, n_frames, out, space=HOST, stream=None):
        self._call("nsb_istft_tf", _ptr(spec, 8 * self.num_freq * sum(n_frames), "spec"), layout, self._lens(n_frames, ctypes.c_int32), len(n_frames),
                   _ptr(out, 4 * sum(self.num_samples_tf(t) for t in n_frames), "out"), space, _ptr(stream))

    def num_frames_tf(self, n):
        return 1 + (int(n) - self.win) // self.hop

    def num_samples_tf(self, T):
        return self.hop * (int(T) - 1) + self.win

    def features_padded(self, wav, n_samples, rows_per_utt, lin_out, mel_out, space=HOST, stream=None):
        rows = int(rows_per_utt) * len(n_samples)
        self._call("nsb_features_padded", _ptr(wav, 4 * sum(n_samples), "wav"), self._lens(n_samples, ctypes.c_int64), len(n_samples), int(rows_per_utt),
                   _ptr(lin_out, 4 * self.num_freq * rows, "lin_out"), _ptr(mel_out, 4 * self.num_mels * rows, "mel_out"), space, _ptr(stream))

    def features_rows(self, wav, n_samples, row_off, total_rows, lin_out, mel_out, space=HOST, stream=None):
        self._call("nsb_features_rows", _ptr(wav, 4 * sum(n_samples), "wav"), self._lens(n_samples, ctypes.c_int64), len(n_samples),
                   self._lens(row_off, ctypes.c_int64), int(total_rows), _ptr(lin_out, 4 * self.num_freq * int(total_rows), "lin_out"),
                   _ptr(mel_out, 4 * self.num_mels * int(total_rows), "mel_out"), space, _ptr(stream))

    def frame_energy(self, wav, n_samples, frame_length, hop_length, out, space=HOST, stream=None):
        self._call("nsb_frame_energy", _ptr(wav, 4 * sum(n_samples), "wav"), self._lens(n_samples, ctypes.c_int64), len(n_samples), int(frame_length),
                   int(hop_length), _ptr(out, 8 * sum(1 + n // int(hop_length) for n in n_samples), "out"), space, _ptr(stream))

    def find_endpoint(self, wav, n_samples, endpoints, dtype=F64, threshold_db=-40.0, min_silence_sec=0.8, space=HOST, stream=None):
        self._call("nsb_find_endpoint", _ptr(wav, (8 if dtype == F64 else 4) * sum(n_samples), "wav"), dtype, self._lens(n_samples, ctypes.c_int64),
                   len(n_samples), float(threshold_db), float(min_silence_sec), _ptr(endpoints, 8 * len(n_samples), "endpoints"), space, _ptr(stream))

    def synthesize(self, spec, n_frames, wav_out, endpoints, iters=-1, threshold_db=-40.0, min_silence_sec=0.8, space=HOST, stream=None,
                   flags=0, out_dtype=F64):
        n = sum(self.num_samples_tf(t) for t in n_frames)
        self._call("nsb_synthesize_ex", _ptr(spec, 4 * self.num_freq * sum(n_frames), "spec"), self._lens(n_frames, ctypes.c_int32), len(n_frames),
                   int(iters), float(threshold_db), float(min_silence_sec), int(flags), _ptr(wav_out, (2 if out_dtype == I16 else 8) * n, "wav_out"),
                   int(out_dtype), _ptr(endpoints, 8 * len(n_frames), "endpoints"), space, _ptr(stream))

    def peak_normalize(self, wav, n_samples, out, wav_dtype=F64, out_dtype=F64, limit=None, space=HOST, stream=None):
        n = sum(n_samples)
        lim = None if limit is None else (self._lens(limit, ctypes.c_int64) if space == HOST else _ptr(limit, 8 * len(n_samples), "limit"))
        self._call("nsb_peak_normalize", _ptr(wav, (8 if wav_dtype == F64 else 4) * n, "wav"), int(wav_dtype), self._lens(n_samples, ctypes.c_int64),
                   len(n_samples), lim, _ptr(out, (2 if out_dtype == I16 else 8) * n, "out"), int(out_dtype), space, _ptr(stream))

    def features(self, wav, n_samples, lin_out, mel_out, space=HOST, stream=None):
        T = self._frames_total(n_samples)
        self._call("nsb_features", _ptr(wav, 4 * sum(n_samples), "wav"), self._lens(n_samples, ctypes.c_int64), len(n_samples),
                   _ptr(lin_out, 4 * self.num_freq * T, "lin_out"), _ptr(mel_out, 4 * self.num_mels * T, "mel_out"), space, _ptr(stream))

    def istft(self, spec, layout, n_frames, out, space=HOST, stream=None):
        self._call("nsb_istft", _ptr(spec, 8 * self.num_freq * sum(n_frames), "spec"), layout, self._lens(n_frames, ctypes.c_int32), len(n_frames),
                   _ptr(out, 4 * sum(self.num_samples(t) for t in n_frames), "out"), space, _ptr(stream))

    def _gl_args(self, spec, layout, n_frames, out, init_phase, seed, iters, flags, out_dtype):
        FT = self.num_freq * sum(n_frames)
        n = sum((self.num_samples_tf(t) if flags & GL_TF_TWIN else self.num_samples(t)) for t in n_frames)
        return (_ptr(spec, 4 * FT, "spec"), layout, self._lens(n_frames, ctypes.c_int32), len(n_frames), _ptr(init_phase, 8 * FT, "init_phase"),
                ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), int(iters), int(flags), _ptr(out, (8 if out_dtype == F64 else 4) * n, "out"), out_dtype)

    def griffin_lim(self, spec, layout, n_frames, out, init_phase=None, seed=0, iters=-1, flags=0, out_dtype=F32,
                    space=HOST, stream=None):
        self._call("nsb_griffin_lim", *(self._gl_args(spec, layout, n_frames, out, init_phase, seed, iters, flags, out_dtype) + (space, _ptr(stream))))

    # ---- asynchronous NSB_HOST calls: submit -> ticket, wait(ticket).  The caller keeps the buffers alive until wait. ----
    def griffin_lim_submit(self, spec, layout, n_frames, out, init_phase=None, seed=0, iters=-1, flags=0, out_dtype=F32):
        t = ctypes.c_uint64(0)
        self._call("nsb_griffin_lim_submit", *(self._gl_args(spec, layout, n_frames, out, init_phase, seed, iters, flags, out_dtype) + (ctypes.byref(t),)))
        return t.value

    def features_submit(self, wav, n_samples, lin_out, mel_out):
        T = self._frames_total(n_samples)
        t = ctypes.c_uint64(0)
        self._call("nsb_features_submit", _ptr(wav, 4 * sum(n_samples), "wav"), self._lens(n_samples, ctypes.c_int64), len(n_samples),
                   _ptr(lin_out, 4 * self.num_freq * T, "lin_out"), _ptr(mel_out, 4 * self.num_mels * T, "mel_out"), ctypes.byref(t))
        return t.value

    def synthesize_submit(self, spec, n_frames, wav_out, endpoints, iters=-1, threshold_db=-40.0, min_silence_sec=0.8, flags=0, out_dtype=F64):
        n = sum(self.num_samples_tf(t) for t in n_frames)
        t = ctypes.c_uint64(0)
        self._call("nsb_synthesize_submit", _ptr(spec, 4 * self.num_freq * sum(n_frames), "spec"), self._lens(n_frames, ctypes.c_int32), len(n_frames),
                   int(iters), float(threshold_db), float(min_silence_sec), int(flags), _ptr(wav_out, (2 if out_dtype == I16 else 8) * n, "wav_out"),
                   int(out_dtype), _ptr(endpoints, 8 * len(n_frames), "endpoints"), ctypes.byref(t))
        return t.value

    def wait(self, ticket):
        self._call("nsb_wait", ctypes.c_uint64(int(ticket)))

    def set_async_slots(self, n):
        self._call("nsb_set_async_slots", int(n))

    def griffin_lim_iterate(self, iters, stream=None):
        self._call("nsb_griffin_lim_iterate", int(iters), _ptr(stream))

    def preemphasis(self, x, n_samples, out, out_dtype=F64, space=HOST, stream=None, inverse=False):
        n = sum(n_samples)
        self._call("nsb_inv_preemphasis" if inverse else "nsb_preemphasis", _ptr(x, 4 * n, "x"), self._lens(n_samples, ctypes.c_int64),
                   len(n_samples), _ptr(out, (8 if out_dtype == F64 else 4) * n, "out"), out_dtype, space, _ptr(stream))

    def linear_to_mel(self, spec, layout, n_frames, out, out_dtype=F64, space=HOST, stream=None):
        T = sum(n_frames)
        self._call("nsb_linear_to_mel", _ptr(spec, 4 * self.num_freq * T, "spec"), layout, self._lens(n_frames, ctypes.c_int32), len(n_frames),
                   _ptr(out, (8 if out_dtype == F64 else 4) * self.num_mels * T, "out"), out_dtype, space, _ptr(stream))

    def mel_basis(self):
        out = np.empty((self.num_mels, self.num_freq), dtype=np.float64)
        self._call("nsb_mel_basis", _ptr(out))
        return out

    def elementwise(self, op, x, out, space=HOST, stream=None, n=None):
        n = int(x.size if n is None else n)
        self._call("nsb_elementwise", int(op), _ptr(x, 4 * n, "x"), n, _ptr(out, 4 * n, "out"), space, _ptr(stream))

    def synchronize(self, stream=None):
        self._call("nsb_synchronize", _ptr(stream))

    def check_status(self, stream=None):
        self._call("nsb_check_status", _ptr(stream))

    def set_tile_hops(self, t):
        self._call("nsb_set_tile_hops", int(t))

    def set_host_chunks(self, n):
        self._call("nsb_set_host_chunks", int(n))

    def set_generic_iteration(self, on):
        self._call("nsb_set_generic_iteration", int(on))

    def set_stream_grid(self, n):
        self._call("nsb_set_stream_grid", int(n))

    def set_option(self, key, value):
        self._call("nsb_set_option", int(key), int(value))

    def stream_trace(self, enable=True, max_ctas=4096):
        """profiling hook: enable tracing / fetch (sm_id, start_ns, end_ns) per CTA of the last k_gl_stream launch"""
        import numpy as np
        buf = np.zeros((max_ctas, 3), dtype=np.uint64)
        n = self.lib.dll.nsb_stream_trace(self._h, int(bool(enable)), buf.ctypes.data_as(ctypes.c_void_p), max_ctas)
        if n < 0:
            raise RuntimeError("nsb_stream_trace failed")
        return buf[:n]

    def kernel_launches(self):
        return int(self.lib.dll.nsb_kernel_launches(self._h))
