"""Array interchange at the Python boundary: numpy (host), ``__cuda_array_interface__`` and DLPack (device or host).

The C ABI takes plain pointers (include/nspeech_b200.h); this module turns whatever array object the caller holds into
(pointer, shape, dtype, strides, where it lives) without copying, and wraps library-owned device results in an object
other frameworks can adopt without copying (``torch.from_dlpack(r)``, ``cupy.asarray(r)``, ``numba.cuda.as_cuda_array(r)``).
Nothing here touches the data: no framework is imported unless the caller handed in one of its arrays.

Natural caller: ``models/tacotron.py:98,107`` of the reference - ``linear_outputs`` [N, T_out, num_freq] already lives on a GPU
when ``inv_spectrogram_tensorflow`` is applied to it.
"""
import ctypes

import numpy as np

kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged = 1, 2, 3, 13
_DL_CODES = {0: "i", 1: "u", 2: "f", 5: "c"}          # DLDataTypeCode: int, uint, float, complex


class DLDevice(ctypes.Structure):
    _fields_ = [("device_type", ctypes.c_int32), ("device_id", ctypes.c_int32)]


class DLDataType(ctypes.Structure):
    _fields_ = [("code", ctypes.c_uint8), ("bits", ctypes.c_uint8), ("lanes", ctypes.c_uint16)]


class DLTensor(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("device", DLDevice), ("ndim", ctypes.c_int32), ("dtype", DLDataType),
                ("shape", ctypes.POINTER(ctypes.c_int64)), ("strides", ctypes.POINTER(ctypes.c_int64)),
                ("byte_offset", ctypes.c_uint64)]


class DLManagedTensor(ctypes.Structure):
    pass


_DELETER = ctypes.CFUNCTYPE(None, ctypes.POINTER(DLManagedTensor))
DLManagedTensor._fields_ = [("dl_tensor", DLTensor), ("manager_ctx", ctypes.c_void_p), ("deleter", _DELETER)]

_capi = ctypes.pythonapi
_capi.PyCapsule_GetPointer.restype = ctypes.c_void_p
_capi.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
_capi.PyCapsule_IsValid.restype = ctypes.c_int
_capi.PyCapsule_IsValid.argtypes = [ctypes.py_object, ctypes.c_char_p]
_capi.PyCapsule_New.restype = ctypes.py_object
_capi.PyCapsule_New.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]


# raw-pointer prototypes for use inside a capsule destructor (the capsule's refcount is 0 there: no py_object conversions)
_IsValidRaw = ctypes.PYFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p)(("PyCapsule_IsValid", ctypes.pythonapi))
_GetPointerRaw = ctypes.PYFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p)(("PyCapsule_GetPointer", ctypes.pythonapi))
_CAPSULE_DTOR = ctypes.PYFUNCTYPE(None, ctypes.c_void_p)


def _capsule_dropped(cap):
    # a capsule nobody consumed (still named "dltensor") is collected: release what __dlpack__ pinned for it
    if _IsValidRaw(cap, b"dltensor"):
        DeviceArray._live.pop(_GetPointerRaw(cap, b"dltensor"), None)


_capsule_dropped_c = _CAPSULE_DTOR(_capsule_dropped)


class Buf(object):
    """A borrowed view of somebody's array: ``ptr`` is valid as long as ``owner`` is alive."""
    __slots__ = ("ptr", "shape", "dtype", "strides", "on_device", "device", "owner", "stream")

    def __init__(self, ptr, shape, dtype, strides, on_device, device, owner, stream=None):
        self.ptr, self.shape, self.dtype = int(ptr or 0), tuple(int(v) for v in shape), np.dtype(dtype)
        if strides is None:                              # C-contiguous
            strides, acc = [], self.dtype.itemsize
            for n in reversed(self.shape):
                strides.append(acc)
                acc *= max(n, 1)
            strides = tuple(reversed(strides))
        self.strides = tuple(int(v) for v in strides)    # bytes
        self.on_device, self.device, self.owner, self.stream = bool(on_device), int(device), owner, stream

    @property
    def size(self):
        return int(np.prod(self.shape)) if self.shape else 1

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def ndim(self):
        return len(self.shape)

    def _dense(self, order):
        acc = self.dtype.itemsize
        dims = list(zip(self.shape, self.strides))
        for n, st in (reversed(dims) if order == "C" else dims):
            if n > 1 and st != acc:
                return False
            acc *= max(n, 1)
        return True

    @property
    def c_contiguous(self):
        return self._dense("C")

    @property
    def f_contiguous(self):
        return self._dense("F")


def is_device_array(a):
    """True for anything that says it lives in CUDA device memory (CUDA array interface or a DLPack CUDA device)."""
    if a is None or isinstance(a, (np.ndarray, int, float, list, tuple)):
        return False
    if hasattr(a, "__cuda_array_interface__"):
        return True
    dev = getattr(a, "__dlpack_device__", None)
    if dev is not None:
        try:
            return int(dev()[0]) in (kDLCUDA, kDLCUDAManaged)
        except Exception:
            return False
    return False


def _from_cai(a):
    d = a.__cuda_array_interface__
    ptr = d["data"][0]
    dev = getattr(getattr(a, "device", None), "index", None)
    if dev is None:
        dev = getattr(getattr(a, "device", None), "id", 0) or 0
    st = d.get("stream")
    return Buf(ptr, d["shape"], np.dtype(d["typestr"]), d.get("strides"), True, dev, a, stream=st if st not in (None, 1, 2) else None)


def _from_dlpack(a):
    try:
        cap = a.__dlpack__()
    except TypeError:
        cap = a.__dlpack__(stream=None)
    if not _capi.PyCapsule_IsValid(cap, b"dltensor"):
        raise TypeError("__dlpack__ did not return a 'dltensor' capsule")
    mt = ctypes.cast(_capi.PyCapsule_GetPointer(cap, b"dltensor"), ctypes.POINTER(DLManagedTensor)).contents
    t = mt.dl_tensor
    if t.dtype.lanes != 1 or t.dtype.code not in _DL_CODES:
        raise TypeError("unsupported DLPack dtype (code %d, %d bits, %d lanes)" % (t.dtype.code, t.dtype.bits, t.dtype.lanes))
    dtype = np.dtype("%s%d" % (_DL_CODES[t.dtype.code], t.dtype.bits // 8))
    shape = [t.shape[i] for i in range(t.ndim)]
    strides = [t.strides[i] * dtype.itemsize for i in range(t.ndim)] if t.strides else None
    if t.device.device_type not in (kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged):
        raise TypeError("DLPack device type %d is neither host nor CUDA memory" % t.device.device_type)
    on_dev = t.device.device_type in (kDLCUDA, kDLCUDAManaged)
    # the capsule stays un-consumed (still named "dltensor"): its own destructor releases the producer's tensor when the
    # Buf (which holds it) goes away
    return Buf((t.data or 0) + t.byte_offset, shape, dtype, strides, on_dev, t.device.device_id if on_dev else 0, (a, cap))


def as_buffer(a):
    """numpy array / CUDA-array-interface object / DLPack producer -> Buf (no copy)."""
    if isinstance(a, Buf):
        return a
    if isinstance(a, np.ndarray):
        return Buf(a.ctypes.data, a.shape, a.dtype, a.strides, False, 0, a)
    if hasattr(a, "__cuda_array_interface__"):
        return _from_cai(a)
    if hasattr(a, "__dlpack__"):
        return _from_dlpack(a)
    if hasattr(a, "__array_interface__") or hasattr(a, "__array__"):
        a = np.asarray(a)
        return Buf(a.ctypes.data, a.shape, a.dtype, a.strides, False, 0, a)
    raise TypeError("unsupported array type %r (numpy, __cuda_array_interface__ or __dlpack__ expected)" % type(a))


class DeviceArray(object):
    """Device memory owned by the library (``nsb_device_alloc``), returned when a caller passed device arrays of a framework
    the library knows nothing about.  Zero-copy hand-over through ``__cuda_array_interface__`` (v3) and ``__dlpack__``;
    ``copy_to_host()`` gives numpy.  Freed when the last reference (including exported DLPack capsules) is gone."""
    _live = {}        # id -> (DeviceArray, ctypes keep-alives) while a DLPack consumer holds the memory

    def __init__(self, lib, shape, dtype, device, order="C"):
        self.lib, self.shape, self.dtype, self.device = lib, tuple(int(v) for v in shape), np.dtype(dtype), int(device)
        self.order = order
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        p = ctypes.c_void_p()
        lib.check(lib.dll.nsb_device_alloc(self.device, ctypes.c_uint64(max(n, 1)), ctypes.byref(p)))
        self._p = p
        self.nbytes = n

    @property
    def ptr(self):
        return self._p.value or 0

    @property
    def strides(self):
        it, dims = self.dtype.itemsize, list(self.shape)
        out, acc = [], it
        for n in (reversed(dims) if self.order == "C" else dims):
            out.append(acc)
            acc *= max(n, 1)
        return tuple(reversed(out)) if self.order == "C" else tuple(out)

    @property
    def T(self):
        """transposed VIEW (shares the memory; keeps this object alive)"""
        v = DeviceArray.__new__(DeviceArray)
        v.lib, v.shape, v.dtype, v.device, v.order = self.lib, self.shape[::-1], self.dtype, self.device, "F" if self.order == "C" else "C"
        v._p, v.nbytes, v._base = ctypes.c_void_p(self.ptr), self.nbytes, self
        return v

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False), "version": 3,
                "strides": None if self.order == "C" else self.strides, "stream": None}

    def __dlpack_device__(self):
        return (kDLCUDA, self.device)

    def __dlpack__(self, stream=None, **kw):
        nd = len(self.shape)
        shape = (ctypes.c_int64 * max(nd, 1))(*self.shape)
        strides = (ctypes.c_int64 * max(nd, 1))(*[s // self.dtype.itemsize for s in self.strides])
        mt = DLManagedTensor()
        code = {"i": 0, "u": 1, "f": 2, "c": 5}[self.dtype.kind]
        mt.dl_tensor = DLTensor(ctypes.c_void_p(self.ptr), DLDevice(kDLCUDA, self.device), nd, DLDataType(code, self.dtype.itemsize * 8, 1),
                                ctypes.cast(shape, ctypes.POINTER(ctypes.c_int64)), ctypes.cast(strides, ctypes.POINTER(ctypes.c_int64)), 0)
        key = ctypes.addressof(mt)

        def _del(_p, key=key):
            DeviceArray._live.pop(key, None)
        deleter = _DELETER(_del)
        mt.manager_ctx, mt.deleter = None, deleter
        DeviceArray._live[key] = (self, mt, shape, strides, deleter)
        return _capi.PyCapsule_New(key, b"dltensor", ctypes.cast(_capsule_dropped_c, ctypes.c_void_p))

    def copy_to_host(self):
        out = np.empty(self.shape, dtype=self.dtype, order=self.order)
        self.lib.check(self.lib.dll.nsb_device_copy(self.device, ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(self.ptr),
                                                    ctypes.c_uint64(self.nbytes), 2))
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.copy_to_host()
        return a if dtype is None else a.astype(dtype)

    def __del__(self):
        try:
            if getattr(self, "_base", None) is None and self._p is not None and self._p.value:
                self.lib.dll.nsb_device_free(self.device, self._p)
                self._p = None
        except Exception:
            pass


def _torch_of(*arrays):
    """the torch module if any of the arrays is a torch tensor (then results are torch tensors), else None"""
    for a in arrays:
        mod = type(a).__module__
        if mod == "torch" or mod.startswith("torch."):
            import torch
            return torch
    return None


def empty_like_source(lib, shape, dtype, device, sources, order="C"):
    """device result buffer: a torch tensor when the caller works in torch (its allocator, its stream semantics), else a DeviceArray"""
    torch = _torch_of(*sources)
    if torch is not None:
        tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.complex64): torch.complex64,
               np.dtype(np.int64): torch.int64, np.dtype(np.int16): torch.int16}[np.dtype(dtype)]
        t = torch.empty(tuple(shape) if order == "C" else tuple(shape)[::-1], dtype=tdt, device=torch.device("cuda", device))
        return t if order == "C" else t.T
    return DeviceArray(lib, shape, dtype, device, order=order)


def stream_of(*sources):
    """the CUDA stream the caller's arrays are being produced on: torch's current stream, else what the array interface names, else NULL"""
    torch = _torch_of(*sources)
    if torch is not None:
        dev = next((a.device for a in sources if hasattr(a, "is_cuda") and a.is_cuda), None)
        return int(torch.cuda.current_stream(dev).cuda_stream)
    for a in sources:
        d = getattr(a, "__cuda_array_interface__", None)
        if d and isinstance(d.get("stream"), int) and d["stream"] > 2:
            return d["stream"]
    return None
