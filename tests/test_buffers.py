"""Array interchange at the Python boundary (nspeech_b200/_buffers.py): numpy, DLPack and the CUDA array interface are
read without copying; library-owned device results export both protocols.  CPU part: host-side producers and the
protocol plumbing (the device part runs under -m gpu in tests/test_gpu_parity.py)."""
import ctypes

import numpy as np
import pytest

from nspeech_b200 import _buffers


def test_numpy_and_dlpack_host_producers_are_borrowed_not_copied():
    a = np.arange(24, dtype=np.float32).reshape(4, 6)
    b = _buffers.as_buffer(a)
    assert (b.ptr, b.shape, b.dtype, b.on_device) == (a.ctypes.data, (4, 6), np.dtype(np.float32), False)
    assert b.c_contiguous and not b.f_contiguous and b.nbytes == a.nbytes
    bt = _buffers.as_buffer(a.T)
    assert bt.f_contiguous and not bt.c_contiguous and bt.ptr == a.ctypes.data
    assert not _buffers.as_buffer(a[:, ::2]).c_contiguous and not _buffers.as_buffer(a[:, ::2]).f_contiguous
    torch = pytest.importorskip("torch")
    t = torch.arange(10, dtype=torch.float64)
    d = _buffers._from_dlpack(t)                      # torch CPU tensors speak DLPack (kDLCPU)
    assert (d.ptr, d.shape, d.dtype, d.on_device) == (t.data_ptr(), (10,), np.dtype(np.float64), False)
    c = torch.zeros((3, 5), dtype=torch.complex64).T
    dc = _buffers._from_dlpack(c)
    assert dc.shape == (5, 3) and dc.dtype == np.dtype(np.complex64) and dc.f_contiguous
    assert not _buffers.is_device_array(a) and not _buffers.is_device_array(t) and not _buffers.is_device_array(None)


class _FakeCuda(object):
    """what a CUDA array interface producer looks like (the pointer is never dereferenced here)"""

    def __init__(self, shape, typestr, ptr=0x7f0000000000, strides=None):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3, "strides": strides}


def test_cuda_array_interface_is_recognised():
    f = _FakeCuda((7, 1025), "<f4")
    assert _buffers.is_device_array(f)
    b = _buffers.as_buffer(f)
    assert b.on_device and b.shape == (7, 1025) and b.dtype == np.dtype(np.float32) and b.c_contiguous and b.ptr == 0x7f0000000000
    g = _buffers.as_buffer(_FakeCuda((1025, 7), "<f4", strides=(4, 4100)))
    assert g.f_contiguous and not g.c_contiguous
    with pytest.raises(TypeError):
        _buffers.as_buffer(object())


def test_exported_dlpack_capsule_round_trips():
    """DeviceArray.__dlpack__ builds the DLManagedTensor by hand: read it back with the importer and run its deleter."""
    class _Lib(object):
        class dll(object):
            @staticmethod
            def nsb_device_alloc(dev, n, out):
                ctypes.cast(out, ctypes.POINTER(ctypes.c_void_p))[0] = 0x7e0000001000
                return 0

            @staticmethod
            def nsb_device_free(dev, p):
                return 0

        @staticmethod
        def check(rc):
            assert rc == 0
    arr = _buffers.DeviceArray(_Lib, (3, 4), np.float64, 0)
    assert arr.__dlpack_device__() == (_buffers.kDLCUDA, 0)
    cai = arr.__cuda_array_interface__
    assert cai["shape"] == (3, 4) and cai["typestr"] == "<f8" and cai["data"][0] == arr.ptr and cai["strides"] is None
    back = _buffers._from_dlpack(arr)
    assert back.on_device and back.shape == (3, 4) and back.dtype == np.dtype(np.float64) and back.ptr == arr.ptr and back.c_contiguous
    tb = _buffers._from_dlpack(arr.T)
    assert tb.shape == (4, 3) and tb.f_contiguous and tb.ptr == arr.ptr
    assert arr.T.__cuda_array_interface__["strides"] == (8, 32)
    import gc
    del back, tb
    gc.collect()
    assert len(_buffers.DeviceArray._live) == 0       # the capsules read back above were never consumed: dropping them released the pins
    cap = arr.__dlpack__()
    assert len(_buffers.DeviceArray._live) == 1
    mt = ctypes.cast(_buffers._capi.PyCapsule_GetPointer(cap, b"dltensor"), ctypes.POINTER(_buffers.DLManagedTensor))
    n = len(_buffers.DeviceArray._live)
    mt.contents.deleter(mt)                           # what a consumer does when it is done with the memory
    assert len(_buffers.DeviceArray._live) == n - 1
