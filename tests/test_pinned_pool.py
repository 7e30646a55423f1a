"""The page-locked result pool of nspeech_b200._lib (host logic only: a fake allocator stands in for cudaHostAlloc)."""
import ctypes
import gc

import numpy as np

from nspeech_b200 import _lib


class _FakeDll(object):
    def __init__(self):
        self.live, self.allocs, self.frees = {}, 0, 0

    def nsb_alloc_pinned(self, size, out):
        buf = ctypes.create_string_buffer(size.value)
        addr = ctypes.addressof(buf)
        self.live[addr] = buf
        out._obj.value = addr
        self.allocs += 1
        return 0

    def nsb_free_pinned(self, p):
        del self.live[p.value]
        self.frees += 1
        return 0


class _FakeLib(object):
    def __init__(self):
        self.dll = _FakeDll()

    def check(self, rc):
        assert rc == 0


def test_blocks_are_recycled_and_the_oldest_free_blocks_go_first():
    lib = _FakeLib()
    pool = _lib.PinnedPool(lib, max_bytes=8 << 20, granule=1 << 20)
    a = pool.empty((3 << 20,), np.uint8)
    a[:] = 7
    view = a[10:20]
    del a
    gc.collect()
    assert pool.kept == 0 and lib.dll.allocs == 1          # a view keeps the block out of the pool
    del view
    gc.collect()
    assert pool.kept == 3 << 20
    b = pool.empty((700000,), np.float32)                  # 2.8 MB -> the 3 MB block again (exact size class)
    assert lib.dll.allocs == 1 and pool.kept == 0 and b.shape == (700000,)
    c = pool.empty((2 << 20,), np.uint8)                   # no free block: a new one
    assert lib.dll.allocs == 2
    del b, c
    gc.collect()
    assert pool.kept == 5 << 20 and lib.dll.frees == 0
    d = pool.empty((1 << 20,), np.uint8)                   # 1 MB: the 2 MB block fits within a factor of two, the 3 MB one would not
    assert lib.dll.allocs == 2 and pool.kept == 3 << 20
    e = pool.empty((4 << 20,), np.uint8)
    f = pool.empty((4 << 20,), np.uint8)
    assert lib.dll.allocs == 4
    del e
    gc.collect()
    assert pool.kept == 7 << 20 and lib.dll.frees == 0
    del f                                                  # 11 MB would be kept: the oldest free block (3 MB) is given back
    gc.collect()
    assert lib.dll.frees == 1 and pool.kept == 8 << 20
    # a workload that moved on to another size keeps recycling its own blocks whatever else sits in the pool
    for _ in range(5):
        g = pool.empty((4 << 20,), np.uint8)
        del g
        gc.collect()
    assert lib.dll.allocs == 4 and lib.dll.frees == 1
    del d
    gc.collect()
    assert pool.kept <= 8 << 20 and len(lib.dll.live) == lib.dll.allocs - lib.dll.frees
