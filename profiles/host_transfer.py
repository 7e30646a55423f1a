"""Host <-> device transfer rates that bound the end-to-end numbers (run on the GPU box):
pinned vs pageable cudaMemcpyAsync at the bench's sizes (262 MB in, 128 MB out), and what cudaHostRegister costs."""
import ctypes
import sys
import time

import numpy as np
import torch

rt = torch.cuda.cudart()
dev = torch.device("cuda")
n_in, n_out = 64 * 1000 * 1025, 64 * 999 * 250
d_in = torch.empty(n_in, dtype=torch.float32, device=dev)
d_out = torch.empty(n_out, dtype=torch.float64, device=dev)
pin_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
pin_out = torch.empty(n_out, dtype=torch.float64).pin_memory()
pag_in = torch.from_numpy(np.random.rand(n_in).astype(np.float32))
pag_out = torch.empty(n_out, dtype=torch.float64)


def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for name, fn, nbytes in (("H2D pinned", lambda: d_in.copy_(pin_in, non_blocking=True), n_in * 4), ("H2D pageable", lambda: d_in.copy_(pag_in, non_blocking=True), n_in * 4),
                         ("D2H pinned", lambda: pin_out.copy_(d_out, non_blocking=True), n_out * 8), ("D2H pageable", lambda: pag_out.copy_(d_out, non_blocking=True), n_out * 8)):
    ms = t(fn)
    print("%-14s %7.2f ms  %6.1f GB/s" % (name, ms, nbytes / ms / 1e6))
# cudaHostRegister / Unregister of the pageable input
t0 = time.perf_counter()
rc = rt.cudaHostRegister(pag_in.data_ptr(), n_in * 4, 0)
t1 = time.perf_counter()
ms = t(lambda: d_in.copy_(pag_in, non_blocking=True))
t2 = time.perf_counter()
rc2 = rt.cudaHostUnregister(pag_in.data_ptr())
t3 = time.perf_counter()
print("cudaHostRegister(262 MB): %.2f ms (rc %s); copy from the registered buffer %.2f ms (%.1f GB/s); unregister %.2f ms" % (
    (t1 - t0) * 1e3, rc, ms, n_in * 4 / ms / 1e6, (t3 - t2) * 1e3))
# host memcpy into a pinned bounce buffer (one thread)
src, dst = pag_in.numpy(), pin_in.numpy()
t0 = time.perf_counter()
for _ in range(3):
    np.copyto(dst, src)
print("host memcpy pageable -> pinned, 1 thread: %.2f ms per 262 MB (%.1f GB/s)" % ((time.perf_counter() - t0) / 3 * 1e3, n_in * 4 / ((time.perf_counter() - t0) / 3) / 1e9))
