"""The asynchronous Griffin-Lim stream on every GPU of the box at once (one rank per GPU under torchrun, bench workload per rank:
64 x 1000 frames, page-locked inputs), waveforms back as float64 (the reference's dtype) against float32: what fewer bytes from the
device buy when the ranks share one host.  Usage: python -m torch.distributed.run --nproc-per-node N profiles/e2e_multi_gpu_dtype.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nspeech_b200 import _lib, audio, batch, hparams  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
audio.DEVICE = local
hparams.load()
N, T, F, STEPS = 64, 1000, 1025, 20
rs = np.random.default_rng(rank)
pinned = [_lib.PinnedArray((N, T, F), np.float32) for _ in range(3)]
for p in pinned:
    p.array[...] = rs.random((N, T, F), dtype=np.float32)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(dtype, n):
    for wavs in batch.inv_spectrogram_stream((pinned[i % 3].array for i in range(n)), seed=1, dtype=dtype):
        del wavs


for dtype in (np.float64, np.float32, np.float64, np.float32):
    run(dtype, 5)
    barrier()
    t0 = time.perf_counter()
    run(dtype, STEPS)
    torch.cuda.synchronize()
    ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / STEPS], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = world * N * 250 * (T - 1) / 20000.0
        print("%d GPUs, %s waveforms back: %.2f ms per step (max over ranks), %.0f audio-s/s" % (world, np.dtype(dtype).name, float(ms), sec / (float(ms) * 1e-3)), flush=True)
if world > 1:
    dist.destroy_process_group()
