"""Multi-GPU path = utterance sharding with no data-path collective.  Host-side logic is
covered here on CPU: the shard assignment, the threaded per-device runner, and a world_size-2 gloo run in which
every rank inverts only its shard (through the CPU-emulated build of the kernels, test infrastructure) and the
gathered result must equal the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from nspeech_b200 import batch


def test_shard_by_frames_balanced_and_deterministic():
    rs = np.random.RandomState(0)
    lens = list(rs.randint(80, 801, size=101))
    for world in (1, 2, 4, 8):
        shards = batch.shard_by_frames(lens, world)
        assert sorted(i for s in shards for i in s) == list(range(len(lens)))
        loads = [sum(lens[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lens)
        assert shards == batch.shard_by_frames(lens, world)
    assert batch.shard_by_frames([5, 3], 4) == [[0], [1], [], []]


def test_run_sharded_restores_order_and_propagates_errors():
    items = list(range(10))
    out = batch.run_sharded(lambda sub, dev: [(x, dev) for x in sub], items, [10 - i for i in items], devices=[0, 1, 2])
    assert [o[0] for o in out] == items and {o[1] for o in out} == {0, 1, 2}

    def boom(sub, dev):
        if dev == 1:
            raise RuntimeError("device 1 failed")
        return sub
    with pytest.raises(RuntimeError):
        batch.run_sharded(boom, items, [1] * 10, devices=[0, 1])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_batch():
    rs = np.random.RandomState(42)
    Ts = [7, 12, 5, 9, 3]
    specs = [rs.rand(1025, T).astype(np.float32) for T in Ts]
    phases = [np.exp(2j * np.pi * rs.rand(1025, T)).astype(np.complex64) for T in Ts]
    return Ts, specs, phases


def _rank_main(rank, world, port, emu_so, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from nspeech_b200 import _lib, audio, batch as b, hparams
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _lib._default = _lib.NativeLib(emu_so)          # test process only: the loader's cached library is the CPU-emulated build
    hparams.load().parse("min_level_db=-100")
    Ts, specs, phases = _make_batch()
    mine = b.shard_by_frames(Ts, world)[rank]
    outs = b.inv_spectrogram_batch([specs[i] for i in mine], init_phase=[phases[i] for i in mine], iters=2) if mine else []
    dist.barrier()                       # the only collective of the multi-GPU path: timing fences
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, {i: o.copy() for i, o in zip(mine, outs)})
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        q.put(merged)
    dist.destroy_process_group()


def test_two_rank_gloo_shards_equal_single_process():
    import test_emulated_kernels as tek
    from nspeech_b200 import _lib, audio, hparams
    stale = (not os.path.exists(tek.EMU_SO)) or any(os.path.getmtime(s) > os.path.getmtime(tek.EMU_SO) for s in tek.SOURCES)
    if stale:
        import subprocess
        subprocess.check_call(["sh", os.path.join(tek.EMU_DIR, "build_emu.sh")])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, tek.EMU_SO, q)) for r in range(2)]
    [p.start() for p in procs]
    merged = q.get(timeout=300)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    old = _lib._default
    _lib._default = _lib.NativeLib(tek.EMU_SO)
    try:
        hparams.load().parse("min_level_db=-100")
        Ts, specs, phases = _make_batch()
        single = batch.inv_spectrogram_batch(specs, init_phase=phases, iters=2)
    finally:
        _lib._default = old
        hparams.load()
    assert sorted(merged) == list(range(len(Ts)))
    for i, ref in enumerate(single):
        np.testing.assert_array_equal(merged[i], ref)


def test_bind_host_to_gpu_is_harmless_without_topology():
    """bind_host_to_gpu pins the calling thread to the CPUs next to a GPU (sysfs local_cpulist of its PCI device); with no
    device or no topology information it must change nothing and return None."""
    import os
    assert batch._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert batch._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    try:
        cpus = batch.bind_host_to_gpu(0)
        if cpus is None:                              # no GPU (the build container) or no topology: nothing changed
            assert os.sched_getaffinity(0) == before
        else:                                         # on a GPU box: a non-empty subset of what was allowed
            assert cpus and set(cpus) <= before and os.sched_getaffinity(0) == set(cpus)
    finally:
        os.sched_setaffinity(0, before)
