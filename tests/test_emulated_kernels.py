"""CPU checks of the ACTUAL kernel + C-ABI sources (nspeech_b200/csrc) through tests/emu's CUDA-on-CPU shim:
the same .cu/.cuh files are compiled with g++ -DNSB_EMULATE (one OS thread per CUDA thread, real barriers)
and driven through the same ctypes binding and Python mirror as on the GPU.  This is test infrastructure for
the GPU-less build container (indexing / tiling / synchronisation); the parity tests proper are the
``-m gpu`` ones.  Nothing here is a product code path."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, make_hp, speechlike
from nspeech_b200 import _lib, audio, hparams
from oracle import audio_oracle as ao

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "_build", "libnspeech_b200_emu.so")
SOURCES = [os.path.join(ROOT, "nspeech_b200", "csrc", f) for f in ("nspeech_b200.cu", "kernels.cuh", "frame_fft.cuh", "fft_core.cuh", "gl_iter.cuh", "gl_stream.cuh")] + \
          [os.path.join(EMU_DIR, f) for f in ("cuda_emu.h", "emu_runtime.cpp")]


@pytest.fixture(scope="module")
def emu():
    stale = (not os.path.exists(EMU_SO)) or any(os.path.getmtime(s) > os.path.getmtime(EMU_SO) for s in SOURCES)
    if stale:
        subprocess.check_call(["sh", os.path.join(EMU_DIR, "build_emu.sh")])
    lib = _lib.NativeLib(EMU_SO)
    old = _lib._default                # the product module has no override hook: the test swaps the loader's cached library
    _lib._default = lib
    yield lib
    _lib._default = old
    hparams.load()


import parity_checks as pc


def test_lane_level_fft(emu):
    import ctypes
    fp = ctypes.POINTER(ctypes.c_float)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(2048).astype(np.float32)
    ore, oim = np.zeros(1025, np.float32), np.zeros(1025, np.float32)
    emu.dll.emu_rfft2048(x.ctypes.data_as(fp), ore.ctypes.data_as(fp), oim.ctypes.data_as(fp))
    assert ao.rel_l2(ore + 1j * oim, np.fft.rfft(x.astype(np.float64))) < 5e-7
    X = (rng.standard_normal(1025) + 1j * rng.standard_normal(1025)).astype(np.complex64)
    xr = np.zeros(2048, np.float32)
    emu.dll.emu_irfft2048(np.ascontiguousarray(X.real).ctypes.data_as(fp), np.ascontiguousarray(X.imag).ctypes.data_as(fp),
                          xr.ctypes.data_as(fp))
    assert ao.rel_l2(xr, np.fft.irfft(X.astype(complex))) < 5e-7



@pytest.mark.parametrize("over", pc.CONFIGS)
def test_single_ops_vs_oracle(emu, over):
    pc.check_single_ops_vs_oracle(over)


@pytest.mark.parametrize("over", [c for c in pc.GENERIC_CONFIGS if c.get("num_freq") != 2049])      # (n_fft 4096, the slowest to emulate, runs on the GPU only)
def test_generic_num_freq_vs_oracle(emu, over):
    pc.check_single_ops_vs_oracle(over)
    pc.check_griffin_lim_vs_oracle(over)


def test_generic_num_freq_tf_twin_and_stages(emu):
    pc.check_generic_tf_twin_and_stages()


@pytest.mark.parametrize("over", pc.CONFIGS)
def test_griffin_lim_vs_oracle(emu, over):
    pc.check_griffin_lim_vs_oracle(over)


def test_golden_fixtures_through_kernels(emu, golden):
    pc.check_golden_fixtures_through_kernels(golden)


def test_ragged_batch_and_tiles(emu):
    pc.check_ragged_batch_and_tiles()


@pytest.mark.parametrize("over", pc.CONFIGS[:2])
def test_tf_twin_vs_oracle(emu, over):
    pc.check_tf_twin_vs_oracle(over)


@pytest.mark.parametrize("tf", [False, True])
def test_streaming_rounds(emu, tf):
    pc.check_streaming_rounds(T=110, iters=2, tf=tf, emulated=True)


def test_find_endpoint_and_synthesis_stage(emu, golden):
    pc.check_find_endpoint_and_synthesis_stage(golden)


def test_trimming(emu, golden_process):
    pc.check_trimming(golden_process)


def test_feeder_targets(emu):
    pc.check_feeder_targets()


def test_errors_and_edge_cases(emu):
    pc.check_errors_and_edge_cases()


def test_device_random_phase_is_deterministic_per_seed(emu):
    pc.check_device_random_phase_is_deterministic_per_seed()


def test_pinned_result_pool_lifecycle(emu):
    """batch.* hand out arrays backed by pooled page-locked blocks: a block goes back to the pool when the LAST view of it
    dies, and the next call reuses it."""
    import gc
    from nspeech_b200 import batch
    hparams.load()
    specs = np.random.RandomState(0).rand(3, 180, 1025).astype(np.float32)      # results just over the 1 MB from which the pool is used
    pool = audio._handle().lib.pinned_pool()
    gc.collect()
    kept0 = pool.kept
    outs = batch.inv_spectrogram_batch(specs, seed=1, iters=1)
    want = outs[2].copy()
    keep = outs[2]
    del outs
    gc.collect()
    assert pool.kept == kept0                      # one view is still alive
    np.testing.assert_array_equal(keep, want)
    del keep
    gc.collect()
    assert pool.kept > kept0                       # the block came back
    again = batch.inv_spectrogram_batch(specs, seed=1, iters=1)
    assert pool.kept == kept0                      # ... and was reused
    np.testing.assert_array_equal(again[2], want)


def test_save_wav_scaling_and_int16(emu):
    pc.check_save_wav_scaling_and_int16()


def test_feeder_groups(emu):
    pc.check_feeder_groups()


def test_async_submit_wait(emu):
    pc.check_async_submit_wait()


def test_api_guards(emu):
    pc.check_api_guards()


def test_features_pipeline(emu):
    pc.check_features_pipeline()


def test_mel_projection_modes(emu):
    pc.check_mel_projection_modes()


def test_stale_griffin_lim_state(emu):
    pc.check_stale_griffin_lim_state(lambda a: a)          # the emulated library's "device" pointers are host pointers


def test_tf_twin_golden(emu, golden_tf):
    pc.check_tf_twin_golden(golden_tf)


def test_length_and_hparams_sweep(emu):
    pc.check_length_and_hparams_sweep([(1, 20000, 50, 1025), (700, 22050, 50, 1025), (1024, 20000, 50, 1025), (1025, 24000, 50, 1025),
                                       (3000, 16000, 50, 1025), (2500, 20000, 25, 1025), (900, 16000, 50, 513), (4100, 20000, 120, 1025)])
