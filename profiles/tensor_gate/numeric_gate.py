#!/usr/bin/env python
"""Numeric go/no-go gate for moving the DFT passes of the Griffin-Lim iteration kernel onto tcgen05.

Zero GPU minutes: a numpy model of the 2048-point frame transform as TWO dense DFT passes (64 x 32, the decomposition of
csrc/frame_fft.cuh) whose operands are what a tensor-core pass would see:

  * the constant DFT matrices as fp16 hi + lo pairs (22 significant bits);
  * the data, after a per-frame power-of-two scale, as ONE fp16 term (11 bits) or as hi + lo (22 bits);
  * products exact, three partial products (hi*hi + hi*lo + lo*hi), fp32 accumulation (TMEM), fp32 twiddles between the passes.

It runs the reference's Griffin-Lim loop (neural_speech/utils/audio.py:77-87 on librosa 0.6.0 semantics) with that transform,
from the same supplied initial phase, and reports the BASELINE.json bars against the float64 loop: single-op relative L2
(bar 1e-5) and waveform SNR after `iters` iterations (bar 40 dB), next to an fp32-FFT loop (what the CUDA-core kernel does).

    python profiles/tensor_gate/numeric_gate.py            # full gate: 60 iterations, the three input families
    python profiles/tensor_gate/numeric_gate.py --quick    # 12 iterations (what tests/test_tensor_gate.py runs)
"""
import argparse
import json
import sys

import numpy as np
import scipy.fft

N, N1, N2, HOP, WIN = 2048, 64, 32, 250, 1000


def _split16(x, terms):
    """x (already scaled to |x| <= 1) -> (hi, lo) as float64 holding fp16-representable values"""
    hi = x.astype(np.float16).astype(np.float64)
    if terms == 1:
        return hi, None
    lo = (x - hi).astype(np.float16).astype(np.float64)
    return hi, lo


def _frame_scale(x):
    m = np.max(np.abs(x), axis=tuple(range(1, x.ndim)), keepdims=True)
    return np.exp2(np.ceil(np.log2(np.maximum(m, 1e-300))))


class SplitDft(object):
    """y[..., k, :] = sum_n W[k, n] x[..., n, :] with split-fp16 operands (complex W, complex or real x)"""

    def __init__(self, W, data_terms):
        self.data_terms = data_terms
        self.Wr = _split16(W.real.copy(), 2)
        self.Wi = _split16(W.imag.copy(), 2)

    def _mm(self, Wp, xp):
        (wh, wl), (xh, xl) = Wp, xp
        acc = np.einsum("kn,tnm->tkm", wh, xh) + np.einsum("kn,tnm->tkm", wl, xh)
        if xl is not None:
            acc = acc + np.einsum("kn,tnm->tkm", wh, xl)
        return acc

    def __call__(self, x):
        """x: [T, n, m] real or complex -> [T, k, m] complex64-rounded"""
        s = _frame_scale(np.concatenate([x.real, x.imag], axis=1) if np.iscomplexobj(x) else x)
        xr = _split16(np.real(x) / s, self.data_terms)
        rr = self._mm(self.Wr, xr)
        ir = self._mm(self.Wi, xr)
        if np.iscomplexobj(x):
            xi = _split16(np.imag(x) / s, self.data_terms)
            re, im = rr - self._mm(self.Wi, xi), ir + self._mm(self.Wr, xi)
        else:
            re, im = rr, ir
        # fp32 accumulators
        return ((re * s).astype(np.float32) + 1j * (im * s).astype(np.float32)).astype(np.complex128)


class TwoPassTransform(object):
    """rfft / irfft of 2048 as 64 x 32 with n = 32 n1 + n2, k = k1 + 64 k2 (all 64 k1 rows: the Hermitian half is a
    layout matter, not a numerical one)"""

    def __init__(self, data_terms):
        n1, n2 = np.arange(N1), np.arange(N2)
        W64 = np.exp(-2j * np.pi * np.outer(n1, n1) / N1)
        W32 = np.exp(-2j * np.pi * np.outer(n2, n2) / N2)
        self.tw = np.exp(-2j * np.pi * np.outer(n1, n2) / N).astype(np.complex64)        # [k1, n2], fp32 like the kernel's table
        self.f1, self.f2 = SplitDft(W64, data_terms), SplitDft(W32, data_terms)
        self.i1, self.i2 = SplitDft(W64.conj(), data_terms), SplitDft(W32.conj(), data_terms)

    def rfft(self, frames):
        T = frames.shape[0]
        x = frames.reshape(T, N1, N2)
        U = self.f1(x)                                                    # [T, k1, n2]
        V = (U.astype(np.complex64) * self.tw[None]).astype(np.complex128)
        X = self.f2(np.swapaxes(V, 1, 2))                                 # contraction over n2: [T, k2, k1]
        return X.reshape(T, N)[:, :N // 2 + 1]                            # k = k1 + 64 k2

    def irfft(self, spec):
        T = spec.shape[0]
        full = np.concatenate([spec, np.conj(spec[:, -2:0:-1])], axis=1)  # Hermitian extension
        full[:, 0] = full[:, 0].real
        full[:, N // 2] = full[:, N // 2].real
        P = full.reshape(T, N2, N1)                                       # [T, k2, k1]
        Z = self.i2(P)                                                    # contraction over k2: [T, n2, k1]
        Z = np.swapaxes(Z, 1, 2)                                          # [T, k1, n2]
        Z = (Z.astype(np.complex64) * np.conj(self.tw)[None]).astype(np.complex128)
        x = self.i1(Z)                                                    # contraction over k1: [T, n1, n2]
        return (x.real / N).reshape(T, N)


class Fp32Transform(object):
    def rfft(self, frames):
        return scipy.fft.rfft(frames.astype(np.float32), axis=1).astype(np.complex128)

    def irfft(self, spec):
        return scipy.fft.irfft(spec.astype(np.complex64), n=N, axis=1).astype(np.float64)


class Fp64Transform(object):
    def rfft(self, frames):
        return np.fft.rfft(frames, axis=1)

    def irfft(self, spec):
        return np.fft.irfft(spec, n=N, axis=1)


def _window():
    w = np.zeros(N)
    w[(N - WIN) // 2:(N - WIN) // 2 + WIN] = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(WIN) / WIN)
    return w


def stft(y, tr):
    """librosa.stft(center=True, reflect) -> [T, 1025] (frame-major)"""
    w = _window()
    yp = np.pad(y, N // 2, mode="reflect")
    T = 1 + (len(yp) - N) // HOP
    idx = np.arange(N)[None, :] + HOP * np.arange(T)[:, None]
    return tr.rfft(yp[idx] * w[None]).astype(np.complex64).astype(np.complex128)       # librosa stores complex64


def istft(D, tr):
    """librosa.istft -> hop*(T-1) samples (float32 overlap-add like librosa's buffer)"""
    w = _window()
    T = D.shape[0]
    fr = tr.irfft(D) * w[None]
    y = np.zeros(N + HOP * (T - 1), dtype=np.float32)
    ws = np.zeros_like(y)
    for t in range(T):
        y[t * HOP:t * HOP + N] += fr[t].astype(np.float32)
        ws[t * HOP:t * HOP + N] += (w * w).astype(np.float32)
    nz = ws > np.finfo(np.float32).tiny
    y[nz] /= ws[nz]
    return y[N // 2:-(N // 2)].astype(np.float64)


def griffin_lim(S, angles, iters, tr):
    """audio.py:77-87 on [T, F] arrays"""
    Sc = np.abs(S).astype(complex)
    y = istft(Sc * angles, tr)
    for _ in range(iters):
        ang = np.exp(1j * np.angle(stft(y, tr)))
        y = istft(Sc * ang, tr)
    return y


def snr_db(a, b):
    return 10 * np.log10(np.sum(b ** 2) / max(np.sum((a - b) ** 2), 1e-300))


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def speechlike(n, seed, sr=20000):
    rng = np.random.RandomState(seed)
    t = np.arange(n) / sr
    f0 = rng.uniform(100, 250)
    x = np.zeros(n)
    for k in range(1, 31):
        x += np.sin(2 * np.pi * f0 * k * t + rng.uniform(0, 2 * np.pi)) / k
    x *= 0.6 + 0.4 * np.sin(2 * np.pi * 3 * t)
    x += 10 ** (-50 / 20) * rng.randn(n)
    return x * 0.9 / np.max(np.abs(x))


def magnitudes(v, min_level_db, ref_level_db=20.0, power=1.5):
    """audio.py:47-48: _db_to_amp(_denormalize(v) + ref) ** power"""
    return (10.0 ** (((np.clip(v, 0, 1) * -min_level_db) + min_level_db + ref_level_db) * 0.05)) ** power


def run(iters=60, T=120, seeds=(0,)):
    trs = {"fp64": Fp64Transform(), "fp32_fft": Fp32Transform(), "split_fp16_data2": TwoPassTransform(2), "split_fp16_data1": TwoPassTransform(1)}
    out = {"iters": iters, "frames": T, "single_op": {}, "griffin_lim": {}}
    # single-op accuracy on a speech-like clip
    y = speechlike(HOP * (T - 1), 3)
    Dref = stft(y, trs["fp64"])
    for name in ("fp32_fft", "split_fp16_data2", "split_fp16_data1"):
        out["single_op"][name] = {"stft_rel_l2": rel_l2(stft(y, trs[name]), Dref), "istft_rel_l2": rel_l2(istft(Dref, trs[name]), istft(Dref, trs["fp64"]))}
    for seed in seeds:
        rs = np.random.RandomState(seed)
        fams = {
            "uniform_yaml_plus100": magnitudes(rs.rand(T, N // 2 + 1), 100.0),          # what bench.py runs: U(0,1), the yaml's +100
            "uniform_minus100": magnitudes(rs.rand(T, N // 2 + 1), -100.0),
            "consistent_speechlike": np.abs(stft(speechlike(HOP * (T - 1), 10 + seed), trs["fp64"])),
        }
        for fam, S in fams.items():
            ang = np.exp(2j * np.pi * rs.rand(*S.shape))
            ref = griffin_lim(S, ang, iters, trs["fp64"])
            for name in ("fp32_fft", "split_fp16_data2", "split_fp16_data1"):
                out["griffin_lim"].setdefault(fam, {}).setdefault(name, []).append(snr_db(griffin_lim(S, ang, iters, trs[name]), ref))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--frames", type=int, default=120)
    a = ap.parse_args()
    res = run(iters=12 if a.quick else a.iters, T=40 if a.quick else a.frames, seeds=(0,) if a.quick else (0, 1))
    json.dump(res, sys.stdout, indent=1)
    print()
