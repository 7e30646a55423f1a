// Generic-size path: the same operators for ANY even n_fft = 2 * (num_freq - 1) (hparams-driven, audio.py:126-130).
//
// The fused kernels of kernels.cuh / gl_iter.cuh / gl_stream.cuh are specialised for the yaml's n_fft = 2048 (64 x 32 in
// registers).  Every other num_freq - 513 (n_fft 1024), 2049 (4096), the yaml's commented alternative 2048 (n_fft 4094 =
// 2 * 23 * 89), 401 (800 = 2^5 * 5^2) ... - runs here: one CTA per frame, the frame's real transform as a complex FFT of
// length M = n_fft / 2 in shared memory (Stockham autosort, MIXED RADIX: the host factors M into 8s, 4s, 2s and its odd
// primes; the power-of-two factors run as butterfly passes - one thread per butterfly, the R-point DFT in registers - an odd
// prime r as a direct r-point DFT per output, twiddles from one table exp(-2 pi i m / n_fft) that sits in shared memory when
// it fits) plus the real-split pass.  The spectrum of an iteration crosses shared memory only, but frames, magnitudes and the
// windowed frames of the overlap-add go through HBM, and the overlap-add is a gather kernel of
// its own (sum over the covering frames in ascending frame order: deterministic, no atomics).
//
// Replaces, for those hparams: librosa.stft / istft / filters.mel as called from audio.py:108, 113, 147 and the
// Griffin-Lim loop audio.py:77-87 (and its TensorFlow twin 90-103 through the Plan's geometry fields).
#pragma once
#include "kernels.cuh"

namespace nsb {

constexpr int kGenThreads = 256;
constexpr int kGenMaxStages = 24;

struct GenPlan {
    int n_fft, M, F;                 // M = n_fft / 2 (complex FFT length), F = M + 1 bins
    int n_stages;
    int radix[kGenMaxStages];        // product = M
    const float2* wt;                // [n_fft] exp(-2 pi i m / n_fft), rounded from double
    int wt_in_smem;                  // 1: the kernels keep a copy of the table in shared memory (it fits beside the frame)
    const float* win;                // [n_fft] window in the frame (librosa: padded centrally; tf: at the start)
    int hop, win_len, lo, origin, norm_wss;
    // sparse mel rows (the Plan's, built for F bins)
    const float* mel_w; const int* mel_lo; const int* mel_n; const int* mel_ptr; int num_mels;
};

// radix-R butterfly pass for R = 2, 4, 8: one thread per butterfly - R inputs, R - 1 twiddle multiplications, the R-point DFT in
// registers (FftC of fft_core.cuh, packed fp32) - instead of R multiply-adds per OUTPUT: 5 x fewer operations at radix 8
template <int DIR, int R>
__device__ __forceinline__ void gen_stage_pow2(const GenPlan& G, const float2* wt, const float2* a, float2* b, int Ns) {
    const int Mr = G.M / R;
    const int tstep = G.n_fft / (Ns * R);                        // exp(-2 pi i / (Ns R)) = wt[tstep]
    for (int j = threadIdx.x; j < Mr; j += blockDim.x) {
        const int k = j & (Ns - 1);                              // the power-of-two factors come first: Ns is a power of two here
        c2 v[R], y[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = a[j + q * Mr];
        if (k) {
#pragma unroll
            for (int q = 1; q < R; ++q) {
                const float2 w = wt[q * k * tstep];               // q k < Ns R: the index stays below n_fft
                v[q] = DIR > 0 ? cmul_conj(v[q], w) : cmul(v[q], w);
            }
        }
        FftC<R, DIR, 1>::run(v, y);
        float2* o = b + (j - k) * R + k;                         // (j / Ns) * Ns * R + k
#pragma unroll
        for (int t = 0; t < R; ++t) o[t * Ns] = y[t];
    }
}

// complex FFT of length M: `a` holds the input, `b` is scratch; returns the array that holds the result (natural order).
// DIR = -1: sum x e^{-i...} (forward), +1: conjugate twiddles (unnormalised inverse).  All threads of the CTA.
// Factors 8, 4, 2 run as butterfly passes; an odd prime factor r as a direct r-point DFT per output (r multiply-adds, the
// twiddle index advanced by a modular add - no per-radix code).
template <int DIR>
__device__ float2* gen_cfft(const GenPlan& G, const float2* wt, float2* a, float2* b) {
    const int M = G.M;
    int Ns = 1;
    for (int s = 0; s < G.n_stages; ++s) {
        const int r = G.radix[s], Mr = M / r;
        if (r == 8) gen_stage_pow2<DIR, 8>(G, wt, a, b, Ns);
        else if (r == 4) gen_stage_pow2<DIR, 4>(G, wt, a, b, Ns);
        else if (r == 2) gen_stage_pow2<DIR, 2>(G, wt, a, b, Ns);
        else {
            const int tstep = G.n_fft / (Ns * r);                    // exp(-2 pi i / (Ns r)) = wt[tstep]
            for (int o = threadIdx.x; o < M; o += blockDim.x) {
                const int t = o / Mr, j = o - t * Mr;
                const int k = j % Ns;
                const int step = (k + t * Ns) * tstep;               // < n_fft: the twiddle index advances by this per input
                int idx = 0;
                float ax = 0.f, ay = 0.f;
                for (int q = 0; q < r; ++q) {
                    const float2 v = a[j + q * Mr];
                    float2 w = wt[idx];
                    if (DIR > 0) w.y = -w.y;
                    ax = fmaf(v.x, w.x, fmaf(-v.y, w.y, ax));
                    ay = fmaf(v.x, w.y, fmaf(v.y, w.x, ay));
                    idx += step;
                    if (idx >= G.n_fft) idx -= G.n_fft;
                }
                b[(j / Ns) * Ns * r + k + t * Ns] = make_float2(ax, ay);
            }
        }
        __syncthreads();
        float2* tmp = a; a = b; b = tmp;
        Ns *= r;
    }
    return a;
}

// the twiddle table the CTA reads: a shared-memory copy behind the frame's arrays when it fits (4 to 8 table reads per butterfly
// otherwise go to L1 / L2), else the global one
__device__ __forceinline__ const float2* gen_twiddles(const GenPlan& G, float2* smem_copy) {
    if (!G.wt_in_smem) return G.wt;
    for (int i = threadIdx.x; i < G.n_fft; i += blockDim.x) smem_copy[i] = __ldg(G.wt + i);
    __syncthreads();
    return smem_copy;
}

// rfft of the packed frame z[m] = (x[2m], x[2m+1]) after gen_cfft: X[k], k = 0..M
__device__ __forceinline__ float2 gen_split_fwd(const GenPlan& G, const float2* wt, const float2* Z, int k) {
    const int M = G.M;
    const float2 zk = Z[k == M ? 0 : k];
    const float2 zc = Z[k == 0 ? 0 : M - k];                    // Z[M - k] (conjugated below)
    const float ex = 0.5f * (zk.x + zc.x), ey = 0.5f * (zk.y - zc.y);         // E = (Z[k] + conj Z[M-k]) / 2
    const float dx = zk.x - zc.x, dy = zk.y + zc.y;                            // D = Z[k] - conj Z[M-k]
    const float ox = 0.5f * dy, oy = -0.5f * dx;                               // O = -i D / 2
    const float2 w = wt[k];                                                    // exp(-2 pi i k / n_fft), k <= M < n_fft
    return make_float2(ex + (ox * w.x - oy * w.y), ey + (ox * w.y + oy * w.x));
}
// the packed spectrum whose inverse complex FFT is (x[2m], x[2m+1]) * M, from the Hermitian half X[0..M] (function of k)
template <typename GetX>
__device__ __forceinline__ float2 gen_split_inv(const GenPlan& G, const float2* wt, GetX X, int k) {
    const int M = G.M;
    float2 xk = X(k), xc = X(M - k);
    if (k == 0) { xk.y = 0.f; xc.y = 0.f; }                    // DC and Nyquist are real (irfft ignores their imaginary parts)
    const float ex = 0.5f * (xk.x + xc.x), ey = 0.5f * (xk.y - xc.y);         // E = (X[k] + conj X[M-k]) / 2
    const float dx = 0.5f * (xk.x - xc.x), dy = 0.5f * (xk.y + xc.y);         // (X[k] - conj X[M-k]) / 2
    const float2 w = wt[k];
    const float ox = dx * w.x + dy * w.y, oy = dy * w.x - dx * w.y;            // O = conj(w) * that
    return make_float2(ex - oy, ey + ox);                                      // Z = E + i O
}

struct GenFrameLoc { int b, k, f_off, T; long long s_off, L; };
__device__ __forceinline__ GenFrameLoc gen_locate(const Batch& B, int f) {
    GenFrameLoc o;
    o.b = find_segment(B.frame_off, B.batch, f);
    o.f_off = __ldg(B.frame_off + o.b);
    o.k = f - o.f_off;
    o.T = __ldg(B.frame_off + o.b + 1) - o.f_off;
    o.s_off = __ldg(B.samp_off + o.b);
    o.L = __ldg(B.samp_off + o.b + 1) - o.s_off;
    return o;
}

// windowed (pre-emphasised) frame k of the utterance, packed into z[m] = (x[2m], x[2m+1]); samples outside the window's
// support are zero; librosa geometry reflects at the utterance's ends (np.pad mode='reflect')
template <bool PREEMPH>
__device__ __forceinline__ void gen_load_frame(const GenPlan& G, float2* z, const float* x, long long L, long long start, float p) {
    const int lo = G.lo, hi = G.lo + G.win_len;
    for (int m = threadIdx.x; m < G.M; m += blockDim.x) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int n = 2 * m + e;
            v[e] = (n >= lo && n < hi) ? sample_at<PREEMPH, true>(x, (int)L, (int)(start + n), p) * __ldg(G.win + n) : 0.f;
        }
        z[m] = make_float2(v[0], v[1]);
    }
    __syncthreads();
}

struct GenAnalysisParams {
    GenPlan plan;
    Batch batch;
    const float* wav;
    float2* out_complex;      // [frames][F] or null
    float* out_lin;           // [rows][F] or null
    float* out_mel;           // [rows][num_mels] or null
    int total_frames;
    int rows_per_utt;
    int preemph_on;
    float preemph;
    float db_scale, db_offset_lin, db_offset_mel;
    int* status;
};

// _stft / spectrogram + melspectrogram for any n_fft: one CTA per frame (grid-stride)
__global__ void __launch_bounds__(kGenThreads) k_gen_analysis(GenAnalysisParams P) {
    NSB_DYN_SMEM(smem_raw);
    const GenPlan& G = P.plan;
    float2* a = reinterpret_cast<float2*>(smem_raw);
    float2* b = a + G.M;
    const float2* wt = gen_twiddles(G, b + G.M);
    bool bad = false;
    for (int f = P.batch.frame_base + (int)blockIdx.x; f < P.batch.frame_base + P.total_frames; f += (int)gridDim.x) {
        const GenFrameLoc loc = gen_locate(P.batch, f);
        const long long start = (long long)loc.k * G.hop - G.origin;
        if (P.preemph_on) gen_load_frame<true>(G, a, P.wav + loc.s_off, loc.L, start, P.preemph);
        else gen_load_frame<false>(G, a, P.wav + loc.s_off, loc.L, start, 0.f);
        float2* Z = gen_cfft<-1>(G, wt, a, b);
        float2* other = (Z == a) ? b : a;
        if (P.out_complex) {
            float2* o = P.out_complex + (size_t)f * G.F;
            for (int k = threadIdx.x; k <= G.M; k += blockDim.x) {
                const float2 X = gen_split_fwd(G, wt, Z, k);
                bad |= !(isfinite(X.x) && isfinite(X.y));
                o[k] = X;
            }
        } else {
            const long long orow = P.batch.row_off ? __ldg(P.batch.row_off + loc.b) + loc.k
                                 : P.rows_per_utt > 0 ? (long long)(P.batch.utt_base + loc.b) * P.rows_per_utt + loc.k : (long long)f;
            float* magrow = reinterpret_cast<float*>(other);                  // F <= 2 M floats
            float* o_lin = P.out_lin ? P.out_lin + (size_t)orow * G.F : nullptr;
            for (int k = threadIdx.x; k <= G.M; k += blockDim.x) {
                const float2 X = gen_split_fwd(G, wt, Z, k);
                const float mg = sqrtf(fmaf(X.x, X.x, X.y * X.y));
                bad |= !isfinite(mg);
                magrow[k] = mg;
                if (o_lin) o_lin[k] = amp_to_db_norm_fast(mg, P.db_scale, P.db_offset_lin);
            }
            __syncthreads();
            if (P.out_mel) {
                float* o = P.out_mel + (size_t)orow * G.num_mels;
                for (int m = threadIdx.x; m < G.num_mels; m += blockDim.x) {
                    const int lo = __ldg(G.mel_lo + m), n = __ldg(G.mel_n + m);
                    const float* w = G.mel_w + __ldg(G.mel_ptr + m);
                    float acc = 0.f;
                    for (int i = 0; i < n; ++i) acc = fmaf(__ldg(w + i), magrow[lo + i], acc);
                    o[m] = amp_to_db_norm_fast(acc, P.db_scale, P.db_offset_mel);
                }
            }
        }
        __syncthreads();
    }
    if (bad) atomicOr(P.status, 1);
}

struct GenSynthParams {
    GenPlan plan;
    Batch batch;
    int src;                  // SRC_* of kernels.cuh
    const float* y_in;        // SRC_Y
    const float* mag;         // [frames][F] natural order (all sources but SRC_SPEC)
    const float2* spec;       // SRC_SPEC / SRC_MAGPHASE: per utterance block at F * frame_off[b], [T][F] or [F][T]
    int spec_bin_major;
    float* frames_out;        // [frames][win_len]: the windowed inverse transform of every frame (before the overlap-add)
    int total_frames;
    int tf_renorm;            // 1: est / max(1e-8, |est|) (the TF twin), 0: exp(i angle(est))
    unsigned long long seed;
    int* status;
};

// one Griffin-Lim half-step (or a plain inverse transform) per frame, spectrum in shared memory only:
// [frame of y -> rfft -> phase renormalise x magnitude | given spectrum | magnitude x phase] -> irfft -> window -> frames_out
__global__ void __launch_bounds__(kGenThreads) k_gen_synth(GenSynthParams P) {
    NSB_DYN_SMEM(smem_raw);
    const GenPlan& G = P.plan;
    const int M = G.M;
    float2* a = reinterpret_cast<float2*>(smem_raw);
    float2* b = a + M;
    float2* X = b + M;                                     // [F] the frame's spectrum
    const float2* wt = gen_twiddles(G, X + G.F);
    bool bad = false;
    for (int f = P.batch.frame_base + (int)blockIdx.x; f < P.batch.frame_base + P.total_frames; f += (int)gridDim.x) {
        const GenFrameLoc loc = gen_locate(P.batch, f);
        const float* mrow = P.mag ? P.mag + (size_t)f * G.F : nullptr;
        if (P.src == SRC_Y) {
            gen_load_frame<false>(G, a, P.y_in + loc.s_off, loc.L, (long long)loc.k * G.hop - G.origin, 0.f);
            const float2* Z = gen_cfft<-1>(G, wt, a, b);
            for (int k = threadIdx.x; k <= M; k += blockDim.x) {
                c2 z = gen_split_fwd(G, wt, Z, k);
                const float S = __ldg(mrow + k);
                if (k == 0 || k == M) z.y = 0.f;
                if (P.tf_renorm) {
                    const float m = sqrtf(fmaf(z.x, z.x, z.y * z.y));
                    const float sc = S / fmaxf(1e-8f, m);
                    z = mk2(z.x * sc, z.y * sc);
                } else {
                    renorm(z, S);
                }
                X[k] = z;
            }
        } else if (P.src == SRC_SPEC || P.src == SRC_MAGPHASE) {
            const float2* sp = P.spec + (size_t)loc.f_off * G.F;
            for (int k = threadIdx.x; k <= M; k += blockDim.x) {
                float2 v = P.spec_bin_major ? __ldg(sp + (size_t)k * loc.T + loc.k) : __ldg(sp + (size_t)loc.k * G.F + k);
                if (P.src == SRC_MAGPHASE) { const float S = __ldg(mrow + k); v = make_float2(v.x * S, v.y * S); }
                X[k] = v;
            }
        } else if (P.src == SRC_MAGZERO) {
            for (int k = threadIdx.x; k <= M; k += blockDim.x) X[k] = make_float2(__ldg(mrow + k), 0.f);
        } else {   // SRC_MAGRAND: Philox keyed by the seed, counter = (frame, bin / 4); one draw serves four bins
            const uint2 key = make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32));
            for (int k4 = threadIdx.x; 4 * k4 <= M; k4 += blockDim.x) {
                const uint4 r = philox4x32(make_uint4((uint32_t)f, (uint32_t)k4, 0x6e737067u, 0u), key);
                const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
                for (int e = 0; e < 4 && 4 * k4 + e <= M; ++e) {
                    float sn, cs;
                    sincospif(2.0f * (float)(rr[e] >> 8) * (1.0f / 16777216.0f), &sn, &cs);
                    const float S = __ldg(mrow + 4 * k4 + e);
                    X[4 * k4 + e] = make_float2(S * cs, S * sn);
                }
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < M; k += blockDim.x)
            a[k] = gen_split_inv(G, wt, [&](int i) { return X[i]; }, k);
        __syncthreads();
        const float2* z = gen_cfft<+1>(G, wt, a, b);
        // x[2m] = Re z[m] / M, x[2m+1] = Im z[m] / M; keep the window's support, windowed
        float* o = P.frames_out + (size_t)f * G.win_len;
        const float inv = 1.0f / (float)M;
        for (int i = threadIdx.x; i < G.win_len; i += blockDim.x) {
            const int n = G.lo + i;
            const float2 v = z[n >> 1];
            const float s = ((n & 1) ? v.y : v.x) * inv * __ldg(G.win + n);
            bad |= !isfinite(s);
            o[i] = s;
        }
        __syncthreads();
    }
    if (bad) atomicOr(P.status, 1);
}

struct GenOlaParams {
    GenPlan plan;
    Batch batch;
    const float* frames;      // [frames][win_len]
    float* y_out;             // packed samples
    long long total_samples;
};

// overlap-add as a gather: sample i of an utterance sums the frames that cover it in ascending frame order (float32 like
// librosa's buffer), then the division by the summed squared window of the frames that exist (librosa.istft; not for tf)
__global__ void __launch_bounds__(256) k_gen_ola(GenOlaParams P) {
    const GenPlan& G = P.plan;
    const int a = G.origin - G.lo;                     // frame k's support starts at sample k*hop - a
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < P.total_samples; g += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = P.batch.batch;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (__ldg(P.batch.samp_off + mid) <= g) lo = mid; else hi = mid; }
        const long long s_off = __ldg(P.batch.samp_off + lo);
        const int f_off = __ldg(P.batch.frame_off + lo);
        const int T = __ldg(P.batch.frame_off + lo + 1) - f_off;
        const long long i = g - s_off;
        // k*hop - a <= i < k*hop - a + win
        const long long num = i + a - G.win_len;                   // k*hop > num
        long long k0 = num >= 0 ? num / G.hop + 1 : -((-num - 1) / G.hop + 1) + 1;
        long long k1 = (i + a) / G.hop;                             // (i + a >= 0 always: a >= 0)
        if (k0 < 0) k0 = 0;
        if (k1 > T - 1) k1 = T - 1;
        float acc = 0.f, ws = 0.f;
        for (long long k = k0; k <= k1; ++k) {
            const int idx = (int)(i - (k * G.hop - a));
            acc += __ldg(P.frames + (size_t)(f_off + k) * G.win_len + idx);
            const float w = __ldg(G.win + G.lo + idx);
            ws = fmaf(w, w, ws);
        }
        if (G.norm_wss && ws > 1.17549435e-38f) acc /= ws;
        P.y_out[g] = acc;
    }
}

// magnitudes in natural order for the generic path (prep_magnitude of kernels.cuh; scale = 1)
__global__ void __launch_bounds__(256) k_gen_prepare_mag(PrepParams P, int F) {
    bool bad = false;
    const long long total = (long long)P.total_frames * F;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int f = P.batch.frame_base + (int)(g / F), k = (int)(g - (long long)(g / F) * F);
        float v;
        if (P.bin_major) {
            const int bb = find_segment(P.batch.frame_off, P.batch.batch, f);
            const int fo = __ldg(P.batch.frame_off + bb), T = __ldg(P.batch.frame_off + bb + 1) - fo;
            v = __ldg(P.in + (size_t)fo * F + (size_t)k * T + (f - fo));
        } else {
            v = __ldg(P.in + (size_t)f * F + k);
        }
        bad |= !isfinite(v);
        P.mag[(size_t)f * F + k] = prep_magnitude(P, v);
    }
    if (bad) atomicOr(P.status, 1);
}

// _linear_to_mel for any F (np.dot with the float64 basis: double accumulation)
__global__ void __launch_bounds__(256) k_gen_linear_to_mel(MelParams P, int F) {
    const long long total = (long long)P.total_frames * P.plan.num_mels;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int f = P.batch.frame_base + (int)(g / P.plan.num_mels), m = (int)(g % P.plan.num_mels);
        const int b = find_segment(P.batch.frame_off, P.batch.batch, f);
        const int fo = __ldg(P.batch.frame_off + b), T = __ldg(P.batch.frame_off + b + 1) - fo;
        const float* base = P.in + (size_t)fo * F;
        const int lo = __ldg(P.plan.mel_lo + m), n = __ldg(P.plan.mel_n + m);
        const float* w = P.plan.mel_w + __ldg(P.plan.mel_ptr + m);
        double acc = 0.0;
        for (int i = 0; i < n; ++i) {
            const float x = P.bin_major ? __ldg(base + (size_t)(lo + i) * T + (f - fo)) : __ldg(base + (size_t)(f - fo) * F + lo + i);
            acc = fma((double)__ldg(w + i), (double)x, acc);
        }
        if (P.out64) P.out64[(size_t)f * P.plan.num_mels + m] = acc; else P.out32[(size_t)f * P.plan.num_mels + m] = (float)acc;
    }
}

}  // namespace nsb
