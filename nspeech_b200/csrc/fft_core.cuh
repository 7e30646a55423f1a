// In-register FFT building blocks for the 2048-point real frame transform, in COMPLEX-PACKED form.
//
// Every complex value lives in one float2 = one aligned 64-bit register pair (x = re, y = im) and all
// arithmetic goes through Blackwell's packed fp32 instructions (FADD2 / FMUL2 / FFMA2, CUDA intrinsics
// __fadd2_rn / __fmul2_rn / __ffma2_rn, sm_100+).  Their operands can be swizzled for free in SASS
// (R.F32x2.LO_HI = halves swapped, per-half negation, R.F32 = scalar broadcast, 32-bit immediate broadcast),
// so a complex add is ONE instruction, a multiplication by +-i costs nothing (it folds into the next add) and a
// complex multiplication is two (FMUL2 + FFMA2) instead of four.  A 32-point complex FFT is 216 instructions
// instead of 420 scalar ones; the packed ops run at the same flop rate in half the issue slots
// (profiles/r1/microbench_fp32_pipes.txt), and issue slots are what bounds the Griffin-Lim kernel.
//
// Everything is straight-line code on register arrays with compile-time indices and twiddles.  The file is
// plain C++17 behind NSB_HD: the same source is compiled by g++ into the warp-emulation harness (tests/emu/),
// where the pair operations fall back to two scalar ones.  That is test infrastructure, not a CPU fallback.
//
// Replaces (together with frame_fft.cuh): scipy.fftpack.fft / ifft as called by librosa.stft / librosa.istft,
// reference call sites neural_speech/utils/audio.py:108 and :113.
#pragma once

#if defined(__CUDACC__)
#define NSB_HD __host__ __device__ __forceinline__
#define NSB_HDC __host__ __device__ constexpr
#else
#define NSB_HD inline __attribute__((always_inline))
#define NSB_HDC constexpr
#endif

namespace nsb {

#if !defined(__CUDACC__) && !defined(NSB_EMULATE)
struct alignas(8) float2 { float x, y; };
#endif
typedef float2 c2;   // complex: x = re, y = im

NSB_HD c2 mk2(float x, float y) { c2 r; r.x = x; r.y = y; return r; }

// ---- element-wise pair operations (one SASS instruction each on sm_100) -------------------------------------
NSB_HD c2 p_add(c2 a, c2 b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    return mk2(a.x + b.x, a.y + b.y);
#endif
}
NSB_HD c2 p_mul(c2 a, c2 b) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, b);
#else
    return mk2(a.x * b.x, a.y * b.y);
#endif
}
NSB_HD c2 p_fma(c2 a, c2 b, c2 c) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return mk2(__builtin_fmaf(a.x, b.x, c.x), __builtin_fmaf(a.y, b.y, c.y));
#endif
}

// ---- complex helpers; swaps / negations / broadcasts become operand modifiers ---------------------------------
NSB_HD c2 cadd(c2 a, c2 b) { return p_add(a, b); }
NSB_HD c2 csub(c2 a, c2 b) { return p_add(a, mk2(-b.x, -b.y)); }
NSB_HD c2 cadd_i(c2 a, c2 b) { return p_add(a, mk2(-b.y, b.x)); }            // a + i*b
NSB_HD c2 csub_i(c2 a, c2 b) { return p_add(a, mk2(b.y, -b.x)); }            // a - i*b
NSB_HD c2 cconj(c2 a) { return mk2(a.x, -a.y); }
NSB_HD c2 cadd_conj(c2 a, c2 b) { return p_add(a, mk2(b.x, -b.y)); }         // a + conj(b)
NSB_HD c2 csub_conj(c2 a, c2 b) { return p_add(a, mk2(-b.x, b.y)); }         // a - conj(b)
NSB_HD c2 cscale(c2 a, float s) { return p_mul(a, mk2(s, s)); }
NSB_HD c2 cmul(c2 z, c2 w) { return p_fma(mk2(-z.y, z.x), mk2(w.y, w.y), p_mul(z, mk2(w.x, w.x))); }        // z * w
NSB_HD c2 cmul_conj(c2 z, c2 w) { return p_fma(mk2(z.y, -z.x), mk2(w.y, w.y), p_mul(z, mk2(w.x, w.x))); }   // z * conj(w)

// ---------------------------------------------------------------------------------------------
// compile-time cos/sin of 2*pi*k/n (double precision Taylor after octant reduction)
// ---------------------------------------------------------------------------------------------
namespace cx {
constexpr double kPi = 3.141592653589793238462643383279502884;

NSB_HDC double tay_sin(double x) {
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 14; ++i) { term *= -x2 / double((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
NSB_HDC double tay_cos(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 14; ++i) { term *= -x2 / double((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}
// cos(2*pi*k/n), sin(2*pi*k/n) for 0 <= k, n > 0
NSB_HDC double cos2pi(long long k, long long n) {
    k %= n;
    if (2 * k > n) k = n - k;                        // cos even about pi
    if (4 * k > n) return -cos2pi(n - 2 * k, 2 * n); // cos(pi - a) = -cos(a)
    if (8 * k > n) return tay_sin(2.0 * kPi * double(n - 4 * k) / double(4 * n)); // cos(pi/2 - a)
    return tay_cos(2.0 * kPi * double(k) / double(n));
}
NSB_HDC double sin2pi(long long k, long long n) {
    k %= n;
    if (2 * k > n) return -sin2pi(n - k, n);
    if (4 * k > n) return sin2pi(n - 2 * k, 2 * n);
    if (8 * k > n) return tay_cos(2.0 * kPi * double(n - 4 * k) / double(4 * n));
    return tay_sin(2.0 * kPi * double(k) / double(n));
}
}  // namespace cx

// z * exp(DIR * 2*pi*i * K / N), K and N compile-time; quarter turns cost nothing
template <int K, int N, int DIR>
NSB_HD c2 ctw(c2 z) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) {
        return z;
    } else if constexpr (2 * k == N) {
        return mk2(-z.x, -z.y);
    } else if constexpr (4 * k == N) {          // exp(DIR*i*pi/2) = DIR*i
        if constexpr (DIR > 0) return mk2(-z.y, z.x); else return mk2(z.y, -z.x);
    } else if constexpr (4 * k == 3 * N) {      // exp(DIR*i*3pi/2) = -DIR*i
        if constexpr (DIR > 0) return mk2(z.y, -z.x); else return mk2(-z.y, z.x);
    } else {
        constexpr float c = float(cx::cos2pi(k, N));
        constexpr float s = float(double(DIR) * cx::sin2pi(k, N));
        return p_fma(mk2(-z.y, z.x), mk2(s, s), p_mul(z, mk2(c, c)));
    }
}

// ---------------------------------------------------------------------------------------------
// out-of-place radix-4 decimation-in-time FFT on register arrays: y[k] = sum_j x[IS*j] exp(DIR*2*pi*i*j*k/N)
// ---------------------------------------------------------------------------------------------
// NZ0, NZ1: only the inputs x[t] with NZ0 <= t < NZ1 (t = absolute index, OFF = index of this sub-transform's x[0]) are
// non-zero - the frame's window support covers half of the 2048 samples, so the first radix-2 layer of the forward
// transform adds zeros: its butterflies degenerate to copies (32 packed adds and 32 zero-initialising moves per frame)
template <int N, int DIR, int IS, int OFF = 0, int NZ0 = 0, int NZ1 = 1 << 30>
struct FftC {
    static NSB_HD void run(const c2* x, c2* y) {
        static_assert(N % 4 == 0, "radix-4 step");
        constexpr int Q = N / 4;
        c2 f[4][Q];
        FftC<Q, DIR, IS * 4, OFF, NZ0, NZ1>::run(x, f[0]);
        FftC<Q, DIR, IS * 4, OFF + IS, NZ0, NZ1>::run(x + IS, f[1]);
        FftC<Q, DIR, IS * 4, OFF + 2 * IS, NZ0, NZ1>::run(x + 2 * IS, f[2]);
        FftC<Q, DIR, IS * 4, OFF + 3 * IS, NZ0, NZ1>::run(x + 3 * IS, f[3]);
        combine<0>(f, y);
    }
    template <int K>
    static NSB_HD void combine(c2 (&f)[4][N / 4], c2* y) {
        constexpr int Q = N / 4;
        if constexpr (K < Q) {
            c2 a = f[0][K], b = ctw<K, N, DIR>(f[1][K]), c = ctw<2 * K, N, DIR>(f[2][K]), d = ctw<3 * K, N, DIR>(f[3][K]);
            c2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
            y[K] = cadd(t0, t2);
            y[K + 2 * Q] = csub(t0, t2);
            if constexpr (DIR < 0) { y[K + Q] = csub_i(t1, t3); y[K + 3 * Q] = cadd_i(t1, t3); }   // forward: t1 -+ i*t3
            else { y[K + Q] = cadd_i(t1, t3); y[K + 3 * Q] = csub_i(t1, t3); }
            combine<K + 1>(f, y);
        }
    }
};
template <int DIR, int IS, int OFF, int NZ0, int NZ1>
struct FftC<1, DIR, IS, OFF, NZ0, NZ1> {
    static NSB_HD void run(const c2* x, c2* y) { y[0] = x[0]; }
};
template <int DIR, int IS, int OFF, int NZ0, int NZ1>
struct FftC<2, DIR, IS, OFF, NZ0, NZ1> {
    static NSB_HD void run(const c2* x, c2* y) {
        constexpr bool a_nz = (OFF >= NZ0 && OFF < NZ1), b_nz = (OFF + IS >= NZ0 && OFF + IS < NZ1);
        if constexpr (a_nz && b_nz) { c2 a = x[0], b = x[IS]; y[0] = cadd(a, b); y[1] = csub(a, b); }
        else if constexpr (a_nz) { y[0] = x[0]; y[1] = x[0]; }
        else if constexpr (b_nz) { y[0] = x[IS]; y[1] = mk2(-x[IS].x, -x[IS].y); }
        else { y[0] = mk2(0.f, 0.f); y[1] = mk2(0.f, 0.f); }
    }
};

// in-place convenience wrapper: 32-point complex FFT (DIR = -1 forward, +1 inverse, unnormalised)
template <int DIR>
NSB_HD void fft32(c2 (&z)[32]) {
    c2 y[32];
    FftC<32, DIR, 1>::run(z, y);
#pragma unroll
    for (int k = 0; k < 32; ++k) z[k] = y[k];
}
// the same for an input whose entries outside [NZ0, NZ1) are zero (they are never read)
template <int DIR, int NZ0, int NZ1>
NSB_HD void fft32_sparse(c2 (&z)[32]) {
    c2 y[32];
    FftC<32, DIR, 1, 0, NZ0, NZ1>::run(z, y);
#pragma unroll
    for (int k = 0; k < 32; ++k) z[k] = y[k];
}

// ---------------------------------------------------------------------------------------------
// real-64 <-> packed-complex-32 split passes (all in registers of one lane)
//
// g[t] = x[2t] + i x[2t+1] (t<32), G = FFT32(g).  The 64-point spectrum of x is
//   Y[q] = E[q] + w64^q O[q],  E[q] = (G[q] + conj G[32-q])/2,  O[q] = (G[q] - conj G[32-q])/(2i).
// real64_post turns G into 2*Y[q] for q = 1..31 in place; slot 0 keeps (Re G[0], Im G[0]) = (sum of even
// samples, sum of odd samples), i.e. Y[0] = x0 + y0 and Y[32] = x0 - y0.
// real64_pre is the inverse map: from V[q] (q = 1..31, Hermitian 64-spectrum) and slot 0 holding
// (V[0] + V[32], V[0] - V[32]) it builds the G' whose inverse FFT32 is the packed real sequence (times 64).
// ---------------------------------------------------------------------------------------------
template <int Q, int DIR>
NSB_HD void real64_pair(c2 (&z)[32]) {
    // S = A + conj(B), D = A - conj(B), T = D * exp(DIR * 2*pi*i*(Q+16)/64)   (= -+i * w64^(-+Q) * D)
    // out[Q] = S + T, out[32-Q] = conj(S - T)
    c2 A = z[Q], B = z[32 - Q];
    c2 S = cadd_conj(A, B), D = csub_conj(A, B);
    c2 T = ctw<Q + 16, 64, DIR>(D);
    z[Q] = cadd(S, T);
    z[32 - Q] = cconj(csub(S, T));
}
template <int DIR>
NSB_HD void real64_split(c2 (&z)[32]) {
    real64_pair<1, DIR>(z);  real64_pair<2, DIR>(z);  real64_pair<3, DIR>(z);  real64_pair<4, DIR>(z);
    real64_pair<5, DIR>(z);  real64_pair<6, DIR>(z);  real64_pair<7, DIR>(z);  real64_pair<8, DIR>(z);
    real64_pair<9, DIR>(z);  real64_pair<10, DIR>(z); real64_pair<11, DIR>(z); real64_pair<12, DIR>(z);
    real64_pair<13, DIR>(z); real64_pair<14, DIR>(z); real64_pair<15, DIR>(z);
    z[16] = mk2(2.0f * z[16].x, -2.0f * z[16].y);     // Q = 16: 2 * conj
}
NSB_HD void real64_post(c2 (&z)[32]) { real64_split<-1>(z); }
NSB_HD void real64_pre(c2 (&z)[32]) { real64_split<+1>(z); }

}  // namespace nsb
