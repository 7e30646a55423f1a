// In-register FFT building blocks for the 2048-point real frame transform.
//
// Everything here is straight-line code on register arrays with compile-time indices and
// compile-time twiddles (they become FFMA/FMUL immediates), so one lane runs a whole 32-point
// complex FFT without touching memory.  The file is plain C++17 behind NSB_HD so that the same
// source is compiled by g++ into the warp-emulation harness (tests/emu/) - that is how the index
// algebra is validated in the GPU-less build container.  It is not a CPU fallback: nothing in the
// product links the harness.
//
// Replaces (together with frame_fft.cuh): scipy.fftpack.fft / ifft as called by librosa.stft /
// librosa.istft, reference call sites neural_speech/utils/audio.py:108 and :113.
#pragma once

#if defined(__CUDACC__)
#define NSB_HD __host__ __device__ __forceinline__
#define NSB_HDC __host__ __device__ constexpr
#else
#define NSB_HD inline __attribute__((always_inline))
#define NSB_HDC constexpr
#endif

namespace nsb {

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

// multiply (r,i) by exp(DIR * 2*pi*i * K / N), K and N compile-time; trivial cases cost no multiplies
template <int K, int N, int DIR>
NSB_HD void tw_mul(float& r, float& i) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) {
    } else if constexpr (2 * k == N) {
        r = -r; i = -i;
    } else if constexpr (4 * k == N) {          // exp(DIR*i*pi/2) = DIR*i
        float t = r;
        if constexpr (DIR > 0) { r = -i; i = t; } else { r = i; i = -t; }
    } else if constexpr (4 * k == 3 * N) {      // exp(DIR*i*3pi/2) = -DIR*i
        float t = r;
        if constexpr (DIR > 0) { r = i; i = -t; } else { r = -i; i = t; }
    } else {
        constexpr float c = float(cx::cos2pi(k, N));
        constexpr float s = float(double(DIR) * cx::sin2pi(k, N));
        float nr = r * c - i * s;
        float ni = r * s + i * c;
        r = nr; i = ni;
    }
}

// ---------------------------------------------------------------------------------------------
// out-of-place decimation-in-time FFT on register arrays: y[k] = sum_j x[IS*j] exp(DIR*2*pi*i*j*k/N)
// ---------------------------------------------------------------------------------------------
template <int N, int DIR, int IS>
struct FftRec {
    static NSB_HD void run(const float* xr, const float* xi, float* yr, float* yi) {
        static_assert(N % 4 == 0, "radix-4 step");
        constexpr int Q = N / 4;
        float fr[4][Q], fi[4][Q];
#pragma unroll
        for (int j = 0; j < 4; ++j) FftRec<Q, DIR, IS * 4>::run(xr + IS * j, xi + IS * j, fr[j], fi[j]);
        combine<0>(fr, fi, yr, yi);
    }
    template <int K>
    static NSB_HD void combine(float (&fr)[4][N / 4], float (&fi)[4][N / 4], float* yr, float* yi) {
        constexpr int Q = N / 4;
        if constexpr (K < Q) {
            float ar = fr[0][K], ai = fi[0][K];
            float br = fr[1][K], bi = fi[1][K];
            float cr = fr[2][K], ci = fi[2][K];
            float dr = fr[3][K], di = fi[3][K];
            tw_mul<K, N, DIR>(br, bi);
            tw_mul<2 * K, N, DIR>(cr, ci);
            tw_mul<3 * K, N, DIR>(dr, di);
            float t0r = ar + cr, t0i = ai + ci;
            float t1r = ar - cr, t1i = ai - ci;
            float t2r = br + dr, t2i = bi + di;
            float t3r = br - dr, t3i = bi - di;
            yr[K] = t0r + t2r;         yi[K] = t0i + t2i;
            yr[K + 2 * Q] = t0r - t2r; yi[K + 2 * Q] = t0i - t2i;
            if constexpr (DIR < 0) {   // forward: y[k+Q] = t1 - i*t3, y[k+3Q] = t1 + i*t3
                yr[K + Q] = t1r + t3i;     yi[K + Q] = t1i - t3r;
                yr[K + 3 * Q] = t1r - t3i; yi[K + 3 * Q] = t1i + t3r;
            } else {
                yr[K + Q] = t1r - t3i;     yi[K + Q] = t1i + t3r;
                yr[K + 3 * Q] = t1r + t3i; yi[K + 3 * Q] = t1i - t3r;
            }
            combine<K + 1>(fr, fi, yr, yi);
        }
    }
};
template <int DIR, int IS>
struct FftRec<1, DIR, IS> {
    static NSB_HD void run(const float* xr, const float* xi, float* yr, float* yi) { yr[0] = xr[0]; yi[0] = xi[0]; }
};
template <int DIR, int IS>
struct FftRec<2, DIR, IS> {
    static NSB_HD void run(const float* xr, const float* xi, float* yr, float* yi) {
        float ar = xr[0], ai = xi[0], br = xr[IS], bi = xi[IS];
        yr[0] = ar + br; yi[0] = ai + bi;
        yr[1] = ar - br; yi[1] = ai - bi;
    }
};

// in-place convenience wrapper: 32-point complex FFT of (re, im)
template <int DIR>
NSB_HD void fft32(float (&re)[32], float (&im)[32]) {
    float yr[32], yi[32];
    FftRec<32, DIR, 1>::run(re, im, yr, yi);
#pragma unroll
    for (int k = 0; k < 32; ++k) { re[k] = yr[k]; im[k] = yi[k]; }
}

// ---------------------------------------------------------------------------------------------
// real-64 <-> packed-complex-32 split passes (all in registers of one lane)
//
// g[t] = x[2t] + i x[2t+1] (t<32), G = FFT32(g).  The 64-point spectrum of x is
//   Y[q] = E[q] + w64^q O[q],  E[q] = (G[q] + conj G[32-q])/2,  O[q] = (G[q] - conj G[32-q])/(2i).
// real64_post turns G (in re/im) into 2*Y[q] for q = 1..31 in place; slot 0 keeps (Re G[0], Im G[0])
// = (sum of even samples, sum of odd samples), i.e. Y[0] = re0 + im0 and Y[32] = re0 - im0.
// real64_pre is the inverse map: from V[q] (q = 1..31, Hermitian 64-spectrum) and slot 0 holding
// (V[0] + V[32], V[0] - V[32]) it builds 2*G so that IFFT32 gives 2*64/... (see frame_fft.cuh for
// the overall scale bookkeeping).
// ---------------------------------------------------------------------------------------------
template <int Q>
NSB_HD void real64_post_pair(float (&re)[32], float (&im)[32]) {
    // A = G[Q], B = conj(G[32-Q]);  S = A + B, D = A - B;  2Y[Q] = S - i*w^Q*D ; 2Y[32-Q] = conj(S + i*w^Q*D)
    float ar = re[Q], ai = im[Q], br = re[32 - Q], bi = -im[32 - Q];
    float sr = ar + br, si = ai + bi;
    float dr = ar - br, di = ai - bi;
    // T = -i * w64^Q * D with w64^Q = exp(-2*pi*i*Q/64):  -i*w = exp(-i*(pi/2 + 2*pi*Q/64)) = exp(-2*pi*i*(Q+16)/64)
    tw_mul<Q + 16, 64, -1>(dr, di);
    re[Q] = sr + dr;      im[Q] = si + di;
    re[32 - Q] = sr - dr; im[32 - Q] = -(si - di);
}
NSB_HD void real64_post(float (&re)[32], float (&im)[32]) {
    real64_post_pair<1>(re, im);  real64_post_pair<2>(re, im);  real64_post_pair<3>(re, im);
    real64_post_pair<4>(re, im);  real64_post_pair<5>(re, im);  real64_post_pair<6>(re, im);
    real64_post_pair<7>(re, im);  real64_post_pair<8>(re, im);  real64_post_pair<9>(re, im);
    real64_post_pair<10>(re, im); real64_post_pair<11>(re, im); real64_post_pair<12>(re, im);
    real64_post_pair<13>(re, im); real64_post_pair<14>(re, im); real64_post_pair<15>(re, im);
    // Q = 16: 2Y[16] = 2*conj(G[16])
    re[16] = 2.0f * re[16]; im[16] = -2.0f * im[16];
}

template <int Q>
NSB_HD void real64_pre_pair(float (&re)[32], float (&im)[32]) {
    // inputs V[Q], V[32-Q];  S = V[Q] + conj V[32-Q] (=2E), D = (V[Q] - conj V[32-Q]) * conj(w64^Q) (=2O)
    // G'[Q] = S + i*D ;  G'[32-Q] = conj(S) + i*conj(D) ... derived from E,O being spectra of real sequences:
    // E[32-Q] = conj E[Q], O[32-Q] = conj O[Q]  =>  G'[32-Q] = conj(S) + i*conj(D)
    float ar = re[Q], ai = im[Q], br = re[32 - Q], bi = -im[32 - Q];
    float sr = ar + br, si = ai + bi;
    float dr = ar - br, di = ai - bi;
    // i * conj(w64^Q) = exp(+i*(pi/2 + 2*pi*Q/64)) = exp(+2*pi*i*(Q+16)/64)
    tw_mul<Q + 16, 64, +1>(dr, di);    // now (dr,di) = i*D
    re[Q] = sr + dr;      im[Q] = si + di;
    // i*conj(D) = conj(-i*D) = -conj(i*D)
    re[32 - Q] = sr - dr; im[32 - Q] = -si + di;
}
NSB_HD void real64_pre(float (&re)[32], float (&im)[32]) {
    real64_pre_pair<1>(re, im);  real64_pre_pair<2>(re, im);  real64_pre_pair<3>(re, im);
    real64_pre_pair<4>(re, im);  real64_pre_pair<5>(re, im);  real64_pre_pair<6>(re, im);
    real64_pre_pair<7>(re, im);  real64_pre_pair<8>(re, im);  real64_pre_pair<9>(re, im);
    real64_pre_pair<10>(re, im); real64_pre_pair<11>(re, im); real64_pre_pair<12>(re, im);
    real64_pre_pair<13>(re, im); real64_pre_pair<14>(re, im); real64_pre_pair<15>(re, im);
    // Q = 16: G'[16] = 2*conj(V[16])
    re[16] = 2.0f * re[16]; im[16] = -2.0f * im[16];
}

}  // namespace nsb
