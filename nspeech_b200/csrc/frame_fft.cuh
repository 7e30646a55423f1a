// One warp = one 2048-point real frame transform, as 64 x 32:
//
//   n = 32*m + r   (r = lane, m = 0..63)          k = q + 64*p   (q = lane, p = 0..31)
//
//   pass 1 (lane r): real 64-point FFT over m, done as a packed complex 32-point FFT in registers
//                    + the real64 split (fft_core.cuh)                        -> Y_r[q], q = 0..32
//   twiddle        : Y_r[q] *= w2048^(r*q)   (table in shared memory, conflict-free columns)
//   transpose      : one trip through a per-warp 32x33 float2 scratch tile
//   pass 2 (lane q): complex 32-point FFT over r                               -> X[q + 64*p]
//
// Hermitian symmetry keeps every pair (k, 2048-k) inside ONE lane: lane q >= 1 ends up with the 32
// bins  k = q + 64p (p < 16, value Z[p])  and  k = (64-q) + 64(31-p) (p >= 16, value conj Z[p]).
// Rows q = 0 and q = 32 only carry real data; they are packed into row 0 (64 floats) and lane 0
// turns them into the 33 bins k = 32*j with one more real64 split.  So the per-bin Griffin-Lim
// update (phase renormalisation times magnitude) needs no cross-lane exchange at all, and the
// inverse transform is the same pipeline run backwards.
//
// Scale conventions (validated by tests/emu): fwd gives 2*rfft(x); inv gives 2048*irfft(X).
// Callers fold the 1/2 and 1/2048 into their window tables.
//
// Lane-0 packing of the "register spectrum": slot 0 = (X[0], X[1024]) (both real), slot j = X[32 j].
#pragma once
#include "fft_core.cuh"

namespace nsb {

typedef float2 f2;

constexpr int kNfft = 2048;
constexpr int kBins = 1025;
constexpr int kRowStride = 33;                      // float2 units; keeps row and column access conflict-free
constexpr int kScratchF2 = 32 * kRowStride;         // per-warp scratch tile (8448 B)
constexpr int kTwF2 = 31 * 32;                      // twiddle table, tw[(j-1)*32 + l] = exp(-2*pi*i*j*l/2048)

// frequency bin held in slot p of lane `lane` (lane 0 slot 0 is the packed DC/Nyquist pair -> returns 0)
NSB_HD int bin_of(int lane, int p) {
    if (lane == 0) return 32 * p;
    return p < 16 ? lane + 64 * p : (64 - lane) + 64 * (31 - p);
}
// true if slot p of lane holds the conjugate of the bin's value
NSB_HD bool slot_is_conj(int lane, int p) { return lane != 0 && p >= 16; }

// ---- forward -------------------------------------------------------------------------------
// in : z[t] = (x[64 t + lane], x[64 t + 32 + lane])
NSB_HD void fwd_phase1(c2 (&z)[32], int lane, f2* scratch, const f2* tw) {
    fft32<-1>(z);
    real64_post(z);
    float* row0 = reinterpret_cast<float*>(scratch);
    row0[lane] = z[0].x;
    row0[lane + 32] = z[0].y;
#pragma unroll
    for (int q = 1; q < 32; ++q) scratch[q * kRowStride + lane] = cmul(z[q], tw[(q - 1) * 32 + lane]);
}
// out: register spectrum (2 * rfft), see header comment
NSB_HD void fwd_phase2(c2 (&z)[32], int lane, const f2* scratch) {
#pragma unroll
    for (int t = 0; t < 32; ++t) z[t] = scratch[lane * kRowStride + t];
    fft32<-1>(z);
    if (lane == 0) {
        real64_post(z);
        float a = z[0].x, b = z[0].y;
        z[0] = mk2(2.0f * (a + b), 2.0f * (a - b));
    }
}

// ---- inverse -------------------------------------------------------------------------------
// in : register spectrum X (Im of DC / Nyquist do not exist in the packing, as in irfft)
NSB_HD void inv_phase1(c2 (&z)[32], int lane, f2* scratch, const f2* tw) {
    if (lane == 0) {
        float a = z[0].x, b = z[0].y;
        z[0] = mk2(a + b, a - b);
        real64_pre(z);
    }
    fft32<+1>(z);
    scratch[lane * kRowStride] = z[0];
#pragma unroll
    for (int r = 1; r < 32; ++r)      // table is symmetric in (j, l): conj(w2048^(lane*r))
        scratch[lane * kRowStride + r] = cmul_conj(z[r], tw[(r - 1) * 32 + lane]);
}
// out: z[t] = 2048 * (x[64 t + lane], x[64 t + 32 + lane])
NSB_HD void inv_phase2(c2 (&z)[32], int lane, const f2* scratch) {
    const float* row0 = reinterpret_cast<const float*>(scratch);
    z[0] = mk2(row0[lane], row0[lane + 32]);
#pragma unroll
    for (int q = 1; q < 32; ++q) z[q] = scratch[q * kRowStride + lane];
    real64_pre(z);
    fft32<+1>(z);
}

#if defined(__CUDACC__) || defined(NSB_EMULATE)
// Twiddles w2048^(q*l) in shared memory in PAIRS of rows: tw4[p*32 + l] = (w^((2p+1) l), w^((2p+2) l)), p = 0..14, row 31 after
// them (tw31 = tw_s + 15*64) - one LDS.128 serves two twiddle multiplications (every instruction less counts: the kernels are
// bound by instruction supply, profiles/r1/microbench_icache.txt).  `src` is the handle's table tw[(q-1)*32 + l].
__device__ __forceinline__ void load_twiddle_pairs(float2* tw_s, const float2* src) {
    for (int i = threadIdx.x; i < kTwF2; i += blockDim.x) {
        const int q = i / 32 + 1, l = i % 32;
        tw_s[q < 31 ? (((q - 1) >> 1) * 32 + l) * 2 + ((q - 1) & 1) : 15 * 64 + l] = src[i];
    }
}

// forward pass 1 with the paired twiddle table
template <int NZ0, int NZ1>
__device__ __forceinline__ void fwd_phase1_tw4(c2 (&z)[32], int lane, f2* scratch, const float4* tw4, const float2* tw31) {
    fft32_sparse<-1, NZ0, NZ1>(z);      // z[t] outside the window support is zero and never read
    real64_post(z);
    float* row0 = reinterpret_cast<float*>(scratch);
    row0[lane] = z[0].x;
    row0[lane + 32] = z[0].y;
    // multiply in place, store afterwards: a product that lives in its own z register is not copied before its store
    // (ptxas moves a store's source aside when the register is about to be reused - two MOVs per twiddle otherwise)
#pragma unroll
    for (int p = 0; p < 15; ++p) {
        const float4 w = tw4[p * 32 + lane];
        z[2 * p + 1] = cmul(z[2 * p + 1], mk2(w.x, w.y));
        z[2 * p + 2] = cmul(z[2 * p + 2], mk2(w.z, w.w));
    }
    z[31] = cmul(z[31], tw31[lane]);
#pragma unroll
    for (int q = 1; q < 32; ++q) scratch[q * kRowStride + lane] = z[q];
}

#endif

}  // namespace nsb
