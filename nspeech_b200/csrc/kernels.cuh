// CUDA kernels of the spectrogram / Griffin-Lim hot path (sm_100a).
//
//   k_analysis<MODE>  : framing + (pre-emphasis) + reflect padding + Hann window + 2048-pt real FFT
//                       -> complex64 STFT, or |.| -> dB -> normalise (+ sparse mel) features
//                       replaces audio.py:31-32,39-42,61-64,106-108,138-151,162-163
//   k_synth<SRC>      : (STFT of y -> phase renormalise x magnitude | given spectrum | magnitude x
//                       given phase) -> inverse real FFT -> window -> overlap-add -> / window-sum
//                       one launch = one `_istft`, or one whole Griffin-Lim iteration
//                       replaces audio.py:77-87,111-113
//   k_prepare_mag     : _denormalize + ref_level_db, _db_to_amp, **power (audio.py:47-48,154-155,166-167)
//   k_deemphasis      : inv_preemphasis (audio.py:35-36) as a blocked scan
//   k_preemphasis, k_elementwise, k_linear_to_mel : the small helpers of the same module
//
// One warp owns one frame; see frame_fft.cuh for the transform itself.
#pragma once
#ifdef NSB_EMULATE            // tests/emu: the same source on CPU threads (test infrastructure, never shipped)
#include "cuda_emu.h"
#define NSB_DYN_SMEM(name) unsigned char* name = nsb_emu::dyn_smem()
#else
#include <cuda_runtime.h>
#define NSB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
#include <stdint.h>
#include "frame_fft.cuh"

namespace nsb {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kMagPitch = 1056;  // floats per frame in the permuted magnitude buffer (1024 slots + Nyquist + pad)

// ragged batch: utterance b owns frames [frame_off[b], frame_off[b+1]) and samples [samp_off[b], samp_off[b+1])
struct Batch {
    const int* frame_off;        // [batch+1]
    const long long* samp_off;   // [batch+1]
    const int* tile_off;         // [batch+1] (synthesis tiles), may be null for analysis
    const int* tile_utt;         // [total tiles] GLOBAL tile index -> GLOBAL utterance index
    const int* group_off;        // [batch+1] groups of C = ceil(win/hop) hops per utterance, GLOBAL prefix sums (k_gl_stream)
    const long long* row_off;    // [batch] or null: output row of utterance b's first feature frame (nsb_features_rows: bucketed, padded batches)
    int group_base;              // global index of this (sub-)batch's first group
    int batch;
    int utt_base;                // global index of this (sub-)batch's first utterance (pointers above are offset by it)
    int frame_base, tile_base;   // global index of this (sub-)batch's first frame / tile: the arrays hold GLOBAL prefix
                                 // sums, so a chunk of a batch is just a pointer offset + these bases (host-side pipelining)
};

struct Plan {                    // device tables owned by the handle
    const float2* tw;            // [31*32]
    const float* win;            // [2048] window padded centrally to n_fft (unscaled)
    const float* rinv;           // [C*hop] (one hop's table repeated over a group) 1 / (n_fft * summed squared window) of the interior,
                                 // 1/n_fft without normalisation (k_gl_stream)
    const float* mel_w;          // mel weights, rows concatenated
    const int* mel_lo;           // [num_mels] first bin of row
    const int* mel_n;            // [num_mels] row length
    const int* mel_ptr;          // [num_mels] offset into mel_w
    // the same filters as straight lines (librosa's triangles ARE linear in the bin index on either side of the peak):
    // segment j = bins [mel_seg[j], mel_seg[j+1]) between band edges j and j+1; row m rises on segment m as
    // c.x + c.y (mel_seg[m+1] - k) and falls on segment m+1 as c.z + c.w (mel_seg[m+2] - k).  Null: use the sparse rows.
    const int* mel_seg;          // [num_mels + 2]
    const float4* mel_coef;      // [num_mels]
    // the same segments cut into at most 96 pieces of balanced length (long segments in two): piece p = bins [x, y), z = float bits of
    // (its segment's end - y); mel_segp[m] = (first, second piece of segment m, first, second piece of segment m+1), 96 = an all-zero slot
    const int4* mel_piece;       // [96] or null
    const int4* mel_segp;        // [num_mels]
    int n_fft, hop, win_len, lo; // window support is n in [lo, lo + win_len) of the n_fft-long frame
    int origin;                  // frame k starts at sample k*hop - origin
    int norm_wss;                // 1: divide the overlap-add by the summed squared window (librosa.istft), 0: do not (tf inverse_stft)
    int num_mels;
    int prune;                   // 0: none, 1: support inside n in [512,1536), 2: support inside n in [0,1024)
    // librosa geometry (audio.py:106-113): window padded centrally, lo = (n_fft - win)/2, origin = n_fft/2, reflect padding.
    // tf.contrib.signal geometry (audio.py:116-123): lo = 0, origin = 0, frames never leave the signal.
};

template <int PRUNE> struct PruneRange { static constexpr int t0 = PRUNE == 1 ? 8 : 0, t1 = PRUNE == 1 ? 24 : (PRUNE == 2 ? 16 : 32); };

// Overlap-add of one frame whose window support n in [LO, LO + 1000) is a compile-time fact (the reference's default
// hparams: LO = 524 for the librosa geometry, 0 for the tf.contrib.signal one).  Lane l holds the samples n = 64 t + l (.x)
// and 64 t + 32 + l (.y); ap = accumulator address of sample n = lane.  Slots fully inside the support are one packed
// multiply-add on (acc[n], acc[n+32]); the two edge slots predicate on the lane.  win_at(t) = (w[64 t + l], w[64 t + 32 + l]).
template <int LO, int T0, int T1, typename WinAt>
__device__ __forceinline__ void ola_fixed_support(float* ap, const c2 (&z)[32], int lane, WinAt win_at) {
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        if (t >= T0 && t < T1) {
            const int r0 = LO - 64 * t, r1 = r0 + 1000;      // lanes whose .x sample is inside the support: [r0, r1)
            const int i0 = r0 - 32, i1 = r1 - 32;            // ... .y sample
            const bool fre = (r0 <= 0 && r1 >= 32), fim = (i0 <= 0 && i1 >= 32);
            const c2 w = win_at(t);
            if (fre && fim) {
                const c2 rr = p_fma(z[t], w, mk2(ap[64 * t], ap[64 * t + 32]));
                ap[64 * t] = rr.x;
                ap[64 * t + 32] = rr.y;
            } else {
                if (r1 > 0 && r0 < 32 && (fre || (lane >= r0 && lane < r1))) ap[64 * t] = fmaf(z[t].x, w.x, ap[64 * t]);
                if (i1 > 0 && i0 < 32 && (fim || (lane >= i0 && lane < i1))) ap[64 * t + 32] = fmaf(z[t].y, w.y, ap[64 * t + 32]);
            }
        }
    }
}

__device__ __forceinline__ int find_segment(const int* __restrict__ off, int n, int v) {
    // largest b in [0,n) with off[b] <= v
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// np.pad(mode='reflect') index map for a signal of length L (any i, multiple reflections); 32-bit on purpose
// (utterances are far below 2^31 samples; 64-bit division is a subroutine call on the GPU)
__device__ __forceinline__ int reflect_index(int i, int L) {
    if (L == 1) return 0;
    int period = 2 * (L - 1);
    int m = i % period;
    if (m < 0) m += period;
    return m < L ? m : period - m;
}

// ---- frame load: windowed samples of frame k into the lane registers -----------------------
// re[t] = w[n] * s(start + n), n = 64 t + lane ; im[t] likewise with n + 32.   s = (pre-emphasised) signal
// COHERENT: the signal is rewritten by other SMs inside the same launch (iteration-fused Griffin-Lim): read it from L2
// (ld.global.cg), never through the non-coherent L1 path
template <bool COHERENT>
__device__ __forceinline__ float ld_sig(const float* p) { return COHERENT ? __ldcg(p) : __ldg(p); }

template <bool PREEMPH, bool COHERENT = false>
__device__ __forceinline__ float sample_at(const float* __restrict__ x, int L, int i, float p) {
    int m = reflect_index(i, L);
    float v = ld_sig<COHERENT>(x + m);
    if (PREEMPH) { if (m > 0) v = fmaf(-p, ld_sig<COHERENT>(x + m - 1), v); }
    return v;
}

// `stage` is the warp's scratch tile (>= 2048 floats): frames that touch the utterance's ends (reflect padding)
// are gathered through it by a rolled loop, so the rare path costs a few dozen instructions of code instead of
// an unrolled copy of the index arithmetic per register
template <bool PREEMPH, int PRUNE, bool COHERENT = false>
__device__ __forceinline__ void load_frame(c2 (&z)[32], const float* __restrict__ x, long long L, long long start,
                                           const float* __restrict__ win_s, int lane, float p, float* stage) {
    constexpr int t0 = PruneRange<PRUNE>::t0, t1 = PruneRange<PRUNE>::t1;
    const long long first = start + 64 * t0, last = start + 64 * t1;   // [first, last) touched
    const bool interior = (first >= (PREEMPH ? 1 : 0)) && (last <= L);
    if (interior) {
        const float* xs = x + start + lane;
        // (x[n-1] for the pre-emphasis by rotating the loaded samples one lane - two shuffles instead of two more loads per 64 samples -
        // measured slower: 0.954 vs 0.901 ms, the shuffle sits between the load and its first use; profiles/r2/features_ab_shuffle_preemphasis.txt)
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            if (t >= t0 && t < t1) {
                float a = ld_sig<COHERENT>(xs + 64 * t), b = ld_sig<COHERENT>(xs + 64 * t + 32);
                if (PREEMPH) {
                    a = fmaf(-p, ld_sig<COHERENT>(xs + 64 * t - 1), a);
                    b = fmaf(-p, ld_sig<COHERENT>(xs + 64 * t + 31), b);
                }
                z[t] = p_mul(mk2(a, b), mk2(win_s[64 * t + lane], win_s[64 * t + 32 + lane]));
            } else { z[t] = mk2(0.f, 0.f); }
        }
    } else {
#pragma unroll 1
        for (int n = 64 * t0 + lane; n < 64 * t1; n += 32) stage[n] = sample_at<PREEMPH, COHERENT>(x, (int)L, (int)start + n, p);
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            if (t >= t0 && t < t1) {
                z[t] = p_mul(mk2(stage[64 * t + lane], stage[64 * t + 32 + lane]), mk2(win_s[64 * t + lane], win_s[64 * t + 32 + lane]));
            } else { z[t] = mk2(0.f, 0.f); }
        }
        __syncwarp();
    }
}

// =============================================================================================
// analysis
// =============================================================================================
enum { ANALYSIS_COMPLEX = 0, ANALYSIS_FEATURES = 1 };

struct AnalysisParams {
    Plan plan;
    Batch batch;
    const float* wav;      // packed samples
    float2* out_complex;   // [frames][1025]
    float* out_lin;        // [frames][1025] or null
    float* out_mel;        // [frames][num_mels] or null
    int total_frames;
    int rows_per_utt;      // > 0: feature row of frame k of utterance b is (utt_base + b) * rows_per_utt + k (padded batch layout of the
                           // feeder, datasets/datafeeder.py:205-220) instead of the packed global frame index
    float preemph;         // coefficient (only read when PREEMPH)
    float ref_level_db, min_level_db;
    // _normalize(_amp_to_db(amp) - ref) = clip(log2(max(1e-5, amp)) * db_scale + db_offset, 0, 1) with
    // db_scale = 20 log10(2) / (-min_level_db), db_offset = (-ref - min_level_db) / (-min_level_db) (host, from double)
    float db_scale, db_offset_lin, db_offset_mel;
    int mel_split;         // 1: the moment loop over the balanced pieces (NSB_OPT_MEL_LINES 3)
    int mel_skew;          // 1: magnitude row with a pad word per 32 bins (production; NSB_OPT_MEL_LINES 1 is the A/B switch without it)
    int* status;           // device flag: bit0 = non-finite input
};

// MUFU approximations for the feature epilogue (1025 logarithms and square roots per frame; log10f + sqrtf were a third of
// the kernel's 4,000 instructions per frame): lg2.approx has an absolute error below 2^-22, i.e. < 2e-6 dB, sqrt.approx
// about one ulp - far inside the 1e-5 relative-L2 bar of the features.  131 -> 173 M frames/s.
// (Tried and rejected, both slower: the mel non-zeros cut into equal lane shares with partial sums in shared memory
// (114 M/s), whole rows dealt to the lanes by non-zero count (165 M/s) - the row-per-lane loop is bound by the latency of
// its dependent loads, not by its ~190 steps.)
__device__ __forceinline__ float lg2_approx(float x) {
#ifdef NSB_EMULATE
    return log2f(x);
#else
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
__device__ __forceinline__ float sqrt_approx(float x) {
#ifdef NSB_EMULATE
    return sqrtf(x);
#else
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
__device__ __forceinline__ float amp_to_db_norm_fast(float amp, float scale, float offset) {
    const float v = fmaf(lg2_approx(fmaxf(1e-5f, amp)), scale, offset);
    return fminf(fmaxf(v, 0.f), 1.f);
}

__device__ __forceinline__ float amp_to_db_norm(float amp, float ref_db, float min_db) {
    // _normalize(_amp_to_db(amp) - ref): 20*log10(max(1e-5, amp)), (S - min)/(-min), clip [0,1]
    float db = 20.0f * log10f(fmaxf(1e-5f, amp)) - ref_db;
    float v = (db - min_db) / (-min_db);
    return fminf(fmaxf(v, 0.f), 1.f);
}

// a kernel parameter pinned in a register: ptxas otherwise re-reads it from the constant bank at every use (LDC / LDCU per bin
// of the feature epilogue: 74 of its instructions per frame)
__device__ __forceinline__ float pin_reg(float v) {
#ifndef NSB_EMULATE
    asm volatile("mov.f32 %0, %0;" : "+f"(v));
#endif
    return v;
}

// position of bin kb in the warp's magnitude row: one pad word per 32 bins.  The mel moment loop gives every lane its own
// segment of consecutive bins; the segments of neighbouring lanes start 4, 8, ... bins apart (the low mel bands), i.e. lanes
// 8 (or 4) apart would hit the same bank at every step - a quarter of the kernel's shared-memory wavefronts were such
// conflicts (profiles/r1/ncu_full_k_analysis_features.txt).  With the pad they walk different banks.
constexpr int kMagSkewLen = 1025 + 32 + 3;          // 1060 floats: skewed row, then the moments
template <bool SKEW> __device__ __forceinline__ int mag_skew(int kb) { return SKEW ? kb + (kb >> 5) : kb; }

// Feature epilogue of one frame (registers of fwd_phase2 -> linear dB row, mel dB row).
// Slot p of lane l >= 1 holds bin l + 64p (p < 16) or (64 - l) + 64(31 - p) (p >= 16): from a per-lane base every address of the
// unrolled loops is an immediate offset (64 floats in the output row, 66 in the skewed magnitude row), stores are predicated
// off for lane 0.  Lane 0's slots are the bins 32 j: it hands them over through shared memory and lane j finishes bin 32 j.
// The linear dB comes from |D|^2 (10 log10 |D|^2 = 20 log10 |D|): the square root is only on the mel path and the two
// MUFU operations of a bin no longer depend on each other.
template <bool LIN, bool SKEW>
__device__ __forceinline__ void feature_epilogue(const AnalysisParams& P, c2 (&z)[32], int lane, float2* scratch, float* o_lin, float* o_mel,
                                                 float db_scale, float db_off_lin, float db_off_mel, bool& bad, const float4* coef_s,
                                                 const int4* piece_s = nullptr, const int4* segp_s = nullptr) {
    float* magrow = reinterpret_cast<float*>(scratch);
    float* mom = magrow + kMagSkewLen;                             // [num_mels + 1][2] (the host checks that it fits the scratch tile)
    float2* xch = reinterpret_cast<float2*>(magrow + kMagSkewLen + 2 * 96 + 4);   // lane 0's 32 slots (8-byte aligned: 1256 floats in)
    const float half_scale = 0.5f * db_scale;
    const bool main = lane != 0;
    constexpr int MS = SKEW ? 66 : 64;                             // stride of the magnitude row per slot
    float* mr_lo = magrow + lane;                                  // bin l + 64p        -> skewed l + 66p
    float* mr_hi = magrow + (SKEW ? (66 * 32 - 1) : 2048) - lane;  // bin 2048 - l - 64p -> skewed 66(32 - p) - l - 1
    float* ol_lo = o_lin + lane;
    float* ol_hi = o_lin + 2048 - lane;
    float chk = 0.f;
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 32; ++q) xch[q] = z[q];
    }
#pragma unroll
    for (int p = 0; p < 32; ++p) {
        const float m2 = fmaf(z[p].x, z[p].x, z[p].y * z[p].y);
        if (main) {
            chk += m2;
            const float mg = sqrt_approx(m2);
            if (p < 16) mr_lo[MS * p] = mg; else mr_hi[-MS * p] = mg;
            if (LIN) {
                const float v = __saturatef(fmaf(lg2_approx(fmaxf(1e-10f, m2)), half_scale, db_off_lin));
                if (p < 16) ol_lo[64 * p] = v; else ol_hi[-64 * p] = v;
            }
        }
    }
    __syncwarp();
    {
        // bins 32 j from lane 0's slots: lane 0 itself takes the real pair (DC, Nyquist), lane j >= 1 the complex X[32 j]
        const float2 v = xch[lane];
        const float m2 = main ? fmaf(v.x, v.x, v.y * v.y) : v.x * v.x;
        chk += m2;
        const float mg = main ? sqrt_approx(m2) : fabsf(v.x);
        magrow[(SKEW ? 33 : 32) * lane] = mg;
        if (LIN) o_lin[32 * lane] = __saturatef(fmaf(lg2_approx(fmaxf(1e-10f, m2)), half_scale, db_off_lin));
        if (!main) {
            const float n2 = v.y * v.y;
            chk += n2;
            magrow[1024 + (SKEW ? 32 : 0)] = fabsf(v.y);
            if (LIN) o_lin[1024] = __saturatef(fmaf(lg2_approx(fmaxf(1e-10f, n2)), half_scale, db_off_lin));
        }
    }
    // non-finite input shows up in every bin of its frames: ONE test per frame and lane on the sum of the squared magnitudes
    // (squares beyond 3e38, i.e. |D| > 1e19, also count as non-finite)
    bad |= !isfinite(chk);
    __syncwarp();
    if (o_mel && P.plan.mel_seg) {
        // mel[m] = sum_k W[m,k] |D[k]| with W[m,.] a triangle: per segment j two moments S0 = sum |D|, S1 = sum (end - k) |D|
        // (every bin read ONCE, no weight loads), then every row is four multiply-adds of its two segments' moments.
        // a1 = sum of the running sums = sum (k1 - k) |D[k]|: the first moment counted from the segment's END costs one add
        // per bin; the host folds the change of origin into the coefficients.  The adds run in bin order whatever the unrolling.
        const int M = P.plan.num_mels;
        if (piece_s) {
            // balanced form: three pieces per lane, piece p = lane + 32 r (the host sorted them by length, so the lanes of a round
            // run about equally long); a piece of a segment cut in two measures its first moment from its own end and moves it
            // to the segment's end with one multiply-add.  Row m then adds the (at most two) pieces of its two segments.
            float2* pm = reinterpret_cast<float2*>(mom);
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
                const int4 t = piece_s[lane + 32 * r];
                float a0 = 0.f, a1 = 0.f;
                // (the piece as two pointer walks, up to the row's next pad word and after it, instead of the skewed index per bin: measured
                //  slower, 0.954 vs 0.846 ms - two loops of divergent trip counts; profiles/r2/features_ab_pieces_pointer_walk.txt)
                for (int kb = t.x; kb < t.y; ++kb) { a0 += magrow[mag_skew<SKEW>(kb)]; a1 += a0; }
                pm[lane + 32 * r] = make_float2(a0, fmaf(__int_as_float(t.z), a0, a1));
            }
            if (lane == 0) pm[96] = make_float2(0.f, 0.f);
            __syncwarp();
            for (int m = lane; m < M; m += 32) {
                const float4 c = coef_s[m];
                const int4 q = segp_s[m];
                const float2 r0 = pm[q.x], r1 = pm[q.y], f0 = pm[q.z], f1 = pm[q.w];
                float acc = c.x * (r0.x + r1.x);
                acc = fmaf(c.y, r0.y + r1.y, acc); acc = fmaf(c.z, f0.x + f1.x, acc); acc = fmaf(c.w, f0.y + f1.y, acc);
                o_mel[m] = amp_to_db_norm_fast(fmaxf(acc, 0.f), db_scale, db_off_mel);
            }
            __syncwarp();
            return;
        }
        for (int j = lane; j <= M; j += 32) {
            const int k0 = __ldg(P.plan.mel_seg + j), k1 = __ldg(P.plan.mel_seg + j + 1);
            float a0 = 0.f, a1 = 0.f;
            // (four bins per trip with a predicated tail - four loads in flight - measured slower: 0.969 vs 0.930 ms, profiles/r2/features_ab.txt)
            for (int kb = k0; kb < k1; ++kb) { a0 += magrow[mag_skew<SKEW>(kb)]; a1 += a0; }
            mom[2 * j] = a0; mom[2 * j + 1] = a1;
        }
        __syncwarp();
        for (int m = lane; m < M; m += 32) {
            const float4 c = coef_s ? coef_s[m] : __ldg(P.plan.mel_coef + m);
            const float2 r = *reinterpret_cast<const float2*>(mom + 2 * m), f = *reinterpret_cast<const float2*>(mom + 2 * m + 2);
            float acc = c.x * r.x;
            acc = fmaf(c.y, r.y, acc); acc = fmaf(c.z, f.x, acc); acc = fmaf(c.w, f.y, acc);
            o_mel[m] = amp_to_db_norm_fast(fmaxf(acc, 0.f), db_scale, db_off_mel);   // melspectrogram subtracts no ref_level_db (audio.py:63)
        }
    } else if (o_mel) {
        for (int m = lane; m < P.plan.num_mels; m += 32) {
            const int lo = __ldg(P.plan.mel_lo + m), n = __ldg(P.plan.mel_n + m);
            const float* w = P.plan.mel_w + __ldg(P.plan.mel_ptr + m);
            float acc = 0.f;
            for (int i = 0; i < n; ++i) acc = fmaf(__ldg(w + i), magrow[mag_skew<SKEW>(lo + i)], acc);
            o_mel[m] = amp_to_db_norm_fast(acc, db_scale, db_off_mel);   // melspectrogram subtracts no ref_level_db (audio.py:63)
        }
    }
    __syncwarp();
}

template <int MODE, bool PREEMPH, int PRUNE>
__global__ void __launch_bounds__(kThreads, 2) k_analysis(AnalysisParams P) {
    NSB_DYN_SMEM(smem_raw);
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float* win_s = reinterpret_cast<float*>(tw_s + kTwF2);
    float2* scratch_all = reinterpret_cast<float2*>(win_s + kNfft);
    // the mel rows' line coefficients, once per CTA (a load from L2 per row and frame otherwise: 4 % of the stall samples)
    float4* coef_s = reinterpret_cast<float4*>(scratch_all + kScratchF2 * kWarpsPerCta);
    const bool mel_tables = MODE == ANALYSIS_FEATURES && P.plan.mel_seg && P.plan.num_mels <= 96;
    int4* piece_s = reinterpret_cast<int4*>(coef_s + 96);
    int4* segp_s = piece_s + 96;
    const bool mel_split = mel_tables && P.mel_split && P.plan.mel_piece;
    if (mel_tables)
        for (int i = threadIdx.x; i < P.plan.num_mels; i += kThreads) coef_s[i] = P.plan.mel_coef[i];
    if (mel_split) {
        for (int i = threadIdx.x; i < 96; i += kThreads) piece_s[i] = P.plan.mel_piece[i];
        for (int i = threadIdx.x; i < P.plan.num_mels; i += kThreads) segp_s[i] = P.plan.mel_segp[i];
    }
    load_twiddle_pairs(tw_s, P.plan.tw);
    const float4* tw4 = reinterpret_cast<const float4*>(tw_s);
    const float2* tw31 = tw_s + 15 * 64;
    // 0.5 folds the forward transform's factor 2 (frame_fft.cuh) so the registers hold rfft exactly
    for (int i = threadIdx.x; i < kNfft; i += kThreads) win_s[i] = 0.5f * P.plan.win[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* scratch = scratch_all + warp * kScratchF2;
    const int hop = P.plan.hop;
    bool bad = false;
    const float db_scale = pin_reg(P.db_scale), db_off_lin = pin_reg(P.db_offset_lin), db_off_mel = pin_reg(P.db_offset_mel);
    // every CTA takes one contiguous block of frames (its 8 warps side by side on 8 consecutive frames, which share three
    // quarters of their samples in L1): the utterance index then only creeps forward.  With the frames dealt round-robin over
    // the grid every step landed in another utterance and paid a binary search of dependent loads (10 % of the stall samples).
    const int per_cta = (P.total_frames + (int)gridDim.x - 1) / (int)gridDim.x;
    const int f_begin = P.batch.frame_base + (int)blockIdx.x * per_cta;
    const int f_end = min(f_begin + per_cta, P.batch.frame_base + P.total_frames);
    int b = (f_begin + warp < f_end) ? find_segment(P.batch.frame_off, P.batch.batch, f_begin + warp) : 0;
    for (int f = f_begin + warp; f < f_end; f += kWarpsPerCta) {
        while (f >= __ldg(P.batch.frame_off + b + 1)) ++b;
        const int k = f - __ldg(P.batch.frame_off + b);
        const long long s_off = __ldg(P.batch.samp_off + b);
        const long long L = __ldg(P.batch.samp_off + b + 1) - s_off;
        c2 z[32];
        load_frame<PREEMPH, PRUNE>(z, P.wav + s_off, L, (long long)k * hop - P.plan.origin, win_s, lane, P.preemph,
                                   reinterpret_cast<float*>(scratch));
        // (an L2 prefetch of the next frame's new hop here: 288.5 -> 289.3 M frames/s, noise - not kept)
        fwd_phase1_tw4<PruneRange<PRUNE>::t0, PruneRange<PRUNE>::t1>(z, lane, scratch, tw4, tw31);
        __syncwarp();
        fwd_phase2(z, lane, scratch);
        __syncwarp();
        if (MODE == ANALYSIS_COMPLEX) {
            float2* o = P.out_complex + (size_t)f * kBins;
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                bad |= !(isfinite(z[p].x) && isfinite(z[p].y));
                if (lane == 0 && p == 0) {
                    o[0] = make_float2(z[0].x, 0.f);
                    o[1024] = make_float2(z[0].y, 0.f);
                } else {
                    int kb = bin_of(lane, p);
                    o[kb] = make_float2(z[p].x, (lane != 0 && p >= 16) ? -z[p].y : z[p].y);
                }
            }
        } else {
            const long long orow = P.batch.row_off ? __ldg(P.batch.row_off + b) + k
                                 : P.rows_per_utt > 0 ? (long long)(P.batch.utt_base + b) * P.rows_per_utt + k : (long long)f;
            float* o_lin = P.out_lin ? P.out_lin + (size_t)orow * kBins : nullptr;
            float* o_mel = P.out_mel ? P.out_mel + (size_t)orow * P.plan.num_mels : nullptr;
            if (P.mel_skew) {
                if (o_lin) feature_epilogue<true, true>(P, z, lane, scratch, o_lin, o_mel, db_scale, db_off_lin, db_off_mel, bad, mel_tables ? coef_s : nullptr, mel_split ? piece_s : nullptr, segp_s);
                else feature_epilogue<false, true>(P, z, lane, scratch, o_lin, o_mel, db_scale, db_off_lin, db_off_mel, bad, mel_tables ? coef_s : nullptr, mel_split ? piece_s : nullptr, segp_s);
            } else {
                if (o_lin) feature_epilogue<true, false>(P, z, lane, scratch, o_lin, o_mel, db_scale, db_off_lin, db_off_mel, bad, mel_tables ? coef_s : nullptr, mel_split ? piece_s : nullptr, segp_s);
                else feature_epilogue<false, false>(P, z, lane, scratch, o_lin, o_mel, db_scale, db_off_lin, db_off_mel, bad, mel_tables ? coef_s : nullptr, mel_split ? piece_s : nullptr, segp_s);
            }
        }
    }
    if (bad) atomicOr(P.status, 1);
}

// =============================================================================================
// synthesis: istft / Griffin-Lim iteration
// =============================================================================================
enum { SRC_Y = 0, SRC_SPEC = 1, SRC_MAGPHASE = 2, SRC_MAGRAND = 3, SRC_MAGZERO = 4 };

struct SynthParams {
    Plan plan;
    Batch batch;
    const float* y_in;        // SRC_Y: packed waveform estimate
    const float* mag;         // permuted magnitudes [frames][kMagPitch]   (SRC_Y, SRC_MAGPHASE, SRC_MAGRAND)
    const float2* spec;       // SRC_SPEC: spectrum, SRC_MAGPHASE: unit phase; utterance b's block of 1025*T_b elements
    int spec_bin_major;       //   starts at 1025*frame_off[b]; inside it [T_b][1025] (0) or [1025][T_b] (1)
    float* y_out;             // packed
    int tile_hops;            // H
    int colours;              // ceil(win/hop)
    unsigned long long seed;
    int* status;
};

__device__ __forceinline__ uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t& hi) {
    unsigned long long r = (unsigned long long)a * b;
    hi = (uint32_t)(r >> 32);
    return (uint32_t)r;
}
// Philox-4x32-10 (Salmon et al. 2011), counter-based: one call -> four 32-bit words
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, hi1;
        uint32_t lo0 = mulhilo(0xD2511F53u, ctr.x, hi0);
        uint32_t lo1 = mulhilo(0xCD9E8D57u, ctr.z, hi1);
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}

// phase renormalisation of one slot: z <- S * z/|z| ; z == 0 -> S (np.angle(0) = 0, audio.py:85)
__device__ __forceinline__ void renorm(c2& z, float S) {
    float m2 = fmaf(z.x, z.x, z.y * z.y);
    if (m2 < 1e-30f || m2 > 1e30f) {              // rare: rescale to dodge under/overflow of the square
        float sc = (m2 < 1e-30f) ? 1.8446744e19f : 5.4210109e-20f;   // 2^64, 2^-64
        c2 w = cscale(z, sc);
        m2 = fmaf(w.x, w.x, w.y * w.y);
        if (m2 == 0.f) { z = mk2(S, 0.f); return; }
        z = cscale(w, rsqrtf(m2) * S);
        return;
    }
    z = cscale(z, rsqrtf(m2) * S);
}

template <int SRC, int PRUNE>
__global__ void __launch_bounds__(kThreads, 2) k_synth(SynthParams P) {
    NSB_DYN_SMEM(smem_raw);
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float* win_s = reinterpret_cast<float*>(tw_s + kTwF2);
    float* rinv_s = win_s + kNfft;                                   // [hop] interior 1/window-sum
    float* acc = rinv_s + P.plan.hop;                                // [H*hop]
    // scratch must be 8-byte aligned: hop and H*hop are arbitrary, so round up
    size_t acc_end = (size_t)(acc + (size_t)P.tile_hops * P.plan.hop - reinterpret_cast<float*>(smem_raw));
    acc_end = (acc_end + 3) & ~(size_t)3;
    float2* scratch_all = reinterpret_cast<float2*>(reinterpret_cast<float*>(smem_raw) + acc_end);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* scratch = scratch_all + warp * kScratchF2;
    const int hop = P.plan.hop, win = P.plan.win_len, lo = P.plan.lo;
    const int a = P.plan.origin - lo;      // frame k's window support starts at sample k*hop - a
    const int C = P.colours, H = P.tile_hops;

    const int tile_g = P.batch.tile_base + (int)blockIdx.x;
    const int b = __ldg(P.batch.tile_utt + tile_g) - P.batch.utt_base;
    const int tile = tile_g - __ldg(P.batch.tile_off + b);
    const int f_off = __ldg(P.batch.frame_off + b);
    const int T = __ldg(P.batch.frame_off + b + 1) - f_off;
    const long long s_off = __ldg(P.batch.samp_off + b);
    const long long L = __ldg(P.batch.samp_off + b + 1) - s_off;       // hop*(T-1) (librosa) or hop*(T-1)+win (tf)
    const int n_hops = (int)((L + hop - 1) / hop);
    const int h0 = tile * H, h1 = min(h0 + H, n_hops);
    const long long s0 = (long long)h0 * hop, s1 = min((long long)h1 * hop, L);
    const int n_out = (int)(s1 - s0);

    for (int i = threadIdx.x; i < kTwF2; i += kThreads) tw_s[i] = P.plan.tw[i];
    for (int i = threadIdx.x; i < kNfft; i += kThreads) win_s[i] = P.plan.win[i];
    for (int i = threadIdx.x; i < n_out; i += kThreads) acc[i] = 0.f;
    __syncthreads();
    // interior reciprocal window-sum for each offset j inside a hop (all covering frames exist)
    for (int j = threadIdx.x; j < hop; j += kThreads) {
        float s = 0.f;
        for (int idx = (j + a) % hop; idx < win; idx += hop) { float w = win_s[lo + idx]; s = fmaf(w, w, s); }
        rinv_s[j] = s > 1.17549435e-38f ? 1.0f / s : 1.0f;
    }

    // frames whose window support [k*hop - a, k*hop - a + win) meets [s0, s1)
    long long num = s0 + a - win;                       // k*hop > num
    int k_min = (int)(num >= 0 ? num / hop + 1 : -((-num - 1) / hop + 1) + 1);
    if (k_min < 0) k_min = 0;
    int k_max = (int)((s1 + a + hop - 1) / hop - 1);     // k*hop < s1 + a
    if (k_max > T - 1) k_max = T - 1;

    bool bad = false;
    for (int c = 0; c < C; ++c) {
        // first frame >= k_min with k % C == c
        int kf = k_min + ((c - k_min % C) + C) % C;
        for (int k = kf + C * warp; k <= k_max; k += C * kWarpsPerCta) {
            const int fg = f_off + k;                   // global frame index
            c2 z[32];
            if (SRC == SRC_Y) {
                load_frame<false, PRUNE>(z, P.y_in + s_off, L, (long long)k * hop - P.plan.origin, win_s, lane, 0.f,
                                         reinterpret_cast<float*>(scratch));
                fwd_phase1(z, lane, scratch, tw_s);
                __syncwarp();
                fwd_phase2(z, lane, scratch);
                __syncwarp();
                const float4* mp = reinterpret_cast<const float4*>(P.mag + (size_t)fg * kMagPitch);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    float4 S = __ldg(mp + g * 32 + lane);
                    if (g == 0) {
                        if (lane == 0) {    // packed real DC / Nyquist: phase of a real number is its sign
                            float Sn = __ldg(P.mag + (size_t)fg * kMagPitch + 1024);
                            z[0] = mk2((z[0].x < 0.f) ? -S.x : S.x, (z[0].y < 0.f) ? -Sn : Sn);
                        } else {
                            renorm(z[0], S.x);
                        }
                    } else {
                        renorm(z[4 * g], S.x);
                    }
                    renorm(z[4 * g + 1], S.y);
                    renorm(z[4 * g + 2], S.z);
                    renorm(z[4 * g + 3], S.w);
                }
            } else if (SRC == SRC_SPEC || SRC == SRC_MAGPHASE) {
                const float2* sp;
                long long st_f;
                if (P.spec_bin_major) { sp = P.spec + (size_t)f_off * kBins + k; st_f = T; }
                else { sp = P.spec + (size_t)f_off * kBins + (size_t)k * kBins; st_f = 1; }
                const float* mrow = P.mag + (size_t)fg * kMagPitch;
#pragma unroll
                for (int p = 0; p < 32; ++p) {
                    if (lane == 0 && p == 0) {
                        float2 v0 = __ldg(sp), v1 = __ldg(sp + 1024 * st_f);
                        z[0] = mk2(v0.x, v1.x);              // imaginary parts of DC / Nyquist are dropped (irfft semantics)
                        if (SRC == SRC_MAGPHASE) z[0] = p_mul(z[0], mk2(__ldg(mrow), __ldg(mrow + 1024)));
                    } else {
                        float2 v = __ldg(sp + (long long)bin_of(lane, p) * st_f);
                        if (lane != 0 && p >= 16) v.y = -v.y;
                        if (SRC == SRC_MAGPHASE) v = cscale(v, __ldg(mrow + ((p >> 2) * 32 + lane) * 4 + (p & 3)));
                        z[p] = v;
                    }
                }
            } else if (SRC == SRC_MAGZERO) {   // zero phase: S_complex = S + 0j (the TF twin's start, audio.py:97-98)
                const float4* mp = reinterpret_cast<const float4*>(P.mag + (size_t)fg * kMagPitch);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    float4 S = __ldg(mp + g * 32 + lane);
                    z[4 * g] = mk2(S.x, 0.f); z[4 * g + 1] = mk2(S.y, 0.f); z[4 * g + 2] = mk2(S.z, 0.f); z[4 * g + 3] = mk2(S.w, 0.f);
                }
                if (lane == 0) z[0].y = __ldg(P.mag + (size_t)fg * kMagPitch + 1024);     // packed (DC, Nyquist)
            } else {  // SRC_MAGRAND: magnitude x exp(2*pi*i*u), u ~ Philox keyed by seed, counter = (frame, lane, group)
                const float* mrow = P.mag + (size_t)fg * kMagPitch;
                uint2 key = make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32));
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    uint4 r = philox4x32(make_uint4((uint32_t)fg, (uint32_t)(lane * 8 + g), 0x6e737062u, 0u), key);
                    float4 S = __ldg(reinterpret_cast<const float4*>(mrow) + g * 32 + lane);
                    uint32_t rr[4] = {r.x, r.y, r.z, r.w};
                    float SS[4] = {S.x, S.y, S.z, S.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float u = (float)(rr[e] >> 8) * (1.0f / 16777216.0f);
                        float sn, cs;
                        sincospif(2.0f * u, &sn, &cs);
                        z[4 * g + e] = mk2(SS[e] * cs, SS[e] * sn);
                    }
                }
                if (lane == 0) {   // packed DC / Nyquist keep only the real part of S*exp(i*phi)
                    uint4 r = philox4x32(make_uint4((uint32_t)fg, 0xffffffffu, 0x6e737062u, 0u), key);
                    float sn, cs;
                    sincospif(2.0f * (float)(r.x >> 8) * (1.0f / 16777216.0f), &sn, &cs);
                    z[0].y = __ldg(mrow + 1024) * cs;
                }
            }
            inv_phase1(z, lane, scratch, tw_s);
            __syncwarp();
            inv_phase2(z, lane, scratch);
            __syncwarp();
            // windowed overlap-add into the tile.  Same-colour frames have disjoint window SUPPORTS, so a plain
            // read-modify-write is race-free only if each warp touches nothing outside its support: per-lane
            // bitmasks of the valid t (n = 64 t + lane [+32] inside [lo, lo + win)).
            const long long base = (long long)k * hop - P.plan.origin - s0;  // tile-local index of n = 0
            constexpr int t0 = PruneRange<PRUNE>::t0, t1 = PruneRange<PRUNE>::t1;
            unsigned mre, mim;
            {
                int a0 = max((lo - lane + 63) >> 6, 0), a1 = min(max((lo + win - lane + 63) >> 6, 0), 32);
                int b0 = max((lo - lane - 32 + 63) >> 6, 0), b1 = min(max((lo + win - lane - 32 + 63) >> 6, 0), 32);
                mre = (a1 > a0) ? ((0xffffffffu >> (32 - (a1 - a0))) << a0) : 0u;
                mim = (b1 > b0) ? ((0xffffffffu >> (32 - (b1 - b0))) << b0) : 0u;
            }
            const bool inside = (base + lo >= 0) && (base + lo + win <= n_out);
            if (inside) {
                float* ap = acc + base + lane;
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    if (t >= t0 && t < t1) {
                        if (mre & (1u << t)) ap[64 * t] = fmaf(z[t].x, win_s[64 * t + lane], ap[64 * t]);
                        if (mim & (1u << t)) ap[64 * t + 32] = fmaf(z[t].y, win_s[64 * t + 32 + lane], ap[64 * t + 32]);
                    }
                }
            } else {
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    if (t >= t0 && t < t1) {
                        long long i0 = base + 64 * t + lane, i1 = i0 + 32;
                        if ((mre & (1u << t)) && i0 >= 0 && i0 < n_out) acc[i0] = fmaf(z[t].x, win_s[64 * t + lane], acc[i0]);
                        if ((mim & (1u << t)) && i1 >= 0 && i1 < n_out) acc[i1] = fmaf(z[t].y, win_s[64 * t + 32 + lane], acc[i1]);
                    }
                }
            }
        }
        __syncthreads();
    }
    // normalise by the summed squared window (librosa.istft) or not at all (tf inverse_stft) and store
    float* yo = P.y_out + s_off + s0;
    for (int j = threadIdx.x; j < hop; j += kThreads) {
        const int dj = (j + a) / hop, rj = (j + a) - dj * hop;     // newest covering frame is h + dj, window index rj
        const float ri = rinv_s[j];
        int ncover = 0;
        for (int idx = rj; idx < win; idx += hop) ++ncover;
        for (int h = h0; h < h1; ++h) {
            const int i = (h - h0) * hop + j;
            if (i >= n_out) break;                                  // the last hop of a tf-geometry signal may be partial
            const int k_hi = h + dj, k_lo = k_hi - (ncover - 1);
            float v = acc[i] * (1.0f / (float)kNfft);
            bad |= !isfinite(v);
            if (P.plan.norm_wss) {
                if (k_lo >= 0 && k_hi <= T - 1) {
                    v *= ri;
                } else {
                    float s = 0.f;
                    int kk = k_hi;
                    for (int idx = rj; idx < win; idx += hop, --kk)
                        if (kk >= 0 && kk <= T - 1) { float w = win_s[lo + idx]; s = fmaf(w, w, s); }
                    if (s > 1.17549435e-38f) v /= s;
                }
            }
            yo[i] = v;
        }
    }
    if (bad) atomicOr(P.status, 1);
}

// =============================================================================================
// magnitude preparation: normalised spectrogram (or raw magnitude) -> permuted float32 magnitudes
// =============================================================================================
struct PrepParams {
    Batch batch;
    const float* in;          // packed per utterance (block of F*T_b elements at F*frame_off[b])
    int bin_major;            // 0: [T][F] frame-major, 1: [F][T] bin-major (per utterance)
    int denorm;               // 1: apply _denormalize + ref, _db_to_amp, **power ; 0: |S| as is
    double min_level_db, ref_level_db, power;
    double e_slope, e_offset; // S * scale = 2 ** (clip(v,0,1) * e_slope + e_offset)  (host: the dB chain times log2(10), + log2(scale))
    float* mag;               // [frames][kMagPitch]
    float scale;              // power-of-two pre-scale of the magnitudes (undone on output; Griffin-Lim is linear in S)
    int total_frames;
    int* status;
};

__device__ __forceinline__ int slot_index(int bin) {
    // position of `bin` inside a permuted magnitude row (inverse of bin_of); bin 1024 -> 1024
    if (bin == 1024) return 1024;
    int q = bin & 63, pp = bin >> 6;
    int lane, p;
    if ((bin & 31) == 0) { lane = 0; p = bin >> 5; }
    else if (q < 32) { lane = q; p = pp; }
    else { lane = 64 - q; p = 31 - pp; }
    return ((p >> 2) * 32 + lane) * 4 + (p & 3);
}

// S = (10 ** ((clip(v,0,1) * -min + min + ref) * 0.05)) ** power: the exponent is formed in double (one FMA) and split
// into integer + fraction, 2 ** fraction is a float exp2 (|fraction| <= 0.5, ~1e-7 relative), the integer part an exact
// ldexp.  (A double exp10 per element made this kernel 5 % of a Griffin-Lim call.)
__device__ __forceinline__ float prep_magnitude(const PrepParams& P, float v) {
    if (!P.denorm) return fabsf(v) * P.scale;
    const double c = fmin(fmax((double)v, 0.0), 1.0);
    const double e = fma(c, P.e_slope, P.e_offset);
    const double n = rint(e);
    return ldexpf(exp2f((float)(e - n)), (int)n);
}

// frame-major input: one warp per frame, row staged in shared memory in slot order, written back as float4
__global__ void __launch_bounds__(kThreads) k_prepare_mag_rows(PrepParams P) {
    __align__(16) __shared__ float row[kWarpsPerCta][kMagPitch];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool bad = false;
    for (int i = 1025 + lane; i < kMagPitch; i += 32) row[warp][i] = 0.f;      // the pad of the row never changes
    for (int f = P.batch.frame_base + blockIdx.x * kWarpsPerCta + warp; f < P.batch.frame_base + P.total_frames; f += gridDim.x * kWarpsPerCta) {
        const float* in = P.in + (size_t)f * kBins;            // frame-major blocks are packed: global frame index addresses the row
        float v[33];
#pragma unroll
        for (int q = 0; q < 33; ++q) { const int kb = lane + 32 * q; v[q] = kb < kBins ? __ldg(in + kb) : 0.f; }   // all loads in flight at once
#pragma unroll
        for (int q = 0; q < 33; ++q) {
            const int kb = lane + 32 * q;
            if (kb < kBins) {
                bad |= !isfinite(v[q]);
                row[warp][slot_index(kb)] = prep_magnitude(P, v[q]);
            }
        }
        __syncwarp();
        float4* out = reinterpret_cast<float4*>(P.mag + (size_t)f * kMagPitch);
        const float4* r4 = reinterpret_cast<const float4*>(row[warp]);
        for (int i = lane; i < kMagPitch / 4; i += 32) out[i] = r4[i];
        __syncwarp();
    }
    if (bad) atomicOr(P.status, 1);
}

__global__ void __launch_bounds__(256) k_prepare_mag(PrepParams P) {
    // one block per (utterance-local) tile of 32 frames x 32 bins, through a padded smem tile so that both the
    // read (contiguous along the input's fast axis) and the per-frame scatter stay inside a few cache lines
    __shared__ float tile[32][33];
    const int f0 = P.batch.frame_base + blockIdx.x * 32;   // global frame index base
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    bool bad = false;
    for (int kb0 = 0; kb0 < kBins; kb0 += 32) {
        // load: tile[frame][bin]
        for (int r = ty; r < 32; r += 8) {
            int f, kb;
            if (P.bin_major) { f = f0 + tx; kb = kb0 + r; } else { f = f0 + r; kb = kb0 + tx; }
            float v = 0.f;
            if (f < P.batch.frame_base + P.total_frames && kb < kBins) {
                int b = find_segment(P.batch.frame_off, P.batch.batch, f);
                int fo = __ldg(P.batch.frame_off + b);
                int T = __ldg(P.batch.frame_off + b + 1) - fo;
                size_t base = (size_t)fo * kBins;
                v = P.bin_major ? __ldg(P.in + base + (size_t)kb * T + (f - fo)) : __ldg(P.in + base + (size_t)(f - fo) * kBins + kb);
            }
            if (P.bin_major) tile[tx][r] = v; else tile[r][tx] = v;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            int f = f0 + r, kb = kb0 + tx;
            if (f < P.batch.frame_base + P.total_frames && kb < kBins) {
                float v = tile[r][tx];
                bad |= !isfinite(v);
                P.mag[(size_t)f * kMagPitch + slot_index(kb)] = prep_magnitude(P, v);
            }
        }
        __syncthreads();
    }
    if (bad) atomicOr(P.status, 1);
}

// =============================================================================================
// (de-)emphasis
// =============================================================================================
struct EmphParams {
    Batch batch;
    const float* in;      // packed float32
    double* out64;        // one of the two outputs is non-null
    float* out32;
    double p;
    double scale;         // output multiplier (undoes the magnitude pre-scale); 0 is treated as 1
    int* status;          // optional device flag: bit0 = non-finite sample (the Griffin-Lim result passes through here)
    int n_seg;            // CTAs per utterance (blockIdx.x = utterance * n_seg + segment), >= 1
    long long seg_len;    // samples per CTA (a multiple of the chunk); 0: one CTA per utterance
    long long warmup;     // samples a segment's CTA runs through before its first output: p^warmup is below double rounding, so
                          // the dropped history cannot be seen in the result (host: smallest chunk multiple with p^w < 1e-24)
};

// y[n] = x[n] + p*y[n-1] per utterance (scipy.signal.lfilter([1],[1,-p]), zero initial state), in double.
// An utterance is cut into segments of seg_len samples, one CTA each: the filter's memory decays like p^n, so
// a CTA that starts `warmup` samples early from a zero state reproduces the sequential result to below double rounding.
// One CTA walks its segment chunk by chunk; inside a chunk each thread owns kPerThread consecutive
// samples and the carries are combined with a warp/CTA scan of the constant-ratio recurrence.
constexpr int kDeemphThreads = 256;
constexpr int kPerThread = 8;

__global__ void __launch_bounds__(kDeemphThreads) k_deemphasis(EmphParams P) {
    __shared__ double warp_tot[kDeemphThreads / 32];
    __shared__ double carry_s;
    const int nseg = P.n_seg > 0 ? P.n_seg : 1;
    const int b = blockIdx.x / nseg, seg = blockIdx.x - b * nseg;
    const long long s_off = __ldg(P.batch.samp_off + b);
    const long long L = __ldg(P.batch.samp_off + b + 1) - s_off;
    const float* x = P.in + s_off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double p = P.p;
    const double oscale = P.scale == 0.0 ? 1.0 : P.scale;
    double ppow[kPerThread + 1];             // p^1 .. p^kPerThread
    ppow[0] = 1.0;
#pragma unroll
    for (int i = 1; i <= kPerThread; ++i) ppow[i] = ppow[i - 1] * p;
    const double pT = ppow[kPerThread];      // decay across one thread's span
    // pT^(2^d) for the warp scan, and pT^32 for the cross-warp step
    double pw[6];
    pw[0] = pT;
#pragma unroll
    for (int d = 1; d < 6; ++d) pw[d] = pw[d - 1] * pw[d - 1];
    if (tid == 0) carry_s = 0.0;
    __syncthreads();
    bool bad = false;
    const long long chunk = (long long)kDeemphThreads * kPerThread;
    const long long seg0 = P.seg_len > 0 ? (long long)seg * P.seg_len : 0;      // first sample this CTA writes
    const long long seg1 = P.seg_len > 0 ? min(L, seg0 + P.seg_len) : L;
    if (seg0 >= L) return;
    for (long long c0 = max(0LL, seg0 - P.warmup); c0 < seg1; c0 += chunk) {
        const long long i0 = c0 + (long long)tid * kPerThread;
        double loc[kPerThread];
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            const float xv = (i0 + i < L) ? __ldg(x + i0 + i) : 0.f;
            bad |= !isfinite(xv);
            double v = (double)xv * oscale;
            s = fma(p, s, v);
            loc[i] = s;
        }
        // inclusive scan of the thread totals: S_i = e_i + pT * S_{i-1}
        double S = s;
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            double o = __shfl_up_sync(0xffffffffu, S, 1 << d);
            if (lane >= (1 << d)) S = fma(pw[d], o, S);
        }
        if (lane == 31) warp_tot[warp] = S;
        __syncthreads();
        // state entering this warp = combination of previous warps' totals and the chunk carry
        double enter = carry_s;
        for (int w = 0; w < warp; ++w) enter = fma(pw[5], enter, warp_tot[w]);
        // state entering this thread = (exclusive) scan value within the warp + decayed warp-entry state
        double prev = __shfl_up_sync(0xffffffffu, S, 1);
        if (lane == 0) prev = 0.0;
        // enter decays by pT per preceding lane
        double dec = 1.0;
        {
            int l = lane;
#pragma unroll
            for (int d = 0; d < 5; ++d) { if (l & 1) dec *= pw[d]; l >>= 1; }
        }
        const double cin = fma(dec, enter, prev);
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
            if (i0 + i < seg1 && i0 + i >= seg0) {
                double v = fma(ppow[i + 1], cin, loc[i]);
                if (P.out64) P.out64[s_off + i0 + i] = v; else P.out32[s_off + i0 + i] = (float)v;
            }
        }
        __syncthreads();
        if (tid == kDeemphThreads - 1) carry_s = fma(ppow[kPerThread], cin, loc[kPerThread - 1]);
        __syncthreads();
    }
    if (bad && P.status) atomicOr(P.status, 1);
}

// y[n] = x[n] - p*x[n-1] (lfilter([1,-p],[1])), double arithmetic like scipy
__global__ void __launch_bounds__(256) k_preemphasis(EmphParams P, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // utterance start lookup by binary search on the 64-bit offsets
        int lo = 0, hi = P.batch.batch;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (__ldg(P.batch.samp_off + mid) <= i) lo = mid; else hi = mid; }
        const long long s_off = __ldg(P.batch.samp_off + lo);
        double v = (double)__ldg(P.in + i);
        if (i > s_off) v -= P.p * (double)__ldg(P.in + i - 1);
        if (P.out64) P.out64[i] = v; else P.out32[i] = (float)v;
    }
}

// =============================================================================================
// element-wise helpers and stand-alone mel projection
// =============================================================================================
enum { EW_AMP_TO_DB = 0, EW_DB_TO_AMP = 1, EW_NORMALIZE = 2, EW_DENORMALIZE = 3 };

__global__ void __launch_bounds__(256) k_elementwise(int op, const float* __restrict__ in, float* __restrict__ out,
                                                     long long n, float min_level_db) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float x = __ldg(in + i), y;
        switch (op) {
            case EW_AMP_TO_DB: y = 20.0f * log10f(fmaxf(1e-5f, x)); break;
            case EW_DB_TO_AMP: y = (float)pow(10.0, (double)x * 0.05); break;
            case EW_NORMALIZE: y = fminf(fmaxf((x - min_level_db) / (-min_level_db), 0.f), 1.f); break;
            default: y = fminf(fmaxf(x, 0.f), 1.f) * (-min_level_db) + min_level_db; break;
        }
        out[i] = y;
    }
}

struct MelParams {
    Plan plan;
    Batch batch;
    const float* in;     // linear magnitudes, packed per utterance, frame-major or bin-major
    int bin_major;
    double* out64;       // [frames][num_mels] frame-major (np.dot(basis f64, S) is float64)
    float* out32;
    int total_frames;
};

__global__ void __launch_bounds__(kThreads) k_linear_to_mel(MelParams P) {
    __shared__ float row[kWarpsPerCta][kBins + 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int f = P.batch.frame_base + blockIdx.x * kWarpsPerCta + warp; f < P.batch.frame_base + P.total_frames; f += gridDim.x * kWarpsPerCta) {
        const int b = find_segment(P.batch.frame_off, P.batch.batch, f);
        const int fo = __ldg(P.batch.frame_off + b);
        const int T = __ldg(P.batch.frame_off + b + 1) - fo;
        const float* base = P.in + (size_t)fo * kBins;
        for (int kb = lane; kb < kBins; kb += 32)
            row[warp][kb] = P.bin_major ? __ldg(base + (size_t)kb * T + (f - fo)) : __ldg(base + (size_t)(f - fo) * kBins + kb);
        __syncwarp();
        for (int m = lane; m < P.plan.num_mels; m += 32) {
            const int lo = __ldg(P.plan.mel_lo + m), n = __ldg(P.plan.mel_n + m);
            const float* w = P.plan.mel_w + __ldg(P.plan.mel_ptr + m);
            double acc = 0.0;
            for (int i = 0; i < n; ++i) acc = fma((double)__ldg(w + i), (double)row[warp][lo + i], acc);
            if (P.out64) P.out64[(size_t)f * P.plan.num_mels + m] = acc; else P.out32[(size_t)f * P.plan.num_mels + m] = (float)acc;
        }
        __syncwarp();
    }
}

// =============================================================================================
// find_endpoint (utils/audio.py:67-74): the first x = hop, 2*hop, ... < len - window with max(wav[x : x + window]) < threshold
// gives x + hop, else len.  window = int(sample_rate * min_silence_sec), hop = int(window / 4); note np.max, not max |.|.
// One CTA per utterance: maxima of the hop-long blocks (shared memory), then every candidate window is the maximum of its
// window/hop whole blocks and the window % hop samples after them; the smallest hit wins (64-bit atomicMin in shared memory).
// =============================================================================================
struct EndpointParams {
    Batch batch;
    const float* in32;        // packed samples: one of the two inputs is non-null
    const double* in64;
    long long* out;           // [batch] endpoints (sample counts)
    long long window, hop;
    double threshold;         // amplitude (the caller applies _db_to_amp)
};
constexpr int kEndpointBlocks = 4096;     // hop-long blocks held in shared memory per pass (4096 * 4000 samples = 13 min at 20 kHz)

__global__ void __launch_bounds__(256) k_find_endpoint(EndpointParams P) {
    __shared__ double bmax[kEndpointBlocks];
    __shared__ unsigned long long best;
    const int b = blockIdx.x;
    const long long s_off = __ldg(P.batch.samp_off + b);
    const long long L = __ldg(P.batch.samp_off + b + 1) - s_off;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto at = [&](long long i) -> double { return P.in64 ? __ldg(P.in64 + s_off + i) : (double)__ldg(P.in32 + s_off + i); };
    if (threadIdx.x == 0) best = (unsigned long long)L;
    const long long nb = (L + P.hop - 1) / P.hop;                 // blocks of the utterance
    const long long q = P.window / P.hop;                         // whole blocks per window (the rest is read directly)
    // candidates x = j * hop, j >= 1, x < L - window; processed in passes of kEndpointBlocks - q blocks
    for (long long j0 = 1; j0 * P.hop < L - P.window; j0 += kEndpointBlocks - q) {
        __syncthreads();
        if (best < (unsigned long long)L) break;                  // an earlier pass found the endpoint (uniform: read after the barrier)
        const long long nblk = min((long long)kEndpointBlocks, nb - j0);
        for (long long k = warp; k < nblk; k += 8) {              // one warp per block
            const long long a0 = (j0 + k) * P.hop, a1 = min(L, a0 + P.hop);
            double m = -INFINITY;
            bool nan = false;
            for (long long i = a0 + lane; i < a1; i += 32) { const double v = at(i); nan |= (v != v); m = fmax(m, v); }
            for (int d = 16; d > 0; d >>= 1) { m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d)); nan |= (__shfl_xor_sync(0xffffffffu, (int)nan, d) != 0); }
            if (lane == 0) bmax[k] = nan ? NAN : m;               // np.max propagates NaN: such a window never compares below
        }
        __syncthreads();
        for (long long k = threadIdx.x; k + q <= nblk; k += blockDim.x) {
            const long long x = (j0 + k) * P.hop;
            if (x >= L - P.window) break;
            double m = -INFINITY;
            bool nan = false;
            for (long long t = 0; t < q; ++t) { const double v = bmax[k + t]; nan |= (v != v); m = fmax(m, v); }
            for (long long i = x + q * P.hop; i < x + P.window; ++i) { const double v = at(i); nan |= (v != v); m = fmax(m, v); }
            if (!nan && m < P.threshold) atomicMin(&best, (unsigned long long)(x + P.hop));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) P.out[b] = (long long)best;
}


// =============================================================================================
// save_wav's scaling (utils/audio.py:17-19): wav *= 32767 / max(0.01, np.max(np.abs(wav))), per utterance, on the first
// limit[b] samples (the caller trims to find_endpoint before saving, synthesizer.py:53 -> eval.py:43); float64 like numpy.
// k_peak_max: per-utterance max |x| (the bit pattern of a non-negative double orders like an unsigned integer: atomicMax);
// k_peak_apply: x * factor -> float64, or the C cast to int16 (numpy's astype(np.int16): truncation toward zero).
// =============================================================================================
struct PeakParams {
    Batch batch;
    const double* in64;       // one of the two inputs is non-null
    const float* in32;
    const long long* limit;   // [batch] samples that count (and are written), null: the whole utterance
    unsigned long long* peak; // [batch] bit pattern of max |x| (zeroed by the host)
    double* out64;            // one of the two outputs is non-null; samples at and beyond limit[b] become 0
    short* out16;
    int n_seg;                // CTAs per utterance
    int* status;
};

__global__ void __launch_bounds__(256) k_peak_max(PeakParams P) {
    __shared__ double wmax[8];
    const int b = blockIdx.x / P.n_seg, seg = blockIdx.x - b * P.n_seg;
    const long long s_off = __ldg(P.batch.samp_off + b);
    long long L = __ldg(P.batch.samp_off + b + 1) - s_off;
    if (P.limit) L = min(L, max(0LL, __ldg(P.limit + b)));
    const long long per = (L + P.n_seg - 1) / P.n_seg;
    const long long i0 = seg * per, i1 = min(L, i0 + per);
    double m = 0.0;
    bool nan = false;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const double v = P.in64 ? __ldg(P.in64 + s_off + i) : (double)__ldg(P.in32 + s_off + i);
        nan |= (v != v);
        m = fmax(m, fabs(v));
    }
    for (int d = 16; d > 0; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
    if (nan && P.status) atomicOr(P.status, 1);         // np.max would propagate the NaN into every sample
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmax(m, wmax[w]);
        atomicMax(P.peak + b, (unsigned long long)__double_as_longlong(m));
    }
}

// samples [i_begin, i_end) of the packed buffers (a whole batch or one chunk of it)
__global__ void __launch_bounds__(256) k_peak_apply(PeakParams P, long long i_begin, long long i_end) {
    for (long long i = i_begin + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < i_end; i += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = P.batch.batch;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (__ldg(P.batch.samp_off + mid) <= i) lo = mid; else hi = mid; }
        const long long s_off = __ldg(P.batch.samp_off + lo);
        const bool live = !P.limit || (i - s_off) < __ldg(P.limit + lo);
        const double peak = __longlong_as_double((long long)P.peak[lo]);
        const double factor = 32767.0 / fmax(0.01, peak);
        const double x = P.in64 ? __ldg(P.in64 + i) : (double)__ldg(P.in32 + i);
        const double v = live ? x * factor : 0.0;
        if (P.out64) P.out64[i] = v; else P.out16[i] = (short)(int)v;
    }
}

// =============================================================================================
// frame energy: mean(|x|^2) of every centred frame (librosa.feature.rmse(y, frame_length, hop_length, center=True,
// pad_mode='reflect') ** 2), the reduction behind trim_wav (librosa.effects.split) and trim_silence
// (datasets/process.py:39-54).  One warp per frame, double accumulation.  frames per utterance = 1 + n // hop.
// =============================================================================================
struct EnergyParams {
    Batch batch;              // frame_off counts THESE frames
    const float* wav;         // packed samples
    double* out;              // [frames] mean square
    int frame_length, hop_length, total_frames;
};

__global__ void __launch_bounds__(256) k_frame_energy(EnergyParams P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int b = 0;
    for (int f = P.batch.frame_base + blockIdx.x * 8 + warp; f < P.batch.frame_base + P.total_frames; f += gridDim.x * 8) {
        if (!(f >= __ldg(P.batch.frame_off + b) && f < __ldg(P.batch.frame_off + b + 1)))
            b = find_segment(P.batch.frame_off, P.batch.batch, f);
        const int k = f - __ldg(P.batch.frame_off + b);
        const long long s_off = __ldg(P.batch.samp_off + b);
        const int L = (int)(__ldg(P.batch.samp_off + b + 1) - s_off);
        const float* x = P.wav + s_off;
        const int start = k * P.hop_length - P.frame_length / 2;
        double acc = 0.0;
        if (start >= 0 && start + P.frame_length <= L) {
            for (int i = lane; i < P.frame_length; i += 32) { const double v = __ldg(x + start + i); acc = fma(v, v, acc); }
        } else {
            for (int i = lane; i < P.frame_length; i += 32) { const double v = __ldg(x + reflect_index(start + i, L)); acc = fma(v, v, acc); }
        }
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) P.out[f] = acc / (double)P.frame_length;
    }
}

}  // namespace nsb
