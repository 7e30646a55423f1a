// k_gl_iter: ONE Griffin-Lim iteration (audio.py:85-86: angles = exp(1j*angle(_stft(y))); y = _istft(S*angles))
// as a single persistent kernel.  Per frame and per launch only the magnitude row (4.1 KB) and the waveform
// hop (1 KB in, 1 KB out) cross HBM; the complex spectrum lives and dies in registers.
//
// Compared with the generic k_synth<SRC_Y> (kept for the other sources) this kernel is specialised for the
// hot loop:
//   * persistent CTAs (grid = 2 x #SMs) walk the tile list, tables are staged in shared memory once;
//   * the frame's magnitude row is prefetched with cp.async into the warp's own scratch tile while the
//     second FFT pass runs (the scratch tile is idle exactly then), so the renormalisation never waits on HBM/L2;
//   * overlap-add ordering uses per-warp progress flags between NEIGHBOUR warps instead of one CTA barrier
//     per colour: warp w owns the C consecutive frames kb + C*w .. kb + C*w + C-1, adds them in the order of the
//     global colour k mod C, and may add its colour-s frame once both neighbours have added their colours < s
//     (frames of non-neighbour warps never overlap; equal colours never overlap);
//   * the lane-0 real-64 split (rows 0/32 of the 64x32 decomposition) is spread over lanes 1..16 through a
//     256-byte exchange area instead of running on one lane while 31 idle;
//   * branch-free renormalisation (the magnitudes are pre-scaled by a power of two in k_prepare_mag so the
//     squares cannot overflow; exact zeros are fixed up in a rare warp-uniform path);
//   * DEFCFG folds hop/win/lo of the reference's default hparams (250/1000/524; lo = 0 for the TF twin) into immediates.
#pragma once
#include "kernels.cuh"

namespace nsb {

#ifdef NSB_EMULATE
__device__ __forceinline__ void cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
__device__ __forceinline__ void cp_async8(void* dst, const void* src) { memcpy(dst, src, 8); }
__device__ __forceinline__ void prefetch_l2(const void*) {}
__device__ __forceinline__ void cp_async_wait_all() {}
__device__ __forceinline__ int flag_load(const int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void flag_store(int* p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ void spin_pause() { std::this_thread::yield(); }
__device__ __forceinline__ int gflag_load(const int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void gflag_store(int* p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ bool warp_any(bool p) {
    auto* c = nsb_emu::g_ctx;
    const int w = threadIdx.x >> 5;
    c->xchg[w][threadIdx.x & 31] = p ? 1.0 : 0.0;
    __syncwarp();
    bool r = false;
    for (int l = 0; l < 32; ++l) r |= c->xchg[w][l] > 0.5;
    __syncwarp();
    return r;
}
#else
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ int flag_load(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void flag_store(int* p, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void spin_pause() { __nanosleep(32); }
__device__ __forceinline__ int gflag_load(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void gflag_store(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ bool warp_any(bool p) { return __any_sync(0xffffffffu, p); }
#endif

// 1/sqrt(x) as ONE MUFU.RSQ: plain rsqrtf() wraps the instruction in a subnormal-input rescue (FSETP + two predicated
// FMULs per call, 96 instructions per frame); the callers below never pass a subnormal that matters
__device__ __forceinline__ float rsqrt_ftz(float x) {
#ifdef NSB_EMULATE
    return 1.0f / sqrtf(x);
#else
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

struct GlParams {
    Plan plan;
    Batch batch;
    const float* y_in;
    float* y_out;
    const float* mag;        // permuted, pre-scaled magnitudes [frames][kMagPitch]
    int tile_hops, colours, total_tiles;
    int wide;                // 1: small batches - frame j of a tile goes to warp j mod 8 (one frame per warp and step instead of C
                             // consecutive ones): a tile of C hops is 1 frame-time long instead of C, and there are H/C times more tiles
                             // to spread over the SMs.  The overlap-add is then ordered by C CTA barriers per step (colour = j mod C).
    // One launch runs `iters` Griffin-Lim iterations: the work items (iteration n, tile t), n-major, are handed out by a
    // global counter; item (n, t) may start once the tiles t-1, t, t+1 of iteration n-1 are stored (per-tile counters
    // `done`, release/acquire at GPU scope).  y ping-pongs between ybuf[0] and ybuf[1]: iteration n reads
    // ybuf[(cur0 + n) & 1] and writes the other one (y_in / y_out above are unused by this kernel).
    float* ybuf[2];
    int cur0, iters;
    int* item_counter;       // [1], zeroed by the host before the launch
    int* done;               // [total_tiles], zeroed by the host before the launch: iterations finished by each tile
    float inv_thr;           // TF twin: the clamp of est / max(1e-8, |est|) (audio.py:101) for register slots, which hold
                             // 2 * g * est (g = magnitude pre-scale): 1 / (2 * g * 1e-8)
    int* status;
};

// TF twin: angles = est / max(1e-8, |est|) -> z * S * min(1/|z|, 1/thr); |z| = 0 gives 0 (not S)
__device__ __forceinline__ void renorm_tf(c2& z, float S, float inv_thr) {
    float m2 = fmaf(z.x, z.x, z.y * z.y);
    z = cscale(z, fminf(rsqrt_ftz(m2), inv_thr) * S);
}

constexpr int kMagBytes = 4096 + 16;                 // slots + the float4 that carries the Nyquist magnitude
constexpr int kXchOffsetF2 = (kMagBytes + 112) / 8;  // exchange area starts at byte 4224 of the scratch tile (8-byte units)

// renormalise one slot, branch-free; `zero` collects exact zeros for the rare fix-up
// (the 1e-36 rides in the second FMA instead of costing an FMNMX per slot: |z|^2 + 1e-36 never reaches rsqrt as 0)
__device__ __forceinline__ void renorm_fast(c2& z, float S, bool& zero) {
    float m2 = fmaf(z.x, z.x, fmaf(z.y, z.y, 1e-36f));
    zero |= (m2 <= 1e-36f);
    z = cscale(z, rsqrt_ftz(m2) * S);
}

// normalise the finished tile by the summed squared window and store it (all threads of the CTA)
// rinv_s[j] already contains 1 / (n_fft * window-sum) for the interior (all covering frames exist).
template <bool DEFCFG>
__device__ __noinline__ void gl_store_tile(const GlParams& P, float* y_out, int tile_g, const float* acc, const float* win_s, const float* rinv_s,
                                           int hop, int win, int lo, int H, bool& bad) {
    const int a = P.plan.origin - lo;
    const int b = __ldg(P.batch.tile_utt + tile_g) - P.batch.utt_base;
    const int tile = tile_g - __ldg(P.batch.tile_off + b);
    const int T = __ldg(P.batch.frame_off + b + 1) - __ldg(P.batch.frame_off + b);
    const long long s_off = __ldg(P.batch.samp_off + b);
    const long long L = __ldg(P.batch.samp_off + b + 1) - s_off;
    const int n_hops = (int)((L + hop - 1) / hop);
    const int h0 = tile * H, h1 = min(h0 + H, n_hops);
    const long long o0 = s_off + (long long)h0 * hop;
    float* yo = y_out + o0;
    const int n_out = (int)(min((long long)h1 * hop, L) - (long long)h0 * hop);
    // interior tile: every sample is covered by all of its ceil(win/hop) frames -> no per-sample frame tests
    // (without window-sum normalisation every tile is "interior": rinv_s is the constant 1/n_fft)
    const int ncov = (win + hop - 1) / hop;
    const bool interior = !P.plan.norm_wss || ((h0 + a / hop - (ncov - 1) >= 0) && (h1 - 1 + (hop - 1 + a) / hop <= T - 1));
    float chk = 0.f;                                  // NaN/Inf detector: v*0 accumulates to NaN iff some v is not finite
    if (interior && (hop & 1) == 0 && (o0 & 1) == 0) {
        const float2* acc2 = reinterpret_cast<const float2*>(acc);
        const float2* r2 = reinterpret_cast<const float2*>(rinv_s);
        float2* yo2 = reinterpret_cast<float2*>(yo);
        const int hop2 = hop >> 1;
        c2 chk2 = mk2(0.f, 0.f);
        for (int i = threadIdx.x; i < (n_out >> 1); i += kThreads) {
            const int j = i % hop2;
            c2 v = p_mul(acc2[i], r2[j]);
            chk2 = p_fma(v, mk2(0.f, 0.f), chk2);
            yo2[i] = v;
        }
        chk = chk2.x + chk2.y;
    } else {
        for (int j = threadIdx.x; j < hop; j += kThreads) {
            const int dj = (j + a) / hop, rj = (j + a) - dj * hop;
            const float ri = rinv_s[j];
            int ncover = 0;
            for (int idx = rj; idx < win; idx += hop) ++ncover;
            for (int h = h0; h < h1; ++h) {
                if ((h - h0) * hop + j >= n_out) break;
                const int k_hi = h + dj, k_lo = k_hi - (ncover - 1);
                float v = acc[(h - h0) * hop + j];
                if (!P.plan.norm_wss || (k_lo >= 0 && k_hi <= T - 1)) {
                    v *= ri;
                } else {
                    v *= (1.0f / (float)kNfft);
                    float sm = 0.f;
                    int kk = k_hi;
                    for (int idx = rj; idx < win; idx += hop, --kk)
                        if (kk >= 0 && kk <= T - 1) { float w = win_s[lo + idx]; sm = fmaf(w, w, sm); }
                    if (sm > 1.17549435e-38f) v /= sm;
                }
                chk = fmaf(v, 0.f, chk);
                yo[(size_t)(h - h0) * hop + j] = v;
            }
        }
    }
    bad |= (chk != 0.f);
}

// Rejected variants (measured on B200, batch 64 x 1000 frames, see profiles/README.md): one shared FFT32 copy for
// the four passes through a rolled pass loop (i-cache stalls 23% -> 6% but +12% instructions from loop-carried
// register shuffling: no gain); software-pipelined tile hand-over (4% slower); global-colour order that rotates
// the warp's frames (12% slower than increasing order, hence the aligned tiles below).
template <int PRUNE, bool DEFCFG, bool TFM>
__global__ void __launch_bounds__(kThreads, 2) k_gl_iter(GlParams P) {
    NSB_DYN_SMEM(smem_raw);
    const int hop = DEFCFG ? 250 : P.plan.hop;
    const int win = DEFCFG ? 1000 : P.plan.win_len;
    constexpr int LO = TFM ? 0 : 524;                // DEFCFG: librosa pads the window centrally, tf.contrib.signal on the right
    const int lo = DEFCFG ? LO : P.plan.lo;
    const int C = DEFCFG ? 4 : P.colours;
    const int H = P.tile_hops;                       // a multiple of C (host guarantees it)
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float* win_s = reinterpret_cast<float*>(tw_s + kTwF2);
    float* rinv_s = win_s + kNfft;
    float* acc = rinv_s + ((hop + 3) & ~3);          // 16-byte aligned (float4 zeroing, float2 epilogue)
    size_t acc_end = (size_t)(acc + (size_t)H * hop - reinterpret_cast<float*>(smem_raw));
    acc_end = (acc_end + 3) & ~(size_t)3;
    int* progress = reinterpret_cast<int*>(reinterpret_cast<float*>(smem_raw) + acc_end);   // [kWarpsPerCta] (+pad to 16 ints)
    float2* scratch_all = reinterpret_cast<float2*>(progress + 16);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* scratch = scratch_all + warp * kScratchF2;
    const int origin = DEFCFG ? (TFM ? 0 : kNfft / 2) : P.plan.origin;
    const int a = origin - lo;
    // first frame (possibly negative = does not exist) whose window support can reach hop h is h + kfirst0
    const int kfirst0 = (a - win >= 0) ? (a - win) / hop + 1 : -((win - a - 1) / hop + 1) + 1;

    load_twiddle_pairs(tw_s, P.plan.tw);
    const float4* tw4 = reinterpret_cast<const float4*>(tw_s);
    const float2* tw31 = tw_s + 15 * 64;
    for (int i = threadIdx.x; i < kNfft; i += kThreads) win_s[i] = P.plan.win[i];
    __syncthreads();
    for (int j = threadIdx.x; j < hop; j += kThreads) {
        float s = 0.f;
        for (int idx = (j + a) % hop; idx < win; idx += hop) { float w = win_s[lo + idx]; s = fmaf(w, w, s); }
        rinv_s[j] = ((P.plan.norm_wss && s > 1.17549435e-38f) ? 1.0f / s : 1.0f) * (1.0f / (float)kNfft);
    }
    bool bad = false;

    // Dynamic scheduling of (iteration, tile) items from one global counter.  SMs do not all run at the same speed and
    // the older of an SM's two CTAs gets the issue slots first (per-CTA timelines: profiles/r1/trace_*.txt): a static
    // round-robin left SMs idle at the end of EVERY iteration's launch.  Here a CTA that runs out of tiles of iteration
    // n simply starts on iteration n+1 - only the last iteration has a tail, and there is no launch gap and no table
    // reload between iterations.
    const long long total_items = (long long)P.iters * P.total_tiles;
    int item_it = 0;
    if (threadIdx.x == 0) progress[8] = atomicAdd(P.item_counter, 1);
    __syncthreads();
    for (long long item = progress[8]; item < total_items; item = progress[8 + (++item_it & 1)]) {
        const int n_it = (int)(item / P.total_tiles);
        const int tile_l = (int)(item - (long long)n_it * P.total_tiles);
        const int tile_g = P.batch.tile_base + tile_l;
        const float* y_in = P.ybuf[(P.cur0 + n_it) & 1];
        float* y_out = P.ybuf[(P.cur0 + n_it + 1) & 1];
        // four lanes of warp 0, four round trips to L2 side by side: the next item, and the three tiles whose iteration n-1 this
        // tile reads (t-1, t, t+1; my store also overwrites what THEY read in iteration n-1 - one wait covers both hazards)
        if (threadIdx.x == 0) progress[8 + ((item_it + 1) & 1)] = atomicAdd(P.item_counter, 1);     // read after this tile's barriers
        if (n_it > 0 && threadIdx.x >= 1 && threadIdx.x <= 3) {
            int t = tile_l + (int)threadIdx.x - 2;
            t = t < 0 ? 0 : (t >= P.total_tiles ? P.total_tiles - 1 : t);
            while (gflag_load(P.done + t) < n_it) spin_pause();
        }
        const int b = __ldg(P.batch.tile_utt + tile_g) - P.batch.utt_base;
        const int tile = tile_g - __ldg(P.batch.tile_off + b);
        const int f_off = __ldg(P.batch.frame_off + b);
        const int T = __ldg(P.batch.frame_off + b + 1) - f_off;
        const long long s_off = __ldg(P.batch.samp_off + b);
        const long long L = __ldg(P.batch.samp_off + b + 1) - s_off;   // hop*(T-1) (librosa) or hop*(T-1)+win (tf)
        const int n_hops = (int)((L + hop - 1) / hop);
        const int h0 = tile * H, h1 = min(h0 + H, n_hops);
        const int s0 = h0 * hop;
        const int n_out = (int)(min((long long)h1 * hop, L) - s0);
        // Frames are grouped from the VIRTUAL first frame h0 + kfirst0 (negative frames simply do not exist).  Because
        // H is a multiple of C, every tile's groups start at the same residue mod C: the warp adds its frames in
        // increasing order AND that order is the global colour (k - kfirst0) mod C, so the summation order of every
        // output sample is independent of the tiling (a batch is bit-identical to one-at-a-time calls).
        const int k_first = h0 + kfirst0;
        int k_max = (s0 + n_out + a + hop - 1) / hop - 1;      // last frame with k*hop - a < s1
        if (k_max > T - 1) k_max = T - 1;
        if (!P.wide && k_max - k_first + 1 > kWarpsPerCta * C) bad = true;   // host sizes tiles so this cannot happen; never drop frames silently
        const int kg = k_first + C * warp;
        {
            float4* a4 = reinterpret_cast<float4*>(acc);            // the tile buffer is 16-byte aligned and a multiple of 4 long (+pad)
            for (int i = threadIdx.x; i < (n_out + 3) / 4; i += kThreads) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (threadIdx.x < kWarpsPerCta) progress[threadIdx.x] = 0;
        __syncthreads();

        bool staged = false;                         // this frame's samples were cp.async'ed into the scratch tile already
        int stage_off = 0;                           // ... shifted by this many floats (16-byte alignment of the copies)
        const bool wide = P.wide != 0;
        const int n_steps = wide ? max(0, k_max - k_first + kWarpsPerCta) / kWarpsPerCta : C;
        for (int s = 0; s < n_steps; ++s) {
            const int k = wide ? k_first + warp + kWarpsPerCta * s : kg + s;
            const bool active = (k >= 0 && k <= k_max);        // warp-uniform
            c2 z[32];
            if (active) {
                const int fg = f_off + k;
                // pull the NEXT frame's magnitude row towards L2 (it streams from HBM) while this frame computes
                if (!wide && k + 1 <= k_max) {
                    const char* nm = reinterpret_cast<const char*>(P.mag + (size_t)(fg + 1) * kMagPitch);
                    prefetch_l2(nm + lane * 128);
                    if (lane == 0) prefetch_l2(nm + 4096);
                }
                if (staged) {
                    // samples were staged by the previous frame of this warp (see below): window them from shared memory
                    cp_async_wait_all();
                    __syncwarp();
                    const float* stage = reinterpret_cast<const float*>(scratch) + stage_off;
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        if (t >= 8 && t < 24) z[t] = p_mul(mk2(stage[64 * t + lane], stage[64 * t + 32 + lane]), mk2(win_s[64 * t + lane], win_s[64 * t + 32 + lane]));
                        else z[t] = mk2(0.f, 0.f);
                    }
                    __syncwarp();
                } else {
                    load_frame<false, PRUNE, true>(z, y_in + s_off, L, (long long)k * hop - origin, win_s, lane, 0.f,
                                                   reinterpret_cast<float*>(scratch));
                }
                fwd_phase1_tw4<PruneRange<PRUNE>::t0, PruneRange<PRUNE>::t1>(z, lane, scratch, tw4, tw31);
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 32; ++t) z[t] = scratch[lane * kRowStride + t];
                __syncwarp();                        // every lane has its row: the scratch tile is free
                {
                    const char* src = reinterpret_cast<const char*>(P.mag + (size_t)fg * kMagPitch);
                    char* dst = reinterpret_cast<char*>(scratch);
#pragma unroll
                    for (int g = 0; g < 8; ++g) cp_async16(dst + (g * 32 + lane) * 16, src + (g * 32 + lane) * 16);
                    if (lane == 0) cp_async16(dst + 4096, src + 4096);
                }
                fft32<-1>(z);
                float2* xch = scratch + kXchOffsetF2;
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) xch[j] = z[j];
                }
                cp_async_wait_all();
                __syncwarp();
                const float4* mrow = reinterpret_cast<const float4*>(scratch);
                const float* mflt = reinterpret_cast<const float*>(scratch);
                bool zero = false;
                // (a) every lane renormalises its 32 slots (lane 0's registers hold the packed-row FFT, not bins:
                //     replaced below)
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    float4 S = mrow[g * 32 + lane];
                    if (TFM) {
                        renorm_tf(z[4 * g], S.x, P.inv_thr); renorm_tf(z[4 * g + 1], S.y, P.inv_thr);
                        renorm_tf(z[4 * g + 2], S.z, P.inv_thr); renorm_tf(z[4 * g + 3], S.w, P.inv_thr);
                    } else {
                        renorm_fast(z[4 * g], S.x, zero);
                        renorm_fast(z[4 * g + 1], S.y, zero);
                        renorm_fast(z[4 * g + 2], S.z, zero);
                        renorm_fast(z[4 * g + 3], S.w, zero);
                    }
                }
                if (lane == 0) zero = false;
                if (!TFM && warp_any(zero)) {                // rare: some bin of y's STFT is exactly 0 -> phase 0 (np.angle(0))
                    if (lane != 0) {
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            float4 S = mrow[g * 32 + lane];
                            const float Ss[4] = {S.x, S.y, S.z, S.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (z[4 * g + e].x == 0.f && z[4 * g + e].y == 0.f) z[4 * g + e].x = Ss[e];
                        }
                    }
                }
                // (b) bins k = 32 j (rows 0/32): lane j in 1..16 does pair (j, 32-j); lane 0 the real DC/Nyquist pair
                if (lane <= 16) {
                    if (lane == 0) {
                        float2 g0 = xch[0];
                        float x0 = g0.x + g0.y, xn = g0.x - g0.y;          // (X[0], X[1024]) up to the factor 2
                        float S0 = mflt[0], Sn = mflt[1024];
                        if (TFM) {                                          // x / max(1e-8, |x|); x0, xn carry g * est (no factor 2)
                            x0 = x0 * fminf(1.0f / fabsf(x0), 2.0f * P.inv_thr) * S0;
                            xn = xn * fminf(1.0f / fabsf(xn), 2.0f * P.inv_thr) * Sn;
                            if (!(fabsf(x0) <= S0)) x0 = 0.f;               // 0 * inf
                            if (!(fabsf(xn) <= Sn)) xn = 0.f;
                        } else {
                            x0 = (x0 < 0.f) ? -S0 : S0;                     // phase of a real number is its sign
                            xn = (xn < 0.f) ? -Sn : Sn;
                        }
                        xch[0] = make_float2(x0 + xn, x0 - xn);
                    } else {
                        const int j = lane, jj = 32 - lane;
                        // u = -i * w64^j with w64^j = w2048^(16 * 2j) from the twiddle table (j = 16: w = -i)
                        const float2 wj = (j <= 15) ? tw_s[(7 * 32 + 2 * j) * 2 + 1] : make_float2(0.f, -1.f);   // row 16 of the paired table
                        const c2 u = mk2(wj.y, -wj.x);
                        const c2 Aj = xch[j], Bj = xch[jj];
                        const c2 S1 = cadd_conj(Aj, Bj), D1 = csub_conj(Aj, Bj);
                        const c2 T1 = cmul(D1, u);
                        c2 c1 = cadd(S1, T1);                               // 2*X[32 j]
                        c2 c2v = cconj(csub(S1, T1));                       // 2*X[32 (32-j)]
                        bool z2 = false;
                        if (TFM) {
                            renorm_tf(c1, mflt[(j >> 2) * 128 + (j & 3)], P.inv_thr);
                            renorm_tf(c2v, mflt[(jj >> 2) * 128 + (jj & 3)], P.inv_thr);
                        } else {
                            renorm_fast(c1, mflt[(j >> 2) * 128 + (j & 3)], z2);
                            renorm_fast(c2v, mflt[(jj >> 2) * 128 + (jj & 3)], z2);
                        }
                        if (z2) {
                            if (c1.x == 0.f && c1.y == 0.f) c1.x = mflt[(j >> 2) * 128 + (j & 3)];
                            if (c2v.x == 0.f && c2v.y == 0.f) c2v.x = mflt[(jj >> 2) * 128 + (jj & 3)];
                        }
                        // inverse split: S' = V[j] + conj V[32-j], D' = V[j] - conj V[32-j], P = D' * conj(u)
                        const c2 S2 = cadd_conj(c1, c2v), D2 = csub_conj(c1, c2v);
                        const c2 Pv = cmul_conj(D2, u);
                        // lane j reads and writes only xch[j] and xch[32-j]: no cross-lane hazard inside this block
                        xch[j] = cadd(S2, Pv);
                        if (j != 16) xch[jj] = cconj(csub(S2, Pv));
                    }
                }
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) z[j] = xch[j];
                }
                __syncwarp();                        // magnitude row and exchange area fully consumed
                // inverse pass 1 (the lane-0 pre-split already happened above)
                fft32<+1>(z);
                scratch[lane * kRowStride] = z[0];
#pragma unroll
                for (int p = 0; p < 15; ++p) {             // multiply in place, store afterwards (see fwd_phase1_tw4)
                    const float4 w = tw4[p * 32 + lane];
                    z[2 * p + 1] = cmul_conj(z[2 * p + 1], mk2(w.x, w.y));
                    z[2 * p + 2] = cmul_conj(z[2 * p + 2], mk2(w.z, w.w));
                }
                z[31] = cmul_conj(z[31], tw31[lane]);
#pragma unroll
                for (int r = 1; r < 32; ++r) scratch[lane * kRowStride + r] = z[r];
                __syncwarp();
                inv_phase2(z, lane, scratch);
                __syncwarp();                        // the scratch tile may be rewritten by this warp's next frame
                // The scratch tile now idles through the overlap-add: stage the warp's NEXT frame (k+1) into it with
                // cp.async so that its load latency hides behind the neighbour wait and the accumulate.  Only for
                // frames that need no reflect padding and (8-byte copies) an even sample offset.
                staged = false;
                if (PRUNE == 1 && !wide && s + 1 < C && k + 1 <= k_max) {
                    const long long nstart = (long long)(k + 1) * hop - origin;
                    if (nstart + 512 >= 0 && nstart + 1536 <= L && ((s_off + nstart) & 1) == 0) {
                        // 16-byte L2-only copies (y is rewritten by other SMs inside this launch: nothing of it may sit in L1);
                        // the frame starts on an 8-byte boundary, so copy from the 16-byte boundary at or below it
                        const float* src0 = y_in + s_off + nstart + 512;
                        stage_off = (int)((reinterpret_cast<uintptr_t>(src0) >> 2) & 3);        // 0 or 2
                        const float* src = src0 - stage_off + 4 * lane;
                        float* dst = reinterpret_cast<float*>(scratch) + 512 + 4 * lane;
#pragma unroll
                        for (int j = 0; j < 8; ++j) cp_async16(dst + 128 * j, src + 128 * j);
                        if (lane == 0) cp_async16(dst + 1024, src + 1024);
                        staged = true;
                    }
                }
            } else {
                staged = false;
            }
            // ---- overlap-add ordering ----
            const int colour = (k - k_first) % C;              // wide mode: every warp passes C barriers per step, `colour` of them first
            if (wide) {
                for (int c = 0; c < colour; ++c) __syncthreads();
            } else if (s > 0) {
                // my colour-s frame overlaps only frames of the two neighbour warps; those of colour < s must be in
                if (lane == 0) {
                    if (warp > 0) while (flag_load(progress + warp - 1) < s) spin_pause();
                    if (warp < kWarpsPerCta - 1) while (flag_load(progress + warp + 1) < s) spin_pause();
                }
                __syncwarp();
            }
            if (active) {
                const int base = k * hop - origin - s0;                       // tile-local index of n = 0
                constexpr int t0 = PruneRange<PRUNE>::t0, t1 = PruneRange<PRUNE>::t1;
                const bool inside = (base + lo >= 0) && (base + lo + win <= n_out);
                if (DEFCFG && inside) {
                    // default hparams: the support n in [LO, LO + 1000) is known at compile time; (acc[n], acc[n+32]) and the
                    // two window values ride in register pairs so the accumulate is one FFMA2
                    ola_fixed_support<LO, t0, t1>(acc + base + lane, z, lane,
                                                  [&](int t) { return mk2(win_s[64 * t + lane], win_s[64 * t + 32 + lane]); });
                } else {
                    // window support [lo, lo+win) clipped to the tile -> per-lane bitmasks of the valid t
                    // (n = 64 t + lane [+32]).  Touching nothing outside the support makes the plain RMW race-free.
                    const int nlo = max(lo, -base), nhi = min(lo + win, n_out - base);
                    const int a0 = min(max((nlo - lane + 63) >> 6, 0), 32), a1 = min(max((nhi - lane + 63) >> 6, 0), 32);
                    const int b0 = min(max((nlo - lane - 32 + 63) >> 6, 0), 32), b1 = min(max((nhi - lane - 32 + 63) >> 6, 0), 32);
                    const unsigned mre = (a1 > a0) ? ((0xffffffffu >> (32 - (a1 - a0))) << a0) : 0u;
                    const unsigned mim = (b1 > b0) ? ((0xffffffffu >> (32 - (b1 - b0))) << b0) : 0u;
                    float* ap = acc + base + lane;
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        if (t >= t0 && t < t1) {
                            if ((mre >> t) & 1u) ap[64 * t] = fmaf(z[t].x, win_s[64 * t + lane], ap[64 * t]);
                            if ((mim >> t) & 1u) ap[64 * t + 32] = fmaf(z[t].y, win_s[64 * t + 32 + lane], ap[64 * t + 32]);
                        }
                    }
                }
            }
            if (wide) {
                for (int c = colour; c < C; ++c) __syncthreads();
            } else {
                __syncwarp();
                if (lane == 0) flag_store(progress + warp, s + 1);
            }
        }
        __syncthreads();
        gl_store_tile<DEFCFG>(P, y_out, tile_g, acc, win_s, rinv_s, hop, win, lo, H, bad);
        __syncthreads();                             // acc and progress are reused by the next tile; every thread's stores are issued
        if (threadIdx.x == 0) { __threadfence(); gflag_store(P.done + tile_l, n_it + 1); }
    }
    if (bad) atomicOr(P.status, 1);
}

}  // namespace nsb
