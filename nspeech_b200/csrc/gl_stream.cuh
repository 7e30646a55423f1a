// k_gl_stream: ONE Griffin-Lim iteration (audio.py:85-86: angles = exp(1j*angle(_stft(y))); y = _istft(S*angles))
// as a persistent STREAMING kernel.  Successor of k_gl_iter (gl_iter.cuh), same per-frame arithmetic, different
// walk over the batch:
//
//   * the batch is ONE stream of GROUPS (C = ceil(win/hop) consecutive hops): utterance b contributes its G_b groups
//     plus one END-HALO group (the frames past its last full group, they only reach back), so the stream needs no
//     special case at utterance boundaries.  CTA n owns the contiguous range [V0, V1) = [n*NV/grid, (n+1)*NV/grid) of
//     it and also transforms the group V1 as its own halo (the C-1 frames of it that reach back into V1-1): a range
//     of 54 groups wastes 1.4 % of its FFTs, the old tile kernel recomputed C-1 frames per 8 groups (32 frames
//     for 28 hops, 12.5 %).  Every CTA gets the same number of positions whatever the utterance lengths are;
//   * the range is walked from its RIGHT end: position i = 0 is the halo group V1, i = V1 - V0 the group V0;
//     position i belongs to warp i mod 8 (round i div 8).  A frame of colour s (= its place in the group) overlaps
//     the right-hand group's frames of colour < s and the left-hand group's frames of colour > s, so adding in colour
//     order needs ONE dependency: (group j, colour s) waits for (j+1, < s) - and j+1 is either the neighbour warp in
//     the same round or a group of the previous round.  Walking leftwards keeps every dependency pointing at work
//     that is in step or long done: the warps stay in lockstep (a first version walked rightwards, needed warp 0 a
//     whole round ahead of warp 7, desynchronised the warps and lost 23 % of the issue slots to instruction-cache
//     misses, profiles/r1/ncu_full_k_gl_stream_ascending.txt).  The left-hand neighbour is waited for as well,
//     purely to pace the warps (one colour of drift at most).  Counters: per-warp events in shared memory
//     (st.release / ld.acquire);
//   * the overlap-add goes into a shared-memory RING of 8 groups + the hops a group reaches back.  When a group's
//     last colour is in, its own C hops are final: the SAME warp normalises them by the window sum, stores them to
//     HBM and zeroes them.  There is no CTA barrier inside a piece and no epilogue phase;
//   * ring reuse: the group at position i writes where positions i-8 (the same warp) and i-9 (the right-hand warp,
//     one round ago) stored - (i, colour 0) waits for that store, which is a round old by then;
//   * while a group is stored, the first frame of the warp's next group (8 positions to the left, possibly another
//     utterance) is already on its way into the scratch tile (cp.async) and its magnitude row towards L2.
//
// The summation order of every output sample is the colour order 0..C-1 with colour = (k - k_first0) mod C, a
// property of the frame index alone: pieces, grids and batches never change a bit of the result.
#pragma once
#include "gl_iter.cuh"

namespace nsb {

struct GlStreamParams {
    Plan plan;
    Batch batch;             // group_off / group_base describe the concatenated group space
    // One launch runs `iters` iterations: items (iteration n, chunk c) - chunk c = groups [c*CH, (c+1)*CH) of the stream -
    // come from a global counter, n-major; (n, c) starts when the chunks c-1, c, c+1 of iteration n-1 are stored.
    // Iteration n reads ybuf[(cur0 + n) & 1] and writes the other buffer.
    float* ybuf[2];
    int cur0, iters;
    int chunk_groups;        // CH
    int* item_counter;       // [1], zeroed by the host before the launch
    int* done;               // [chunks], zeroed by the host: iterations finished by each chunk
    const float* mag;        // permuted, pre-scaled magnitudes [frames][kMagPitch]
    int colours;             // C
    int total_groups;        // groups in this (sub-)batch (without the end-halo groups)
    int sync_mode;           // CTA barriers that re-align the warps (instruction-cache locality): 0 none, 1 per round, 2 per frame
    float inv_thr;           // TF twin clamp, see GlParams
    int* status;
    unsigned long long* trace;   // optional [grid][3]: SM id, start and end time (globaltimer ns) of every CTA
};

__device__ __forceinline__ unsigned long long global_ns() {
#ifdef NSB_EMULATE
    return 0;
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#endif
}
__device__ __forceinline__ unsigned sm_id() {
#ifdef NSB_EMULATE
    return 0;
#else
    unsigned v;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
    return v;
#endif
}

// barrier among the 4 warps of one half of the CTA (warps 0-3: barrier 1, warps 4-7: barrier 2); an experiment in how many
// instruction streams per SM the kernel should run (the counters keep the overlap-add ordered whatever the barriers do)
__device__ __forceinline__ void half_cta_barrier(int warp) {
#ifndef NSB_EMULATE
    if (warp < 4) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
#endif
}

__device__ __forceinline__ void wait_events(const int* flag, int target, int lane) {
    if (lane == 0) while (flag_load(flag) < target) spin_pause();
    __syncwarp();
}

// window value of sample n from the paired table wp[(t - t0) * 32 + lane] = (w[64 t + lane], w[64 t + 32 + lane])
__device__ __forceinline__ float win_at(const float* wtab, int t0, int n) {
    return wtab[(((n >> 6) - t0) * 32 + (n & 31)) * 2 + ((n >> 5) & 1)];
}

// normalise the warp's finished group by the summed squared window, store it and clear the ring slot.
// rinv[j] (global memory, Plan::rinv) holds 1 / (n_fft * window-sum) of the interior (all covering frames exist);
// wtab is the paired window table (win_at).
// The group's samples sit at ring[(r0 + i) mod RS].
__device__ __noinline__ void gl_store_group(float* ring, int r0, int RS, float* yo, int n_store, int h_a, int T, const float* wtab, int wt0,
                                            const float* __restrict__ rinv_s, int hop, int win, int lo, int a, int norm_wss, int lane) {
    const int ncov = (win + hop - 1) / hop;
    const int h_b = h_a + (n_store - 1) / hop;
    // interior group: every sample is covered by all of its frames -> no per-sample frame tests
    const bool interior = !norm_wss || ((h_a + a / hop - (ncov - 1) >= 0) && (h_b + (hop - 1 + a) / hop <= T - 1));
    // (no NaN/Inf test here: the inputs are tested by k_prepare_mag / k_synth, the result by k_deemphasis)
    if (interior && ((hop | n_store | r0 | RS) & 1) == 0 && (reinterpret_cast<uintptr_t>(yo) & 7) == 0) {
        float2* s2 = reinterpret_cast<float2*>(ring);          // the ring is 16-byte aligned
        const float2* r2 = reinterpret_cast<const float2*>(rinv_s);   // indexed by the sample inside the group
        float2* yo2 = reinterpret_cast<float2*>(yo);
        const int n2 = n_store >> 1, RS2 = RS >> 1;
        const int n_a = min(n2, RS2 - (r0 >> 1));              // pairs before the ring wraps
        const float2* src = s2 + (r0 >> 1);
        for (int i = lane; i < n_a; i += 32) {
            yo2[i] = p_mul(src[i], __ldg(r2 + i));
            const_cast<float2*>(src)[i] = mk2(0.f, 0.f);
        }
        for (int i = n_a + lane; i < n2; i += 32) {            // the part that wrapped to the ring's start
            yo2[i] = p_mul(s2[i - n_a], __ldg(r2 + i));
            s2[i - n_a] = mk2(0.f, 0.f);
        }
    } else {
        for (int i = lane; i < n_store; i += 32) {
            const int hh = i / hop, j = i - hh * hop, h = h_a + hh;
            const int dj = (j + a) / hop, rj = (j + a) - dj * hop;
            int ncover = 0;
            for (int idx = rj; idx < win; idx += hop) ++ncover;
            const int k_hi = h + dj, k_lo = k_hi - (ncover - 1);
            int ri = r0 + i;
            if (ri >= RS) ri -= RS;
            float v = ring[ri];
            if (!norm_wss || (k_lo >= 0 && k_hi <= T - 1)) {
                v *= __ldg(rinv_s + j);
            } else {
                v *= (1.0f / (float)kNfft);
                float sm = 0.f;
                int kk = k_hi;
                for (int idx = rj; idx < win; idx += hop, --kk)
                    if (kk >= 0 && kk <= T - 1) { float w = win_at(wtab, wt0, lo + idx); sm = fmaf(w, w, sm); }
                if (sm > 1.17549435e-38f) v /= sm;
            }
            yo[i] = v;
            ring[ri] = 0.f;
        }
    }
}

template <int PRUNE> struct WinTable {          // the part of the n_fft-long window the kernel can touch
    static constexpr int n0 = PRUNE == 1 ? 512 : 0;
    static constexpr int len = PRUNE == 0 ? kNfft : 1024;
};

// where a group of the stream lives: utterance, utterance-local group index and the utterance's geometry
struct GroupLoc {
    int b;                   // utterance (index into the (sub-)batch), -1: past the end of the stream
    int g, Gb;               // group index inside the utterance, number of (real) groups of the utterance; g == Gb: end-halo
    int f_off, T, L;
    long long s_off;
};

// stream coordinate of utterance b's first group: its groups before it + one end-halo per earlier utterance
__device__ __forceinline__ int stream_off(const Batch& B, int b) { return __ldg(B.group_off + b) - B.group_base + b; }

__device__ __forceinline__ void fill_loc(const Batch& B, int V, int b, GroupLoc& o) {
    o.b = b;
    o.g = V - stream_off(B, b);
    o.Gb = __ldg(B.group_off + b + 1) - __ldg(B.group_off + b);
    o.f_off = __ldg(B.frame_off + b);
    o.T = __ldg(B.frame_off + b + 1) - o.f_off;
    o.s_off = __ldg(B.samp_off + b);
    o.L = (int)(__ldg(B.samp_off + b + 1) - o.s_off);          // hop*(T-1) (librosa) or hop*(T-1)+win (tf)
}
// binary search
__device__ __forceinline__ void locate(const Batch& B, int NV, int V, GroupLoc& o) {
    if (V >= NV) { o.b = -1; o.g = 0; o.Gb = 0; o.f_off = 0; o.T = 0; o.L = 0; o.s_off = 0; return; }
    int lo = 0, hi = B.batch;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (stream_off(B, mid) <= V) lo = mid; else hi = mid;
    }
    fill_loc(B, V, lo, o);
}
// walking leftwards from a known place: the utterance is the hint's or a few before it
__device__ __forceinline__ void locate_left(const Batch& B, int V, int b_hint, GroupLoc& o) {
    int b = b_hint;
    while (b > 0 && stream_off(B, b) > V) --b;
    fill_loc(B, V, b, o);
}

// frame load (kernels.cuh load_frame) for the paired window table, reading the waveform from L2 (it is rewritten by
// other SMs inside the launch)
template <int PRUNE>
__device__ __forceinline__ void load_frame_wp(c2 (&z)[32], const float* x, int L, int start, const float2* wp, int lane, float* stage) {
    constexpr int t0 = PruneRange<PRUNE>::t0, t1 = PruneRange<PRUNE>::t1;
    if (start + 64 * t0 >= 0 && start + 64 * t1 <= L) {
        const float* xs = x + start + lane;
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            if (t >= t0 && t < t1) z[t] = p_mul(mk2(__ldcg(xs + 64 * t), __ldcg(xs + 64 * t + 32)), wp[t * 32 + lane]);
            else z[t] = mk2(0.f, 0.f);
        }
    } else {
#pragma unroll 1
        for (int n = 64 * t0 + lane; n < 64 * t1; n += 32) stage[n] = sample_at<false, true>(x, L, start + n, 0.f);
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            if (t >= t0 && t < t1) z[t] = p_mul(mk2(stage[64 * t + lane], stage[64 * t + 32 + lane]), wp[t * 32 + lane]);
            else z[t] = mk2(0.f, 0.f);
        }
        __syncwarp();
    }
}

// cp.async the samples n in [64*T0, 64*T1) of a frame into the warp's scratch tile.  16-byte L2-only copies: the waveform
// is rewritten by other SMs inside the same launch, nothing of it may sit in L1.  `src` (the frame's sample 64*T0) is
// 8-byte aligned; the copy starts at the 16-byte boundary at or below it and `off` (0 or 2 floats) tells the reader
// where sample n landed: stage[n + off].
template <int T0, int T1>
__device__ __forceinline__ void stage_frame(float* stage, const float* src0, int lane, int& off) {
    off = (int)((reinterpret_cast<uintptr_t>(src0) >> 2) & 3);
    const float* src = src0 - off + 4 * lane;
    float* dst = stage + 64 * T0 + 4 * lane;
#pragma unroll
    for (int q = 0; q < (T1 - T0) / 2; ++q) cp_async16(dst + 128 * q, src + 128 * q);
    if (lane == 0) cp_async16(dst + 64 * (T1 - T0), src + 64 * (T1 - T0));
}

// FB: one CTA barrier after every colour step (the production mode).  It keeps the 8 warps in the same stretch of code, which
// is what the instruction caches need, and it orders the overlap-add by itself: colour s starts when every warp has added
// its colours < s, and a group is stored before the barrier of its last colour, so ring space is never reused before it is
// free.  Without FB the per-warp event counters do both (kept for window/hop ratios where a group's hops need ALL colours of
// the next group, and for the barrier experiments of profiles/r1/sweep_kernels.txt).
template <int PRUNE, bool DEFCFG, bool TFM, bool FB>
__global__ void __launch_bounds__(kThreads, 2) k_gl_stream(GlStreamParams P) {
    NSB_DYN_SMEM(smem_raw);
    const int hop = DEFCFG ? 250 : P.plan.hop;
    const int win = DEFCFG ? 1000 : P.plan.win_len;
    constexpr int LO = TFM ? 0 : 524;                // DEFCFG: librosa pads the window centrally, tf.contrib.signal on the right
    const int lo = DEFCFG ? LO : P.plan.lo;
    const int C = DEFCFG ? 4 : P.colours;
    const int origin = DEFCFG ? (TFM ? 0 : kNfft / 2) : P.plan.origin;
    const int a = origin - lo;                       // frame k's window support starts at sample k*hop - a
    // hop h is touched by the frames h + kfirst0 .. h + klast0
    const int kfirst0 = (a - win >= 0) ? (a - win) / hop + 1 : -((win - a - 1) / hop + 1) + 1;
    const int klast0 = (hop - 1 + a) / hop;
    const int back = min(klast0 - kfirst0, C);       // hops a group reaches back = colours of group j+1 that group j's hops need
    const int GH = C * hop;                          // samples per group
    const int RS = (kWarpsPerCta * C + back) * hop;  // ring size: 8 groups + the hops the leftmost one reaches back
    constexpr int t0 = PruneRange<PRUNE>::t0, t1 = PruneRange<PRUNE>::t1;

    // twiddles w2048^(q*l) in PAIRS of rows: tw4[p*32 + l] = (w^((2p+1) l), w^((2p+2) l)), p = 0..14, row 31 after them -
    // one LDS.128 serves two twiddle multiplications (every instruction less counts: the kernel is bound by instruction
    // supply, profiles/r1/microbench_icache.txt).  The window likewise in the pairs the registers want:
    // wp[(t - t0)*32 + lane] = (w[64 t + lane], w[64 t + 32 + lane]).
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    const float4* tw4 = reinterpret_cast<const float4*>(tw_s);
    const float2* tw31 = tw_s + 15 * 64;
    float2* wp_tab = tw_s + kTwF2;
    const float2* wp = wp_tab - t0 * 32;                     // wp[t*32 + lane] for t in [t0, t1)
    float* ring = reinterpret_cast<float*>(wp_tab + (t1 - t0) * 32);   // 16-byte aligned
    int* progress = reinterpret_cast<int*>(ring + ((RS + 3) & ~3));    // [kWarpsPerCta] (+ pad to 16 ints)
    float2* scratch_all = reinterpret_cast<float2*>(progress + 16);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* scratch = scratch_all + warp * kScratchF2;
    float* stage = reinterpret_cast<float*>(scratch);
    if (P.trace && threadIdx.x == 0) { P.trace[3 * blockIdx.x] = sm_id(); P.trace[3 * blockIdx.x + 1] = global_ns(); }

    load_twiddle_pairs(tw_s, P.plan.tw);
    for (int i = threadIdx.x; i < (t1 - t0) * 32; i += kThreads) {
        const int n = 64 * (t0 + i / 32) + (i & 31);
        wp_tab[i] = make_float2(P.plan.win[n], P.plan.win[n + 32]);
    }
    {
        float4* r4 = reinterpret_cast<float4*>(ring);
        for (int i = threadIdx.x; i < (RS + 3) / 4; i += kThreads) r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (threadIdx.x < 16) progress[threadIdx.x] = 0;
    __syncthreads();
    // (no non-finite test in this kernel: the inputs are tested by k_prepare_mag / k_synth, the result by k_deemphasis)
    const int E = C + 1;                              // events per round: C adds + 1 store

    // ---- (iteration, chunk) items from one global counter (see k_gl_iter for why) ----
    const int NV = P.total_groups + P.batch.batch;
    const int CH = P.chunk_groups;
    const int n_chunks = (NV + CH - 1) / CH;
    const long long total_items = (long long)P.iters * n_chunks;
    int item_it = 0;
    if (threadIdx.x == 0) progress[8] = atomicAdd(P.item_counter, 1);
    __syncthreads();
    for (long long item = progress[8]; item < total_items; item = progress[8 + (++item_it & 1)]) {
    const int n_it = (int)(item / n_chunks);
    const int chunk = (int)(item - (long long)n_it * n_chunks);
    const float* y_in = P.ybuf[(P.cur0 + n_it) & 1];
    float* y_out = P.ybuf[(P.cur0 + n_it + 1) & 1];
    // four lanes, four round trips to L2 side by side: the next item, and the three chunks whose iteration n-1 this chunk reads
    // (c-1, c, c+1; my stores also overwrite what THEY read in iteration n-1 - one wait covers both hazards)
    if (threadIdx.x == 0) progress[8 + ((item_it + 1) & 1)] = atomicAdd(P.item_counter, 1);     // read after this chunk's barriers
    if (n_it > 0 && threadIdx.x >= 1 && threadIdx.x <= 3) {
        int c = chunk + (int)threadIdx.x - 2;
        c = c < 0 ? 0 : (c >= n_chunks ? n_chunks - 1 : c);
        while (gflag_load(P.done + c) < n_it) spin_pause();
    }
    if (threadIdx.x < kWarpsPerCta) progress[threadIdx.x] = 0;     // the ring is all zero again after every chunk
    __syncthreads();
    const int V0 = chunk * CH, V1 = min(NV, V0 + CH);
    const int n_pos = V1 - V0 + 1;                    // positions: i = 0 is the halo group V1, i = n_pos - 1 the group V0
    // samples left of the range's first group belong to the CTA on the left: frames of V0 must not add there
    int b_low, x_low;
    {
        GroupLoc l0;
        locate(P.batch, NV, V0, l0);
        b_low = l0.b; x_low = l0.g * GH;
    }

    bool staged = false;                         // this frame's samples were cp.async'ed into the scratch tile already
    int stage_off = 0;                           // ... shifted by this many floats (16-byte alignment of the copies)
    GroupLoc cur;
    const int n_pad = (n_pos + kWarpsPerCta - 1) & ~(kWarpsPerCta - 1);   // phantom positions fill the last round (barriers are CTA-wide)
    locate(P.batch, NV, warp < n_pos ? V1 - warp : NV, cur);
    for (int i = warp; i < n_pad; i += kWarpsPerCta) {      // position i counts groups from the range's right end
        const int u = n_pos - 1 - i;                        // stream coordinate of the group relative to V0
        const int r = i >> 3;
        const bool has_right = (i >= 1);                    // group V+1: neighbour warp, this round or (warp 0) the previous one
        const int wr = (warp + 7) & 7, r_r = (i - 1) >> 3;
        const bool has_left = (warp < kWarpsPerCta - 1) && (i + 1 < n_pos);   // group V-1 in the same round: pacing only
        const int kbase = C * cur.g + kfirst0;
        // frames of this group that exist and that this range needs (the halo group V1: only those that reach back, and
        // none at all if V1 starts an utterance)
        int k_hi = cur.T - 1;
        if (i == 0) k_hi = (cur.g == 0) ? -1 : min(k_hi, kbase + back - 1);
        if (cur.b < 0) k_hi = -1;
        const int T = cur.T, L = cur.L;
        const float* yin = y_in + cur.s_off;
        const float* mag0 = P.mag + (size_t)cur.f_off * kMagPitch;
        const int xmin = (cur.b == b_low) ? x_low : 0;      // lowest utterance sample this CTA accumulates
        const int ubase = (u - cur.g) * GH;                 // utterance sample x sits at ring[(ubase + x) mod RS]
        // the group's own C hops are final once its last colour is in: normalise, store, clear the ring
        auto store_group = [&]() {
            if (i >= 1 && cur.b >= 0 && cur.g < cur.Gb) {
                const int x0 = cur.g * GH;                               // utterance sample of the group's first hop
                const int n_store = min(GH, L - x0);
                if (n_store > 0)
                    gl_store_group(ring, (u * GH) % RS, RS, y_out + cur.s_off + x0, n_store, C * cur.g, T, reinterpret_cast<const float*>(wp_tab), t0,
                                   P.plan.rinv, hop, win, lo, a, P.plan.norm_wss, lane);
            } else if (i == 0 && k_hi >= 0) {
                // the halo group's own hops belong to the chunk on the right: nothing to store, but positions 8 and 9 reuse the space
                const int r0 = (u * GH) % RS;
                for (int q = lane; q < GH; q += 32) { int ri = r0 + q; if (ri >= RS) ri -= RS; ring[ri] = 0.f; }
            }
        };
        for (int s = 0; s < C; ++s) {
            const int k = kbase + s;
            const bool active = (k >= 0 && k <= k_hi);        // warp-uniform
            const bool next_active = (s + 1 < C) && (k + 1 >= 0 && k + 1 <= k_hi);   // the group's next frame
            c2 z[32];
            if (active) {
                // pull the NEXT frame's magnitude row towards L2 (it streams from HBM) while this frame computes
                if (next_active) {
                    const char* nm = reinterpret_cast<const char*>(mag0 + (size_t)(k + 1) * kMagPitch);
                    prefetch_l2(nm + lane * 128);
                    if (lane == 0) prefetch_l2(nm + 4096);
                }
                const float* magrow = mag0 + (size_t)k * kMagPitch;
                if (staged) {
                    // samples were staged by the previous frame of this warp (see below): window them from shared memory
                    cp_async_wait_all();
                    __syncwarp();
                    const float* sg = stage + stage_off;
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        if (t >= t0 && t < t1) z[t] = p_mul(mk2(sg[64 * t + lane], sg[64 * t + 32 + lane]), wp[t * 32 + lane]);
                        else z[t] = mk2(0.f, 0.f);
                    }
                    __syncwarp();
                } else {
                    load_frame_wp<PRUNE>(z, yin, L, k * hop - origin, wp, lane, stage);
                }
                fwd_phase1_tw4<PruneRange<PRUNE>::t0, PruneRange<PRUNE>::t1>(z, lane, scratch, tw4, tw31);
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 32; ++t) z[t] = scratch[lane * kRowStride + t];
                __syncwarp();                        // every lane has its row: the scratch tile is free
                {
                    const char* src = reinterpret_cast<const char*>(magrow);
                    char* dst = reinterpret_cast<char*>(scratch);
#pragma unroll
                    for (int q = 0; q < 8; ++q) cp_async16(dst + (q * 32 + lane) * 16, src + (q * 32 + lane) * 16);
                    if (lane == 0) cp_async16(dst + 4096, src + 4096);
                }
                fft32<-1>(z);
                float2* xch = scratch + kXchOffsetF2;
                if (lane == 0) {       // (128-bit stores would need the two pairs in four consecutive registers: four MOVs each)
#pragma unroll
                    for (int q = 0; q < 32; ++q) xch[q] = z[q];
                }
                cp_async_wait_all();
                __syncwarp();
                const float4* mrow = reinterpret_cast<const float4*>(scratch);
                const float* mflt = reinterpret_cast<const float*>(scratch);
                bool zero = false;
                // (a) every lane renormalises its 32 slots (lane 0's registers hold the packed-row FFT, not bins:
                //     replaced below)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 S = mrow[q * 32 + lane];
                    if (TFM) {
                        renorm_tf(z[4 * q], S.x, P.inv_thr); renorm_tf(z[4 * q + 1], S.y, P.inv_thr);
                        renorm_tf(z[4 * q + 2], S.z, P.inv_thr); renorm_tf(z[4 * q + 3], S.w, P.inv_thr);
                    } else {
                        renorm_fast(z[4 * q], S.x, zero);
                        renorm_fast(z[4 * q + 1], S.y, zero);
                        renorm_fast(z[4 * q + 2], S.z, zero);
                        renorm_fast(z[4 * q + 3], S.w, zero);
                    }
                }
                if (lane == 0) zero = false;
                if (!TFM && warp_any(zero)) {                // rare: some bin of y's STFT is exactly 0 -> phase 0 (np.angle(0))
                    if (lane != 0) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 S = mrow[q * 32 + lane];
                            const float Ss[4] = {S.x, S.y, S.z, S.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (z[4 * q + e].x == 0.f && z[4 * q + e].y == 0.f) z[4 * q + e].x = Ss[e];
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
                        const int jq = lane, jj = 32 - lane;
                        // u = -i * w64^j with w64^j = w2048^(16 * 2j) from the twiddle table (j = 16: w = -i)
                        const float2 wj = (jq <= 15) ? tw_s[(7 * 32 + 2 * jq) * 2 + 1] : make_float2(0.f, -1.f);   // row 16 of the table
                        const c2 u = mk2(wj.y, -wj.x);
                        const c2 Aj = xch[jq], Bj = xch[jj];
                        const c2 S1 = cadd_conj(Aj, Bj), D1 = csub_conj(Aj, Bj);
                        const c2 T1 = cmul(D1, u);
                        c2 c1 = cadd(S1, T1);                               // 2*X[32 j]
                        c2 c2v = cconj(csub(S1, T1));                       // 2*X[32 (32-j)]
                        bool z2 = false;
                        if (TFM) {
                            renorm_tf(c1, mflt[(jq >> 2) * 128 + (jq & 3)], P.inv_thr);
                            renorm_tf(c2v, mflt[(jj >> 2) * 128 + (jj & 3)], P.inv_thr);
                        } else {
                            renorm_fast(c1, mflt[(jq >> 2) * 128 + (jq & 3)], z2);
                            renorm_fast(c2v, mflt[(jj >> 2) * 128 + (jj & 3)], z2);
                        }
                        if (z2) {
                            if (c1.x == 0.f && c1.y == 0.f) c1.x = mflt[(jq >> 2) * 128 + (jq & 3)];
                            if (c2v.x == 0.f && c2v.y == 0.f) c2v.x = mflt[(jj >> 2) * 128 + (jj & 3)];
                        }
                        // inverse split: S' = V[j] + conj V[32-j], D' = V[j] - conj V[32-j], P = D' * conj(u)
                        const c2 S2 = cadd_conj(c1, c2v), D2 = csub_conj(c1, c2v);
                        const c2 Pv = cmul_conj(D2, u);
                        // lane j reads and writes only xch[j] and xch[32-j]: no cross-lane hazard inside this block
                        xch[jq] = cadd(S2, Pv);
                        if (jq != 16) xch[jj] = cconj(csub(S2, Pv));
                    }
                }
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) z[q] = xch[q];
                }
                __syncwarp();                        // magnitude row and exchange area fully consumed
                // inverse pass 1 (the lane-0 pre-split already happened above)
                fft32<+1>(z);
                scratch[lane * kRowStride] = z[0];
#pragma unroll
                for (int p = 0; p < 15; ++p) {
                    const float4 w = tw4[p * 32 + lane];
                    z[2 * p + 1] = cmul_conj(z[2 * p + 1], mk2(w.x, w.y));
                    z[2 * p + 2] = cmul_conj(z[2 * p + 2], mk2(w.z, w.w));
                }
                z[31] = cmul_conj(z[31], tw31[lane]);
#pragma unroll
                for (int q = 1; q < 32; ++q) scratch[lane * kRowStride + q] = z[q];
                __syncwarp();
                inv_phase2(z, lane, scratch);
                __syncwarp();                        // the scratch tile may be rewritten by this warp's next frame
            }
            // The scratch tile now idles through the overlap-add: stage the group's NEXT frame into it with cp.async so
            // that its load latency hides behind the neighbour wait and the accumulate.  Only frames that need no
            // reflect padding and (8-byte copies) start at an even sample.
            staged = false;
            if (PRUNE != 0 && next_active) {
                const long long nstart = (long long)(k + 1) * hop - origin;
                if (nstart + 64 * t0 >= 0 && nstart + 64 * t1 <= L && ((cur.s_off + nstart) & 1) == 0) {
                    stage_frame<t0, t1>(stage, yin + nstart + 64 * t0, lane, stage_off);
                    staged = true;
                }
            }
            // ---- overlap-add ordering ----
            if (FB) {
                // the barrier at the end of the previous step ordered everything - or, split-phase (sync_mode bit 3): every warp
                // published its finished steps in progress[], and only now, with the frame's transforms done, waits for the others
                if (P.sync_mode & 8) {
                    if (lane < kWarpsPerCta) while (flag_load(progress + lane) < (i >> 3) * C + s) spin_pause();
                    __syncwarp();
                }
            } else if (s == 0) {
                // ring reuse: this group writes where positions i-8 (this warp) and i-9 (the right-hand warp) stored
                if (i >= 9) wait_events(progress + wr, E * (((i - 9) >> 3) + 1), lane);
            } else {
                // my colour-s frame overlaps the right-hand group's frames of colour < s: they go first
                if (has_right) wait_events(progress + wr, E * r_r + s, lane);
                // pacing: stay within one colour of the left-hand neighbour (keeps the CTA's warps in the same code)
                if (has_left && !(P.sync_mode & 4)) wait_events(progress + warp + 1, E * r + s, lane);
            }
            if (active) {
                const int base = k * hop - origin;                            // utterance sample of n = 0
                int q0 = (ubase + base + lo) % RS;                            // ring index of the first support sample
                if (q0 < 0) q0 += RS;
                const bool inside = (base + lo >= xmin) && (base + lo + win <= L);
                if (DEFCFG && inside && q0 + win <= RS) {
                    // default hparams: the support n in [LO, LO + 1000) is known at compile time; (acc[n], acc[n+32]) and the
                    // two window values ride in register pairs so the accumulate is one FFMA2
                    ola_fixed_support<LO, t0, t1>(ring + (q0 - lo) + lane, z, lane, [&](int t) { return wp[t * 32 + lane]; });
                } else {
                    // window support [lo, lo+win) clipped to the piece -> per-lane bitmasks of the valid t
                    // (n = 64 t + lane [+32]); indices wrap around the ring.  Touching nothing outside the support
                    // makes the plain read-modify-write race-free.
                    const int nlo = max(lo, xmin - base), nhi = min(lo + win, L - base);
                    const int a0 = min(max((nlo - lane + 63) >> 6, 0), 32), a1 = min(max((nhi - lane + 63) >> 6, 0), 32);
                    const int b0 = min(max((nlo - lane - 32 + 63) >> 6, 0), 32), b1 = min(max((nhi - lane - 32 + 63) >> 6, 0), 32);
                    const unsigned mre = (a1 > a0) ? ((0xffffffffu >> (32 - (a1 - a0))) << a0) : 0u;
                    const unsigned mim = (b1 > b0) ? ((0xffffffffu >> (32 - (b1 - b0))) << b0) : 0u;
                    const int qb = q0 - lo + lane;
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        if (t >= t0 && t < t1) {
                            int i0 = qb + 64 * t, i1 = i0 + 32;
                            if (i0 >= RS) i0 -= RS;
                            if (i1 >= RS) i1 -= RS;
                            const c2 w = wp[t * 32 + lane];
                            if ((mre >> t) & 1u) ring[i0] = fmaf(z[t].x, w.x, ring[i0]);
                            if ((mim >> t) & 1u) ring[i1] = fmaf(z[t].y, w.y, ring[i1]);
                        }
                    }
                }
            }
            if (FB) {
                if (s == C - 1) { __syncwarp(); store_group(); }      // the lanes read each other's accumulates
                if (P.sync_mode & 8) {
                    __syncwarp();
                    if (lane == 0) flag_store(progress + warp, (i >> 3) * C + s + 1);
                } else {
                    __syncthreads();
                }
            } else {
                __syncwarp();
                if (lane == 0) flag_store(progress + warp, E * r + s + 1);
                if ((P.sync_mode & 3) == 3) half_cta_barrier(warp);
            }
        }
        // ---- the group is complete (z is dead from here on) ----
        // the warp's next group sits 8 positions to the left: find it, send its first frame on its way
        GroupLoc nxt;
        const bool more = (i + kWarpsPerCta < n_pos);
        if (!more) locate(P.batch, NV, NV, nxt);              // phantom
        if (more) {
            locate_left(P.batch, V1 - i - kWarpsPerCta, cur.b < 0 ? P.batch.batch - 1 : cur.b, nxt);
            const int kn = C * nxt.g + kfirst0;               // its colour-0 frame (does not exist at an utterance's start)
            if (PRUNE != 0 && kn >= 0 && kn <= nxt.T - 1) {
                const char* nm = reinterpret_cast<const char*>(P.mag + ((size_t)nxt.f_off + kn) * kMagPitch);
                prefetch_l2(nm + lane * 128);
                if (lane == 0) prefetch_l2(nm + 4096);
                const long long nstart = (long long)kn * hop - origin;
                if (nstart + 64 * t0 >= 0 && nstart + 64 * t1 <= nxt.L && ((nxt.s_off + nstart) & 1) == 0) {
                    stage_frame<t0, t1>(stage, y_in + nxt.s_off + nstart + 64 * t0, lane, stage_off);
                    staged = true;
                }
            }
        }
        if (!FB) {
            // my hops also need the right-hand group's first `back` colours; the colour C-1 wait covered back <= C-1
            if (back > C - 1 && i >= 1) wait_events(progress + wr, E * r_r + back, lane);
            store_group();
            __syncwarp();
            if (lane == 0) flag_store(progress + warp, E * (r + 1));
        }
        cur = nxt;
        if (!FB && (P.sync_mode & 3) == 1) __syncthreads();
    }
    __syncthreads();                                 // every thread's stores of this chunk are issued
    if (threadIdx.x == 0) { __threadfence(); gflag_store(P.done + chunk, n_it + 1); }
    }   // items
    if (P.trace) {
        __syncthreads();
        if (threadIdx.x == 0) P.trace[3 * blockIdx.x + 2] = global_ns();
    }
}

}  // namespace nsb
