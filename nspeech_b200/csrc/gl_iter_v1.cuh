// EXPERIMENT ONLY (A/B timing): the first, fully inlined version of the Griffin-Lim iteration kernel
// (four FFT32 copies, CTA barrier per tile).  Selected with nsb_set_generic_iteration(h, 2 / 3).
#pragma once
#include "gl_iter.cuh"

namespace nsb {

template <bool PRUNE, bool DEFCFG, bool ROTATE>
__global__ void __launch_bounds__(kThreads, 2) k_gl_iter_v1(GlParams P) {
    NSB_DYN_SMEM(smem_raw);
    const int hop = DEFCFG ? 250 : P.plan.hop;
    const int win = DEFCFG ? 1000 : P.plan.win_len;
    const int lo = DEFCFG ? 524 : P.plan.lo;
    const int C = DEFCFG ? 4 : P.colours;
    const int H = P.tile_hops;
    float2* tw_s = reinterpret_cast<float2*>(smem_raw);
    float* win_s = reinterpret_cast<float*>(tw_s + kTwF2);
    float* rinv_s = win_s + kNfft;
    float* acc = rinv_s + hop;
    size_t acc_end = (size_t)(acc + (size_t)H * hop - reinterpret_cast<float*>(smem_raw));
    acc_end = (acc_end + 3) & ~(size_t)3;
    int* progress = reinterpret_cast<int*>(reinterpret_cast<float*>(smem_raw) + acc_end);   // [kWarpsPerCta] (+pad to 16 ints)
    float2* scratch_all = reinterpret_cast<float2*>(progress + 16);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* scratch = scratch_all + warp * kScratchF2;
    const int a = kNfft / 2 - lo;

    for (int i = threadIdx.x; i < kTwF2; i += kThreads) tw_s[i] = P.plan.tw[i];
    for (int i = threadIdx.x; i < kNfft; i += kThreads) win_s[i] = P.plan.win[i];
    __syncthreads();
    for (int j = threadIdx.x; j < hop; j += kThreads) {
        float s = 0.f;
        for (int idx = (j + a) % hop; idx < win; idx += hop) { float w = win_s[lo + idx]; s = fmaf(w, w, s); }
        rinv_s[j] = s > 1.17549435e-38f ? 1.0f / s : 1.0f;
    }
    // per-lane validity of the window support for the overlap-add (generic configs)
    unsigned mre = 0xffffffffu, mim = 0xffffffffu;
    if (!DEFCFG) {
        int a0 = max((lo - lane + 63) >> 6, 0), a1 = min(max((lo + win - lane + 63) >> 6, 0), 32);
        int b0 = max((lo - lane - 32 + 63) >> 6, 0), b1 = min(max((lo + win - lane - 32 + 63) >> 6, 0), 32);
        mre = (a1 > a0) ? ((0xffffffffu >> (32 - (a1 - a0))) << a0) : 0u;
        mim = (b1 > b0) ? ((0xffffffffu >> (32 - (b1 - b0))) << b0) : 0u;
    }
    // twiddle of the distributed real-64 split for this lane: u = -i * w64^lane  (lanes 1..16)
    float ur, ui;
    {
        float2 w = (lane >= 1 && lane <= 15) ? tw_s[15 * 32 + 2 * lane] : make_float2(0.f, -1.f);   // w2048^(16*2*lane) = w64^lane
        ur = w.y; ui = -w.x;
    }
    bool bad = false;

    for (int tile_g = blockIdx.x; tile_g < P.total_tiles; tile_g += gridDim.x) {
        const int b = find_segment(P.batch.tile_off, P.batch.batch, tile_g);
        const int tile = tile_g - __ldg(P.batch.tile_off + b);
        const int f_off = __ldg(P.batch.frame_off + b);
        const int T = __ldg(P.batch.frame_off + b + 1) - f_off;
        const long long s_off = __ldg(P.batch.samp_off + b);
        const long long L = (long long)hop * (T - 1);
        const int h0 = tile * H, h1 = min(h0 + H, T - 1);
        const long long s0 = (long long)h0 * hop, s1 = (long long)h1 * hop;
        const int n_out = (int)(s1 - s0);
        for (int i = threadIdx.x; i < n_out; i += kThreads) acc[i] = 0.f;
        if (threadIdx.x < 16) progress[threadIdx.x] = 0;
        long long num = s0 + a - win;
        int k_min = (int)(num >= 0 ? num / hop + 1 : -((-num - 1) / hop + 1) + 1);
        if (k_min < 0) k_min = 0;
        int k_max = (int)((s1 + a + hop - 1) / hop - 1);
        if (k_max > T - 1) k_max = T - 1;
        if (k_max - k_min + 1 > kWarpsPerCta * C) bad = true;   // host sizes tiles so this cannot happen; never drop frames silently
        __syncthreads();

        for (int s = 0; s < C; ++s) {
            // the warp's group is the C consecutive frames kg .. kg+C-1; it adds them in the order of the GLOBAL colour
            // k mod C (step s takes the frame with k % C == s), so the summation order of every output sample is
            // independent of how the utterance was tiled: a batch is bit-identical to one-at-a-time calls
            const int kg = k_min + C * warp;
            const int k = ROTATE ? kg + ((s - kg % C) + C) % C : kg + s;
            const bool active = (k <= k_max);        // warp-uniform
            float re[32], im[32];
            if (active) {
                const int fg = f_off + k;
                load_frame<false, PRUNE>(re, im, P.y_in + s_off, L, (long long)k * hop - kNfft / 2, win_s, lane, 0.f, 1.0f,
                                         reinterpret_cast<float*>(scratch));
                fwd_phase1(re, im, lane, scratch, tw_s);
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 32; ++t) { float2 v = scratch[lane * kRowStride + t]; re[t] = v.x; im[t] = v.y; }
                __syncwarp();                        // every lane has its row: the scratch tile is free
                {
                    const char* src = reinterpret_cast<const char*>(P.mag + (size_t)fg * kMagPitch);
                    char* dst = reinterpret_cast<char*>(scratch);
#pragma unroll
                    for (int g = 0; g < 8; ++g) cp_async16(dst + (g * 32 + lane) * 16, src + (g * 32 + lane) * 16);
                    if (lane == 0) cp_async16(dst + 4096, src + 4096);
                }
                fft32<-1>(re, im);
                float2* xch = scratch + kXchOffsetF2;
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) xch[j] = make_float2(re[j], im[j]);
                }
                cp_async_wait_all();
                __syncwarp();
                const float4* mrow = reinterpret_cast<const float4*>(scratch);
                const float* mflt = reinterpret_cast<const float*>(scratch);
                bool zero = false;
                // (a) every lane renormalises its 32 slots (lane 0's registers hold the packed-row FFT, not bins:
                //     its results are discarded below)
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    float4 S = mrow[g * 32 + lane];
                    renorm_fast(re[4 * g], im[4 * g], S.x, zero);
                    renorm_fast(re[4 * g + 1], im[4 * g + 1], S.y, zero);
                    renorm_fast(re[4 * g + 2], im[4 * g + 2], S.z, zero);
                    renorm_fast(re[4 * g + 3], im[4 * g + 3], S.w, zero);
                }
                if (lane == 0) zero = false;
                if (warp_any(zero)) {                // rare: some bin of y's STFT is exactly 0 -> phase 0 (np.angle(0))
                    if (lane != 0) {
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            float4 S = mrow[g * 32 + lane];
                            if (re[4 * g] == 0.f && im[4 * g] == 0.f) re[4 * g] = S.x;
                            if (re[4 * g + 1] == 0.f && im[4 * g + 1] == 0.f) re[4 * g + 1] = S.y;
                            if (re[4 * g + 2] == 0.f && im[4 * g + 2] == 0.f) re[4 * g + 2] = S.z;
                            if (re[4 * g + 3] == 0.f && im[4 * g + 3] == 0.f) re[4 * g + 3] = S.w;
                        }
                    }
                }
                // (b) bins k = 32 j (rows 0/32): lane j in 1..16 does pair (j, 32-j); lane 0 the real DC/Nyquist pair
                if (lane <= 16) {
                    if (lane == 0) {
                        float2 g0 = xch[0];
                        float x0 = g0.x + g0.y, xn = g0.x - g0.y;          // (X[0], X[1024]) up to the factor 2
                        float S0 = mflt[0], Sn = mflt[1024];
                        x0 = (x0 < 0.f) ? -S0 : S0;                         // phase of a real number is its sign
                        xn = (xn < 0.f) ? -Sn : Sn;
                        xch[0] = make_float2(x0 + xn, x0 - xn);
                    } else {
                        const int j = lane, jj = 32 - lane;
                        float2 A = xch[j], B = xch[jj];
                        float sr = A.x + B.x, si = A.y - B.y, dr = A.x - B.x, di = A.y + B.y;
                        float tr = dr * ur - di * ui, ti = dr * ui + di * ur;
                        float c1r = sr + tr, c1i = si + ti;                 // 2*X[32 j]
                        float c2r = sr - tr, c2i = -(si - ti);              // 2*X[32 (32-j)]
                        bool z2 = false;
                        renorm_fast(c1r, c1i, mflt[(j >> 2) * 128 + (j & 3)], z2);
                        renorm_fast(c2r, c2i, mflt[(jj >> 2) * 128 + (jj & 3)], z2);
                        if (z2) {
                            if (c1r == 0.f && c1i == 0.f) c1r = mflt[(j >> 2) * 128 + (j & 3)];
                            if (c2r == 0.f && c2i == 0.f) c2r = mflt[(jj >> 2) * 128 + (jj & 3)];
                        }
                        // inverse split: S' = V[j] + conj V[32-j], D' = V[j] - conj V[32-j], P = D' * conj(u)
                        float s2r = c1r + c2r, s2i = c1i - c2i, d2r = c1r - c2r, d2i = c1i + c2i;
                        float pr = d2r * ur + d2i * ui, pi = d2i * ur - d2r * ui;
                        // lane j reads and writes only xch[j] and xch[32-j]: no cross-lane hazard inside this block
                        xch[j] = make_float2(s2r + pr, s2i + pi);
                        if (j != 16) xch[jj] = make_float2(s2r - pr, -s2i + pi);
                    }
                }
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) { float2 v = xch[j]; re[j] = v.x; im[j] = v.y; }
                }
                __syncwarp();                        // magnitude row and exchange area fully consumed
                // inverse pass 1 (the lane-0 pre-split already happened above)
                fft32<+1>(re, im);
                scratch[lane * kRowStride] = make_float2(re[0], im[0]);
#pragma unroll
                for (int r = 1; r < 32; ++r) {
                    float2 w = tw_s[(r - 1) * 32 + lane];
                    scratch[lane * kRowStride + r] = make_float2(re[r] * w.x + im[r] * w.y, im[r] * w.x - re[r] * w.y);
                }
                __syncwarp();
                inv_phase2(re, im, lane, scratch);
            }
            // overlap-add ordering: my frame number s overlaps the neighbours' frames < s (warp+1) and > s (warp-1)
            if (s > 0) {
                if (lane == 0) {
                    if (warp > 0) while (flag_load(progress + warp - 1) < s) spin_pause();
                    if (warp < kWarpsPerCta - 1) while (flag_load(progress + warp + 1) < s) spin_pause();
                }
                __syncwarp();
            }
            if (active) {
                const long long base = (long long)k * hop - kNfft / 2 - s0;
                constexpr int t0 = PRUNE ? 8 : 0, t1 = PRUNE ? 24 : 32;
                const bool inside = (base + lo >= 0) && (base + lo + win <= n_out);
                if (inside) {
                    float* ap = acc + base + lane;
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        if (t >= t0 && t < t1) {
                            const bool vr = DEFCFG ? (t > 8 || lane >= 12) : ((mre >> t) & 1u);
                            const bool vi = DEFCFG ? (t < 23 || lane < 20) : ((mim >> t) & 1u);
                            if (vr) ap[64 * t] = fmaf(re[t], win_s[64 * t + lane], ap[64 * t]);
                            if (vi) ap[64 * t + 32] = fmaf(im[t], win_s[64 * t + 32 + lane], ap[64 * t + 32]);
                        }
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        if (t >= t0 && t < t1) {
                            const bool vr = DEFCFG ? (t > 8 || lane >= 12) : ((mre >> t) & 1u);
                            const bool vi = DEFCFG ? (t < 23 || lane < 20) : ((mim >> t) & 1u);
                            long long i0 = base + 64 * t + lane, i1 = i0 + 32;
                            if (vr && i0 >= 0 && i0 < n_out) acc[i0] = fmaf(re[t], win_s[64 * t + lane], acc[i0]);
                            if (vi && i1 >= 0 && i1 < n_out) acc[i1] = fmaf(im[t], win_s[64 * t + 32 + lane], acc[i1]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) flag_store(progress + warp, s + 1);
        }
        __syncthreads();
        float* yo = P.y_out + s_off + s0;
        for (int j = threadIdx.x; j < hop; j += kThreads) {
            const int dj = (j + a) / hop, rj = (j + a) - dj * hop;
            const float ri = rinv_s[j];
            int ncover = 0;
            for (int idx = rj; idx < win; idx += hop) ++ncover;
            for (int h = h0; h < h1; ++h) {
                const int k_hi = h + dj, k_lo = k_hi - (ncover - 1);
                float v = acc[(h - h0) * hop + j] * (1.0f / (float)kNfft);
                bad |= !isfinite(v);
                if (k_lo >= 0 && k_hi <= T - 1) {
                    v *= ri;
                } else {
                    float sm = 0.f;
                    int kk = k_hi;
                    for (int idx = rj; idx < win; idx += hop, --kk)
                        if (kk >= 0 && kk <= T - 1) { float w = win_s[lo + idx]; sm = fmaf(w, w, sm); }
                    if (sm > 1.17549435e-38f) v /= sm;
                }
                yo[(size_t)(h - h0) * hop + j] = v;
            }
        }
        __syncthreads();                             // acc and progress are reused by the next tile
    }
    if (bad) atomicOr(P.status, 1);
}

}  // namespace nsb
