// C-ABI implementation: handle, plans, workspaces, launches.  See include/nspeech_b200.h.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/nspeech_b200.h"
#include "kernels.cuh"
#include "gl_stream.cuh"
#include "gen_kernels.cuh"

using namespace nsb;

// kernel launch: <<<>>> on the GPU; under NSB_EMULATE (tests/emu only) the CPU thread emulator
#ifdef NSB_EMULATE
#define NSB_LAUNCH(kern, grid, block, smem, st, ...) nsb_emu::launch((grid), (block), (smem), [&] { (kern)(__VA_ARGS__); })
#else
#define NSB_LAUNCH(kern, grid, block, smem, st, ...) (kern)<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__)
#endif

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? NSB_ERR_OOM : NSB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return NSB_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail(NSB_ERR_OOM, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e)); }
        cap = want;
        return NSB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct nsb_handle_s {
    int device = 0;
    nsb_hparams hp{};
    int n_fft = 0, hop = 0, win = 0, lo = 0, colours = 0, prune = 0, defcfg = 0, num_mels = 0;
    int num_sms = 0;
    int F = 0;                       // num_freq = n_fft / 2 + 1
    int generic = 0;                 // n_fft != 2048: every operator runs the generic-size kernels (gen_kernels.cuh)
    GenPlan gen{}, gen_tf{};         // their plan in the librosa and the tf.contrib.signal geometry
    float2* d_wt = nullptr;          // [n_fft] exp(-2 pi i m / n_fft)
    // pageable host input: staged through a ring of page-locked slots by the calling thread (copy_h2d)
    void* bounce[8] = {}; cudaEvent_t bounce_ev[8] = {};
    DevBuf ws_frames;                // generic path: windowed frames before the overlap-add [frames][win]
    int user_tile_hops = 0, user_stream_grid = 0;
    cudaStream_t own_stream = nullptr, copy_in = nullptr, copy_out = nullptr;
    cudaStream_t chunk_stream = nullptr;   // odd chunks of a pipelined NSB_HOST Griffin-Lim call (even ones run on the call's stream)
    cudaEvent_t desc_done = nullptr, chunk_fork = nullptr, chunk_join = nullptr;
    cudaEvent_t last_done = nullptr;       // recorded on the stream of every call when it returns (CallScope)
    cudaStream_t last_stream = nullptr; bool have_last = false;
    // tables
    float2* d_tw = nullptr;
    float* d_win = nullptr;          // periodic Hann padded centrally to n_fft (librosa geometry)
    float* d_win_tf = nullptr;       // the same window at n in [0, win) (tf.contrib.signal geometry)
    float *d_rinv = nullptr, *d_rinv_tf = nullptr;   // [hop] reciprocal interior window sums of the two geometries
    int stream_sync_mode = 2;        // 2: CTA barrier per colour step of k_gl_stream (production); 0, 1, 3, +4: the event-counter variants (experiments)
    DevBuf d_trace; int trace_on = 0, trace_grid = 0;
    DevBuf d_done;                   // k_gl_iter: item counter + per-tile completion counters
    DevBuf d_done2;                  // the same for launches on chunk_stream (two chunks' iteration launches overlap)
    int wave_schedule = 1;           // NSB_HOST Griffin-Lim on long batches: wave schedule (griffin_lim_impl), 0 = plain chunk pipeline
    int overlap_chunks = 0;          // NSB_HOST chunk pipeline on two streams (a chunk's tail overlaps the next chunk's start); off: the delayed de-emphasis of the earlier chunk costs more than the tails (profiles/r1/e2e_wave_schedule.txt)
    int wide_mode = -1;              // k_gl_iter wide mode: -1 automatic (small batches), 0 off, 1 forced (tests)
    int fuse_iterations = 1;         // all Griffin-Lim iterations of a call in ONE launch (0: one launch per iteration, A/B hook)
    int stream_ctas_per_sm = 1;      // resident k_gl_stream CTAs per SM (occupancy query at creation)
    int prune_tf = 0, colours_tf = 0;
    int specialize = 1;              // NSB_OPT_SPECIALIZE: 1 = the instantiations with the default hparams' geometry as immediates when it applies (defcfg), 0 = never
    float* d_mel_w = nullptr;
    int *d_mel_lo = nullptr, *d_mel_n = nullptr, *d_mel_ptr = nullptr;
    int4* d_mel_piece = nullptr; int4* d_mel_segp = nullptr;     // the segments as 96 balanced pieces (null: no such cut exists)
    int* d_mel_seg = nullptr; float4* d_mel_coef = nullptr;      // the filters as line segments (null: not representable, sparse rows are used)
    int mel_lines = 3;               // 3 = line segments on the skewed magnitude row, long segments cut in two (production; 2 where no such cut exists), 2 = whole segments, 1 = on the plain row, 0 = sparse mel rows (A/B hooks)
    int* d_status = nullptr;
    std::vector<double> mel_dense;   // [num_mels][num_freq]
    // descriptors
    DevBuf d_desc;
    void* h_desc = nullptr;          // pinned
    size_t h_desc_cap = 0;
    // workspaces
    DevBuf ws_mag, ws_y0, ws_y1, ws_in, ws_in2, ws_out, ws_out2, ws_ep;
    // state of the last device-resident Griffin-Lim (for nsb_griffin_lim_iterate)
    struct { bool valid = false; Batch batch{}; int total_frames = 0; int tile_hops = 0; int total_tiles = 0; int total_groups = 0; int cur = 0; bool tf = false; float inv_thr = 0.f; } gl;
    std::vector<int> h_frame_off, h_tile_off, h_group_off;       // host copies of the last descriptors (chunking)
    std::vector<long long> h_samp_off;
    int host_chunks = 0;             // 0 = automatic chunking of NSB_HOST Griffin-Lim calls, n > 0 = force n chunks
    int use_generic_iter = -1;       // -1 = automatic (k_gl_stream for long batches, k_gl_iter otherwise), 0 = k_gl_stream, 1 = generic k_synth<SRC_Y>, 2 = tile kernel k_gl_iter; A-B hook: run the iterations with k_synth<SRC_Y> instead of k_gl_iter
    unsigned long long launches = 0;
    unsigned long long launches_async = 0;   // kernels launched by the worker slots of the asynchronous entry points (guarded by async->m)
    struct nsb_async_s* async = nullptr;     // worker slots of nsb_*_submit / nsb_wait, created by the first submit
    int async_slots = 3;
    std::mutex mu;
};
static void async_shutdown(nsb_handle_s* h);

static Plan make_plan(nsb_handle_s* h, bool tf = false) {
    Plan p;
    p.tw = h->d_tw; p.mel_w = h->d_mel_w; p.mel_lo = h->d_mel_lo; p.mel_n = h->d_mel_n; p.mel_ptr = h->d_mel_ptr;
    p.mel_seg = h->mel_lines ? h->d_mel_seg : nullptr; p.mel_coef = h->d_mel_coef;
    p.mel_piece = h->d_mel_piece; p.mel_segp = h->d_mel_segp;
    p.n_fft = h->n_fft; p.hop = h->hop; p.win_len = h->win; p.num_mels = h->num_mels;
    if (tf) { p.win = h->d_win_tf; p.rinv = h->d_rinv_tf; p.lo = 0; p.origin = 0; p.norm_wss = 0; p.prune = h->prune_tf; }
    else { p.win = h->d_win; p.rinv = h->d_rinv; p.lo = h->lo; p.origin = h->n_fft / 2; p.norm_wss = 1; p.prune = h->prune; }
    return p;
}

// ---- librosa.filters.mel restated (Slaney scale, area normalisation) ---------------------------
static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}
static void build_mel(int sr, int n_fft, int n_mels, std::vector<double>& dense) {
    const int F = 1 + n_fft / 2;
    dense.assign((size_t)n_mels * F, 0.0);
    std::vector<double> mel_f(n_mels + 2), fftf(F);
    const double fmax = sr / 2.0, mmin = hz_to_mel(0.0), mmax = hz_to_mel(fmax);
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(mmin + (mmax - mmin) * i / (n_mels + 1));
    for (int k = 0; k < F; ++k) fftf[k] = fmax * k / (F - 1);
    for (int i = 0; i < n_mels; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int k = 0; k < F; ++k) {
            double lower = -(mel_f[i] - fftf[k]) / fd0, upper = (mel_f[i + 2] - fftf[k]) / fd1;
            double w = std::fmax(0.0, std::fmin(lower, upper));
            dense[(size_t)i * F + k] = w * enorm;
        }
    }
}

// The same filter bank as straight lines: between band edges j and j+1 (segment j, bins [seg[j], seg[j+1])) row j rises and row
// j-1 falls, both linearly in the bin index.  coef[m] = (value one bin past the END of segment m, slope per bin counted back from
// there) of the rising side and the same of the falling side on segment m+1.  Returns false when the lines do not reproduce `dense` (then the sparse rows are used).
static bool build_mel_lines(int sr, int n_fft, int n_mels, const std::vector<double>& dense, std::vector<int>& seg, std::vector<float>& coef) {
    const int F = 1 + n_fft / 2;
    std::vector<double> mel_f(n_mels + 2), c(4 * (size_t)n_mels);
    const double fmax = sr / 2.0, mmin = hz_to_mel(0.0), mmax = hz_to_mel(fmax), df = fmax / (F - 1);
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(mmin + (mmax - mmin) * i / (n_mels + 1));
    seg.assign(n_mels + 2, F);
    for (int j = 0, k = 0; j <= n_mels; ++j) {                  // first bin at or above edge j; the last segment keeps the Nyquist bin
        while (k < F && fmax * k / (F - 1) < mel_f[j]) ++k;
        seg[j] = k;
    }
    seg[n_mels + 1] = F;
    for (int m = 0; m < n_mels; ++m) {
        const double fd0 = mel_f[m + 1] - mel_f[m], fd1 = mel_f[m + 2] - mel_f[m + 1], enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
        if (!(fd0 > 0.0) || !(fd1 > 0.0)) return false;
        c[4 * m + 0] = enorm * (fmax * seg[m] / (F - 1) - mel_f[m]) / fd0;           c[4 * m + 1] = enorm * df / fd0;
        c[4 * m + 2] = enorm * (mel_f[m + 2] - fmax * seg[m + 1] / (F - 1)) / fd1;   c[4 * m + 3] = -enorm * df / fd1;
    }
    double wmax = 0.0, err = 0.0;
    for (int m = 0; m < n_mels; ++m)
        for (int k = 0; k < F; ++k) {
            double w = 0.0;
            if (k >= seg[m] && k < seg[m + 1]) w = c[4 * m] + c[4 * m + 1] * (k - seg[m]);
            else if (k >= seg[m + 1] && k < seg[m + 2]) w = c[4 * m + 2] + c[4 * m + 3] * (k - seg[m + 1]);
            wmax = std::fmax(wmax, dense[(size_t)m * F + k]);
            err = std::fmax(err, std::fabs(w - dense[(size_t)m * F + k]));
        }
    if (!(err <= 1e-12 * wmax)) return false;
    // the kernel measures the first moment from the segment's END (sum of running sums): a + b (k - start) = (a + b len) - b (end - k)
    coef.resize(c.size());
    for (int m = 0; m < n_mels; ++m) {
        const int len0 = seg[m + 1] - seg[m], len1 = seg[m + 2] - seg[m + 1];
        coef[4 * m + 0] = (float)(c[4 * m + 0] + c[4 * m + 1] * len0); coef[4 * m + 1] = (float)(-c[4 * m + 1]);
        coef[4 * m + 2] = (float)(c[4 * m + 2] + c[4 * m + 3] * len1); coef[4 * m + 3] = (float)(-c[4 * m + 3]);
    }
    return true;
}

// The segments of build_mel_lines cut into at most 96 pieces (three per lane) so that the lanes of the moment loop run about equally
// long: the longest segments (the top mel bands: 42 bins at the default hparams, where the first 32 have 4 or 5) are cut in two.
// Pieces sorted by length, longest first: piece p goes to lane p % 32 in round p / 32.  false: no cut with <= 2 pieces per segment fits.
static bool build_mel_pieces(const std::vector<int>& seg, int n_mels, std::vector<int4>& piece, std::vector<int4>& segp) {
    const int S = n_mels + 1;                                      // segments 0 .. n_mels
    int lmax = 0;
    for (int limit = 1; limit <= 4096 && !lmax; ++limit) {         // smallest cap on the piece length that needs <= 96 pieces
        int n = 0;
        bool ok = true;
        for (int j = 0; j < S; ++j) { const int len = seg[j + 1] - seg[j]; if (len > 2 * limit) ok = false; n += len > limit ? 2 : 1; }
        if (ok && n <= 96) lmax = limit;
    }
    if (!lmax) return false;
    struct Pc { int k0, k1, end, segment; };
    std::vector<Pc> pcs;
    for (int j = 0; j < S; ++j) {
        const int k0 = seg[j], k1 = seg[j + 1], len = k1 - k0;
        if (len > lmax) { const int kc = k0 + (len + 1) / 2; pcs.push_back({k0, kc, k1, j}); pcs.push_back({kc, k1, k1, j}); }
        else pcs.push_back({k0, k1, k1, j});
    }
    std::stable_sort(pcs.begin(), pcs.end(), [](const Pc& a, const Pc& b) { return a.k1 - a.k0 > b.k1 - b.k0; });
    piece.assign(96, make_int4(0, 0, 0, 0));
    std::vector<int> first(S, 96), second(S, 96);                  // 96 = the all-zero slot
    for (size_t p = 0; p < pcs.size(); ++p) {
        const float shift = (float)(pcs[p].end - pcs[p].k1);
        int bits;
        std::memcpy(&bits, &shift, sizeof(bits));
        piece[p] = make_int4(pcs[p].k0, pcs[p].k1, bits, pcs[p].segment);
        if (first[pcs[p].segment] == 96) first[pcs[p].segment] = (int)p; else second[pcs[p].segment] = (int)p;
    }
    segp.resize(n_mels);
    for (int m = 0; m < n_mels; ++m) segp[m] = make_int4(first[m], second[m], first[m + 1], second[m + 1]);
    return true;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return NSB_OK;
}

static size_t analysis_smem() {
    return sizeof(float2) * kTwF2 + sizeof(float) * kNfft + sizeof(float2) * kScratchF2 * kWarpsPerCta + sizeof(float4) * 96 + sizeof(int4) * 192;   // ... + mel line coefficients + piece tables
}
static size_t synth_smem(int hop, int H) {
    size_t fl = 2 * kTwF2 + kNfft + hop + (size_t)H * hop;
    fl = (fl + 3) & ~(size_t)3;
    return fl * sizeof(float) + sizeof(float2) * kScratchF2 * kWarpsPerCta;
}
static size_t gl_smem(int hop, int H) { return synth_smem(hop, H) + 64 + 16; }   // + neighbour progress flags + rinv padding   // + neighbour progress flags
// k_gl_stream: twiddles | window table (the part the pruned transform touches) | rinv | ring of 8 groups | counters | scratch
static size_t stream_smem(int hop, int win, int a, int colours, int prune) {
    const int kfirst0 = (a - win >= 0) ? (a - win) / hop + 1 : -((win - a - 1) / hop + 1) + 1, klast0 = (hop - 1 + a) / hop;
    const int back = klast0 - kfirst0 < colours ? klast0 - kfirst0 : colours;
    size_t fl = 2 * kTwF2 + (prune == 0 ? kNfft : 1024) + (((size_t)(kWarpsPerCta * colours + back) * hop + 3) & ~(size_t)3) + 16;
    return fl * sizeof(float) + sizeof(float2) * kScratchF2 * kWarpsPerCta;
}
static const size_t kSmemPerCtaTwoResident = (227 * 1024) / 2 - 1024;   // two CTAs per SM, 1 KB reserved each

static int max_tile_hops(const nsb_handle_s* h, bool tf = false) {
    // k_gl_iter gives each of its 8 warps C consecutive frames starting from the tile's (virtual) first frame, so a
    // tile may span at most 8*C frame indices.  The frames meeting H hops are the multiples of hop inside an open
    // interval of length H*hop + win: at most H + C of them (H + C - 1 when win is a multiple of hop and the interval
    // ends fall on multiples of hop, the default config).  H is kept a multiple of C: then every tile's frame groups
    // start at the same residue mod C and the summation order of the overlap-add does not depend on the tiling.
    const int a = tf ? 0 : kNfft / 2 - h->lo;      // frame k's window support starts at sample k*hop - a
    const bool aligned = (h->win % h->hop == 0) && ((a - h->win) % h->hop == 0);
    int H = 8 * h->colours - h->colours + (aligned ? 1 : 0);
    H -= H % h->colours;
    if (H < h->colours) H = h->colours;
    while (H > h->colours && gl_smem(h->hop, H) > kSmemPerCtaTwoResident) H -= h->colours;
    return H;
}

extern "C" int nsb_abi_version(void) { return NSB_ABI_VERSION; }
extern "C" const char* nsb_last_error(void) { return g_err.c_str(); }

extern "C" int nsb_device_count(int* count) {
    if (!count) return fail(NSB_ERR_INVALID, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return fail(NSB_ERR_NODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *count = n;
    return NSB_OK;
}

extern "C" int nsb_device_pci_bus_id(int device, char* out, int32_t len) {
    if (!out || len < 16) return fail(NSB_ERR_INVALID, "out is null or shorter than 16 bytes");
    out[0] = 0;
#ifdef NSB_EMULATE
    return fail(NSB_ERR_NODEVICE, "no device in the emulated build");
#else
    cudaError_t e = cudaDeviceGetPCIBusId(out, len, device);
    if (e != cudaSuccess) return fail(NSB_ERR_NODEVICE, "cudaDeviceGetPCIBusId(%d): %s", device, cudaGetErrorString(e));
    return NSB_OK;
#endif
}

extern "C" int nsb_alloc_pinned(uint64_t bytes, void** out) {
    if (!out) return fail(NSB_ERR_INVALID, "out is null");
    *out = nullptr;
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return NSB_OK;
}
extern "C" int nsb_free_pinned(void* p) {
    if (p) CU(cudaFreeHost(p));
    return NSB_OK;
}

extern "C" int nsb_destroy(nsb_handle_t h) {
    if (!h) return NSB_OK;
    async_shutdown(h);                       // runs what was submitted, joins the workers, destroys their child handles
    cudaSetDevice(h->device);
    if (h->own_stream) { cudaStreamSynchronize(h->own_stream); cudaStreamDestroy(h->own_stream); }
    if (h->copy_in) cudaStreamDestroy(h->copy_in);
    if (h->copy_out) cudaStreamDestroy(h->copy_out);
    if (h->chunk_stream) { cudaStreamSynchronize(h->chunk_stream); cudaStreamDestroy(h->chunk_stream); }
    if (h->desc_done) cudaEventDestroy(h->desc_done);
    if (h->chunk_fork) cudaEventDestroy(h->chunk_fork);
    if (h->chunk_join) cudaEventDestroy(h->chunk_join);
    if (h->last_done) cudaEventDestroy(h->last_done);
    cudaFree(h->d_tw); cudaFree(h->d_win); cudaFree(h->d_win_tf); cudaFree(h->d_rinv); cudaFree(h->d_rinv_tf); cudaFree(h->d_mel_w); cudaFree(h->d_mel_lo); cudaFree(h->d_mel_n); cudaFree(h->d_mel_ptr); cudaFree(h->d_mel_seg); cudaFree(h->d_mel_coef); cudaFree(h->d_mel_piece); cudaFree(h->d_mel_segp);
    cudaFree(h->d_status); cudaFree(h->d_wt); h->ws_frames.release();
    for (int i = 0; i < 8; ++i) { if (h->bounce[i]) cudaFreeHost(h->bounce[i]); if (h->bounce_ev[i]) cudaEventDestroy(h->bounce_ev[i]); }
    if (h->h_desc) cudaFreeHost(h->h_desc);
    h->d_desc.release(); h->d_trace.release(); h->d_done.release(); h->d_done2.release(); h->ws_mag.release(); h->ws_y0.release(); h->ws_y1.release();
    h->ws_in.release(); h->ws_in2.release(); h->ws_out.release(); h->ws_out2.release(); h->ws_ep.release();
    delete h;
    return NSB_OK;
}

extern "C" int nsb_create(const nsb_hparams* hp, int device, nsb_handle_t* out) {
    if (!hp || !out) return fail(NSB_ERR_INVALID, "null argument");
    *out = nullptr;
    // _stft_parameters, audio.py:126-130 (truncating int())
    const int n_fft = (hp->num_freq - 1) * 2;
    const int hop = (int)(hp->frame_shift_ms / 1000 * hp->sample_rate);
    const int win = (int)(hp->frame_length_ms / 1000 * hp->sample_rate);
    if (hp->num_freq < 3 || n_fft > 16384) return fail(NSB_ERR_UNSUPPORTED, "num_freq=%d (n_fft=%d): n_fft must be in [4, 16384]", hp->num_freq, n_fft);
    if (win < 1 || win > n_fft) return fail(NSB_ERR_UNSUPPORTED, "win_length=%d must be in [1, n_fft]", win);
    if (hop < 1) return fail(NSB_ERR_INVALID, "hop_length=%d must be >= 1", hop);
    if (hp->num_mels < 1 || hp->num_mels > 1024) return fail(NSB_ERR_INVALID, "num_mels=%d out of range", hp->num_mels);
    if (hp->min_level_db == 0.0) return fail(NSB_ERR_INVALID, "min_level_db must be non-zero");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(NSB_ERR_NODEVICE, "no CUDA device (%s); this library has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
    if (device < 0 || device >= ndev) return fail(NSB_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(NSB_ERR_NODEVICE, "device %d is sm_%d%d; the kernels are built for sm_100a only", device, prop.major, prop.minor);
    CU(cudaSetDevice(device));

    nsb_handle_s* h = new nsb_handle_s();
    h->device = device; h->hp = *hp; h->n_fft = n_fft; h->hop = hop; h->win = win; h->lo = (n_fft - win) / 2;
    h->colours = (win + hop - 1) / hop;
    h->F = n_fft / 2 + 1;
    h->generic = (n_fft != kNfft) ? 1 : 0;
    h->prune = (!h->generic && h->lo >= 512 && h->lo + win <= 1536) ? 1 : 0;
    h->num_mels = hp->num_mels;
    h->num_sms = prop.multiProcessorCount;
    int rc = NSB_OK;
    auto bail = [&](int code) { nsb_destroy(h); return code; };
#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(NSB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); return bail(NSB_ERR_CUDA); } } while (0)
    CUB(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&h->chunk_stream, cudaStreamNonBlocking));
    CUB(cudaEventCreateWithFlags(&h->desc_done, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&h->chunk_fork, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&h->chunk_join, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&h->last_done, cudaEventDisableTiming));
    // twiddles w2048^(j*l), j = 1..31, l = 0..31, rounded from double
    {
        std::vector<float2> tw(kTwF2);
        for (int j = 1; j < 32; ++j)
            for (int l = 0; l < 32; ++l) {
                double a = -2.0 * M_PI * (double)(j * l) / (double)kNfft;
                tw[(j - 1) * 32 + l] = make_float2((float)std::cos(a), (float)std::sin(a));
            }
        CUB(cudaMalloc(&h->d_tw, sizeof(float2) * kTwF2));
        CUB(cudaMemcpy(h->d_tw, tw.data(), sizeof(float2) * kTwF2, cudaMemcpyHostToDevice));
    }
    // periodic Hann (scipy.signal.get_window('hann', win, fftbins=True)) padded centrally (librosa.util.pad_center)
    {
        std::vector<float> w(n_fft, 0.f);
        for (int i = 0; i < win; ++i) w[h->lo + i] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * i / win));
        CUB(cudaMalloc(&h->d_win, sizeof(float) * n_fft));
        CUB(cudaMemcpy(h->d_win, w.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
        // tf.contrib.signal.stft / inverse_stft: the same periodic Hann on the first `win` samples, zero-padded at the END
        std::vector<float> wt(n_fft, 0.f);
        for (int i = 0; i < win; ++i) wt[i] = w[h->lo + i];
        CUB(cudaMalloc(&h->d_win_tf, sizeof(float) * n_fft));
        CUB(cudaMemcpy(h->d_win_tf, wt.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
        h->prune_tf = (!h->generic && win <= 1024) ? 2 : 0;
        h->colours_tf = h->colours;
        // reciprocal of the interior window sum per offset inside a hop, in the arithmetic the kernels use
        // (float fmaf chain over the covering frames, IEEE division), times 1/n_fft of the unnormalised inverse FFT
        const int GHs = h->colours * hop;                    // k_gl_stream indexes the table by the sample inside a group of C hops
        std::vector<float> ri(GHs), rt(GHs);
        for (int geo = 0; geo < 2; ++geo) {
            const int lo_g = geo ? 0 : h->lo, a = geo ? 0 : n_fft / 2 - h->lo;
            const std::vector<float>& wg = geo ? wt : w;
            for (int j = 0; j < hop; ++j) {
                float sum = 0.f;
                for (int idx = (j + a) % hop; idx < win; idx += hop) sum = std::fmaf(wg[lo_g + idx], wg[lo_g + idx], sum);
                const float v = ((geo == 0 && sum > 1.17549435e-38f) ? 1.0f / sum : 1.0f) * (1.0f / (float)n_fft);
                for (int c = 0; c < h->colours; ++c) (geo ? rt : ri)[c * hop + j] = v;
            }
        }
        CUB(cudaMalloc(&h->d_rinv, sizeof(float) * GHs));
        CUB(cudaMemcpy(h->d_rinv, ri.data(), sizeof(float) * GHs, cudaMemcpyHostToDevice));
        CUB(cudaMalloc(&h->d_rinv_tf, sizeof(float) * GHs));
        CUB(cudaMemcpy(h->d_rinv_tf, rt.data(), sizeof(float) * GHs, cudaMemcpyHostToDevice));
    }
    // sparse mel rows
    {
        build_mel(hp->sample_rate, n_fft, hp->num_mels, h->mel_dense);
        std::vector<float> wts; std::vector<int> lo(hp->num_mels), n(hp->num_mels), ptr(hp->num_mels);
        for (int m = 0; m < hp->num_mels; ++m) {
            int first = -1, last = -1;
            for (int k = 0; k < h->F; ++k) if (h->mel_dense[(size_t)m * h->F + k] != 0.0) { if (first < 0) first = k; last = k; }
            if (first < 0) { first = 0; last = -1; }
            lo[m] = first; n[m] = last - first + 1; ptr[m] = (int)wts.size();
            for (int k = first; k <= last; ++k) wts.push_back((float)h->mel_dense[(size_t)m * h->F + k]);
        }
        if (wts.empty()) wts.push_back(0.f);
        CUB(cudaMalloc(&h->d_mel_w, sizeof(float) * wts.size()));
        CUB(cudaMemcpy(h->d_mel_w, wts.data(), sizeof(float) * wts.size(), cudaMemcpyHostToDevice));
        CUB(cudaMalloc(&h->d_mel_lo, sizeof(int) * lo.size()));
        CUB(cudaMemcpy(h->d_mel_lo, lo.data(), sizeof(int) * lo.size(), cudaMemcpyHostToDevice));
        CUB(cudaMalloc(&h->d_mel_n, sizeof(int) * n.size()));
        CUB(cudaMemcpy(h->d_mel_n, n.data(), sizeof(int) * n.size(), cudaMemcpyHostToDevice));
        CUB(cudaMalloc(&h->d_mel_ptr, sizeof(int) * ptr.size()));
        CUB(cudaMemcpy(h->d_mel_ptr, ptr.data(), sizeof(int) * ptr.size(), cudaMemcpyHostToDevice));
        std::vector<int> seg; std::vector<float> coef;
        // (the two moments per segment live behind the skewed magnitude row in the warp's scratch tile, room for 95 mel bands; lane 0's exchange area after them)
        if (!h->generic && hp->num_mels <= 95 && build_mel_lines(hp->sample_rate, n_fft, hp->num_mels, h->mel_dense, seg, coef)) {
            CUB(cudaMalloc(&h->d_mel_seg, sizeof(int) * seg.size()));
            CUB(cudaMemcpy(h->d_mel_seg, seg.data(), sizeof(int) * seg.size(), cudaMemcpyHostToDevice));
            CUB(cudaMalloc(&h->d_mel_coef, sizeof(float) * coef.size()));
            CUB(cudaMemcpy(h->d_mel_coef, coef.data(), sizeof(float) * coef.size(), cudaMemcpyHostToDevice));
            std::vector<int4> piece, segp;
            if (build_mel_pieces(seg, hp->num_mels, piece, segp)) {
                CUB(cudaMalloc(&h->d_mel_piece, sizeof(int4) * piece.size()));
                CUB(cudaMemcpy(h->d_mel_piece, piece.data(), sizeof(int4) * piece.size(), cudaMemcpyHostToDevice));
                CUB(cudaMalloc(&h->d_mel_segp, sizeof(int4) * segp.size()));
                CUB(cudaMemcpy(h->d_mel_segp, segp.data(), sizeof(int4) * segp.size(), cudaMemcpyHostToDevice));
            }
        }
    }
    CUB(cudaMalloc(&h->d_status, sizeof(int)));
    CUB(cudaMemset(h->d_status, 0, sizeof(int)));
    // opt in to the large dynamic shared memory of every instantiation
    // The opt-in limit is a per-FUNCTION attribute shared by every handle of the process: always raise it to the
    // device maximum (a handle with a smaller hop must not lower it under another handle's launches).
    const size_t as = prop.sharedMemPerBlockOptin, ss = prop.sharedMemPerBlockOptin, gs = prop.sharedMemPerBlockOptin;
#define SET(k, b) do { rc = set_smem(k, b); if (rc) return bail(rc); } while (0)
    h->defcfg = (!h->generic && hop == 250 && win == 1000 && h->lo == 524) ? 1 : 0;
    if (h->generic) {
        // mixed-radix plan of the complex FFT of length M = n_fft / 2: 8s, 4s, 2s, then the odd prime factors
        GenPlan& G = h->gen;
        G.n_fft = n_fft; G.M = n_fft / 2; G.F = h->F; G.n_stages = 0;
        int m = G.M;
        auto take = [&](int r) { while (m % r == 0 && G.n_stages < kGenMaxStages) { G.radix[G.n_stages++] = r; m /= r; } };
        take(8); take(4); take(2);
        for (int pfac = 3; m > 1; pfac += 2) take(pfac);
        if (m != 1) return bail(fail(NSB_ERR_UNSUPPORTED, "n_fft=%d has too many prime factors", n_fft));
        std::vector<float2> wt(n_fft);
        for (int i = 0; i < n_fft; ++i) { const double a = -2.0 * M_PI * (double)i / (double)n_fft; wt[i] = make_float2((float)std::cos(a), (float)std::sin(a)); }
        CUB(cudaMalloc(&h->d_wt, sizeof(float2) * n_fft));
        CUB(cudaMemcpy(h->d_wt, wt.data(), sizeof(float2) * n_fft, cudaMemcpyHostToDevice));
        G.wt = h->d_wt; G.hop = hop; G.win_len = win; G.num_mels = hp->num_mels;
        G.wt_in_smem = ((size_t)(3 * G.M + 1 + n_fft) * sizeof(float2) <= 100 * 1024) ? 1 : 0;          // (up to n_fft = 4096: 81 KB, two CTAs per SM)
        G.mel_w = h->d_mel_w; G.mel_lo = h->d_mel_lo; G.mel_n = h->d_mel_n; G.mel_ptr = h->d_mel_ptr;
        h->gen_tf = G;
        G.win = h->d_win; G.lo = h->lo; G.origin = n_fft / 2; G.norm_wss = 1;
        h->gen_tf.win = h->d_win_tf; h->gen_tf.lo = 0; h->gen_tf.origin = 0; h->gen_tf.norm_wss = 0;
        SET(k_gen_analysis, as); SET(k_gen_synth, as);
        if ((size_t)(3 * G.M + 1) * sizeof(float2) > prop.sharedMemPerBlockOptin)
            return bail(fail(NSB_ERR_UNSUPPORTED, "n_fft=%d needs more shared memory than the device has", n_fft));
        *out = h;
        return NSB_OK;
    }
    SET((k_analysis<ANALYSIS_COMPLEX, false, 0>), as); SET((k_analysis<ANALYSIS_COMPLEX, false, 1>), as); SET((k_analysis<ANALYSIS_COMPLEX, false, 2>), as);
    SET((k_analysis<ANALYSIS_COMPLEX, true, 0>), as);  SET((k_analysis<ANALYSIS_COMPLEX, true, 1>), as);
    SET((k_analysis<ANALYSIS_FEATURES, true, 0>), as); SET((k_analysis<ANALYSIS_FEATURES, true, 1>), as);
    SET((k_synth<SRC_Y, 0>), ss);        SET((k_synth<SRC_Y, 1>), ss);
    SET((k_synth<SRC_SPEC, 0>), ss);     SET((k_synth<SRC_SPEC, 1>), ss);     SET((k_synth<SRC_SPEC, 2>), ss);
    SET((k_synth<SRC_MAGPHASE, 0>), ss); SET((k_synth<SRC_MAGPHASE, 1>), ss);
    SET((k_synth<SRC_MAGRAND, 0>), ss);  SET((k_synth<SRC_MAGRAND, 1>), ss);
    SET((k_synth<SRC_MAGZERO, 0>), ss);  SET((k_synth<SRC_MAGZERO, 2>), ss);
    SET((k_gl_iter<1, true, false>), gs); SET((k_gl_iter<1, false, false>), gs); SET((k_gl_iter<0, false, false>), gs);
    SET((k_gl_iter<2, false, true>), gs); SET((k_gl_iter<0, false, true>), gs); SET((k_gl_iter<2, true, true>), gs);
    SET((k_gl_stream<1, true, false, true>), gs); SET((k_gl_stream<1, false, false, true>), gs); SET((k_gl_stream<0, false, false, true>), gs);
    SET((k_gl_stream<2, false, true, true>), gs); SET((k_gl_stream<0, false, true, true>), gs); SET((k_gl_stream<2, true, true, true>), gs);
    SET((k_gl_stream<1, false, false, false>), gs); SET((k_gl_stream<0, false, false, false>), gs);
    SET((k_gl_stream<2, false, true, false>), gs); SET((k_gl_stream<0, false, true, false>), gs);
    {
        // two CTAs per SM is what the shared-memory layout of k_gl_stream is sized for; ask the runtime instead of trusting the sum
        int nb = 0;
        const size_t sm = stream_smem(hop, win, kNfft / 2 - h->lo, h->colours, h->prune);
        cudaError_t oe = h->defcfg ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gl_stream<1, true, false, true>, kThreads, sm)
                       : h->prune == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gl_stream<1, false, false, true>, kThreads, sm)
                                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gl_stream<0, false, false, true>, kThreads, sm);
        if (oe != cudaSuccess) { cudaGetLastError(); nb = 1; }
        h->stream_ctas_per_sm = nb < 1 ? 1 : (nb > 2 ? 2 : nb);
    }
#undef SET
#undef CUB
    *out = h;
    return NSB_OK;
}

extern "C" int nsb_stft_parameters(nsb_handle_t h, int32_t* n_fft, int32_t* hop, int32_t* win) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (n_fft) *n_fft = h->n_fft;
    if (hop) *hop = h->hop;
    if (win) *win = h->win;
    return NSB_OK;
}
extern "C" int64_t nsb_num_frames(nsb_handle_t h, int64_t n) { return h ? 1 + n / h->hop : -1; }
extern "C" int64_t nsb_num_samples(nsb_handle_t h, int64_t T) { return h ? (int64_t)h->hop * (T - 1) : -1; }
extern "C" uint64_t nsb_kernel_launches(nsb_handle_t h) { return h ? h->launches + h->launches_async : 0; }
extern "C" int nsb_set_host_chunks(nsb_handle_t h, int32_t n) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (n < 0 || n > 64) return fail(NSB_ERR_INVALID, "host_chunks %d outside [0,64]", n);
    h->host_chunks = n;
    return NSB_OK;
}
extern "C" int nsb_set_generic_iteration(nsb_handle_t h, int32_t on) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    h->use_generic_iter = on < 0 ? -1 : on;     // < 0: the library's choice
    return NSB_OK;
}
// profiling hook: per-CTA (SM id, start ns, end ns) of the LAST k_gl_stream launch; returns the number of CTAs written
extern "C" int nsb_stream_trace(nsb_handle_t h, int32_t enable, uint64_t* out, int32_t max_ctas) {
    if (!h) return -1;
    std::lock_guard<std::mutex> lk(h->mu);
    cudaSetDevice(h->device);
    h->trace_on = enable;
    if (!out || !h->d_trace.p || h->trace_grid <= 0) return 0;
    const int n = h->trace_grid < max_ctas ? h->trace_grid : max_ctas;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpy(out, h->d_trace.p, sizeof(uint64_t) * 3 * n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return n;
}
extern "C" int nsb_set_stream_grid(nsb_handle_t h, int32_t n) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (n < 0) return fail(NSB_ERR_INVALID, "stream grid %d < 0", n);
    h->user_stream_grid = n;
    return NSB_OK;
}
// A/B switches of the Griffin-Lim iteration kernels (profiles/sweep.py, tests): see the enum in the header
extern "C" int nsb_set_option(nsb_handle_t h, int32_t key, int32_t value) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    switch (key) {
        case NSB_OPT_STREAM_SYNC_MODE: if (value < 0 || (value > 7 && value != 10)) return fail(NSB_ERR_INVALID, "sync mode %d outside [0,7] and not 10", value); h->stream_sync_mode = value; return NSB_OK;
        case NSB_OPT_FUSE_ITERATIONS: h->fuse_iterations = value != 0; return NSB_OK;
        case NSB_OPT_WIDE_MODE: if (value < -1 || value > 1) return fail(NSB_ERR_INVALID, "wide mode %d outside [-1,1]", value); h->wide_mode = value; return NSB_OK;
        case NSB_OPT_OVERLAP_CHUNKS: h->overlap_chunks = value != 0; return NSB_OK;
        case NSB_OPT_WAVE_SCHEDULE: h->wave_schedule = value != 0; return NSB_OK;
        case NSB_OPT_SPECIALIZE: h->specialize = value != 0; return NSB_OK;
        case NSB_OPT_MEL_LINES: h->mel_lines = value < 0 ? 0 : (value > 3 ? 3 : value); return NSB_OK;
        default: return fail(NSB_ERR_INVALID, "unknown option %d", key);
    }
}
extern "C" int nsb_set_tile_hops(nsb_handle_t h, int32_t t) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (t < 0 || t > max_tile_hops(h)) return fail(NSB_ERR_INVALID, "tile_hops %d outside [0,%d]", t, max_tile_hops(h));
    h->user_tile_hops = t;
    return NSB_OK;
}
extern "C" int nsb_mel_basis(nsb_handle_t h, double* out) {
    if (!h || !out) return fail(NSB_ERR_INVALID, "null argument");
    memcpy(out, h->mel_dense.data(), sizeof(double) * h->mel_dense.size());
    return NSB_OK;
}

// NSB_DEVICE calls are stream-ordered on exactly the cudaStream_t given (NULL = CUDA's default stream, as in the
// runtime API), so that they compose with the caller's own work and events.  NSB_HOST calls are synchronous; with
// NULL they run on the handle's private non-blocking stream so that host threads with separate handles overlap.
static cudaStream_t pick_stream(nsb_handle_s* h, void* stream, int space = NSB_DEVICE) {
    if (stream) return (cudaStream_t)stream;
    return space == NSB_HOST ? h->own_stream : (cudaStream_t)0;
}

// The device state of a handle (descriptors, scheduling counters, workspaces, status flag) is shared by all of its calls, and
// an NSB_DEVICE call returns while its kernels still run.  One stream per handle is the rule; when a call nevertheless arrives
// on ANOTHER stream (or an NSB_HOST call, which uses the handle's private streams, follows an NSB_DEVICE call) its streams
// first wait for the event the previous call recorded when it returned - the handle mutex only orders the host side.
struct CallScope {
    nsb_handle_s* h; cudaStream_t st;
    CallScope(nsb_handle_s* h_, cudaStream_t st_, int space) : h(h_), st(st_) {
        if (h->have_last) {
            if (h->last_stream != st) cudaStreamWaitEvent(st, h->last_done, 0);
            if (space == NSB_HOST) cudaStreamWaitEvent(h->copy_in, h->last_done, 0);
        }
    }
    ~CallScope() { if (cudaEventRecord(h->last_done, st) == cudaSuccess) { h->last_stream = st; h->have_last = true; } }
};

extern "C" int nsb_synchronize(nsb_handle_t h, void* stream) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    if (stream) CU(cudaStreamSynchronize((cudaStream_t)stream));
    else { CU(cudaStreamSynchronize(h->own_stream)); CU(cudaStreamSynchronize((cudaStream_t)0)); }
    return NSB_OK;
}

static int read_status(nsb_handle_s* h, cudaStream_t st) {
    int flag = 0;
    CU(cudaMemcpyAsync(&flag, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (flag) {
        CU(cudaMemsetAsync(h->d_status, 0, sizeof(int), st));
        return fail(NSB_ERR_NONFINITE, "Audio buffer is not finite everywhere");
    }
    return NSB_OK;
}
extern "C" int nsb_check_status(nsb_handle_t h, void* stream) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    CallScope scope(h, pick_stream(h, stream), NSB_DEVICE);
    return read_status(h, scope.st);
}

// ---------------------------------------------------------------------------------------------
// ragged-batch descriptors: frame_off[B+1] (int), tile_off[B+1] (int), samp_off[B+1] (int64)
// ---------------------------------------------------------------------------------------------
struct Desc {
    Batch dev;
    long long total_samples = 0;
    int total_frames = 0;
    int total_tiles = 0;
    int total_groups = 0;
};

static int upload_desc(nsb_handle_s* h, cudaStream_t st, const std::vector<int>& frames, const std::vector<long long>& samples,
                       int tile_hops, Desc* d, const int64_t* rows = nullptr) {
    const int B = (int)frames.size();
    h->gl.valid = false;            // the Griffin-Lim state of nsb_griffin_lim_iterate points into the descriptor buffer rewritten below
    long long n_tiles = 0;
    if (tile_hops > 0) for (int b = 0; b < B; ++b) { long long hops = (samples[b] + h->hop - 1) / h->hop; n_tiles += (hops + tile_hops - 1) / tile_hops; }
    if (n_tiles > 2000000000LL) return fail(NSB_ERR_INVALID, "batch too large: more than 2e9 tiles");
    const size_t n_int = 2 * (size_t)(B + 1);
    const size_t off_samp = (n_int * sizeof(int) + 7) & ~(size_t)7;
    const size_t off_tutt = off_samp + (size_t)(B + 1) * sizeof(long long);
    const size_t off_goff = off_tutt + (size_t)(n_tiles > 0 ? n_tiles : 1) * sizeof(int);
    const size_t off_rows = (off_goff + (size_t)(B + 1) * sizeof(int) + 7) & ~(size_t)7;
    const size_t bytes = off_rows + (rows ? (size_t)B * sizeof(long long) : 0);
    if (bytes > h->h_desc_cap) {
        if (h->h_desc) { CU(cudaEventSynchronize(h->desc_done)); cudaFreeHost(h->h_desc); h->h_desc = nullptr; h->h_desc_cap = 0; }
        CU(cudaHostAlloc(&h->h_desc, bytes * 2, cudaHostAllocDefault));
        h->h_desc_cap = bytes * 2;
    } else {
        CU(cudaEventSynchronize(h->desc_done));   // the previous upload must have left the pinned buffer
    }
    int rc = h->d_desc.reserve(bytes);
    if (rc) return rc;
    char* hb = reinterpret_cast<char*>(h->h_desc);
    int* fo = reinterpret_cast<int*>(hb);
    int* to = fo + (B + 1);
    long long* so = reinterpret_cast<long long*>(hb + off_samp);
    int* tu = reinterpret_cast<int*>(hb + off_tutt);
    int* go = reinterpret_cast<int*>(hb + off_goff);
    fo[0] = 0; to[0] = 0; so[0] = 0; go[0] = 0;
    if (rows) { long long* ro = reinterpret_cast<long long*>(hb + off_rows); for (int b = 0; b < B; ++b) ro[b] = rows[b]; }
    for (int b = 0; b < B; ++b) {
        long long nf = (long long)fo[b] + frames[b];
        if (nf > 2000000000LL) return fail(NSB_ERR_INVALID, "batch too large: more than 2e9 frames");
        fo[b + 1] = (int)nf;
        int tiles = 0;
        if (tile_hops > 0) { long long hops = (samples[b] + h->hop - 1) / h->hop; tiles = (int)((hops + tile_hops - 1) / tile_hops); }
        to[b + 1] = to[b] + tiles;
        for (int t = to[b]; t < to[b + 1]; ++t) tu[t] = b;       // tile -> utterance (saves a binary search per tile on the GPU)
        so[b + 1] = so[b] + samples[b];
        const long long hops_b = (samples[b] + h->hop - 1) / h->hop;
        go[b + 1] = go[b] + (int)((hops_b + h->colours - 1) / h->colours);    // <= frames: cannot overflow after the check above
    }
    h->h_frame_off.assign(fo, fo + B + 1); h->h_tile_off.assign(to, to + B + 1); h->h_samp_off.assign(so, so + B + 1);
    h->h_group_off.assign(go, go + B + 1);
    CU(cudaMemcpyAsync(h->d_desc.p, h->h_desc, bytes, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(h->desc_done, st));
    const char* db = reinterpret_cast<const char*>(h->d_desc.p);
    d->dev.frame_off = reinterpret_cast<const int*>(db);
    d->dev.tile_off = d->dev.frame_off + (B + 1);
    d->dev.samp_off = reinterpret_cast<const long long*>(db + off_samp);
    d->dev.tile_utt = reinterpret_cast<const int*>(db + off_tutt);
    d->dev.group_off = reinterpret_cast<const int*>(db + off_goff);
    d->dev.row_off = rows ? reinterpret_cast<const long long*>(db + off_rows) : nullptr;
    d->dev.batch = B;
    d->dev.frame_base = 0; d->dev.tile_base = 0; d->dev.utt_base = 0; d->dev.group_base = 0;
    d->total_groups = go[B];
    d->total_frames = fo[B];
    d->total_tiles = to[B];
    d->total_samples = so[B];
    return NSB_OK;
}

static int check_launch(nsb_handle_s* h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(NSB_ERR_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
    ++h->launches;
    return NSB_OK;
}

// Host -> device copy of a caller's buffer.  cudaMemcpyAsync from PAGEABLE memory is staged by the driver at a fifth of the PCIe
// rate (11.6 against 55 GB/s here, profiles/r2/host_transfer.txt) and blocks the caller meanwhile - and a drop-in caller hands plain
// numpy arrays.  So large pageable sources go through the handle's own page-locked slots: four short-lived threads (one memcpy
// stream moves 14.7 GB/s, PCIe needs four) each copy every fourth 4 MB piece into one of their two slots and queue its DMA; a
// slot is reused when its DMA is done.  The pieces reach the stream in any order; the caller's event behind the call orders them.
// Page-locked sources are copied directly.
constexpr size_t kBounceBytes = 4u << 20;
constexpr int kBounceThreads = 4;
static bool host_is_pageable(const void* p) {
#ifdef NSB_EMULATE
    (void)p;
    return false;
#else
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
#endif
}
static cudaError_t copy_h2d(nsb_handle_s* h, void* dst, const void* src, size_t bytes, cudaStream_t st, bool pageable) {
    if (!pageable || bytes < (8u << 20)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    for (int i = 0; i < 2 * kBounceThreads; ++i)
        if (!h->bounce[i]) {
            cudaError_t e;
            if ((e = cudaHostAlloc(&h->bounce[i], kBounceBytes, cudaHostAllocDefault)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&h->bounce_ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(h->bounce_ev[i], st)) != cudaSuccess) return e;
        }
    // one staging at a time per process: when several calls arrive together (the first batches of an asynchronous stream) their
    // memcpy threads would only share the cores and the memory bus, and the first batch - the one the GPU waits for - would be
    // ready last
    static std::mutex staging_mu;
    std::lock_guard<std::mutex> staging(staging_mu);
    const size_t pieces = (bytes + kBounceBytes - 1) / kBounceBytes;
    cudaError_t errs[kBounceThreads];
    auto work = [&](int t) {
        cudaError_t e = cudaSetDevice(h->device);
        unsigned turn = 0;
        for (size_t i = (size_t)t; i < pieces && e == cudaSuccess; i += kBounceThreads, ++turn) {
            const int slot = 2 * t + (int)(turn & 1u);
            if ((e = cudaEventSynchronize(h->bounce_ev[slot])) != cudaSuccess) break;
            const size_t off = i * kBounceBytes, n = bytes - off < kBounceBytes ? bytes - off : kBounceBytes;
            memcpy(h->bounce[slot], static_cast<const char*>(src) + off, n);
            if ((e = cudaMemcpyAsync(static_cast<char*>(dst) + off, h->bounce[slot], n, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
            e = cudaEventRecord(h->bounce_ev[slot], st);
        }
        errs[t] = e;
    };
    std::thread th[kBounceThreads - 1];
    for (int t = 1; t < kBounceThreads; ++t) th[t - 1] = std::thread(work, t);
    work(0);
    for (int t = 1; t < kBounceThreads; ++t) th[t - 1].join();
    for (int t = 0; t < kBounceThreads; ++t) if (errs[t] != cudaSuccess) return errs[t];
    return cudaSuccess;
}

static int grid_1d(long long n, int threads, int max_blocks) {
    long long g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}

// ---- generic-size path (gen_kernels.cuh): launch helpers ----
static size_t gen_smem_analysis(const GenPlan& G) { return (size_t)(2 * G.M + (G.wt_in_smem ? G.n_fft : 0)) * sizeof(float2); }
static size_t gen_smem_synth(const GenPlan& G) { return (size_t)(3 * G.M + 1 + (G.wt_in_smem ? G.n_fft : 0)) * sizeof(float2); }
// threads per frame: a quarter of the complex length (a radix-8 pass has M / 8 butterflies), 64 to 256
static int gen_threads(const GenPlan& G) { int t = G.M / 4; t = (t + 31) & ~31; return t < 64 ? 64 : (t > kGenThreads ? kGenThreads : t); }
static int gen_grid(const nsb_handle_s* h, long long frames) {
    const long long cap = 16LL * h->num_sms;
    return (int)(frames < 1 ? 1 : (frames < cap ? frames : cap));
}
// frames of the (sub-)batch B from `src` -> windowed inverse transforms -> overlap-add into y_out (packed samples of the batch)
static int gen_synthesize(nsb_handle_s* h, const GenPlan& G, const Batch& B, int frames, long long total_samples, int src, const float* y_in,
                          const float* mag, const float2* spec, int spec_bin_major, int tf_renorm, unsigned long long seed, float* y_out, cudaStream_t st) {
    int rc = h->ws_frames.reserve(sizeof(float) * (size_t)frames * G.win_len);
    if (rc) return rc;
    GenSynthParams S{};
    S.plan = G; S.batch = B; S.src = src; S.y_in = y_in; S.mag = mag; S.spec = spec; S.spec_bin_major = spec_bin_major;
    S.frames_out = reinterpret_cast<float*>(h->ws_frames.p) - (size_t)B.frame_base * G.win_len;     // indexed by the GLOBAL frame number
    S.total_frames = frames; S.tf_renorm = tf_renorm; S.seed = seed; S.status = h->d_status;
    NSB_LAUNCH(k_gen_synth, gen_grid(h, frames), gen_threads(G), gen_smem_synth(G), st, S);
    if ((rc = check_launch(h, "k_gen_synth"))) return rc;
    GenOlaParams O{};
    O.plan = G; O.batch = B; O.frames = S.frames_out; O.y_out = y_out; O.total_samples = total_samples;
    NSB_LAUNCH(k_gen_ola, grid_1d(total_samples, 256, 16 * h->num_sms), 256, 0, st, O);
    return check_launch(h, "k_gen_ola");
}

// ---------------------------------------------------------------------------------------------
// analysis entry points
// ---------------------------------------------------------------------------------------------
static int run_analysis(nsb_handle_s* h, int mode, bool preemph, const float* wav, const int64_t* n_samples, int batch,
                        float* out_complex, float* lin_out, float* mel_out, int space, void* stream, bool tf = false, int rows_per_utt = 0,
                        const int64_t* row_off = nullptr, int64_t total_rows = 0) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (!wav || !n_samples || batch < 1) return fail(NSB_ERR_INVALID, "null/empty input");
    if (mode == ANALYSIS_COMPLEX && !out_complex) return fail(NSB_ERR_INVALID, "out_complex is null");
    if (mode == ANALYSIS_FEATURES && !lin_out && !mel_out) return fail(NSB_ERR_INVALID, "both outputs are null");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = pick_stream(h, stream, space);
    CallScope scope(h, st, space);
    std::vector<int> frames(batch);
    std::vector<long long> samples(batch);
    for (int b = 0; b < batch; ++b) {
        if (n_samples[b] < 1) return fail(NSB_ERR_INVALID, "utterance %d is empty", b);
        if (tf && n_samples[b] < h->win) return fail(NSB_ERR_INVALID, "utterance %d is shorter than one frame (%d samples)", b, h->win);
        samples[b] = n_samples[b];
        frames[b] = tf ? (int)(1 + (n_samples[b] - h->win) / h->hop) : (int)(1 + n_samples[b] / h->hop);
        if (rows_per_utt > 0 && frames[b] > rows_per_utt) return fail(NSB_ERR_INVALID, "utterance %d has %d frames, more than rows_per_utt = %d", b, frames[b], rows_per_utt);
        if (row_off && (row_off[b] < 0 || row_off[b] + frames[b] > total_rows))
            return fail(NSB_ERR_INVALID, "utterance %d: rows [%lld, %lld) leave the output of %lld rows", b, (long long)row_off[b], (long long)row_off[b] + frames[b], (long long)total_rows);
    }
    Desc d;
    int rc = upload_desc(h, st, frames, samples, 0, &d, row_off);
    if (rc) return rc;
    const size_t F = h->F, M = h->num_mels;
    const size_t out_rows = row_off ? (size_t)total_rows : rows_per_utt > 0 ? (size_t)rows_per_utt * batch : (size_t)d.total_frames;     // scattered, padded or packed
    const float* d_wav = wav;
    float2* d_c = reinterpret_cast<float2*>(out_complex);
    float *d_lin = lin_out, *d_mel = mel_out;
    if (space == NSB_HOST) {
        if ((rc = h->ws_in.reserve(sizeof(float) * d.total_samples))) return rc;
        d_wav = reinterpret_cast<const float*>(h->ws_in.p);
        if (mode == ANALYSIS_COMPLEX) {
            if ((rc = h->ws_out.reserve(sizeof(float2) * F * d.total_frames))) return rc;
            d_c = reinterpret_cast<float2*>(h->ws_out.p);
        } else {
            if (lin_out) { if ((rc = h->ws_out.reserve(sizeof(float) * F * out_rows))) return rc; d_lin = reinterpret_cast<float*>(h->ws_out.p); }
            if (mel_out) { if ((rc = h->ws_out2.reserve(sizeof(float) * M * out_rows))) return rc; d_mel = reinterpret_cast<float*>(h->ws_out2.p); }
        }
    }
    AnalysisParams P;
    P.plan = make_plan(h, tf); P.batch = d.dev; P.wav = d_wav; P.out_complex = d_c; P.out_lin = d_lin; P.out_mel = d_mel;
    P.total_frames = d.total_frames; P.preemph = (float)h->hp.preemphasis; P.rows_per_utt = rows_per_utt;
    P.ref_level_db = (float)h->hp.ref_level_db; P.min_level_db = (float)h->hp.min_level_db; P.status = h->d_status; P.mel_skew = h->mel_lines >= 2; P.mel_split = h->mel_lines == 3;
    {
        const double inv = 1.0 / (-h->hp.min_level_db);
        P.db_scale = (float)(20.0 * 0.30102999566398119521 * inv);      // 20 log10(2) / (-min_level_db)
        P.db_offset_lin = (float)((-h->hp.ref_level_db - h->hp.min_level_db) * inv);
        P.db_offset_mel = (float)((-h->hp.min_level_db) * inv);
    }
    const size_t smem = analysis_smem();
    const int prune = tf ? h->prune_tf : h->prune;

    // Host buffers: the results are 4.4x the input (1025 + 80 floats out per 250 samples in), so the call is bound by the copy
    // out.  Chunks of whole utterances keep both PCIe directions and the GPU busy at once: chunk c+1 is copied in (copy_in
    // stream) and transformed while chunk c is copied out (copy_out stream).  Device buffers: one chunk.
    std::vector<int> cuts(1, 0);
    if (space == NSB_HOST && batch > 1 && !row_off) {          // (scattered rows: one chunk, the output rows of a chunk are not one range)
        const size_t out_bytes = (mode == ANALYSIS_COMPLEX ? sizeof(float2) * F : sizeof(float) * ((lin_out ? F : 0) + (mel_out ? M : 0))) * (size_t)d.total_frames;
        int want = h->host_chunks > 0 ? h->host_chunks : (int)(out_bytes / (24u << 20));      // ~24 MB of results per chunk
        if (want > 32) want = 32;
        for (int c = 1; c < want; ++c) {
            const long long target = (long long)d.total_frames * c / want;
            int bb = cuts.back() + 1;
            while (bb < batch && h->h_frame_off[bb] < target) ++bb;
            if (bb < batch) cuts.push_back(bb);
        }
    }
    cuts.push_back(batch);
    const int n_chunks = (int)cuts.size() - 1;
    std::vector<cudaEvent_t> ev_in(n_chunks, nullptr), ev_done(n_chunks, nullptr);
    auto cleanup = [&]() { for (auto e : ev_in) if (e) cudaEventDestroy(e); for (auto e : ev_done) if (e) cudaEventDestroy(e); };
#define CUA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(NSB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)
    if (space == NSB_HOST) {
        for (int c = 0; c < n_chunks; ++c) {
            CUA(cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming));
            CUA(cudaEventCreateWithFlags(&ev_done[c], cudaEventDisableTiming));
        }
        const bool pageable = host_is_pageable(wav);
        for (int c = 0; c < n_chunks; ++c) {                     // all input copies are queued up front, in chunk order
            const long long s0 = h->h_samp_off[cuts[c]], s1 = h->h_samp_off[cuts[c + 1]];
            CUA(copy_h2d(h, reinterpret_cast<float*>(h->ws_in.p) + s0, wav + s0, sizeof(float) * (size_t)(s1 - s0), h->copy_in, pageable));
            CUA(cudaEventRecord(ev_in[c], h->copy_in));
        }
    }
    if (rows_per_utt > 0 || row_off) {  // the padding rows are zeros (_pad = 0, datafeeder.py:216)
        if (d_lin) CUA(cudaMemsetAsync(d_lin, 0, sizeof(float) * F * out_rows, st));
        if (d_mel) CUA(cudaMemsetAsync(d_mel, 0, sizeof(float) * M * out_rows, st));
    }
    for (int c = 0; c < n_chunks; ++c) {
        const int b0 = cuts[c], b1 = cuts[c + 1];
        P.batch = d.dev;
        P.batch.frame_off += b0; P.batch.samp_off += b0; P.batch.batch = b1 - b0;
        P.batch.frame_base = h->h_frame_off[b0]; P.batch.utt_base = b0;
        P.total_frames = h->h_frame_off[b1] - h->h_frame_off[b0];
        if (space == NSB_HOST) CUA(cudaStreamWaitEvent(st, ev_in[c], 0));
        const int grid = grid_1d(P.total_frames, kWarpsPerCta, 2 * h->num_sms);
        if (h->generic) {
            GenAnalysisParams Q{};
            Q.plan = tf ? h->gen_tf : h->gen; Q.batch = P.batch; Q.wav = P.wav; Q.total_frames = P.total_frames; Q.rows_per_utt = rows_per_utt;
            Q.out_complex = mode == ANALYSIS_COMPLEX ? d_c : nullptr; Q.out_lin = mode == ANALYSIS_COMPLEX ? nullptr : d_lin;
            Q.out_mel = mode == ANALYSIS_COMPLEX ? nullptr : d_mel;
            Q.preemph_on = preemph ? 1 : 0; Q.preemph = P.preemph;
            Q.db_scale = P.db_scale; Q.db_offset_lin = P.db_offset_lin; Q.db_offset_mel = P.db_offset_mel; Q.status = h->d_status;
            NSB_LAUNCH(k_gen_analysis, gen_grid(h, P.total_frames), gen_threads(Q.plan), gen_smem_analysis(Q.plan), st, Q);
        } else if (mode == ANALYSIS_COMPLEX) {
            if (preemph) {
                if (prune == 1) NSB_LAUNCH((k_analysis<ANALYSIS_COMPLEX, true, 1>), grid, kThreads, smem, st, P);
                else NSB_LAUNCH((k_analysis<ANALYSIS_COMPLEX, true, 0>), grid, kThreads, smem, st, P);
            } else {
                if (prune == 1) NSB_LAUNCH((k_analysis<ANALYSIS_COMPLEX, false, 1>), grid, kThreads, smem, st, P);
                else if (prune == 2) NSB_LAUNCH((k_analysis<ANALYSIS_COMPLEX, false, 2>), grid, kThreads, smem, st, P);
                else NSB_LAUNCH((k_analysis<ANALYSIS_COMPLEX, false, 0>), grid, kThreads, smem, st, P);
            }
        } else {
            if (prune == 1) NSB_LAUNCH((k_analysis<ANALYSIS_FEATURES, true, 1>), grid, kThreads, smem, st, P);
            else NSB_LAUNCH((k_analysis<ANALYSIS_FEATURES, true, 0>), grid, kThreads, smem, st, P);
        }
        if ((rc = check_launch(h, "k_analysis"))) { cleanup(); return rc; }
        if (space == NSB_HOST) {
            CUA(cudaEventRecord(ev_done[c], st));
            CUA(cudaStreamWaitEvent(h->copy_out, ev_done[c], 0));
            // rows of this chunk: packed = its frames, padded = rows_per_utt per utterance
            const size_t r0 = row_off ? 0 : rows_per_utt > 0 ? (size_t)rows_per_utt * b0 : (size_t)h->h_frame_off[b0];
            const size_t r1 = row_off ? out_rows : rows_per_utt > 0 ? (size_t)rows_per_utt * b1 : (size_t)h->h_frame_off[b1];
            if (mode == ANALYSIS_COMPLEX)
                CUA(cudaMemcpyAsync(reinterpret_cast<float2*>(out_complex) + F * r0, d_c + F * r0, sizeof(float2) * F * (r1 - r0), cudaMemcpyDeviceToHost, h->copy_out));
            else {
                if (lin_out) CUA(cudaMemcpyAsync(lin_out + F * r0, d_lin + F * r0, sizeof(float) * F * (r1 - r0), cudaMemcpyDeviceToHost, h->copy_out));
                if (mel_out) CUA(cudaMemcpyAsync(mel_out + M * r0, d_mel + M * r0, sizeof(float) * M * (r1 - r0), cudaMemcpyDeviceToHost, h->copy_out));
            }
        }
    }
    if (space == NSB_HOST) {
        CUA(cudaStreamSynchronize(h->copy_out));
        cleanup();
        return read_status(h, st);
    }
#undef CUA
    cleanup();
    return NSB_OK;
}

extern "C" int nsb_stft(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch, int32_t apply_preemphasis,
                        float* out_complex, int32_t space, void* stream) {
    return run_analysis(h, ANALYSIS_COMPLEX, apply_preemphasis != 0, wav, n_samples, batch, out_complex, nullptr, nullptr, space, stream);
}
extern "C" int nsb_stft_tf(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch, float* out_complex, int32_t space, void* stream) {
    return run_analysis(h, ANALYSIS_COMPLEX, false, wav, n_samples, batch, out_complex, nullptr, nullptr, space, stream, true);
}
extern "C" int nsb_features(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch,
                            float* lin_out, float* mel_out, int32_t space, void* stream) {
    return run_analysis(h, ANALYSIS_FEATURES, true, wav, n_samples, batch, nullptr, lin_out, mel_out, space, stream);
}

// ---------------------------------------------------------------------------------------------
// synthesis entry points
// ---------------------------------------------------------------------------------------------
// small batches: tiles of C hops with one frame per warp (k_gl_iter's wide mode) when that still gives at most four tiles per
// resident CTA - a tile is then one frame-time long instead of C, which is what a single utterance needs (demo_server.py)
static bool choose_wide(nsb_handle_s* h, const std::vector<long long>& samples) {
    if (h->wide_mode == 0 || h->user_tile_hops > 0) return false;
    if (h->wide_mode == 1) return true;
    long long tiles = 0;
    for (size_t b = 0; b < samples.size(); ++b) { long long hops = (samples[b] + h->hop - 1) / h->hop; tiles += (hops + h->colours - 1) / h->colours; }
    return tiles <= 8LL * h->num_sms;            // measured crossover ~1,300 tiles (profiles/r1/sweep_wide.txt)
}

static int choose_tile_hops(nsb_handle_s* h, const std::vector<long long>& samples, bool tf = false) {
    const int Hmax = max_tile_hops(h, tf), C = h->colours;
    if (choose_wide(h, samples)) return C;
    if (h->user_tile_hops > 0) {
        int H = h->user_tile_hops < Hmax ? h->user_tile_hops : Hmax;
        H -= H % C;                  // multiples of C only (tiling-independent summation order, see max_tile_hops)
        return H < C ? C : H;
    }
    // persistent grid of 2*#SMs CTAs; a CTA spends ~C frame-times per tile whatever H is (one frame per warp and
    // colour), so the launch takes ceil(tiles / CTAs) * C frame-times: pick the multiple of C that minimises the number
    // of waves, and among equals the largest tile (least halo recomputation)
    const long long ctas = 2LL * h->num_sms;
    int best = C;
    long long best_waves = -1;
    for (int H = C; H <= Hmax; H += C) {
        long long tiles = 0;
        for (size_t b = 0; b < samples.size(); ++b) { long long hops = (samples[b] + h->hop - 1) / h->hop; tiles += (hops + H - 1) / H; }
        const long long waves = (tiles + ctas - 1) / ctas;
        if (best_waves < 0 || waves <= best_waves) { best = H; best_waves = waves; }
    }
    return best;
}

template <int SRC>
static void launch_synth(nsb_handle_s* h, const SynthParams& P, int grid, size_t smem, cudaStream_t st) {
    if (P.plan.prune == 1) NSB_LAUNCH((k_synth<SRC, 1>), grid, kThreads, smem, st, P);
    else NSB_LAUNCH((k_synth<SRC, 0>), grid, kThreads, smem, st, P);
}
// tf geometry variants exist for the two sources the TF twin needs
static void launch_synth_tf(nsb_handle_s* h, int src, const SynthParams& P, int grid, size_t smem, cudaStream_t st) {
    if (src == SRC_SPEC) {
        if (P.plan.prune == 2) NSB_LAUNCH((k_synth<SRC_SPEC, 2>), grid, kThreads, smem, st, P);
        else NSB_LAUNCH((k_synth<SRC_SPEC, 0>), grid, kThreads, smem, st, P);
    } else {
        if (P.plan.prune == 2) NSB_LAUNCH((k_synth<SRC_MAGZERO, 2>), grid, kThreads, smem, st, P);
        else NSB_LAUNCH((k_synth<SRC_MAGZERO, 0>), grid, kThreads, smem, st, P);
    }
}

// librosa.istft gives hop*(T-1) samples (centre trimmed); tf inverse_stft gives hop*(T-1) + win
static int validate_frames(nsb_handle_s* h, const int32_t* n_frames, int batch, std::vector<int>& frames, std::vector<long long>& samples,
                           bool tf = false) {
    frames.resize(batch); samples.resize(batch);
    for (int b = 0; b < batch; ++b) {
        if (n_frames[b] < (tf ? 1 : 2))
            return fail(NSB_ERR_INVALID, "utterance %d has %d frame(s); inversion needs at least %d", b, n_frames[b], tf ? 1 : 2);
        frames[b] = n_frames[b];
        samples[b] = (long long)h->hop * (n_frames[b] - 1) + (tf ? h->win : 0);
    }
    return NSB_OK;
}

static int run_istft(nsb_handle_t h, const float* spec, int32_t layout, const int32_t* n_frames, int32_t batch,
                     float* wav_out, int32_t space, void* stream, bool tf) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (!spec || !n_frames || !wav_out || batch < 1) return fail(NSB_ERR_INVALID, "null/empty argument");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = pick_stream(h, stream, space);
    CallScope scope(h, st, space);
    std::vector<int> frames; std::vector<long long> samples;
    int rc = validate_frames(h, n_frames, batch, frames, samples, tf);
    if (rc) return rc;
    const int H = h->generic ? 0 : choose_tile_hops(h, samples, tf);
    Desc d;
    if ((rc = upload_desc(h, st, frames, samples, H, &d))) return rc;
    const float2* d_spec = reinterpret_cast<const float2*>(spec);
    float* d_out = wav_out;
    if (space == NSB_HOST) {
        if ((rc = h->ws_in.reserve(sizeof(float2) * h->F * (size_t)d.total_frames))) return rc;
        if ((rc = h->ws_out.reserve(sizeof(float) * d.total_samples))) return rc;
        CU(cudaMemcpyAsync(h->ws_in.p, spec, sizeof(float2) * h->F * (size_t)d.total_frames, cudaMemcpyHostToDevice, st));
        d_spec = reinterpret_cast<const float2*>(h->ws_in.p);
        d_out = reinterpret_cast<float*>(h->ws_out.p);
    }
    h->gl.valid = false;
    if (h->generic) {
        if ((rc = gen_synthesize(h, tf ? h->gen_tf : h->gen, d.dev, d.total_frames, d.total_samples, SRC_SPEC, nullptr, nullptr, d_spec,
                                 layout == NSB_BIN_MAJOR, 0, 0, d_out, st))) return rc;
        if (space == NSB_HOST) {
            CU(cudaMemcpyAsync(wav_out, d_out, sizeof(float) * d.total_samples, cudaMemcpyDeviceToHost, st));
            return read_status(h, st);
        }
        return NSB_OK;
    }
    SynthParams P{};
    P.plan = make_plan(h, tf); P.batch = d.dev; P.spec = d_spec; P.spec_bin_major = (layout == NSB_BIN_MAJOR);
    P.y_out = d_out; P.tile_hops = H; P.colours = h->colours; P.status = h->d_status; P.mag = nullptr;
    if (tf) launch_synth_tf(h, SRC_SPEC, P, d.total_tiles, synth_smem(h->hop, H), st);
    else launch_synth<SRC_SPEC>(h, P, d.total_tiles, synth_smem(h->hop, H), st);
    if ((rc = check_launch(h, "k_synth<SPEC>"))) return rc;
    if (space == NSB_HOST) {
        CU(cudaMemcpyAsync(wav_out, d_out, sizeof(float) * d.total_samples, cudaMemcpyDeviceToHost, st));
        return read_status(h, st);
    }
    return NSB_OK;
}
extern "C" int nsb_istft(nsb_handle_t h, const float* spec, int32_t layout, const int32_t* n_frames, int32_t batch,
                         float* wav_out, int32_t space, void* stream) {
    return run_istft(h, spec, layout, n_frames, batch, wav_out, space, stream, false);
}
extern "C" int nsb_istft_tf(nsb_handle_t h, const float* spec, int32_t layout, const int32_t* n_frames, int32_t batch,
                            float* wav_out, int32_t space, void* stream) {
    return run_istft(h, spec, layout, n_frames, batch, wav_out, space, stream, true);
}

// k_deemphasis launch geometry: segments of 16 chunks (32768 samples) per CTA once the filter's memory is short enough
static int deemph_grid(double p, long long max_len, EmphParams& E, int batch) {
    const long long chunk = (long long)kDeemphThreads * kPerThread;
    E.seg_len = 0; E.warmup = 0; E.n_seg = 1;
    const double ap = std::fabs(p);
    if (ap > 0.0 && ap < 1.0) {
        const long long w = (long long)std::ceil(std::log(1e-24) / std::log(ap));
        const long long wc = (w + chunk - 1) / chunk * chunk;
        const long long seg = 16 * chunk;
        if (wc <= seg / 4 && max_len > seg) { E.seg_len = seg; E.warmup = wc; E.n_seg = (int)((max_len + seg - 1) / seg); }
    } else if (ap == 0.0) {
        const long long seg = 16 * chunk;
        if (max_len > seg) { E.seg_len = seg; E.warmup = 0; E.n_seg = (int)((max_len + seg - 1) / seg); }
    }
    return batch * E.n_seg;
}

// `iters` Griffin-Lim iterations on the (sub-)batch B; y ping-pongs between ws_y0 / ws_y1, `cur` says which holds y
static int gl_iterations(nsb_handle_s* h, const Batch& B, int total_tiles, int total_groups, int H, int& cur, int iters, cudaStream_t st,
                         bool tf = false, float inv_thr = 0.f, DevBuf* done_buf = nullptr) {
    DevBuf& d_done = done_buf ? *done_buf : h->d_done;     // scheduling counters of this stream's launches
    float* y[2] = {reinterpret_cast<float*>(h->ws_y0.p), reinterpret_cast<float*>(h->ws_y1.p)};
    if (total_tiles <= 0) return NSB_OK;
    // hop 250 / window 1000 / 4 colours as immediates (the window sits at n = 524 in the librosa geometry, at 0 in the TF twin's)
    const bool defcfg = h->defcfg && h->specialize;
    // automatic choice: the streaming kernel wins from about 7,000 groups (28k frames) per launch on; below that the tile
    // kernel, which has less to set up per work item (measured: profiles/r1/sweep_kernels.txt)
    int which = h->use_generic_iter;
    if (which < 0) which = (2 * ((long long)total_groups + B.batch) >= 49LL * h->stream_ctas_per_sm * h->num_sms && h->stream_ctas_per_sm >= 2) ? 0 : 2;   // >= 3.5 chunks of 7 groups per CTA
    if (which == 0) {
        // the production path for long batches: streaming kernel, CTA i owns the contiguous group range [i*N/n, (i+1)*N/n)
        GlStreamParams S{};
        S.plan = make_plan(h, tf); S.batch = B; S.mag = reinterpret_cast<const float*>(h->ws_mag.p);
        S.colours = h->colours; S.total_groups = total_groups; S.status = h->d_status; S.inv_thr = inv_thr;
        S.sync_mode = h->stream_sync_mode;
        S.trace = nullptr;
        const size_t smem = stream_smem(h->hop, h->win, S.plan.origin - S.plan.lo, h->colours, S.plan.prune);
        const int ctas = h->stream_ctas_per_sm * h->num_sms;
        // the stream: every utterance's groups + one end-halo group each, cut into chunks of CH groups; a chunk and its
        // halo group fill whole rounds of the 8 warps when CH + 1 is a multiple of 8.  Long chunks waste fewer halo frames
        // (3 of 4*CH+3) but the launch needs a few chunks per CTA and iteration to keep every SM busy.
        const long long NV = (long long)total_groups + B.batch;
        int CH = 7;
        for (int c = 31; c > 7; c -= 8) if ((NV + c - 1) / c >= 2LL * ctas) { CH = c; break; }
        if (h->user_tile_hops > 0) CH = h->user_tile_hops / h->colours;                       // tests: chunks of that many groups
        if (h->user_stream_grid > 0) CH = (int)((NV + h->user_stream_grid - 1) / h->user_stream_grid);   // tests: that many chunks
        if (CH < 1) CH = 1;
        const long long n_chunks = (NV + CH - 1) / CH;
        int rc = d_done.reserve(sizeof(int) * ((size_t)n_chunks + 1));
        if (rc) return rc;
        S.item_counter = reinterpret_cast<int*>(d_done.p);
        S.done = S.item_counter + 1;
        S.ybuf[0] = y[0]; S.ybuf[1] = y[1];
        S.chunk_groups = CH;
        bool fb;
        {
            const int a = S.plan.origin - S.plan.lo, hop = h->hop, win = h->win;
            const int kfirst0 = (a - win >= 0) ? (a - win) / hop + 1 : -((win - a - 1) / hop + 1) + 1, klast0 = (hop - 1 + a) / hop;
            fb = (klast0 - kfirst0 <= h->colours - 1) && (h->stream_sync_mode & 7) == 2;
        }
        const int per_launch = h->fuse_iterations ? iters : 1;
        for (int it = 0; it < iters; it += per_launch) {
            const int n = iters - it < per_launch ? iters - it : per_launch;
            CU(cudaMemsetAsync(d_done.p, 0, sizeof(int) * ((size_t)n_chunks + 1), st));
            S.cur0 = cur; S.iters = n;
            const long long items = (long long)n * n_chunks;
            const int grid = items < ctas ? (int)items : ctas;
            if (h->trace_on) {
                if ((rc = h->d_trace.reserve(sizeof(unsigned long long) * 3 * (size_t)grid))) return rc;
                S.trace = reinterpret_cast<unsigned long long*>(h->d_trace.p);
                h->trace_grid = grid;
            }
            // the per-frame barrier variant whenever a group's hops need at most C-1 colours of the next group (always at the
            // default hparams) and no barrier experiment is selected
            if (fb) {
                if (tf) {
                    if (defcfg && S.plan.prune == 2) NSB_LAUNCH((k_gl_stream<2, true, true, true>), grid, kThreads, smem, st, S);
                    else if (S.plan.prune == 2) NSB_LAUNCH((k_gl_stream<2, false, true, true>), grid, kThreads, smem, st, S);
                    else NSB_LAUNCH((k_gl_stream<0, false, true, true>), grid, kThreads, smem, st, S);
                } else if (defcfg) NSB_LAUNCH((k_gl_stream<1, true, false, true>), grid, kThreads, smem, st, S);
                else if (h->prune == 1) NSB_LAUNCH((k_gl_stream<1, false, false, true>), grid, kThreads, smem, st, S);
                else NSB_LAUNCH((k_gl_stream<0, false, false, true>), grid, kThreads, smem, st, S);
            } else {
                if (tf) {
                    if (S.plan.prune == 2) NSB_LAUNCH((k_gl_stream<2, false, true, false>), grid, kThreads, smem, st, S);
                    else NSB_LAUNCH((k_gl_stream<0, false, true, false>), grid, kThreads, smem, st, S);
                } else if (h->prune == 1) NSB_LAUNCH((k_gl_stream<1, false, false, false>), grid, kThreads, smem, st, S);
                else NSB_LAUNCH((k_gl_stream<0, false, false, false>), grid, kThreads, smem, st, S);
            }
            if ((rc = check_launch(h, "k_gl_stream"))) return rc;
            cur ^= (n & 1);
        }
        return NSB_OK;
    }
    if (which == 1 && !tf) {
        SynthParams P{};
        P.plan = make_plan(h); P.batch = B; P.mag = reinterpret_cast<const float*>(h->ws_mag.p);
        P.tile_hops = H; P.colours = h->colours; P.status = h->d_status;
        const size_t smem = synth_smem(h->hop, H);
        for (int it = 0; it < iters; ++it) {
            P.y_in = y[cur]; P.y_out = y[cur ^ 1];
            launch_synth<SRC_Y>(h, P, total_tiles, smem, st);
            int rc = check_launch(h, "k_synth<Y>");
            if (rc) return rc;
            cur ^= 1;
        }
        return NSB_OK;
    }
    GlParams G{};
    G.plan = make_plan(h, tf); G.batch = B; G.mag = reinterpret_cast<const float*>(h->ws_mag.p);
    G.tile_hops = H; G.colours = h->colours; G.total_tiles = total_tiles; G.status = h->d_status; G.inv_thr = inv_thr;
    G.wide = (h->wide_mode != 0 && h->user_tile_hops == 0 && H == h->colours && (h->wide_mode == 1 || total_tiles <= 8 * h->num_sms)) ? 1 : 0;
    const size_t smem = gl_smem(h->hop, H);
    // one launch runs all the iterations: (iteration, tile) items from a global counter, per-tile completion counters
    int rc = d_done.reserve(sizeof(int) * ((size_t)total_tiles + 1));
    if (rc) return rc;
    G.item_counter = reinterpret_cast<int*>(d_done.p);
    G.done = G.item_counter + 1;
    G.ybuf[0] = y[0]; G.ybuf[1] = y[1];
    const int per_launch = h->fuse_iterations ? iters : 1;
    for (int it = 0; it < iters; it += per_launch) {
        const int n = iters - it < per_launch ? iters - it : per_launch;
        CU(cudaMemsetAsync(d_done.p, 0, sizeof(int) * ((size_t)total_tiles + 1), st));
        G.cur0 = cur; G.iters = n;
        const long long items = (long long)n * total_tiles;
        const int grid = items < 2LL * h->num_sms ? (int)items : 2 * h->num_sms;   // persistent: 2 CTAs per SM
        if (tf) {
            if (defcfg && G.plan.prune == 2) NSB_LAUNCH((k_gl_iter<2, true, true>), grid, kThreads, smem, st, G);
            else if (G.plan.prune == 2) NSB_LAUNCH((k_gl_iter<2, false, true>), grid, kThreads, smem, st, G);
            else NSB_LAUNCH((k_gl_iter<0, false, true>), grid, kThreads, smem, st, G);
        } else if (defcfg) NSB_LAUNCH((k_gl_iter<1, true, false>), grid, kThreads, smem, st, G);
        else if (h->prune == 1) NSB_LAUNCH((k_gl_iter<1, false, false>), grid, kThreads, smem, st, G);
        else NSB_LAUNCH((k_gl_iter<0, false, false>), grid, kThreads, smem, st, G);
        if ((rc = check_launch(h, "k_gl_iter"))) return rc;
        cur ^= (n & 1);
    }
    return NSB_OK;
}

extern "C" int nsb_griffin_lim_iterate(nsb_handle_t h, int32_t iters, void* stream) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->gl.valid) return fail(NSB_ERR_INVALID, "no device-resident Griffin-Lim state (call nsb_griffin_lim with NSB_DEVICE first)");
    CU(cudaSetDevice(h->device));
    CallScope scope(h, pick_stream(h, stream), NSB_DEVICE);
    return gl_iterations(h, h->gl.batch, h->gl.total_tiles, h->gl.total_groups, h->gl.tile_hops, h->gl.cur, iters, scope.st, h->gl.tf, h->gl.inv_thr);
}

// endpoint search fused into the Griffin-Lim pipeline (nsb_synthesize): per chunk, on the device result
struct EndpointReq {
    int64_t* out = nullptr; double threshold_db = -40.0, min_silence_sec = 0.8;
    int peak = 0;                     // save_wav's scaling (audio.py:17-19) on the trimmed waveform
    int final_dtype = NSB_F64;        // NSB_F64, or NSB_I16 (the scaled waveform cast like numpy's astype(np.int16))
};
static size_t dtype_size(int dt) { return dt == NSB_F64 ? sizeof(double) : dt == NSB_I16 ? sizeof(short) : sizeof(float); }
static void endpoint_params(const nsb_handle_s* h, double threshold_db, double min_silence_sec, EndpointParams& E) {
    E.window = (long long)(h->hp.sample_rate * min_silence_sec);            // int(sample_rate * min_silence_sec), audio.py:68
    E.hop = E.window / 4;                                                    // int(window_length / 4), audio.py:69
    if (E.hop < 1) E.hop = 1;
    E.threshold = std::pow(10.0, threshold_db * 0.05);                       // _db_to_amp, audio.py:154-155
}

static int griffin_lim_impl(nsb_handle_t h, const float* spec, int32_t layout, const int32_t* n_frames, int32_t batch,
                            const float* init_phase, uint64_t seed, int32_t iters, int32_t flags,
                            void* wav_out, int32_t out_dtype, int32_t space, void* stream, const EndpointReq* ep) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (!spec || !n_frames || !wav_out || batch < 1) return fail(NSB_ERR_INVALID, "null/empty argument");
    if (out_dtype != NSB_F32 && out_dtype != NSB_F64) return fail(NSB_ERR_INVALID, "bad out_dtype");
    if (iters < 0) iters = h->hp.griffin_lim_iters;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = pick_stream(h, stream, space);
    CallScope scope(h, st, space);
    const bool tf = (flags & NSB_GL_TF_TWIN) != 0;
    if (tf) init_phase = nullptr;            // the TF twin always starts from zero phase (audio.py:97-98)
    std::vector<int> frames; std::vector<long long> samples;
    int rc = validate_frames(h, n_frames, batch, frames, samples, tf);
    if (rc) return rc;
    const int H = h->generic ? 0 : choose_tile_hops(h, samples, tf);
    Desc d;
    if ((rc = upload_desc(h, st, frames, samples, H, &d))) return rc;
    // out_dtype is what the de-emphasis writes; the synthesis stage may end in another type (peak-normalised int16)
    const bool to_i16 = ep && ep->final_dtype == NSB_I16;
    const size_t out_elt = to_i16 ? sizeof(short) : dtype_size(out_dtype);
    if (!(flags & NSB_GL_DEEMPHASIS) && out_dtype != NSB_F32)
        return fail(NSB_ERR_INVALID, "NSB_F64 output without NSB_GL_DEEMPHASIS is not provided (_griffin_lim returns float32)");
    const size_t n_spec = (size_t)h->F * d.total_frames;
    const size_t mag_pitch = h->generic ? (size_t)h->F : (size_t)kMagPitch;
    const float* d_spec = spec;
    const float2* d_phase = reinterpret_cast<const float2*>(init_phase);
    char* d_out = reinterpret_cast<char*>(wav_out);           // de-emphasis output (out_dtype)
    char* d_final = d_out;                                      // what the caller receives
    if (to_i16) {
        if ((rc = h->ws_out.reserve(dtype_size(out_dtype) * d.total_samples))) return rc;
        d_out = reinterpret_cast<char*>(h->ws_out.p);
        if (space == NSB_HOST) { if ((rc = h->ws_out2.reserve(sizeof(short) * d.total_samples))) return rc; d_final = reinterpret_cast<char*>(h->ws_out2.p); }
    }
    if (space == NSB_HOST) {
        if ((rc = h->ws_in.reserve(sizeof(float) * n_spec))) return rc;
        d_spec = reinterpret_cast<const float*>(h->ws_in.p);
        if (init_phase) {
            if ((rc = h->ws_in2.reserve(sizeof(float2) * n_spec))) return rc;
            d_phase = reinterpret_cast<const float2*>(h->ws_in2.p);
        }
        if (!to_i16) {
            if ((rc = h->ws_out.reserve(out_elt * d.total_samples))) return rc;
            d_out = d_final = reinterpret_cast<char*>(h->ws_out.p);
        }
    }
    if ((rc = h->ws_mag.reserve(sizeof(float) * mag_pitch * (size_t)d.total_frames))) return rc;
    if ((rc = h->ws_y0.reserve(sizeof(float) * d.total_samples))) return rc;
    if ((rc = h->ws_y1.reserve(sizeof(float) * d.total_samples))) return rc;

    // Griffin-Lim is linear in S: run it on g*S with g a power of two that brings the largest possible magnitude
    // below 1 (the yaml's +100 dB floor gives S up to 1e9 whose squares would overflow fp32), undo g on output.
    double gscale = 1.0;
    if ((flags & NSB_GL_DENORMALIZE) && !h->generic) {       // (the generic kernels renormalise with a guarded square: no pre-scale needed)
        const double e0 = (h->hp.min_level_db + h->hp.ref_level_db) * 0.05 * h->hp.power, e1 = h->hp.ref_level_db * 0.05 * h->hp.power;
        const double smax = std::pow(10.0, e0 > e1 ? e0 : e1);
        gscale = std::ldexp(1.0, -(int)std::ceil(std::log2(smax)));
        if (!(gscale > 0.0) || !std::isfinite(gscale)) gscale = 1.0;
    }

    // Host buffers: cut the batch into a few chunks of whole utterances and pipeline them - chunk c+1 is copied in
    // (copy_in stream) and chunk c-1 copied out (copy_out stream) while chunk c computes.  Device buffers: one chunk.
    std::vector<int> cuts;           // utterance index where each chunk starts, plus the end
    cuts.push_back(0);
    if (space == NSB_HOST && batch > 1 && !h->generic && (h->host_chunks > 0 || n_spec * sizeof(float) > (8u << 20))) {
        // Only the first chunk's copy-in and the last chunk's copy-out are exposed, so those two chunks should be small -
        // but below ~16k frames the iteration kernels are latency-bound, and the streaming kernel (the fastest) needs
        // ~35k frames.  Automatic mode: long batches (>= 60k frames) are cut 16k | rest | 12k frames, shorter ones into up
        // to 4 equal chunks of >= 16k frames (measured: profiles/r1/e2e_chunks.txt).
        std::vector<long long> targets;
        const char* ecuts = std::getenv("NSB_CHUNK_CUTS");        // tuning hook: comma-separated frame positions of the cuts
        if (ecuts && *ecuts) {
            for (const char* q = ecuts; *q;) { char* end; long long v = std::strtoll(q, &end, 10); if (end == q) break; targets.push_back(v); q = (*end == ',') ? end + 1 : end; }
        } else if (h->host_chunks > 0) {
            for (int c = 1; c < h->host_chunks; ++c) targets.push_back((long long)d.total_frames * c / h->host_chunks);
        } else if (d.total_frames >= 60000) {
            const char* ef = std::getenv("NSB_CHUNK_FIRST"); const char* el = std::getenv("NSB_CHUNK_LAST");     // tuning hooks
            targets.push_back(ef ? std::atoll(ef) : 16000);
            targets.push_back((long long)d.total_frames - (el ? std::atoll(el) : 12000));
        } else {
            int want = d.total_frames / 16000;
            if (want > 4) want = 4;
            for (int c = 1; c < want; ++c) targets.push_back((long long)d.total_frames * c / want);
        }
        for (long long target : targets) {
            int b = cuts.back() + 1;
            while (b < batch && h->h_frame_off[b] < target) ++b;
            if (b < batch && b > cuts.back()) cuts.push_back(b);
        }
    }
    // Wave schedule for long NSB_HOST batches.  The input arrives over PCIe ~3.5x faster than Griffin-Lim consumes it, but
    // a chunk small enough to arrive quickly cannot fill the GPU for 60 iterations on its own.  So the iterations are cut into
    // `waves` launches of `wave_iters` (even) iterations; chunk g joins at wave g, and every launch runs ALL chunks that have
    // arrived and are not finished: the launches grow with the data, only the very first ones are small, and a finished chunk's
    // de-emphasis and copy-out hide behind the later waves.  Chunk g+1 is sized to arrive while wave g runs.
    int waves = 0, wave_iters = 0;
    if (space == NSB_HOST && batch > 1 && !h->generic && h->wave_schedule && h->host_chunks == 0 && !h->trace_on && h->fuse_iterations
        && !(std::getenv("NSB_CHUNK_CUTS") && *std::getenv("NSB_CHUNK_CUTS")) && d.total_frames >= 40000) {
        const int ie = iters - (iters & 1);
        const char* ew = std::getenv("NSB_WAVES");                 // tuning hooks
        // few long waves beat many short ones (every launch has its ramp and tail): 3 waves if the count divides, else 4, 5, 6, 2
        const int pref[5] = {ew ? std::atoi(ew) : 3, 4, 5, 6, 2};
        for (int K : pref) if (K >= 2 && ie % (2 * K) == 0 && ie / K >= 4) { waves = K; wave_iters = ie / K; break; }
    }
    if (waves > 0) {
        const char* ef = std::getenv("NSB_WAVE_FIRST"); const char* eg = std::getenv("NSB_WAVE_GROWTH");
        const double first = ef ? std::atof(ef) : 4000.0, growth = eg ? std::atof(eg) : 1.4;     // (measured: profiles/r1/e2e_wave_schedule.txt)
        const double in_ms_per_frame = (init_phase ? 3.0 : 1.0) * kBins * sizeof(float) / 53.0e6;     // ~53 GB/s host to device
        const double it_ms_per_frame = 4.3e-6, it_ms_floor = 0.045;                                   // measured: profiles/r1/sweep_kernels.txt
        const std::vector<int> chunk_cuts = cuts;     // the plain chunk pipeline's cuts, should the batch not split into >= 3 chunks
        cuts.assign(1, 0);
        std::vector<long long> joined;       // frames of the chunks so far
        double target = first;
        while (cuts.back() < batch) {
            int b = cuts.back() + 1;
            const long long f0 = h->h_frame_off[cuts.back()];
            while (b < batch && h->h_frame_off[b] - f0 < (long long)target) ++b;
            if (batch - b < 2 || (long long)d.total_frames - h->h_frame_off[b] < 2000 || (int)cuts.size() >= 12) b = batch;     // no crumbs at the end
            cuts.push_back(b);
            joined.push_back(h->h_frame_off[b] - f0);
            // the wave that starts now runs the last `waves` chunks; the next chunk should arrive while it runs
            long long active = 0;
            for (int g = (int)joined.size() - 1; g >= 0 && g > (int)joined.size() - 1 - waves; --g) active += joined[g];
            const double wave_ms = wave_iters * std::max(it_ms_floor, active * it_ms_per_frame);
            target = growth * wave_ms / in_ms_per_frame;
        }
        cuts.pop_back();                     // re-added below
        if ((int)cuts.size() < 3) { waves = 0; wave_iters = 0; cuts = chunk_cuts; }       // a few very long utterances: chunk pipeline
    }
    cuts.push_back(batch);
    const int n_chunks = (int)cuts.size() - 1;
    std::vector<cudaEvent_t> ev_in(n_chunks, nullptr), ev_done(n_chunks, nullptr);
    auto cleanup = [&]() { for (auto e : ev_in) if (e) cudaEventDestroy(e); for (auto e : ev_done) if (e) cudaEventDestroy(e); };
#define CUE(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(NSB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)
    if (space == NSB_HOST) {
        for (int c = 0; c < n_chunks; ++c) {
            CUE(cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming));
            CUE(cudaEventCreateWithFlags(&ev_done[c], cudaEventDisableTiming));
        }
        // all input copies are queued up front on the copy-in stream, in chunk order
        const bool pg_spec = host_is_pageable(spec), pg_phase = init_phase && host_is_pageable(init_phase);
        for (int c = 0; c < n_chunks; ++c) {
            const size_t f0 = h->h_frame_off[cuts[c]], f1 = h->h_frame_off[cuts[c + 1]];
            const size_t Fb = h->F;
            CUE(copy_h2d(h, reinterpret_cast<float*>(h->ws_in.p) + f0 * Fb, spec + f0 * Fb, sizeof(float) * (f1 - f0) * Fb, h->copy_in, pg_spec));
            if (init_phase)
                CUE(copy_h2d(h, reinterpret_cast<float2*>(h->ws_in2.p) + f0 * Fb, reinterpret_cast<const float2*>(init_phase) + f0 * Fb,
                             sizeof(float2) * (f1 - f0) * Fb, h->copy_in, pg_phase));
            CUE(cudaEventRecord(ev_in[c], h->copy_in));
        }
    }
    h->gl.valid = false;
    int cur = 0;
    const float inv_thr = (float)(1.0 / (2.0e-8 * gscale));      // est / max(1e-8, |est|) on slots that hold 2*g*est
    if (ep && (rc = h->ws_ep.reserve(2 * sizeof(long long) * (size_t)batch))) { cleanup(); return rc; }      // endpoints | peaks
    // the utterances [b0, b1) as a batch of their own (descriptor pointers shifted, bases remembered)
    struct Sub { Batch B; int groups, frames, tiles; };
    auto sub_batch = [&](int b0, int b1) {
        Sub s;
        s.B = d.dev;
        s.B.frame_off += b0; s.B.tile_off += b0; s.B.samp_off += b0; s.B.batch = b1 - b0;
        s.B.frame_base = h->h_frame_off[b0]; s.B.tile_base = h->h_tile_off[b0]; s.B.utt_base = b0;
        s.B.group_off += b0; s.B.group_base = h->h_group_off[b0];
        s.groups = h->h_group_off[b1] - h->h_group_off[b0];
        s.frames = h->h_frame_off[b1] - h->h_frame_off[b0];
        s.tiles = h->h_tile_off[b1] - h->h_tile_off[b0];
        return s;
    };
#define CUL(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(NSB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)
    // chunk c arrives: magnitudes in the kernels' layout and the initial waveform y0 = _istft(S * angles)
    auto stage_in = [&](int c, cudaStream_t s_) -> int {
        const Sub u = sub_batch(cuts[c], cuts[c + 1]);
        if (space == NSB_HOST) CUL(cudaStreamWaitEvent(s_, ev_in[c], 0));
        // S = _db_to_amp(_denormalize(spec) + ref_level_db) ** power   (or |S| for _griffin_lim), permuted layout
        PrepParams Q{};
        Q.batch = u.B; Q.in = d_spec; Q.bin_major = (layout == NSB_BIN_MAJOR); Q.denorm = (flags & NSB_GL_DENORMALIZE) ? 1 : 0;
        Q.min_level_db = h->hp.min_level_db; Q.ref_level_db = h->hp.ref_level_db; Q.power = h->hp.power;
        {
            const double l2_10 = 3.321928094887362347870319429489390175864831;     // log2(10)
            Q.e_slope = -h->hp.min_level_db * 0.05 * h->hp.power * l2_10;
            Q.e_offset = (h->hp.min_level_db + h->hp.ref_level_db) * 0.05 * h->hp.power * l2_10 + std::log2(gscale);
        }
        Q.mag = reinterpret_cast<float*>(h->ws_mag.p); Q.total_frames = u.frames; Q.status = h->d_status; Q.scale = (float)gscale;
        if (h->generic) {
            NSB_LAUNCH(k_gen_prepare_mag, grid_1d((long long)u.frames * h->F, 256, 16 * h->num_sms), 256, 0, s_, Q, h->F);
            int r_ = check_launch(h, "k_gen_prepare_mag");
            if (r_) return r_;
            const long long s_cnt = h->h_samp_off[cuts[c + 1]] - h->h_samp_off[cuts[c]];
            return gen_synthesize(h, tf ? h->gen_tf : h->gen, u.B, u.frames, s_cnt, tf ? SRC_MAGZERO : (init_phase ? SRC_MAGPHASE : SRC_MAGRAND), nullptr,
                                  Q.mag, d_phase, layout == NSB_BIN_MAJOR, 0, seed, reinterpret_cast<float*>(h->ws_y0.p), s_);
        }
        if (Q.bin_major) {
            CUL(cudaMemsetAsync(Q.mag + (size_t)u.B.frame_base * kMagPitch, 0, sizeof(float) * kMagPitch * (size_t)u.frames, s_));
            NSB_LAUNCH(k_prepare_mag, (u.frames + 31) / 32, 256, 0, s_, Q);
        } else {
            NSB_LAUNCH(k_prepare_mag_rows, grid_1d(u.frames, kWarpsPerCta, 8 * h->num_sms), kThreads, 0, s_, Q);
        }
        int r_ = check_launch(h, "k_prepare_mag");
        if (r_) return r_;
        SynthParams P{};
        P.plan = make_plan(h, tf); P.batch = u.B; P.mag = Q.mag; P.spec = d_phase; P.spec_bin_major = (layout == NSB_BIN_MAJOR);
        P.y_out = reinterpret_cast<float*>(h->ws_y0.p); P.tile_hops = H; P.colours = h->colours; P.seed = seed; P.status = h->d_status;
        const size_t smem = synth_smem(h->hop, H);
        if (tf) launch_synth_tf(h, SRC_MAGZERO, P, u.tiles, smem, s_);
        else if (init_phase) launch_synth<SRC_MAGPHASE>(h, P, u.tiles, smem, s_);
        else launch_synth<SRC_MAGRAND>(h, P, u.tiles, smem, s_);
        return check_launch(h, "k_synth<init>");
    };
    // n iterations on the utterances [b0, b1); cur_ says which of the two waveform buffers holds their y
    auto iterate = [&](int b0, int b1, int n, int& cur_, cudaStream_t s_, DevBuf* done_buf) -> int {
        const Sub u = sub_batch(b0, b1);
        if (h->generic) {
            float* yb[2] = {reinterpret_cast<float*>(h->ws_y0.p), reinterpret_cast<float*>(h->ws_y1.p)};
            const long long s_cnt = h->h_samp_off[b1] - h->h_samp_off[b0];
            for (int it = 0; it < n; ++it) {
                int r_ = gen_synthesize(h, tf ? h->gen_tf : h->gen, u.B, u.frames, s_cnt, SRC_Y, yb[cur_], reinterpret_cast<const float*>(h->ws_mag.p),
                                        nullptr, 0, tf ? 1 : 0, 0, yb[cur_ ^ 1], s_);
                if (r_) return r_;
                cur_ ^= 1;
            }
            return NSB_OK;
        }
        return gl_iterations(h, u.B, u.tiles, u.groups, H, cur_, n, s_, tf, inv_thr, done_buf);
    };
    // chunk c is finished: de-emphasis (+ endpoint search) and the copy out
    auto stage_out = [&](int c, int cur_, cudaStream_t s_) -> int {
        const int b0 = cuts[c], b1 = cuts[c + 1];
        const Sub u = sub_batch(b0, b1);
        const float* y_fin = reinterpret_cast<const float*>(cur_ ? h->ws_y1.p : h->ws_y0.p);
        const long long s_base = h->h_samp_off[b0], s_cnt = h->h_samp_off[b1] - s_base;
        if ((flags & NSB_GL_DEEMPHASIS) || gscale != 1.0) {
            EmphParams E{};
            E.batch = u.B; E.in = y_fin; E.p = (flags & NSB_GL_DEEMPHASIS) ? h->hp.preemphasis : 0.0; E.scale = 1.0 / gscale;
            E.status = h->d_status;
            if (out_dtype == NSB_F64) E.out64 = reinterpret_cast<double*>(d_out); else E.out32 = reinterpret_cast<float*>(d_out);
            long long max_len = 0;
            for (int b = b0; b < b1; ++b) max_len = std::max(max_len, (long long)(h->h_samp_off[b + 1] - h->h_samp_off[b]));
            const int dgrid = deemph_grid(E.p, max_len, E, b1 - b0);
            NSB_LAUNCH(k_deemphasis, dgrid, kDeemphThreads, 0, s_, E);
            int r_ = check_launch(h, "k_deemphasis");
            if (r_) return r_;
        } else {
            CUL(cudaMemcpyAsync(d_out + s_base * sizeof(float), y_fin + s_base, sizeof(float) * s_cnt, cudaMemcpyDeviceToDevice, s_));
        }
        if (ep) {
            EndpointParams EP{};
            EP.batch = u.B;
            if (out_dtype == NSB_F64) EP.in64 = reinterpret_cast<const double*>(d_out); else EP.in32 = reinterpret_cast<const float*>(d_out);
            EP.out = reinterpret_cast<long long*>(h->ws_ep.p) + b0;
            endpoint_params(h, ep->threshold_db, ep->min_silence_sec, EP);
            NSB_LAUNCH(k_find_endpoint, b1 - b0, 256, 0, s_, EP);
            int r_ = check_launch(h, "k_find_endpoint");
            if (r_) return r_;
            if (ep->peak) {
                // save_wav's scaling on what the caller keeps: wav[:endpoint] *= 32767 / max(0.01, max |wav[:endpoint]|)  (audio.py:17-19)
                PeakParams K{};
                K.batch = u.B; K.limit = EP.out; K.peak = reinterpret_cast<unsigned long long*>(h->ws_ep.p) + batch + b0; K.status = h->d_status;
                if (out_dtype == NSB_F64) K.in64 = reinterpret_cast<const double*>(d_out); else K.in32 = reinterpret_cast<const float*>(d_out);
                if (to_i16) K.out16 = reinterpret_cast<short*>(d_final); else K.out64 = reinterpret_cast<double*>(d_final);
                long long max_len = 0;
                for (int b = b0; b < b1; ++b) max_len = std::max(max_len, (long long)(h->h_samp_off[b + 1] - h->h_samp_off[b]));
                K.n_seg = (int)std::min<long long>(64, std::max<long long>(1, max_len / 16384));
                CUL(cudaMemsetAsync(K.peak, 0, sizeof(unsigned long long) * (size_t)(b1 - b0), s_));
                NSB_LAUNCH(k_peak_max, (b1 - b0) * K.n_seg, 256, 0, s_, K);
                if ((r_ = check_launch(h, "k_peak_max"))) return r_;
                NSB_LAUNCH(k_peak_apply, grid_1d(s_cnt, 256, 8 * h->num_sms), 256, 0, s_, K, s_base, s_base + s_cnt);
                if ((r_ = check_launch(h, "k_peak_apply"))) return r_;
            }
        }
        if (space == NSB_HOST) {
            CUL(cudaEventRecord(ev_done[c], s_));
            CUL(cudaStreamWaitEvent(h->copy_out, ev_done[c], 0));
            CUL(cudaMemcpyAsync(reinterpret_cast<char*>(wav_out) + s_base * out_elt, d_final + s_base * out_elt, out_elt * s_cnt,
                                cudaMemcpyDeviceToHost, h->copy_out));
        }
        return NSB_OK;
    };
#undef CUL
    const cudaStream_t st_call = st;
    bool overlap = false;
    if (waves > 0) {
        // ---- wave schedule (see wave_plan): chunk g joins at wave g and runs `waves` launches of `wave_iters` iterations,
        // every launch covers the chunks that have arrived and are not finished - a contiguous utterance range ----
        const int G = n_chunks;
        // the launches grow from wave to wave: size the scheduling counters for the largest one now (a reallocation in the
        // middle of the pipeline would synchronise the device)
        if ((rc = h->d_done.reserve(sizeof(int) * ((size_t)d.total_tiles + (size_t)d.total_groups + (size_t)batch + 2)))) { cleanup(); return rc; }
        cur = iters & 1;                    // an odd iteration count: every chunk does ONE iteration on its own when it arrives,
                                            // so that all chunks of a launch agree on which buffer holds y
        for (int w = 0; w < G + waves - 1; ++w) {
            if (w < G) {
                if ((rc = stage_in(w, st))) { cleanup(); return rc; }
                if (iters & 1) { int c1 = 0; if ((rc = iterate(cuts[w], cuts[w + 1], 1, c1, st, nullptr))) { cleanup(); return rc; } }
            }
            const int g_lo = w - waves + 1 > 0 ? w - waves + 1 : 0, g_hi = w < G - 1 ? w : G - 1;
            int cw = cur;
            if ((rc = iterate(cuts[g_lo], cuts[g_hi + 1], wave_iters, cw, st, nullptr))) { cleanup(); return rc; }   // wave_iters is even: cw == cur
            if (w - waves + 1 >= 0 && (rc = stage_out(w - waves + 1, cur, st))) { cleanup(); return rc; }
        }
    } else {
    // Two compute streams take the chunks in turn: every chunk's iteration launch is persistent with dynamic work items, so
    // its CTAs retire one by one in its tail and the NEXT chunk's kernels (other stream) move into the freed SMs - the
    // tails and ramps of consecutive chunks overlap instead of adding up.  A launch only ever waits for items taken
    // earlier by CTAs that are already running, so sharing the SMs cannot deadlock.  Chunks touch disjoint ranges of every
    // workspace; the scheduling counters are per stream (d_done / d_done2).
    overlap = space == NSB_HOST && n_chunks > 1 && h->overlap_chunks && !h->trace_on;
    if (overlap) {
        const size_t cnt = sizeof(int) * ((size_t)d.total_tiles + (size_t)d.total_groups + (size_t)batch + 2);
        if ((rc = h->d_done.reserve(cnt)) || (rc = h->d_done2.reserve(cnt))) { cleanup(); return rc; }
        CUE(cudaEventRecord(h->chunk_fork, st));                 // descriptors uploaded, earlier work of this stream done
        CUE(cudaStreamWaitEvent(h->chunk_stream, h->chunk_fork, 0));
    }
    for (int c = 0; c < n_chunks; ++c) {
        st = (overlap && (c & 1)) ? h->chunk_stream : st_call;
        DevBuf* done_buf = (overlap && (c & 1)) ? &h->d_done2 : &h->d_done;
        if ((rc = stage_in(c, st))) { cleanup(); return rc; }
        cur = 0;
        if ((rc = iterate(cuts[c], cuts[c + 1], iters, cur, st, done_buf))) { cleanup(); return rc; }
        if ((rc = stage_out(c, cur, st))) { cleanup(); return rc; }
    }
    }
    st = st_call;
    if (overlap) {                                               // join: the call's stream continues after both
        CUE(cudaEventRecord(h->chunk_join, h->chunk_stream));
        CUE(cudaStreamWaitEvent(st, h->chunk_join, 0));
    }
    // device-resident state for nsb_griffin_lim_iterate: the whole batch
    h->gl.valid = (space == NSB_DEVICE) && !h->generic; h->gl.batch = d.dev; h->gl.total_frames = d.total_frames; h->gl.tile_hops = H;
    h->gl.total_tiles = d.total_tiles; h->gl.total_groups = d.total_groups; h->gl.cur = cur; h->gl.tf = tf; h->gl.inv_thr = (float)(1.0 / (2.0e-8 * gscale));
    if (ep) CUE(cudaMemcpyAsync(ep->out, h->ws_ep.p, sizeof(long long) * (size_t)batch,
                                space == NSB_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
    if (space == NSB_HOST) {
        CUE(cudaStreamSynchronize(h->copy_out));
        cleanup();
        return read_status(h, st);
    }
#undef CUE
    cleanup();
    return NSB_OK;
}

extern "C" int nsb_griffin_lim(nsb_handle_t h, const float* spec, int32_t layout, const int32_t* n_frames, int32_t batch,
                               const float* init_phase, uint64_t seed, int32_t iters, int32_t flags,
                               void* wav_out, int32_t out_dtype, int32_t space, void* stream) {
    return griffin_lim_impl(h, spec, layout, n_frames, batch, init_phase, seed, iters, flags, wav_out, out_dtype, space, stream, nullptr);
}

extern "C" int nsb_synthesize_ex(nsb_handle_t h, const float* spec, const int32_t* n_frames, int32_t batch, int32_t iters,
                                 double threshold_db, double min_silence_sec, int32_t flags, void* wav_out, int32_t out_dtype,
                                 int64_t* endpoints, int32_t space, void* stream) {
    if (!endpoints) return fail(NSB_ERR_INVALID, "null endpoints");
    if (!(min_silence_sec > 0.0)) return fail(NSB_ERR_INVALID, "min_silence_sec must be positive");
    if (out_dtype != NSB_F64 && out_dtype != NSB_I16) return fail(NSB_ERR_INVALID, "out_dtype must be NSB_F64 or NSB_I16");
    if (out_dtype == NSB_I16 && !(flags & NSB_SYNTH_PEAK_NORMALIZE)) return fail(NSB_ERR_INVALID, "NSB_I16 output needs NSB_SYNTH_PEAK_NORMALIZE (save_wav's scaling)");
    EndpointReq ep;
    ep.out = endpoints; ep.threshold_db = threshold_db; ep.min_silence_sec = min_silence_sec;
    ep.peak = (flags & NSB_SYNTH_PEAK_NORMALIZE) ? 1 : 0; ep.final_dtype = out_dtype;
    return griffin_lim_impl(h, spec, NSB_FRAME_MAJOR, n_frames, batch, nullptr, 0, iters,
                            NSB_GL_TF_TWIN | NSB_GL_DENORMALIZE | NSB_GL_DEEMPHASIS, wav_out, NSB_F64, space, stream, &ep);
}
extern "C" int nsb_synthesize(nsb_handle_t h, const float* spec, const int32_t* n_frames, int32_t batch, int32_t iters,
                              double threshold_db, double min_silence_sec, double* wav_out, int64_t* endpoints, int32_t space, void* stream) {
    return nsb_synthesize_ex(h, spec, n_frames, batch, iters, threshold_db, min_silence_sec, 0, wav_out, NSB_F64, endpoints, space, stream);
}

extern "C" int nsb_features_padded(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch, int32_t rows_per_utt,
                                   float* lin_out, float* mel_out, int32_t space, void* stream) {
    if (rows_per_utt < 1) return fail(NSB_ERR_INVALID, "rows_per_utt must be positive");
    return run_analysis(h, ANALYSIS_FEATURES, true, wav, n_samples, batch, nullptr, lin_out, mel_out, space, stream, false, rows_per_utt);
}

extern "C" int nsb_features_rows(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch, const int64_t* row_off, int64_t total_rows,
                                 float* lin_out, float* mel_out, int32_t space, void* stream) {
    if (!row_off || total_rows < 1) return fail(NSB_ERR_INVALID, "row_off is null or total_rows < 1");
    if (total_rows > 2000000000LL) return fail(NSB_ERR_INVALID, "more than 2e9 output rows");
    return run_analysis(h, ANALYSIS_FEATURES, true, wav, n_samples, batch, nullptr, lin_out, mel_out, space, stream, false, 0, row_off, total_rows);
}

extern "C" int nsb_frame_energy(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch, int32_t frame_length,
                                int32_t hop_length, double* out, int32_t space, void* stream) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (!wav || !n_samples || !out || batch < 1) return fail(NSB_ERR_INVALID, "null/empty argument");
    if (frame_length < 1 || hop_length < 1) return fail(NSB_ERR_INVALID, "frame_length and hop_length must be positive");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = pick_stream(h, stream, space);
    CallScope scope(h, st, space);
    std::vector<int> frames(batch); std::vector<long long> samples(batch);
    for (int b = 0; b < batch; ++b) {
        if (n_samples[b] < 1) return fail(NSB_ERR_INVALID, "utterance %d is empty", b);
        samples[b] = n_samples[b];
        frames[b] = (int)(1 + n_samples[b] / hop_length);
    }
    Desc d;
    int rc = upload_desc(h, st, frames, samples, 0, &d);
    if (rc) return rc;
    const float* d_wav = wav;
    double* d_out = out;
    if (space == NSB_HOST) {
        if ((rc = h->ws_in.reserve(sizeof(float) * (size_t)d.total_samples))) return rc;
        if ((rc = h->ws_out.reserve(sizeof(double) * (size_t)d.total_frames))) return rc;
        CU(cudaMemcpyAsync(h->ws_in.p, wav, sizeof(float) * (size_t)d.total_samples, cudaMemcpyHostToDevice, st));
        d_wav = reinterpret_cast<const float*>(h->ws_in.p);
        d_out = reinterpret_cast<double*>(h->ws_out.p);
    }
    EnergyParams E{};
    E.batch = d.dev; E.wav = d_wav; E.out = d_out; E.frame_length = frame_length; E.hop_length = hop_length; E.total_frames = d.total_frames;
    NSB_LAUNCH(k_frame_energy, grid_1d(d.total_frames, 8, 8 * h->num_sms), 256, 0, st, E);
    if ((rc = check_launch(h, "k_frame_energy"))) return rc;
    if (space == NSB_HOST) {
        CU(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)d.total_frames, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return NSB_OK;
}

extern "C" int nsb_find_endpoint(nsb_handle_t h, const void* wav, int32_t wav_dtype, const int64_t* n_samples, int32_t batch,
                                 double threshold_db, double min_silence_sec, int64_t* endpoints, int32_t space, void* stream) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (!wav || !n_samples || !endpoints || batch < 1) return fail(NSB_ERR_INVALID, "null/empty argument");
    if (wav_dtype != NSB_F32 && wav_dtype != NSB_F64) return fail(NSB_ERR_INVALID, "bad wav_dtype");
    if (!(min_silence_sec > 0.0)) return fail(NSB_ERR_INVALID, "min_silence_sec must be positive");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = pick_stream(h, stream, space);
    CallScope scope(h, st, space);
    std::vector<int> frames(batch, 0); std::vector<long long> samples(batch);
    for (int b = 0; b < batch; ++b) {
        if (n_samples[b] < 0) return fail(NSB_ERR_INVALID, "utterance %d has a negative length", b);
        samples[b] = n_samples[b];
    }
    Desc d;
    int rc = upload_desc(h, st, frames, samples, 0, &d);
    if (rc) return rc;
    const size_t elt = wav_dtype == NSB_F64 ? sizeof(double) : sizeof(float);
    const void* d_wav = wav;
    if (space == NSB_HOST) {
        if ((rc = h->ws_in.reserve(elt * (size_t)d.total_samples + 16))) return rc;
        CU(cudaMemcpyAsync(h->ws_in.p, wav, elt * (size_t)d.total_samples, cudaMemcpyHostToDevice, st));
        d_wav = h->ws_in.p;
    }
    if ((rc = h->ws_ep.reserve(sizeof(long long) * (size_t)batch))) return rc;
    EndpointParams EP{};
    EP.batch = d.dev;
    if (wav_dtype == NSB_F64) EP.in64 = reinterpret_cast<const double*>(d_wav); else EP.in32 = reinterpret_cast<const float*>(d_wav);
    EP.out = reinterpret_cast<long long*>(h->ws_ep.p);
    endpoint_params(h, threshold_db, min_silence_sec, EP);
    NSB_LAUNCH(k_find_endpoint, batch, 256, 0, st, EP);
    if ((rc = check_launch(h, "k_find_endpoint"))) return rc;
    CU(cudaMemcpyAsync(endpoints, h->ws_ep.p, sizeof(long long) * (size_t)batch,
                       space == NSB_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
    if (space == NSB_HOST) CU(cudaStreamSynchronize(st));
    return NSB_OK;
}

// ---------------------------------------------------------------------------------------------
// small helpers of the module
// ---------------------------------------------------------------------------------------------
static int run_emph(nsb_handle_s* h, bool inverse, const float* x, const int64_t* n_samples, int batch, void* out, int out_dtype,
                    int space, void* stream) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (!x || !n_samples || !out || batch < 1) return fail(NSB_ERR_INVALID, "null/empty argument");
    if (out_dtype != NSB_F32 && out_dtype != NSB_F64) return fail(NSB_ERR_INVALID, "bad out_dtype");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = pick_stream(h, stream, space);
    CallScope scope(h, st, space);
    std::vector<int> frames(batch, 1);
    std::vector<long long> samples(batch);
    for (int b = 0; b < batch; ++b) { if (n_samples[b] < 0) return fail(NSB_ERR_INVALID, "negative length"); samples[b] = n_samples[b]; }
    Desc d;
    int rc = upload_desc(h, st, frames, samples, 0, &d);
    if (rc) return rc;
    if (d.total_samples == 0) return NSB_OK;
    const size_t out_elt = out_dtype == NSB_F64 ? sizeof(double) : sizeof(float);
    const float* d_x = x;
    void* d_out = out;
    if (space == NSB_HOST) {
        if ((rc = h->ws_in.reserve(sizeof(float) * d.total_samples))) return rc;
        if ((rc = h->ws_out.reserve(out_elt * d.total_samples))) return rc;
        CU(cudaMemcpyAsync(h->ws_in.p, x, sizeof(float) * d.total_samples, cudaMemcpyHostToDevice, st));
        d_x = reinterpret_cast<const float*>(h->ws_in.p);
        d_out = h->ws_out.p;
    }
    EmphParams E{};
    E.batch = d.dev; E.in = d_x; E.p = h->hp.preemphasis;
    if (out_dtype == NSB_F64) E.out64 = reinterpret_cast<double*>(d_out); else E.out32 = reinterpret_cast<float*>(d_out);
    if (inverse) {
        long long max_len = 0;
        for (int b = 0; b < batch; ++b) max_len = std::max(max_len, (long long)n_samples[b]);
        const int dgrid = deemph_grid(E.p, max_len, E, batch);
        NSB_LAUNCH(k_deemphasis, dgrid, kDeemphThreads, 0, st, E);
    }
    else NSB_LAUNCH(k_preemphasis, grid_1d(d.total_samples, 256, 4 * h->num_sms), 256, 0, st, E, d.total_samples);
    if ((rc = check_launch(h, inverse ? "k_deemphasis" : "k_preemphasis"))) return rc;
    if (space == NSB_HOST) {
        CU(cudaMemcpyAsync(out, d_out, out_elt * d.total_samples, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return NSB_OK;
}
extern "C" int nsb_preemphasis(nsb_handle_t h, const float* x, const int64_t* n, int32_t batch, void* out, int32_t dt, int32_t space, void* stream) {
    return run_emph(h, false, x, n, batch, out, dt, space, stream);
}
extern "C" int nsb_inv_preemphasis(nsb_handle_t h, const float* x, const int64_t* n, int32_t batch, void* out, int32_t dt, int32_t space, void* stream) {
    return run_emph(h, true, x, n, batch, out, dt, space, stream);
}

extern "C" int nsb_linear_to_mel(nsb_handle_t h, const float* spec, int32_t layout, const int32_t* n_frames, int32_t batch,
                                 void* out, int32_t out_dtype, int32_t space, void* stream) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (!spec || !n_frames || !out || batch < 1) return fail(NSB_ERR_INVALID, "null/empty argument");
    if (out_dtype != NSB_F32 && out_dtype != NSB_F64) return fail(NSB_ERR_INVALID, "bad out_dtype");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = pick_stream(h, stream, space);
    CallScope scope(h, st, space);
    std::vector<int> frames(batch);
    std::vector<long long> samples(batch, 0);
    for (int b = 0; b < batch; ++b) { if (n_frames[b] < 1) return fail(NSB_ERR_INVALID, "utterance %d has no frames", b); frames[b] = n_frames[b]; }
    Desc d;
    int rc = upload_desc(h, st, frames, samples, 0, &d);
    if (rc) return rc;
    const size_t out_elt = out_dtype == NSB_F64 ? sizeof(double) : sizeof(float);
    const size_t n_in = (size_t)h->F * d.total_frames, n_out = (size_t)h->num_mels * d.total_frames;
    const float* d_in = spec;
    void* d_out = out;
    if (space == NSB_HOST) {
        if ((rc = h->ws_in.reserve(sizeof(float) * n_in))) return rc;
        if ((rc = h->ws_out.reserve(out_elt * n_out))) return rc;
        CU(cudaMemcpyAsync(h->ws_in.p, spec, sizeof(float) * n_in, cudaMemcpyHostToDevice, st));
        d_in = reinterpret_cast<const float*>(h->ws_in.p);
        d_out = h->ws_out.p;
    }
    MelParams M{};
    M.plan = make_plan(h); M.batch = d.dev; M.in = d_in; M.bin_major = (layout == NSB_BIN_MAJOR); M.total_frames = d.total_frames;
    if (out_dtype == NSB_F64) M.out64 = reinterpret_cast<double*>(d_out); else M.out32 = reinterpret_cast<float*>(d_out);
    if (h->generic) NSB_LAUNCH(k_gen_linear_to_mel, grid_1d((long long)d.total_frames * h->num_mels, 256, 8 * h->num_sms), 256, 0, st, M, h->F);
    else NSB_LAUNCH(k_linear_to_mel, grid_1d(d.total_frames, kWarpsPerCta, 4 * h->num_sms), kThreads, 0, st, M);
    if ((rc = check_launch(h, "k_linear_to_mel"))) return rc;
    if (space == NSB_HOST) {
        CU(cudaMemcpyAsync(out, d_out, out_elt * n_out, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return NSB_OK;
}

extern "C" int nsb_elementwise(nsb_handle_t h, int32_t op, const float* in, int64_t n, float* out, int32_t space, void* stream) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (op < 0 || op > 3) return fail(NSB_ERR_INVALID, "bad op %d", op);
    if (n < 0 || (n > 0 && (!in || !out))) return fail(NSB_ERR_INVALID, "null/negative argument");
    if (n == 0) return NSB_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = pick_stream(h, stream, space);
    CallScope scope(h, st, space);
    const float* d_in = in;
    float* d_out = out;
    int rc;
    if (space == NSB_HOST) {
        if ((rc = h->ws_in.reserve(sizeof(float) * n))) return rc;
        if ((rc = h->ws_out.reserve(sizeof(float) * n))) return rc;
        CU(cudaMemcpyAsync(h->ws_in.p, in, sizeof(float) * n, cudaMemcpyHostToDevice, st));
        d_in = reinterpret_cast<const float*>(h->ws_in.p);
        d_out = reinterpret_cast<float*>(h->ws_out.p);
    }
    NSB_LAUNCH(k_elementwise, grid_1d(n, 256, 8 * h->num_sms), 256, 0, st, op, d_in, d_out, n, (float)h->hp.min_level_db);
    if ((rc = check_launch(h, "k_elementwise"))) return rc;
    if (space == NSB_HOST) {
        CU(cudaMemcpyAsync(out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return NSB_OK;
}

// ---------------------------------------------------------------------------------------------
// save_wav's scaling as an entry point of its own (utils/audio.py:17-19)
// ---------------------------------------------------------------------------------------------
extern "C" int nsb_peak_normalize(nsb_handle_t h, const void* wav, int32_t wav_dtype, const int64_t* n_samples, int32_t batch,
                                  const int64_t* limit, void* out, int32_t out_dtype, int32_t space, void* stream) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (!wav || !n_samples || !out || batch < 1) return fail(NSB_ERR_INVALID, "null/empty argument");
    if (wav_dtype != NSB_F32 && wav_dtype != NSB_F64) return fail(NSB_ERR_INVALID, "bad wav_dtype");
    if (out_dtype != NSB_F64 && out_dtype != NSB_I16) return fail(NSB_ERR_INVALID, "out_dtype must be NSB_F64 or NSB_I16");
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = pick_stream(h, stream, space);
    CallScope scope(h, st, space);
    std::vector<int> frames(batch, 0); std::vector<long long> samples(batch);
    long long max_len = 0;
    for (int b = 0; b < batch; ++b) {
        if (n_samples[b] < 0) return fail(NSB_ERR_INVALID, "utterance %d has a negative length", b);
        samples[b] = n_samples[b];
        max_len = std::max(max_len, samples[b]);
    }
    Desc d;
    int rc = upload_desc(h, st, frames, samples, 0, &d);
    if (rc) return rc;
    if (d.total_samples == 0) return NSB_OK;
    const size_t in_elt = dtype_size(wav_dtype), out_elt = dtype_size(out_dtype);
    const void* d_in = wav;
    void* d_out = out;
    if ((rc = h->ws_ep.reserve(2 * sizeof(long long) * (size_t)batch))) return rc;
    long long* d_limit = nullptr;
    if (space == NSB_HOST) {
        if ((rc = h->ws_in.reserve(in_elt * (size_t)d.total_samples))) return rc;
        if ((rc = h->ws_out.reserve(out_elt * (size_t)d.total_samples))) return rc;
        CU(cudaMemcpyAsync(h->ws_in.p, wav, in_elt * (size_t)d.total_samples, cudaMemcpyHostToDevice, st));
        d_in = h->ws_in.p; d_out = h->ws_out.p;
        if (limit) {
            d_limit = reinterpret_cast<long long*>(h->ws_ep.p);
            CU(cudaMemcpyAsync(d_limit, limit, sizeof(long long) * (size_t)batch, cudaMemcpyHostToDevice, st));
        }
    } else if (limit) {
        d_limit = const_cast<long long*>(reinterpret_cast<const long long*>(limit));
    }
    PeakParams K{};
    K.batch = d.dev; K.limit = d_limit; K.peak = reinterpret_cast<unsigned long long*>(h->ws_ep.p) + batch; K.status = h->d_status;
    if (wav_dtype == NSB_F64) K.in64 = reinterpret_cast<const double*>(d_in); else K.in32 = reinterpret_cast<const float*>(d_in);
    if (out_dtype == NSB_I16) K.out16 = reinterpret_cast<short*>(d_out); else K.out64 = reinterpret_cast<double*>(d_out);
    K.n_seg = (int)std::min<long long>(64, std::max<long long>(1, max_len / 16384));
    CU(cudaMemsetAsync(K.peak, 0, sizeof(unsigned long long) * (size_t)batch, st));
    NSB_LAUNCH(k_peak_max, batch * K.n_seg, 256, 0, st, K);
    if ((rc = check_launch(h, "k_peak_max"))) return rc;
    NSB_LAUNCH(k_peak_apply, grid_1d(d.total_samples, 256, 8 * h->num_sms), 256, 0, st, K, 0LL, d.total_samples);
    if ((rc = check_launch(h, "k_peak_apply"))) return rc;
    if (space == NSB_HOST) {
        CU(cudaMemcpyAsync(out, d_out, out_elt * (size_t)d.total_samples, cudaMemcpyDeviceToHost, st));
        return read_status(h, st);
    }
    return NSB_OK;
}

// ---------------------------------------------------------------------------------------------
// device memory for results handed to other frameworks (nspeech_b200/_buffers.py: DLPack / CUDA array interface)
// ---------------------------------------------------------------------------------------------
extern "C" int nsb_device_alloc(int device, uint64_t bytes, void** out) {
    if (!out) return fail(NSB_ERR_INVALID, "out is null");
    *out = nullptr;
    CU(cudaSetDevice(device));
    CU(cudaMalloc(out, bytes ? bytes : 1));
    return NSB_OK;
}
extern "C" int nsb_device_free(int device, void* p) {
    if (!p) return NSB_OK;
    CU(cudaSetDevice(device));
    CU(cudaFree(p));
    return NSB_OK;
}
// kind: 1 host -> device, 2 device -> host, 3 device -> device; synchronous
extern "C" int nsb_device_copy(int device, void* dst, const void* src, uint64_t bytes, int32_t kind) {
    if (kind < 1 || kind > 3) return fail(NSB_ERR_INVALID, "bad copy kind %d", kind);
    if (bytes == 0) return NSB_OK;
    if (!dst || !src) return fail(NSB_ERR_INVALID, "null pointer");
    CU(cudaSetDevice(device));
    CU(cudaMemcpy(dst, src, bytes, kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice));
    return NSB_OK;
}

// ---------------------------------------------------------------------------------------------
// asynchronous host calls: submit -> ticket, wait(ticket)
//
// A synchronous NSB_HOST call keeps its caller inside the library for the whole batch, so consecutive batches cannot overlap
// one's copy-out and pipeline tail with the next one's copy-in and ramp.  The asynchronous entry points hand the call to one
// of `slots` (default 3) worker threads of the handle.  Every slot owns a CHILD handle - private streams, descriptors,
// scheduling counters and workspaces, nothing shared with its siblings - and runs the ordinary synchronous call on it, so two
// batches are in flight on the GPU at once: the second one's H2D copies and first waves fill what the first one's tail leaves
// idle.  The caller's buffers must stay valid until nsb_wait(ticket) returns; n_frames / n_samples are copied at submit.
// (The feeder threads of datasets/datafeeder.py:110-152 are this pattern on the reference's side.)
// ---------------------------------------------------------------------------------------------
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <thread>
#include <unordered_map>

struct AsyncJob {
    std::function<int(nsb_handle_s*)> run;
    int status = NSB_OK;
    std::string err;
    bool done = false;
};
struct AsyncSlot {
    nsb_handle_s* child = nullptr;
    std::thread th;
    std::deque<std::shared_ptr<AsyncJob>> q;
};
struct nsb_async_s {
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    std::vector<std::unique_ptr<AsyncSlot>> slots;
    std::unordered_map<uint64_t, std::shared_ptr<AsyncJob>> jobs;
    uint64_t next_ticket = 1;
    bool stop = false;
};

static void async_worker(nsb_handle_s* parent, nsb_async_s* A, AsyncSlot* S) {
    for (;;) {
        std::shared_ptr<AsyncJob> job;
        {
            std::unique_lock<std::mutex> lk(A->m);
            A->cv_work.wait(lk, [&] { return A->stop || !S->q.empty(); });
            if (S->q.empty()) return;              // stop requested and nothing left to run
            job = S->q.front();
            S->q.pop_front();
        }
        int rc = NSB_OK;
        if (!S->child) {
            rc = nsb_create(&parent->hp, parent->device, &S->child);
            if (rc) S->child = nullptr;
        }
        if (!rc) {
            // the tuning state of the parent at the time the job runs
            S->child->host_chunks = parent->host_chunks; S->child->wave_schedule = parent->wave_schedule; S->child->specialize = parent->specialize;
            S->child->overlap_chunks = parent->overlap_chunks; S->child->use_generic_iter = parent->use_generic_iter;
            S->child->stream_sync_mode = parent->stream_sync_mode; S->child->fuse_iterations = parent->fuse_iterations;
            S->child->wide_mode = parent->wide_mode; S->child->mel_lines = parent->mel_lines;
            rc = job->run(S->child);
        }
        {
            std::lock_guard<std::mutex> lk(A->m);
            job->status = rc;
            if (rc) job->err = g_err;
            job->done = true;
            if (S->child) parent->launches_async += S->child->launches, S->child->launches = 0;
        }
        A->cv_done.notify_all();
    }
}

static int async_submit(nsb_handle_s* h, std::function<int(nsb_handle_s*)> fn, uint64_t* ticket) {
    if (!ticket) return fail(NSB_ERR_INVALID, "ticket is null");
    nsb_async_s* A;
    {
        std::lock_guard<std::mutex> lk(h->mu);
        if (!h->async) {
            h->async = new nsb_async_s();
            const int n = h->async_slots < 1 ? 1 : h->async_slots;
            for (int i = 0; i < n; ++i) {
                h->async->slots.emplace_back(new AsyncSlot());
                AsyncSlot* S = h->async->slots.back().get();
                S->th = std::thread(async_worker, h, h->async, S);
            }
        }
        A = h->async;
    }
    auto job = std::make_shared<AsyncJob>();
    job->run = std::move(fn);
    {
        std::lock_guard<std::mutex> lk(A->m);
        const uint64_t t = A->next_ticket++;
        A->jobs[t] = job;
        A->slots[t % A->slots.size()]->q.push_back(job);
        *ticket = t;
    }
    A->cv_work.notify_all();
    return NSB_OK;
}

static void async_shutdown(nsb_handle_s* h) {
    nsb_async_s* A = h->async;
    if (!A) return;
    { std::lock_guard<std::mutex> lk(A->m); A->stop = true; }
    A->cv_work.notify_all();
    for (auto& S : A->slots) if (S->th.joinable()) S->th.join();          // the workers drain their queues first
    for (auto& S : A->slots) if (S->child) nsb_destroy(S->child);
    delete A;
    h->async = nullptr;
}

extern "C" int nsb_wait(nsb_handle_t h, uint64_t ticket) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    nsb_async_s* A = h->async;
    if (!A) return fail(NSB_ERR_INVALID, "ticket %llu: nothing was submitted on this handle", (unsigned long long)ticket);
    std::shared_ptr<AsyncJob> job;
    {
        std::unique_lock<std::mutex> lk(A->m);
        auto it = A->jobs.find(ticket);
        if (it == A->jobs.end()) return fail(NSB_ERR_INVALID, "unknown or already collected ticket %llu", (unsigned long long)ticket);
        job = it->second;
        A->cv_done.wait(lk, [&] { return job->done; });
        A->jobs.erase(ticket);
    }
    if (job->status) g_err = job->err;
    return job->status;
}

extern "C" int nsb_set_async_slots(nsb_handle_t h, int32_t n) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (n < 1 || n > 8) return fail(NSB_ERR_INVALID, "async slots %d outside [1,8]", n);
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->async) return fail(NSB_ERR_INVALID, "the worker slots already run; set their number before the first submit");
    h->async_slots = n;
    return NSB_OK;
}

extern "C" int nsb_griffin_lim_submit(nsb_handle_t h, const float* spec, int32_t layout, const int32_t* n_frames, int32_t batch,
                                      const float* init_phase_complex, uint64_t seed, int32_t iters, int32_t flags,
                                      void* wav_out, int32_t out_dtype, uint64_t* ticket) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (!spec || !n_frames || !wav_out || batch < 1) return fail(NSB_ERR_INVALID, "null/empty argument");
    std::vector<int32_t> nf(n_frames, n_frames + batch);
    return async_submit(h, [=](nsb_handle_s* c) {
        // with several batches in flight the overlap comes from the neighbours: one chunk per call (copy in, ONE launch of all
        // iterations, copy out) beats the wave schedule, whose small first launches only pay when a call runs alone
        // (profiles/r2/e2e_async.txt: 17.45 against 18.67 ms per step)
        if (c->host_chunks == 0) c->host_chunks = 1;
        return nsb_griffin_lim(c, spec, layout, nf.data(), batch, init_phase_complex, seed, iters, flags, wav_out, out_dtype, NSB_HOST, nullptr);
    }, ticket);
}

extern "C" int nsb_features_submit(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch,
                                   float* lin_out, float* mel_out, uint64_t* ticket) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (!wav || !n_samples || batch < 1) return fail(NSB_ERR_INVALID, "null/empty argument");
    std::vector<int64_t> ns(n_samples, n_samples + batch);
    return async_submit(h, [=](nsb_handle_s* c) {
        return nsb_features(c, wav, ns.data(), batch, lin_out, mel_out, NSB_HOST, nullptr);
    }, ticket);
}

extern "C" int nsb_synthesize_submit(nsb_handle_t h, const float* spec, const int32_t* n_frames, int32_t batch, int32_t iters,
                                     double threshold_db, double min_silence_sec, int32_t flags, void* wav_out, int32_t out_dtype,
                                     int64_t* endpoints, uint64_t* ticket) {
    if (!h) return fail(NSB_ERR_INVALID, "null handle");
    if (!spec || !n_frames || !wav_out || !endpoints || batch < 1) return fail(NSB_ERR_INVALID, "null/empty argument");
    std::vector<int32_t> nf(n_frames, n_frames + batch);
    return async_submit(h, [=](nsb_handle_s* c) {
        if (c->host_chunks == 0) c->host_chunks = 1;
        return nsb_synthesize_ex(c, spec, nf.data(), batch, iters, threshold_db, min_silence_sec, flags, wav_out, out_dtype, endpoints, NSB_HOST, nullptr);
    }, ticket);
}
