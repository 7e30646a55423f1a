/*
 * nspeech_b200 - C ABI of the B200-native spectrogram / Griffin-Lim hot path.
 *
 * The reference (MLCogUP/nspeech) has no FFI layer: the boundary of this path is the module-level
 * Python API of neural_speech/utils/audio.py plus neural_speech/hparams.get_hparams().  Each entry
 * point below names the reference function(s) it replaces (file:line relative to /root/reference).
 * The Python mirror that binds these with ctypes is nspeech_b200/audio.py; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes, no framework types; every call returns an nsb_status (0 = OK) and
 *    nsb_last_error() gives the thread-local message of the last failure;
 *  - `space` says where ALL data buffers of the call live: NSB_HOST (the library copies in/out and
 *    synchronises before returning) or NSB_DEVICE (pointers into the handle's GPU; the call is
 *    stream-ordered on `stream` and returns without synchronising);
 *  - `stream` is a cudaStream_t used as given; NULL means CUDA's default stream for NSB_DEVICE calls and the
 *    handle's private non-blocking stream for the (synchronous) NSB_HOST calls.  A handle's device state (descriptors,
 *    scheduling counters, workspaces) is shared by its calls: use ONE stream per handle.  A call that arrives on another
 *    stream than the previous one first waits (on the device) for the event the previous call recorded when it returned;
 *  - nsb_*_submit / nsb_wait are the asynchronous form of the NSB_HOST calls (up to three batches in flight per handle);
 *  - ragged batches: per-utterance lengths are a HOST int array; utterance blocks are packed back to
 *    back in every buffer (a uniform [N,T,F] or [N,n] C-contiguous array is already in that form);
 *  - spectra are float32, 1025 = num_freq bins per frame; layout NSB_FRAME_MAJOR = [T][F] per
 *    utterance (the memory order of the reference's Fortran-ordered [F,T] arrays and of Tacotron's
 *    [T,F] outputs), NSB_BIN_MAJOR = [F][T] per utterance (a C-contiguous numpy [F,T]);
 *  - complex data is interleaved float32 (re, im) = numpy complex64;
 *  - no CPU fallback: without an sm_100 device nsb_create fails with NSB_ERR_NODEVICE.
 */
#ifndef NSPEECH_B200_H
#define NSPEECH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSB_ABI_VERSION 2

typedef struct nsb_handle_s* nsb_handle_t;

/* the audio keys of the hparams yaml (neural_speech/hparams/audio.yaml:6-18) */
typedef struct nsb_hparams {
    int32_t num_freq;
    int32_t num_mels;
    int32_t sample_rate;
    int32_t griffin_lim_iters;
    double frame_shift_ms;
    double frame_length_ms;
    double preemphasis;
    double ref_level_db;
    double min_level_db;
    double power;
} nsb_hparams;

typedef enum nsb_status {
    NSB_OK = 0,
    NSB_ERR_INVALID = 1,      /* bad argument (null pointer, wrong length, T < 2 for inversion, ...) */
    NSB_ERR_CUDA = 2,         /* a CUDA runtime call failed */
    NSB_ERR_UNSUPPORTED = 3,  /* hparams outside what the kernels implement (n_fft != 2048, win > n_fft) */
    NSB_ERR_NONFINITE = 4,    /* NaN/Inf in the audio or spectrogram (librosa.util.valid_audio's ParameterError) */
    NSB_ERR_NODEVICE = 5,     /* no usable sm_100 GPU */
    NSB_ERR_OOM = 6
} nsb_status;

enum { NSB_HOST = 0, NSB_DEVICE = 1 };
enum { NSB_FRAME_MAJOR = 0, NSB_BIN_MAJOR = 1 };
enum { NSB_F32 = 0, NSB_F64 = 1, NSB_I16 = 2 };
enum { NSB_EW_AMP_TO_DB = 0, NSB_EW_DB_TO_AMP = 1, NSB_EW_NORMALIZE = 2, NSB_EW_DENORMALIZE = 3 };

/* flags of nsb_griffin_lim */
enum {
    NSB_GL_DENORMALIZE = 1,   /* input is a normalised dB spectrogram: apply audio.py:47-48 first */
    NSB_GL_DEEMPHASIS = 2,    /* apply inv_preemphasis (audio.py:35-36) to the result */
    NSB_GL_TF_TWIN = 4        /* the TensorFlow twin's semantics, inv_spectrogram_tensorflow / _griffin_lim_tensorflow
                               * (audio.py:51-58, 90-103): zero initial phase, tf.contrib.signal framing (frame k =
                               * samples [k*hop, k*hop+win), window on them, zero-padded at the END to n_fft, no centring,
                               * no reflect padding), inverse without window-sum normalisation, phase = est/max(1e-8,|est|),
                               * output win + hop*(T-1) samples per utterance.  init_phase and seed are ignored. */
};

int nsb_abi_version(void);
const char* nsb_last_error(void);
int nsb_device_count(int* count);
/* PCI bus id ("0000:1b:00.0") of a device: lets the host side place its threads and page-locked buffers on the GPU's NUMA node (8-GPU sharding) */
int nsb_device_pci_bus_id(int device, char* out, int32_t len);

/* handle = hparams + device tables (window, twiddles, sparse mel basis) + grow-only workspaces.
 * Replaces the module-global state of the reference: get_hparams() (hparams/__init__.py:25-26) and the
 * cached _mel_basis (utils/audio.py:135-147). Thread-safe; calls on one handle serialise. */
int nsb_create(const nsb_hparams* hp, int device, nsb_handle_t* out);
int nsb_destroy(nsb_handle_t h);
int nsb_synchronize(nsb_handle_t h, void* stream);

/* _stft_parameters (utils/audio.py:126-130): n_fft, hop_length, win_length with the truncating int() */
int nsb_stft_parameters(nsb_handle_t h, int32_t* n_fft, int32_t* hop_length, int32_t* win_length);
/* frames of an n-sample signal (1 + n // hop) and samples of a T-frame inversion (hop * (T - 1)) */
int64_t nsb_num_frames(nsb_handle_t h, int64_t n_samples);
int64_t nsb_num_samples(nsb_handle_t h, int64_t n_frames);

/* _stft(y) or _stft(preemphasis(y)) (utils/audio.py:106-108, 31-32): centred, reflect-padded, periodic
 * Hann STFT.  wav: packed float32 samples; out: complex64, frame-major [sum T][num_freq]. */
int nsb_stft(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch, int32_t apply_preemphasis,
             float* out_complex, int32_t space, void* stream);

/* _stft_tensorflow(signals) (utils/audio.py:116-118): tf.contrib.signal.stft(signals, win, hop, n_fft, pad_end=False);
 * needs n >= win; T = 1 + (n - win) // hop frames; out complex64 frame-major [sum T][num_freq]. */
int nsb_stft_tf(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch, float* out_complex,
                int32_t space, void* stream);

/* spectrogram(y) and melspectrogram(y) (utils/audio.py:39-42, 61-64) from ONE STFT pass.
 * lin_out: [sum T][num_freq] or NULL; mel_out: [sum T][num_mels] or NULL (frame-major float32). */
int nsb_features(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch,
                 float* lin_out, float* mel_out, int32_t space, void* stream);

/* _istft(D) (utils/audio.py:111-113): spec complex64 in `layout`; wav_out float32, hop*(T-1) per utterance. */
int nsb_istft(nsb_handle_t h, const float* spec_complex, int32_t layout, const int32_t* n_frames, int32_t batch,
              float* wav_out, int32_t space, void* stream);

/* _istft_tensorflow(stfts) (utils/audio.py:121-123): tf.contrib.signal.inverse_stft(stfts, win, hop, n_fft) with its
 * TF-1.7 default window (periodic Hann, no normalisation); wav_out float32, win + hop*(T-1) per utterance. */
int nsb_istft_tf(nsb_handle_t h, const float* spec_complex, int32_t layout, const int32_t* n_frames, int32_t batch,
                 float* wav_out, int32_t space, void* stream);

/* _griffin_lim(S) / inv_spectrogram(spectrogram) (utils/audio.py:77-87, 45-48).
 *   spec        float32 magnitudes S (flags without NSB_GL_DENORMALIZE) or normalised spectrogram in [0,1]
 *   init_phase  complex64 unit phasors exp(i*phi) in the same layout, or NULL: then the phase is drawn on
 *               the device (Philox-4x32 keyed by `seed`) in place of np.random.rand (utils/audio.py:81)
 *   iters       < 0 -> hparams.griffin_lim_iters
 *   wav_out     out_dtype NSB_F32 (what _griffin_lim returns) or NSB_F64 (what inv_preemphasis returns) */
int nsb_griffin_lim(nsb_handle_t h, const float* spec, int32_t layout, const int32_t* n_frames, int32_t batch,
                    const float* init_phase_complex, uint64_t seed, int32_t iters, int32_t flags,
                    void* wav_out, int32_t out_dtype, int32_t space, void* stream);

/* preemphasis(x) / inv_preemphasis(x) (utils/audio.py:31-36): scipy.signal.lfilter with zero initial state. */
int nsb_preemphasis(nsb_handle_t h, const float* x, const int64_t* n_samples, int32_t batch,
                    void* out, int32_t out_dtype, int32_t space, void* stream);
int nsb_inv_preemphasis(nsb_handle_t h, const float* x, const int64_t* n_samples, int32_t batch,
                        void* out, int32_t out_dtype, int32_t space, void* stream);

/* _linear_to_mel(S) (utils/audio.py:138-147): librosa.filters.mel (Slaney, area-normalised) applied to
 * linear magnitudes in `layout`; out frame-major [sum T][num_mels]. */
int nsb_linear_to_mel(nsb_handle_t h, const float* spec, int32_t layout, const int32_t* n_frames, int32_t batch,
                      void* out, int32_t out_dtype, int32_t space, void* stream);
/* the dense basis itself, float64 [num_mels][num_freq] into a HOST buffer (_build_mel_basis, utils/audio.py:145-147) */
int nsb_mel_basis(nsb_handle_t h, double* out_host);

/* _amp_to_db, _db_to_amp, _normalize, _denormalize (utils/audio.py:150-167), element-wise float32 */
int nsb_elementwise(nsb_handle_t h, int32_t op, const float* in, int64_t n, float* out, int32_t space, void* stream);

/* deferred device-side error flag of NSB_DEVICE calls (bit 0: non-finite data seen); synchronises; clears it */
int nsb_check_status(nsb_handle_t h, void* stream);

/* find_endpoint(wav, threshold_db=-40, min_silence_sec=0.8) (utils/audio.py:67-74), batched: wav packed per utterance,
 * wav_dtype NSB_F32 / NSB_F64; endpoints[b] = the sample count to keep (x + hop of the first silent window, else the length). */
int nsb_find_endpoint(nsb_handle_t h, const void* wav, int32_t wav_dtype, const int64_t* n_samples, int32_t batch,
                      double threshold_db, double min_silence_sec, int64_t* endpoints, int32_t space, void* stream);

/* The feeder's target tensors in one pass (datasets/datafeeder.py:190-220, _prepare_targets): spectrogram(y).T and
 * melspectrogram(y).T of every utterance written time-major into zero-padded batch tensors
 * lin_out [batch][rows_per_utt][num_freq], mel_out [batch][rows_per_utt][num_mels] (either may be NULL);
 * rows_per_utt >= every utterance's frame count (the feeder uses round_up(max frames + 1, outputs_per_step)). */
int nsb_features_padded(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch, int32_t rows_per_utt,
                        float* lin_out, float* mel_out, int32_t space, void* stream);

/* mean(|x|^2) of every centred frame = librosa.feature.rmse(y, frame_length, hop_length) ** 2 (reflect padding), the
 * reduction behind trim_wav / trim_silence (datasets/process.py:39-54).  out: float64, 1 + n // hop_length values per
 * utterance, packed. */
int nsb_frame_energy(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch, int32_t frame_length,
                     int32_t hop_length, double* out, int32_t space, void* stream);

/* The spectrogram -> waveform stage of Synthesizer.synthesize (synthesizer.py:30, 51-53) for a batch, in one pipeline:
 * inv_spectrogram_tensorflow (NSB_GL_TF_TWIN semantics) -> inv_preemphasis -> find_endpoint.  spec: normalised linear
 * spectrograms, frame-major [sum T][num_freq]; wav_out float64, win + hop*(T-1) samples per utterance (the caller keeps the
 * first endpoints[b] of them, synthesizer.py:53). */
int nsb_synthesize(nsb_handle_t h, const float* spec, const int32_t* n_frames, int32_t batch, int32_t iters,
                   double threshold_db, double min_silence_sec, double* wav_out, int64_t* endpoints, int32_t space, void* stream);

/* The same stage with save_wav's scaling fused in (utils/audio.py:17-19, applied by eval.py:43 / train.py:108 to the trimmed
 * waveform): flags NSB_SYNTH_PEAK_NORMALIZE multiplies utterance b's first endpoints[b] samples by
 * 32767 / max(0.01, max |wav[:endpoints[b]]|) (float64, like numpy) and zeroes the rest; out_dtype NSB_F64, or NSB_I16 = the
 * scaled samples through the C cast numpy's astype(np.int16) performs (needs the flag; a quarter of the bytes to copy out). */
enum { NSB_SYNTH_PEAK_NORMALIZE = 1 };
int nsb_synthesize_ex(nsb_handle_t h, const float* spec, const int32_t* n_frames, int32_t batch, int32_t iters,
                      double threshold_db, double min_silence_sec, int32_t flags, void* wav_out, int32_t out_dtype,
                      int64_t* endpoints, int32_t space, void* stream);
/* save_wav's scaling alone (utils/audio.py:17-19): wav (NSB_F32 / NSB_F64, packed per utterance) times
 * 32767 / max(0.01, max |wav[:limit[b]]|) per utterance (limit NULL: the whole utterance; a DEVICE pointer for NSB_DEVICE);
 * out NSB_F64 or NSB_I16, samples at and beyond limit[b] are written as 0. */
int nsb_peak_normalize(nsb_handle_t h, const void* wav, int32_t wav_dtype, const int64_t* n_samples, int32_t batch,
                       const int64_t* limit, void* out, int32_t out_dtype, int32_t space, void* stream);

/* The feeder's bucketed batches in one pass (datasets/datafeeder.py:130-158: a group of batch_size * batch_group_size
 * examples is sorted by length, cut into batches, every batch padded on its own): feature row k of utterance b goes to row
 * row_off[b] + k (HOST int64 array) of lin_out [total_rows][num_freq] / mel_out [total_rows][num_mels]; all other rows are
 * zeroed (_pad = 0, datafeeder.py:216).  With row_off[b] = b * rows_per_utt this is nsb_features_padded. */
int nsb_features_rows(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch, const int64_t* row_off,
                      int64_t total_rows, float* lin_out, float* mel_out, int32_t space, void* stream);

/* Asynchronous NSB_HOST calls.  submit returns at once with a ticket; the call runs on one of the handle's worker slots
 * (default 3, nsb_set_async_slots before the first submit), each with private streams and workspaces, so consecutive
 * batches overlap on the GPU and on PCIe.  Data buffers must stay valid until nsb_wait(ticket), which returns the call's
 * status (nsb_last_error has its message); lengths arrays are copied at submit.  Tickets may be waited for in any order,
 * each exactly once.  nsb_destroy runs whatever is still queued.  Caller pattern: the feeder threads of
 * datasets/datafeeder.py:110-152; eval.py:36-59 synthesising sentence after sentence. */
int nsb_griffin_lim_submit(nsb_handle_t h, const float* spec, int32_t layout, const int32_t* n_frames, int32_t batch,
                           const float* init_phase_complex, uint64_t seed, int32_t iters, int32_t flags,
                           void* wav_out, int32_t out_dtype, uint64_t* ticket);
int nsb_features_submit(nsb_handle_t h, const float* wav, const int64_t* n_samples, int32_t batch,
                        float* lin_out, float* mel_out, uint64_t* ticket);
int nsb_synthesize_submit(nsb_handle_t h, const float* spec, const int32_t* n_frames, int32_t batch, int32_t iters,
                          double threshold_db, double min_silence_sec, int32_t flags, void* wav_out, int32_t out_dtype,
                          int64_t* endpoints, uint64_t* ticket);
int nsb_wait(nsb_handle_t h, uint64_t ticket);
int nsb_set_async_slots(nsb_handle_t h, int32_t n);

/* device memory owned by the library, for results handed to other frameworks through DLPack / the CUDA array interface
 * (nspeech_b200/_buffers.py); nsb_device_copy kind: 1 host->device, 2 device->host, 3 device->device, synchronous */
int nsb_device_alloc(int device, uint64_t bytes, void** out);
int nsb_device_free(int device, void* p);
int nsb_device_copy(int device, void* dst, const void* src, uint64_t bytes, int32_t kind);

/* page-locked host buffers for the NSB_HOST entry points (plain cudaHostAlloc / cudaFreeHost): copies from
 * pageable numpy memory are staged by the driver and reach only a fraction of PCIe bandwidth */
int nsb_alloc_pinned(uint64_t bytes, void** out);
int nsb_free_pinned(void* p);

/* tuning / accounting hooks (not in the reference) */
int nsb_set_tile_hops(nsb_handle_t h, int32_t tile_hops);       /* 0 = automatic */
int nsb_set_host_chunks(nsb_handle_t h, int32_t n);             /* pipelining of NSB_HOST calls (Griffin-Lim, features, STFT): 0 = automatic (Griffin-Lim: wave schedule on long batches, else a few chunks; analysis: ~24 MB of results per chunk), n = force n equal chunks of whole utterances */
int nsb_set_generic_iteration(nsb_handle_t h, int32_t on);     /* A/B hook: Griffin-Lim iterations with 0 = k_gl_stream (production), 1 = the generic k_synth<SRC_Y>, 2 = the tile kernel k_gl_iter */
enum {                          /* nsb_set_option keys: A/B switches of the iteration kernels (defaults are the production settings) */
    NSB_OPT_STREAM_SYNC_MODE = 1,   /* k_gl_stream: 2 = CTA barrier per colour step (default), 10 = the same barrier split-phase (publish after the overlap-add, wait before the next one); 0 none, 1 per round, 3 per half CTA (+4: no pacing) use the event counters */
    NSB_OPT_FUSE_ITERATIONS = 2,    /* 1 = all iterations of a call in one launch (default), 0 = one launch per iteration */
    NSB_OPT_WIDE_MODE = 3,          /* k_gl_iter one-frame-per-warp tiles: -1 automatic for a few utterances (default), 0 off, 1 forced */
    NSB_OPT_OVERLAP_CHUNKS = 4,     /* NSB_HOST Griffin-Lim chunk pipeline: 1 = consecutive chunks on two streams so that tails and ramps overlap, 0 = one stream (default) */
    NSB_OPT_MEL_LINES = 6,          /* mel projection: 3 = filters as line segments, two moments per band, on a magnitude row with one pad word per 32 bins, the long segments cut in two so that the lanes run equally long (default when representable), 2 = whole segments, 1 = 2 on the plain row, 0 = sparse rows */
    NSB_OPT_SPECIALIZE = 7,         /* Griffin-Lim iteration kernels: 1 = at the reference's default hparams (hop 250, window 1000, n_fft 2048) run the instantiations with that geometry as immediates (default), 0 = the general ones; results are identical bit for bit */
    NSB_OPT_WAVE_SCHEDULE = 5       /* NSB_HOST Griffin-Lim on long batches: 1 = wave schedule (chunk g joins at wave g, every launch runs all chunks in flight; default), 0 = plain chunk pipeline */
};
int nsb_set_option(nsb_handle_t h, int32_t key, int32_t value);
int nsb_set_stream_grid(nsb_handle_t h, int32_t ctas);          /* k_gl_stream grid: 0 = automatic, n = force n CTAs (tests: long pieces, many rounds of the ring) */
int nsb_stream_trace(nsb_handle_t h, int32_t enable, uint64_t* out, int32_t max_ctas);   /* profiling: per-CTA (SM id, start ns, end ns) of the last k_gl_stream launch -> number of CTAs written */
uint64_t nsb_kernel_launches(nsb_handle_t h);                    /* kernels launched through this handle so far */
/* the Griffin-Lim iteration kernel alone on device-resident state, for roofline timing: runs `iters`
 * iterations on the state left by the last nsb_griffin_lim(NSB_DEVICE) call */
int nsb_griffin_lim_iterate(nsb_handle_t h, int32_t iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NSPEECH_B200_H */
