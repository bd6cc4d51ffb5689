/* b200rx — C ABI of the B200-native batched 802.11a receive hot path.
 *
 * Drop-in boundary for bmorgan5/fun_ofdm's receiver path after timing_sync: the four blocks
 *     fft_symbols -> channel_est -> phase_tracker -> frame_decoder
 * (reference src/receiver_chain.cpp:33-36 creates them, :47-50 chains them, :106-126 drives them)
 * are replaced by one batched call that takes already-synchronised frames (the samples from each
 * frame's LTS1 tag onwards) and returns payload bytes, rate, length and a status per frame.  One step wider, the
 * b200rx_receive* entry points also replace frame_detector and timing_sync (receiver_chain.cpp:31-32, 45-46), i.e. all of
 * receiver_chain::process_samples for a contiguous capture of raw samples.
 *
 * Plain pointers and sizes only; no C++/torch types; never throws; never keeps a caller pointer
 * after the call returns, except where an entry point says so (the asynchronous ones: b200rx_submit_batch until
 * b200rx_wait, b200rx_pass_put until b200rx_pass_scan, b200rx_pass_decode until b200rx_pass_wait).  Every entry point returns 0 on success or a negative B200RX_E_* code;
 * b200rx_last_error() gives the text.  There is NO CPU fallback: without a CUDA device of compute
 * capability 10.x b200rx_create() fails with B200RX_E_DEVICE.
 *
 * The library (fun_ofdm_b200/lib/libb200rx.so) is built by `python -c "import __graft_entry__ as g; g.build()"`
 * or `make -C fun_ofdm_b200/csrc`.
 */
#ifndef B200RX_H
#define B200RX_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200RX_API __attribute__((visibility("default")))
#else
#define B200RX_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- return codes ---- */
#define B200RX_OK 0
#define B200RX_E_ARG (-1)    /* bad argument (NULL, zero size, exceeds the handle's limits) */
#define B200RX_E_CUDA (-2)   /* a CUDA call failed; see b200rx_last_error() */
#define B200RX_E_NOMEM (-3)  /* device or pinned-host allocation failed */
#define B200RX_E_DEVICE (-4) /* no usable sm_100 device */

/* ---- per-frame status byte ---- (the reference only drops such frames silently or with a stderr
 * line: frame_decoder.cpp:78, ppdu.cpp:187-203, ppdu.cpp:274-279) */
#define B200RX_ST_OK 0          /* header valid, CRC-32 matches: payload delivered */
#define B200RX_ST_HDR_PARITY 1  /* SIGNAL parity check failed (ppdu.cpp:187-191) */
#define B200RX_ST_HDR_RATE 2    /* rate field not in VALID_RATES (ppdu.cpp:197-203, rates.h:21) */
#define B200RX_ST_CRC_FAIL 3    /* CRC-32 mismatch (ppdu.cpp:274-279) */
#define B200RX_ST_TRUNCATED 4   /* fewer samples than 128 + 80*(1 + nsym) were supplied for the frame */
#define B200RX_ST_TOO_LONG 5    /* decoded LENGTH exceeds the handle's max_payload_bytes */
#define B200RX_ST_NO_FRAME 255  /* b200rx_receive*: output slot beyond the number of frames found in the capture */

/* ---- fun::Rate enum values (reference src/rates.h:31-44), as written to rate_out ---- */
#define B200RX_RATE_1_2_BPSK 0
#define B200RX_RATE_2_3_BPSK 1
#define B200RX_RATE_3_4_BPSK 2
#define B200RX_RATE_1_2_QPSK 3
#define B200RX_RATE_2_3_QPSK 4
#define B200RX_RATE_3_4_QPSK 5
#define B200RX_RATE_1_2_QAM16 6
#define B200RX_RATE_2_3_QAM16 7
#define B200RX_RATE_3_4_QAM16 8
#define B200RX_RATE_2_3_QAM64 9
#define B200RX_RATE_3_4_QAM64 10
#define B200RX_RATE_INVALID 255

typedef struct b200rx_handle b200rx_handle;

/* Capacity of a handle.  Device scratch (branch metrics 4 B and survivor words 8 B per trellis
 * step per frame) is sized from these once, at create time. */
typedef struct {
    uint32_t max_frames;        /* frames per decode call */
    uint32_t max_payload_bytes; /* largest LENGTH to decode, <= 4095 (12-bit field, ppdu.cpp:195) */
    uint32_t reserved[6];       /* zero */
} b200rx_limits;

/* Optional taps for parity tests (DEVICE pointers; any may be NULL).  Not a product feature. */
typedef struct {
    double *equalized;          /* [n_frames][eq_vectors][48][2]: phase_tracker output (SIGNAL first) */
    uint32_t eq_vectors;        /* vectors reserved per frame */
    uint8_t *decoded;           /* [n_frames][decoded_stride]: Viterbi output bytes before descrambling */
    uint32_t decoded_stride;
    uint32_t *header_field;     /* [n_frames]: 24-bit decoded SIGNAL field */
    uint8_t *depunct;           /* [n_frames][depunct_stride]: depunctured soft symbols, 2 per trellis step */
    uint32_t depunct_stride;
} b200rx_debug;

/* Device time of each stage of the most recent batch (CUDA events on the handle's stream). */
typedef struct {
    float frontend_ms;  /* CP strip + FFT + channel estimate + equalise + phase track + demap + branch metrics */
    float viterbi_ms;   /* add-compare-select */
    float traceback_ms; /* traceback + descramble + CRC-32 + payload write */
    float total_ms;
    uint64_t trellis_steps; /* sum over frames of nsym * dbps actually decoded */
    uint32_t frames_ok;
    uint32_t frames_failed;
    uint64_t payload_bytes; /* of CRC-OK frames */
    uint64_t traceback_rewalks; /* traceback tiles whose speculative pre-roll had not merged and were redone */
} b200rx_stats;

/* Lifetime.  One handle per GPU per host thread; a handle is not re-entrant. */
B200RX_API int b200rx_create(int device, const b200rx_limits *limits, b200rx_handle **out);
B200RX_API int b200rx_destroy(b200rx_handle *h);
B200RX_API const char *b200rx_last_error(const b200rx_handle *h); /* h may be NULL: last create() error */
B200RX_API const char *b200rx_version(void);

/* Use the caller's CUDA stream (a cudaStream_t passed as void*) instead of the handle's own.
 * NULL restores the handle-owned stream. */
B200RX_API int b200rx_set_stream(b200rx_handle *h, void *cuda_stream);
B200RX_API int b200rx_synchronize(b200rx_handle *h);

/* ---- sample format of every `iq` argument (wire-format ingestion) ----
 * FC64 (default) is the reference's own std::complex<double> (tagged_vector.h:82-94; usrp.cpp:43-44 asks UHD for the cpu
 * format "fc64", i.e. UHD widens its 16-bit over-the-wire samples on the host).  FC32 and SC16 are what the same samples
 * are before that widening: taking them as they are halves / quarters the bytes that cross PCIe, which is what bounds
 * the host-buffer entry points.  Samples are
 * widened to double in the kernels' loads - (double)float exactly, (double)int16 * sc16_scale with one rounding - and
 * everything after that is the same fp64 arithmetic, so results are bit-identical to the reference fed the widened
 * samples.  Applies to all later calls on the handle; `iq` pointers are then float[2] / int16_t[2] per sample. */
#define B200RX_FMT_FC64 0
#define B200RX_FMT_FC32 1
#define B200RX_FMT_SC16 2
/* The reference's own element type between timing_sync and fft_symbols: fun::tagged_sample (tagged_vector.h:82-94),
 * { std::complex<double> sample; vector_tag tag; } = 24 bytes.  Accepted by the decode entry points (the tags are
 * ignored there, lts1_index says where frames start) and by b200rx_pass_scan_tagged (which reads them); the raw-capture
 * entry points do not take it - a tagged stream has been through frame_detector and timing_sync already. */
#define B200RX_FMT_TAGGED_FC64 3
B200RX_API int b200rx_set_sample_format(b200rx_handle *h, int format, double sc16_scale);

/* Implementation knobs, per handle (nothing is read from the environment).  Results never depend on them; they select
 * between kernel variants and pipeline geometries that are bit-identical by construction and by test.  Keys:
 *   "acs_gen"       3 (default): add-compare-select on soft-symbol pairs, two frames per register; 2: round-1 kernel
 *   "acs_lb"        log2 of the lanes per frame (gen 2: 2..5) / per frame pair (gen 3: 2..3); 0 = from the batch size
 *   "acs_warps"     warps per ACS CTA (1, 2, 4); 0 = default
 *   "acs_rn"        renormalisation variant: gen 2 cross-lane minimum in 1 (0) or 2 (1) lane bits per round; gen 3
 *                   subtract the minimum (0) or keep a per-frame offset (1, default)
 *   "h2d_chunk", "h2d_chunk_min"   frames per pipelined chunk of b200rx_submit_batch
 *   "pull_mode"     host-buffer ingest: 0 DMA copy, 1 GPU pull from pinned memory, k >= 2 every k-th chunk by DMA and the
 *                   rest pulled, -1 by sample format (fc64: pull, narrower formats: DMA)
 *   "fe_split"      front end: 1 header kernel + data kernel with 8 lanes per OFDM symbol (default), 0 one CTA per frame
 *   "scan_graph"    b200rx_pass_scan: 1 replay the scan's launches as one CUDA graph (default), 0 issue them one by one
 * The call drains the handle first. */
B200RX_API int b200rx_set_tuning(b200rx_handle *h, const char *key, int64_t value);

#define B200RX_MAX_PIPELINE_DEPTH 12

/* Pipelining of consecutive b200rx_decode_batch_dev calls.  depth = 1 (default): every call runs in order on
 * the handle's stream.  depth = 2 .. B200RX_MAX_PIPELINE_DEPTH: calls rotate over `depth` lanes, each with its own scratch set and
 * stream, so that batch j+1 is already in its front end while batch j is still in its Viterbi kernel (the
 * Viterbi kernel of one 4096-frame batch cannot fill a B200: +32 % / +43 % / +63 % throughput measured at depth 2 / 3 / 6).
 * A call still only reads its inputs after everything queued before it on the handle's stream, but its
 * results are ordered on that stream only after b200rx_join(); calls that may overlap (any `depth` consecutive
 * ones) must be given distinct output buffers.  b200rx_join(h, 0) makes the handle's stream wait for every
 * call issued so far; b200rx_join(h, k) with 0 < k < depth for the call issued k calls before the latest one.
 * b200rx_synchronize() always waits for everything. */
B200RX_API int b200rx_set_pipeline_depth(b200rx_handle *h, uint32_t depth);
B200RX_API int b200rx_join(b200rx_handle *h, uint32_t calls_back);
/* Same, but the waiting stream is `cuda_stream` (e.g. a communication stream that gathers a finished batch's
 * status while the handle's stream keeps issuing new batches). */
B200RX_API int b200rx_join_on(b200rx_handle *h, uint32_t calls_back, void *cuda_stream);

/* Pinned host memory for the host-buffer entry point (plain malloc'ed memory also works, slower). */
B200RX_API int b200rx_host_alloc(void **ptr, size_t bytes);
B200RX_API int b200rx_host_free(void *ptr);
/* 1 if ptr lies in pinned (page-locked) host memory known to CUDA, 0 if not. */
B200RX_API int b200rx_host_is_pinned(const void *ptr);

/* Replaces fft_symbols::work + channel_est::work + phase_tracker::work + frame_decoder::work
 * (fft_symbols.cpp:33-80, channel_est.cpp:36-85, phase_tracker.cpp:70-105, frame_decoder.cpp:45-91,
 * ppdu.cpp:168-295) for n_frames isolated frames.  HOST buffers; copies in, decodes, copies out,
 * returns when the results are in the caller's arrays.
 *
 *   iq            interleaved (re, im) doubles == std::complex<double>[] (the `sample` members of the
 *                 tagged_sample stream, tagged_vector.h:82-94), or the format set by b200rx_set_sample_format;
 *                 iq_samples complex samples in total.  If the buffer is pinned (b200rx_host_alloc, cudaHostRegister)
 *                 the GPU reads it in place and fetches only the samples the path uses (no cyclic prefixes).
 *   lts1_index[f] index into iq of the sample timing_sync tagged LTS1 for frame f (timing_sync.cpp:105);
 *                 LTS2 is implied 64 samples later (timing_sync.cpp:106)
 *   avail[f]      complex samples available to frame f from lts1_index[f] on
 *   payload_out   [n_frames][payload_stride] bytes; frame f's payload (LENGTH bytes) starts at f*payload_stride
 *   payload_len   [n_frames] decoded LENGTH (0 if the header failed)
 *   rate_out      [n_frames] fun::Rate value or B200RX_RATE_INVALID
 *   status        [n_frames] B200RX_ST_*
 */
B200RX_API int b200rx_decode_batch(b200rx_handle *h, const void *iq, uint64_t iq_samples,
                        const uint64_t *lts1_index, const uint32_t *avail, uint32_t n_frames,
                        uint8_t *payload_out, uint32_t payload_stride,
                        uint16_t *payload_len, uint8_t *rate_out, uint8_t *status);

/* The same call split in two, so that several batches are in flight and the PCIe link never idles: submit queues
 * H2D + decode + D2H and returns a ticket; b200rx_wait(h, ticket) returns when that call's outputs are in the caller's
 * arrays (ticket 0: every outstanding call).  Every buffer passed to submit - inputs and outputs - must stay valid and
 * untouched until the wait.  At most B200RX_MAX_INFLIGHT calls overlap (a further submit first waits for the oldest);
 * each call in flight owns a full set of device scratch.  b200rx_decode_batch == submit + wait. */
#define B200RX_MAX_INFLIGHT 3
B200RX_API int b200rx_submit_batch(b200rx_handle *h, const void *iq, uint64_t iq_samples,
                        const uint64_t *lts1_index, const uint32_t *avail, uint32_t n_frames,
                        uint8_t *payload_out, uint32_t payload_stride,
                        uint16_t *payload_len, uint8_t *rate_out, uint8_t *status, uint64_t *ticket);
B200RX_API int b200rx_wait(b200rx_handle *h, uint64_t ticket);

/* Same contract with every array already in DEVICE memory; asynchronous on the handle's stream
 * (call b200rx_synchronize or sync the stream yourself).  dbg may be NULL. */
B200RX_API int b200rx_decode_batch_dev(b200rx_handle *h, const void *iq_dev, uint64_t iq_samples,
                            const uint64_t *lts1_index_dev, const uint32_t *avail_dev, uint32_t n_frames,
                            uint8_t *payload_out_dev, uint32_t payload_stride,
                            uint16_t *payload_len_dev, uint8_t *rate_out_dev, uint8_t *status_dev,
                            const b200rx_debug *dbg);

/* ---- frame detection + timing synchronisation + decode from a raw sample stream ----
 * Replaces frame_detector::work (frame_detector.cpp:41-92) and timing_sync::work (timing_sync.cpp:51-139) in
 * front of the four hot-path blocks, i.e. together with the decode it is receiver_chain::process_samples
 * (receiver_chain.cpp:106-126) for one contiguous capture.  Tags are the reference's vector_tag values
 * (tagged_vector.h:25-34: 1 STS_START, 2 STS_END, 4 LTS1, 5 LTS2), indexed by the sample they sit on (the
 * reference's timing_sync output is the same stream delayed by 160 samples).  STS_END tags in the last 160
 * samples are left for the next capture, as in the reference (timing_sync.cpp:68). */
typedef struct b200rx_sync_result {
    uint32_t n_events;    /* STS_END tags examined */
    uint32_t n_frames;    /* LTS1 tags placed = frames handed to the decoder, in stream order */
    uint32_t overflow;    /* events / frames dropped because the handle's max_frames was too small */
    uint32_t reserved;
    double last_phase;    /* m_phase_acc after the capture (timing_sync.h:45); feed it to the next call */
    uint32_t phase_valid; /* 0: no frame was found, the phase is still phase_in */
    uint32_t pad;
} b200rx_sync_result;

/* Chunked streams.  The reference's timing_sync::work() sees the caller's chunk behind 160 carried-over samples and
 * discards an LTS whose start would lie before that buffer (timing_sync.cpp:102 `if(lts_offset < 0) break;`), so what
 * it finds depends on where the caller cut the stream.  One capture is treated as one such buffer.  A caller that
 * replays a chunked stream in larger captures (fun::b200_receiver_chain does) passes, before the call, the sample
 * indices - relative to the capture, ascending, possibly negative - at which the reference's buffers would have started
 * (chunk start - 160); the STS_END tag at x is then judged against the last origin <= x.  Consumed by the next
 * b200rx_sync_dev / b200rx_receive / b200rx_receive_dev call. */
B200RX_API int b200rx_set_receive_origins(b200rx_handle *h, const int64_t *origins, uint32_t n);

/* Tags and frame list only.  tags_dev [n_samples] (nullable); lts1_index_dev / avail_dev [max_frames] as consumed by
 * b200rx_decode_batch_dev; phase_dev [max_frames] (nullable) the rotation phase of each frame.  phase_in is
 * m_phase_acc before the capture (0 for a fresh chain).  Synchronous: returns when `res` (host) is filled. */
B200RX_API int b200rx_sync_dev(b200rx_handle *h, const void *iq_dev, uint64_t n_samples, double phase_in,
                               uint8_t *tags_dev, uint64_t *lts1_index_dev, uint32_t *avail_dev, double *phase_dev,
                               b200rx_sync_result *res);

/* Raw samples in HBM -> payloads: detection, synchronisation, phase rotation and the hot path, with no host
 * round trip in between (the decode kernels are launched for max_frames slots and read the number of frames
 * found from device memory).  Outputs are device arrays sized for max_frames: slots [0, n_frames) hold the frames
 * in stream order, the rest get status B200RX_ST_NO_FRAME.  lts1_out_dev [max_frames] and n_frames_dev [1] are
 * optional.  res != NULL: the call returns when the summary is on the host (outputs complete).  res == NULL: fully
 * asynchronous, pipelined over the lanes of b200rx_set_pipeline_depth like b200rx_decode_batch_dev. */
B200RX_API int b200rx_receive_dev(b200rx_handle *h, const void *iq_dev, uint64_t n_samples, double phase_in,
                                  uint8_t *payload_out_dev, uint32_t payload_stride, uint16_t *payload_len_dev,
                                  uint8_t *rate_out_dev, uint8_t *status_dev, uint64_t *lts1_out_dev,
                                  uint32_t *n_frames_dev, b200rx_sync_result *res);

/* Same with host buffers (pinned buffers recommended); synchronous. */
B200RX_API int b200rx_receive(b200rx_handle *h, const void *iq, uint64_t n_samples, double phase_in,
                              uint8_t *payload_out, uint32_t payload_stride, uint16_t *payload_len, uint8_t *rate_out,
                              uint8_t *status, uint64_t *lts1_out, b200rx_sync_result *res);

/* ---- the same, as a two-phase pass for streaming callers (fun::b200_receiver_chain) ----
 * receiver_chain::process_samples (receiver_chain.cpp:106-126) returns what its last block finished in EARLIER rounds: a
 * frame surfaces up to five calls after the one that brought its samples (:118-125 swap the buffers one block down per
 * call).  A pass uses that slack.  Phase one (open / put / scan) is short and synchronous - detection, synchronisation
 * and SIGNAL decode of every frame in the capture, i.e. everything the caller's streaming state depends on (which
 * frames exist, where, how long they are, m_phase_acc).  Phase two (decode) is the long dependent chain - data
 * symbols, Viterbi, descrambler, CRC - and runs asynchronously: the caller collects it during a later call.  Up to
 * `pipeline depth` passes (b200rx_set_pipeline_depth, 1..12) are in flight, each on its own lane (stream, scratch,
 * sample staging), so single-frame Viterbi latency (12 096 dependent steps ~ 0.7 ms) overlaps with the following calls.
 *
 *   b200rx_pass_open    takes the next lane (waits for the pass that used it `depth` passes ago) and empties its staging
 *   b200rx_pass_put     appends n_samples samples (handle's sample format) to the staging, asynchronously; pinned memory
 *                       is copied by the DMA engine while the caller prepares the next slice; `iq` must stay untouched
 *                       until b200rx_pass_scan returns
 *   b200rx_pass_scan    frame_detector + timing_sync + ppdu::decode_header over the staged capture; returns when the
 *                       frame list is on the host: frames[f] for f < res->n_frames (at most frames_cap are written).
 *                       status: B200RX_ST_OK (header valid, all 128 + 80 * (1 + nsym) samples present), HDR_PARITY,
 *                       HDR_RATE, TOO_LONG, or TRUNCATED (frame cut short by the next LTS1 or by the end of the capture)
 *   b200rx_pass_decode  decodes the frames with select[f] != 0 (others are skipped); asynchronous.  Outputs are indexed
 *                       by frame like b200rx_receive's and must stay valid until the pass is waited for; allocate them
 *                       with b200rx_host_alloc.  A pass whose decode is never requested needs no wait.
 *   b200rx_pass_poll    1 = that pass has finished (outputs are in place), 0 = still running
 *   b200rx_pass_wait    blocks until it has
 * b200rx_set_receive_origins applies to the next b200rx_pass_scan.  Passes and the other entry points of a handle must
 * not be interleaved without b200rx_synchronize in between. */
typedef struct b200rx_pass_frame {
    uint64_t lts1;    /* capture index of the sample tagged LTS1 */
    uint32_t avail;   /* samples from there to the next frame's LTS1 / the end of the capture */
    uint16_t length;  /* LENGTH field (0 when the header failed) */
    uint8_t rate;     /* fun::Rate or B200RX_RATE_INVALID */
    uint8_t status;   /* header verdict, see above */
    uint64_t sts_end; /* capture index of the STS_END tag that led to this frame (b200rx_pass_scan; 0 for tagged passes) */
    double phase;     /* m_phase_acc as this frame's synchronisation left it (timing_sync.cpp:114-115; 0 for tagged passes) */
} b200rx_pass_frame;
B200RX_API int b200rx_pass_open(b200rx_handle *h);
B200RX_API int b200rx_pass_put(b200rx_handle *h, const void *iq, uint64_t n_samples);
B200RX_API int b200rx_pass_scan(b200rx_handle *h, double phase_in, b200rx_pass_frame *frames, uint32_t frames_cap,
                                b200rx_sync_result *res);
/* Phase one for a stream that timing_sync has tagged already (fun::b200_rx, the block that replaces fft_symbols ..
 * frame_decoder): the handle's sample format must be B200RX_FMT_TAGGED_FC64; a frame starts at every sample whose tag is
 * LTS1 (fft_symbols.cpp:42-51) and extends to the next such sample or the end of the staged stream.  Same outputs as
 * b200rx_pass_scan (res->n_events = LTS1 tags seen, res->overflow = those beyond max_frames); the structs are unpacked
 * on the GPU, the host never walks them. */
B200RX_API int b200rx_pass_scan_tagged(b200rx_handle *h, b200rx_pass_frame *frames, uint32_t frames_cap, b200rx_sync_result *res);
B200RX_API int b200rx_pass_decode(b200rx_handle *h, const uint8_t *select, uint8_t *payload_out, uint32_t payload_stride,
                                  uint8_t *status, uint64_t *ticket);
B200RX_API int b200rx_pass_poll(b200rx_handle *h, uint64_t ticket);
B200RX_API int b200rx_pass_wait(b200rx_handle *h, uint64_t ticket);

/* Replaces ppdu::decode_header (ppdu.cpp:168-218) for n_frames frames: only the two LTS symbols and the
 * SIGNAL symbol are read (208 samples from lts1_index[f]); gives the streaming adapter the frame length
 * before the frame has fully arrived.  HOST buffers, synchronous.  status: B200RX_ST_OK (header valid; the
 * frame needs 128 + 80 * (1 + nsym) samples), HDR_PARITY, HDR_RATE, TOO_LONG, or TRUNCATED (< 208 samples). */
B200RX_API int b200rx_decode_headers(b200rx_handle *h, const void *iq, uint64_t iq_samples,
                                     const uint64_t *lts1_index, const uint32_t *avail, uint32_t n_frames,
                                     uint16_t *payload_len, uint8_t *rate_out, uint8_t *status);

/* Replaces viterbi::conv_decode (viterbi.cpp:31-37: alloc/init/FULL_SPIRAL/chainback) for a batch.
 * DEVICE buffers.  symbols: depunctured soft symbols (0..255, erasure 127), frame f at
 * symbols_dev + f*symbols_stride, 2*(data_bits[f] + 6) bytes each.  out: frame f's decoded bytes at
 * out_dev + f*out_stride, ceil(data_bits[f] / 8) bytes, MSB first.  data_bits[f] + 6 must be even
 * (it is for every rate: viterbi.cpp:209 processes two steps per pass) and <= the handle's capacity. */
B200RX_API int b200rx_viterbi_batch_dev(b200rx_handle *h, const uint8_t *symbols_dev, uint64_t symbols_stride,
                             const uint32_t *data_bits_dev, uint32_t max_data_bits, uint32_t n_frames,
                             uint8_t *out_dev, uint32_t out_stride);

/* Counters and per-stage device times of the most recent decode call on this handle
 * (synchronises the handle's stream). */
B200RX_API int b200rx_get_stats(b200rx_handle *h, b200rx_stats *out);

/* Per-stage device times summed over the next `slots` decode calls (one CUDA-event quadruple per call,
 * recorded on the handle's stream, so the kernels are timed in place inside whatever region the
 * caller is timing).  b200rx_profile_read synchronises the stream, returns the number of calls
 * recorded and the three sums in milliseconds, and ends the profiling window. */
B200RX_API int b200rx_profile_begin(b200rx_handle *h, uint32_t slots);
B200RX_API int b200rx_profile_read(b200rx_handle *h, uint32_t *calls, float *frontend_ms, float *viterbi_ms,
                                   float *traceback_ms);

/* DEVICE address of the four 64-bit counters of the most recent decode call (with a pipeline depth > 1 every
 * lane has its own: ask right after the call)
 * {frames ok, frames failed, payload bytes of ok frames, trellis steps}: lets a multi-GPU driver
 * reduce them with one collective (NCCL all-reduce) without a host round trip. */
B200RX_API int b200rx_device_counters(b200rx_handle *h, void **dev_ptr);
/* Copies those four counters of the most recent device-buffer call to dst_dev (4 x uint64, device memory) in stream order
 * right behind that call - before its lane can be given the next batch, which resets them - and inside what b200rx_join*
 * waits for.  The race-free way to keep per-call counters when calls are pipelined. */
B200RX_API int b200rx_copy_counters(b200rx_handle *h, void *dst_dev);

/* ---- several GPUs of one box from one process ----
 * Frames are independent (all per-frame state is rebuilt from the frame's own LTS and SIGNAL symbols), so a batch is
 * cut into contiguous shards, one per device, and there is no exchange on the data path.  A group owns one handle and
 * one host thread per device; group calls fan out to the threads and return when every device has queued (device-buffer
 * calls) or finished (host-buffer calls) its share.  The only collective is the gather of the per-frame status bytes
 * and the sum of the counters after a decode (NCCL over NVLink: ncclCommInitAll, one communicator and one communication
 * stream per device, ordered behind that device's decode by an event).  NCCL is loaded on first use (libnccl.so.2);
 * everything except b200rx_gather_status works without it. */
typedef struct b200rx_group b200rx_group;
B200RX_API int b200rx_group_create(const int *devices, uint32_t n_devices, const b200rx_limits *limits_per_device,
                                   b200rx_group **out);
B200RX_API int b200rx_group_destroy(b200rx_group *g);
B200RX_API uint32_t b200rx_group_size(const b200rx_group *g);
B200RX_API b200rx_handle *b200rx_group_handle(b200rx_group *g, uint32_t i); /* for per-device settings (format, tuning, depth) */
B200RX_API const char *b200rx_group_last_error(const b200rx_group *g);      /* g may be NULL: last group_create() error */
B200RX_API int b200rx_group_synchronize(b200rx_group *g);

/* Shard plan: first[i] .. first[i+1] is device i's range of frames (first has n_devices + 1 entries), contiguous and
 * balanced by weight[f] (pass avail[]: a frame's samples are proportional to its trellis steps for a given rate; NULL =
 * equal weights). */
B200RX_API int b200rx_group_plan(const b200rx_group *g, const uint32_t *weight, uint32_t n_frames, uint32_t *first);

/* b200rx_decode_batch over all devices of the group: same arguments, same outputs (frame f's results at index f); every
 * device decodes its shard of the plan from the caller's host buffers with its own copy pipeline.  Returns when all
 * outputs are in place.  n_frames may be up to n_devices * max_frames as long as no shard exceeds max_frames. */
B200RX_API int b200rx_group_decode_batch(b200rx_group *g, const void *iq, uint64_t iq_samples, const uint64_t *lts1_index,
                                         const uint32_t *avail, uint32_t n_frames, uint8_t *payload_out,
                                         uint32_t payload_stride, uint16_t *payload_len, uint8_t *rate_out, uint8_t *status);

/* b200rx_decode_batch_dev on every device at once: argument i of each array belongs to device i (pointers into that
 * device's memory).  Asynchronous like the single-device call; b200rx_group_synchronize or b200rx_gather_status order
 * the results. */
B200RX_API int b200rx_group_decode_batch_dev(b200rx_group *g, const void *const *iq_dev, const uint64_t *iq_samples,
                                             const uint64_t *const *lts1_index_dev, const uint32_t *const *avail_dev,
                                             const uint32_t *n_frames, uint8_t *const *payload_out_dev, uint32_t payload_stride,
                                             uint16_t *const *payload_len_dev, uint8_t *const *rate_out_dev,
                                             uint8_t *const *status_dev);

/* After a decode on every device: all-gather of frames_per_device status bytes from each device's status_dev[i] into
 * gathered_dev[i] (n_devices * frames_per_device bytes on EVERY device, device-major; entries may be NULL to skip the
 * gather) and all-reduce of the four counters of each device's most recent call {frames ok, frames failed, payload bytes
 * of ok frames, trellis steps}, returned in counters_sum (host, 4 values; may be NULL).  Returns when both are complete. */
B200RX_API int b200rx_gather_status(b200rx_group *g, const uint8_t *const *status_dev, uint32_t frames_per_device,
                                    uint8_t *const *gathered_dev, uint64_t *counters_sum);

/* Number of kernels launched by this handle since creation (for gpu_launches accounting). */
B200RX_API uint64_t b200rx_launch_count(const b200rx_handle *h);

/* Trellis capacity (steps per frame) the handle was sized for. */
B200RX_API uint32_t b200rx_max_steps(const b200rx_handle *h);

/* Bytes per complex sample in the handle's current sample format (16, 8 or 4). */
B200RX_API size_t b200rx_sample_bytes(const b200rx_handle *h);

#ifdef __cplusplus
}
#endif

#endif /* B200RX_H */
