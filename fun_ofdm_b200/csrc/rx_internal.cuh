// Internal declarations shared by the b200rx translation units (kernels + C ABI).
// Device layout of everything that lives in HBM between kernels is defined here.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200rx.h"

namespace b200rx {

// ---- rate parameters (reference src/rates.h:52-196), indexed by fun::Rate value ----
enum { PUNC_1_2 = 0, PUNC_2_3 = 1, PUNC_3_4 = 2 };
enum { ACS_RN_DEFAULT = 1 };

struct RateRow { uint16_t cbps, dbps; uint8_t bpsc, punc, rate_field, pad; };

__host__ __device__ __forceinline__ RateRow rate_row(int r)
{
    // cbps, dbps, bpsc, punc, rate_field
    switch (r) {
        case 0: return {48, 24, 1, PUNC_1_2, 0xD, 0};
        case 1: return {48, 32, 1, PUNC_2_3, 0xE, 0};
        case 2: return {48, 36, 1, PUNC_3_4, 0xF, 0};
        case 3: return {96, 48, 2, PUNC_1_2, 0x5, 0};
        case 4: return {96, 64, 2, PUNC_2_3, 0x6, 0};
        case 5: return {96, 72, 2, PUNC_3_4, 0x7, 0};
        case 6: return {192, 96, 4, PUNC_1_2, 0x9, 0};
        case 7: return {192, 128, 4, PUNC_2_3, 0xA, 0};
        case 8: return {192, 144, 4, PUNC_3_4, 0xB, 0};
        case 9: return {288, 192, 6, PUNC_2_3, 0x1, 0};
        default: return {288, 216, 6, PUNC_3_4, 0x3, 0};
    }
}

// rate field (4 bits) -> fun::Rate, 255 when not in VALID_RATES (rates.h:21, :208-249)
__host__ __device__ __forceinline__ int rate_from_field(int f)
{
    switch (f & 0xF) {
        case 0xD: return 0; case 0xE: return 1; case 0xF: return 2;
        case 0x5: return 3; case 0x6: return 4; case 0x7: return 5;
        case 0x9: return 6; case 0xA: return 7; case 0xB: return 8;
        case 0x1: return 9; case 0x3: return 10;
        default: return 255;
    }
}

// ppdu.cpp:38-40: ceil((16 + 8*(len + 4) + 6) / dbps)
__host__ __device__ __forceinline__ uint32_t num_symbols(int rate, uint32_t length)
{
    const uint32_t dbps = rate_row(rate).dbps;
    return (16u + 8u * (length + 4u) + 6u + dbps - 1u) / dbps;
}

// Largest trellis (steps) any rate needs for a payload of max_len bytes, rounded up to a multiple of 96
// (the ACS kernel works in blocks of 24 steps, metric words are moved 4 at a time).
inline uint32_t max_steps_for(uint32_t max_len)
{
    uint32_t m = 0;
    for (int r = 0; r < 11; r++) {
        uint32_t s = num_symbols(r, max_len) * rate_row(r).dbps;
        if (s > m) m = s;
    }
    return ((m + 95u) / 96u) * 96u;
}

// ---- per-frame descriptor written by the front end, read by the Viterbi and traceback kernels ----
struct FrameDesc {
    uint32_t n_steps;   // trellis steps = nsym * dbps (0 when the frame is not decoded)
    uint32_t data_bits; // n_steps - 6: bits produced by the traceback (viterbi.cpp:31-37)
    uint32_t field;     // 24-bit SIGNAL field as decoded
    uint16_t length;    // LENGTH
    uint8_t rate;       // fun::Rate or 255
    uint8_t status;     // B200RX_ST_* so far (traceback turns OK into CRC_FAIL where needed)
};

// ---- branch-metric word (one per trellis step) ----
// byte c = ((s0 ^ b0) + (s1 ^ b1) + 1) >> 3 with b0 = (c & 2) ? 255 : 0, b1 = (c & 1) ? 255 : 0:
// the four distinct values pavgb/psrlw produce in viterbi.cpp:234-248 (Branchtab entries are 0 or 255).
__host__ __device__ __forceinline__ uint32_t bm_word(uint32_t s0, uint32_t s1)
{
#ifdef __CUDA_ARCH__
    // the same four values from two sums: with a = s0 + s1 + 1 and d = s0 - s1 + 256 (both 1..511),
    // m00 = a >> 3, m01 = d >> 3, m10 = (512 - d) >> 3, m11 = (512 - a) >> 3   (255 - s = s ^ 255)
    const uint32_t a = s0 + s1 + 1u;
    const uint32_t d = s0 - s1 + 256u;
    const uint32_t P = d * 65536u + a;
    const uint32_t Q = 0x02000200u - P;
    const uint32_t A = (P >> 3) & 0x003F003Fu; // m00 | m01 << 16
    const uint32_t B = (Q >> 3) & 0x003F003Fu; // m11 | m10 << 16
    return __byte_perm(A, B, 0x4620u);         // m00 | m01 << 8 | m10 << 16 | m11 << 24
#else
    const uint32_t n0 = s0 ^ 255u, n1 = s1 ^ 255u;
    const uint32_t m00 = (s0 + s1 + 1u) >> 3;
    const uint32_t m01 = (s0 + n1 + 1u) >> 3;
    const uint32_t m10 = (n0 + s1 + 1u) >> 3;
    const uint32_t m11 = (n0 + n1 + 1u) >> 3;
    return m00 | (m01 << 8) | (m10 << 16) | (m11 << 24);
#endif
}

// Branch class of butterfly j (0..31): Branchtab[0][j] = parity(2j & 121), Branchtab[1][j] = parity(2j & 91)
// (viterbi.cpp:87-91).  2j & 121 keeps j bits 2,3,4; 2j & 91 keeps j bits 0,2,3.
__host__ __device__ __forceinline__ uint32_t branch_class(uint32_t j)
{
    const uint32_t b0 = ((j >> 2) ^ (j >> 3) ^ (j >> 4)) & 1u;
    const uint32_t b1 = (j ^ (j >> 2) ^ (j >> 3)) & 1u;
    return (b0 << 1) | b1;
}

// ---- sample formats of the ingest side (b200rx_set_sample_format) ----
// FC64 is the reference's own std::complex<double> (tagged_vector.h:82-94, usrp.cpp:43-44 cpu format "fc64"); FC32 and SC16
// are the narrower formats the same samples have before UHD widens them on the host.  The widening to
// double happens in the load, exactly (float -> double) or with one rounding ((double)int16 * scale), so the kernels
// compute what the reference computes when it is handed the widened samples.
enum { FMT_FC64 = B200RX_FMT_FC64, FMT_FC32 = B200RX_FMT_FC32, FMT_SC16 = B200RX_FMT_SC16, FMT_TAGGED = B200RX_FMT_TAGGED_FC64 };

// FMT_TAGGED: the reference's own element type between timing_sync and fft_symbols, fun::tagged_sample
// (tagged_vector.h:82-94): { std::complex<double> sample; vector_tag tag; } = 16 + 4 + 4 padding bytes.
constexpr size_t TAGGED_SAMPLE_BYTES = 24;
constexpr int TAG_LTS1 = 4; // tagged_vector.h:25-34

__host__ __device__ __forceinline__ size_t sample_bytes(int fmt)
{
    return fmt == FMT_FC64 ? 16 : (fmt == FMT_FC32 ? 8 : (fmt == FMT_SC16 ? 4 : TAGGED_SAMPLE_BYTES));
}

#ifdef __CUDACC__
template <int FMT>
__device__ __forceinline__ double2 load_sample(const void *base, uint64_t i, double scale)
{
    if constexpr (FMT == FMT_FC64) {
        return reinterpret_cast<const double2 *>(base)[i];
    } else if constexpr (FMT == FMT_FC32) {
        const float2 v = reinterpret_cast<const float2 *>(base)[i];
        return make_double2((double)v.x, (double)v.y);
    } else if constexpr (FMT == FMT_TAGGED) { // 24-byte structs: 8-byte aligned, two loads
        const double *q = reinterpret_cast<const double *>(reinterpret_cast<const char *>(base) + i * TAGGED_SAMPLE_BYTES);
        return make_double2(q[0], q[1]);
    } else {
        const short2 v = reinterpret_cast<const short2 *>(base)[i];
        return make_double2(__dmul_rn((double)v.x, scale), __dmul_rn((double)v.y, scale));
    }
}
#endif

// ---- frame detection / timing synchronisation (sync.cu) ----
struct SyncRec {      // outcome of one STS_END event (timing_sync.cpp:71-118)
    uint64_t x;       // stream index of the STS_END tag
    int64_t lts1;     // stream index tagged LTS1 (valid when found)
    double2 rot;      // (cos, sin) of the phase set by this event
    double phase;
    uint32_t found, n_peaks;
};

struct FrameRot {     // constant rotation timing_sync applies to the samples of a frame (timing_sync.cpp:121-125)
    double2 rot_new;  // samples at stream index >= from
    double2 rot_old;  // samples before (the phase the previous frame left behind)
    uint64_t from;    // stream index of the frame's STS_END tag
};

struct CtaEvents {    // STS_END tags one detector CTA found, in stream order (offsets within the CTA's range)
    enum { CAP = 15 };
    uint16_t count;   // all found, even if more than CAP
    uint16_t off[CAP];
};

typedef b200rx_sync_result SyncSummary;

// Per-call scalars of a scan (b200rx_pass_scan) when its launches are replayed as a CUDA graph: the captured kernels have
// fixed arguments and read what changes from call to call here (device memory, refreshed by a copy node at the head of the
// graph from the pinned block the host fills in).
struct ScanParams {
    uint64_t n_samples;     // samples staged for this pass
    uint64_t x_limit;       // STS_END tags at or beyond it wait for the next capture (timing_sync.cpp:68)
    double2 rot_in;         // (cos, sin) of m_phase_acc in front of the capture
    int64_t origins[16];    // work() buffer origins (b200rx_set_receive_origins), ascending
    uint32_t n_origins;
    uint32_t pad;
};

struct SyncArgs {
    const void *iq;       // samples in format `fmt`
    int fmt;
    double scale;         // FMT_SC16: value of one LSB
    uint64_t n_samples;
    double2 rot_in;       // (cos, sin) of m_phase_acc before the stream
    uint32_t max_frames;
    uint8_t *tags;        // [n_samples] or null
    CtaEvents *cta_ev;    // scratch [sync_cta_count(n_samples)]
    uint8_t *cta_cnt;     // scratch [sync_cta_count(n_samples)]: events per detector CTA (saturating)
    uint64_t *ev_x;       // scratch [ev_cap]: STS_END positions in stream order
    uint32_t *ev_count;   // scratch [2]: events kept, events lost
    uint32_t ev_cap;
    const int64_t *origins; // device, ascending: work() buffer origins of a chunked stream (null: one buffer at -160)
    uint32_t n_origins;
    const ScanParams *sp;   // non-null (device): n_samples, rot_in and the origins come from there, grids cover grid_samples
    uint64_t grid_samples;
    int64_t origins_inline[16]; // the same list when it has at most 16 entries: travels in the kernel arguments (origins
    uint32_t n_inline;          // is then null) - a streaming call per 4096 samples has no copy to spare
    SyncRec *rec;         // scratch [ev_cap]
    uint64_t *lts1;       // [max_frames]
    uint32_t *avail;      // [max_frames]
    FrameRot *rot;        // [max_frames]
    double *phase;        // [max_frames] or null
    SyncSummary *summary; // device
};

cudaError_t launch_sync(const SyncArgs &a, cudaStream_t s);
uint32_t sync_cta_count(uint64_t n_samples);
cudaError_t upload_sync_tables();

// ---- per-handle tuning (b200rx_set_tuning); nothing in the launchers is process-wide ----
struct Tuning {
    int sm_count = 148;     // multiprocessors of the handle's device
    int acs_gen = 3;        // ACS kernel generation: 2 = viterbi_acs2.cuh (metric words), 3 = viterbi_acs3.cuh (soft pairs)
    int acs_lb = 0;         // lanes per frame (gen 2) / frame pair (gen 3) = 2^acs_lb; 0 = chosen from the batch size
    int acs_warps = 0;      // warps per ACS CTA; 0 = default
    int acs_rn = ACS_RN_DEFAULT; // gen 2: renormalisation variant
    int h2d_chunk = 1024;   // b200rx_submit_batch: frames per pipelined chunk, and the size the last chunks shrink to
    int h2d_chunk_min = 256;
    int pull_mode = -1;     // host-buffer ingest: 0 DMA copy, 1 GPU pull when the buffer is pinned, k >= 2 every k-th chunk by DMA, -1 by format
    int inflight = 1;       // batches the caller keeps in flight on this handle (pipeline depth): the ACS launcher sizes its
                            // warps for the GPU being shared, not for one batch alone
    int scan_graph = 1;     // b200rx_pass_scan: replay the scan launches as a CUDA graph (1) or issue them one by one (0)
    int fe_split = 1;       // front end: 1 = header kernel + one warp per OFDM symbol over all frames, 0 = one CTA per frame
};

// ---- launchers (each returns the cudaError_t of the launch) ----
struct FrontendArgs {
    const void *iq;          // samples in format `fmt`
    int fmt;
    double scale;            // FMT_SC16: value of one LSB
    uint64_t iq_samples;
    const uint64_t *lts1;
    const uint32_t *avail;
    uint32_t n_frames;
    FrameDesc *desc;
    uint32_t *bm;            // [n_frames][bm_stride] branch-metric words (ACS generation 2), or
    uint32_t bm_stride;      // words per frame (>= max_steps, multiple of 32)
    int emit_pairs;          // 1: the same buffer receives soft-symbol pairs, 2 bytes per trellis step (ACS generation 3)
    uint32_t max_steps;
    uint32_t max_len;
    int header_only;         // 1: stop after the SIGNAL symbol (descriptor only, no length check against `avail`);
                             // 2: header pass of the split front end (full status logic, exports H^-1 to hinv_out)
    double2 *hinv_out;       // [n_frames][64] inverse channel of each frame (shifted order), or null
    const FrameRot *rot;     // per frame, or null: samples are used as they are
    const uint32_t *n_live;  // device count of valid frames (slots beyond it become B200RX_ST_NO_FRAME), or null
    int sm_count;            // multiprocessors of the device (persistent grids); 0 = 148
    const ScanParams *sp;    // non-null (device): iq_samples comes from there (graph-replayed scan)
    const uint8_t *select;   // per frame, or null: frames with select[f] == 0 are skipped (B200RX_ST_NO_FRAME; b200rx_pass_decode)
    // taps (may be null)
    double2 *dbg_eq;
    uint32_t dbg_eq_vectors;
    uint8_t *dbg_depunct;
    uint32_t dbg_depunct_stride;
};

cudaError_t launch_frontend(const FrontendArgs &a, cudaStream_t s);
cudaError_t launch_frontend_data(const FrontendArgs &a, cudaStream_t s); // data symbols after launch_frontend(header_only = 2)

cudaError_t launch_bm_from_symbols(const uint8_t *symbols, uint64_t symbols_stride, const uint32_t *data_bits,
                                   uint32_t max_data_bits, uint32_t n_frames, FrameDesc *desc, uint32_t *bm,
                                   uint32_t bm_stride, uint32_t max_steps, cudaStream_t s);

cudaError_t launch_desc_from_bits(const uint32_t *data_bits, uint32_t n_frames, FrameDesc *desc, uint32_t max_steps, cudaStream_t s);

cudaError_t launch_viterbi_acs(const FrameDesc *desc, const uint32_t *bm, uint32_t bm_stride, uint32_t *dec,
                               uint32_t dec_stride_words, uint32_t n_frames, const Tuning &tn, cudaStream_t s);

// ACS generation 3: soft-symbol pairs (2 bytes per trellis step, frame f at soft + f * soft_stride bytes).  guard = the
// buffer is the caller's: 2-byte loads and nothing is read beyond a frame's own steps.
cudaError_t launch_viterbi_acs3(const FrameDesc *desc, const uint8_t *soft, uint64_t soft_stride, uint32_t *dec,
                                uint32_t dec_stride_words, uint32_t n_frames, bool guard, const Tuning &tn, cudaStream_t s);

struct TracebackArgs {
    FrameDesc *desc;
    const uint32_t *dec;     // survivor words, layout of viterbi_acs2.cuh
    uint32_t dec_stride;     // uint32 words per frame
    uint32_t n_frames;
    int raw_mode;            // 1: Viterbi-only entry point — write decoded bytes, no descramble/CRC
    uint8_t *payload;
    uint32_t payload_stride;
    uint16_t *payload_len;
    uint8_t *rate_out;
    uint8_t *status_out;
    unsigned long long *counters; // [0] frames ok, [1] frames failed, [2] payload bytes, [3] trellis steps, [4] traceback re-walks
    uint8_t *dbg_decoded;
    uint32_t dbg_decoded_stride;
    uint32_t *dbg_field;
};

cudaError_t launch_traceback(const TracebackArgs &a, const Tuning &tn, cudaStream_t s);
cudaError_t prepare_device_functions(); // per-device function attributes; called once per device by b200rx_create

cudaError_t launch_export_headers(const FrameDesc *desc, uint32_t n, uint16_t *len, uint8_t *rate, uint8_t *status,
                                  cudaStream_t s);

cudaError_t launch_pull(const void *src, void *dst, int fmt, uint64_t n_samples, const uint64_t *lts1, const uint32_t *avail,
                        uint32_t n_frames, int sm_count, cudaStream_t s);

cudaError_t upload_tables();

} // namespace b200rx
