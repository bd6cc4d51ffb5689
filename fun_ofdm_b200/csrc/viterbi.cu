// Viterbi decode of a batch of frames: add-compare-select over the whole trellis (survivor words to
// HBM), then traceback + descrambler + CRC-32 + payload delivery.
//
//   viterbi::conv_decode    (src/viterbi.cpp:31-37, 71-78, 166-181, 190-197, 208-459)  -> viterbi_acs_kernel
//   viterbi_chainback       (src/viterbi.cpp:108-146)                                  -> traceback_kernel
//   descrambler             (src/ppdu.cpp:255-264)                                     -> traceback_kernel
//   boost::crc_32_type use  (src/ppdu.cpp:267-279)                                     -> traceback_kernel
//   payload extraction      (src/ppdu.cpp:284-289)                                     -> traceback_kernel
#include "rx_internal.cuh"
#include "viterbi_core.cuh"

namespace b200rx {

namespace {

__constant__ uint32_t c_crc_tab[4][256]; // slice-by-4 tables of CRC-32/ISO-HDLC (reflected 0xEDB88320)
__constant__ uint8_t c_scramble[127];    // descrambler bit per byte index mod 127 (ppdu.cpp:257-263)

// ------------------------------------------------------------------------------------------------
// Branch metrics from depunctured soft symbols (Viterbi-only entry point).
// ------------------------------------------------------------------------------------------------
__global__ void bm_from_symbols_kernel(const uint8_t *symbols, uint64_t symbols_stride, const uint32_t *data_bits,
                                       uint32_t n_frames, FrameDesc *desc, uint32_t *bm, uint32_t bm_stride,
                                       uint32_t max_steps)
{
    const uint32_t frame = blockIdx.y;
    if (frame >= n_frames) return;
    const uint32_t nb = data_bits[frame];
    const uint32_t steps = nb + 6u;
    const bool ok = steps <= max_steps && (steps & 1u) == 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        FrameDesc d;
        d.n_steps = ok ? steps : 0u; d.data_bits = ok ? nb : 0u; d.field = 0; d.length = 0;
        d.rate = B200RX_RATE_INVALID; d.status = ok ? B200RX_ST_OK : B200RX_ST_TOO_LONG;
        desc[frame] = d;
    }
    if (!ok) return;
    const uint8_t *in = symbols + (size_t)frame * symbols_stride;
    uint32_t *out = bm + (size_t)frame * bm_stride;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < steps; t += gridDim.x * blockDim.x) {
        const uchar2 s = *reinterpret_cast<const uchar2 *>(in + 2 * (size_t)t);
        out[t] = bm_word(s.x, s.y);
    }
}

// ------------------------------------------------------------------------------------------------
// ACS: one warp per frame (see viterbi_core.cuh for the lane layout).  Per 32 steps: one coalesced
// 128 B load of metric words, one coalesced 256 B store of survivor words.
// ------------------------------------------------------------------------------------------------
constexpr int ACS_WARPS = 4;

__global__ void __launch_bounds__(ACS_WARPS * 32) viterbi_acs_kernel(const FrameDesc *desc, const uint32_t *bm,
                                                                      uint32_t bm_stride, uint2 *dec,
                                                                      uint32_t dec_stride, uint32_t n_frames)
{
    const int lane = threadIdx.x & 31;
    const uint32_t frame = blockIdx.x * ACS_WARPS + (threadIdx.x >> 5);
    if (frame >= n_frames) return;
    const uint32_t n_steps = desc[frame].n_steps;
    if (n_steps == 0) return;
    const uint32_t *w_in = bm + (size_t)frame * bm_stride;
    uint2 *d_out = dec + (size_t)frame * dec_stride;

    const AcsLane ln = acs_lane_init(lane);
    uint32_t R = acs_initial_metrics(lane);
    uint32_t cur = ((uint32_t)lane < n_steps) ? __ldg(w_in + lane) : 0u;
    for (uint32_t t0 = 0; t0 < n_steps; t0 += 32) {
        const uint32_t tn = t0 + 32 + lane;
        const uint32_t nxt = (tn < n_steps) ? __ldg(w_in + tn) : 0u;
        uint32_t my_e = 0, my_o = 0;
        const int cnt = min(32u, n_steps - t0);
        if (cnt == 32) {
#pragma unroll
            for (int i = 0; i < 32; i++) {
                uint32_t de, dod;
                const uint32_t Y = acs_step(R, __shfl_sync(VIT_FULL, cur, i), ln, lane, de, dod);
                if (lane == i) { my_e = de; my_o = dod; }
                R = acs_next(Y, ln);
            }
        } else {
            for (int i = 0; i < cnt; i++) {
                uint32_t de, dod;
                const uint32_t Y = acs_step(R, __shfl_sync(VIT_FULL, cur, i), ln, lane, de, dod);
                if (lane == i) { my_e = de; my_o = dod; }
                R = acs_next(Y, ln);
            }
        }
        if (lane < cnt) d_out[t0 + lane] = make_uint2(my_e, my_o);
        cur = nxt;
    }
}

// ------------------------------------------------------------------------------------------------
// Traceback: one CTA per frame.  The chainback (viterbi.cpp:131-142) is a 12 000-step dependent
// chain per frame; it is cut into tiles of TB_TILE decoded bits walked by different threads.
// A tile starts TB_PRE steps later than it has to, from an arbitrary state (0), and relies on
// survivor paths merging; this is then VERIFIED, not assumed: tile k is exact iff the state it had
// at its upper boundary equals the state tile k+1 (already exact, by induction from the last tile,
// which starts from the true end state 0) ended in.  Any tile failing the check is re-walked from
// the right state, so the output always equals the sequential chainback bit for bit.
// ------------------------------------------------------------------------------------------------
constexpr int TB_THREADS = 128;
constexpr int TB_TILE = 96;  // decoded bits per tile (multiple of 8)
constexpr int TB_PRE = 96;   // speculative pre-roll
constexpr int TB_MAX_BYTES = 4224; // >= (8*(4095+6)+6+215)/8

__device__ __forceinline__ uint32_t walk_tile(const uint2 *dec, uint32_t e, int n_from, int n_to, int out_below,
                                              uint8_t *bytes, int entry_at, uint32_t *entry_state)
{
    // processes decoded-bit indices n = n_from-1 ... n_to (descending); survivor word of bit n is step n+6
    for (int n = n_from - 1; n >= n_to; n--) {
        if (n == entry_at) *entry_state = e >> 2;
        const uint32_t k = decision_bit(dec, (uint32_t)n + 6u, e >> 2);
        e = (e >> 1) | (k << 7);
        if (n < out_below && (n & 7) == 0) bytes[n >> 3] = (uint8_t)e;
    }
    return e;
}

__global__ void __launch_bounds__(TB_THREADS) traceback_kernel(TracebackArgs a)
{
    __shared__ uint8_t s_bytes[TB_MAX_BYTES];
    __shared__ uint8_t s_entry[512], s_exit[512];
    __shared__ uint32_t s_crc[4][256];
    __shared__ int s_status;

    const uint32_t frame = blockIdx.x;
    const int tid = threadIdx.x;
    FrameDesc d = a.desc[frame];
    const int nbits = (int)d.data_bits;

    if (d.status != B200RX_ST_OK || nbits <= 0) {
        if (tid == 0) {
            if (a.status_out) a.status_out[frame] = d.status;
            if (a.payload_len) a.payload_len[frame] = (d.status == B200RX_ST_TRUNCATED || d.status == B200RX_ST_TOO_LONG) ? d.length : 0;
            if (a.rate_out) a.rate_out[frame] = d.rate;
            if (a.dbg_field) a.dbg_field[frame] = d.field;
            if (a.counters) atomicAdd(&a.counters[1], 1ull);
        }
        return;
    }

    if (!a.raw_mode)
        for (int i = tid; i < 1024; i += TB_THREADS) (&s_crc[0][0])[i] = (&c_crc_tab[0][0])[i];

    const uint2 *dec = a.dec + (size_t)frame * a.dec_stride;
    const int nbytes = (nbits + 7) >> 3;
    const int ntiles = (nbits + TB_TILE - 1) / TB_TILE;

    for (int k = tid; k < ntiles; k += TB_THREADS) {
        const int lo = k * TB_TILE;
        const int hi = min(lo + TB_TILE, nbits);
        const int from = min(hi + TB_PRE, nbits);
        uint32_t entry = 0; // state at the tile's upper boundary (true value 0 when hi == nbits)
        const uint32_t e = walk_tile(dec, 0u, from, lo, hi, s_bytes, hi - 1, &entry);
        s_entry[k] = (uint8_t)entry;
        s_exit[k] = (uint8_t)(e >> 2);
    }
    __syncthreads();

    if (tid == 0) {
        for (int k = ntiles - 2; k >= 0; k--) {
            if (s_entry[k] != s_exit[k + 1]) { // paths had not merged: redo from the verified state
                const int lo = k * TB_TILE, hi = lo + TB_TILE;
                uint32_t dummy;
                const uint32_t e = walk_tile(dec, (uint32_t)s_exit[k + 1] << 2, hi, lo, hi, s_bytes, -1, &dummy);
                s_exit[k] = (uint8_t)(e >> 2);
            }
        }
    }
    __syncthreads();

    if (a.dbg_decoded)
        for (int i = tid; i < nbytes && i < (int)a.dbg_decoded_stride; i += TB_THREADS)
            a.dbg_decoded[(size_t)frame * a.dbg_decoded_stride + i] = s_bytes[i];

    if (a.raw_mode) {
        for (int i = tid; i < nbytes && i < (int)a.payload_stride; i += TB_THREADS)
            a.payload[(size_t)frame * a.payload_stride + i] = s_bytes[i];
        if (tid == 0) {
            if (a.status_out) a.status_out[frame] = B200RX_ST_OK;
            if (a.counters) { atomicAdd(&a.counters[0], 1ull); atomicAdd(&a.counters[3], (unsigned long long)d.n_steps); }
        }
        return;
    }

    // ppdu.cpp:255-264: num_data_bytes = (nsym*dbps)/8 bytes descrambled; bit 0 of byte x flipped by
    // the LFSR output of step x (state 93, one step per byte, period 127)
    const int num_data_bytes = (int)(d.n_steps >> 3);
    for (int i = tid; i < num_data_bytes && i < nbytes; i += TB_THREADS) s_bytes[i] ^= c_scramble[i % 127];
    __syncthreads();

    // ppdu.cpp:267-279: CRC-32 over service(2) + payload, against the little-endian word behind it
    const int len = d.length;
    if (tid == 0) {
        uint32_t r = 0xFFFFFFFFu;
        const int n = 2 + len;
        int i = 0;
        for (; i + 4 <= n; i += 4) {
            r ^= (uint32_t)s_bytes[i] | ((uint32_t)s_bytes[i + 1] << 8) | ((uint32_t)s_bytes[i + 2] << 16) |
                 ((uint32_t)s_bytes[i + 3] << 24);
            r = s_crc[3][r & 0xFF] ^ s_crc[2][(r >> 8) & 0xFF] ^ s_crc[1][(r >> 16) & 0xFF] ^ s_crc[0][r >> 24];
        }
        for (; i < n; i++) r = s_crc[0][(r ^ s_bytes[i]) & 0xFF] ^ (r >> 8);
        r ^= 0xFFFFFFFFu;
        const uint32_t given = (uint32_t)s_bytes[n] | ((uint32_t)s_bytes[n + 1] << 8) |
                               ((uint32_t)s_bytes[n + 2] << 16) | ((uint32_t)s_bytes[n + 3] << 24);
        s_status = (r == given) ? B200RX_ST_OK : B200RX_ST_CRC_FAIL;
    }
    __syncthreads();
    const int status = s_status;

    // payload = descrambled[2 .. 2+len) (ppdu.cpp:284-285); written for CRC failures too (status says so)
    if (a.payload)
        for (int i = tid; i < len && i < (int)a.payload_stride; i += TB_THREADS)
            a.payload[(size_t)frame * a.payload_stride + i] = s_bytes[2 + i];
    if (tid == 0) {
        if (a.status_out) a.status_out[frame] = (uint8_t)status;
        if (a.payload_len) a.payload_len[frame] = (uint16_t)len;
        if (a.rate_out) a.rate_out[frame] = d.rate;
        if (a.dbg_field) a.dbg_field[frame] = d.field;
        a.desc[frame].status = (uint8_t)status;
        if (a.counters) {
            atomicAdd(&a.counters[status == B200RX_ST_OK ? 0 : 1], 1ull);
            if (status == B200RX_ST_OK) atomicAdd(&a.counters[2], (unsigned long long)len);
            atomicAdd(&a.counters[3], (unsigned long long)d.n_steps);
        }
    }
}

} // namespace

cudaError_t upload_viterbi_tables(const uint32_t *crc, const uint8_t *scr)
{
    cudaError_t e = cudaMemcpyToSymbol(c_crc_tab, crc, sizeof(uint32_t) * 1024);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_scramble, scr, 127);
}

cudaError_t launch_bm_from_symbols(const uint8_t *symbols, uint64_t symbols_stride, const uint32_t *data_bits,
                                   uint32_t max_data_bits, uint32_t n_frames, FrameDesc *desc, uint32_t *bm,
                                   uint32_t bm_stride, uint32_t max_steps, cudaStream_t s)
{
    if (n_frames == 0) return cudaSuccess;
    const uint32_t steps = max_data_bits + 6u;
    dim3 grid((steps + 1023u) / 1024u, n_frames);
    if (grid.x == 0) grid.x = 1;
    bm_from_symbols_kernel<<<grid, 256, 0, s>>>(symbols, symbols_stride, data_bits, n_frames, desc, bm, bm_stride,
                                                 max_steps);
    return cudaGetLastError();
}

cudaError_t launch_viterbi_acs(const FrameDesc *desc, const uint32_t *bm, uint32_t bm_stride, uint2 *dec,
                               uint32_t dec_stride, uint32_t n_frames, cudaStream_t s)
{
    if (n_frames == 0) return cudaSuccess;
    viterbi_acs_kernel<<<(n_frames + ACS_WARPS - 1) / ACS_WARPS, ACS_WARPS * 32, 0, s>>>(desc, bm, bm_stride, dec,
                                                                                           dec_stride, n_frames);
    return cudaGetLastError();
}

cudaError_t launch_traceback(const TracebackArgs &a, cudaStream_t s)
{
    if (a.n_frames == 0) return cudaSuccess;
    traceback_kernel<<<a.n_frames, TB_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

} // namespace b200rx
