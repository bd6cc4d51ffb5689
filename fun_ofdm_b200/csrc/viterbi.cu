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
#include "viterbi_acs2.cuh"
#include "viterbi_acs3.cuh"

namespace b200rx {

namespace {

__constant__ uint32_t c_crc_tab[4][256]; // slice-by-4 tables of CRC-32/ISO-HDLC (reflected 0xEDB88320)
__constant__ uint8_t c_scramble[127];    // descrambler bit per byte index mod 127 (ppdu.cpp:257-263)
__constant__ uint32_t c_xpow[5][132];    // x^(8 * L * 2^lvl) mod P, reflected: CRC combine factors, L <= 129

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
// ACS, second generation (viterbi_acs2.cuh): 8 lanes per frame, 4 frames per warp, metrics in place.
// Per 24 steps and frame: one 96 B read of metric words (staged through shared memory, double
// buffered, prefetched one block ahead) and three 64 B survivor stores.
// ------------------------------------------------------------------------------------------------
template <int LB, int RN, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) viterbi_acs2_kernel(const FrameDesc *desc, const uint32_t *bm, uint32_t bm_stride,
                                                          uint32_t *dec, uint32_t dec_stride_words, uint32_t n_frames,
                                                          uint32_t neg1)
{
    using A = Acs2<LB>;
    constexpr int T = A::T, NR = A::NR, FPW = A::FPW;
    constexpr int LOADERS = 6;             // 16-byte pieces per 24 metric words
    constexpr int PER_LANE = (LOADERS + T - 1) / T; // pieces each lane of a group moves (T = 4: 2, else 1)
    __shared__ __align__(16) uint32_t s_w[WARPS][2][FPW][ACS2_BLK];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int group = lane >> LB, glane = lane & (T - 1);
    const uint32_t frame0 = (blockIdx.x * WARPS + warp) * FPW;
    if (frame0 >= n_frames) return;
    const uint32_t frame = min(frame0 + group, n_frames - 1); // surplus groups shadow the last frame (no stores)
    const bool live = frame0 + group < n_frames;
    const uint32_t n_steps = live ? desc[frame].n_steps : 0u;
    const uint32_t my_blocks = (n_steps + ACS2_BLK - 1) / ACS2_BLK;
    const uint32_t n_blocks = __reduce_max_sync(0xFFFFFFFFu, my_blocks);
    if (n_blocks == 0) return;

    const uint32_t *w_in = bm + (size_t)frame * bm_stride;
    uint32_t *d_out = dec + (size_t)frame * dec_stride_words;

    typename A::Lane L;
    A::lane_init(L, glane, neg1);
    uint32_t R[NR];
    A::init_metrics(R, glane);

    // the lanes of each group move the group's 24 metric words (6 x 16 B) per block
    uint4 pre[PER_LANE];
#pragma unroll
    for (int j = 0; j < PER_LANE; j++) {
        const int piece = glane * PER_LANE + j;
        pre[j] = make_uint4(0, 0, 0, 0);
        if (piece < LOADERS) pre[j] = __ldg(reinterpret_cast<const uint4 *>(w_in) + piece);
    }
    for (uint32_t b = 0; b < n_blocks; b++) {
        uint32_t *sw = &s_w[warp][b & 1][group][0];
#pragma unroll
        for (int j = 0; j < PER_LANE; j++) {
            const int piece = glane * PER_LANE + j;
            if (piece < LOADERS) reinterpret_cast<uint4 *>(sw)[piece] = pre[j];
        }
        __syncwarp();
        if (b + 1 < my_blocks) {
#pragma unroll
            for (int j = 0; j < PER_LANE; j++) {
                const int piece = glane * PER_LANE + j;
                if (piece < LOADERS)
                    pre[j] = __ldg(reinterpret_cast<const uint4 *>(w_in + (size_t)(b + 1) * ACS2_BLK) + piece);
            }
        }
        uint32_t *d_blk = d_out + (size_t)b * 3 * ACS2_WORDS_PER_8 + (NR >= 2 ? glane * (NR / 2) : (glane >> 1));
        const bool store = b < my_blocks;
#pragma unroll
        for (int o = 0; o < 3; o++) {
            const uint4 wa = reinterpret_cast<const uint4 *>(sw)[2 * o];
            const uint4 wb = reinterpret_cast<const uint4 *>(sw)[2 * o + 1];
            uint32_t acc[A::NA];
#pragma unroll
            for (int j = 0; j < A::NA; j++) acc[j] = 0;
            // 8 steps; phase = (8 * o + i) % 6
            if (o == 0) {
                A::template one<0, RN>(R, acc, wa.x, L, glane, group); A::template one<1, RN>(R, acc, wa.y, L, glane, group);
                A::template one<2, RN>(R, acc, wa.z, L, glane, group); A::template one<3, RN>(R, acc, wa.w, L, glane, group);
                A::template one<4, RN>(R, acc, wb.x, L, glane, group); A::template one<5, RN>(R, acc, wb.y, L, glane, group);
                A::template one<0, RN>(R, acc, wb.z, L, glane, group); A::template one<1, RN>(R, acc, wb.w, L, glane, group);
            } else if (o == 1) {
                A::template one<2, RN>(R, acc, wa.x, L, glane, group); A::template one<3, RN>(R, acc, wa.y, L, glane, group);
                A::template one<4, RN>(R, acc, wa.z, L, glane, group); A::template one<5, RN>(R, acc, wa.w, L, glane, group);
                A::template one<0, RN>(R, acc, wb.x, L, glane, group); A::template one<1, RN>(R, acc, wb.y, L, glane, group);
                A::template one<2, RN>(R, acc, wb.z, L, glane, group); A::template one<3, RN>(R, acc, wb.w, L, glane, group);
            } else {
                A::template one<4, RN>(R, acc, wa.x, L, glane, group); A::template one<5, RN>(R, acc, wa.y, L, glane, group);
                A::template one<0, RN>(R, acc, wa.z, L, glane, group); A::template one<1, RN>(R, acc, wa.w, L, glane, group);
                A::template one<2, RN>(R, acc, wb.x, L, glane, group); A::template one<3, RN>(R, acc, wb.y, L, glane, group);
                A::template one<4, RN>(R, acc, wb.z, L, glane, group); A::template one<5, RN>(R, acc, wb.w, L, glane, group);
            }
#pragma unroll
            for (int j = 0; j < A::NA; j++) acc[j] ^= L.flip[o];
            if constexpr (NR == 1) { // lanes 2k and 2k + 1 share survivor word k: low and high 16 bits
                const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, acc[0], 1);
                acc[0] = (acc[0] & 0xFFFFu) | (other << 16);
            }
            if (store) {
                uint32_t *dst = d_blk + o * ACS2_WORDS_PER_8;
                if constexpr (NR / 2 == 4) *reinterpret_cast<uint4 *>(dst) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
                else if constexpr (NR / 2 == 2) *reinterpret_cast<uint2 *>(dst) = make_uint2(acc[0], acc[1]);
                else if constexpr (NR == 2) dst[0] = acc[0];
                else if ((glane & 1) == 0) dst[0] = acc[0];
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// ACS, third generation (viterbi_acs3.cuh): two frames per register, 2^LB lanes per frame pair, 32 >> LB pairs per warp.
// Input: soft-symbol pairs, 2 bytes per trellis step (frame f at soft + f * soft_stride bytes).  Per 24 steps and pair:
// 96 B of pairs are read (prefetched one block ahead), turned into 48 metric words in shared memory (double buffered),
// and 2 x 3 survivor rows of 64 B are stored.
// GUARD: the buffer is the caller's (Viterbi-only entry point): 2-byte loads, nothing is read beyond a frame's n_steps.
// ------------------------------------------------------------------------------------------------
// Registers: 128 per thread at LB = 2 (16 warps per SM), 96 at LB = 3.  Capping LB = 2 at 96 as well compiles without spills
// but schedules worse: 0.892 instead of 0.823 ms per pipelined step (bench.py, 12 batches in flight).
constexpr int acs3_min_warps_per_sm(int lb) { return lb <= 2 ? 16 : 20; }
template <int LB, int WARPS, bool GUARD, bool LAZY>
__global__ void __launch_bounds__(32 * WARPS, acs3_min_warps_per_sm(LB) / WARPS) viterbi_acs3_kernel(const FrameDesc *desc, const uint8_t *soft, uint64_t soft_stride,
                                                                  uint32_t *dec, uint32_t dec_stride_words, uint32_t n_frames,
                                                                  uint32_t neg1)
{
    using A = Acs3<LB, LAZY>;
    constexpr int T = A::T, NR = A::NR, PPW = A::PPW, NA = A::NA;
    constexpr int SPL = 2 * ACS2_BLK / T;   // steps each lane of a pair stages per block (both frames: 48 steps)
    static_assert(SPL % 2 == 0, "a lane stages whole 32-bit words (two steps)");
    constexpr int WPL = SPL / 2;            // 32-bit words of pairs per lane per block
    // 2 x 24 words per pair + 4 words of padding: the pairs of a warp read the same word index at the same time, and a
    // stride of 48 words puts every other pair on the same banks (ncu: 4-way conflicts on the metric loads)
    constexpr int PAIR_W = 2 * ACS2_BLK + 4;
    __shared__ __align__(16) uint32_t s_w[WARPS][2][PPW][PAIR_W];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int group = lane >> LB, glane = lane & (T - 1);
    const uint32_t pair0 = (blockIdx.x * WARPS + warp) * PPW;
    if (2u * pair0 >= n_frames) return;
    const uint32_t fA = 2u * (pair0 + group), fB = fA + 1u;
    const uint32_t nA = fA < n_frames ? desc[fA].n_steps : 0u;
    const uint32_t nB = fB < n_frames ? desc[fB].n_steps : 0u;
    const uint32_t blocksA = (nA + ACS2_BLK - 1) / ACS2_BLK, blocksB = (nB + ACS2_BLK - 1) / ACS2_BLK;
    const uint32_t my_blocks = max(blocksA, blocksB);
    const uint32_t n_blocks = __reduce_max_sync(0xFFFFFFFFu, my_blocks);
    if (n_blocks == 0) return;

    // staging role of this lane: the first half of the pair's lanes moves frame A, the second half frame B
    const int st_frame = glane / (T / 2), st_part = glane % (T / 2);
    const uint32_t st_f = st_frame ? fB : fA;
    const uint32_t st_n = st_frame ? nB : nA;           // steps of the staged frame
    const uint32_t st_blocks = st_frame ? blocksB : blocksA;
    const uint8_t *st_src = soft + (size_t)min(st_f, n_frames - 1) * soft_stride + (size_t)st_part * SPL * 2;

    typename A::Lane L;
    A::lane_init(L, glane, neg1);
    uint32_t R[NR];
    A::init_metrics(R, glane);

    uint32_t pre[WPL];
    auto fetch = [&](uint32_t b) {
        const uint8_t *src = st_src + (size_t)b * (ACS2_BLK * 2);
        if constexpr (GUARD) {
            const uint32_t s0 = b * ACS2_BLK + (uint32_t)st_part * SPL;
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                uint32_t lo = 0, hi = 0;
                if (s0 + 2 * j < st_n) lo = __ldg(reinterpret_cast<const unsigned short *>(src) + 2 * j);
                if (s0 + 2 * j + 1 < st_n) hi = __ldg(reinterpret_cast<const unsigned short *>(src) + 2 * j + 1);
                pre[j] = lo | (hi << 16);
            }
        } else if constexpr (WPL % 2 == 0) {
#pragma unroll
            for (int j = 0; j < WPL / 2; j++) {
                const uint2 v = __ldg(reinterpret_cast<const uint2 *>(src) + j);
                pre[2 * j] = v.x; pre[2 * j + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < WPL; j++) pre[j] = __ldg(reinterpret_cast<const uint32_t *>(src) + j);
        }
    };
#pragma unroll
    for (int j = 0; j < WPL; j++) pre[j] = 0;
    if (st_blocks > 0) fetch(0);

    uint32_t acc[NA];
#pragma unroll
    for (int j = 0; j < NA; j++) acc[j] = 0;
    uint32_t *dA = dec + (size_t)min(fA, n_frames - 1) * dec_stride_words + glane * (NR / 4);
    uint32_t *dB = dec + (size_t)min(fB, n_frames - 1) * dec_stride_words + glane * (NR / 4);

    for (uint32_t b = 0; b < n_blocks; b++) {
        uint32_t *sw = &s_w[warp][b & 1][group][st_frame * ACS2_BLK + st_part * SPL];
#pragma unroll
        for (int j = 0; j < WPL; j++) acs3_bm_words2(pre[j], sw[2 * j], sw[2 * j + 1]);
        __syncwarp();
        if (b + 1 < st_blocks) fetch(b + 1);
        A::rebase(R, L, 20000u); // + 24 steps x 255 stays far below 0x7F2D, where the threshold constant would wrap
        const uint32_t *swA = &s_w[warp][b & 1][group][0], *swB = &s_w[warp][b & 1][group][ACS2_BLK];
        const bool storeA = b < blocksA, storeB = b < blocksB;
        uint32_t *rowA = dA + (size_t)b * 3 * ACS2_WORDS_PER_8, *rowB = dB + (size_t)b * 3 * ACS2_WORDS_PER_8;
        // acc[j] = bytes [B of register 2j, A of 2j, B of 2j+1, A of 2j+1]; survivor word w of a frame holds positions
        // 4w .. 4w+3 of this lane as bytes [4w+1, 4w, 4w+3, 4w+2]
        auto store = [&](int o) {
            uint32_t wa[NR / 4], wb[NR / 4];
#pragma unroll
            for (int w = 0; w < NR / 4; w++) {
                wa[w] = __byte_perm(acc[2 * w], acc[2 * w + 1], 0x5713u) ^ L.flip[o];
                wb[w] = __byte_perm(acc[2 * w], acc[2 * w + 1], 0x4602u) ^ L.flip[o];
            }
            uint32_t *ra = rowA + o * ACS2_WORDS_PER_8, *rb = rowB + o * ACS2_WORDS_PER_8;
            if constexpr (NR / 4 == 8) {
                if (storeA) { reinterpret_cast<uint4 *>(ra)[0] = make_uint4(wa[0], wa[1], wa[2], wa[3]); reinterpret_cast<uint4 *>(ra)[1] = make_uint4(wa[4], wa[5], wa[6], wa[7]); }
                if (storeB) { reinterpret_cast<uint4 *>(rb)[0] = make_uint4(wb[0], wb[1], wb[2], wb[3]); reinterpret_cast<uint4 *>(rb)[1] = make_uint4(wb[4], wb[5], wb[6], wb[7]); }
            } else if constexpr (NR / 4 == 4) {
                if (storeA) *reinterpret_cast<uint4 *>(ra) = make_uint4(wa[0], wa[1], wa[2], wa[3]);
                if (storeB) *reinterpret_cast<uint4 *>(rb) = make_uint4(wb[0], wb[1], wb[2], wb[3]);
            } else {
                if (storeA) *reinterpret_cast<uint2 *>(ra) = make_uint2(wa[0], wa[1]);
                if (storeB) *reinterpret_cast<uint2 *>(rb) = make_uint2(wb[0], wb[1]);
            }
#pragma unroll
            for (int j = 0; j < NA; j++) acc[j] = 0;
        };
        // The unrolled body is one phase cycle (6 steps), not the 24-step block: with 64 >> LB state registers per lane a
        // 24-step body is 30-50 KB of code and ncu showed the warps waiting for instruction fetches (stall_no_instruction
        // 0.87 per issue at LB = 2).  Survivor rows are 8 steps, so a row ends inside cycles 1, 2 and at the end of cycle 3.
#pragma unroll 1
        for (int c = 0; c < 4; c++) {
            const uint2 a01 = reinterpret_cast<const uint2 *>(swA)[3 * c], a23 = reinterpret_cast<const uint2 *>(swA)[3 * c + 1],
                        a45 = reinterpret_cast<const uint2 *>(swA)[3 * c + 2];
            const uint2 b01 = reinterpret_cast<const uint2 *>(swB)[3 * c], b23 = reinterpret_cast<const uint2 *>(swB)[3 * c + 1],
                        b45 = reinterpret_cast<const uint2 *>(swB)[3 * c + 2];
            A::template one<0>(R, acc, a01.x, b01.x, L, lane);
            A::template one<1>(R, acc, a01.y, b01.y, L, lane);
            if (c == 1) store(0);
            A::template one<2>(R, acc, a23.x, b23.x, L, lane);
            A::template one<3>(R, acc, a23.y, b23.y, L, lane);
            if (c == 2) store(1);
            A::template one<4>(R, acc, a45.x, b45.x, L, lane);
            A::template one<5>(R, acc, a45.y, b45.y, L, lane);
            if (c == 3) store(2);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Traceback.  The chainback (viterbi.cpp:131-142) is a dependent chain of one step per decoded bit
// (12 090 for a 1500-byte frame); it is cut into tiles of TB_TILE bits walked by different threads.
// A tile starts TB_PRE steps later than it has to, from an arbitrary state (0), and relies on survivor
// paths merging; this is then VERIFIED, not assumed: tile k is exact iff the state it had at its upper
// boundary equals the state tile k+1 (already exact, by induction from the last tile, which starts from
// the true end state 0) ended in.  A tile failing the check is re-walked from the right state, so the
// output always equals the sequential chainback bit for bit (re-walks are counted: stats.traceback_rewalks).
//
// The walk runs in POSITION space: with q = rotr6^(tau % 6)(state at time tau) the survivor bit of the step
// into tau sits at position q of row (tau-1)>>3, and going one step back only replaces bit
// (6 - tau % 6) % 6 of q by that bit (the in-place butterfly of viterbi_acs2.cuh seen backwards).  Unrolled
// over 24 steps every shift, row offset and byte store is a compile-time constant: ~12 instructions per step.
//
// Memory: 32 tiles of a warp sit 768 B apart, so reading survivor words straight from global memory costs
// 32 L1 wavefronts per warp load and the kernel is L1-pipeline bound (ncu r01: issue 16 %, long_scoreboard
// 30).  Instead the CTA works in rounds of 24 steps: the four 64-byte survivor rows every tile needs for its
// next 24 steps are copied to shared memory with coalesced 16-byte cp.async (256 B contiguous per tile),
// then each thread walks its 24 steps out of shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int TB_THREADS = 128;
constexpr int TB_TILE = 96;         // decoded bits per tile (multiple of 24)
constexpr int TB_PRE = 72;          // speculative pre-roll (multiple of 24)
constexpr int TB_ROUNDS = (TB_TILE + TB_PRE) / 24;
constexpr int TB_CTAS_PER_SM = 5;   // what 44 KB of shared memory per CTA allow: one full wave, no straggler CTAs
constexpr int TB_ROW_W = 16;        // words per staged row (64 B, all used)
constexpr int TB_TILE_W = 4 * TB_ROW_W + 4; // words per tile slot; 68 = 4 mod 32 keeps a quarter warp's 16-byte stores on distinct banks
constexpr int TB_MAX_BYTES = 4224;  // >= (8*(4095+6)+6+215)/8
constexpr int TB_MAX_TILES = (TB_MAX_BYTES * 8 + TB_TILE - 1) / TB_TILE; // 352

template <bool GLOBAL>
__device__ __forceinline__ uint32_t tb_lookup(const uint32_t *row, uint32_t q, int t_and_7)
{
    // word lane*2 + (reg>>1) = q >> 2; byte (reg&1)*2 + (1 - low) = (q ^ 1) & 3; bit 7 - (t & 7)
    const uint32_t w = GLOBAL ? __ldg(row + (q >> 2)) : row[q >> 2];
    return (w >> ((((q ^ 1u) & 3u) << 3) + (7u - (uint32_t)t_and_7))) & 1u;
}

// One backward step from time tau (dynamic phase), survivor words read from global memory.
__device__ __forceinline__ void tb_step_generic(const uint32_t *dec, int tau, uint32_t &q, uint32_t &e)
{
    const int t = tau - 1;
    const uint32_t k = tb_lookup<true>(dec + (size_t)(t >> 3) * ACS2_WORDS_PER_8, q, t & 7);
    const int b = (6 - tau % 6) % 6;
    q = (q & ~(1u << b)) | (k << b);
    e = (e >> 1) | (k << 7);
}

// 24 backward steps for decoded bits n = 24m+23 ... 24m (times tau = 24m+30 ... 24m+7), all constants static.
// `row0` = survivor row 3m (the row of step 24m), rows ROW_W words apart.  Bytes 3m+2, 3m+1, 3m are complete at
// n = 24m+16, +8, +0.
template <bool GLOBAL, int ROW_W, bool STORE>
__device__ __forceinline__ void tb_block24(const uint32_t *row0, uint32_t &q, uint32_t &e, uint8_t *bytes3m)
{
#pragma unroll
    for (int i = 23; i >= 0; i--) {
        const int t = i + 6;                 // step index relative to 24m
        const int tau = t + 1;
        const uint32_t k = tb_lookup<GLOBAL>(row0 + (t >> 3) * ROW_W, q, t & 7);
        const int b = (6 - tau % 6) % 6;     // 24m is a multiple of 6
        q = (q & ~(1u << b)) | (k << b);
        e = (e >> 1) | (k << 7);
        if (STORE && (i & 7) == 0) bytes3m[i >> 3] = (uint8_t)e;
    }
}

__device__ __forceinline__ uint32_t tb_state_to_pos(uint32_t state, int tau)
{
    const int r = tau % 6;
    return ((state >> r) | (state << (6 - r))) & 63u;
}
__device__ __forceinline__ uint32_t tb_pos_to_state(uint32_t q, int tau)
{
    const int r = tau % 6;
    return ((q << r) | (q >> (6 - r))) & 63u;
}

// Serial repair walk (rare): decoded bits n = n_from-1 ... n_to from a known state, straight from global memory.
__device__ __forceinline__ uint32_t tb_walk_global(const uint32_t *dec, uint32_t state, int n_from, int n_to,
                                                   uint8_t *bytes)
{
    uint32_t q = tb_state_to_pos(state, n_from + 6), e = state << 2;
    int n = n_from - 1;
    for (; (n + 1) % 24 != 0 && n >= n_to; n--) {
        tb_step_generic(dec, n + 7, q, e);
        if ((n & 7) == 0) bytes[n >> 3] = (uint8_t)e;
    }
    for (; n >= n_to; n -= 24) {
        const int m3 = (n - 23) >> 3;
        tb_block24<true, ACS2_WORDS_PER_8, true>(dec + (size_t)m3 * ACS2_WORDS_PER_8, q, e, bytes + m3);
    }
    return tb_pos_to_state(q, n_to + 6);
}

// a(x) * b(x) mod P(x) in the reflected CRC-32 representation (bit 31 = x^0)
__device__ __forceinline__ uint32_t crc_multmodp(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
#pragma unroll 4
    for (int i = 31; i >= 0; i--) {
        p ^= ((a >> i) & 1u) ? b : 0u;
        b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
    }
    return p;
}

__global__ void __launch_bounds__(TB_THREADS) traceback_kernel(TracebackArgs a)
{
    extern __shared__ __align__(16) uint32_t s_rows[]; // TB_THREADS * TB_TILE_W words (34 KB; 44 KB with the static arrays)
    __shared__ uint8_t s_bytes[TB_MAX_BYTES];
    __shared__ uint8_t s_entry[TB_MAX_TILES + 1], s_exit[TB_MAX_TILES + 1];
    __shared__ uint32_t s_crc[4][256];
    __shared__ uint32_t s_flag;
    __shared__ short s_mtop[TB_THREADS];

    const int tid = threadIdx.x;
    const uint32_t s_rows_u32 = (uint32_t)__cvta_generic_to_shared(s_rows);
    if (!a.raw_mode)
        for (int i = tid; i < 1024; i += TB_THREADS) (&s_crc[0][0])[i] = (&c_crc_tab[0][0])[i];

    for (uint32_t frame = blockIdx.x; frame < a.n_frames; frame += gridDim.x) {
        __syncthreads(); // shared buffers of the previous frame are free
        const FrameDesc d = a.desc[frame];
        const int nbits = (int)d.data_bits;

        if (d.status != B200RX_ST_OK || nbits <= 0) {
            if (tid == 0) {
                if (a.status_out) a.status_out[frame] = d.status;
                if (a.payload_len) a.payload_len[frame] = (d.status == B200RX_ST_TRUNCATED || d.status == B200RX_ST_TOO_LONG) ? d.length : 0;
                if (a.rate_out) a.rate_out[frame] = d.rate;
                if (a.dbg_field) a.dbg_field[frame] = d.field;
                if (a.counters && d.status != B200RX_ST_NO_FRAME) atomicAdd(&a.counters[1], 1ull);
            }
            continue;
        }

        const uint32_t *dec = a.dec + (size_t)frame * a.dec_stride;
        const int nbytes = (nbits + 7) >> 3;
        const int ntiles = (nbits + TB_TILE - 1) / TB_TILE;

        for (int g = 0; g * TB_THREADS < ntiles; g++) {
            const int k = g * TB_THREADS + tid;
            const bool active = k < ntiles;
            const int lo = k * TB_TILE;
            const int hi = min(lo + TB_TILE, nbits);
            const int from = min(hi + TB_PRE, nbits);
            uint32_t q = 0, e = 0, entry = 0; // start state 0 (true when from == nbits, a guess otherwise)
            int n = from - 1;
            if (active) {
                q = tb_state_to_pos(0u, from + 6);
                // ragged top (only tiles within TB_PRE of the end of the frame): single steps from global memory
                for (; (n + 1) % 24 != 0 && n >= lo; n--) {
                    if (n == hi - 1) entry = tb_pos_to_state(q, n + 7);
                    tb_step_generic(dec, n + 7, q, e);
                    if (n < hi && (n & 7) == 0) s_bytes[n >> 3] = (uint8_t)e;
                }
            }
            // Each warp stages and walks its own 32 tiles; warps drift apart, so one warp's wait for its rows
            // is covered by the other warps' walking (no CTA-wide barrier inside the rounds).
            // Lane l moves piece (l & 15) of tiles (l >> 4), (l >> 4) + 2, ... : 16 cp.async of 16 B per round.
            const int wtile0 = tid & ~31, lane = tid & 31;
            s_mtop[tid] = active ? (short)((n + 1) / 24 - 1) : (short)-1; // first block of my tile (n + 1 is a multiple of 24 now)
            __syncwarp();
            const int part = lane & 15;
            const uint32_t dst0 = s_rows_u32 + (uint32_t)((wtile0 + (lane >> 4)) * TB_TILE_W + (part >> 2) * TB_ROW_W + (part & 3) * 4) * 4u;
            const uint32_t *src0 = dec + part * 4;
            for (int round = 0; round < TB_ROUNDS; round++) {
                __syncwarp(); // the warp is done with the rows of the previous round
#pragma unroll 4
                for (int j = 0; j < 16; j++) {
                    const int tile = wtile0 + (lane >> 4) + 2 * j;
                    const int m = (int)s_mtop[tile] - round;          // block this tile walks in this round
                    if (m >= 4 * (g * TB_THREADS + tile))              // still inside [lo, from): lo / 24 = 4 k
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)(2 * j * TB_TILE_W * 4)),
                                     "l"(src0 + (size_t)(3 * m) * ACS2_WORDS_PER_8));
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();
                if (active && n >= lo) {
                    if (n == hi - 1) entry = tb_pos_to_state(q, n + 7);
                    const int m3 = (n - 23) >> 3;
                    const uint32_t *row0 = s_rows + tid * TB_TILE_W;
                    if (n < hi) tb_block24<false, TB_ROW_W, true>(row0, q, e, s_bytes + m3);
                    else tb_block24<false, TB_ROW_W, false>(row0, q, e, s_bytes + m3);
                    n -= 24;
                }
            }
            if (active) {
                s_entry[k] = (uint8_t)entry;
                s_exit[k] = (uint8_t)tb_pos_to_state(q, lo + 6);
            }
        }
        __syncthreads();

        // Verify the chain in parallel.  A tile whose entry state differs from the exit state of the tile above
        // had not merged within its pre-roll: it is re-walked from that exit state.  All such tiles are repaired
        // at once and the check repeated (a repaired tile may end in a different state than before and invalidate
        // the tile below it); exactness spreads downwards from the top tile at least one tile per pass, in
        // practice one or two passes suffice even for frames that are pure noise.
        unsigned rewalks = 0;
        for (int pass = 0; pass < ntiles; pass++) {
            uint32_t fix_from[(TB_MAX_TILES + TB_THREADS - 1) / TB_THREADS];
            int bad = 0, j = 0;
            for (int k = tid; k < ntiles - 1; k += TB_THREADS, j++) {
                const uint32_t above = s_exit[k + 1];
                fix_from[j] = (s_entry[k] != above) ? above : 0xFFu;
                bad |= (fix_from[j] != 0xFFu);
            }
            if (!__syncthreads_or(bad)) break; // (also orders the reads above before the writes below)
            j = 0;
            for (int k = tid; k < ntiles - 1; k += TB_THREADS, j++) {
                if (fix_from[j] == 0xFFu) continue;
                const int lo = k * TB_TILE, hi = lo + TB_TILE;
                s_exit[k] = (uint8_t)tb_walk_global(dec, fix_from[j], hi, lo, s_bytes);
                s_entry[k] = (uint8_t)fix_from[j];
                rewalks++;
            }
            __syncthreads();
        }
        if (rewalks && a.counters) atomicAdd(&a.counters[4], (unsigned long long)rewalks);

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
            continue;
        }

        // ppdu.cpp:255-264: num_data_bytes = (nsym*dbps)/8 bytes descrambled; bit 0 of byte x flipped by
        // the LFSR output of step x (state 93, one step per byte, period 127)
        const int num_data_bytes = (int)(d.n_steps >> 3);
        for (int i = tid; i < num_data_bytes && i < nbytes; i += TB_THREADS) s_bytes[i] ^= c_scramble[i % 127];
        __syncthreads();

        // ppdu.cpp:267-279: CRC-32 over service(2) + payload, against the little-endian word behind it.
        // Warp-parallel: 32 right-aligned segments of L bytes (the first may be shorter or empty), each lane
        // takes the standard CRC of its segment, then a 5-level tree of crc(A||B) = crc(A)*x^(8|B|) ^ crc(B).
        const int len = d.length;
        const int n = 2 + len;
        if (tid < 32) {
            const int L = (n + 31) >> 5;
            const int beg = max(0, n - (32 - tid) * L), end = max(0, n - (31 - tid) * L);
            uint32_t r = 0xFFFFFFFFu;
            int i = beg;
            for (; i + 4 <= end; i += 4) {
                r ^= (uint32_t)s_bytes[i] | ((uint32_t)s_bytes[i + 1] << 8) | ((uint32_t)s_bytes[i + 2] << 16) |
                     ((uint32_t)s_bytes[i + 3] << 24);
                r = s_crc[3][r & 0xFF] ^ s_crc[2][(r >> 8) & 0xFF] ^ s_crc[1][(r >> 16) & 0xFF] ^ s_crc[0][r >> 24];
            }
            for (; i < end; i++) r = s_crc[0][(r ^ s_bytes[i]) & 0xFF] ^ (r >> 8);
            uint32_t crc = (end > beg) ? (r ^ 0xFFFFFFFFu) : 0u; // CRC of the empty string is 0
#pragma unroll
            for (int lvl = 0; lvl < 5; lvl++) {
                const uint32_t right = __shfl_down_sync(0xFFFFFFFFu, crc, 1 << lvl);
                if ((tid & ((2 << lvl) - 1)) == 0) crc = crc_multmodp(c_xpow[lvl][L], crc) ^ right;
            }
            if (tid == 0) {
                const uint32_t given = (uint32_t)s_bytes[n] | ((uint32_t)s_bytes[n + 1] << 8) |
                                       ((uint32_t)s_bytes[n + 2] << 16) | ((uint32_t)s_bytes[n + 3] << 24);
                s_flag = (crc == given) ? B200RX_ST_OK : B200RX_ST_CRC_FAIL;
            }
        }
        __syncthreads();
        const int status = (int)s_flag;

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
}

} // namespace

cudaError_t upload_viterbi_tables(const uint32_t *crc, const uint8_t *scr, const uint32_t *xpow)
{
    cudaError_t e = cudaMemcpyToSymbol(c_crc_tab, crc, sizeof(uint32_t) * 1024);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_xpow, xpow, sizeof(uint32_t) * 5 * 132);
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

// Descriptors of the Viterbi-only entry point (ACS generation 3 reads the caller's symbols in place).
__global__ void desc_from_bits_kernel(const uint32_t *data_bits, uint32_t n_frames, FrameDesc *desc, uint32_t max_steps)
{
    const uint32_t frame = blockIdx.x * blockDim.x + threadIdx.x;
    if (frame >= n_frames) return;
    const uint32_t nb = data_bits[frame];
    const uint32_t steps = nb + 6u;
    const bool ok = steps <= max_steps && (steps & 1u) == 0u;
    FrameDesc d;
    d.n_steps = ok ? steps : 0u; d.data_bits = ok ? nb : 0u; d.field = 0; d.length = 0;
    d.rate = B200RX_RATE_INVALID; d.status = ok ? B200RX_ST_OK : B200RX_ST_TOO_LONG;
    desc[frame] = d;
}

cudaError_t launch_desc_from_bits(const uint32_t *data_bits, uint32_t n_frames, FrameDesc *desc, uint32_t max_steps, cudaStream_t s)
{
    if (n_frames == 0) return cudaSuccess;
    desc_from_bits_kernel<<<(n_frames + 127) / 128, 128, 0, s>>>(data_bits, n_frames, desc, max_steps);
    return cudaGetLastError();
}

cudaError_t launch_viterbi_acs(const FrameDesc *desc, const uint32_t *bm, uint32_t bm_stride, uint32_t *dec,
                               uint32_t dec_stride_words, uint32_t n_frames, const Tuning &tn, cudaStream_t s)
{
    if (n_frames == 0) return cudaSuccess;
    // Lanes per frame (measured on B200, 4096 frames x 12 096 steps: 4 lanes 1.19 ms, 8 lanes 0.98 ms, 16 lanes
    // 1.07 ms): 8 by default, 16 when the batch is too small to give every scheduler (4 per SM) a warp.
    const uint32_t sched = 4u * (uint32_t)tn.sm_count;
    int lb = (n_frames * 8u >= 32u * sched * 3u / 4u) ? 3 : 4;
    if (tn.acs_lb >= 2 && tn.acs_lb <= 5) lb = tn.acs_lb;
    int cta_warps = tn.acs_warps;
    if (cta_warps != 1 && cta_warps != 2 && cta_warps != 4) cta_warps = ACS2_DEFAULT_WARPS;
    const uint32_t per_cta = (uint32_t)cta_warps * (32u >> lb);
    const uint32_t grid = (n_frames + per_cta - 1) / per_cta;
    const int rn = (tn.acs_rn == 0) ? 0 : 1;
#define ACS2_LAUNCH(LBV, RNV, WV) viterbi_acs2_kernel<LBV, RNV, WV><<<grid, 32 * WV, 0, s>>>(desc, bm, bm_stride, dec, dec_stride_words, n_frames, 0xFFFFFFFFu)
#define ACS2_LAUNCH_W(LBV, RNV) do { if (cta_warps == 1) ACS2_LAUNCH(LBV, RNV, 1); else if (cta_warps == 2) ACS2_LAUNCH(LBV, RNV, 2); else ACS2_LAUNCH(LBV, RNV, 4); } while (0)
#define ACS2_LAUNCH_RN(LBV) do { if (rn == 0) ACS2_LAUNCH_W(LBV, 0); else ACS2_LAUNCH_W(LBV, 1); } while (0)
    if (lb == 2) ACS2_LAUNCH_RN(2);
    else if (lb == 3) ACS2_LAUNCH_RN(3);
    else if (lb == 4) ACS2_LAUNCH_RN(4);
    else ACS2_LAUNCH_RN(5);
#undef ACS2_LAUNCH_RN
#undef ACS2_LAUNCH_W
#undef ACS2_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_viterbi_acs3(const FrameDesc *desc, const uint8_t *soft, uint64_t soft_stride, uint32_t *dec,
                                uint32_t dec_stride_words, uint32_t n_frames, bool guard, const Tuning &tn, cudaStream_t s)
{
    if (n_frames == 0) return cudaSuccess;
    // Lanes per frame pair: 4 (16 frames per warp) once the batch gives every scheduler a warp that way, else 8.
    const uint32_t sched = 4u * (uint32_t)tn.sm_count;
    // (bench.py, config 2, 12 batches in flight: LB = 2 with 4-warp CTAs 0.797 ms per step, 2-warp CTAs 0.82-0.84, LB = 3 0.844)
    const uint64_t eff = (uint64_t)n_frames * (uint64_t)(tn.inflight > 1 ? tn.inflight : 1);
    int lb = (eff >= 16u * sched * 3u / 4u) ? 2 : 3;
    if (tn.acs_lb >= 2 && tn.acs_lb <= 3) lb = tn.acs_lb;
    int cta_warps = tn.acs_warps;
    if (cta_warps != 1 && cta_warps != 2 && cta_warps != 4) cta_warps = 4;
    const uint32_t per_cta = (uint32_t)cta_warps * (64u >> lb); // frames per CTA
    const uint32_t grid = (n_frames + per_cta - 1) / per_cta;
#define ACS3_LAUNCH(LBV, WV, GV, ZV) viterbi_acs3_kernel<LBV, WV, GV, ZV><<<grid, 32 * WV, 0, s>>>(desc, soft, soft_stride, dec, dec_stride_words, n_frames, 0xFFFFFFFFu)
#define ACS3_LAUNCH_Z(LBV, WV, GV) do { if (tn.acs_rn == 0) ACS3_LAUNCH(LBV, WV, GV, false); else ACS3_LAUNCH(LBV, WV, GV, true); } while (0)
#define ACS3_LAUNCH_G(LBV, WV) do { if (guard) ACS3_LAUNCH_Z(LBV, WV, true); else ACS3_LAUNCH_Z(LBV, WV, false); } while (0)
#define ACS3_LAUNCH_W(LBV) do { if (cta_warps == 1) ACS3_LAUNCH_G(LBV, 1); else if (cta_warps == 2) ACS3_LAUNCH_G(LBV, 2); else ACS3_LAUNCH_G(LBV, 4); } while (0)
    if (lb == 2) ACS3_LAUNCH_W(2);
    else ACS3_LAUNCH_W(3);
#undef ACS3_LAUNCH_W
#undef ACS3_LAUNCH_G
#undef ACS3_LAUNCH_Z
#undef ACS3_LAUNCH
    return cudaGetLastError();
}

// Header-only decode: copy (rate, length, status) out of the descriptors the front end wrote.
__global__ void export_headers_kernel(const FrameDesc *desc, uint32_t n, uint16_t *len, uint8_t *rate, uint8_t *status)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const FrameDesc d = desc[f];
    if (len) len[f] = d.length;
    if (rate) rate[f] = d.rate;
    if (status) status[f] = d.status;
}

cudaError_t launch_export_headers(const FrameDesc *desc, uint32_t n, uint16_t *len, uint8_t *rate, uint8_t *status,
                                  cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    export_headers_kernel<<<(n + 127) / 128, 128, 0, s>>>(desc, n, len, rate, status);
    return cudaGetLastError();
}

cudaError_t prepare_device_functions()
{
    constexpr size_t dyn = sizeof(uint32_t) * TB_THREADS * TB_TILE_W;
    return cudaFuncSetAttribute(traceback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
}

cudaError_t launch_traceback(const TracebackArgs &a, const Tuning &tn, cudaStream_t s)
{
    if (a.n_frames == 0) return cudaSuccess;
    constexpr size_t dyn = sizeof(uint32_t) * TB_THREADS * TB_TILE_W;
    const uint32_t grid = min((uint32_t)(tn.sm_count * TB_CTAS_PER_SM), a.n_frames);
    traceback_kernel<<<grid, TB_THREADS, dyn, s>>>(a);
    return cudaGetLastError();
}

} // namespace b200rx
