// Front end of the batched receive path: everything the reference does between the tagged sample
// stream and the Viterbi decoder, fused into one kernel so that no intermediate touches HBM:
//
//   fft_symbols::work   (src/fft_symbols.cpp:33-80)   CP strip + 64-point forward FFT, shifted order
//   channel_est::work   (src/channel_est.cpp:36-85)   H^-1 = mean of the two LTS inverses; equalise
//   phase_tracker::work (src/phase_tracker.cpp:70-105) 4-pilot common phase, derotate, keep 48 carriers
//   ppdu::decode_header (src/ppdu.cpp:168-218)        SIGNAL: BPSK demap, deinterleave, 24-step Viterbi
//   modulator::demodulate (src/modulator.cpp:108-164, src/qam.h:110-125)   soft bits 0..255
//   interleaver::deinterleave (src/interleaver.cpp:28-38)                  48-element permutation
//   puncturer::depuncture (src/puncturer.cpp:78-123)                       erasures = 127
//   + the branch-metric arithmetic of viterbi.cpp:234-248 (4 distinct values per trellis step)
//
// One CTA per frame, one warp per OFDM symbol.  A warp loads the 64 useful samples of its symbol with
// one coalesced 16-byte load per lane pair (2 samples per lane), runs the FFT in registers
// (radix-2 DIF, 1 in-lane + 5 shuffle stages), and hands the bins to the rest of the warp through a
// 1 KB shared-memory tile.  HBM traffic per symbol: 1024 B of samples in, 4*dbps B of branch metrics out.
// All arithmetic before the quantiser is fp64 like the reference.
#include "rx_internal.cuh"
#include "viterbi_core.cuh"
#include "fft64.cuh"

#include <math.h>

namespace b200rx {

namespace {

// Warps per frame.  Measured with six batches in flight (bench.py): 8 warps 0.936 ms per step, 4 warps 0.907 ms, 2 warps
// 0.912 ms - small CTAs (10 per SM) interleave better with the Viterbi kernel's, and fewer warps idle while warp 0 decodes SIGNAL.
constexpr int FE_WARPS = 4;
constexpr int SOFT_STRIDE = 292;   // 288 soft bits + the erasure sentinel, padded to a multiple of 4
constexpr int ERASURE_AT = 288;

// Per rate: for trellis step t of an OFDM symbol the two soft-bit indices (demodulator order) it consumes,
// i.e. depuncture (puncturer.cpp:94-118) composed with deinterleave (interleaver.cpp:31-37); ERASURE_AT marks a
// re-inserted erasure (value 127).  Filled by upload_frontend_tables from step_index_pair() below.
__device__ uint32_t c_step_idx[11][216]; // global memory, not __constant__: every thread reads its own entry (a constant
                                         // bank serves one address per cycle, 32 replays per warp-wide load)

// Inverse of c_step_idx for the data kernel: soft byte i (demodulator order) of an OFDM symbol -> its byte offset 2 * t + slot
// in the symbol's run of depunctured soft-symbol pairs; the offsets no soft byte maps to are the re-inserted erasures.
__device__ uint16_t c_scatter[11][288];
// QAM<N>::decode (qam.h:110-125) of one axis as a table: soft bits of pt = clamp(u, -320, 320), byte i = bit i, for
// N = 1, 2, 3 (beyond +-320 every bit has reached its final value 0 or 255).
constexpr int SOFT_TAB_HALF = 320, SOFT_TAB_N = 2 * SOFT_TAB_HALF + 1;
__device__ uint32_t c_soft_tab[3][SOFT_TAB_N];

__device__ double2 c_twiddle[64];    // exp(-2 pi i k / 64)
__constant__ int8_t c_polarity[127]; // pilot polarity sequence (phase_tracker.cpp:23-32)

// LTS_FREQ_DOMAIN (preamble.h:363-429; IEEE 802.11a 17.3.3 L_{-26..26}), index s <-> subcarrier s - 32.
// Bit s of NONZERO: carrier is used; bit s of NEG: value is -1.
// (pinned against the reference's table by tests/test_tables.py)
constexpr unsigned long long LTS_NONZERO = 0x07FFFFFEFFFFFFC0ull;
constexpr unsigned long long LTS_NEG = 0x00567D4C0A605300ull;

__device__ __forceinline__ int shifted_index(int lane, int slot) { return ((slot ^ 1) << 5) | (int)(__brev((unsigned)lane) >> 27); }

// QAM<N>::decode (qam.h:110-125) for one axis: pt = (int)(x * scale) truncates toward zero
template <int NBITS>
__device__ __forceinline__ void qam_decode_axis(double x, double scale, uint8_t *bits)
{
    // qam.h:112-122 is: bit_i = clamp(flip * pt + 128); b = pt < 0 ? -1 : 1; pt -= b * amp; flip = -b; amp /= 2.
    // With u_0 = pt the quantity flip * pt obeys u_{i+1} = amp_i - |u_i| (pt - b * amp = b * (|pt| - amp), times -b),
    // so each further soft bit is one abs, one subtract and one clamp.
    int u = __double2int_rz(x * scale);
    int amp = (1 << (NBITS - 1)) << (8 - NBITS);
#pragma unroll
    for (int i = 0; i < NBITS; i++) {
        bits[i] = (uint8_t)min(max(u + 128, 0), 255);
        u = amp - abs(u);
        amp >>= 1;
    }
}

// d_scale_d of qam.h:50 for the four constellations the reference instantiates (modulator.cpp:120-153)
__device__ __forceinline__ double demap_scale(int bpsc)
{
    switch (bpsc) {
        case 1: return 128.0;                         // QAM<1>(1.0): 2^7 / sqrt(1)
        case 2: return 128.0 / sqrt(0.5);             // QAM<1>(0.5): 2^7 / sqrt(0.5)
        case 4: return 64.0 / sqrt(0.5 * 2.0 / 10.0); // QAM<2>(0.5): 2^6 / sqrt(0.5 * 2 / 10)
        default: return 32.0 / sqrt(0.5 * 4.0 / 84.0); // QAM<3>(0.5): 2^5 / sqrt(0.5 * 4 / 84)
    }
}

// Deinterleaved soft bit k of the current OFDM symbol (interleaver.cpp:31-37: BitInterleave(48, 1) always):
// out[blk + k'] = in[blk + 3 * (k' % 16) + k' / 16]
__device__ __forceinline__ uint32_t deint_at(const uint8_t *soft, int k)
{
    const int blk = (k / 48) * 48, kk = k - blk;
    return soft[blk + 3 * (kk & 15) + (kk >> 4)];
}

__host__ __device__ inline uint32_t deint_index(int k)
{
    const int blk = (k / 48) * 48, kk = k - blk;
    return (uint32_t)(blk + 3 * (kk & 15) + (kk >> 4));
}

// (i0 | i1 << 16) for trellis step t under puncturing mode punc
__host__ __device__ inline uint32_t step_index_pair(int punc, int t)
{
    uint32_t i0, i1;
    if (punc == PUNC_1_2) { i0 = deint_index(2 * t); i1 = deint_index(2 * t + 1); }
    else if (punc == PUNC_3_4) {
        const int g = t / 3, r = t - 3 * g;
        if (r == 0) { i0 = deint_index(4 * g); i1 = deint_index(4 * g + 1); }
        else { i0 = ERASURE_AT; i1 = deint_index(4 * g + 1 + r); }
    } else {
        const int g = t >> 1;
        if ((t & 1) == 0) { i0 = deint_index(3 * g); i1 = ERASURE_AT; }
        else { i0 = deint_index(3 * g + 1); i1 = deint_index(3 * g + 2); }
    }
    return i0 | (i1 << 16);
}

// The two depunctured soft symbols of trellis step t (0-based within the OFDM symbol), puncturer.cpp:94-118
__device__ __forceinline__ void step_symbols(const uint8_t *soft, int punc, int t, uint32_t &s0, uint32_t &s1)
{
    if (punc == PUNC_1_2) {
        s0 = deint_at(soft, 2 * t);
        s1 = deint_at(soft, 2 * t + 1);
    } else if (punc == PUNC_3_4) { // d0 d1 127 d2 127 d3 per 4 coded bits -> 3 steps
        const int g = t / 3, r = t - 3 * g;
        if (r == 0) { s0 = deint_at(soft, 4 * g); s1 = deint_at(soft, 4 * g + 1); }
        else if (r == 1) { s0 = 127u; s1 = deint_at(soft, 4 * g + 2); }
        else { s0 = 127u; s1 = deint_at(soft, 4 * g + 3); }
    } else { // d0 127 d1 d2 per 3 coded bits -> 2 steps
        const int g = t >> 1;
        if ((t & 1) == 0) { s0 = deint_at(soft, 3 * g); s1 = 127u; }
        else { s0 = deint_at(soft, 3 * g + 1); s1 = deint_at(soft, 3 * g + 2); }
    }
}

// Sample fetch.  ROT: multiply by the constant rotation timing_sync applies (timing_sync.cpp:121-125:
// input[x].sample *= (cos(m_phase_acc), sin(m_phase_acc)); std::complex operator*= evaluated without contraction).
struct RotCtx {
    double2 rn, ro;  // rotation for relative index >= from / < from
    int64_t from;    // relative to the frame's LTS1
};

// The samples of one frame: `base` in format FMT, the frame's LTS1 tag at sample index `p`.
struct Window {
    const void *base;
    uint64_t p;
    double scale;
};

template <bool ROT, int FMT>
__device__ __forceinline__ double2 fetch(const Window &win, int k, const RotCtx &rc)
{
    double2 v = load_sample<FMT>(win.base, win.p + (uint64_t)k, win.scale);
    if constexpr (ROT) {
        const double2 r = ((int64_t)k >= rc.from) ? rc.rn : rc.ro;
        v = make_double2(__dsub_rn(__dmul_rn(v.x, r.x), __dmul_rn(v.y, r.y)),
                         __dadd_rn(__dmul_rn(v.x, r.y), __dmul_rn(v.y, r.x)));
    }
    return v;
}

struct SymbolCtx {
    const double2 *tw;   // smem twiddles
    const double2 *hinv; // smem inverse channel
    double2 *xs;         // smem, this warp's 64 bins (shifted order)
    uint8_t *soft;       // smem, this warp's soft bits (up to 288)
};

// One OFDM symbol: samples -> equalised, derotated data carriers -> soft bits in ctx.soft.
// v = symbol index from SIGNAL (0) on: selects the pilot polarity (phase_tracker.cpp:77-86).
template <bool ROT, int FMT, bool DBG>
__device__ __forceinline__ void process_symbol(const SymbolCtx &ctx, const Window &win, int off, const RotCtx &rc, int v,
                                               int bpsc, int lane, double2 *dbg_eq)
{
    double2 v0 = fetch<ROT, FMT>(win, off + lane, rc), v1 = fetch<ROT, FMT>(win, off + lane + 32, rc);
    warp_fft64(v0, v1, ctx.tw, lane);
    ctx.xs[shifted_index(lane, 0)] = v0;
    ctx.xs[shifted_index(lane, 1)] = v1;
    __syncwarp();

    // phase_tracker.cpp:83-92: e = sum_p rec_p * conj(ref_p) / 4, ref_p = sign_p * POLARITY[v % 127] (real).
    // Lane l takes pilot l & 3; two butterfly additions give every lane the sum of the four terms.
    const double pol = (double)c_polarity[v % 127];
    double2 e;
    {
        const int p = lane & 3, bin = 11 + 14 * p;
        const double ref = (p == 3) ? -pol : pol;
        const double2 rec = cmul(ctx.hinv[bin], ctx.xs[bin]);
        e = make_double2(rec.x * ref / 4.0, rec.y * ref / 4.0);
        e = cadd(e, shfl_xor2(e, 1));
        e = cadd(e, shfl_xor2(e, 2));
    }
    // phase_tracker.cpp:92-98 rotates by exp(-i arg(e)) = conj(e) / |e|
    const double m2 = e.x * e.x + e.y * e.y;
    const double inv = rsqrt(m2); // 1 / |e|
    const double2 rot = (m2 > 0.0) ? make_double2(e.x * inv, -e.y * inv) : make_double2(1.0, 0.0);

    const double scale = demap_scale(bpsc);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int c = lane + 32 * h;
        if (c < 48) {
            const int bin = data_bin(c);
            const double2 y = cmul(cmul(ctx.hinv[bin], ctx.xs[bin]), rot);
            if constexpr (DBG) { if (dbg_eq) dbg_eq[c] = y; }
            uint8_t *o = ctx.soft + c * bpsc;
            switch (bpsc) { // modulator.cpp:118-160: real axis first, then imaginary
                case 1: qam_decode_axis<1>(y.x, scale, o); break;
                case 2: qam_decode_axis<1>(y.x, scale, o); qam_decode_axis<1>(y.y, scale, o + 1); break;
                case 4: qam_decode_axis<2>(y.x, scale, o); qam_decode_axis<2>(y.y, scale, o + 2); break;
                default: qam_decode_axis<3>(y.x, scale, o); qam_decode_axis<3>(y.y, scale, o + 3); break;
            }
        }
    }
    __syncwarp();
}

// DBG: the parity-test taps (equalised points, depunctured soft symbols) are compiled in only for calls that ask for them
template <bool ROT, int FMT, bool DBG>
__global__ void __launch_bounds__(FE_WARPS * 32, 40 / FE_WARPS) frontend_kernel(FrontendArgs a)
{
    __shared__ double2 s_tw[64];
    __shared__ double2 s_hinv[64];
    __shared__ double2 s_xs[FE_WARPS][64];
    __shared__ uint8_t s_soft[FE_WARPS][SOFT_STRIDE]; // [288] = 127: the erasure the index table points at
    __shared__ uint32_t s_idx[216];                   // soft-bit indices (i0 | i1 << 16) of each trellis step of a symbol
    __shared__ uint32_t s_hdr_bm[32];
    __shared__ FrameDesc s_desc;

    const int frame = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if ((a.n_live && (uint32_t)frame >= *a.n_live) || (a.select && !a.select[frame])) { // slot without a frame (raw-capture
                                                                                       // entry points) / frame not asked for
        if (tid == 0) {
            FrameDesc d;
            d.n_steps = 0; d.data_bits = 0; d.field = 0; d.length = 0;
            d.rate = B200RX_RATE_INVALID; d.status = B200RX_ST_NO_FRAME;
            a.desc[frame] = d;
        }
        return;
    }
    const uint64_t p = a.lts1[frame];
    uint32_t avail = a.avail[frame];
    const uint64_t iq_samples = a.sp ? a.sp->n_samples : a.iq_samples;
    if (p >= iq_samples) avail = 0;
    else if ((uint64_t)avail > iq_samples - p) avail = (uint32_t)(iq_samples - p);
    const Window win{a.iq, p, a.scale};
    RotCtx rc{};
    if constexpr (ROT) {
        const FrameRot fr = a.rot[frame];
        rc.rn = fr.rot_new;
        rc.ro = fr.rot_old;
        rc.from = (int64_t)fr.from - (int64_t)p;
    }

    if (tid < 64) s_tw[tid] = c_twiddle[tid];
    if (tid < FE_WARPS) s_soft[tid][ERASURE_AT] = 127;
    if (tid == 0) {
        s_desc.n_steps = 0; s_desc.data_bits = 0; s_desc.field = 0; s_desc.length = 0;
        s_desc.rate = B200RX_RATE_INVALID; s_desc.status = B200RX_ST_TRUNCATED;
    }
    __syncthreads();

    // fft_symbols.cpp:42-71 with LTS1 at sample 0 and LTS2 at 64: windows [0,64) and [64,128);
    // the SIGNAL symbol occupies [144, 208).  Nothing is decodable below 208 samples.
    if (avail < 208) {
        if (tid == 0) a.desc[frame] = s_desc;
        return;
    }

    SymbolCtx ctx{s_tw, s_hinv, s_xs[warp], s_soft[warp]};

    // ---- channel estimate: channel_est.cpp:53-58, H^-1[j] = sum_{2 LTS} L[j] / R[j] / 2 for all 64 bins ----
    if (warp < 2) {
        double2 v0 = fetch<ROT, FMT>(win, 64 * warp + lane, rc), v1 = fetch<ROT, FMT>(win, 64 * warp + lane + 32, rc);
        warp_fft64(v0, v1, s_tw, lane);
#pragma unroll
        for (int slot = 0; slot < 2; slot++) {
            const double2 r = slot ? v1 : v0;
            const int s = shifted_index(lane, slot);
            double l = ((LTS_NONZERO >> s) & 1ull) ? (((LTS_NEG >> s) & 1ull) ? -1.0 : 1.0) : 0.0;
            // L / R with L real: L * conj(R) / |R|^2, then / 2
            const double den = r.x * r.x + r.y * r.y;
            s_xs[warp][s] = make_double2(l * r.x / den / 2.0, -l * r.y / den / 2.0);
        }
    }
    __syncthreads();
    if (tid < 64) {
        s_hinv[tid] = cadd(s_xs[0][tid], s_xs[1][tid]);
        if (a.hinv_out) a.hinv_out[(size_t)frame * 64 + tid] = s_hinv[tid]; // for data_kernel (split front end)
    }
    __syncthreads();

    // ---- SIGNAL: ppdu.cpp:168-218 ----
    if (warp == 0) {
        double2 *dbg = nullptr;
        if constexpr (DBG) { if (a.dbg_eq && a.dbg_eq_vectors > 0) dbg = a.dbg_eq + ((size_t)frame * a.dbg_eq_vectors) * 48; }
        process_symbol<ROT, FMT, DBG>(ctx, win, 128 + 16, rc, 0, 1, lane, dbg);
        if (lane < 24) {
            uint32_t s0, s1;
            step_symbols(ctx.soft, PUNC_1_2, lane, s0, s1);
            s_hdr_bm[lane] = bm_word(s0, s1);
        }
        __syncwarp();
        // 18 data bits -> 24 trellis steps -> 3 bytes MSB first (ppdu.cpp:181-186)
        const uint32_t field = warp_viterbi_short(s_hdr_bm, 24, 18, lane);
        if (lane == 0) {
            s_desc.field = field;
            uint32_t x = field; // parity.h:43-48
            x ^= x >> 16; x ^= x >> 8;
            const int par = __popc(x & 0xFFu) & 1;
            const int rate = rate_from_field((field >> 19) & 0xF); // ppdu.cpp:194
            const uint32_t len = (field >> 6) & 0xFFF;             // ppdu.cpp:195
            if (par) s_desc.status = B200RX_ST_HDR_PARITY;
            else if (rate == 255) s_desc.status = B200RX_ST_HDR_RATE;
            else {
                const uint32_t nsym = num_symbols(rate, len);
                const uint32_t steps = nsym * rate_row(rate).dbps;
                s_desc.rate = (uint8_t)rate;
                s_desc.length = (uint16_t)len;
                if (len > a.max_len || steps > a.max_steps) s_desc.status = B200RX_ST_TOO_LONG;
                else if (a.header_only != 1 && (uint64_t)avail < 128ull + 80ull * (1ull + nsym)) s_desc.status = B200RX_ST_TRUNCATED;
                else {
                    s_desc.status = B200RX_ST_OK;
                    s_desc.n_steps = steps;
                    s_desc.data_bits = steps - 6;
                }
            }
        }
    }
    __syncthreads();
    const FrameDesc d = s_desc;
    if (tid == 0) a.desc[frame] = d;
    if (d.status != B200RX_ST_OK || a.header_only) return;

    // ---- data symbols ----
    const RateRow rr = rate_row(d.rate);
    for (int t = tid; t < rr.dbps; t += FE_WARPS * 32) s_idx[t] = c_step_idx[d.rate][t];
    __syncthreads();
    const uint32_t nsym = d.n_steps / rr.dbps;
    uint32_t *bm_out = a.bm + (size_t)frame * a.bm_stride;
    for (uint32_t s = warp; s < nsym; s += FE_WARPS) {
        double2 *dbg = nullptr;
        if constexpr (DBG) { if (a.dbg_eq && s + 1 < a.dbg_eq_vectors) dbg = a.dbg_eq + ((size_t)frame * a.dbg_eq_vectors + s + 1) * 48; }
        process_symbol<ROT, FMT, DBG>(ctx, win, 128 + 80 * (int)(s + 1) + 16, rc, (int)(s + 1), rr.bpsc, lane, dbg);
        uint32_t *sym_out = bm_out + (size_t)s * rr.dbps;
        unsigned short *pair_out = reinterpret_cast<unsigned short *>(bm_out) + (size_t)s * rr.dbps;
        for (int t = lane; t < rr.dbps; t += 32) {
            const uint32_t pair = s_idx[t];
            const uint32_t s0 = ctx.soft[pair & 0xFFFFu], s1 = ctx.soft[pair >> 16];
            const size_t step = (size_t)s * rr.dbps + t;
            if (a.emit_pairs) pair_out[t] = (unsigned short)(s0 | (s1 << 8));
            else sym_out[t] = bm_word(s0, s1);
            if constexpr (DBG) {
                if (a.dbg_depunct && 2 * step + 1 < a.dbg_depunct_stride) {
                    uint8_t *dp = a.dbg_depunct + (size_t)frame * a.dbg_depunct_stride + 2 * step;
                    dp[0] = (uint8_t)s0; dp[1] = (uint8_t)s1;
                }
            }
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// Data symbols of the split front end (Tuning::fe_split): everything frontend_kernel does after SIGNAL, for frames whose
// descriptor and inverse channel the header pass (frontend_kernel, header_only = 2) has written.
//
// ncu on frontend_kernel (profiles/r01_v9_ncu_full.md): 559 warp instructions per OFDM symbol, 2.0 TB/s - bound by
// instruction issue, not by HBM.  Here a symbol costs about a quarter of that:
//   * 8 lanes per symbol, 8 points per lane, four symbols per warp: the 64-point DFT is 8-point DFTs in registers
//     (constant twiddles, half of them trivial), one transposition through shared memory, 8-point DFTs again -
//     no shuffles (the two-points-per-lane radix-2 version spends 40 SHFL + 5 twiddle multiplies per lane per symbol);
//   * equalise, pilots, derotation and demapping run on the registers the DFT left behind (lane r holds bins r + 8 k);
//   * QAM<N>::decode is a table look-up on the truncated integer (qam.h:112), built from the same recurrence;
//   * deinterleave + depuncture is a scatter: every soft byte is stored once, at the place it has in the symbol's run of
//     (s0, s1) pairs (c_scatter), into a tile whose erasure bytes were set to 127 once; the tile leaves with 8-byte
//     coalesced stores.
// One CTA (4 warps) per frame; no barrier after the setup.
// ------------------------------------------------------------------------------------------------
constexpr int DK_WARPS = 4;
#ifndef DK_MIN_BLOCKS
#define DK_MIN_BLOCKS 5 // CTAs per SM the register allocation aims at (5: 96 registers, no spills)
#endif
constexpr int DK_TR_STRIDE = 9;      // double2 per transposition row (8 + 1: 16-byte accesses of 8 lanes hit 8 bank groups)
constexpr int DK_PAIR_BYTES = 432;   // 2 * max dbps

// In-place 8-point forward DFT, natural order in and out.
__device__ __forceinline__ void dft8(double2 (&x)[8])
{
    const double h = 0.70710678118654752440; // 1 / sqrt(2)
    const double2 a0 = cadd(x[0], x[4]), b0 = csub(x[0], x[4]);
    const double2 a1 = cadd(x[1], x[5]), t1 = csub(x[1], x[5]);
    const double2 a2 = cadd(x[2], x[6]), t2 = csub(x[2], x[6]);
    const double2 a3 = cadd(x[3], x[7]), t3 = csub(x[3], x[7]);
    const double2 b1 = make_double2((t1.x + t1.y) * h, (t1.y - t1.x) * h);   // * (1 - i) / sqrt 2
    const double2 b2 = make_double2(t2.y, -t2.x);                             // * -i
    const double2 b3 = make_double2((t3.y - t3.x) * h, (-t3.x - t3.y) * h);  // * (-1 - i) / sqrt 2
    {
        const double2 s0 = cadd(a0, a2), d0 = csub(a0, a2), s1 = cadd(a1, a3), u = csub(a1, a3);
        const double2 d1 = make_double2(u.y, -u.x);
        x[0] = cadd(s0, s1); x[4] = csub(s0, s1); x[2] = cadd(d0, d1); x[6] = csub(d0, d1);
    }
    {
        const double2 s0 = cadd(b0, b2), d0 = csub(b0, b2), s1 = cadd(b1, b3), u = csub(b1, b3);
        const double2 d1 = make_double2(u.y, -u.x);
        x[1] = cadd(s0, s1); x[5] = csub(s0, s1); x[3] = cadd(d0, d1); x[7] = csub(d0, d1);
    }
}

struct DataArgs {
    const void *iq;
    double scale;
    uint64_t iq_samples;
    const uint64_t *lts1;
    uint32_t n_frames;
    const FrameDesc *desc;
    const double2 *hinv;     // [n_frames][64], shifted order
    uint8_t *pairs;          // [n_frames][pair_stride] bytes: (s0, s1) per trellis step
    uint64_t pair_stride;
    const FrameRot *rot;
    double2 *dbg_eq;
    uint32_t dbg_eq_vectors;
    uint8_t *dbg_depunct;
    uint32_t dbg_depunct_stride;
};

// The symbol loop of data_kernel for one modulation (BPSC soft bits per carrier): specialised so that the per-carrier
// work has no branches.
template <bool ROT, int FMT, bool DBG, int BPSC>
__device__ __forceinline__ void data_symbols(const DataArgs &a, int frame, const FrameDesc &d, const Window &win, const RotCtx &rc,
                                             double2 *tr_row, const double2 *tr_col, double2 *pil, const double2 *hinv_r,
                                             const double2 *tw_r, const uint32_t *soft_tab, const uint16_t *scatter,
                                             uint8_t *tile, int lane, int warp)
{
    const int q = lane >> 3, r = lane & 7;
    const int dbps = rate_row(d.rate).dbps;
    const uint32_t nsym = d.n_steps / (uint32_t)dbps;
    const double scale = demap_scale(BPSC);
    // which of this lane's bins r + 8 k2 carry data, and where the carrier's soft bytes go in the symbol's run of pairs
    const uint16_t *sc[8];
    bool is_data[8];
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) {
        const int s = (r + 8 * k2 + 32) & 63; // shifted index: subcarrier s - 32
        is_data[k2] = s >= 6 && s <= 58 && s != 11 && s != 25 && s != 32 && s != 39 && s != 53;
        const int c = s - 6 - (s > 11) - (s > 25) - (s > 32) - (s > 39) - (s > 53);
        sc[k2] = scatter + (is_data[k2] ? c * BPSC : 0);
    }
    // pilot of this lane (odd lanes): pilot number pp (phase_tracker.cpp:37-43: bins 11, 25, 39, 53 = r 3, 1, 7, 5)
    const int pp = r == 3 ? 0 : (r == 1 ? 1 : (r == 7 ? 2 : 3));
    const double psign = (pp == 3) ? -0.25 : 0.25; // sign_p / 4: an exact scaling, so (x * ref) / 4 == x * (ref / 4)
    uint8_t *st = tile + q * 2 * dbps;
    uint8_t *out = a.pairs + (size_t)frame * a.pair_stride;

    for (uint32_t g = warp; 4 * g < nsym; g += DK_WARPS) {
        const uint32_t sym = 4 * g + q;
        const bool live = sym < nsym;
        const int off = 128 + 80 * (int)(live ? sym + 1 : 1) + 16 + r; // dead lanes of the last group re-read symbol 0
        double2 x[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; n1++) x[n1] = fetch<ROT, FMT>(win, off + 8 * n1, rc);
        // X[k1 + 8 k2] = sum_r W64^(r k1) W8^(r k2) [ sum_n1 x[8 n1 + r] W8^(n1 k1) ]
        dft8(x);
#pragma unroll
        for (int k1 = 1; k1 < 8; k1++) x[k1] = cmul(x[k1], tw_r[8 * k1]);
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) tr_row[k1] = x[k1];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = tr_col[j * DK_TR_STRIDE];
        dft8(x); // x[k2] = X[r + 8 k2]

        // channel_est.cpp:77-81: out = H^-1 * in   (hinv_r[k2] = H^-1 of bin r + 8 k2)
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) x[k2] = cmul(hinv_r[8 * k2], x[k2]);
        // phase_tracker.cpp:83-92: e = sum_p rec_p * (sign_p * POLARITY[n % 127]) / 4, summed as (e0 + e1) + (e2 + e3)
        {
            const double ref4 = (double)c_polarity[(live ? sym + 1 : 1) % 127] * psign;
            const double2 pv = r == 3 ? x[5] : (r == 1 ? x[7] : (r == 7 ? x[0] : x[2]));
            if (r & 1) pil[pp] = make_double2(pv.x * ref4, pv.y * ref4);
        }
        __syncwarp();
        const double2 e = cadd(cadd(pil[0], pil[1]), cadd(pil[2], pil[3]));
        // phase_tracker.cpp:92-98 rotates by exp(-i arg(e)) = conj(e) / |e|
        const double m2 = e.x * e.x + e.y * e.y;
        const double inv = rsqrt(m2);
        const double2 rot = (m2 > 0.0) ? make_double2(e.x * inv, -e.y * inv) : make_double2(1.0, 0.0);
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            const double2 y = cmul(x[k2], rot);
            const bool on = is_data[k2] && live;
            if constexpr (DBG) {
                if (a.dbg_eq && on && sym + 1 < a.dbg_eq_vectors)
                    a.dbg_eq[((size_t)frame * a.dbg_eq_vectors + sym + 1) * 48 + (sc[k2] - scatter) / BPSC] = y;
            }
            // modulator.cpp:118-160: real axis first, then imaginary (BPSK: real only); qam.h:112 truncates toward zero
            const int ur = min(max(__double2int_rz(y.x * scale), -SOFT_TAB_HALF), SOFT_TAB_HALF);
            const uint32_t br = soft_tab[ur];
            const uint16_t *p = sc[k2];
            if constexpr (BPSC == 1) {
                if (on) st[p[0]] = (uint8_t)br;
            } else {
                const int ui = min(max(__double2int_rz(y.y * scale), -SOFT_TAB_HALF), SOFT_TAB_HALF);
                const uint32_t bi = soft_tab[ui];
                if constexpr (BPSC == 2) {
                    if (on) { st[p[0]] = (uint8_t)br; st[p[1]] = (uint8_t)bi; }
                } else if constexpr (BPSC == 4) {
                    if (on) {
                        st[p[0]] = (uint8_t)br; st[p[1]] = (uint8_t)(br >> 8);
                        st[p[2]] = (uint8_t)bi; st[p[3]] = (uint8_t)(bi >> 8);
                    }
                } else {
                    if (on) {
                        st[p[0]] = (uint8_t)br; st[p[1]] = (uint8_t)(br >> 8); st[p[2]] = (uint8_t)(br >> 16);
                        st[p[3]] = (uint8_t)bi; st[p[4]] = (uint8_t)(bi >> 8); st[p[5]] = (uint8_t)(bi >> 16);
                    }
                }
            }
        }
        __syncwarp();
        // the group's run of pairs: symbols 4g .. 4g+3 are contiguous in the tile and in the frame's pair buffer
        const uint32_t cnt = min(4u, nsym - 4 * g);
        const uint32_t n8 = cnt * 2 * dbps / 8;
        const size_t gofs = (size_t)4 * g * 2 * dbps;
        for (uint32_t i = lane; i < n8; i += 32)
            reinterpret_cast<uint2 *>(out + gofs)[i] = reinterpret_cast<const uint2 *>(tile)[i];
        if constexpr (DBG) {
            if (a.dbg_depunct)
                for (uint32_t i = lane; i < cnt * 2 * dbps; i += 32)
                    if (gofs + i < a.dbg_depunct_stride) a.dbg_depunct[(size_t)frame * a.dbg_depunct_stride + gofs + i] = tile[i];
        }
        __syncwarp();
    }
}

template <bool ROT, int FMT, bool DBG>
__global__ void __launch_bounds__(DK_WARPS * 32, DK_MIN_BLOCKS) data_kernel(DataArgs a)
{
    __shared__ double2 s_hinv[64];      // [k2][r] = H^-1 of DFT bin r + 8 k2 (natural DFT order): lane r reads [k2][r], conflict-free
    __shared__ double2 s_tw[8][8];      // [k1][r] = W64^(r k1)
    __shared__ __align__(16) double2 s_tr[DK_WARPS][4][8][DK_TR_STRIDE];
    __shared__ __align__(16) double2 s_pil[DK_WARPS][4][4];
    __shared__ __align__(16) uint8_t s_pairs[DK_WARPS][4 * DK_PAIR_BYTES];
    __shared__ uint16_t s_scatter[288];
    __shared__ uint32_t s_soft[SOFT_TAB_N];

    // Persistent CTAs: each takes every gridDim.x-th frame, so the tables (672 words) are staged once per CTA and again
    // only when the rate changes, not once per frame (ncu, one CTA per frame: 18 % of the kernel's stall samples sat on
    // the stores of this staging, waiting for divergent constant-bank reads).
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = lane >> 3, r = lane & 7;
    if (tid < 64) s_tw[tid >> 3][tid & 7] = c_twiddle[((tid >> 3) * (tid & 7)) & 63];
    int cur_rate = -1, cur_tab = -1;
    double2 *tr_row = &s_tr[warp][q][r][0];
    const double2 *tr_col = &s_tr[warp][q][0][r];
    double2 *pil = &s_pil[warp][q][0];
    const double2 *hinv_r = &s_hinv[r];
    const double2 *tw_r = &s_tw[0][r];
    const uint32_t *soft_tab = s_soft + SOFT_TAB_HALF;
    uint8_t *tile = &s_pairs[warp][0];
    for (uint32_t frame = blockIdx.x; frame < a.n_frames; frame += gridDim.x) {
        const FrameDesc d = a.desc[frame];
        if (d.status != B200RX_ST_OK || d.n_steps == 0) continue; // the same for every thread of the CTA
        const int bpsc = rate_row(d.rate).bpsc;
        const int tab = bpsc <= 2 ? 0 : (bpsc == 4 ? 1 : 2);
        __syncthreads(); // every warp has finished the previous frame
        if (tid < 64) s_hinv[tid] = a.hinv[(size_t)frame * 64 + ((tid + 32) & 63)]; // a.hinv is in the reference's shifted order
        if ((int)d.rate != cur_rate) {
            for (int i = tid; i < 288; i += DK_WARPS * 32) s_scatter[i] = c_scatter[d.rate][i];
            if (tab != cur_tab)
                for (int i = tid; i < SOFT_TAB_N; i += DK_WARPS * 32) s_soft[i] = c_soft_tab[tab][i];
            for (int i = tid; i < DK_WARPS * 4 * DK_PAIR_BYTES / 4; i += DK_WARPS * 32)
                reinterpret_cast<uint32_t *>(&s_pairs[0][0])[i] = 0x7F7F7F7Fu; // erasures (puncturer.cpp:98): data bytes are overwritten per symbol
            cur_rate = d.rate;
            cur_tab = tab;
        }
        __syncthreads();

        const Window win{a.iq, a.lts1[frame], a.scale};
        RotCtx rc{};
        if constexpr (ROT) {
            const FrameRot fr = a.rot[frame];
            rc.rn = fr.rot_new;
            rc.ro = fr.rot_old;
            rc.from = (int64_t)fr.from - (int64_t)win.p;
        }
        switch (bpsc) {
            case 1: data_symbols<ROT, FMT, DBG, 1>(a, (int)frame, d, win, rc, tr_row, tr_col, pil, hinv_r, tw_r, soft_tab, s_scatter, tile, lane, warp); break;
            case 2: data_symbols<ROT, FMT, DBG, 2>(a, (int)frame, d, win, rc, tr_row, tr_col, pil, hinv_r, tw_r, soft_tab, s_scatter, tile, lane, warp); break;
            case 4: data_symbols<ROT, FMT, DBG, 4>(a, (int)frame, d, win, rc, tr_row, tr_col, pil, hinv_r, tw_r, soft_tab, s_scatter, tile, lane, warp); break;
            default: data_symbols<ROT, FMT, DBG, 6>(a, (int)frame, d, win, rc, tr_row, tr_col, pil, hinv_r, tw_r, soft_tab, s_scatter, tile, lane, warp); break;
        }
    }
}

} // namespace

cudaError_t upload_frontend_tables(const double2 *tw, const int8_t *pol)
{
    static uint32_t idx[11][216];
    static uint16_t scatter[11][288];
    for (int r = 0; r < 11; r++) {
        const RateRow rr = rate_row(r);
        for (int i = 0; i < 288; i++) scatter[r][i] = 0;
        for (int t = 0; t < 216; t++) {
            idx[r][t] = t < rr.dbps ? step_index_pair(rr.punc, t) : (ERASURE_AT | (ERASURE_AT << 16));
            if (t < rr.dbps) {
                const uint32_t i0 = idx[r][t] & 0xFFFFu, i1 = idx[r][t] >> 16;
                if (i0 != ERASURE_AT) scatter[r][i0] = (uint16_t)(2 * t);
                if (i1 != ERASURE_AT) scatter[r][i1] = (uint16_t)(2 * t + 1);
            }
        }
    }
    // QAM<N>::decode (qam.h:110-125) of one axis on the truncated integer pt (see qam_decode_axis)
    static uint32_t soft[3][SOFT_TAB_N];
    for (int nb = 1; nb <= 3; nb++)
        for (int i = 0; i < SOFT_TAB_N; i++) {
            int u = i - SOFT_TAB_HALF, amp = 128;
            uint32_t w = 0;
            for (int b = 0; b < nb; b++) {
                const int v = u + 128;
                w |= (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v)) << (8 * b);
                u = amp - (u < 0 ? -u : u);
                amp >>= 1;
            }
            soft[nb - 1][i] = w;
        }
    cudaError_t e = cudaMemcpyToSymbol(c_step_idx, idx, sizeof(idx));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_scatter, scatter, sizeof(scatter));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_soft_tab, soft, sizeof(soft));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_twiddle, tw, sizeof(double2) * 64);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_polarity, pol, 127);
}

// Data symbols of the split front end; a.desc / a.hinv_out were written by launch_frontend(header_only = 2).
cudaError_t launch_frontend_data(const FrontendArgs &a, cudaStream_t s)
{
    if (a.n_frames == 0) return cudaSuccess;
    DataArgs d{};
    d.iq = a.iq; d.scale = a.scale; d.iq_samples = a.iq_samples; d.lts1 = a.lts1; d.n_frames = a.n_frames; d.desc = a.desc;
    d.hinv = a.hinv_out; d.pairs = reinterpret_cast<uint8_t *>(a.bm); d.pair_stride = (uint64_t)a.bm_stride * 4; d.rot = a.rot;
    d.dbg_eq = a.dbg_eq; d.dbg_eq_vectors = a.dbg_eq_vectors; d.dbg_depunct = a.dbg_depunct; d.dbg_depunct_stride = a.dbg_depunct_stride;
    const uint32_t resident = (uint32_t)(a.sm_count > 0 ? a.sm_count : 148) * DK_MIN_BLOCKS; // persistent grid: what the GPU holds at once
    const dim3 grid(a.n_frames < resident ? a.n_frames : resident), block(DK_WARPS * 32);
    const bool dbg = a.dbg_eq != nullptr || a.dbg_depunct != nullptr;
#define DK_LAUNCH(ROTV, FMTV) do { if (dbg) data_kernel<ROTV, FMTV, true><<<grid, block, 0, s>>>(d); \
                                   else data_kernel<ROTV, FMTV, false><<<grid, block, 0, s>>>(d); } while (0)
    switch (a.fmt * 2 + (a.rot ? 1 : 0)) {
        case FMT_FC64 * 2: DK_LAUNCH(false, FMT_FC64); break;
        case FMT_FC64 * 2 + 1: DK_LAUNCH(true, FMT_FC64); break;
        case FMT_FC32 * 2: DK_LAUNCH(false, FMT_FC32); break;
        case FMT_FC32 * 2 + 1: DK_LAUNCH(true, FMT_FC32); break;
        case FMT_SC16 * 2: DK_LAUNCH(false, FMT_SC16); break;
        case FMT_SC16 * 2 + 1: DK_LAUNCH(true, FMT_SC16); break;
        case FMT_TAGGED * 2: DK_LAUNCH(false, FMT_TAGGED); break; // a tagged stream has been rotated by timing_sync already
        default: return cudaErrorInvalidValue;
    }
#undef DK_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_frontend(const FrontendArgs &a, cudaStream_t s)
{
    if (a.n_frames == 0) return cudaSuccess;
    // the header pass only has work for two warps (LTS1 / LTS2, then SIGNAL on warp 0): half-size CTAs, twice as many resident
    const dim3 grid(a.n_frames), block(a.header_only ? 64 : FE_WARPS * 32);
    const bool dbg = a.dbg_eq != nullptr || a.dbg_depunct != nullptr;
#define FE_LAUNCH(ROTV, FMTV) do { if (dbg) frontend_kernel<ROTV, FMTV, true><<<grid, block, 0, s>>>(a); \
                                   else frontend_kernel<ROTV, FMTV, false><<<grid, block, 0, s>>>(a); } while (0)
    switch (a.fmt * 2 + (a.rot ? 1 : 0)) {
        case FMT_FC64 * 2: FE_LAUNCH(false, FMT_FC64); break;
        case FMT_FC64 * 2 + 1: FE_LAUNCH(true, FMT_FC64); break;
        case FMT_FC32 * 2: FE_LAUNCH(false, FMT_FC32); break;
        case FMT_FC32 * 2 + 1: FE_LAUNCH(true, FMT_FC32); break;
        case FMT_SC16 * 2: FE_LAUNCH(false, FMT_SC16); break;
        case FMT_SC16 * 2 + 1: FE_LAUNCH(true, FMT_SC16); break;
        case FMT_TAGGED * 2: FE_LAUNCH(false, FMT_TAGGED); break;
        default: return cudaErrorInvalidValue;
    }
#undef FE_LAUNCH
    return cudaGetLastError();
}

} // namespace b200rx
