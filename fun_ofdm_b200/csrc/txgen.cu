// Device-side frame generator + channel: the synthetic corpus of the benchmarks built directly in HBM,
// so that large corpora (BASELINE config 5: 2^20 frames) need no host generation and no H2D copy.
//
// What the reference computes for one frame (restated, never copied):
//   frame_builder::build_frame src/frame_builder.cpp:53-82   ppdu::encode -> symbol_mapper::map -> fft::inverse -> CP -> preamble
//   ppdu::encoder_header       src/ppdu.cpp:81-110           24-bit field, parity, K=7 r=1/2 code, interleave, BPSK
//   ppdu::encode_data          src/ppdu.cpp:112-165          service|payload|crc32|pad -> per-byte scrambler -> code ->
//                                                            puncture -> interleave -> modulate
//   viterbi::conv_encode       src/viterbi.cpp:39-62         polys {121, 91}, newest bit in the LSB, tail not forced to zero
//   puncturer::puncture        src/puncturer.cpp:26-70       keeps {0,1,3,5}/6 (3/4) and {0,2,3}/4 (2/3)
//   interleaver::interleave    src/interleaver.cpp:15-26     out[3 * (y % 16) + y / 16] = in[y] per 48, always
//   modulator::modulate        src/modulator.cpp:30-99, QAM<N>::encode src/qam.h:87-99
//   symbol_mapper::map         src/symbol_mapper.cpp:81-119  48 data + 4 pilots (x polarity) + 12 nulls
//   fft::inverse               src/fft.cpp:68-96             x[n] = (1/64) sum_k X[k] e^{+2 pi i k n / 64}
// The channel (multipath FIR + AWGN from a counter-based generator) is this repo's own, the same as the host generator's
// (fun_ofdm_b200/host/txgen.cpp), so that both produce the same stream: coded bits identical, samples to ~1e-15
// (tests/test_gpu_txgen.py).
//
// tx_frames_kernel: one CTA per frame, one warp per OFDM symbol.  Every coded bit is computed where it is needed from
// the 7 data bits it depends on (no sequential encoder), the 64-point inverse DFT runs in registers with warp shuffles,
// and each symbol is written once, coalesced, with its cyclic prefix.  tx_channel_kernel: one CTA per frame, in place.
#include "rx_internal.cuh"
#include "fft64.cuh"

#include <math.h>

#include "../../include/b200tx.h"

namespace b200rx {

namespace {

constexpr int TX_WARPS = 8;
constexpr int TX_THREADS = TX_WARPS * 32;
constexpr int TX_MAX_DATA_BYTES = 4160; // ceil(nsym * dbps / 8) <= 4131 for LENGTH <= 4095

__constant__ double2 c_tx_tw[64];        // exp(-2 pi i k / 64)
__constant__ double2 c_tx_preamble[320]; // preamble.h:24-359 (computed from 802.11a 17.3.3 + the table's two edge samples)
__constant__ int8_t c_tx_pol[127];       // pilot polarity
__constant__ uint8_t c_tx_scr[127];      // per-byte scrambler bit (ppdu.cpp:141-153), period 127
__constant__ uint32_t c_tx_crc[256];     // CRC-32/ISO-HDLC byte table

__device__ __forceinline__ uint32_t multmodp(uint32_t a, uint32_t b) // a(x) b(x) mod P(x), reflected (bit 31 = x^0)
{
    uint32_t p = 0;
#pragma unroll 4
    for (int i = 31; i >= 0; i--) {
        p ^= ((a >> i) & 1u) ? b : 0u;
        b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
    }
    return p;
}

// QAM<N>::encode for one axis (qam.h:87-99): recursive Gray mapping
__device__ __forceinline__ double qam_axis(uint32_t bits, int nbits, double scale)
{
    int pt = 0, flip = 1;
    for (int i = 0; i < nbits; i++) {
        const int bit = (int)((bits >> i) & 1u) * 2 - 1;
        pt = bit * flip + pt * 2;
        flip *= -bit;
    }
    return (double)pt * scale;
}

struct BitSource {
    const uint8_t *bytes; // MSB-first bit string (bits before the start are zero)
    int punc;
};

// coded bit `pc` (index after puncturing) of the stream: viterbi.cpp:48-60 + puncturer.cpp:41-63
__device__ __forceinline__ uint32_t coded_bit(const BitSource &src, uint32_t pc)
{
    uint32_t uc;
    if (src.punc == PUNC_1_2) uc = pc;
    else if (src.punc == PUNC_3_4) { const uint32_t g = pc >> 2, r = pc & 3u; uc = 6u * g + (r ? 2u * r - 1u : 0u); }
    else { const uint32_t g = pc / 3u, r = pc - 3u * g; uc = 4u * g + (r ? r + 1u : 0u); }
    const uint32_t i = uc >> 1; // input bit that produced it
    const uint32_t byte = i >> 3;
    const uint32_t w = ((byte ? (uint32_t)src.bytes[byte - 1] : 0u) << 8) | src.bytes[byte];
    const uint32_t sr = (w >> (7u - (i & 7u))) & 0x7Fu; // bit d = input bit i - d
    return (uint32_t)__popc(sr & ((uc & 1u) ? 91u : 121u)) & 1u;
}

struct TxArgs {
    const uint8_t *payloads;
    const uint64_t *payload_off;
    const uint32_t *lengths;
    const uint8_t *rates;
    uint32_t n_frames;
    double2 *iq_out;
    const uint64_t *out_off;
    double snr_db;
    uint32_t taps;
    uint32_t lead_in;
    uint64_t seed;
};

__global__ void __launch_bounds__(TX_THREADS) tx_frames_kernel(TxArgs a)
{
    __shared__ uint8_t s_data[TX_MAX_DATA_BYTES];
    __shared__ uint8_t s_hdr[4];
    __shared__ uint32_t s_crc[256];
    __shared__ double2 s_tw[64];
    __shared__ double2 s_bins[TX_WARPS][64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t f = blockIdx.x;
    const int rate = a.rates[f] <= 10 ? a.rates[f] : 10;
    const int len = (int)min(a.lengths[f], 4095u);
    const RateRow rr = rate_row(rate);
    const int nsym = (int)num_symbols(rate, (uint32_t)len);
    const int nbytes = nsym * rr.dbps / 8; // ppdu.cpp:118-119
    double2 *out = a.iq_out + a.out_off[f] + a.lead_in;

    for (int i = tid; i < 320; i += TX_THREADS) out[i] = c_tx_preamble[i]; // frame_builder.cpp:74-76
    if (tid < 64) s_tw[tid] = c_tx_tw[tid];
    s_crc[tid] = c_tx_crc[tid];
    // service (2 zero bytes) | payload | crc32 | zero pad (ppdu.cpp:121-137)
    const uint8_t *pl = a.payloads ? a.payloads + a.payload_off[f] : nullptr;
    for (int i = tid; i <= nbytes; i += TX_THREADS) s_data[i] = (i >= 2 && i < 2 + len) ? pl[i - 2] : 0;
    if (tid == 0) { // ppdu.cpp:84-95: rate(4) | reserved | length(12), parity in bit 17, six tail bits
        uint32_t field = ((uint32_t)(rr.rate_field & 0xF) << 13) | ((uint32_t)len & 0xFFFu);
        if (__popc(field) & 1) field |= 131072u;
        field <<= 6;
        s_hdr[0] = (uint8_t)(field >> 16); s_hdr[1] = (uint8_t)(field >> 8); s_hdr[2] = (uint8_t)field; s_hdr[3] = 0;
    }
    __syncthreads();

    // CRC-32 over service + payload: 32 right-aligned segments, one per lane, then crc(A||B) = crc(A) x^(8|B|) ^ crc(B)
    if (warp == 0) {
        const int n = 2 + len;
        const int L = (n + 31) >> 5;
        const int beg = max(0, n - (32 - lane) * L), end = max(0, n - (31 - lane) * L);
        uint32_t r = 0xFFFFFFFFu;
        for (int i = beg; i < end; i++) r = s_crc[(r ^ s_data[i]) & 0xFF] ^ (r >> 8);
        uint32_t crc = (end > beg) ? (r ^ 0xFFFFFFFFu) : 0u; // CRC of the empty string is 0
        uint32_t xp = 1u << 31, base = 1u << 23;             // x^0, x^8
        for (int e = L; e; e >>= 1) { if (e & 1) xp = multmodp(xp, base); base = multmodp(base, base); } // x^(8L)
#pragma unroll
        for (int lvl = 0; lvl < 5; lvl++) {
            const uint32_t right = __shfl_down_sync(FULL, crc, 1 << lvl);
            if ((lane & ((2 << lvl) - 1)) == 0) crc = multmodp(xp, crc) ^ right;
            xp = multmodp(xp, xp);
        }
        if (lane == 0) { // host byte order (ppdu.cpp:134-137)
            s_data[n] = (uint8_t)crc; s_data[n + 1] = (uint8_t)(crc >> 8);
            s_data[n + 2] = (uint8_t)(crc >> 16); s_data[n + 3] = (uint8_t)(crc >> 24);
        }
    }
    __syncthreads();
    for (int i = tid; i < nbytes; i += TX_THREADS) s_data[i] ^= c_tx_scr[i % 127]; // ppdu.cpp:141-153
    __syncthreads();

    double scale = 1.0; // QAM<N>(power) scale: sqrt(power * nn / sum2), qam.h:35-51
    const int nb = rr.bpsc == 1 ? 1 : rr.bpsc / 2;
    if (rr.bpsc > 1) {
        const int nn = 1 << (nb - 1), sum2 = (4 * nn * nn * nn - nn) / 3;
        scale = sqrt(0.5 * (double)nn / (double)sum2);
    }

    for (int v = warp; v <= nsym; v += TX_WARPS) { // v = 0: SIGNAL (BPSK 1/2, never scrambled)
        const BitSource src{v ? s_data : s_hdr, v ? (int)rr.punc : (int)PUNC_1_2};
        const int bpsc = v ? rr.bpsc : 1, nbv = v ? nb : 1;
        const double sc = v ? scale : 1.0;
        const uint32_t sym_base = v ? (uint32_t)(v - 1) * rr.cbps : 0u;
        double2 *bins = s_bins[warp];
        // nulls 0-5, 32, 59-63; pilots 11, 25, 39 = +polarity, 53 = -polarity (symbol_mapper.cpp:24-29, 97-115)
        const double pol = (double)c_tx_pol[v % 127];
        for (int s = lane; s < 64; s += 32) {
            double val = 0.0;
            if (s == 11 || s == 25 || s == 39) val = pol;
            else if (s == 53) val = -pol;
            if (s < 6 || s == 32 || s > 58 || val != 0.0) bins[s] = make_double2(val, 0.0);
        }
        for (int c = lane; c < 48; c += 32) {
            uint32_t bits = 0;
            for (int b = 0; b < bpsc; b++) {
                const int j = c * bpsc + b;             // interleaved position within the symbol
                const int blk = (j / 48) * 48, jj = j - blk;
                const int y = blk + 16 * (jj % 3) + jj / 3; // interleaver.cpp:21-24 inverted
                bits |= coded_bit(src, sym_base + (uint32_t)y) << b;
            }
            double2 pt;
            if (bpsc == 1) pt = make_double2(qam_axis(bits, 1, sc), 0.0); // modulator.cpp:43-50
            else pt = make_double2(qam_axis(bits, nbv, sc), qam_axis(bits >> nbv, nbv, sc));
            bins[data_bin(c)] = pt;
        }
        __syncwarp();
        // inverse DFT = conj(forward DFT of the conjugate) / 64; X[k] = bins[(k + 32) % 64] (fft.cpp:72-95)
        double2 v0 = bins[(lane + 32) & 63], v1 = bins[lane]; // X[lane], X[lane + 32]
        v0.y = -v0.y; v1.y = -v1.y;
        warp_fft64(v0, v1, s_tw, lane);
        __syncwarp();
        const int n0 = (int)(__brev((unsigned)(lane << 1)) >> 26), n1 = (int)(__brev((unsigned)((lane << 1) | 1)) >> 26);
        bins[n0] = make_double2(v0.x / 64.0, -v0.y / 64.0);
        bins[n1] = make_double2(v1.x / 64.0, -v1.y / 64.0);
        __syncwarp();
        double2 *o = out + 320 + 80 * (size_t)v; // cyclic prefix = last 16 samples (frame_builder.cpp:62-70)
        for (int k = lane; k < 80; k += 32) o[k] = bins[(k + 48) & 63];
        __syncwarp();
    }
}

// ---- channel: the host generator's (txgen.cpp), same counter-based Gaussian generator ----
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ double2 gauss_pair(uint64_t seed, uint64_t stream, uint64_t ctr)
{
    const uint64_t k = mix64(seed ^ mix64(stream * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull));
    const uint64_t x = mix64(k + 2 * ctr), y = mix64(k + 2 * ctr + 1);
    const double u1 = ((double)(x >> 11) + 1.0) * (1.0 / 9007199254740992.0); // (0, 1]
    const double u2 = (double)(y >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
    const double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincos(2.0 * M_PI * u2, &sn, &cs);
    return make_double2(r * cs, r * sn);
}

constexpr int CH_TILE = 2048;

__global__ void __launch_bounds__(TX_THREADS) tx_channel_kernel(TxArgs a)
{
    __shared__ double2 s_tile[CH_TILE + 8];
    __shared__ double2 s_taps[8];
    __shared__ double s_red[TX_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t f = blockIdx.x;
    const int rate = a.rates[f] <= 10 ? a.rates[f] : 10;
    const int len = (int)min(a.lengths[f], 4095u);
    const int ns = 320 + 80 * (1 + (int)num_symbols(rate, (uint32_t)len));
    double2 *base = a.iq_out + a.out_off[f];
    double2 *frame = base + a.lead_in;
    const int nt = a.taps > 8 ? 8 : (int)a.taps;

    if (nt > 1) { // random FIR, exponential power profile, unit energy; in place, last tile first
        if (tid == 0) {
            double e = 0.0;
            for (int k = 0; k < nt; k++) {
                double2 t = make_double2(1.0, 0.0);
                if (k) {
                    const double2 g = gauss_pair(a.seed, 0x7A9500000000ull + f, (uint64_t)k);
                    const double w = exp(-0.7 * k);
                    t = make_double2(g.x * w, g.y * w);
                }
                s_taps[k] = t;
                e += t.x * t.x + t.y * t.y;
            }
            const double inv = sqrt(e);
            for (int k = 0; k < nt; k++) s_taps[k] = make_double2(s_taps[k].x / inv, s_taps[k].y / inv);
        }
        for (int t0 = ((ns - 1) / CH_TILE) * CH_TILE; t0 >= 0; t0 -= CH_TILE) {
            const int cnt = min(CH_TILE, ns - t0);
            __syncthreads();
            for (int k = tid; k < cnt + 8; k += TX_THREADS) {
                const int n = t0 - 8 + k;
                s_tile[k] = n >= 0 ? frame[n] : make_double2(0.0, 0.0);
            }
            __syncthreads();
            for (int k = tid; k < cnt; k += TX_THREADS) {
                double2 acc = make_double2(0.0, 0.0);
                for (int j = 0; j < nt; j++) acc = cadd(acc, cmul(s_taps[j], s_tile[k + 8 - j]));
                frame[t0 + k] = acc;
            }
        }
        __syncthreads();
    }
    double sigma = 0.0;
    if (a.snr_db < 200.0) { // sigma from the mean power of the frame behind its preamble
        double p = 0.0;
        for (int n = 320 + tid; n < ns; n += TX_THREADS) { const double2 v = frame[n]; p += v.x * v.x + v.y * v.y; }
#pragma unroll
        for (int d = 16; d; d >>= 1) p += __shfl_xor_sync(FULL, p, d);
        if (lane == 0) s_red[warp] = p;
        __syncthreads();
        p = 0.0;
        for (int w = 0; w < TX_WARPS; w++) p += s_red[w];
        p /= (double)(ns - 320);
        sigma = sqrt(p / pow(10.0, a.snr_db / 10.0) / 2.0);
    }
    const int total = (int)a.lead_in + ns;
    for (int n = tid; n < total; n += TX_THREADS) {
        double2 v = n >= (int)a.lead_in ? base[n] : make_double2(0.0, 0.0);
        if (sigma > 0.0) {
            const double2 g = gauss_pair(a.seed, f, (uint64_t)n);
            v.x += sigma * g.x;
            v.y += sigma * g.y;
        }
        if (sigma > 0.0 || n < (int)a.lead_in) base[n] = v;
    }
}

bool g_tx_tables[64] = {false};

// 320 preamble samples from their definition (802.11a 17.3.3): ten short symbols, a 32-sample guard, two long symbols;
// plus the two window-edge samples the reference's table carries (preamble.h:26, :186)
void make_preamble(double2 *pre)
{
    double st[64][2], lt[64][2];
    const double a = sqrt(13.0 / 6.0);
    const int sk[12] = {-24, -20, -16, -12, -8, -4, 4, 8, 12, 16, 20, 24};
    const int ss[12] = {1, -1, 1, -1, -1, 1, -1, -1, 1, 1, 1, 1};
    const char *lts = "++--++-+-++++++--++-+-++++0+--++-+-+-----++--+-+-++++";
    for (int n = 0; n < 64; n++) {
        double sr = 0, si = 0, lr = 0, li = 0;
        for (int i = 0; i < 12; i++) {
            const int ph = ((sk[i] * n) % 64 + 64) % 64;
            const double c = cos(2.0 * M_PI * ph / 64.0), s = sin(2.0 * M_PI * ph / 64.0);
            sr += a * ss[i] * (c - s); // (1 + i) e^{i phi}
            si += a * ss[i] * (c + s);
        }
        for (int k = -26; k <= 26; k++) {
            const double l = lts[k + 26] == '+' ? 1.0 : (lts[k + 26] == '-' ? -1.0 : 0.0);
            const int ph = ((k * n) % 64 + 64) % 64;
            lr += l * cos(2.0 * M_PI * ph / 64.0);
            li += l * sin(2.0 * M_PI * ph / 64.0);
        }
        st[n][0] = sr / 64.0; st[n][1] = si / 64.0;
        lt[n][0] = lr / 64.0; lt[n][1] = li / 64.0;
    }
    for (int i = 0; i < 160; i++) pre[i] = make_double2(st[i & 15][0], st[i & 15][1]);
    for (int i = 0; i < 32; i++) pre[160 + i] = make_double2(lt[32 + i][0], lt[32 + i][1]);
    for (int i = 0; i < 64; i++) pre[192 + i] = pre[256 + i] = make_double2(lt[i][0], lt[i][1]);
    pre[0].x *= 0.5; pre[0].y *= 0.5;
    pre[160] = make_double2(-0.078, 0.0);
}

cudaError_t upload_txgen_tables()
{
    double2 tw[64], pre[320];
    for (int k = 0; k < 64; k++) { // exact octant symmetry, like the receive side's table
        const int q = k % 16, quad = k / 16;
        double c, s;
        if (q == 0) { c = 1.0; s = 0.0; }
        else if (q == 8) { c = s = sqrt(0.5); }
        else if (q < 8) { c = cos(2.0 * M_PI * q / 64.0); s = sin(2.0 * M_PI * q / 64.0); }
        else { c = sin(2.0 * M_PI * (16 - q) / 64.0); s = cos(2.0 * M_PI * (16 - q) / 64.0); }
        double re, im;
        switch (quad) {
            case 0: re = c; im = s; break;
            case 1: re = -s; im = c; break;
            case 2: re = -c; im = -s; break;
            default: re = s; im = -c; break;
        }
        tw[k] = make_double2(re, -im);
    }
    make_preamble(pre);
    int8_t pol[127];
    int st = 0x7F; // 802.11a 17.3.5.9: x^7 + x^4 + 1 from all ones, 0 -> +1, 1 -> -1
    for (int i = 0; i < 127; i++) {
        const int fb = ((st >> 6) ^ (st >> 3)) & 1;
        st = ((st << 1) | fb) & 0x7F;
        pol[i] = fb ? -1 : 1;
    }
    uint8_t scr[127];
    int state = 93; // ppdu.cpp:141-153
    for (int x = 0; x < 127; x++) {
        const int fb = ((state >> 6) & 1) ^ ((state >> 3) & 1);
        scr[x] = (uint8_t)fb;
        state = ((state << 1) & 0x7E) | fb;
    }
    uint32_t crc[256];
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
        crc[i] = c;
    }
    cudaError_t e = cudaMemcpyToSymbol(c_tx_tw, tw, sizeof(tw));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_tx_preamble, pre, sizeof(pre));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_tx_pol, pol, sizeof(pol));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_tx_scr, scr, sizeof(scr));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_tx_crc, crc, sizeof(crc));
    return e;
}

} // namespace

} // namespace b200rx

extern "C" __attribute__((visibility("default")))
int b200tx_build_batch_dev(int device, void *cuda_stream, const uint8_t *payloads_dev, const uint64_t *payload_off_dev,
                           const uint32_t *lengths_dev, const uint8_t *rates_dev, uint32_t n_frames, double *iq_out_dev,
                           const uint64_t *out_off_dev, const b200tx_channel *ch)
{
    using namespace b200rx;
    if (!payload_off_dev || !lengths_dev || !rates_dev || !iq_out_dev || !out_off_dev || !ch) return -1;
    if (device < 0 || device >= 64) return -1;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    if (!g_tx_tables[device]) {
        if (upload_txgen_tables() != cudaSuccess) return -2;
        g_tx_tables[device] = true;
    }
    if (n_frames == 0) return 0;
    TxArgs a{};
    a.payloads = payloads_dev; a.payload_off = payload_off_dev; a.lengths = lengths_dev; a.rates = rates_dev;
    a.n_frames = n_frames;
    a.iq_out = reinterpret_cast<double2 *>(iq_out_dev);
    a.out_off = out_off_dev;
    a.snr_db = ch->snr_db; a.taps = ch->multipath_taps; a.lead_in = ch->lead_in; a.seed = ch->seed;
    cudaStream_t s = (cudaStream_t)cuda_stream;
    tx_frames_kernel<<<n_frames, TX_THREADS, 0, s>>>(a);
    if (cudaGetLastError() != cudaSuccess) return -2;
    if (a.taps > 1 || a.snr_db < 200.0 || a.lead_in > 0) {
        tx_channel_kernel<<<n_frames, TX_THREADS, 0, s>>>(a);
        if (cudaGetLastError() != cudaSuccess) return -2;
    }
    return 0;
}
