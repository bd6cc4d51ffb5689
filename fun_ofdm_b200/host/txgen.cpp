// Synthetic-corpus generator (include/b200tx.h): frames as the reference's frame_builder emits them,
// then a seeded multipath + AWGN channel.  Host C++; used by bench.py to make the benchmark input
// and by the tests (pinned against the compiled reference).
//
// Reference behaviour restated here (never copied):
//   ppdu::encoder_header  src/ppdu.cpp:81-110    24-bit field, parity, K=7 r=1/2 code, interleave, BPSK
//   ppdu::encode_data     src/ppdu.cpp:112-165   service|payload|crc32 -> per-byte scrambler -> code ->
//                                                puncture -> interleave -> modulate
//   viterbi::conv_encode  src/viterbi.cpp:39-62  polys {121, 91}, tail NOT forced to zero
//   puncturer::puncture   src/puncturer.cpp:26-70
//   interleaver::interleave src/interleaver.cpp:15-26  (48-element permutation, always)
//   modulator::modulate   src/modulator.cpp:30-99, QAM<N>::encode src/qam.h:87-99
//   symbol_mapper::map    src/symbol_mapper.cpp:81-119
//   fft::inverse          src/fft.cpp:68-96
//   frame_builder::build_frame src/frame_builder.cpp:53-82
#include "../../include/b200tx.h"

#include <cmath>
#include <complex>
#include <cstring>
#include <thread>
#include <vector>

namespace {

typedef std::complex<double> cd;

struct RateRow { int rate_field, cbps, dbps, bpsc, punc; }; // punc: 0 = 1/2, 1 = 2/3, 2 = 3/4
const RateRow RATES[11] = {
    {0xD, 48, 24, 1, 0},  {0xE, 48, 32, 1, 1},  {0xF, 48, 36, 1, 2},
    {0x5, 96, 48, 2, 0},  {0x6, 96, 64, 2, 1},  {0x7, 96, 72, 2, 2},
    {0x9, 192, 96, 4, 0}, {0xA, 192, 128, 4, 1}, {0xB, 192, 144, 4, 2},
    {0x1, 288, 192, 6, 1}, {0x3, 288, 216, 6, 2},
};

inline int parity32(unsigned v) { return __builtin_parity(v); }

int num_symbols(int rate, int length)
{
    const int dbps = RATES[rate].dbps;
    return (16 + 8 * (length + 4) + 6 + dbps - 1) / dbps;
}

uint32_t crc32_iso_hdlc(const uint8_t *p, size_t n)
{
    static uint32_t tab[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
            tab[i] = c;
        }
        ready = true;
    }
    uint32_t r = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) r = tab[(r ^ p[i]) & 0xFF] ^ (r >> 8);
    return r ^ 0xFFFFFFFFu;
}

// K=7 encoder: shift register takes the new bit in the LSB, outputs parity(sr & 121), parity(sr & 91)
void conv_encode(const uint8_t *data, int nbits_in, std::vector<uint8_t> &out)
{
    out.resize((size_t)2 * nbits_in);
    unsigned sr = 0;
    for (int i = 0; i < nbits_in; i++) {
        const unsigned bit = (data[i >> 3] >> (7 - (i & 7))) & 1u;
        sr = (sr << 1) | bit;
        out[2 * i] = (uint8_t)parity32(sr & 121u);
        out[2 * i + 1] = (uint8_t)parity32(sr & 91u);
    }
}

void puncture(const std::vector<uint8_t> &in, int punc, std::vector<uint8_t> &out)
{
    out.clear();
    if (punc == 0) { out = in; return; }
    if (punc == 2) { // keep 0, 1, 3, 5 of every 6
        for (size_t x = 0; x + 5 < in.size() + 0; x += 6) { out.push_back(in[x]); out.push_back(in[x + 1]); out.push_back(in[x + 3]); out.push_back(in[x + 5]); }
    } else {         // keep 0, 2, 3 of every 4
        for (size_t x = 0; x + 3 < in.size() + 0; x += 4) { out.push_back(in[x]); out.push_back(in[x + 2]); out.push_back(in[x + 3]); }
    }
}

void interleave48(const std::vector<uint8_t> &in, std::vector<uint8_t> &out)
{
    out.resize(in.size());
    for (size_t x = 0; x < in.size(); x += 48)
        for (int y = 0; y < 48; y++) out[x + 3 * (y % 16) + y / 16] = in[x + y];
}

// QAM<N>::encode: recursive Gray mapping; scale sqrt(power * nn / sum2)
double qam_axis(const uint8_t *bits, int nbits, double scale)
{
    int pt = 0, flip = 1;
    for (int i = 0; i < nbits; i++) {
        const int bit = (int)bits[i] * 2 - 1;
        pt = bit * flip + pt * 2;
        flip *= -bit;
    }
    return pt * scale;
}

void modulate(const std::vector<uint8_t> &bits, int bpsc, std::vector<cd> &out)
{
    const size_t n = bits.size() / bpsc;
    out.resize(n);
    if (bpsc == 1) {
        for (size_t x = 0; x < n; x++) out[x] = cd(qam_axis(&bits[x], 1, 1.0), 0.0);
        return;
    }
    const int nb = bpsc / 2, nn = 1 << (nb - 1), sum2 = (4 * nn * nn * nn - nn) / 3;
    const double sf = std::sqrt(0.5 * double(nn) / double(sum2));
    for (size_t x = 0; x < n; x++)
        out[x] = cd(qam_axis(&bits[x * bpsc], nb, sf), qam_axis(&bits[x * bpsc + nb], nb, sf));
}

// pilot polarity: IEEE 802.11a 17.3.5.9 (scrambler x^7+x^4+1 from all ones; 0 -> +1, 1 -> -1)
struct Tables {
    double pol[127];
    cd tw[64];       // exp(+2 pi i k / 64)
    cd preamble[320];
    Tables()
    {
        int st = 0x7F;
        for (int i = 0; i < 127; i++) {
            const int fb = ((st >> 6) ^ (st >> 3)) & 1;
            st = ((st << 1) | fb) & 0x7F;
            pol[i] = fb ? -1.0 : 1.0;
        }
        for (int k = 0; k < 64; k++) {
            const int q = k % 16, quad = k / 16;
            double c, s;
            if (q == 0) { c = 1.0; s = 0.0; }
            else if (q == 8) { c = s = std::sqrt(0.5); }
            else if (q < 8) { c = std::cos(2.0 * M_PI * q / 64.0); s = std::sin(2.0 * M_PI * q / 64.0); }
            else { c = std::sin(2.0 * M_PI * (16 - q) / 64.0); s = std::cos(2.0 * M_PI * (16 - q) / 64.0); }
            switch (quad) {
                case 0: tw[k] = cd(c, s); break;
                case 1: tw[k] = cd(-s, c); break;
                case 2: tw[k] = cd(-c, -s); break;
                default: tw[k] = cd(s, -c); break;
            }
        }
        // Preamble from its definition (17.3.3): 10 short symbols (period 16), then a 32-sample guard
        // and two long symbols.  Index s <-> subcarrier s - 32.
        cd S[64], L[64], st_t[64], lt_t[64];
        for (int s = 0; s < 64; s++) { S[s] = 0; L[s] = 0; }
        const double a = std::sqrt(13.0 / 6.0);
        const int sk[12] = {-24, -20, -16, -12, -8, -4, 4, 8, 12, 16, 20, 24};
        const int ss[12] = {1, -1, 1, -1, -1, 1, -1, -1, 1, 1, 1, 1};
        for (int i = 0; i < 12; i++) S[sk[i] + 32] = cd(a * ss[i], a * ss[i]);
        const char *lts = "++--++-+-++++++--++-+-++++0+--++-+-+-----++--+-+-++++";
        for (int k = -26; k <= 26; k++) L[k + 32] = lts[k + 26] == '+' ? 1.0 : (lts[k + 26] == '-' ? -1.0 : 0.0);
        ifft64(S, st_t);
        ifft64(L, lt_t);
        for (int i = 0; i < 160; i++) preamble[i] = st_t[i & 15];
        for (int i = 0; i < 32; i++) preamble[160 + i] = lt_t[32 + i];
        for (int i = 0; i < 64; i++) preamble[192 + i] = preamble[256 + i] = lt_t[i];
        // window edges as carried by the reference's table (preamble.h:26 and :186): first sample halved,
        // first guard sample -0.078 (the table's 3-digit rounding of -0.078125)
        preamble[0] *= 0.5;
        preamble[160] = cd(-0.078, 0.0);
    }
    // x[n] = (1/64) sum_k X[k] exp(+2 pi i k n / 64) with X[k] = in[(k + 32) % 64]  (fft.cpp:72-95)
    void ifft64(const cd *in_shifted, cd *out) const
    {
        cd a[64];
        for (int k = 0; k < 64; k++) a[k] = in_shifted[(k + 32) & 63];
        // radix-2 decimation in time, bit-reversed input
        cd x[64];
        for (int k = 0; k < 64; k++) {
            int r = 0;
            for (int b = 0; b < 6; b++) if (k & (1 << b)) r |= 1 << (5 - b);
            x[k] = a[r];
        }
        for (int len = 2; len <= 64; len <<= 1) {
            const int h = len >> 1, step = 64 / len;
            for (int base = 0; base < 64; base += len)
                for (int j = 0; j < h; j++) {
                    const cd t = x[base + j + h] * tw[j * step];
                    x[base + j + h] = x[base + j] - t;
                    x[base + j] += t;
                }
        }
        for (int n = 0; n < 64; n++) out[n] = x[n] / 64.0;
    }
};

const Tables &tables()
{
    static const Tables t;
    return t;
}

// ppdu::encode -> constellation points, 48 per symbol, SIGNAL first
int ppdu_encode(const uint8_t *payload, int length, int rate, std::vector<cd> &pts)
{
    const RateRow &rp = RATES[rate];
    const int nsym = num_symbols(rate, length);
    std::vector<uint8_t> coded, punct, inter;
    std::vector<cd> mod;

    // header (ppdu.cpp:84-107)
    unsigned field = ((unsigned)(rp.rate_field & 0xF) << 13) | ((unsigned)length & 0xFFFu);
    if (parity32(field) == 1) field |= 131072u;
    field <<= 6;
    const uint8_t hb[4] = {(uint8_t)(field >> 16), (uint8_t)(field >> 8), (uint8_t)field, 0};
    conv_encode(hb, 24, coded); // 18 data bits + 6
    interleave48(coded, inter);
    modulate(inter, 1, mod);
    pts.assign(mod.begin(), mod.end());

    // data (ppdu.cpp:118-163)
    const int num_data_bits = nsym * rp.dbps, num_data_bytes = num_data_bits / 8;
    std::vector<uint8_t> data((size_t)num_data_bytes + 1, 0);
    if (length) memcpy(&data[2], payload, (size_t)length);
    const uint32_t crc = crc32_iso_hdlc(data.data(), 2 + (size_t)length);
    data[2 + length] = (uint8_t)crc; data[3 + length] = (uint8_t)(crc >> 8);
    data[4 + length] = (uint8_t)(crc >> 16); data[5 + length] = (uint8_t)(crc >> 24);
    int state = 93;
    for (int x = 0; x < num_data_bytes; x++) { // per-BYTE scrambler, flips bit 0 only
        const int fb = ((state >> 6) & 1) ^ ((state >> 3) & 1);
        data[x] ^= (uint8_t)fb;
        state = ((state << 1) & 0x7E) | fb;
    }
    conv_encode(data.data(), num_data_bits, coded); // (num_data_bits - 6) + 6 input bits
    puncture(coded, rp.punc, punct);
    interleave48(punct, inter);
    modulate(inter, rp.bpsc, mod);
    pts.insert(pts.end(), mod.begin(), mod.end());
    return nsym;
}

int build_frame(const uint8_t *payload, int length, int rate, cd *out)
{
    const Tables &T = tables();
    std::vector<cd> pts;
    const int nsym = ppdu_encode(payload, length, rate, pts);
    memcpy(out, T.preamble, sizeof(cd) * 320);
    cd bins[64], td[64];
    for (int v = 0; v <= nsym; v++) {
        // symbol_mapper.cpp:24-29,97-115: nulls 0-5, 32, 59-63; pilots 11, 25, 39, 53 = {1,1,1,-1} * polarity
        const double pol = T.pol[v % 127];
        int di = 0;
        for (int s = 0; s < 64; s++) {
            if (s < 6 || s == 32 || s > 58) bins[s] = 0.0;
            else if (s == 11 || s == 25 || s == 39) bins[s] = pol;
            else if (s == 53) bins[s] = -pol;
            else bins[s] = pts[(size_t)v * 48 + di++];
        }
        T.ifft64(bins, td);
        cd *o = out + 320 + 80 * (size_t)v;
        memcpy(o, td + 48, sizeof(cd) * 16); // cyclic prefix = last 16 samples
        memcpy(o + 16, td, sizeof(cd) * 64);
    }
    return 320 + 80 * (1 + nsym);
}

// ---- counter-based Gaussian generator: splitmix64 of (seed, stream, counter) + Box-Muller ----
inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

inline cd gauss_pair(uint64_t seed, uint64_t stream, uint64_t ctr)
{
    const uint64_t k = mix64(seed ^ mix64(stream * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull));
    const uint64_t a = mix64(k + 2 * ctr), b = mix64(k + 2 * ctr + 1);
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0); // (0, 1]
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
    const double r = std::sqrt(-2.0 * std::log(u1));
    return cd(r * std::cos(2.0 * M_PI * u2), r * std::sin(2.0 * M_PI * u2));
}

} // namespace

extern "C" {

int b200tx_num_symbols(int rate, int length)
{
    if (rate < 0 || rate > 10 || length < 0 || length > 4095) return -1;
    return num_symbols(rate, length);
}

int b200tx_frame_samples(int rate, int length)
{
    const int n = b200tx_num_symbols(rate, length);
    return n < 0 ? -1 : 320 + 80 * (1 + n);
}

int b200tx_build_frame(const uint8_t *payload, int length, int rate, double *iq_out)
{
    if (rate < 0 || rate > 10 || length < 0 || length > 4095 || !iq_out || (length && !payload)) return -1;
    return build_frame(payload, length, rate, reinterpret_cast<cd *>(iq_out));
}

int b200tx_ppdu_encode(const uint8_t *payload, int length, int rate, double *out)
{
    if (rate < 0 || rate > 10 || length < 0 || length > 4095 || !out || (length && !payload)) return -1;
    std::vector<cd> pts;
    ppdu_encode(payload, length, rate, pts);
    memcpy(out, pts.data(), pts.size() * sizeof(cd));
    return (int)pts.size();
}

void b200tx_preamble(double *iq_out_320) { memcpy(iq_out_320, tables().preamble, sizeof(cd) * 320); }

int b200tx_build_batch(const uint8_t *payloads, const uint64_t *payload_off, const uint32_t *lengths,
                       const uint8_t *rates, uint32_t n_frames, double *iq_out, const uint64_t *out_off,
                       const b200tx_channel *ch)
{
    if (!payload_off || !lengths || !rates || !iq_out || !out_off || !ch) return -1;
    for (uint32_t f = 0; f < n_frames; f++)
        if (rates[f] > 10 || lengths[f] > 4095) return -1;
    (void)tables();
    (void)crc32_iso_hdlc(nullptr, 0);
    unsigned nt = ch->n_threads ? ch->n_threads : 1;
    if (nt > n_frames) nt = n_frames ? n_frames : 1;
    auto work = [&](unsigned t) {
        std::vector<cd> frame, tmp;
        for (uint32_t f = t; f < n_frames; f += nt) {
            const int ns = 320 + 80 * (1 + num_symbols(rates[f], (int)lengths[f]));
            frame.resize(ns);
            build_frame(payloads ? payloads + payload_off[f] : nullptr, (int)lengths[f], rates[f], frame.data());
            if (ch->multipath_taps > 1) {
                const unsigned nt_ = ch->multipath_taps > 8 ? 8 : ch->multipath_taps;
                cd taps[8];
                double e = 0.0;
                for (unsigned k = 0; k < nt_; k++) {
                    taps[k] = (k == 0) ? cd(1.0, 0.0) : gauss_pair(ch->seed, 0x7A9500000000ull + f, k) * std::exp(-0.7 * k);
                    e += std::norm(taps[k]);
                }
                for (unsigned k = 0; k < nt_; k++) taps[k] /= std::sqrt(e);
                tmp = frame;
                for (int n = 0; n < ns; n++) {
                    cd acc = 0.0;
                    for (unsigned k = 0; k < nt_ && (int)k <= n; k++) acc += taps[k] * tmp[n - k];
                    frame[n] = acc;
                }
            }
            double sigma = 0.0;
            if (ch->snr_db < 200.0) {
                double p = 0.0;
                for (int n = 320; n < ns; n++) p += std::norm(frame[n]);
                p /= (double)(ns - 320);
                sigma = std::sqrt(p / std::pow(10.0, ch->snr_db / 10.0) / 2.0);
            }
            cd *o = reinterpret_cast<cd *>(iq_out) + out_off[f];
            const int total = (int)ch->lead_in + ns;
            for (int n = 0; n < total; n++) {
                cd v = (n >= (int)ch->lead_in) ? frame[n - ch->lead_in] : cd(0.0, 0.0);
                if (sigma > 0.0) v += sigma * gauss_pair(ch->seed, f, (uint64_t)n);
                o[n] = v;
            }
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    return 0;
}

} // extern "C"
