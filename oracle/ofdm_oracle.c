/* TEST INFRASTRUCTURE — not product code.  See ofdm_oracle.h for scope and pinning status.
 *
 * Every function cites the reference file:line (relative to /root/reference/) it restates.
 * Written as straight scalar C: no SIMD, no tables copied from the reference (constant
 * tables are regenerated from their defining rule and pinned by tests against the
 * reference's own tables / behaviour).
 */
#include "ofdm_oracle.h"

#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef double complex cd;

/* ------------------------------------------------------------------------------------------
 * Rates — src/rates.h:21 (VALID_RATES), :31-44 (enum), :52-196 (RateParams)
 * ------------------------------------------------------------------------------------------ */
enum { PUNC_1_2 = 0, PUNC_2_3 = 1, PUNC_3_4 = 2 };

typedef struct { int rate_field, cbps, dbps, bpsc, punc; } rate_params;

static const rate_params RATE_TABLE[11] = {
    {0xD, 48, 24, 1, PUNC_1_2},  {0xE, 48, 32, 1, PUNC_2_3},  {0xF, 48, 36, 1, PUNC_3_4},
    {0x5, 96, 48, 2, PUNC_1_2},  {0x6, 96, 64, 2, PUNC_2_3},  {0x7, 96, 72, 2, PUNC_3_4},
    {0x9, 192, 96, 4, PUNC_1_2}, {0xA, 192, 128, 4, PUNC_2_3}, {0xB, 192, 144, 4, PUNC_3_4},
    {0x1, 288, 192, 6, PUNC_2_3}, {0x3, 288, 216, 6, PUNC_3_4},
};

/* rates.h:208-249 FromRateField; -1 when the field is not in VALID_RATES (rates.h:21) */
static int rate_from_field(int field)
{
    for (int r = 0; r < 11; r++) if (RATE_TABLE[r].rate_field == field) return r;
    return -1;
}

/* ppdu.cpp:38-40, 207-209: ceil((16 service + 8*(len + 4 crc) + 6 tail) / dbps) */
static int num_symbols(int rate, int length)
{
    int bits = 16 + 8 * (length + 4) + 6, dbps = RATE_TABLE[rate].dbps;
    return (bits + dbps - 1) / dbps;
}

/* ------------------------------------------------------------------------------------------
 * parity — src/parity.h:43-48, parity.cpp:21-35 (odd parity of the low 32 bits, folded to a byte)
 * ------------------------------------------------------------------------------------------ */
int orc_parity(int x)
{
    unsigned v = (unsigned)x;
    v ^= v >> 16;
    v ^= v >> 8;
    v &= 0xFF;
    int cnt = 0;
    while (v) { cnt += v & 1; v >>= 1; }
    return cnt & 1;
}

/* ------------------------------------------------------------------------------------------
 * CRC-32 — boost::crc_32_type as used at src/ppdu.cpp:134-136, 267-269.  Third-party
 * (Boost >= 1.59, CMakeLists.txt:71, not vendored): published CRC-32/ISO-HDLC, reflected
 * poly 0xEDB88320, init and xorout 0xFFFFFFFF.  Bitwise on purpose (independent of the shim).
 * ------------------------------------------------------------------------------------------ */
uint32_t orc_crc32(const uint8_t *data, int n)
{
    uint32_t r = 0xFFFFFFFFu;
    for (int i = 0; i < n; i++) {
        r ^= data[i];
        for (int k = 0; k < 8; k++) r = (r & 1u) ? (r >> 1) ^ 0xEDB88320u : (r >> 1);
    }
    return r ^ 0xFFFFFFFFu;
}

/* ------------------------------------------------------------------------------------------
 * Convolutional code — src/viterbi.h:23-29 (K=7, polys {121, 91}), viterbi.cpp:39-62 (encode)
 * ------------------------------------------------------------------------------------------ */
void orc_conv_encode(const uint8_t *data, uint8_t *symbols, int data_bits)
{
    /* viterbi.cpp:48-60: data_bits + 6 input bits are READ FROM data (tail not forced to zero),
     * MSB first within each byte; shift register takes the new bit in its LSB. */
    int sr = 0, idx = 0;
    for (int i = 0; i < data_bits + 6; i++) {
        int bit = (data[i / 8] >> (7 - (i % 8))) & 1;
        sr = (sr << 1) | bit;
        symbols[idx++] = (uint8_t)orc_parity(sr & 121);
        symbols[idx++] = (uint8_t)orc_parity(sr & 91);
    }
}

/* Viterbi decoder — scalar restatement of the Spiral SSE2 kernel.
 *   viterbi.cpp:31-37   conv_decode = alloc + init + decode
 *   viterbi.cpp:71-78   init: all metrics 63, state 0 -> 0
 *   viterbi.cpp:87-91   Branchtab[i*32 + s] = parity(2s & poly_i) ? 255 : 0
 *   viterbi.cpp:190-197 decisions zeroed for all nbits steps
 *   viterbi.cpp:208-459 FULL_SPIRAL: two trellis steps per loop iteration, nbits/2 iterations
 *   viterbi.cpp:108-146 chainback from state 0
 */
void orc_conv_decode(const uint8_t *syms, uint8_t *data, int data_bits)
{
    const int T = data_bits + 6;
    uint8_t B0[32], B1[32];
    for (int s = 0; s < 32; s++) {
        B0[s] = orc_parity((2 * s) & 121) ? 255 : 0;
        B1[s] = orc_parity((2 * s) & 91) ? 255 : 0;
    }
    uint64_t *dec = (uint64_t *)calloc((size_t)(T > 0 ? T : 1), sizeof(uint64_t));
    uint8_t X[64], Y[64];
    for (int i = 0; i < 64; i++) X[i] = 63;
    X[0] = 0;

    const int steps = 2 * (T / 2); /* loop bound i9 <= nbits/2 - 1, two steps per pass (:209) */
    for (int t = 0; t < steps; t++) {
        const int s0 = syms[2 * t], s1 = syms[2 * t + 1];
        uint64_t d = 0;
        for (int j = 0; j < 32; j++) {
            /* :234-248  _mm_avg_epu8 (rounds up), >>2 within the byte, mask 63; 63 - m by psubusb */
            const int m = (((s0 ^ B0[j]) + (s1 ^ B1[j]) + 1) >> 1) >> 2;
            const int mi = 63 - m;
            /* :252-255  paddusb: saturate at 255 */
            int a = X[j] + m;       if (a > 255) a = 255;
            int b = X[j + 32] + mi; if (b > 255) b = 255;
            int c = X[j] + mi;      if (c > 255) c = 255;
            int e = X[j + 32] + m;  if (e > 255) e = 255;
            /* :256-273  min, decision = (min == metric through predecessor j+32): ties -> 1.
             * unpacklo/hi interleave puts new state 2j at bit 2j, 2j+1 at bit 2j+1 */
            if (b <= a) { d |= 1ull << (2 * j); Y[2 * j] = (uint8_t)b; } else Y[2 * j] = (uint8_t)a;
            if (e <= c) { d |= 1ull << (2 * j + 1); Y[2 * j + 1] = (uint8_t)e; } else Y[2 * j + 1] = (uint8_t)c;
        }
        dec[t] = d;
        /* :314-332, :438-456  renormalise only when the metric of state 0 exceeds 210 */
        if (Y[0] > 210) {
            uint8_t mn = 255;
            for (int i = 0; i < 64; i++) if (Y[i] < mn) mn = Y[i];
            for (int i = 0; i < 64; i++) Y[i] = (uint8_t)(Y[i] - mn);
        }
        memcpy(X, Y, 64);
    }

    /* :131-142  endstate kept in the top 6 bits of a byte (ADDSHIFT = 2); decisions looked up
     * past the 6 tail steps; every step rewrites data[n >> 3] with the current byte register */
    unsigned e = 0;
    for (int n = data_bits - 1; n >= 0; n--) {
        unsigned k = (unsigned)((dec[n + 6] >> (e >> 2)) & 1u);
        e = (e >> 1) | (k << 7);
        data[n >> 3] = (uint8_t)e;
    }
    free(dec);
}

/* ------------------------------------------------------------------------------------------
 * Puncturing — src/puncturer.cpp:26-70 (puncture), :78-123 (depuncture, erasure value 127)
 * ------------------------------------------------------------------------------------------ */
int orc_puncture(const uint8_t *in, int n, int rate, uint8_t *out)
{
    int o = 0;
    switch (RATE_TABLE[rate].punc) {
        case PUNC_1_2: memcpy(out, in, (size_t)n); return n;                 /* :33-36 */
        case PUNC_3_4:                                                        /* :39-52 keep 0,1,3,5 of 6 */
            for (int x = 0; x < n; x += 6) { out[o++] = in[x]; out[o++] = in[x + 1]; out[o++] = in[x + 3]; out[o++] = in[x + 5]; }
            return o;
        default:                                                              /* :55-67 keep 0,2,3 of 4 */
            for (int x = 0; x < n; x += 4) { out[o++] = in[x]; out[o++] = in[x + 2]; out[o++] = in[x + 3]; }
            return o;
    }
}

int orc_depuncture(const uint8_t *in, int n, int rate, uint8_t *out)
{
    int o = 0;
    switch (RATE_TABLE[rate].punc) {
        case PUNC_1_2: memcpy(out, in, (size_t)n); return n;                 /* :85-88 */
        case PUNC_3_4:                                                        /* :91-105 */
            for (int x = 0; x < n; x += 4) {
                out[o++] = in[x]; out[o++] = in[x + 1]; out[o++] = 127;
                out[o++] = in[x + 2]; out[o++] = 127; out[o++] = in[x + 3];
            }
            return o;
        default:                                                              /* :108-121 */
            for (int x = 0; x < n; x += 3) { out[o++] = in[x]; out[o++] = 127; out[o++] = in[x + 1]; out[o++] = in[x + 2]; }
            return o;
    }
}

/* ------------------------------------------------------------------------------------------
 * Interleaver — src/interleaver.cpp:15-38; BitInterleave(48, 1) ALWAYS (interleaver.cpp:18,31),
 * so interleaver.h:66-75 reduces to index(k) = 3*(k % 16) + k / 16  (s = 1, j = i).
 * ------------------------------------------------------------------------------------------ */
static int ilv_index(int k) { return 3 * (k % 16) + k / 16; }

int orc_interleave(const uint8_t *in, int n, uint8_t *out)
{
    /* interleaver.cpp:21-24: out[x + map[y]] = in[x + y], map = forward index */
    for (int x = 0; x < n; x += 48)
        for (int y = 0; y < 48 && x + y < n; y++) out[x + ilv_index(y)] = in[x + y];
    return n;
}

int orc_deinterleave(const uint8_t *in, int n, uint8_t *out)
{
    /* interleaver.cpp:34-37 with the inverse map (interleaver.h:88-92: v[index(i)] = i):
     * out[s + inv[t]] = in[s + t]  <=>  out[s + k] = in[s + index(k)] */
    for (int s = 0; s < n; s += 48)
        for (int k = 0; k < 48 && s + k < n; k++) out[s + k] = in[s + ilv_index(k)];
    return n;
}

/* ------------------------------------------------------------------------------------------
 * QAM — src/qam.h:35-51 (scales), :87-99 (encode), :110-125 (decode);
 * modulation order per rate src/modulator.cpp:30-99 (modulate), :108-164 (demodulate)
 * ------------------------------------------------------------------------------------------ */
typedef struct { int nbits; int gain; double scale_e, scale_d; } qam_t;

static qam_t qam_make(int nbits, double power)
{
    qam_t q;
    q.nbits = nbits;
    q.gain = 8 - nbits;                                  /* qam.h:37 d_gain = gain(0) + CHAR_BIT - NumBits */
    int nn = 1 << (nbits - 1);
    int sum2 = (4 * nn * nn * nn - nn) / 3;              /* qam.h:46 */
    double sf = sqrt(power * (double)nn / (double)sum2); /* qam.h:48 */
    q.scale_e = sf;
    q.scale_d = (double)(1 << q.gain) / sf;              /* qam.h:50 */
    return q;
}

static double qam_encode(const qam_t *q, const uint8_t *bits)
{
    int pt = 0, flip = 1;
    for (int i = 0; i < q->nbits; i++) { /* qam.h:91-97; bits are read as (signed) char 0/1 */
        int bit = (int)(signed char)bits[i] * 2 - 1;
        pt = bit * flip + pt * 2;
        flip *= -bit;
    }
    return pt * q->scale_e;
}

static void qam_decode(const qam_t *q, double sym, uint8_t *bits)
{
    int pt = (int)(sym * q->scale_d); /* qam.h:112: C conversion, truncates toward zero */
    int flip = 1;
    int amp = (1 << (q->nbits - 1)) << q->gain;
    for (int i = 0; i < q->nbits; i++) { /* qam.h:116-124 */
        int v = flip * pt + 128;
        bits[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        int sgn = pt < 0 ? -1 : 1;       /* qam.h:62-65: +1 | (v >> 31) */
        pt -= sgn * amp;
        flip = -sgn;
        amp /= 2;
    }
}

static qam_t qam_for_rate(int rate)
{
    switch (RATE_TABLE[rate].bpsc) {
        case 1: return qam_make(1, 1.0);  /* modulator.cpp:37,120 QAM<1> bpsk(1.0) */
        case 2: return qam_make(1, 0.5);  /* :50,129 QAM<1> qpsk(0.5): one instance per axis */
        case 4: return qam_make(2, 0.5);  /* :65,141 */
        default: return qam_make(3, 0.5); /* :80,153 */
    }
}

int orc_modulate(const uint8_t *bits, int n, int rate, double *out)
{
    qam_t q = qam_for_rate(rate);
    int bpsc = RATE_TABLE[rate].bpsc, ns = n / bpsc;
    for (int x = 0; x < ns; x++) {
        if (bpsc == 1) { out[2 * x] = qam_encode(&q, bits + x); out[2 * x + 1] = 0.0; }
        else {
            out[2 * x] = qam_encode(&q, bits + x * bpsc);
            out[2 * x + 1] = qam_encode(&q, bits + x * bpsc + bpsc / 2);
        }
    }
    return ns;
}

int orc_demodulate(const double *iq, int nsamp, int rate, uint8_t *out)
{
    qam_t q = qam_for_rate(rate);
    int bpsc = RATE_TABLE[rate].bpsc;
    for (int s = 0; s < nsamp; s++) {
        if (bpsc == 1) qam_decode(&q, iq[2 * s], out + s);            /* modulator.cpp:121-122 real only */
        else {
            qam_decode(&q, iq[2 * s], out + s * bpsc);                 /* real -> first half of the bits */
            qam_decode(&q, iq[2 * s + 1], out + s * bpsc + bpsc / 2);  /* imag -> second half */
        }
    }
    return nsamp * bpsc;
}

/* ------------------------------------------------------------------------------------------
 * FFT — src/fft.cpp:50-59 + fft_map :20-24.  FFTW3 is third-party (fftw3 >= 3.0, unpinned,
 * cmake/Modules/FindFFTW3.cmake:5): restated from its published definition, the unnormalised
 * forward DFT  X[k] = sum_n x[n] exp(-2 pi i k n / 64), then stored shifted so that
 * data[s] = X[(s + 32) mod 64].  Direct O(N^2) summation with an exact-symmetry twiddle table.
 * ------------------------------------------------------------------------------------------ */
static double TW_RE[64], TW_IM[64];
static int tables_ready = 0;

static void tables_init(void)
{
    if (tables_ready) return;
    for (int k = 0; k < 64; k++) {
        /* octant reduction keeps cos/sin arguments in [0, pi/4] so the table is exactly symmetric */
        int q = k % 16, quad = k / 16;
        double c, s;
        if (q <= 8) { c = cos(2.0 * M_PI * q / 64.0); s = sin(2.0 * M_PI * q / 64.0); }
        else { c = sin(2.0 * M_PI * (16 - q) / 64.0); s = cos(2.0 * M_PI * (16 - q) / 64.0); }
        if (q == 0) { c = 1.0; s = 0.0; }
        if (q == 8) { c = s = sqrt(0.5); }
        double re, im; /* exp(+i theta), theta = 2 pi k / 64 */
        switch (quad) {
            case 0: re = c; im = s; break;
            case 1: re = -s; im = c; break;
            case 2: re = -c; im = -s; break;
            default: re = s; im = -c; break;
        }
        TW_RE[k] = re; TW_IM[k] = im;
    }
    tables_ready = 1;
}

void orc_fft_forward(double *iq64)
{
    tables_init();
    double outr[64], outi[64];
    for (int s = 0; s < 64; s++) {
        int k = (s + 32) & 63;
        double ar = 0.0, ai = 0.0;
        for (int n = 0; n < 64; n++) {
            int idx = (k * n) & 63;
            double wr = TW_RE[idx], wi = -TW_IM[idx]; /* exp(-i ...) */
            ar += iq64[2 * n] * wr - iq64[2 * n + 1] * wi;
            ai += iq64[2 * n] * wi + iq64[2 * n + 1] * wr;
        }
        outr[s] = ar; outi[s] = ai;
    }
    for (int s = 0; s < 64; s++) { iq64[2 * s] = outr[s]; iq64[2 * s + 1] = outi[s]; }
}

/* ------------------------------------------------------------------------------------------
 * Constant sequences, regenerated from IEEE 802.11a-1999 rules and pinned by tests:
 *   LTS_FREQ_DOMAIN  src/preamble.h:363-429  (index s <-> subcarrier s - 32; L_{-26..26} of 17.3.3)
 *   POLARITY[127]    src/phase_tracker.cpp:23-32 (17.3.5.9: scrambler x^7 + x^4 + 1 seeded all-ones;
 *                    output bit 0 -> +1, 1 -> -1)
 *   pilots           src/phase_tracker.cpp:37-43: bins 11, 25, 39, 53 with base signs +,+,+,-
 *   data carriers    src/phase_tracker.cpp:46-50: bins 6..58 except 11, 25, 32, 39, 53
 * ------------------------------------------------------------------------------------------ */
static double lts_freq(int s)
{
    /* bit (k + 26) set => L_k = -1, for k = -26..26 (L_0 = 0) */
    static const char *L = "++--++-+-++++++--++-+-++++0+--++-+-+-----++--+-+-++++";
    int k = s - 32;
    if (k < -26 || k > 26) return 0.0;
    char c = L[k + 26];
    return c == '+' ? 1.0 : (c == '-' ? -1.0 : 0.0);
}

static int polarity(int n)
{
    static int seq[127];
    static int ready = 0;
    if (!ready) {
        int st = 0x7F;
        for (int i = 0; i < 127; i++) {
            int fb = ((st >> 6) ^ (st >> 3)) & 1;
            st = ((st << 1) | fb) & 0x7F;
            seq[i] = fb ? -1 : 1;
        }
        ready = 1;
    }
    return seq[n % 127];
}

static const int PILOT_BIN[4] = {11, 25, 39, 53};
static const int PILOT_SIGN[4] = {1, 1, 1, -1};

static int data_bin(int s)
{
    /* 48 data carriers in ascending bin order */
    int bin = 6;
    for (int i = 0;; bin++) {
        if (bin == 11 || bin == 25 || bin == 32 || bin == 39 || bin == 53) continue;
        if (i == s) return bin;
        i++;
    }
}

/* ------------------------------------------------------------------------------------------
 * Header + payload codec — src/ppdu.cpp:168-218 (decode_header), :223-295 (decode_data)
 * ------------------------------------------------------------------------------------------ */
static int decode_header_detail(const double *iq48, int *field_out, int *parity_out, int *rate_valid,
                                int *rate, int *length, int *nsym)
{
    uint8_t dm[48], di[48], hb[4] = {0, 0, 0, 0};
    orc_demodulate(iq48, 48, 0, dm);          /* :173 BPSK */
    orc_deinterleave(dm, 48, di);             /* :176 */
    orc_conv_decode(di, hb, 18);              /* :181 18 data bits -> 24 trellis steps */
    /* :184-186 three bytes, MSB first */
    unsigned field = ((unsigned)hb[0] << 16) | ((unsigned)hb[1] << 8) | hb[2];
    int par = orc_parity((int)field);         /* :187 */
    int rf = (field >> 19) & 0xF;             /* :194 */
    int len = (field >> 6) & 0xFFF;           /* :195 */
    int r = rate_from_field(rf);              /* :198-203 */
    if (field_out) *field_out = (int)field;
    if (parity_out) *parity_out = par;
    if (rate_valid) *rate_valid = r >= 0;
    if (par == 1) return 1;                   /* HDR_PARITY */
    if (r < 0) return 2;                      /* HDR_RATE */
    *rate = r; *length = len; *nsym = num_symbols(r, len);
    return 0;
}

int orc_decode_header(const double *iq48, int *rate, int *length, int *nsym)
{
    return decode_header_detail(iq48, NULL, NULL, NULL, rate, length, nsym) == 0;
}

/* Returns 1 when the CRC matches.  Stage dumps optional (NULL to skip). */
static int decode_data_detail(const double *iq, int rate, int length, uint8_t *payload,
                              uint8_t *soft, uint8_t *deint, uint8_t *depunct, uint8_t *decoded, uint8_t *descr)
{
    const rate_params *rp = &RATE_TABLE[rate];
    int nsym = num_symbols(rate, length);          /* :229-231 */
    int num_data_bits = nsym * rp->dbps;           /* :234 */
    int num_data_bytes = num_data_bits / 8;        /* :235 */
    int ncoded = nsym * rp->cbps;

    uint8_t *dm = (uint8_t *)malloc((size_t)ncoded);
    uint8_t *di = (uint8_t *)malloc((size_t)ncoded);
    uint8_t *dp = (uint8_t *)malloc((size_t)num_data_bits * 2 + 8);
    uint8_t *dec = (uint8_t *)calloc((size_t)num_data_bytes + 8, 1);
    uint8_t *ds = (uint8_t *)calloc((size_t)num_data_bytes + 8, 1);

    orc_demodulate(iq, nsym * 48, rate, dm);       /* :238 */
    orc_deinterleave(dm, ncoded, di);              /* :241 */
    orc_depuncture(di, ncoded, rate, dp);          /* :244 */
    orc_conv_decode(dp, dec, num_data_bits - 6);   /* :247-253 data_bits = num_data_bits - 6 */

    /* :255-264 descrambler: LFSR state 93, stepped once per BYTE, flips bit 0 of the byte */
    int state = 93;
    for (int x = 0; x < num_data_bytes; x++) {
        int fb = ((state >> 6) & 1) ^ ((state >> 3) & 1);
        ds[x] = (uint8_t)(fb ^ dec[x]);
        state = ((state << 1) & 0x7E) | fb;
    }

    /* :267-279 CRC over service(2) + payload, compared with the little-endian word that follows */
    uint32_t calc = orc_crc32(ds, 2 + length);
    uint32_t given = (uint32_t)ds[2 + length] | ((uint32_t)ds[3 + length] << 8) |
                     ((uint32_t)ds[4 + length] << 16) | ((uint32_t)ds[5 + length] << 24);
    int ok = calc == given;
    if (ok && payload && length) memcpy(payload, ds + 2, (size_t)length); /* :284-285 */

    if (soft) memcpy(soft, dm, (size_t)ncoded);
    if (deint) memcpy(deint, di, (size_t)ncoded);
    if (depunct) memcpy(depunct, dp, (size_t)num_data_bits * 2);
    if (decoded) memcpy(decoded, dec, (size_t)num_data_bytes);
    if (descr) memcpy(descr, ds, (size_t)num_data_bytes);
    free(dm); free(di); free(dp); free(dec); free(ds);
    return ok;
}

int orc_decode_data(const double *iq, int nsamp, int rate, int length, uint8_t *payload_out)
{
    (void)nsamp;
    return decode_data_detail(iq, rate, length, payload_out, NULL, NULL, NULL, NULL, NULL);
}

/* ------------------------------------------------------------------------------------------
 * Front end for one genie-tagged frame: LTS1 tag on sample 0, LTS2 tag on sample 64.
 *   fft_symbols.cpp:39-71  windows [0,64) [64,128) then [128 + 80 s + 16, +64) for s = 0, 1, ...
 *                          (a vector is emitted only when all 64 slots are filled)
 *   channel_est.cpp:44-65  Hinv[j] = sum over the two LTS vectors of L[j] / R[j] / 2   (all 64 j)
 *   channel_est.cpp:77-81  out[j] = Hinv[j] * in[j]
 *   phase_tracker.cpp:77-102  n = 0 at SIGNAL; e = sum_p in[bin_p] * conj(sign_p * POLARITY[n % 127]) / 4;
 *                          theta = arg(e); out[s] = in[data_bin(s)] * (cos(-theta) + i sin(-theta)); n++
 * Returns the number of equalised 48-carrier vectors written to eq (SIGNAL first).
 * ------------------------------------------------------------------------------------------ */
static int front_end(const double *iq, int n_samples, double *eq /* [cap][48][2] */, int cap)
{
    if (n_samples < 128) return 0;
    cd hinv[64];
    double buf[128];
    for (int j = 0; j < 64; j++) hinv[j] = 0.0;
    for (int l = 0; l < 2; l++) {
        memcpy(buf, iq + 2 * 64 * l, sizeof(buf));
        orc_fft_forward(buf);
        for (int j = 0; j < 64; j++) {
            cd ref = lts_freq(j) + 0.0 * I;
            cd rec = buf[2 * j] + buf[2 * j + 1] * I;
            hinv[j] += ref / rec / 2.0; /* channel_est.cpp:57 — same operator order */
        }
    }
    int nvec = (n_samples - 128) / 80;
    if (nvec > cap) nvec = cap;
    for (int v = 0; v < nvec; v++) {
        memcpy(buf, iq + 2 * (128 + 80 * v + 16), sizeof(buf));
        orc_fft_forward(buf);
        cd sym[64];
        for (int j = 0; j < 64; j++) sym[j] = hinv[j] * (buf[2 * j] + buf[2 * j + 1] * I);
        cd err = 0.0;
        for (int p = 0; p < 4; p++) {
            int pilot = PILOT_SIGN[p] * polarity(v);
            cd refp = (double)pilot + 0.0 * I;
            err += sym[PILOT_BIN[p]] * conj(refp) / 4.0;
        }
        double angle = carg(err);
        cd rot = cos(-angle) + sin(-angle) * I;
        for (int s = 0; s < 48; s++) {
            cd o = sym[data_bin(s)] * rot;
            eq[((size_t)v * 48 + s) * 2] = creal(o);
            eq[((size_t)v * 48 + s) * 2 + 1] = cimag(o);
        }
    }
    return nvec;
}

/* frame_decoder.cpp:45-91 for an isolated frame: header from the first vector, then nsym vectors.
 * Status: 0 OK, 1 HDR_PARITY, 2 HDR_RATE, 3 CRC_FAIL, 4 TRUNCATED (not enough samples). */
static int decode_one(const double *iq, int n_samples, orc_frame_info *info, double *eq_out, int eq_cap,
                      uint8_t *soft, uint8_t *deint, uint8_t *depunct, uint8_t *decoded, uint8_t *descr,
                      uint8_t *payload)
{
    orc_frame_info local;
    if (!info) info = &local;
    memset(info, 0, sizeof(*info));
    info->rate = -1;
    int cap = n_samples >= 128 ? (n_samples - 128) / 80 : 0;
    double *eq = (double *)malloc(sizeof(double) * 96 * (size_t)(cap > 0 ? cap : 1));
    int nvec = front_end(iq, n_samples, eq, cap);
    info->n_vectors = nvec;
    if (eq_out) memcpy(eq_out, eq, sizeof(double) * 96 * (size_t)(nvec < eq_cap ? nvec : eq_cap));
    int status = 4;
    if (nvec >= 1) {
        int rate = -1, length = 0, nsym = 0;
        int h = decode_header_detail(eq, &info->hdr_field, &info->hdr_parity, &info->rate_valid, &rate, &length, &nsym);
        if (h != 0) status = h;
        else {
            info->hdr_ok = 1; info->rate = rate; info->length = length; info->nsym = nsym;
            if (nvec >= 1 + nsym) {
                int ok = decode_data_detail(eq + 96, rate, length, payload, soft, deint, depunct, decoded, descr);
                info->crc_ok = ok;
                info->payload_from_blocks = ok;
                status = ok ? 0 : 3;
            }
        }
    }
    free(eq);
    return status;
}

int orc_decode_frame(const double *iq, int n_samples, orc_frame_info *info,
                     double *eq, int eq_cap_vectors,
                     uint8_t *soft, uint8_t *deint, uint8_t *depunct,
                     uint8_t *decoded, uint8_t *descrambled, uint8_t *payload)
{
    decode_one(iq, n_samples, info, eq, eq_cap_vectors, soft, deint, depunct, decoded, descrambled, payload);
    return info->hdr_ok;
}

typedef struct {
    const double *iq; const int64_t *off; const int32_t *avail; int f0, f1;
    uint8_t *payload; int stride; int32_t *len; uint8_t *status;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    uint8_t *tmp = (uint8_t *)malloc(4096 + 8);
    for (int f = j->f0; f < j->f1; f++) {
        orc_frame_info info;
        int st = decode_one(j->iq + 2 * j->off[f], j->avail[f], &info, NULL, 0, NULL, NULL, NULL, NULL, NULL, tmp);
        if (j->status) j->status[f] = (uint8_t)st;
        if (j->len) j->len[f] = info.hdr_ok ? info.length : 0;
        if (st == 0 && j->payload) {
            int n = info.length < j->stride ? info.length : j->stride;
            memcpy(j->payload + (size_t)f * j->stride, tmp, (size_t)n);
        }
    }
    free(tmp);
    return NULL;
}

double orc_decode_batch(const double *iq, const int64_t *lts1_off, const int32_t *n_avail, int n_frames,
                        uint8_t *payload_out, int payload_stride, int32_t *len_out, uint8_t *status_out,
                        int n_threads)
{
    orc_init();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t th[256];
    batch_job jobs[256];
    for (int t = 0; t < n_threads; t++) {
        batch_job j = {iq, lts1_off, n_avail, (int)((long long)n_frames * t / n_threads),
                       (int)((long long)n_frames * (t + 1) / n_threads), payload_out, payload_stride, len_out, status_out};
        jobs[t] = j;
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void orc_init(void)
{
    tables_init();
    (void)polarity(0);
}
