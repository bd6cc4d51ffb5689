/* TEST INFRASTRUCTURE — not product code.
 *
 * C-ABI harness around the UNMODIFIED reference sources (compiled in place from
 * /root/reference/src by oracle/Makefile into oracle/_ref/libfunref.so).  Nothing
 * here re-implements the reference: every function instantiates the reference's
 * own classes (fun::frame_builder, fun::fft_symbols, fun::channel_est,
 * fun::phase_tracker, fun::frame_decoder, fun::ppdu, fun::viterbi, ...) and calls
 * their public members.  Used by tests/ (parity checker, golden-vector
 * generator) and by bench.py's cpu_baseline / --impl reference legs only.
 *
 * The reference needs FFTW3, Boost.CRC and Boost.DateTime, none of which exist
 * in this image; oracle/shims/ provides API-compatible stand-ins (see the
 * header comment of each shim for what arithmetic they implement).
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <complex>
#include <cstdint>
#include <cstring>
#include <thread>
#include <time.h>
#include <vector>

#include "block.h"
#include "tagged_vector.h"
#include "rates.h"
#include "parity.h"
#include "viterbi.h"
#include "interleaver.h"
#include "puncturer.h"
#include "modulator.h"
#include "ppdu.h"
#include "fft.h"
#include "frame_builder.h"
#include "frame_detector.h"
#include "timing_sync.h"
#include "fft_symbols.h"
#include "channel_est.h"
#include "phase_tracker.h"
#include "frame_decoder.h"
#include "receiver_chain.h"
#include "preamble.h"
#include "symbol_mapper.h"

#include <boost/crc.hpp>

using namespace fun;

typedef std::complex<double> cd;

namespace {

struct hot_path {
    fft_symbols ffts;
    channel_est chest;
    phase_tracker phase;
    frame_decoder dec;

    /* One round through the four hot-path blocks, chained by hand in the order
     * receiver_chain.cpp:33-36 adds them (no threads, no one-round delay). */
    void run(std::vector<tagged_sample> &in, std::vector<tagged_vector<48> > *capture)
    {
        ffts.input_buffer.swap(in);
        ffts.work();
        chest.input_buffer.swap(ffts.output_buffer);
        chest.work();
        phase.input_buffer.swap(chest.output_buffer);
        phase.work();
        if (capture) capture->insert(capture->end(), phase.output_buffer.begin(), phase.output_buffer.end());
        dec.input_buffer.swap(phase.output_buffer);
        dec.work();
    }
};

void make_genie_stream(const double *iq, int n, std::vector<tagged_sample> &out)
{
    out.resize(n);
    for (int i = 0; i < n; i++) {
        out[i].sample = cd(iq[2 * i], iq[2 * i + 1]);
        out[i].tag = NONE;
    }
    if (n > 0) out[0].tag = LTS1;
    if (n > 64) out[64].tag = LTS2;
}

int nsym_for(int rate, int length)
{
    RateParams rp = RateParams((Rate)rate);
    return (int)std::ceil(double(16 + 8 * (length + 4) + 6) / double(rp.dbps));
}

} // namespace

extern "C" {

typedef struct {
    int32_t hdr_ok;      /* ppdu::decode_header() result */
    int32_t hdr_field;   /* 24-bit decoded SIGNAL field */
    int32_t hdr_parity;  /* fun::parity(field) (1 = reject) */
    int32_t rate_valid;  /* rate field is in VALID_RATES */
    int32_t rate;        /* fun::Rate enum, -1 if header failed */
    int32_t length;
    int32_t nsym;
    int32_t crc_ok;      /* ppdu::decode_data() result (0 if not attempted) */
    int32_t n_vectors;   /* 48-carrier vectors produced by phase_tracker */
    int32_t payload_from_blocks; /* frame_decoder pushed a payload (must equal crc_ok) */
} ref_frame_info;

void ref_init(void) { (void)parity(0); }

int ref_rate_params(int rate, int *cbps, int *dbps, int *bpsc, int *rate_field)
{
    if (rate < 0 || rate > 10) return -1;
    RateParams rp = RateParams((Rate)rate);
    *cbps = rp.cbps; *dbps = rp.dbps; *bpsc = rp.bpsc; *rate_field = rp.rate_field;
    return 0;
}

int ref_num_symbols(int rate, int length) { return nsym_for(rate, length); }

int ref_sizeof(int which)
{
    switch (which) {
        case 0: return (int)sizeof(tagged_sample);
        case 1: return (int)sizeof(tagged_vector<64>);
        case 2: return (int)sizeof(tagged_vector<48>);
    }
    return -1;
}

/* ---- TX (frame_builder.cpp:53-82) ---- */
int ref_build_frame(const uint8_t *payload, int len, int rate, double *iq_out, int max_samples)
{
    frame_builder fb;
    std::vector<unsigned char> p(payload, payload + len);
    std::vector<cd> s = fb.build_frame(p, (Rate)rate);
    if ((int)s.size() > max_samples) return -(int)s.size();
    memcpy(iq_out, s.data(), s.size() * sizeof(cd));
    return (int)s.size();
}

/* ppdu::encode(): 48*(1+nsym) constellation points before symbol mapping */
int ref_ppdu_encode(const uint8_t *payload, int len, int rate, double *out, int max_samples)
{
    std::vector<unsigned char> p(payload, payload + len);
    ppdu f(p, (Rate)rate);
    std::vector<cd> s = f.encode();
    if ((int)s.size() > max_samples) return -(int)s.size();
    memcpy(out, s.data(), s.size() * sizeof(cd));
    return (int)s.size();
}

/* ---- codec stages ---- */
void ref_conv_encode(const uint8_t *data, uint8_t *symbols, int data_bits)
{
    viterbi v;
    v.conv_encode(const_cast<uint8_t *>(data), symbols, data_bits);
}

void ref_conv_decode(const uint8_t *symbols, uint8_t *data, int data_bits)
{
    viterbi v;
    v.conv_decode(const_cast<uint8_t *>(symbols), data, data_bits);
}

int ref_puncture(const uint8_t *in, int n, int rate, uint8_t *out)
{
    std::vector<unsigned char> d(in, in + n);
    std::vector<unsigned char> o = puncturer::puncture(d, RateParams((Rate)rate));
    memcpy(out, o.data(), o.size());
    return (int)o.size();
}

int ref_depuncture(const uint8_t *in, int n, int rate, uint8_t *out)
{
    std::vector<unsigned char> d(in, in + n);
    std::vector<unsigned char> o = puncturer::depuncture(d, RateParams((Rate)rate));
    memcpy(out, o.data(), o.size());
    return (int)o.size();
}

int ref_interleave(const uint8_t *in, int n, uint8_t *out)
{
    std::vector<unsigned char> d(in, in + n);
    std::vector<unsigned char> o = interleaver::interleave(d);
    memcpy(out, o.data(), o.size());
    return (int)o.size();
}

int ref_deinterleave(const uint8_t *in, int n, uint8_t *out)
{
    std::vector<unsigned char> d(in, in + n);
    std::vector<unsigned char> o = interleaver::deinterleave(d);
    memcpy(out, o.data(), o.size());
    return (int)o.size();
}

int ref_modulate(const uint8_t *bits, int n, int rate, double *out)
{
    std::vector<unsigned char> d(bits, bits + n);
    std::vector<cd> o = modulator::modulate(d, (Rate)rate);
    memcpy(out, o.data(), o.size() * sizeof(cd));
    return (int)o.size();
}

int ref_demodulate(const double *iq, int nsamp, int rate, uint8_t *out)
{
    std::vector<cd> d(nsamp);
    memcpy(d.data(), iq, nsamp * sizeof(cd));
    std::vector<unsigned char> o = modulator::demodulate(d, (Rate)rate);
    memcpy(out, o.data(), o.size());
    return (int)o.size();
}

void ref_fft_forward(double *iq64)
{
    static thread_local fft *f = new fft(64);
    f->forward(reinterpret_cast<cd *>(iq64));
}

int ref_fft_inverse(double *iq, int n)
{
    static thread_local fft *f = new fft(64);
    std::vector<cd> d(n);
    memcpy(d.data(), iq, n * sizeof(cd));
    f->inverse(d);
    memcpy(iq, d.data(), n * sizeof(cd));
    return n;
}

uint32_t ref_crc32(const uint8_t *data, int n)
{
    boost::crc_32_type crc;
    crc.process_bytes(data, n);
    return crc.checksum();
}

int ref_parity(int x) { return parity(x); }

/* Constant tables of the reference, dumped so tests can pin the product's own (computed) tables.
 * which: 0 PREAMBLE_SAMPLES[320], 1 LTS_FREQ_DOMAIN[64], 2 LTS_TIME_DOMAIN_CONJ[64] (preamble.h) */
int ref_table(int which, double *out)
{
    const cd *t = NULL; int n = 0;
    switch (which) {
        case 0: t = PREAMBLE_SAMPLES; n = 320; break;
        case 1: t = LTS_FREQ_DOMAIN; n = 64; break;
        case 2: t = LTS_TIME_DOMAIN_CONJ; n = 64; break;
        default: return -1;
    }
    memcpy(out, t, n * sizeof(cd));
    return n;
}

int ref_decode_header(const double *iq48, int *rate, int *length, int *nsym)
{
    std::vector<cd> s(48);
    memcpy(s.data(), iq48, 48 * sizeof(cd));
    ppdu h;
    if (!h.decode_header(s)) return 0;
    *rate = (int)h.get_rate(); *length = h.get_length(); *nsym = h.get_num_symbols();
    return 1;
}

int ref_decode_data(const double *iq, int nsamp, int rate, int length, uint8_t *payload_out)
{
    std::vector<cd> s(nsamp);
    memcpy(s.data(), iq, nsamp * sizeof(cd));
    ppdu f((Rate)rate, length);
    if (!f.decode_data(s)) return 0;
    std::vector<unsigned char> p = f.get_payload();
    if (!p.empty()) memcpy(payload_out, p.data(), p.size());
    return 1;
}

/* ---- one frame through the four hot-path blocks, with every intermediate ----
 * iq: n_samples complex doubles starting AT the LTS1-tagged sample (genie tags:
 * LTS1 at 0, LTS2 at 64; SURVEY.md section 4, last bullet).  Any output pointer
 * may be NULL.  Stage dumps are produced by calling the reference's own stage
 * functions in the order ppdu::decode_data uses them (ppdu.cpp:238-264). */
int ref_decode_frame(const double *iq, int n_samples, ref_frame_info *info,
                     double *eq, int eq_cap_vectors,
                     uint8_t *soft, uint8_t *deint, uint8_t *depunct,
                     uint8_t *decoded, uint8_t *descrambled, uint8_t *payload)
{
    memset(info, 0, sizeof(*info));
    info->rate = -1;

    std::vector<tagged_sample> stream;
    make_genie_stream(iq, n_samples, stream);
    hot_path hp;
    std::vector<tagged_vector<48> > vec;
    hp.run(stream, &vec);
    info->n_vectors = (int)vec.size();
    if (eq) {
        int nv = std::min((int)vec.size(), eq_cap_vectors);
        for (int v = 0; v < nv; v++) memcpy(eq + (size_t)v * 96, vec[v].samples, 48 * sizeof(cd));
    }
    if (!hp.dec.output_buffer.empty()) {
        info->payload_from_blocks = 1;
        const std::vector<unsigned char> &p = hp.dec.output_buffer[0];
        if (payload && !p.empty()) memcpy(payload, p.data(), p.size());
    }
    if (vec.empty() || vec[0].tag != START_OF_FRAME) return 0;

    /* header internals via the reference's public stage functions (ppdu.cpp:173-203) */
    std::vector<cd> hs(vec[0].samples, vec[0].samples + 48);
    {
        std::vector<unsigned char> dm = modulator::demodulate(hs, RATE_1_2_BPSK);
        std::vector<unsigned char> di = interleaver::deinterleave(dm);
        unsigned char hb[4] = {0, 0, 0, 0};
        viterbi v;
        v.conv_decode(di.data(), hb, 18);
        unsigned int field = ((unsigned)hb[0] << 16) | ((unsigned)hb[1] << 8) | hb[2];
        info->hdr_field = (int)field;
        info->hdr_parity = parity((int)field);
        unsigned char rf = (field >> 19) & 0xF;
        for (size_t x = 0; x < VALID_RATES.size(); x++) if (VALID_RATES[x] == rf) info->rate_valid = 1;
    }
    ppdu h;
    if (!h.decode_header(hs)) return 0;
    info->hdr_ok = 1;
    info->rate = (int)h.get_rate();
    info->length = h.get_length();
    info->nsym = h.get_num_symbols();
    if ((int)vec.size() < 1 + info->nsym) return 0; /* truncated */

    RateParams rp = RateParams(h.get_rate());
    std::vector<cd> ds((size_t)info->nsym * 48);
    for (int s = 0; s < info->nsym; s++) memcpy(&ds[(size_t)s * 48], vec[1 + s].samples, 48 * sizeof(cd));
    std::vector<unsigned char> dm = modulator::demodulate(ds, h.get_rate());
    std::vector<unsigned char> di = interleaver::deinterleave(dm);
    std::vector<unsigned char> dp = puncturer::depuncture(di, rp);
    int num_data_bits = info->nsym * rp.dbps;
    int num_data_bytes = num_data_bits / 8;
    std::vector<unsigned char> dec(num_data_bytes + 1, 0);
    viterbi v;
    v.conv_decode(dp.data(), dec.data(), num_data_bits - 6);
    std::vector<unsigned char> ds2(num_data_bytes + 1, 0);
    int state = 93, feedback = 0;
    for (int x = 0; x < num_data_bytes; x++) { /* ppdu.cpp:257-263, executed here only to dump it */
        feedback = (!!(state & 64)) ^ (!!(state & 8));
        ds2[x] = feedback ^ dec[x];
        state = ((state << 1) & 0x7E) | feedback;
    }
    if (soft) memcpy(soft, dm.data(), dm.size());
    if (deint) memcpy(deint, di.data(), di.size());
    if (depunct) memcpy(depunct, dp.data(), dp.size());
    if (decoded) memcpy(decoded, dec.data(), num_data_bytes);
    if (descrambled) memcpy(descrambled, ds2.data(), num_data_bytes);

    ppdu f(h.get_rate(), info->length);
    std::vector<unsigned char> pl(4096);
    info->crc_ok = f.decode_data(ds) ? 1 : 0;
    return 1;
}

/* ---- batch of genie-tagged frames through the hot-path blocks, N threads ----
 * Each thread owns private fft_symbols/channel_est/phase_tracker/frame_decoder
 * instances and a contiguous range of frames.
 * The tagged_sample streams (the blocks' input format) are built before the
 * clock starts; the timed region is the four work() calls per frame only.
 * status: 0 = payload produced, 255 = no payload.  Returns wall seconds. */
double ref_decode_batch(const double *iq, const int64_t *lts1_off, const int32_t *n_avail, int n_frames,
                        uint8_t *payload_out, int payload_stride, int32_t *len_out, uint8_t *status_out,
                        int n_threads)
{
    ref_init();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_frames) n_threads = n_frames > 0 ? n_frames : 1;
    std::vector<std::vector<tagged_sample> > streams(n_frames);
    auto range = [&](int t, int &a, int &b) {
        a = (int)((long long)n_frames * t / n_threads);
        b = (int)((long long)n_frames * (t + 1) / n_threads);
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++)
            th.emplace_back([&, t]() {
                int a, b; range(t, a, b);
                for (int f = a; f < b; f++) make_genie_stream(iq + 2 * lts1_off[f], n_avail[f], streams[f]);
            });
        for (auto &t : th) t.join();
    }
    auto worker = [&](int t) {
        int a, b; range(t, a, b);
        hot_path *hp = new hot_path();
        for (int f = a; f < b; f++) {
            hp->run(streams[f], nullptr);
            if (!hp->dec.output_buffer.empty()) {
                const std::vector<unsigned char> &p = hp->dec.output_buffer[0];
                int n = std::min((int)p.size(), payload_stride);
                if (payload_out && n) memcpy(payload_out + (size_t)f * payload_stride, p.data(), n);
                if (len_out) len_out[f] = (int)p.size();
                if (status_out) status_out[f] = 0;
            } else {
                if (len_out) len_out[f] = 0;
                if (status_out) status_out[f] = 255;
                /* a frame that did not complete may leave block state behind: start clean */
                delete hp;
                hp = new hot_path();
            }
        }
        delete hp;
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(worker, t);
    worker(0);
    for (auto &t : th) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* ---- frame_detector + timing_sync over a raw stream (producer of the boundary) ----
 * Output stream is delayed 160 samples (timing_sync.cpp:58-66,130-137); n_out = n. */
long ref_sync(const double *iq, long n, int chunk, double *iq_out, uint8_t *tags_out)
{
    frame_detector fd;
    timing_sync ts;
    long produced = 0;
    for (long x = 0; x < n; x += chunk) {
        long e = std::min(n, x + (long)chunk);
        fd.input_buffer.assign(reinterpret_cast<const cd *>(iq) + x, reinterpret_cast<const cd *>(iq) + e);
        fd.work();
        ts.input_buffer.swap(fd.output_buffer);
        ts.work();
        for (size_t i = 0; i < ts.output_buffer.size(); i++) {
            iq_out[2 * produced] = ts.output_buffer[i].sample.real();
            iq_out[2 * produced + 1] = ts.output_buffer[i].sample.imag();
            tags_out[produced] = (uint8_t)ts.output_buffer[i].tag;
            produced++;
        }
    }
    return produced;
}

/* Same, with the work() calls cut at arbitrary stream positions: call i gets samples [bounds[i], bounds[i+1]).
 * flush_zeros > 0 appends one more call of that many zero samples, which pushes the last 160 samples out of
 * timing_sync's carry-over.  Returns the number of output samples (n + flush_zeros). */
long ref_sync_bounds(const double *iq, long n, const long *bounds, int n_bounds, int flush_zeros, double *iq_out,
                     uint8_t *tags_out)
{
    frame_detector fd;
    timing_sync ts;
    long produced = 0;
    std::vector<cd> zeros((size_t)(flush_zeros > 0 ? flush_zeros : 0));
    for (int i = 0; i + 1 < n_bounds + (flush_zeros > 0 ? 1 : 0); i++) {
        if (i + 1 < n_bounds) {
            long a = std::max(0L, std::min(n, bounds[i])), e = std::max(a, std::min(n, bounds[i + 1]));
            if (e == a) continue;
            fd.input_buffer.assign(reinterpret_cast<const cd *>(iq) + a, reinterpret_cast<const cd *>(iq) + e);
        } else {
            fd.input_buffer.assign(zeros.begin(), zeros.end());
        }
        fd.work();
        ts.input_buffer.swap(fd.output_buffer);
        ts.work();
        for (size_t k = 0; k < ts.output_buffer.size(); k++) {
            iq_out[2 * produced] = ts.output_buffer[k].sample.real();
            iq_out[2 * produced + 1] = ts.output_buffer[k].sample.imag();
            tags_out[produced] = (uint8_t)ts.output_buffer[k].tag;
            produced++;
        }
    }
    return produced;
}

/* ---- tagged stream through the four hot-path blocks in chunks (streaming semantics) ---- */
int ref_hotpath_stream(const double *iq, const uint8_t *tags, long n, int chunk,
                       uint8_t *payload_out, int payload_stride, int32_t *len_out, int max_frames)
{
    hot_path hp;
    int count = 0;
    std::vector<tagged_sample> in;
    for (long x = 0; x < n; x += chunk) {
        long e = std::min(n, x + (long)chunk);
        in.resize(e - x);
        for (long i = x; i < e; i++) {
            in[i - x].sample = cd(iq[2 * i], iq[2 * i + 1]);
            in[i - x].tag = (vector_tag)tags[i];
        }
        hp.run(in, nullptr);
        for (size_t k = 0; k < hp.dec.output_buffer.size(); k++) {
            if (count < max_frames) {
                const std::vector<unsigned char> &p = hp.dec.output_buffer[k];
                int m = std::min((int)p.size(), payload_stride);
                if (m) memcpy(payload_out + (size_t)count * payload_stride, p.data(), m);
                len_out[count] = (int)p.size();
            }
            count++;
        }
    }
    return count;
}

/* ---- the full 6-thread receiver_chain (BASELINE config #1) ---- */
void *ref_chain_new(void) { ref_init(); return new receiver_chain(); } /* never freed: its threads never join */

int ref_chain_process(void *chain, const double *iq, int n, uint8_t *payload_out, int payload_stride,
                      int32_t *len_out, int max_frames)
{
    receiver_chain *rc = static_cast<receiver_chain *>(chain);
    std::vector<cd> s(reinterpret_cast<const cd *>(iq), reinterpret_cast<const cd *>(iq) + n);
    std::vector<std::vector<unsigned char> > out = rc->process_samples(s);
    int count = 0;
    for (size_t k = 0; k < out.size(); k++) {
        if (count < max_frames) {
            int m = std::min((int)out[k].size(), payload_stride);
            if (m) memcpy(payload_out + (size_t)count * payload_stride, out[k].data(), m);
            len_out[count] = (int)out[k].size();
        }
        count++;
    }
    return count;
}

/* The feed loop of examples/test_sim.cpp:77-97 in native code: chunks of `chunk` samples built as a std::vector and
 * passed by value (test_sim.cpp:84-87), then `drain` zero chunks of 4096 samples (untimed) for the frames still inside the
 * six-stage pipeline.  seconds = the feed loop only.  Returns the number of payloads; the first max_frames are copied. */
int ref_chain_run(void *chain, const double *iq, long n, long chunk, int drain, uint8_t *payload_out, int payload_stride,
                  int32_t *len_out, int max_frames, double *seconds)
{
    receiver_chain *rc = static_cast<receiver_chain *>(chain);
    const cd *x = reinterpret_cast<const cd *>(iq);
    int count = 0;
    auto take = [&](const std::vector<std::vector<unsigned char> > &out) {
        for (size_t k = 0; k < out.size(); k++, count++) {
            if (count >= max_frames) continue;
            int m = std::min((int)out[k].size(), payload_stride);
            if (m) memcpy(payload_out + (size_t)count * payload_stride, out[k].data(), m);
            len_out[count] = (int)out[k].size();
        }
    };
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (long s = 0; s < n; s += chunk) {
        std::vector<cd> v(x + s, x + std::min(n, s + chunk));
        take(rc->process_samples(v));
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (int i = 0; i < drain; i++) take(rc->process_samples(std::vector<cd>(4096)));
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    return count;
}

} /* extern "C" */
