// C entry points that drive fun::b200_rx from a test harness (ctypes): feed a tagged stream in chunks the
// way receiver_chain::process_samples does (receiver_chain.cpp:106-126), collect the payloads.
#include "b200_rx.h"
#include "b200_receiver_chain.h"
#include "b200_receiver.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>

#define API extern "C" __attribute__((visibility("default")))

API void *b200host_rx_block_new(int device, unsigned max_frames, unsigned max_payload)
{
    fun::b200_rx *b = new fun::b200_rx(device, max_frames, max_payload);
    if (!b->ok()) { delete b; return nullptr; }
    return b;
}

API void b200host_rx_block_delete(void *blk) { delete static_cast<fun::b200_rx *>(blk); }

// One work() round.  Returns the number of payloads produced; copies them (up to max_out) to payload_out
// (stride bytes each) / len_out.  flush != 0 also decodes what is complete and drops partial frames.
API int b200host_rx_block_work(void *blk, const double *iq, const uint8_t *tags, long n, int flush,
                               uint8_t *payload_out, int stride, int32_t *len_out, int max_out)
{
    fun::b200_rx *b = static_cast<fun::b200_rx *>(blk);
    b->input_buffer.resize(n);
    for (long i = 0; i < n; i++) {
        b->input_buffer[i].sample = std::complex<double>(iq[2 * i], iq[2 * i + 1]);
        b->input_buffer[i].tag = (fun::vector_tag)tags[i];
    }
    b->work();
    std::vector<std::vector<unsigned char> > out;
    out.swap(b->output_buffer);
    if (flush) {
        b->flush();
        out.insert(out.end(), b->output_buffer.begin(), b->output_buffer.end());
    }
    int count = 0;
    for (size_t k = 0; k < out.size(); k++, count++) {
        if (count >= max_out) continue;
        int m = (int)out[k].size() < stride ? (int)out[k].size() : stride;
        if (m) std::memcpy(payload_out + (size_t)count * stride, out[k].data(), m);
        len_out[count] = (int32_t)out[k].size();
    }
    return count;
}

API void *b200host_rx_block_new2(int device, unsigned max_frames, unsigned max_payload, unsigned depth, unsigned max_lag)
{
    fun::b200_rx *b = new fun::b200_rx(device, max_frames, max_payload, depth, max_lag);
    if (!b->ok()) { delete b; return nullptr; }
    return b;
}

// The block inside a chain, in native code: the tagged stream (iq, tags)[0, n) is cut into rounds of `chunk` samples,
// each round's std::vector<tagged_sample> is built beforehand (in the reference it is timing_sync's output_buffer, swapped
// into this block's input_buffer by receiver_chain.cpp:119 - no copy), then the timed loop swaps it in and calls work().
// seconds[0] = the rounds, seconds[1] = rounds + flush.  Returns the number of payloads.
API int b200host_rx_block_run(void *blk, const double *iq, const uint8_t *tags, long n, long chunk, uint8_t *payload_out,
                              int stride, int32_t *len_out, int max_out, double *seconds)
{
    fun::b200_rx *b = static_cast<fun::b200_rx *>(blk);
    std::vector<std::vector<fun::tagged_sample> > rounds;
    for (long s = 0; s < n; s += chunk) {
        const long len = n - s < chunk ? n - s : chunk;
        rounds.push_back(std::vector<fun::tagged_sample>((size_t)len));
        for (long i = 0; i < len; i++) {
            rounds.back()[i].sample = std::complex<double>(iq[2 * (s + i)], iq[2 * (s + i) + 1]);
            rounds.back()[i].tag = (fun::vector_tag)tags[s + i];
        }
    }
    int count = 0;
    auto take = [&](const std::vector<std::vector<unsigned char> > &out) {
        for (size_t k = 0; k < out.size(); k++, count++) {
            if (count >= max_out) continue;
            int m = (int)out[k].size() < stride ? (int)out[k].size() : stride;
            if (m) std::memcpy(payload_out + (size_t)count * stride, out[k].data(), m);
            len_out[count] = (int32_t)out[k].size();
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t r = 0; r < rounds.size(); r++) {
        b->input_buffer.swap(rounds[r]);
        b->work();
        take(b->output_buffer);
    }
    const auto t1 = std::chrono::steady_clock::now();
    b->output_buffer.resize(0);
    b->flush();
    take(b->output_buffer);
    const auto t2 = std::chrono::steady_clock::now();
    seconds[0] = std::chrono::duration<double>(t1 - t0).count();
    seconds[1] = std::chrono::duration<double>(t2 - t0).count();
    return count;
}

API void b200host_rx_block_counters(void *blk, uint64_t *out5)
{
    fun::b200_rx::counters_t c = static_cast<fun::b200_rx *>(blk)->counters();
    out5[0] = c.frames_seen; out5[1] = c.headers_bad; out5[2] = c.frames_ok; out5[3] = c.frames_crc_fail; out5[4] = c.frames_abandoned;
}

// ---- fun::b200_receiver_chain (raw samples in, payloads out: receiver_chain::process_samples) ----
API void *b200host_chain_new(int device, unsigned max_frames, unsigned max_payload)
{
    fun::b200_receiver_chain *c = new fun::b200_receiver_chain(device, max_frames, max_payload);
    if (!c->ok()) { delete c; return nullptr; }
    return c;
}

API void *b200host_chain_new2(int device, unsigned max_frames, unsigned max_payload, unsigned depth, unsigned max_lag)
{
    fun::b200_receiver_chain *c = new fun::b200_receiver_chain(device, max_frames, max_payload, depth, max_lag);
    if (!c->ok()) { delete c; return nullptr; }
    return c;
}

API int b200host_chain_set_tuning(void *chain, const char *key, long long value)
{
    return static_cast<fun::b200_receiver_chain *>(chain)->set_tuning(key, value);
}

API void b200host_chain_set_copy_threads(void *chain, unsigned n) { static_cast<fun::b200_receiver_chain *>(chain)->set_copy_threads(n); }

API double *b200host_alloc_samples(long n) { return reinterpret_cast<double *>(fun::b200_receiver_chain::alloc_samples((size_t)n)); }
API void b200host_free_samples(double *p) { fun::b200_receiver_chain::free_samples(reinterpret_cast<std::complex<double> *>(p)); }

// The feed loop of examples/test_sim.cpp:77-97 in native code (a Python loop would add its own per-call cost): the stream
// iq[0, n) goes through process_samples in chunks of `chunk` samples, then flush().  by_value != 0 builds a
// std::vector per chunk and passes it by value exactly like test_sim.cpp:84-87; 0 uses the pointer overload.
// seconds[0] = the feed loop, seconds[1] = feed loop + flush.  Returns the number of payloads (all of them are counted,
// the first max_out are copied out).
API int b200host_chain_run(void *chain, const double *iq, long n, long chunk, int by_value, uint8_t *payload_out, int stride,
                           int32_t *len_out, int max_out, double *seconds)
{
    fun::b200_receiver_chain *c = static_cast<fun::b200_receiver_chain *>(chain);
    const std::complex<double> *x = reinterpret_cast<const std::complex<double> *>(iq);
    int count = 0;
    auto take = [&](const std::vector<std::vector<unsigned char> > &out) {
        for (size_t k = 0; k < out.size(); k++, count++) {
            if (count >= max_out) continue;
            int m = (int)out[k].size() < stride ? (int)out[k].size() : stride;
            if (m) std::memcpy(payload_out + (size_t)count * stride, out[k].data(), m);
            len_out[count] = (int32_t)out[k].size();
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    for (long s = 0; s < n; s += chunk) {
        const long len = n - s < chunk ? n - s : chunk;
        if (by_value) {
            std::vector<std::complex<double> > v(x + s, x + s + len);
            take(c->process_samples(v));
        } else {
            take(c->process_samples(x + s, (size_t)len));
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    take(c->flush());
    const auto t2 = std::chrono::steady_clock::now();
    seconds[0] = std::chrono::duration<double>(t1 - t0).count();
    seconds[1] = std::chrono::duration<double>(t2 - t0).count();
    return count;
}

API void b200host_chain_delete(void *chain) { delete static_cast<fun::b200_receiver_chain *>(chain); }

// One process_samples() call (n >= 0) or flush (n < 0).  Returns the number of payloads; copies them (up to max_out).
API int b200host_chain_process(void *chain, const double *iq, long n, uint8_t *payload_out, int stride, int32_t *len_out,
                               int max_out)
{
    fun::b200_receiver_chain *c = static_cast<fun::b200_receiver_chain *>(chain);
    std::vector<std::vector<unsigned char> > out;
    if (n < 0) out = c->flush();
    else {
        out = c->process_samples(reinterpret_cast<const std::complex<double> *>(iq), (size_t)n);
    }
    int count = 0;
    for (size_t k = 0; k < out.size(); k++, count++) {
        if (count >= max_out) continue;
        int m = (int)out[k].size() < stride ? (int)out[k].size() : stride;
        if (m) std::memcpy(payload_out + (size_t)count * stride, out[k].data(), m);
        len_out[count] = (int32_t)out[k].size();
    }
    return count;
}

API void b200host_chain_counters(void *chain, uint64_t *out7)
{
    fun::b200_receiver_chain::counters_t c = static_cast<fun::b200_receiver_chain *>(chain)->counters();
    out7[0] = c.samples; out7[1] = c.calls; out7[2] = c.frames_found; out7[3] = c.frames_ok; out7[4] = c.frames_crc_fail;
    out7[5] = c.headers_bad; out7[6] = c.frames_truncated;
}

API int b200host_sizeof(int which)
{
    switch (which) {
        case 0: return (int)sizeof(fun::tagged_sample);
        case 1: return (int)sizeof(fun::tagged_vector<64>);
        case 2: return (int)sizeof(fun::tagged_vector<48>);
    }
    return -1;
}

// ---- fun::b200_receiver (receiver.h:36-104: callback + thread + pause/resume) over an in-memory sample source ----
namespace
{
    std::mutex g_rx_mu;
    std::vector<std::vector<unsigned char> > g_rx_packets; // what the callback has been handed so far
    long g_rx_rounds = 0;
    void collect_packets(std::vector<std::vector<unsigned char> > packets)
    {
        std::lock_guard<std::mutex> l(g_rx_mu);
        g_rx_rounds++;
        for (size_t i = 0; i < packets.size(); i++) g_rx_packets.push_back(packets[i]);
    }
}

// Streams iq[0, n) through a b200_receiver in rounds of `chunk` samples.  pause_after >= 0: after that many rounds the
// caller's thread pauses the receiver for pause_ms, checks that no round ran meanwhile (rounds_while_paused) and resumes.
// Returns the number of payloads delivered to the callback (the first max_out are copied out).
API int b200host_receiver_run(const double *iq, long n, long chunk, long pause_after, int pause_ms, unsigned max_frames,
                              unsigned max_payload, uint8_t *payload_out, int stride, int32_t *len_out, int max_out,
                              long *rounds_while_paused)
{
    {
        std::lock_guard<std::mutex> l(g_rx_mu);
        g_rx_packets.clear();
        g_rx_rounds = 0;
    }
    const std::complex<double> *x = reinterpret_cast<const std::complex<double> *>(iq);
    std::atomic<long> pos(0); // advanced by the receiver's thread, read by this one
    fun::b200_receiver::source_t source = [&](std::complex<double> *buf, size_t m) -> bool {
        if (pos >= n) return false;
        const long len = n - pos < (long)m ? n - pos : (long)m;
        std::memcpy(static_cast<void *>(buf), static_cast<const void *>(x + pos), (size_t)len * sizeof(std::complex<double>));
        if (len < (long)m) std::memset(static_cast<void *>(buf + len), 0, (m - (size_t)len) * sizeof(std::complex<double>));
        pos += len;
        return true;
    };
    if (rounds_while_paused) *rounds_while_paused = -1;
    {
        fun::b200_receiver rx(collect_packets, source, (size_t)chunk, 0, max_frames, max_payload);
        if (!rx.ok()) return -1;
        if (pause_after >= 0) {
            for (;;) { // wait until the loop has done pause_after rounds (or has ended)
                long r;
                { std::lock_guard<std::mutex> l(g_rx_mu); r = g_rx_rounds; }
                if (r >= pause_after || pos >= n) break;
                std::this_thread::sleep_for(std::chrono::microseconds(200));
            }
            rx.pause();
            long before;
            { std::lock_guard<std::mutex> l(g_rx_mu); before = g_rx_rounds; }
            std::this_thread::sleep_for(std::chrono::milliseconds(pause_ms));
            long after;
            { std::lock_guard<std::mutex> l(g_rx_mu); after = g_rx_rounds; }
            if (rounds_while_paused) *rounds_while_paused = after - before;
            rx.resume();
        }
        rx.wait();
    }
    std::lock_guard<std::mutex> l(g_rx_mu);
    int count = 0;
    for (size_t k = 0; k < g_rx_packets.size(); k++, count++) {
        if (count >= max_out) continue;
        int m = (int)g_rx_packets[k].size() < stride ? (int)g_rx_packets[k].size() : stride;
        if (m) std::memcpy(payload_out + (size_t)count * stride, g_rx_packets[k].data(), m);
        len_out[count] = (int32_t)g_rx_packets[k].size();
    }
    return count;
}
