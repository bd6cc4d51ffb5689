// C entry points that drive fun::b200_rx from a test harness (ctypes): feed a tagged stream in chunks the
// way receiver_chain::process_samples does (receiver_chain.cpp:106-126), collect the payloads.
#include "b200_rx.h"
#include "b200_receiver_chain.h"

#include <cstring>

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
