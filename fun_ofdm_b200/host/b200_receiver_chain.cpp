// fun::b200_receiver_chain implementation (see b200_receiver_chain.h).  Bookkeeping only: all signal processing is
// b200rx_receive on the GPU.
#include "b200_receiver_chain.h"

#include "../../include/b200rx.h"

#include <cstring>
#include <iostream>

namespace fun
{
    namespace
    {
        const uint64_t SYNC_CARRYOVER = 160; // timing_sync.h: CARRYOVER_LENGTH; tags in the last 160 samples wait (timing_sync.cpp:68)
        const uint64_t KEEP_BEFORE = 512;    // history kept in front of the first unexamined tag (detector needs 48, a tag's
                                             // LTS1 lies within [-8, +88) of it)
        // samples per GPU pass when a caller hands over a very long vector: no more than can hold max_frames frames
        // (the shortest frame, BPSK 1/2 with an empty payload, is 320 + 80 * 3 = 560 samples; 400 leaves a margin)
        inline uint64_t max_capture(unsigned max_frames)
        {
            const uint64_t m = (uint64_t)max_frames * 400u;
            return m < 65536u ? 65536u : (m > (1u << 22) ? (1u << 22) : m);
        }
    }

    b200_receiver_chain::b200_receiver_chain(int device, unsigned max_frames, unsigned max_payload) :
        m_handle(nullptr),
        m_max_frames(max_frames ? max_frames : 1),
        m_max_payload(max_payload > 4095 ? 4095 : (max_payload ? max_payload : 1)),
        m_buf(nullptr), m_buf_n(0), m_buf_cap(0),
        m_base(0), m_handled(0), m_last_lts1(-1), m_pending_lts1(-1), m_phase(0.0)
    {
        std::memset(&m_counters, 0, sizeof(m_counters));
        b200rx_limits lim;
        std::memset(&lim, 0, sizeof(lim));
        lim.max_frames = m_max_frames;
        lim.max_payload_bytes = m_max_payload;
        int rc = b200rx_create(device, &lim, &m_handle);
        if (rc != B200RX_OK) {
            m_error = b200rx_last_error(nullptr);
            m_handle = nullptr;
            std::cerr << "b200_receiver_chain: " << m_error << std::endl; // no CPU fallback: the chain stays inert
            return;
        }
        m_payload.resize((size_t)m_max_frames * m_max_payload);
        m_rate.resize(m_max_frames);
        m_status.resize(m_max_frames);
        m_len.resize(m_max_frames);
        m_lts1.resize(m_max_frames);
    }

    b200_receiver_chain::~b200_receiver_chain()
    {
        if (m_handle) b200rx_destroy(m_handle);
        if (m_buf) b200rx_host_free(m_buf);
    }

    bool b200_receiver_chain::reserve(size_t n)
    {
        if (n <= m_buf_cap) return true;
        size_t cap = m_buf_cap ? m_buf_cap : 65536;
        while (cap < n) cap *= 2;
        void *p = nullptr;
        if (b200rx_host_alloc(&p, cap * sizeof(std::complex<double>)) != B200RX_OK) return false;
        if (m_buf_n) std::memcpy(p, static_cast<const void *>(m_buf), m_buf_n * sizeof(std::complex<double>));
        if (m_buf) b200rx_host_free(m_buf);
        m_buf = static_cast<std::complex<double> *>(p);
        m_buf_cap = cap;
        return true;
    }

    std::vector<std::vector<unsigned char> > b200_receiver_chain::process_samples(std::vector<std::complex<double> > samples)
    {
        return process_samples(samples.data(), samples.size());
    }

    std::vector<std::vector<unsigned char> > b200_receiver_chain::process_samples(const std::complex<double> *samples, size_t n)
    {
        std::vector<std::vector<unsigned char> > out;
        if (!m_handle) return out;
        m_counters.calls++;
        m_counters.samples += n;
        m_calls.push_back(m_base + m_buf_n);
        size_t fed = 0;
        do { // one GPU pass per MAX_CAPTURE new samples (normally one pass per call)
            const size_t cap = (size_t)max_capture(m_max_frames);
            const size_t take = n - fed < cap ? n - fed : cap;
            if (!reserve(m_buf_n + take)) {
                std::cerr << "b200_receiver_chain: out of pinned host memory" << std::endl;
                return out;
            }
            if (take) std::memcpy(static_cast<void *>(m_buf + m_buf_n), static_cast<const void *>(samples + fed),
                                  take * sizeof(std::complex<double>));
            m_buf_n += take;
            fed += take;
            run_capture(out);
        } while (fed < n);
        return out;
    }

    std::vector<std::vector<unsigned char> > b200_receiver_chain::flush(unsigned pad)
    {
        std::vector<std::vector<unsigned char> > out = process_samples(std::vector<std::complex<double> >(pad));
        m_base += m_buf_n;
        m_buf_n = 0;
        m_handled = m_base;
        m_pending_lts1 = -1;
        return out;
    }

    void b200_receiver_chain::run_capture(std::vector<std::vector<unsigned char> > &out)
    {
        const uint64_t n = m_buf_n;
        if (n == 0) return;
        // origins of the reference's work() buffers inside the retained samples: chunk start - 160, relative to m_buf[0];
        // the last one at or before the buffer start and every later one
        while (m_calls.size() > 1 && (int64_t)m_calls[1] - 160 <= (int64_t)m_base) m_calls.pop_front();
        std::vector<int64_t> origins(m_calls.size());
        for (size_t i = 0; i < m_calls.size(); i++) origins[i] = (int64_t)m_calls[i] - 160 - (int64_t)m_base;
        b200rx_set_receive_origins(m_handle, origins.data(), (uint32_t)origins.size());
        b200rx_sync_result res;
        std::memset(&res, 0, sizeof(res));
        int rc = b200rx_receive(m_handle, m_buf, n, m_phase, m_payload.data(),
                                m_max_payload, m_len.data(), m_rate.data(), m_status.data(), m_lts1.data(), &res);
        if (rc != B200RX_OK) {
            std::cerr << "b200_receiver_chain: " << b200rx_last_error(m_handle) << std::endl;
            m_base += n;
            m_buf_n = 0;
            m_handled = m_base;
            return;
        }
        if (res.overflow)
            std::cerr << "b200_receiver_chain: " << res.overflow << " frames dropped, more than max_frames in one call" << std::endl;

        // Frames in stream order.  The retained part of the buffer was examined by earlier calls: frames up to the last one
        // settled are skipped, and so is anything "found" in the first samples of a buffer that does not start the stream
        // (there the detector's history is missing; every tag not yet examined lies >= KEEP_BEFORE - 8 samples further on).
        // The last frame, if its samples end with the buffer, waits for the next call.
        int64_t pending_lts1 = -1;
        for (uint32_t f = 0; f < res.n_frames; f++) {
            const int64_t lts1 = (int64_t)(m_base + m_lts1[f]);
            if (lts1 <= m_last_lts1) continue;
            if (m_base > 0 && m_lts1[f] < KEEP_BEFORE / 2 && lts1 != m_pending_lts1) continue;
            const uint8_t st = m_status[f];
            if (st == B200RX_ST_TRUNCATED && f + 1 == res.n_frames) { // may still be arriving
                pending_lts1 = lts1;
                break;
            }
            m_last_lts1 = lts1;
            m_counters.frames_found++;
            if (st == B200RX_ST_OK) {
                const uint8_t *p = m_payload.data() + (size_t)f * m_max_payload;
                out.push_back(std::vector<unsigned char>(p, p + m_len[f]));
                m_counters.frames_ok++;
            } else if (st == B200RX_ST_CRC_FAIL) {
                std::cerr << "Invalid CRC (length " << m_len[f] << ")" << std::endl; // ppdu.cpp:276
                m_counters.frames_crc_fail++;
            } else if (st == B200RX_ST_TRUNCATED) {
                m_counters.frames_truncated++; // cut short by the next frame's LTS1 (fft_symbols.cpp:42-51)
            } else {
                m_counters.headers_bad++;     // frame_decoder.cpp:78: skipped silently
            }
        }

        m_pending_lts1 = pending_lts1;

        // What the next call must see again.
        const uint64_t end = m_base + n;
        m_handled = end > SYNC_CARRYOVER ? end - SYNC_CARRYOVER : 0;
        if (m_handled < m_base) m_handled = m_base;
        uint64_t keep_from = m_handled > KEEP_BEFORE ? m_handled - KEEP_BEFORE : 0;
        if (pending_lts1 >= 0) {
            const uint64_t p = (uint64_t)pending_lts1 > KEEP_BEFORE ? (uint64_t)pending_lts1 - KEEP_BEFORE : 0;
            if (p < keep_from) keep_from = p;
        }
        if (keep_from < m_base) keep_from = m_base;
        if (keep_from > m_base) {
            // m_phase_acc in front of the new buffer: the last synchronised frame's (it either lies in the dropped part,
            // or it is recomputed from the retained samples and this value is not used)
            if (res.phase_valid) m_phase = res.last_phase;
            const size_t drop = (size_t)(keep_from - m_base);
            m_buf_n -= drop;
            if (m_buf_n) std::memmove(static_cast<void *>(m_buf), static_cast<const void *>(m_buf + drop),
                                      m_buf_n * sizeof(std::complex<double>));
            m_base = keep_from;
        }
    }
}
