// fun::b200_rx implementation (see b200_rx.h).  All signal processing happens in libb200rx.so on the GPU;
// this file is bookkeeping only.  Control flow mirrored from the reference:
//   fft_symbols.cpp:42-56   an LTS1 tag starts a frame (LTS2 is implied 64 samples later)
//   frame_decoder.cpp:72-89 a valid SIGNAL fixes the frame's length; an invalid one is skipped silently
//   frame_decoder.cpp:61-69 the frame is decoded when its last symbol has arrived; payload pushed on CRC ok
//   ppdu.cpp:274-279        "Invalid CRC (length N)" on stderr for a failed frame
#include "b200_rx.h"

#include "../../include/b200rx.h"

#include <cstring>
#include <iostream>
#include <new>

namespace fun
{
    static_assert(sizeof(tagged_sample) == 24, "the GPU unpacks fun::tagged_sample as 16 + 4 + 4 bytes (tagged_vector.h:82-94)");

    b200_rx::b200_rx(int device, unsigned max_frames_per_call, unsigned max_payload, unsigned depth, unsigned max_lag) :
        block("b200_rx"),
        m_handle(nullptr),
        m_max_frames(max_frames_per_call ? max_frames_per_call : 1),
        m_max_payload(max_payload > 4095 ? 4095 : (max_payload ? max_payload : 1)),
        m_depth(depth < 1 ? 1 : (depth > B200RX_MAX_PIPELINE_DEPTH ? B200RX_MAX_PIPELINE_DEPTH : depth)),
        m_max_lag(max_lag),
        m_buf(nullptr), m_buf_n(0), m_buf_cap(0), m_base(0), m_last_lts1(-1), m_round(0), m_pass_seq(0), m_frames(nullptr)
    {
        std::memset(&m_counters, 0, sizeof(m_counters));
        b200rx_limits lim;
        std::memset(&lim, 0, sizeof(lim));
        lim.max_frames = m_max_frames;
        lim.max_payload_bytes = m_max_payload;
        int rc = b200rx_create(device, &lim, &m_handle);
        if (rc == B200RX_OK) rc = b200rx_set_sample_format(m_handle, B200RX_FMT_TAGGED_FC64, 1.0);
        if (rc == B200RX_OK && m_depth > 1 && b200rx_set_pipeline_depth(m_handle, m_depth) != B200RX_OK) {
            std::cerr << "b200_rx: " << b200rx_last_error(m_handle) << " - running with one pass at a time" << std::endl;
            m_depth = 1;
            rc = b200rx_set_pipeline_depth(m_handle, 1);
        }
        bool ok = rc == B200RX_OK;
        if (ok) {
            m_payload.assign(m_depth, nullptr);
            m_status.assign(m_depth, nullptr);
            for (unsigned i = 0; i < m_depth && ok; i++) {
                void *p = nullptr, *s = nullptr;
                ok = b200rx_host_alloc(&p, (size_t)m_max_frames * m_max_payload) == B200RX_OK && b200rx_host_alloc(&s, m_max_frames) == B200RX_OK;
                m_payload[i] = static_cast<uint8_t *>(p);
                m_status[i] = static_cast<uint8_t *>(s);
            }
            m_frames = new (std::nothrow) b200rx_pass_frame[m_max_frames];
            ok = ok && m_frames;
            m_select.resize(m_max_frames);
        }
        if (!ok) {
            m_error = rc != B200RX_OK ? b200rx_last_error(m_handle) : "out of pinned host memory";
            if (m_handle) b200rx_destroy(m_handle);
            m_handle = nullptr;
            std::cerr << "b200_rx: " << m_error << std::endl; // no CPU fallback: the block stays inert
        }
    }

    b200_rx::~b200_rx()
    {
        if (m_handle) {
            b200rx_pass_wait(m_handle, 0);
            b200rx_destroy(m_handle);
        }
        for (uint8_t *p : m_payload) if (p) b200rx_host_free(p);
        for (uint8_t *p : m_status) if (p) b200rx_host_free(p);
        if (m_buf) b200rx_host_free(m_buf);
        delete[] static_cast<b200rx_pass_frame *>(m_frames);
    }

    bool b200_rx::reserve(size_t n)
    {
        if (n <= m_buf_cap) return true;
        size_t cap = m_buf_cap ? m_buf_cap : 16384;
        while (cap < n) cap *= 2;
        void *p = nullptr;
        if (b200rx_host_alloc(&p, cap * sizeof(tagged_sample)) != B200RX_OK) return false;
        if (m_buf_n) std::memcpy(p, static_cast<const void *>(m_buf), m_buf_n * sizeof(tagged_sample));
        if (m_buf) b200rx_host_free(m_buf);
        m_buf = static_cast<tagged_sample *>(p);
        m_buf_cap = cap;
        return true;
    }

    void b200_rx::work()
    {
        output_buffer.resize(0);
        if (!m_handle) return;
        m_round++;
        if (input_buffer.size() == 0) {
            collect(false);
            return;
        }
        // one pass per max_frames * 400 new samples (the shortest frame is 560 samples), normally one per round
        const size_t cap = (size_t)m_max_frames * 400u > 4096u ? (size_t)m_max_frames * 400u : 4096u;
        for (size_t fed = 0; fed < input_buffer.size();) {
            const size_t take = input_buffer.size() - fed < cap ? input_buffer.size() - fed : cap;
            pass(input_buffer.data() + fed, take);
            fed += take;
        }
        if (m_depth == 1 || m_max_lag == 0) collect(true);
    }

    void b200_rx::flush()
    {
        collect(true);
        if (m_buf_n) m_counters.frames_abandoned++; // the frame still arriving
        m_base += m_buf_n;
        m_buf_n = 0;
    }

    void b200_rx::collect(bool all)
    {
        while (!m_inflight.empty()) {
            const pending_pass &p = m_inflight.front();
            const bool must = all || p.seq + m_depth <= m_pass_seq || p.round + m_max_lag <= m_round;
            if (must) {
                if (b200rx_pass_wait(m_handle, p.ticket) != B200RX_OK) {
                    std::cerr << "b200_rx: " << b200rx_last_error(m_handle) << std::endl;
                    m_inflight.pop_front();
                    continue;
                }
            } else {
                const int done = b200rx_pass_poll(m_handle, p.ticket);
                if (done == 0) break;
                if (done < 0) {
                    std::cerr << "b200_rx: " << b200rx_last_error(m_handle) << std::endl;
                    m_inflight.pop_front();
                    continue;
                }
            }
            const size_t slot = (size_t)(p.seq % m_depth);
            for (size_t k = 0; k < p.frames.size(); k++) {
                const uint32_t f = p.frames[k];
                const uint8_t st = m_status[slot][f];
                if (st == B200RX_ST_OK) {
                    const uint8_t *q = m_payload[slot] + (size_t)f * m_max_payload;
                    output_buffer.push_back(std::vector<unsigned char>(q, q + p.len[k]));
                    m_counters.frames_ok++;
                } else if (st == B200RX_ST_CRC_FAIL) {
                    std::cerr << "Invalid CRC (length " << p.len[k] << ")" << std::endl; // ppdu.cpp:276
                    m_counters.frames_crc_fail++;
                } else {
                    m_counters.headers_bad++;
                }
            }
            m_inflight.pop_front();
        }
    }

    // One GPU pass over (the frame still arriving) + n new tagged samples.
    void b200_rx::pass(const tagged_sample *fresh, size_t n)
    {
        collect(false);
        if (!reserve(m_buf_n + n)) {
            std::cerr << "b200_rx: out of pinned host memory" << std::endl;
            return;
        }
        std::memcpy(static_cast<void *>(m_buf + m_buf_n), static_cast<const void *>(fresh), n * sizeof(tagged_sample));
        m_buf_n += n;
        b200rx_pass_frame *frames = static_cast<b200rx_pass_frame *>(m_frames);
        b200rx_sync_result res;
        std::memset(&res, 0, sizeof(res));
        int rc = b200rx_pass_open(m_handle);
        if (rc == B200RX_OK) rc = b200rx_pass_put(m_handle, m_buf, m_buf_n);
        if (rc == B200RX_OK) {
            m_pass_seq++;
            rc = b200rx_pass_scan_tagged(m_handle, frames, m_max_frames, &res);
        }
        if (rc != B200RX_OK) {
            std::cerr << "b200_rx: " << b200rx_last_error(m_handle) << std::endl;
            m_base += m_buf_n;
            m_buf_n = 0;
            return;
        }
        if (res.overflow)
            std::cerr << "b200_rx: " << res.overflow << " frames dropped, more than max_frames_per_call in one round" << std::endl;

        const uint32_t nf = res.n_frames < m_max_frames ? res.n_frames : m_max_frames;
        if (nf) std::memset(m_select.data(), 0, nf);
        pending_pass pend;
        int64_t pending_at = -1; // buffer index of the LTS1 tag of the frame still arriving
        for (uint32_t f = 0; f < nf; f++) {
            const b200rx_pass_frame &fr = frames[f];
            const int64_t lts1 = (int64_t)(m_base + fr.lts1);
            const bool last = f + 1 == nf && res.overflow == 0;
            if (fr.status == B200RX_ST_TRUNCATED && last) { // its samples may still be arriving
                pending_at = (int64_t)fr.lts1;
                if (lts1 > m_last_lts1) { m_counters.frames_seen++; m_last_lts1 = lts1; }
                break;
            }
            if (lts1 > m_last_lts1) { m_counters.frames_seen++; m_last_lts1 = lts1; }
            else if (!(f == 0 && fr.lts1 == 0)) continue; // (only the retained frame is ever seen twice)
            if (fr.status == B200RX_ST_OK) {
                m_select[f] = 1;
                pend.frames.push_back(f);
                pend.len.push_back(fr.length);
            } else if (fr.status == B200RX_ST_TRUNCATED) {
                m_counters.frames_abandoned++;  // a new LTS1 arrived inside it
            } else {
                m_counters.headers_bad++;       // frame_decoder.cpp:78: skipped silently
            }
        }
        if (!pend.frames.empty()) {
            const size_t slot = (size_t)((m_pass_seq - 1) % m_depth);
            uint64_t ticket = 0;
            rc = b200rx_pass_decode(m_handle, m_select.data(), m_payload[slot], m_max_payload, m_status[slot], &ticket);
            if (rc != B200RX_OK) {
                std::cerr << "b200_rx: " << b200rx_last_error(m_handle) << std::endl;
            } else {
                pend.ticket = ticket;
                pend.seq = m_pass_seq - 1;
                pend.round = m_round;
                m_inflight.push_back(std::move(pend));
            }
        }
        // keep the frame still arriving, drop the rest
        if (pending_at >= 0) {
            const size_t drop = (size_t)pending_at;
            m_buf_n -= drop;
            if (drop && m_buf_n) std::memmove(static_cast<void *>(m_buf), static_cast<const void *>(m_buf + drop), m_buf_n * sizeof(tagged_sample));
            m_base += drop;
        } else {
            m_base += m_buf_n;
            m_buf_n = 0;
        }
    }
}
