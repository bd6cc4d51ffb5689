// fun::b200_rx implementation (see b200_rx.h).  All signal processing happens in libb200rx.so on the GPU;
// this file is bookkeeping only.  Control-flow mirrored from the reference:
//   fft_symbols.cpp:42-56   an LTS1 tag starts a frame (LTS2 is implied 64 samples later)
//   frame_decoder.cpp:72-89 a valid SIGNAL fixes the frame's length; an invalid one is skipped silently
//   frame_decoder.cpp:61-69 the frame is decoded when its last symbol has arrived; payload pushed on CRC ok
//   ppdu.cpp:274-279        "Invalid CRC (length N)" on stderr for a failed frame
#include "b200_rx.h"

#include "../../include/b200rx.h"

#include <cstring>
#include <iostream>

namespace fun
{
    b200_rx::b200_rx(int device, unsigned max_frames_per_call, unsigned max_payload) :
        block("b200_rx"),
        m_handle(nullptr),
        m_max_frames(max_frames_per_call ? max_frames_per_call : 1),
        m_max_payload(max_payload > 4095 ? 4095 : max_payload)
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
            std::cerr << "b200_rx: " << m_error << std::endl; // no CPU fallback: the block stays inert
        }
    }

    b200_rx::~b200_rx()
    {
        if (m_handle) b200rx_destroy(m_handle);
    }

    void b200_rx::work()
    {
        output_buffer.resize(0);
        if (input_buffer.size() == 0 || !m_handle) return;

        for (size_t x = 0; x < input_buffer.size(); x++) {
            const tagged_sample &s = input_buffer[x];
            if (s.tag == LTS1) {
                // a frame still arriving is abandoned by a new LTS1 (see header comment)
                while (!m_open.empty()) {
                    if (m_open.front().needed < 0 || (int)m_open.front().samples.size() < m_open.front().needed)
                        m_counters.frames_abandoned++;
                    m_open.pop_front();
                }
                m_open.push_back(capture());
                m_open.back().needed = -1;
                m_open.back().header_tried = false;
                m_open.back().samples.reserve(8192);
                m_counters.frames_seen++;
            }
            if (!m_open.empty()) {
                capture &c = m_open.back();
                if (c.needed < 0 || (int)c.samples.size() < c.needed) c.samples.push_back(s.sample);
                if (c.needed > 0 && (int)c.samples.size() >= c.needed) {
                    m_ready.push_back(capture());
                    m_ready.back().samples.swap(c.samples);
                    m_ready.back().needed = c.needed;
                    m_open.pop_back();
                    if (m_ready.size() >= m_max_frames) decode_ready();
                } else if (c.needed < 0 && !c.header_tried && c.samples.size() >= 208) {
                    decode_headers();
                }
            }
        }
        decode_ready();
    }

    void b200_rx::flush()
    {
        decode_ready();
        m_counters.frames_abandoned += m_open.size();
        m_open.clear();
    }

    // SIGNAL decode for every open frame that has its first 208 samples (fft_symbols windows [0,64), [64,128),
    // [144,208)): sets capture::needed = 128 + 80 * (1 + nsym), or drops the frame on a bad header.
    void b200_rx::decode_headers()
    {
        std::vector<double> iq;
        std::vector<uint64_t> off;
        std::vector<uint32_t> avail;
        std::vector<size_t> which;
        for (size_t i = 0; i < m_open.size(); i++) {
            capture &c = m_open[i];
            if (c.needed >= 0 || c.header_tried || c.samples.size() < 208) continue;
            off.push_back(iq.size() / 2);
            avail.push_back(208);
            const double *p = reinterpret_cast<const double *>(c.samples.data());
            iq.insert(iq.end(), p, p + 2 * 208);
            which.push_back(i);
            c.header_tried = true;
        }
        if (which.empty()) return;
        std::vector<uint16_t> len(which.size());
        std::vector<uint8_t> rate(which.size()), status(which.size());
        int rc = b200rx_decode_headers(m_handle, iq.data(), iq.size() / 2, off.data(), avail.data(), (uint32_t)which.size(),
                                       len.data(), rate.data(), status.data());
        if (rc != B200RX_OK) {
            std::cerr << "b200_rx: " << b200rx_last_error(m_handle) << std::endl;
            return;
        }
        static const int DBPS[11] = {24, 32, 36, 48, 64, 72, 96, 128, 144, 192, 216}; // rates.h:52-196
        std::vector<size_t> drop;
        for (size_t k = 0; k < which.size(); k++) {
            capture &c = m_open[which[k]];
            if (status[k] == B200RX_ST_OK && rate[k] <= 10) {
                const int nsym = (16 + 8 * ((int)len[k] + 4) + 6 + DBPS[rate[k]] - 1) / DBPS[rate[k]]; // ppdu.cpp:207-209
                c.needed = 128 + 80 * (1 + nsym);
            } else {
                m_counters.headers_bad++;
                drop.push_back(which[k]);
            }
        }
        for (size_t k = drop.size(); k-- > 0;) m_open.erase(m_open.begin() + drop[k]);
    }

    void b200_rx::decode_ready()
    {
        while (!m_ready.empty()) {
            const size_t n = m_ready.size() < m_max_frames ? m_ready.size() : m_max_frames;
            std::vector<uint64_t> off(n);
            std::vector<uint32_t> avail(n);
            size_t total = 0;
            for (size_t i = 0; i < n; i++) { off[i] = total; avail[i] = (uint32_t)m_ready[i].samples.size(); total += avail[i]; }
            std::vector<std::complex<double> > iq(total);
            for (size_t i = 0; i < n; i++)
                std::memcpy(&iq[off[i]], m_ready[i].samples.data(), sizeof(std::complex<double>) * avail[i]);
            const uint32_t stride = m_max_payload ? m_max_payload : 1;
            std::vector<uint8_t> payload(n * (size_t)stride), rate(n), status(n);
            std::vector<uint16_t> len(n);
            int rc = b200rx_decode_batch(m_handle, reinterpret_cast<const double *>(iq.data()), total, off.data(), avail.data(),
                                         (uint32_t)n, payload.data(), stride, len.data(), rate.data(), status.data());
            if (rc != B200RX_OK) {
                std::cerr << "b200_rx: " << b200rx_last_error(m_handle) << std::endl;
            } else {
                for (size_t i = 0; i < n; i++) {
                    if (status[i] == B200RX_ST_OK) {
                        output_buffer.push_back(std::vector<unsigned char>(payload.begin() + i * stride,
                                                                           payload.begin() + i * stride + len[i]));
                        m_counters.frames_ok++;
                    } else if (status[i] == B200RX_ST_CRC_FAIL) {
                        std::cerr << "Invalid CRC (length " << len[i] << ")" << std::endl; // ppdu.cpp:276
                        m_counters.frames_crc_fail++;
                    } else {
                        m_counters.headers_bad++;
                    }
                }
            }
            m_ready.erase(m_ready.begin(), m_ready.begin() + n);
        }
    }
}
