// fun::b200_receiver_chain implementation (see b200_receiver_chain.h).  Bookkeeping only: all signal processing is the
// b200rx_pass_* calls on the GPU.
#include "b200_receiver_chain.h"

#include "../../include/b200rx.h"

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <iostream>
#include <mutex>
#include <thread>

namespace fun
{
    namespace
    {
        const uint64_t SYNC_CARRYOVER = 160; // timing_sync.h: CARRYOVER_LENGTH; tags in the last 160 samples wait (timing_sync.cpp:68)
        const uint64_t KEEP_BEFORE = 512;    // history kept in front of the first unexamined tag (detector needs 48, a tag's
                                             // LTS1 lies within [-8, +88) of it)
        const size_t DIRECT_MIN = 65536;     // calls at least this long are copied to the GPU straight from pinned caller memory
        const size_t SLICE = (size_t)1 << 18; // samples per staging slice of a long call (4 MB: copy of slice i+1 overlaps H2D of slice i)
        // samples per GPU pass when a caller hands over a very long vector: no more than can hold max_frames frames
        // (the shortest frame, BPSK 1/2 with an empty payload, is 320 + 80 * 3 = 560 samples; 400 leaves a margin)
        inline uint64_t max_capture(unsigned max_frames)
        {
            const uint64_t m = (uint64_t)max_frames * 400u;
            return m < 4096u ? 4096u : (m > (1u << 22) ? (1u << 22) : m);
        }
    }

    // A few host threads that copy one large block in parallel (a single core moves pageable memory into pinned memory
    // at 6-10 GB/s, a quarter of what the PCIe link takes).
    class b200_copy_pool
    {
    public:
        explicit b200_copy_pool(unsigned helpers) : m_gen(0), m_left(0), m_stop(false), m_dst(nullptr), m_src(nullptr), m_bytes(0)
        {
            for (unsigned i = 0; i < helpers; i++) m_threads.push_back(std::thread([this, i] { loop(i); }));
        }
        ~b200_copy_pool()
        {
            {
                std::lock_guard<std::mutex> l(m_mu);
                m_stop = true;
                m_gen++;
            }
            m_cv.notify_all();
            for (auto &t : m_threads) t.join();
        }
        void copy(void *dst, const void *src, size_t bytes)
        {
            const size_t parts = m_threads.size() + 1;
            if (parts == 1 || bytes < ((size_t)1 << 20)) {
                std::memcpy(dst, src, bytes);
                return;
            }
            {
                std::lock_guard<std::mutex> l(m_mu);
                m_dst = static_cast<char *>(dst);
                m_src = static_cast<const char *>(src);
                m_bytes = bytes;
                m_left = (unsigned)m_threads.size();
                m_gen++;
            }
            m_cv.notify_all();
            part(parts - 1, parts); // the caller's share
            std::unique_lock<std::mutex> l(m_mu);
            m_done.wait(l, [this] { return m_left == 0; });
        }
        unsigned helpers() const { return (unsigned)m_threads.size(); }

    private:
        void part(size_t i, size_t parts)
        {
            const size_t per = ((m_bytes / parts) + 4095) & ~(size_t)4095;
            const size_t a = i * per < m_bytes ? i * per : m_bytes;
            const size_t b = (i + 1 == parts) ? m_bytes : (a + per < m_bytes ? a + per : m_bytes);
            if (b > a) std::memcpy(m_dst + a, m_src + a, b - a);
        }
        void loop(unsigned i)
        {
            uint64_t seen = 0;
            for (;;) {
                {
                    std::unique_lock<std::mutex> l(m_mu);
                    m_cv.wait(l, [&] { return m_gen != seen; });
                    seen = m_gen;
                    if (m_stop) return;
                }
                part(i, m_threads.size() + 1);
                {
                    std::lock_guard<std::mutex> l(m_mu);
                    m_left--;
                }
                m_done.notify_all();
            }
        }
        std::vector<std::thread> m_threads;
        std::mutex m_mu;
        std::condition_variable m_cv, m_done;
        uint64_t m_gen;
        unsigned m_left;
        bool m_stop;
        char *m_dst;
        const char *m_src;
        size_t m_bytes;
    };

    b200_receiver_chain::b200_receiver_chain(int device, unsigned max_frames, unsigned max_payload, unsigned depth, unsigned max_lag) :
        m_handle(nullptr),
        m_max_frames(max_frames ? max_frames : 1),
        m_max_payload(max_payload > 4095 ? 4095 : (max_payload ? max_payload : 1)),
        m_depth(depth < 1 ? 1 : (depth > B200RX_MAX_PIPELINE_DEPTH ? B200RX_MAX_PIPELINE_DEPTH : depth)),
        m_max_lag(max_lag),
        m_buf(nullptr), m_buf_n(0), m_buf_cap(0),
        m_base(0), m_handled(0), m_stream_start(0), m_last_lts1(-1), m_pending_lts1(-1), m_phase(0.0),
        m_pass_seq(0), m_frames(nullptr), m_pool(nullptr)
    {
        std::memset(&m_counters, 0, sizeof(m_counters));
        b200rx_limits lim;
        std::memset(&lim, 0, sizeof(lim));
        lim.max_frames = m_max_frames;
        lim.max_payload_bytes = m_max_payload;
        int rc = b200rx_create(device, &lim, &m_handle);
        if (rc == B200RX_OK && m_depth > 1) {
            rc = b200rx_set_pipeline_depth(m_handle, m_depth);
            if (rc != B200RX_OK) { // not enough device memory for that many lanes: fall back to what the handle has
                std::cerr << "b200_receiver_chain: " << b200rx_last_error(m_handle) << " - running with one pass at a time" << std::endl;
                m_depth = 1;
                rc = b200rx_set_pipeline_depth(m_handle, 1);
            }
        }
        if (rc != B200RX_OK) {
            m_error = b200rx_last_error(m_handle);
            if (m_handle) b200rx_destroy(m_handle);
            m_handle = nullptr;
            std::cerr << "b200_receiver_chain: " << m_error << std::endl; // no CPU fallback: the chain stays inert
            return;
        }
        m_payload.assign(m_depth, nullptr);
        m_status.assign(m_depth, nullptr);
        bool ok = true;
        for (unsigned i = 0; i < m_depth && ok; i++) {
            void *p = nullptr, *s = nullptr;
            ok = b200rx_host_alloc(&p, (size_t)m_max_frames * m_max_payload) == B200RX_OK && b200rx_host_alloc(&s, m_max_frames) == B200RX_OK;
            m_payload[i] = static_cast<uint8_t *>(p);
            m_status[i] = static_cast<uint8_t *>(s);
        }
        m_frames = new (std::nothrow) b200rx_pass_frame[m_max_frames];
        if (!ok || !m_frames) {
            m_error = "out of pinned host memory";
            std::cerr << "b200_receiver_chain: " << m_error << std::endl;
            b200rx_destroy(m_handle);
            m_handle = nullptr;
            return;
        }
        m_select.resize(m_max_frames);
        set_copy_threads(4); // measured on a 16-core host: 4 threads stage 1 Mi-sample calls at 12-14 GB/s, 8 are no faster
    }

    b200_receiver_chain::~b200_receiver_chain()
    {
        if (m_handle) {
            b200rx_pass_wait(m_handle, 0);
            b200rx_destroy(m_handle);
        }
        for (uint8_t *p : m_payload) if (p) b200rx_host_free(p);
        for (uint8_t *p : m_status) if (p) b200rx_host_free(p);
        if (m_buf) b200rx_host_free(m_buf);
        delete[] static_cast<b200rx_pass_frame *>(m_frames);
        delete m_pool;
    }

    int b200_receiver_chain::set_tuning(const char *key, long long value)
    {
        if (!m_handle) return B200RX_E_DEVICE;
        std::vector<std::vector<unsigned char> > sink;
        collect(sink, true); // nothing in flight may see two settings (payloads collected here are dropped: call it between streams)
        return b200rx_set_tuning(m_handle, key, (int64_t)value);
    }

    void b200_receiver_chain::set_copy_threads(unsigned n)
    {
        if (n < 1) n = 1;
        if (n > 16) n = 16;
        if (m_pool && m_pool->helpers() == n - 1) return;
        delete m_pool;
        m_pool = nullptr;
        if (n > 1) m_pool = new b200_copy_pool(n - 1);
    }

    std::complex<double> *b200_receiver_chain::alloc_samples(size_t n)
    {
        void *p = nullptr;
        if (b200rx_host_alloc(&p, (n ? n : 1) * sizeof(std::complex<double>)) != B200RX_OK) return nullptr;
        return static_cast<std::complex<double> *>(p);
    }

    void b200_receiver_chain::free_samples(std::complex<double> *p) { b200rx_host_free(p); }

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

    // Copies n samples behind the retained ones (capacity reserved by the caller) and queues their H2D copy slice by slice.
    void b200_receiver_chain::stage(const std::complex<double> *src, size_t n)
    {
        for (size_t off = 0; off < n; off += SLICE) {
            const size_t len = n - off < SLICE ? n - off : SLICE;
            std::complex<double> *dst = m_buf + m_buf_n + off;
            if (m_pool) m_pool->copy(static_cast<void *>(dst), static_cast<const void *>(src + off), len * sizeof(std::complex<double>));
            else std::memcpy(static_cast<void *>(dst), static_cast<const void *>(src + off), len * sizeof(std::complex<double>));
            b200rx_pass_put(m_handle, dst, len);
        }
        m_buf_n += n;
    }

    std::vector<std::vector<unsigned char> > b200_receiver_chain::process_samples(const std::complex<double> *samples, size_t n)
    {
        std::vector<std::vector<unsigned char> > out;
        if (!m_handle) return out;
        m_counters.calls++;
        m_counters.samples += n;
        m_calls.push_back(m_base + m_buf_n);
        const bool pinned = n >= DIRECT_MIN && b200rx_host_is_pinned(samples) == 1;
        size_t fed = 0;
        do { // one GPU pass per max_capture new samples (normally one pass per call)
            const size_t cap = (size_t)max_capture(m_max_frames);
            const size_t take = n - fed < cap ? n - fed : cap;
            const bool direct = pinned && take >= DIRECT_MIN;
            collect(out, false);
            if (!reserve(m_buf_n + (direct ? 0 : take))) {
                std::cerr << "b200_receiver_chain: out of pinned host memory" << std::endl;
                return out;
            }
            int rc = b200rx_pass_open(m_handle);
            if (rc == B200RX_OK) rc = b200rx_pass_put(m_handle, m_buf, m_buf_n); // the retained tail
            if (rc != B200RX_OK) {
                std::cerr << "b200_receiver_chain: " << b200rx_last_error(m_handle) << std::endl;
                return out;
            }
            m_pass_seq++;
            if (direct) {
                b200rx_pass_put(m_handle, samples + fed, take);
                run_capture(samples + fed, take, out);
            } else {
                stage(samples + fed, take);
                run_capture(nullptr, 0, out);
            }
            fed += take;
        } while (fed < n);
        if (m_depth == 1 || m_max_lag == 0) collect(out, true);
        return out;
    }

    std::vector<std::vector<unsigned char> > b200_receiver_chain::flush(unsigned pad)
    {
        std::vector<std::vector<unsigned char> > out = process_samples(std::vector<std::complex<double> >(pad));
        collect(out, true);
        // whatever comes next is a new stream: nothing retained, no phase carried over, no chunk boundaries of the old one
        m_base += m_buf_n;
        m_buf_n = 0;
        m_handled = m_base;
        m_stream_start = m_base;
        m_pending_lts1 = -1;
        m_phase = 0.0;
        m_calls.clear();
        return out;
    }

    // Hands out the passes that have finished, oldest first; waits for those that must not stay in flight any longer:
    // everything (all), a pass whose output slot the next pass reuses, a pass older than max_lag calls.
    void b200_receiver_chain::collect(std::vector<std::vector<unsigned char> > &out, bool all)
    {
        while (!m_inflight.empty()) {
            const pending_pass &p = m_inflight.front();
            const bool must = all || p.seq + m_depth <= m_pass_seq || p.call + m_max_lag <= m_counters.calls;
            if (must) {
                if (b200rx_pass_wait(m_handle, p.ticket) != B200RX_OK) {
                    std::cerr << "b200_receiver_chain: " << b200rx_last_error(m_handle) << std::endl;
                    m_inflight.pop_front();
                    continue;
                }
            } else {
                const int done = b200rx_pass_poll(m_handle, p.ticket);
                if (done == 0) break;
                if (done < 0) {
                    std::cerr << "b200_receiver_chain: " << b200rx_last_error(m_handle) << std::endl;
                    m_inflight.pop_front();
                    continue;
                }
            }
            deliver(p, out);
            m_inflight.pop_front();
        }
    }

    void b200_receiver_chain::deliver(const pending_pass &p, std::vector<std::vector<unsigned char> > &out)
    {
        const size_t slot = (size_t)(p.seq % m_depth);
        for (size_t k = 0; k < p.frames.size(); k++) {
            const uint32_t f = p.frames[k];
            const uint8_t st = m_status[slot][f];
            if (st == B200RX_ST_OK) {
                const uint8_t *q = m_payload[slot] + (size_t)f * m_max_payload;
                out.push_back(std::vector<unsigned char>(q, q + p.len[k]));
                m_counters.frames_ok++;
            } else if (st == B200RX_ST_CRC_FAIL) {
                std::cerr << "Invalid CRC (length " << p.len[k] << ")" << std::endl; // ppdu.cpp:276
                m_counters.frames_crc_fail++;
            } else {
                m_counters.headers_bad++;
            }
        }
    }

    // The pass is open and holds m_buf[0, m_buf_n) followed by direct[0, n_direct): scan it, settle every frame whose
    // fate is known, queue the decode of the complete ones, and keep what the next call must see again.
    void b200_receiver_chain::run_capture(const std::complex<double> *direct, size_t n_direct,
                                          std::vector<std::vector<unsigned char> > &out)
    {
        (void)out;
        const uint64_t n = m_buf_n + n_direct;
        if (n == 0) return;
        // origins of the reference's work() buffers inside the retained samples: chunk start - 160, relative to m_buf[0];
        // the last one at or before the buffer start and every later one
        while (m_calls.size() > 1 && (int64_t)m_calls[1] - 160 <= (int64_t)m_base) m_calls.pop_front();
        std::vector<int64_t> origins(m_calls.size());
        for (size_t i = 0; i < m_calls.size(); i++) origins[i] = (int64_t)m_calls[i] - 160 - (int64_t)m_base;
        b200rx_set_receive_origins(m_handle, origins.data(), (uint32_t)origins.size());
        b200rx_sync_result res;
        std::memset(&res, 0, sizeof(res));
        b200rx_pass_frame *frames = static_cast<b200rx_pass_frame *>(m_frames);
        int rc = b200rx_pass_scan(m_handle, m_phase, frames, m_max_frames, &res);
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
        pending_pass pend;
        const uint32_t nf = res.n_frames < m_max_frames ? res.n_frames : m_max_frames;
        if (nf) std::memset(m_select.data(), 0, nf);
        for (uint32_t f = 0; f < nf; f++) {
            const b200rx_pass_frame &fr = frames[f];
            const int64_t lts1 = (int64_t)(m_base + fr.lts1);
            if (lts1 <= m_last_lts1) continue;
            if (m_base > m_stream_start && fr.lts1 < KEEP_BEFORE / 2 && lts1 != m_pending_lts1) continue;
            const uint8_t st = fr.status;
            if (st == B200RX_ST_TRUNCATED && f + 1 == nf) { // may still be arriving
                pending_lts1 = lts1;
                break;
            }
            m_last_lts1 = lts1;
            m_counters.frames_found++;
            if (st == B200RX_ST_OK) {             // verdict (payload / "Invalid CRC") when the decode has finished
                m_select[f] = 1;
                pend.frames.push_back(f);
                pend.len.push_back(fr.length);
            } else if (st == B200RX_ST_TRUNCATED) {
                m_counters.frames_truncated++;    // cut short by the next frame's LTS1 (fft_symbols.cpp:42-51)
            } else {
                m_counters.headers_bad++;         // frame_decoder.cpp:78: skipped silently
            }
        }
        if (!pend.frames.empty()) {
            const size_t slot = (size_t)((m_pass_seq - 1) % m_depth);
            uint64_t ticket = 0;
            rc = b200rx_pass_decode(m_handle, m_select.data(), m_payload[slot], m_max_payload, m_status[slot], &ticket);
            if (rc != B200RX_OK) {
                std::cerr << "b200_receiver_chain: " << b200rx_last_error(m_handle) << std::endl;
            } else {
                pend.ticket = ticket;
                pend.seq = m_pass_seq - 1;
                pend.call = m_counters.calls;
                m_inflight.push_back(std::move(pend));
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
        // m_phase_acc in front of the new buffer: what the last frame whose STS_END tag lies in the dropped part left behind.
        // (A frame retained for the next call is synchronised again there; the samples in front of its STS_END tag must
        // then be rotated by its predecessor's phase, timing_sync.cpp:121-125, not by its own.)
        if (keep_from > m_base)
            for (uint32_t f = 0; f < nf; f++)
                if (m_base + frames[f].sts_end < keep_from) m_phase = frames[f].phase;
        const size_t drop = (size_t)(keep_from - m_base);
        size_t kept = 0;
        if (drop < m_buf_n) {
            kept = m_buf_n - drop;
            if (drop) std::memmove(static_cast<void *>(m_buf), static_cast<const void *>(m_buf + drop), kept * sizeof(std::complex<double>));
        }
        if (n_direct) { // the part of the caller's buffer that must be seen again moves behind it
            const size_t from = drop > m_buf_n ? drop - m_buf_n : 0;
            const size_t cnt = n_direct - from;
            m_buf_n = kept;
            if (!reserve(kept + cnt)) {
                std::cerr << "b200_receiver_chain: out of pinned host memory" << std::endl;
                m_base += n;
                m_buf_n = 0;
                m_handled = m_base;
                return;
            }
            if (cnt) std::memcpy(static_cast<void *>(m_buf + kept), static_cast<const void *>(direct + from), cnt * sizeof(std::complex<double>));
            kept += cnt;
        }
        m_buf_n = kept;
        m_base = keep_from;
    }
}
