// fun::b200_receiver implementation (see b200_receiver.h); control flow of the reference's receiver.cpp:42-77.
#include "b200_receiver.h"

namespace fun
{
    b200_receiver::b200_receiver(callback_t callback, source_t source, size_t num_rx_samples, int device, unsigned max_frames,
                                 unsigned max_payload) :
        m_callback(callback),
        m_source(source),
        m_n(num_rx_samples ? num_rx_samples : 4096),
        m_chain(device, max_frames, max_payload),
        m_samples(b200_receiver_chain::alloc_samples(num_rx_samples ? num_rx_samples : 4096)),
        m_stop(false),
        m_done(false)
    {
        if (!m_chain.ok() || !m_samples || !m_callback || !m_source) { // inert, like a chain without a GPU
            m_done = true;
            return;
        }
        m_thread = std::thread(&b200_receiver::receiver_chain_loop, this); // receiver.cpp:33
    }

    b200_receiver::~b200_receiver()
    {
        stop();
        if (m_samples) b200_receiver_chain::free_samples(m_samples);
    }

    void b200_receiver::receiver_chain_loop()
    {
        while (!m_stop.load()) {
            std::vector<std::vector<unsigned char> > packets;
            bool more;
            {
                std::lock_guard<std::mutex> round(m_pause);          // sem_wait(&m_pause) ... sem_post(&m_pause)
                more = m_source(m_samples, m_n);                     // m_usrp.get_samples(NUM_RX_SAMPLES, m_samples)
                if (more) packets = m_chain.process_samples(m_samples, m_n);
                if (more) m_callback(packets);                       // called every round, also with no packets (receiver.cpp:53)
            }
            if (!more) break;
            std::this_thread::yield(); // gives a pause() caller waiting for the round to end its turn
        }
        {
            std::lock_guard<std::mutex> round(m_pause);              // a paused receiver hands nothing to the callback
            m_callback(m_chain.flush());                             // the payloads still in flight
        }
        {
            std::lock_guard<std::mutex> l(m_done_mu);
            m_done = true;
        }
        m_done_cv.notify_all();
    }

    void b200_receiver::pause() { m_pause.lock(); }

    void b200_receiver::resume() { m_pause.unlock(); }

    void b200_receiver::stop()
    {
        m_stop.store(true);
        if (m_thread.joinable()) m_thread.join();
    }

    void b200_receiver::wait()
    {
        std::unique_lock<std::mutex> l(m_done_mu);
        m_done_cv.wait(l, [this] { return m_done; });
    }
}
