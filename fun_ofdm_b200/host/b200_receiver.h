// fun::b200_receiver — the reference's fun::receiver (receiver.h:36-104, receiver.cpp:42-77) on the GPU chain.
//
// The reference's receiver owns a USRP, a receiver_chain and a thread that loops forever:
//     sem_wait(pause) -> usrp.get_samples(4096, samples) -> chain.process_samples(samples) -> callback(packets) -> sem_post(pause)
// and pause() / resume() take and release that semaphore from the user's thread.  This class is the same loop, the same
// callback signature and the same pause / resume contract, with two substitutions:
//   * the chain is fun::b200_receiver_chain (all six blocks on the GPU, payloads handed out up to max_lag calls late -
//     the reference's own chain hands them out up to five calls late);
//   * the radio is a sample source the caller supplies: a functor that fills a buffer of `n` samples and returns false
//     when the stream has ended (the UHD wrapper usrp::get_samples(n, buf), usrp.h:93, fits that shape; UHD itself and
//     the radio hardware are outside this repository).  The buffer handed to the source is pinned memory.
// Unlike the reference's thread, which is never joined (the process must be killed), the loop ends when the source
// returns false or stop() is called; the last payloads still in flight are flushed to the callback.
#ifndef B200_RECEIVER_H
#define B200_RECEIVER_H

#include "b200_receiver_chain.h"

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace fun
{
    class b200_receiver
    {
    public:
        typedef void (*callback_t)(std::vector<std::vector<unsigned char> > packets);                       // receiver.h:58
        typedef std::function<bool(std::complex<double> *buffer, size_t n)> source_t;                       // usrp.h:93

        // num_rx_samples: samples asked of the source per loop (receiver.h:16: NUM_RX_SAMPLES 4096)
        b200_receiver(callback_t callback, source_t source, size_t num_rx_samples = 4096, int device = 0,
                      unsigned max_frames = 1024, unsigned max_payload = 4095);
        ~b200_receiver();

        void pause();   // receiver.cpp:64-67: blocks until the loop is between two rounds, then holds it there
        void resume();  // receiver.cpp:74-77
        void stop();    // ends the loop (after the round in progress), flushes, joins the thread
        void wait();    // returns when the source has ended and everything has been delivered

        bool ok() const { return m_chain.ok(); }
        b200_receiver_chain::counters_t counters() const { return m_chain.counters(); }

    private:
        void receiver_chain_loop(); // receiver.cpp:42-58

        callback_t m_callback;
        source_t m_source;
        size_t m_n;
        b200_receiver_chain m_chain;
        std::complex<double> *m_samples; // pinned
        std::mutex m_pause;              // the reference's binary semaphore m_pause
        std::atomic<bool> m_stop;
        std::mutex m_done_mu;
        std::condition_variable m_done_cv;
        bool m_done;
        std::thread m_thread;
    };
}

#endif
