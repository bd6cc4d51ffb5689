// fun::b200_receiver_chain — the reference's receiver_chain with every block on the GPU.
//
// Same call as the reference (receiver_chain.h:56, receiver_chain.cpp:106-126):
//     std::vector<std::vector<unsigned char> > process_samples(std::vector<std::complex<double> > samples);
// raw base-band samples in, CRC-valid payloads (MPDUs) out, in the order their frames start in the stream.  Where the
// reference pushes the samples through six block threads (frame_detector, timing_sync, fft_symbols, channel_est,
// phase_tracker, frame_decoder), this class hands them to the GPU as two-phase passes (include/b200rx.h, b200rx_pass_*):
//
//   * phase one, inside the call: the samples cross PCIe, frame detection + timing synchronisation + SIGNAL decode run
//     over the retained tail of the stream plus the new samples, and the frame list comes back (tens of microseconds).
//     Everything the streaming state depends on is known at that point: which frames exist, where they start, how long
//     they are, whether their last sample has arrived, m_phase_acc.
//   * phase two, asynchronous: the complete new frames are decoded (data symbols, Viterbi, descrambler, CRC) on one of
//     `depth` lanes while the caller is already preparing its next chunk.  Their payloads are handed out by a LATER
//     call - like the reference, whose payloads surface up to five calls after the samples went in
//     (receiver_chain.cpp:118-125: one buffer swap per block and call).  `max_lag` bounds that delay in calls (default 5,
//     the reference's own; a frame older than that is waited for), flush() hands out everything still in flight.
//
// Streaming state carried across calls (what the reference keeps inside its blocks: frame_detector's and timing_sync's
// carry-over buffers, timing_sync's m_phase_acc, fft_symbols' partial vector, frame_decoder's frame in progress):
//   * the tail of the stream that may still matter - the last 672 samples (an STS_END tag in the last 160 samples is
//     left for the next call exactly like timing_sync.cpp:68, and the detector needs 48 samples of history), or
//     everything from 512 samples before the LTS1 tag of a frame whose samples have not all arrived yet;
//   * m_phase_acc of the last synchronised frame;
//   * the stream position below which frames have already been settled (re-examined frames are recognised by the
//     absolute index of their LTS1 tag);
//   * where the caller cut the stream: the reference's timing_sync discards an LTS that starts before the 160 samples it
//     carried over into the current work() buffer (timing_sync.cpp:102), so its output depends on the chunking; the
//     boundaries are handed to the GPU pass (b200rx_set_receive_origins) and the same frames are dropped.
// Sequences of payloads are identical to the reference's; per-call alignment is not.
#ifndef B200_RECEIVER_CHAIN_H
#define B200_RECEIVER_CHAIN_H

#include <complex>
#include <cstdint>
#include <deque>
#include <string>
#include <vector>

struct b200rx_handle;

namespace fun
{
    class b200_copy_pool; // host threads that move a large call's samples into pinned memory in parallel

    class b200_receiver_chain
    {
    public:
        // device: CUDA device index; max_frames: most frames one GPU pass may contain (device scratch is sized for it);
        // max_payload: largest LENGTH decoded (longer frames are dropped, status TOO_LONG);
        // depth: passes in flight (1 = every call waits for its own frames, like round 1); max_lag: see above
        explicit b200_receiver_chain(int device = 0, unsigned max_frames = 1024, unsigned max_payload = 4095,
                                     unsigned depth = 6, unsigned max_lag = 5);
        ~b200_receiver_chain();

        std::vector<std::vector<unsigned char> > process_samples(std::vector<std::complex<double> > samples);
        // same, without the by-value vector (one host copy less per call).  If `samples` lies in pinned memory
        // (alloc_samples) a long call is copied to the GPU straight from it.
        std::vector<std::vector<unsigned char> > process_samples(const std::complex<double> *samples, size_t n);

        // End of stream: pushes `pad` zero samples through (the reference needs trailing samples just the same:
        // test_sim.cpp:75-77 pads its stream with zeros), waits for every frame still being decoded and drops whatever
        // is still incomplete.
        std::vector<std::vector<unsigned char> > flush(unsigned pad = 1024);

        // Pinned sample buffers for callers that fill them directly (e.g. a radio driver's receive buffer).
        static std::complex<double> *alloc_samples(size_t n);
        static void free_samples(std::complex<double> *p);

        void set_max_lag(unsigned calls) { m_max_lag = calls; }
        int set_tuning(const char *key, long long value); // b200rx_set_tuning on the chain's handle (never changes results)
        void set_copy_threads(unsigned n); // host threads used to stage calls of >= 64 Ki samples (default 4; 1 = caller's thread only)

        struct counters_t { uint64_t samples, calls, frames_found, frames_ok, frames_crc_fail, headers_bad, frames_truncated; };
        counters_t counters() const { return m_counters; }
        bool ok() const { return m_handle != nullptr; } // false: no GPU / library error at construction
        const std::string &error() const { return m_error; }

    private:
        struct pending_pass {            // a pass whose frames are still being decoded
            uint64_t ticket, seq, call;  // b200rx ticket; pass number (output slot = seq % depth); call that submitted it
            std::vector<uint32_t> frames; // slots of the selected frames within the pass, in stream order
            std::vector<uint16_t> len;    // their LENGTH fields
        };
        void run_capture(const std::complex<double> *direct, size_t n_direct, std::vector<std::vector<unsigned char> > &out);
        void collect(std::vector<std::vector<unsigned char> > &out, bool all);
        void deliver(const pending_pass &p, std::vector<std::vector<unsigned char> > &out);
        void stage(const std::complex<double> *src, size_t n);

        b200rx_handle *m_handle;
        std::string m_error;
        unsigned m_max_frames, m_max_payload, m_depth, m_max_lag;
        // retained tail of the stream (+ the new samples of a call that is staged through it), in pinned host memory (the
        // H2D copy of a pageable std::vector would be staged by the driver at a fraction of the link rate)
        std::complex<double> *m_buf;
        size_t m_buf_n, m_buf_cap;
        bool reserve(size_t n);
        uint64_t m_base;                           // stream index of m_buf[0]
        uint64_t m_handled;                        // STS_END tags below this stream index have been examined
        uint64_t m_stream_start;                   // stream index at which the current stream began (0, or where flush() cut)
        int64_t m_last_lts1;                       // LTS1 index of the last frame settled (-1: none)
        std::deque<uint64_t> m_calls;              // stream index at which each recent process_samples() call started
        int64_t m_pending_lts1;                    // LTS1 index of the frame still arriving (-1: none)
        double m_phase;                            // timing_sync's m_phase_acc in front of m_buf
        counters_t m_counters;
        // passes
        uint64_t m_pass_seq;
        std::deque<pending_pass> m_inflight;
        std::vector<uint8_t *> m_payload, m_status; // per output slot, pinned
        std::vector<uint8_t> m_select;
        void *m_frames;                             // b200rx_pass_frame[max_frames]
        b200_copy_pool *m_pool;
    };
}

#endif
