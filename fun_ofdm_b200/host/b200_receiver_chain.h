// fun::b200_receiver_chain — the reference's receiver_chain with every block on the GPU.
//
// Same call as the reference (receiver_chain.h:56, receiver_chain.cpp:106-126):
//     std::vector<std::vector<unsigned char> > process_samples(std::vector<std::complex<double> > samples);
// raw base-band samples in, CRC-valid payloads (MPDUs) out, in the order their frames start in the stream.  Where the
// reference pushes the samples through six block threads (frame_detector, timing_sync, fft_symbols, channel_est,
// phase_tracker, frame_decoder), this class hands them to b200rx_receive (include/b200rx.h): detection, synchronisation
// and decoding of a whole capture in one GPU pass.
//
// Streaming state carried across calls (what the reference keeps inside its blocks: frame_detector's and timing_sync's
// carry-over buffers, timing_sync's m_phase_acc, fft_symbols' partial vector, frame_decoder's frame in progress):
//   * the tail of the stream that may still matter - the last 672 samples (an STS_END tag in the last 160 samples is
//     left for the next call exactly like timing_sync.cpp:68, and the detector needs 48 samples of history), or
//     everything from 224 samples before the LTS1 tag of a frame whose samples have not all arrived yet;
//   * m_phase_acc of the last synchronised frame;
//   * the stream position below which frames have already been delivered (re-examined frames are recognised by the
//     absolute index of their LTS1 tag);
//   * where the caller cut the stream: the reference's timing_sync discards an LTS that starts before the 160 samples it
//     carried over into the current work() buffer (timing_sync.cpp:102), so its output depends on the chunking; the
//     boundaries are handed to the GPU pass (b200rx_set_receive_origins) and the same frames are dropped.
// A frame is delivered in the call that brings its last sample; the reference delivers it up to five calls later
// (one per block still in front of the payload).  Sequences of payloads are identical; per-call alignment is not.
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
    class b200_receiver_chain
    {
    public:
        // device: CUDA device index; max_frames: most frames one call may contain (device scratch is sized for it);
        // max_payload: largest LENGTH decoded (longer frames are dropped, status TOO_LONG)
        explicit b200_receiver_chain(int device = 0, unsigned max_frames = 1024, unsigned max_payload = 4095);
        ~b200_receiver_chain();

        std::vector<std::vector<unsigned char> > process_samples(std::vector<std::complex<double> > samples);
        // same, without the by-value vector (one host copy less per call)
        std::vector<std::vector<unsigned char> > process_samples(const std::complex<double> *samples, size_t n);

        // End of stream: pushes `pad` zero samples through (the reference needs trailing samples just the same:
        // test_sim.cpp:75-77 pads its stream with zeros) and drops whatever is still incomplete.
        std::vector<std::vector<unsigned char> > flush(unsigned pad = 1024);

        struct counters_t { uint64_t samples, calls, frames_found, frames_ok, frames_crc_fail, headers_bad, frames_truncated; };
        counters_t counters() const { return m_counters; }
        bool ok() const { return m_handle != nullptr; } // false: no GPU / library error at construction
        const std::string &error() const { return m_error; }

    private:
        void run_capture(std::vector<std::vector<unsigned char> > &out);

        b200rx_handle *m_handle;
        std::string m_error;
        unsigned m_max_frames, m_max_payload;
        // retained tail of the stream + the new samples, in pinned host memory (the H2D copy of a pageable std::vector
        // would be staged by the driver at a fraction of the link rate)
        std::complex<double> *m_buf;
        size_t m_buf_n, m_buf_cap;
        bool reserve(size_t n);
        uint64_t m_base;                           // stream index of m_buf[0]
        uint64_t m_handled;                        // STS_END tags below this stream index have been examined
        int64_t m_last_lts1;                       // LTS1 index of the last frame delivered or dropped for good (-1: none)
        std::deque<uint64_t> m_calls;              // stream index at which each recent process_samples() call started
        int64_t m_pending_lts1;                    // LTS1 index of the frame still arriving (-1: none)
        double m_phase;                            // timing_sync's m_phase_acc in front of m_buf
        counters_t m_counters;
        // output staging reused across calls
        std::vector<uint8_t> m_payload, m_rate, m_status;
        std::vector<uint16_t> m_len;
        std::vector<uint64_t> m_lts1;
    };
}

#endif
