// fun::b200_rx — the GPU receive block.
//
// Replaces, as ONE block, the reference's fft_symbols -> channel_est -> phase_tracker -> frame_decoder
// (receiver_chain.cpp:33-36, 47-50): input type of fft_symbols (fft_symbols.h:32), output type of
// frame_decoder (frame_decoder.h:80).  work() consumes input_buffer (the tagged_sample stream
// timing_sync produces), cuts it into frames at the LTS1 tags, learns each frame's length from a batched
// SIGNAL decode on the GPU as soon as 208 samples are in, and when a frame's last sample has arrived
// decodes all completed frames in one b200rx_decode_batch call.  Payloads of CRC-OK frames are appended
// to output_buffer in completion order, which is the order frame_decoder would emit them.
//
// Streaming state carried across work() calls: the samples of frames still arriving.  Differences from the
// reference pipeline, by design: frames surface up to three process_samples() rounds earlier (one block
// instead of four); an LTS1 tag arriving inside a frame abandons that frame (in the reference its remaining
// symbols would be re-sliced and re-equalised against the new LTS and the CRC fails).
#ifndef B200_RX_BLOCK_H
#define B200_RX_BLOCK_H

#ifdef B200_USE_REFERENCE_HEADERS
#include "block.h"
#include "tagged_vector.h"
#else
#include "fun_api.h"
#endif

#include <cstdint>
#include <deque>
#include <vector>

struct b200rx_handle;

namespace fun
{
    class b200_rx : public fun::block<tagged_sample, std::vector<unsigned char> >
    {
    public:
        // device: CUDA device index; max_frames_per_call: capacity of one GPU batch;
        // max_payload: largest LENGTH decoded (longer frames are dropped)
        explicit b200_rx(int device = 0, unsigned max_frames_per_call = 256, unsigned max_payload = 4095);
        virtual ~b200_rx();
        virtual void work();

        // Decode whatever is complete and drop partial frames (end of stream).
        void flush();

        struct counters_t { uint64_t frames_seen, headers_bad, frames_ok, frames_crc_fail, frames_abandoned; };
        counters_t counters() const { return m_counters; }
        bool ok() const { return m_handle != nullptr; }          // false: no GPU / library error at construction
        const std::string &error() const { return m_error; }

    private:
        struct capture {
            std::vector<std::complex<double> > samples; // from the LTS1-tagged sample on
            int needed;                                  // -1 until the header is known
            bool header_tried;
        };
        void decode_headers();
        void decode_ready();

        b200rx_handle *m_handle;
        std::string m_error;
        unsigned m_max_frames, m_max_payload;
        std::deque<capture> m_open;   // frames still arriving, in stream order (normally 0 or 1)
        std::vector<capture> m_ready; // complete frames awaiting the batch decode
        counters_t m_counters;
    };
}

#endif
