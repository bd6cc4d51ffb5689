// fun::b200_rx — the GPU receive block.
//
// Replaces, as ONE block, the reference's fft_symbols -> channel_est -> phase_tracker -> frame_decoder
// (receiver_chain.cpp:33-36, 47-50): input type of fft_symbols (fft_symbols.h:32), output type of
// frame_decoder (frame_decoder.h:80).  work() hands input_buffer - the tagged_sample stream timing_sync produces, 24-byte
// structs, exactly as they lie in the vector - to the GPU as a two-phase pass (include/b200rx.h, b200rx_pass_*,
// sample format B200RX_FMT_TAGGED_FC64): the structs are unpacked and the LTS1 tags found on the device (the host never
// walks the samples), every frame's SIGNAL symbol is decoded inside the call, and the frames whose last sample has
// arrived are decoded asynchronously on one of `depth` lanes.  Payloads of CRC-OK frames are appended to output_buffer by
// a later work() round in stream order - at most max_lag rounds later (the reference's four blocks deliver three rounds
// after the samples went in: one buffer swap per block, receiver_chain.cpp:118-125); flush() hands out the rest.
//
// Streaming state carried across work() calls: the samples of the frame still arriving (everything from its LTS1 tag on).
// Differences from the reference pipeline, by design: an LTS1 tag arriving inside a frame abandons that frame (in the
// reference its remaining symbols would be re-sliced and re-equalised against the new LTS and the CRC fails).
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
#include <string>
#include <vector>

struct b200rx_handle;

namespace fun
{
    class b200_rx : public fun::block<tagged_sample, std::vector<unsigned char> >
    {
    public:
        // device: CUDA device index; max_frames_per_call: capacity of one GPU pass;
        // max_payload: largest LENGTH decoded (longer frames are dropped); depth: passes in flight; max_lag: see above
        explicit b200_rx(int device = 0, unsigned max_frames_per_call = 256, unsigned max_payload = 4095, unsigned depth = 4,
                         unsigned max_lag = 3);
        virtual ~b200_rx();
        virtual void work();

        // Waits for everything still being decoded, appends it to output_buffer and drops the frame still arriving
        // (end of stream).
        void flush();

        struct counters_t { uint64_t frames_seen, headers_bad, frames_ok, frames_crc_fail, frames_abandoned; };
        counters_t counters() const { return m_counters; }
        bool ok() const { return m_handle != nullptr; }          // false: no GPU / library error at construction
        const std::string &error() const { return m_error; }

    private:
        struct pending_pass {
            uint64_t ticket, seq, round;
            std::vector<uint32_t> frames;
            std::vector<uint16_t> len;
        };
        void pass(const tagged_sample *fresh, size_t n);
        void collect(bool all);
        bool reserve(size_t n);

        b200rx_handle *m_handle;
        std::string m_error;
        unsigned m_max_frames, m_max_payload, m_depth, m_max_lag;
        tagged_sample *m_buf;          // pinned: the frame still arriving (from its LTS1 tag on) + the round's new samples
        size_t m_buf_n, m_buf_cap;
        uint64_t m_base;               // stream index of m_buf[0]
        int64_t m_last_lts1;           // stream index of the last frame settled
        uint64_t m_round, m_pass_seq;
        std::deque<pending_pass> m_inflight;
        std::vector<uint8_t *> m_payload, m_status; // per output slot, pinned
        std::vector<uint8_t> m_select;
        void *m_frames;                // b200rx_pass_frame[max_frames]
        counters_t m_counters;
    };
}

#endif
