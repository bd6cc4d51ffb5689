/* b200tx — C ABI of the synthetic-corpus generator (host side).
 *
 * Benchmark input for the receive hot path: 802.11a-like frames exactly as the reference's
 * frame_builder::build_frame produces them (src/frame_builder.cpp:53-82: ppdu::encode ->
 * symbol_mapper::map -> fft::inverse -> cyclic prefix -> 320-sample preamble), followed by a
 * seeded channel (the reference has none: examples/test_sim.cpp is noiseless): optional
 * exponential-profile multipath FIR (<= 8 taps) and complex AWGN at a given SNR.
 *
 * This is the "f3" row of SURVEY.md section 8 on the host; coded bits are bit-exact with the
 * reference, samples agree to ~1e-15 (tests/test_txgen.py pins both against oracle/_ref).
 * Library: fun_ofdm_b200/lib/libb200host.so.
 */
#ifndef B200TX_H
#define B200TX_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200TX_API __attribute__((visibility("default")))
#else
#define B200TX_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Number of complex samples of a frame: 320 + 80 * (1 + nsym); negative on a bad rate/length. */
B200TX_API int b200tx_frame_samples(int rate, int length);

/* Data OFDM symbols (ppdu.cpp:38-40). */
B200TX_API int b200tx_num_symbols(int rate, int length);

/* Replaces frame_builder::build_frame(payload, rate) (frame_builder.cpp:53-82).  iq_out receives
 * b200tx_frame_samples() interleaved (re, im) doubles.  Returns the sample count or a negative code. */
B200TX_API int b200tx_build_frame(const uint8_t *payload, int length, int rate, double *iq_out);

/* The 48 * (1 + nsym) constellation points of ppdu::encode (ppdu.cpp:65-73), for tests. */
B200TX_API int b200tx_ppdu_encode(const uint8_t *payload, int length, int rate, double *out);

/* The 320 preamble samples this generator emits (computed from IEEE 802.11a 17.3.3, with the two
 * window-edge samples the reference's table carries: preamble.h:24-359). */
B200TX_API void b200tx_preamble(double *iq_out_320);

typedef struct {
    double snr_db;        /* AWGN: per-component sigma = sqrt(P / 10^(snr/10) / 2), P = mean |x|^2 of the
                             frame after its 320-sample preamble; >= 200 disables noise */
    uint32_t multipath_taps; /* 0 or 1: none; 2..8: random complex FIR, exponential power profile, unit energy */
    uint32_t lead_in;     /* noise-only samples written before each frame */
    uint64_t seed;        /* channel seed (payloads are the caller's) */
    uint32_t n_threads;   /* host threads */
    uint32_t reserved;
} b200tx_channel;

/* Build n_frames frames into one stream.  Frame f: payload bytes at payloads + payload_off[f],
 * lengths[f] bytes, rates[f]; written at iq_out + 2 * out_off[f] (complex-sample offsets) preceded by
 * ch->lead_in noise samples, i.e. the frame itself starts at out_off[f] + lead_in and occupies
 * b200tx_frame_samples() samples.  Deterministic in (seed, f) regardless of n_threads. */
B200TX_API int b200tx_build_batch(const uint8_t *payloads, const uint64_t *payload_off, const uint32_t *lengths,
                                  const uint8_t *rates, uint32_t n_frames, double *iq_out, const uint64_t *out_off,
                                  const b200tx_channel *ch);

/* The same generator on the GPU: every pointer except `ch` is a DEVICE pointer, the stream is written
 * straight into HBM (fun_ofdm_b200/csrc/txgen.cu, exported by libb200rx.so - there is no host implementation behind
 * this entry point).  Same layout, same seeds, same counter-based noise as b200tx_build_batch: coded bits are identical,
 * samples agree to ~1e-15.  Asynchronous on `cuda_stream` (a cudaStream_t, may be NULL); ch->n_threads is ignored.
 * Returns 0, -1 (bad argument) or -2 (CUDA error). */
#if defined(__GNUC__)
#define B200TX_DEV_API __attribute__((visibility("default")))
#else
#define B200TX_DEV_API
#endif
B200TX_DEV_API int b200tx_build_batch_dev(int device, void *cuda_stream, const uint8_t *payloads_dev,
                                          const uint64_t *payload_off_dev, const uint32_t *lengths_dev,
                                          const uint8_t *rates_dev, uint32_t n_frames, double *iq_out_dev,
                                          const uint64_t *out_off_dev, const b200tx_channel *ch);

#ifdef __cplusplus
}
#endif

#endif /* B200TX_H */
