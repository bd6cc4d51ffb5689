/* TEST INFRASTRUCTURE — not product code.
 *
 * Plain-C restatement ("port") of the reference receive hot path
 * fft_symbols -> channel_est -> phase_tracker -> frame_decoder (bmorgan5/fun_ofdm,
 * src/receiver_chain.cpp:33-36) and of the codec stage functions it calls.
 * Parity status: PINNED against the reference itself — tests/test_oracle_*.py
 * compare every function here with oracle/_ref/libfunref.so (the unmodified
 * reference sources compiled in this container) and with the committed golden
 * vectors under tests/golden/ (generated from that same library by
 * tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link
 * or load this file's library.
 */
#ifndef OFDM_ORACLE_H
#define OFDM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t hdr_ok;
    int32_t hdr_field;
    int32_t hdr_parity;
    int32_t rate_valid;
    int32_t rate;
    int32_t length;
    int32_t nsym;
    int32_t crc_ok;
    int32_t n_vectors;
    int32_t payload_from_blocks;
} orc_frame_info;

void orc_init(void);

void orc_conv_encode(const uint8_t *data, uint8_t *symbols, int data_bits);
void orc_conv_decode(const uint8_t *symbols, uint8_t *data, int data_bits);
int orc_puncture(const uint8_t *in, int n, int rate, uint8_t *out);
int orc_depuncture(const uint8_t *in, int n, int rate, uint8_t *out);
int orc_interleave(const uint8_t *in, int n, uint8_t *out);
int orc_deinterleave(const uint8_t *in, int n, uint8_t *out);
int orc_modulate(const uint8_t *bits, int n, int rate, double *out);
int orc_demodulate(const double *iq, int nsamp, int rate, uint8_t *out);
void orc_fft_forward(double *iq64);
uint32_t orc_crc32(const uint8_t *data, int n);
int orc_parity(int x);
int orc_decode_header(const double *iq48, int *rate, int *length, int *nsym);
int orc_decode_data(const double *iq, int nsamp, int rate, int length, uint8_t *payload_out);

int orc_decode_frame(const double *iq, int n_samples, orc_frame_info *info,
                     double *eq, int eq_cap_vectors,
                     uint8_t *soft, uint8_t *deint, uint8_t *depunct,
                     uint8_t *decoded, uint8_t *descrambled, uint8_t *payload);

double orc_decode_batch(const double *iq, const int64_t *lts1_off, const int32_t *n_avail, int n_frames,
                        uint8_t *payload_out, int payload_stride, int32_t *len_out, uint8_t *status_out,
                        int n_threads);

#ifdef __cplusplus
}
#endif

#endif
