// Warp-level add-compare-select core of the K=7, 64-state Viterbi decoder, bit-exact with the
// reference's Spiral SSE2 kernel (src/viterbi.cpp:208-459) including its quirks:
//   * path metrics are unsigned bytes with SATURATING adds (paddusb, :252-255),
//   * ties choose the predecessor j+32 (pminub/pcmpeqb against the j+32 candidate, :256-259),
//   * renormalisation (subtract the minimum) happens only when the metric of STATE 0 exceeds 210
//     (:314-332),
//   * initial metrics: state 0 -> 0, all others 63 (:71-78).
//
// Layout (one warp per frame): lane l owns butterfly l, i.e. old states l and l+32, packed as two
// 16-bit halves of one register (values stay <= 255, so u16x2 DPX instructions implement the u8
// saturating arithmetic exactly: VIADDMNMX.U16x2 = min(x + m, 255)).  The butterfly produces new
// states 2l and 2l+1; two shuffles bring states l and l+32 back for the next step.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200rx {

constexpr unsigned VIT_FULL = 0xFFFFFFFFu;

struct AcsLane {
    uint32_t cls_sel;  // PRMT selector extracting this lane's branch-class byte from a metric word
    uint32_t half_sel; // PRMT selector gathering (A.half, B.half) for the next step
    int src_a, src_b;  // lanes holding new states l and l+32
};

__device__ __forceinline__ AcsLane acs_lane_init(int lane)
{
    AcsLane s;
    const uint32_t j = (uint32_t)lane;
    const uint32_t b0 = ((j >> 2) ^ (j >> 3) ^ (j >> 4)) & 1u; // parity(2j & 121)
    const uint32_t b1 = (j ^ (j >> 2) ^ (j >> 3)) & 1u;        // parity(2j & 91)
    s.cls_sel = 0x4440u | ((b0 << 1) | b1);
    s.half_sel = (lane & 1) ? 0x7632u : 0x5410u;
    s.src_a = lane >> 1;
    s.src_b = 16 + (lane >> 1);
    return s;
}

__device__ __forceinline__ uint32_t acs_initial_metrics(int lane)
{
    return (lane == 0 ? 0u : 63u) | (63u << 16); // (X[l], X[l+32])
}

// One trellis step.  R = (X[l], X[l+32]); w = branch-metric word of the step.
// Returns (Y[2l], Y[2l+1]) after the reference's conditional renormalisation and sets the two
// decision ballots (bit l of d_even = decision of new state 2l, of d_odd = new state 2l+1).
__device__ __forceinline__ uint32_t acs_step(uint32_t R, uint32_t w, const AcsLane &ln, int lane,
                                             uint32_t &d_even, uint32_t &d_odd)
{
    const uint32_t m = __byte_perm(w, 0u, ln.cls_sel); // 0..63
    const uint32_t mi = m ^ 63u;                       // 63 - m (psubusb, viterbi.cpp:246-248)
    const uint32_t P = __viaddmin_u16x2(R, m | (mi << 16), 0x00FF00FFu); // (X[l]+m , X[l+32]+mi)
    const uint32_t Q = __viaddmin_u16x2(R, mi | (m << 16), 0x00FF00FFu); // (X[l]+mi, X[l+32]+m )
    const uint32_t U = __byte_perm(P, Q, 0x5410u); // candidates through predecessor l
    const uint32_t V = __byte_perm(P, Q, 0x7632u); // candidates through predecessor l+32
    const uint32_t D = U + 0x01000100u - V;        // bit 8 / 24 set <=> V <= U (ties -> l+32)
    uint32_t Y = __vminu2(U, V);
    d_even = __ballot_sync(VIT_FULL, (D & 0x00000100u) != 0u);
    d_odd = __ballot_sync(VIT_FULL, (D & 0x01000000u) != 0u);
    // renormalise iff Y[0] > 210 (state 0 = low half of lane 0)
    if (__any_sync(VIT_FULL, lane == 0 && (Y & 0xFFFFu) > 210u)) {
        const uint32_t mn = __reduce_min_sync(VIT_FULL, min(Y & 0xFFFFu, Y >> 16));
        Y -= mn * 0x00010001u;
    }
    return Y;
}

__device__ __forceinline__ uint32_t acs_next(uint32_t Y, const AcsLane &ln)
{
    const uint32_t A = __shfl_sync(VIT_FULL, Y, ln.src_a);
    const uint32_t B = __shfl_sync(VIT_FULL, Y, ln.src_b);
    return __byte_perm(A, B, ln.half_sel);
}

// Whole decode of a short block (<= 32 steps) by one warp, metric words in shared memory.
// Returns the first 32 decoded bits MSB-first packed as bytes b0<<16 | b1<<8 | b2 (3 bytes),
// as viterbi_chainback writes them (viterbi.cpp:131-142).  Used for the SIGNAL field
// (ppdu.cpp:181: 18 data bits, 24 steps).
__device__ __forceinline__ uint32_t warp_viterbi_short(const uint32_t *bm, int n_steps, int data_bits, int lane)
{
    const AcsLane ln = acs_lane_init(lane);
    uint32_t R = acs_initial_metrics(lane);
    uint32_t my_e = 0, my_o = 0;
    for (int t = 0; t < n_steps; t++) {
        uint32_t de, dod;
        const uint32_t Y = acs_step(R, bm[t], ln, lane, de, dod);
        if (lane == t) { my_e = de; my_o = dod; }
        R = acs_next(Y, ln);
    }
    uint32_t e = 0, bytes = 0;
    for (int n = data_bits - 1; n >= 0; n--) {
        const uint32_t st = e >> 2;
        const uint32_t we = __shfl_sync(VIT_FULL, my_e, n + 6);
        const uint32_t wo = __shfl_sync(VIT_FULL, my_o, n + 6);
        const uint32_t k = (((st & 1u) ? wo : we) >> (st >> 1)) & 1u;
        e = (e >> 1) | (k << 7);
        const int byte = n >> 3; // data[n >> 3] = e, rewritten every step
        if (byte < 3) bytes = (bytes & ~(0xFFu << (16 - 8 * byte))) | (e << (16 - 8 * byte));
    }
    return bytes;
}

} // namespace b200rx
