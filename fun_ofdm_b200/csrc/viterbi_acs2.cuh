// Viterbi add-compare-select, second generation: 8 lanes per frame, 8 states per lane, IN PLACE.
//
// Semantics are the reference's Spiral SSE2 kernel exactly (src/viterbi.cpp:208-459, see viterbi_core.cuh
// for the list of quirks: saturating u8 metrics, ties -> predecessor j+32, renormalise only when the
// metric of state 0 exceeds 210, initial metrics 0/63).
//
// Why this layout.  ncu on the first (warp-per-frame) kernel showed the ALU pipe 87 % busy with ~35
// warp-instructions per trellis step, most of them not arithmetic on path metrics but plumbing: two
// shuffles + a PRMT to re-pair states every step, two ballots to collect decisions, metric selection.
// Here the 64 path metrics never move between steps.  A butterfly maps old states (j, j+32) to new
// states (2j, 2j+1) = (rotl6(j), rotl6(j+32)); if the result is written back where the inputs were,
// position p simply holds state rotl6^t(p) at time t.  The butterfly partner of a position is then the
// position differing in bit (5 - t mod 6).  With position = [lane:3 | register:2 | half:1] that axis is
//     t mod 6 = 0,1,2 : a lane bit      -> one shuffle per register, candidates exchanged pre-saturated
//     t mod 6 = 3,4   : a register bit  -> pure register arithmetic, 6 DPX ops per 4 states
//     t mod 6 = 5     : the half bit    -> 2 PRMT + 2 DPX ops per 2 states
// Arithmetic is u16x2 DPX: VIADDMNMX.U16x2 does min(x + m, cap) (the saturating add, and with
// cap = the other candidate the add + compare-select in one instruction).  The decision bit is
// bit 8 / 24 of (survivor + 0x0100 - candidate_via_j+32) (set iff equal, i.e. iff j+32 won or tied) and is
// computed with two IMADs on the otherwise idle FMA pipe where possible.  Decisions are accumulated over
// 8 steps per position (acc = 2*acc + bits, one IMAD per 4 states) and stored as 64 B per frame per 8 steps.
//
// Survivor layout in HBM (per frame, uint32 words):  word[(t >> 3) * 16 + lane * 2 + (reg >> 1)],
// byte (reg & 1) * 2 + (1 - half_is_high... see acs2_decision_bit), bit 7 - (t & 7).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200rx {

constexpr int ACS2_LB = 3;                 // lane bits per frame
constexpr int ACS2_T = 1 << ACS2_LB;       // lanes per frame
constexpr int ACS2_NR = 32 >> ACS2_LB;     // u16x2 registers per lane
constexpr int ACS2_FPW = 32 / ACS2_T;      // frames per warp
constexpr int ACS2_BLK = 24;               // steps per unrolled block = lcm(6 phases, 8-step store period)
constexpr int ACS2_WORDS_PER_8 = ACS2_T * ACS2_NR / 2; // survivor words per frame per 8 steps (16)

__host__ __device__ __forceinline__ uint32_t acs2_rotl6(uint32_t x, int r)
{
    r %= 6;
    return ((x << r) | (x >> (6 - r))) & 63u;
}

__host__ __device__ __forceinline__ uint32_t acs2_class(uint32_t j)
{
    const uint32_t b0 = ((j >> 2) ^ (j >> 3) ^ (j >> 4)) & 1u; // parity(2j & 121)
    const uint32_t b1 = (j ^ (j >> 2) ^ (j >> 3)) & 1u;        // parity(2j & 91)
    return (b0 << 1) | b1;
}

// Decision of new state `state` produced by trellis step t (1 = survivor came from predecessor j+32).
__device__ __forceinline__ uint32_t acs2_decision_bit(const uint32_t *dec, uint32_t t, uint32_t state)
{
    const uint32_t r = (t + 1u) % 6u;
    const uint32_t p = ((state >> r) | (state << (6u - r))) & 63u; // rotr6^r
    const uint32_t lane = p >> (6 - ACS2_LB), reg = (p >> 1) & (ACS2_NR - 1), low = p & 1u;
    const uint32_t w = __ldg(dec + (size_t)(t >> 3) * ACS2_WORDS_PER_8 + lane * (ACS2_NR / 2) + (reg >> 1));
    const uint32_t byte = (reg & 1u) * 2u + (low ^ 1u); // byte 0/2 = low half (half bit 1), byte 1/3 = high half
    return (w >> (byte * 8u + (7u - (t & 7u)))) & 1u;
}

// PTX prmt in its default mode: selector nibble bit 3 replicates the sign bit of the selected byte, which
// turns a byte known to be < 128 into a zero byte.  (__byte_perm masks the selector to 3 bits per nibble.)
__device__ __forceinline__ uint32_t acs2_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// a * b + c forced onto the FMA pipe (IMAD); with b = -1 this is c - a.  The ALU pipe is the bottleneck
// of this kernel (ncu: 61 % busy vs 9 % for the FMA pipe), so every op that can move does.
__device__ __forceinline__ uint32_t acs2_fma(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

struct Acs2Lane {
    uint32_t sel[6][ACS2_NR]; // PRMT selector (hi class, lo class) per phase and register
    uint32_t selB[ACS2_NR];   // second selector of the half-bit phase
    int sgn[ACS2_LB], nsgn[ACS2_LB];
    uint32_t neg1, one;       // 0xFFFFFFFF and 1 derived from a kernel argument so that ptxas keeps the IMADs
};

__device__ __forceinline__ void acs2_lane_init(Acs2Lane &L, int glane, uint32_t neg1)
{
    L.neg1 = neg1;
    L.one = 0u - neg1;
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int i = 0; i < ACS2_NR; i++) {
            const uint32_t p_hi = ((uint32_t)glane << (6 - ACS2_LB)) | ((uint32_t)i << 1);
            const uint32_t c_hi = acs2_class(acs2_rotl6(p_hi, r) & 31u);
            const uint32_t c_lo = acs2_class(acs2_rotl6(p_hi | 1u, r) & 31u);
            if (r < 5) L.sel[r][i] = 0x8080u | (c_hi << 8) | c_lo;
            else {
                L.sel[r][i] = 0x8080u | (c_hi << 8) | (4u + c_hi);   // (m high, 63-m low)
                L.selB[i] = 0x8080u | ((4u + c_hi) << 8) | c_hi;     // (63-m high, m low)
            }
        }
    }
#pragma unroll
    for (int k = 0; k < ACS2_LB; k++) {
        // phase r = k uses lane bit (LB-1-k); the lane holding predecessor j (bit 0) decides with
        // "partner <= own", the lane holding j+32 with "own <= partner"
        const int bit = (glane >> (ACS2_LB - 1 - k)) & 1;
        L.sgn[k] = bit ? -1 : 1;
        L.nsgn[k] = bit ? 1 : -1;
    }
}

constexpr uint32_t ACS2_CAP = 0x00FF00FFu;
constexpr uint32_t ACS2_C = 0x01000100u;

// One trellis step of phase PH (= t mod 6).  R: path metrics; w: the step's branch-metric word;
// D: raw decision words (bit 8 / 24 valid, bits 9-15 / 25-31 zero).
template <int PH>
__device__ __forceinline__ void acs2_step(uint32_t (&R)[ACS2_NR], uint32_t (&D)[ACS2_NR], uint32_t w, const Acs2Lane &L)
{
    constexpr int axis = 5 - PH;         // position bit separating the butterfly partners
    if constexpr (axis >= 6 - ACS2_LB) {
        // ---- partners in another lane ----
        constexpr int lbit = axis - (6 - ACS2_LB);
        constexpr int k = ACS2_LB - 1 - lbit; // index into sgn[]
        static_assert(k == PH, "lane phases come first");
#pragma unroll
        for (int i = 0; i < ACS2_NR; i++) {
            const uint32_t M = acs2_prmt(w, w, L.sel[PH][i]);
            const uint32_t Mi = acs2_fma(M, L.neg1, 0x003F003Fu);           // 63 - m per half, on the FMA pipe
            const uint32_t G = __viaddmin_u16x2(R[i], Mi, ACS2_CAP);            // my candidate for the partner's new state
            const uint32_t S = __shfl_xor_sync(0xFFFFFFFFu, G, 1 << lbit);      // partner's candidate for mine
            const uint32_t V = __viaddmin_u16x2(R[i], M, ACS2_CAP);             // my own candidate
            R[i] = __vminu2(V, S);
            D[i] = (uint32_t)((int)V * L.sgn[k] + ((int)S * L.nsgn[k] + (int)ACS2_C));
        }
    } else if constexpr (axis >= 1) {
        // ---- partners in another register of this lane ----
        constexpr int q = axis - 1;
#pragma unroll
        for (int a = 0; a < ACS2_NR; a++) {
            if ((a >> q) & 1) continue;
            const int b = a | (1 << q);
            const uint32_t M = acs2_prmt(w, w, L.sel[PH][a]);
            const uint32_t Mi = acs2_fma(M, L.neg1, 0x003F003Fu);
            const uint32_t B = __viaddmin_u16x2(R[b], Mi, ACS2_CAP); // via j+32 -> new state 2j
            const uint32_t E = __viaddmin_u16x2(R[b], M, ACS2_CAP);  // via j+32 -> new state 2j+1
            const uint32_t Ya = __viaddmin_u16x2(R[a], M, B);        // min(X[j] + m, B); ties keep B's value
            const uint32_t Yb = __viaddmin_u16x2(R[a], Mi, E);
            D[a] = acs2_fma(Ya, L.one, acs2_fma(B, L.neg1, ACS2_C));
            D[b] = acs2_fma(Yb, L.one, acs2_fma(E, L.neg1, ACS2_C));
            R[a] = Ya;
            R[b] = Yb;
        }
    } else {
        // ---- partners are the two halves of one register: high = state j, low = state j+32 ----
        const uint32_t wc = w ^ 0x3F3F3F3Fu; // 63 - m per byte (psubusb 63, m: viterbi.cpp:246-248)
#pragma unroll
        for (int i = 0; i < ACS2_NR; i++) {
            const uint32_t W = __byte_perm(R[i], R[i], 0x3232u); // (X[j], X[j])
            const uint32_t Z = __byte_perm(R[i], R[i], 0x1010u); // (X[j+32], X[j+32])
            const uint32_t MA = acs2_prmt(w, wc, L.sel[PH][i]); // (m, 63-m)
            const uint32_t MB = acs2_prmt(w, wc, L.selB[i]);    // (63-m, m)
            const uint32_t V = __viaddmin_u16x2(Z, MB, ACS2_CAP); // via j+32: (-> 2j, -> 2j+1)
            const uint32_t Y = __viaddmin_u16x2(W, MA, V);
            D[i] = acs2_fma(Y, L.one, acs2_fma(V, L.neg1, ACS2_C));
            R[i] = Y;
        }
    }
}

// Reference renormalisation (viterbi.cpp:314-332): if metric of state 0 > 210, subtract the minimum of
// all 64 metrics.  State 0 always sits in the high half of register 0 of the frame's lane 0.
__device__ __forceinline__ void acs2_renorm(uint32_t (&R)[ACS2_NR], int glane, int group)
{
    const bool hot = (glane == 0) && (R[0] > 0x00D2FFFFu);
    if (__any_sync(0xFFFFFFFFu, hot)) {
        const uint32_t any = __ballot_sync(0xFFFFFFFFu, hot);
        uint32_t m = R[0];
#pragma unroll
        for (int i = 1; i < ACS2_NR; i++) m = __vminu2(m, R[i]);
#pragma unroll
        for (int s = 1; s < ACS2_T; s <<= 1) m = __vminu2(m, __shfl_xor_sync(0xFFFFFFFFu, m, s));
        m = __vminu2(m, __byte_perm(m, m, 0x1032u)); // both halves = min over all 64 states
        const uint32_t sub = ((any >> (group * ACS2_T)) & 1u) ? m : 0u;
#pragma unroll
        for (int i = 0; i < ACS2_NR; i++) R[i] -= sub;
    }
}

} // namespace b200rx
