// Viterbi add-compare-select, third generation: TWO FRAMES per u16x2 register, 64 >> LB states per lane, in place.
//
// Semantics are the reference's Spiral SSE2 kernel exactly (src/viterbi.cpp:208-459; quirks listed in viterbi_core.cuh).
//
// What changed against viterbi_acs2.cuh, and why (profiles/r02_ubench_pipes.txt: ALU pipe - DPX, PRMT, LOP3, HSET2 - and
// FMA pipe - IMAD - both take one warp instruction per 2 cycles per scheduler, SHFL one per 4; the scheduler issues 1 per
// cycle, so the kernel is as fast as its instruction count allows only while ALU and FMA instructions are balanced):
//   * the two 16-bit halves of a register no longer hold two states of one frame but the SAME state of two frames
//     (A high, B low).  The position of a state is then [lane:LB | register:6-LB]: the butterfly partner is never
//     "the other half of the register", so the phase that cost 6 instructions per register (4 PRMT + 2 DPX) is gone;
//     with LB = 2 four of the six phases are pure register phases (2 DPX + 2 IMAD per state register).
//   * one set of branch metrics (4 PRMT + 4 IMAD per step) and one renormalisation test now serve 2 << (5 - LB) frames
//     per warp instead of half as many, and every lane carries 64 >> LB independent dependency chains (16 for LB = 2)
//     instead of 4, so one warp keeps its scheduler busy on its own.
//   * input is the depunctured soft-symbol pair of each trellis step (2 bytes, what puncturer::depuncture hands to
//     viterbi::conv_decode, viterbi.cpp:31-37) instead of a precomputed 4-byte metric word: half the bytes between the
//     front end and this kernel, and the Viterbi-only entry point reads the caller's symbols in place.  The metric
//     word (viterbi.cpp:234-248: four distinct values per step) is formed while the block is staged into shared memory.
//
// Position p holds state rotl6^t(p) at time t (the in-place butterfly of viterbi_acs2.cuh); the survivor layout in HBM
// is unchanged, so the traceback kernel serves both generations:
//     word (t >> 3) * 16 + (p >> 2), byte (p ^ 1) & 3, bit 7 - (t & 7)
#pragma once

#include "viterbi_acs2.cuh"

namespace b200rx {

// LAZY: renormalisation keeps a per-frame offset instead of subtracting the minimum from all 64 metrics (see renorm()).
template <int LB, bool LAZY>
struct Acs3 {
    static constexpr int T = 1 << LB;       // lanes per frame pair
    static constexpr int NR = 64 >> LB;     // u16x2 registers per lane: one per position, (frame A, frame B)
    static constexpr int PPW = 32 / T;      // frame pairs per warp
    static constexpr int NA = NR / 2;       // decision accumulators per lane
    static_assert(LB >= 1 && LB <= 4, "1 to 4 lane bits");

    // The branch class is GF(2)-linear in the position bits: register i of a lane sees the lane's class XOR xreg(phase, i).
    static __host__ __device__ constexpr uint32_t xreg(int r, int i) { return acs2_class(acs2_rotl6((uint32_t)i, r) & 31u); }
    static __host__ __device__ constexpr bool xused(int r, uint32_t x)
    {
        for (int i = 0; i < NR; i++)
            if (xreg(r, i) == x) return true;
        return false;
    }

    struct Lane {
        uint32_t sel[6][4];  // PRMT selector (A's metric of class c ^ x -> high half, B's -> low half) per phase and offset x
        uint32_t ck[LB];     // lane phases: 0x01000100 - (this lane holds predecessor j+32 ? 0x00010001 : 0)
        uint32_t flip[3];    // decision bits this lane records inverted, per 8-step store of a 24-step block
        uint32_t thr_add;    // lane 0 of the pair: 0x7F2D7F2D - offset (x + it has bit 15 set <=> x > 210), else 0x80008000
        uint32_t trig;       // lane 0 of the pair: 0x80008000, else 0: the halves of x + thr_add that may trigger
        uint32_t cap;        // saturation value of both frames: 0x00FF00FF + offset
        uint32_t off;        // LAZY: what the stored metrics of (frame A, frame B) exceed the reference's by
        uint32_t neg1, one, two; // derived from a kernel argument so that ptxas keeps the IMADs on the FMA pipe
    };

    static __device__ __forceinline__ void lane_init(Lane &L, int glane, uint32_t neg1)
    {
        L.neg1 = neg1;
        L.one = 0u - neg1;
        L.two = L.one + L.one;
        L.thr_add = glane == 0 ? 0x7F2D7F2Du : 0x80008000u;
        L.trig = glane == 0 ? 0x80008000u : 0u;
        L.cap = ACS2_CAP;
        L.off = 0u;
        const uint32_t p_lane = (uint32_t)glane << (6 - LB);
#pragma unroll
        for (int r = 0; r < 6; r++) {
            const uint32_t c = acs2_class(acs2_rotl6(p_lane, r) & 31u);
#pragma unroll
            for (uint32_t x = 0; x < 4; x++) {
                const uint32_t a = c ^ x;       // byte of frame A's metric word
                const uint32_t b = 4u + (c ^ x); // byte of frame B's metric word
                // result bytes: [B's metric, 0, A's metric, 0]; selector bit 3 replicates the sign bit of a byte < 128 = 0
                L.sel[r][x] = ((0x8u | a) << 12) | (a << 8) | ((0x8u | b) << 4) | b;
            }
        }
        uint32_t inv[LB];
#pragma unroll
        for (int k = 0; k < LB; k++) {
            inv[k] = (glane >> (LB - 1 - k)) & 1u;
            L.ck[k] = ACS2_C - inv[k] * 0x00010001u;
        }
#pragma unroll
        for (int o = 0; o < 3; o++) {
            uint32_t m = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int ph = (8 * o + i) % 6;
                if (ph < LB) m |= inv[ph < LB ? ph : 0] << (7 - i);
            }
            L.flip[o] = m * 0x01010101u;
        }
    }

    static __device__ __forceinline__ void init_metrics(uint32_t (&R)[NR], int glane)
    {
#pragma unroll
        for (int i = 0; i < NR; i++) R[i] = 0x003F003Fu;
        if (glane == 0) R[0] = 0u; // state 0 of both frames starts at 0 (viterbi.cpp:71-78)
    }

    // One trellis step of phase PH (= t mod 6) for the lane's NR positions of both frames.  wA / wB: the step's branch-
    // metric words of frame A / B; D: raw decision words (bit 8 = frame B, bit 24 = frame A; bits 9-15 / 25-31 zero).
    template <int PH>
    static __device__ __forceinline__ void step(uint32_t (&R)[NR], uint32_t (&D)[NR], uint32_t wA, uint32_t wB, const Lane &L)
    {
        constexpr int axis = 5 - PH; // position bit separating the butterfly partners
        uint32_t Mv[4], Miv[4];      // (m, m) of both frames per class offset, and 63 - m (psubusb 63, m)
#pragma unroll
        for (uint32_t x = 0; x < 4; x++) {
            Mv[x] = Miv[x] = 0;
            if (xused(PH, x)) {
                Mv[x] = acs2_prmt(wA, wB, L.sel[PH][x]);
                Miv[x] = acs2_fma(Mv[x], L.neg1, 0x003F003Fu);
            }
        }
        if constexpr (axis >= 6 - LB) {
            // ---- partners in another lane ----
            constexpr int lbit = axis - (6 - LB);
            constexpr int k = LB - 1 - lbit;
            static_assert(k == PH, "lane phases come first");
#pragma unroll
            for (int i = 0; i < NR; i++) {
                const uint32_t M = Mv[xreg(PH, i)], Mi = Miv[xreg(PH, i)];
                const uint32_t G = __viaddmin_u16x2(R[i], Mi, L.cap);          // my candidate for the partner's new state
                const uint32_t S = __shfl_xor_sync(0xFFFFFFFFu, G, 1 << lbit); // the partner's candidate for mine
                const uint32_t V = __viaddmin_u16x2(R[i], M, L.cap);           // my own candidate
                R[i] = __vminu2(V, S);
                D[i] = acs2_fma(V, L.one, acs2_fma(S, L.neg1, L.ck[k]));       // bit 8/24: own >= partner (or >, see ck)
            }
        } else {
            // ---- partners in another register of this lane ----
            constexpr int q = axis;
#pragma unroll
            for (int a = 0; a < NR; a++) {
                if ((a >> q) & 1) continue;
                const int b = a | (1 << q);
                const uint32_t M = Mv[xreg(PH, a)], Mi = Miv[xreg(PH, a)];
                const uint32_t B = __viaddmin_u16x2(R[b], Mi, L.cap);    // via j+32 -> new state 2j
                const uint32_t E = __viaddmin_u16x2(R[b], M, L.cap);     // via j+32 -> new state 2j+1
                const uint32_t Ya = __viaddmin_u16x2(R[a], M, B);        // min(X[j] + m, B); ties keep B's value
                const uint32_t Yb = __viaddmin_u16x2(R[a], Mi, E);
                D[a] = acs2_fma(Ya, L.one, acs2_fma(B, L.neg1, ACS2_C));
                D[b] = acs2_fma(Yb, L.one, acs2_fma(E, L.neg1, ACS2_C));
                R[a] = Ya;
                R[b] = Yb;
            }
        }
    }

    // Reference renormalisation (viterbi.cpp:314-332) for both frames of the pair at once: a frame whose state-0 metric
    // (register 0 of the pair's lane 0) exceeds 210 has the minimum of its 64 metrics subtracted from all of them.
    // The minimum is taken over the lane's registers AND a mask that is 0 in the halves of frames that are not hot (hm;
    // all ones in the lanes that do not hold state 0), so the cross-lane minimum is already "minimum if hot, else 0".
    // LAZY: the subtraction is not carried out.  The stored metrics exceed the reference's by a per-frame offset `off`; the
    // saturation value (cap), and the threshold constant move with it, differences (decisions, the candidates two lanes
    // exchange) do not notice.  A hot frame's new offset is the minimum of its stored metrics (= reference minimum + old
    // offset >= old offset), i.e. off = max(off, minimum-if-hot): 4 instructions instead of one subtraction per register.
    // rebase() brings the offsets back to 0 long before a 16-bit half could overflow.
    static __device__ __forceinline__ void renorm(uint32_t (&R)[NR], Lane &L, int lane)
    {
        (void)lane;
        const uint32_t y = acs2_fma(R[0], L.one, L.thr_add);
        if (__any_sync(0xFFFFFFFFu, (y & L.trig) != 0u)) {
            const uint32_t hm = acs2_prmt(y, y, 0xBB99u); // 0xFFFF in the halves whose frame is hot; lanes without state 0: all ones
            uint32_t m;
            if constexpr (NR == 32) {
                uint32_t t[10];
#pragma unroll
                for (int i = 0; i < 10; i++) t[i] = __vimin3_u16x2(R[3 * i], R[3 * i + 1], R[3 * i + 2]);
                const uint32_t u0 = __vimin3_u16x2(t[0], t[1], t[2]), u1 = __vimin3_u16x2(t[3], t[4], t[5]);
                const uint32_t u2 = __vimin3_u16x2(t[6], t[7], t[8]), u3 = __vimin3_u16x2(t[9], R[30], R[31]);
                m = __vimin3_u16x2(__vimin3_u16x2(u0, u1, u2), u3, hm);
            } else if constexpr (NR == 16) {
                const uint32_t a = __vimin3_u16x2(R[0], R[1], R[2]), b = __vimin3_u16x2(R[3], R[4], R[5]);
                const uint32_t c = __vimin3_u16x2(R[6], R[7], R[8]), d = __vimin3_u16x2(R[9], R[10], R[11]);
                const uint32_t e = __vimin3_u16x2(R[12], R[13], R[14]);
                m = __vimin3_u16x2(__vimin3_u16x2(a, b, c), __vimin3_u16x2(d, e, R[15]), hm);
            } else if constexpr (NR == 8) {
                const uint32_t a = __vimin3_u16x2(R[0], R[1], R[2]), b = __vimin3_u16x2(R[3], R[4], R[5]);
                m = __vimin3_u16x2(a, b, __vimin3_u16x2(R[6], R[7], hm));
            } else {
                m = hm;
#pragma unroll
                for (int i = 0; i < NR; i++) m = __vminu2(m, R[i]);
            }
#pragma unroll
            for (int b = 0; b < LB; b += 2) {
                if (b + 1 < LB) {
                    const uint32_t m1 = __shfl_xor_sync(0xFFFFFFFFu, m, 1 << b);
                    const uint32_t m2 = __shfl_xor_sync(0xFFFFFFFFu, m, 2 << b);
                    const uint32_t m3 = __shfl_xor_sync(0xFFFFFFFFu, m, 3 << b);
                    m = __vminu2(__vimin3_u16x2(m, m1, m2), m3);
                } else {
                    m = __vminu2(m, __shfl_xor_sync(0xFFFFFFFFu, m, 1 << b));
                }
            }
            // m: per frame, the minimum of its 64 stored metrics if it is hot, else 0
            if constexpr (LAZY) {
                L.off = __vmaxu2(L.off, m);
                L.cap = acs2_fma(L.off, L.one, ACS2_CAP);
                if (L.trig) L.thr_add = 0x7F2D7F2Du - L.off;
            } else {
#pragma unroll
                for (int i = 0; i < NR; i++) R[i] = acs2_fma(m, L.neg1, R[i]);
            }
        }
    }

    // LAZY: back to offset 0 (call between blocks; an offset grows by at most 255 per step).
    static __device__ __forceinline__ void rebase(uint32_t (&R)[NR], Lane &L, uint32_t limit)
    {
        if constexpr (LAZY) {
            if (__any_sync(0xFFFFFFFFu, (L.off >> 16) > limit || (L.off & 0xFFFFu) > limit)) {
#pragma unroll
                for (int i = 0; i < NR; i++) R[i] -= L.off;
                L.off = 0u;
                L.cap = ACS2_CAP;
                if (L.trig) L.thr_add = 0x7F2D7F2Du;
            }
        }
    }

    // step + decision history + renormalisation; acc[j] collects registers 2j, 2j+1 of both frames:
    // bytes [B of 2j, A of 2j, B of 2j+1, A of 2j+1]
    template <int PH>
    static __device__ __forceinline__ void one(uint32_t (&R)[NR], uint32_t (&acc)[NA], uint32_t wA, uint32_t wB, Lane &L,
                                               int lane)
    {
        uint32_t D[NR];
        step<PH>(R, D, wA, wB, L);
#pragma unroll
        for (int j = 0; j < NA; j++) acc[j] = acs2_fma(acc[j], L.two, __byte_perm(D[2 * j], D[2 * j + 1], 0x7531u));
        renorm(R, L, lane);
    }
};

// Two trellis steps' soft-symbol pairs (bytes s0, s1, s0', s1') -> their two branch-metric words.  Byte c of a word is
// ((s0 ^ b0) + (s1 ^ b1) + 1) >> 3 with c = (b0, b1): the four distinct values of viterbi.cpp:234-248 (bm_word in
// rx_internal.cuh, here for two steps at a time in 16-bit halves).
__device__ __forceinline__ void acs3_bm_words2(uint32_t x, uint32_t &w0, uint32_t &w1)
{
    const uint32_t X0 = __byte_perm(x, 0u, 0x4240u);          // s0 | s0' << 16
    const uint32_t X1 = __byte_perm(x, 0u, 0x4341u);          // s1 | s1' << 16
    const uint32_t a = X0 + X1 + 0x00010001u;                 // s0 + s1 + 1            (1 .. 511 per half)
    const uint32_t d = X0 - X1 + 0x01000100u;                 // s0 - s1 + 256          (1 .. 511 per half)
    // (x >> 3) per half: the 6-bit result sits in the low byte of the half; what the shift drags into the byte above it is
    // never selected below
    const uint32_t m00 = a >> 3;
    const uint32_t m11 = (0x02000200u - a) >> 3;
    const uint32_t m01 = d >> 3;
    const uint32_t m10 = (0x02000200u - d) >> 3;
    const uint32_t U = __byte_perm(m00, m01, 0x6240u);        // m00, m01, m00', m01'
    const uint32_t V = __byte_perm(m10, m11, 0x6240u);        // m10, m11, m10', m11'
    w0 = __byte_perm(U, V, 0x5410u);
    w1 = __byte_perm(U, V, 0x7632u);
}

} // namespace b200rx
