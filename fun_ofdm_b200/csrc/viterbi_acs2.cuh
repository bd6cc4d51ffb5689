// Viterbi add-compare-select, second generation: T = 2^LB lanes per frame, 64/T states per lane, IN PLACE.
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
// position differing in bit (5 - t mod 6).  With position = [lane:LB | register:5-LB | half:1] that axis is
//     a lane bit      (LB phases)   -> one shuffle per register, candidates exchanged pre-saturated
//     a register bit  (5-LB phases) -> pure register arithmetic, 4 DPX ops + 1 PRMT per 4 states
//     the half bit    (1 phase)     -> 4 PRMT + 2 DPX ops per 2 states
// Arithmetic is u16x2 DPX: VIADDMNMX.U16x2 does min(x + m, cap) (the saturating add, and with
// cap = the other candidate the add + compare-select in one instruction).  The decision bit is
// bit 8 / 24 of (survivor + 0x0100 - candidate_via_j+32) (set iff equal, i.e. iff j+32 won or tied) and is
// computed with IMADs on the otherwise idle FMA pipe.  Decisions are accumulated over 8 steps per position
// (acc = 2*acc + bits, one IMAD per 4 states) and stored as 64 B per frame per 8 steps; in lane phases the
// lanes holding predecessor j+32 accumulate the complement and flip it back with one XOR before the store.
//
// LB = 5 (one frame per warp, two states per lane) is the low-latency end: most instructions per frame, shortest
// dependency chain per step.  Measured on 32 .. 512 frames: 0.651 .. 0.686 ms against 0.680 .. 0.707 ms for LB = 4 - the
// ~105 cycles a step takes are latency of the dependent DPX / shuffle / vote chain, not work, so it is only reachable
// through B200RX_ACS_LB=5 and the launcher keeps choosing 3 or 4.
// LB trades instructions for parallelism (ncu, 4096 frames x 12 096 steps): LB = 3 needs 11.0
// warp-instructions per trellis step but gives 1024 warps for 592 schedulers (1 or 2 per scheduler, the
// doubly loaded ones set the time); LB = 2 needs fewer instructions per step, has twice the independent
// work per warp and puts exactly one warp on 512 schedulers.  The launcher picks LB from the batch size.
//
// Survivor layout in HBM (per frame, uint32 words), independent of LB:
//     word (t >> 3) * 16 + (p >> 2), byte (p ^ 1) & 3, bit 7 - (t & 7),   p = rotr6^((t+1) % 6)(new state)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200rx {

constexpr int ACS2_BLK = 24;          // steps per unrolled block = lcm(6 phases, 8-step store period)
constexpr int ACS2_WORDS_PER_8 = 16;  // survivor words per frame per 8 steps
constexpr int ACS2_DEFAULT_RN = 1;
constexpr int ACS2_DEFAULT_WARPS = 2; // warps per CTA of the ACS kernel    // renormalisation variant (Acs2::group_min)

__host__ __device__ constexpr uint32_t acs2_rotl6(uint32_t x, int r)
{
    return ((x << (r % 6)) | (x >> (6 - r % 6))) & 63u;
}

// Branch class of butterfly j: (parity(2j & 121), parity(2j & 91)).  GF(2)-linear in the bits of j.
__host__ __device__ constexpr uint32_t acs2_class(uint32_t j)
{
    return ((((j >> 2) ^ (j >> 3) ^ (j >> 4)) & 1u) << 1) | ((j ^ (j >> 2) ^ (j >> 3)) & 1u);
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

constexpr uint32_t ACS2_CAP = 0x00FF00FFu;
constexpr uint32_t ACS2_C = 0x01000100u;

template <int LB>
struct Acs2 {
    static constexpr int T = 1 << LB;       // lanes per frame
    static constexpr int NR = 32 >> LB;     // u16x2 registers per lane
    static constexpr int FPW = 32 / T;      // frames per warp
    static constexpr int NA = NR >= 2 ? NR / 2 : 1; // decision accumulators per lane (NR = 1: one, 16 bits used)
    static_assert(LB >= 2 && LB <= 5, "2 to 5 lane bits");

    // The class is linear in the position bits, so register i of a lane sees the lane's base class pair XOR
    // xreg(phase, i) on both halves: registers with the same xreg share one metric pair (one PRMT).
    static __host__ __device__ constexpr uint32_t xreg(int r, int i) { return acs2_class(acs2_rotl6((uint32_t)i << 1, r) & 31u); }
    static __host__ __device__ constexpr bool xused(int r, uint32_t x)
    {
        for (int i = 0; i < NR; i++)
            if (xreg(r, i) == x) return true;
        return false;
    }

    struct Lane {
        uint32_t sel[6][4];  // PRMT selector (hi class, lo class) per phase and class offset x
        uint32_t ck[LB];     // lane phases: 0x01000100 - (this lane holds predecessor j+32 ? 0x00010001 : 0)
        uint32_t flip[3];    // decision bits this lane records inverted, per 8-step store of a 24-step block
        uint32_t thr;        // renormalisation threshold on register 0 (lane 0 of the frame only)
        uint32_t neg1, one;  // 0xFFFFFFFF and 1 derived from a kernel argument so that ptxas keeps the IMADs
    };

    static __device__ __forceinline__ void lane_init(Lane &L, int glane, uint32_t neg1)
    {
        L.neg1 = neg1;
        L.one = 0u - neg1;
        L.thr = glane == 0 ? 0x00D2FFFFu : 0xFFFFFFFFu;
        const uint32_t p_hi = (uint32_t)glane << (6 - LB);
#pragma unroll
        for (int r = 0; r < 6; r++) {
            const uint32_t c_hi = acs2_class(acs2_rotl6(p_hi, r) & 31u);
            const uint32_t c_lo = acs2_class(acs2_rotl6(p_hi | 1u, r) & 31u);
#pragma unroll
            for (uint32_t x = 0; x < 4; x++) {
                if (r < 5) L.sel[r][x] = 0x8080u | ((c_hi ^ x) << 8) | (c_lo ^ x);
                else L.sel[r][x] = 0x8080u | ((c_hi ^ x) << 8) | (4u + (c_hi ^ x)); // (m high, 63-m low)
            }
        }
        // phase k < LB pairs lanes differing in lane bit (LB-1-k).  The lane holding predecessor j records
        // "partner <= own" directly; the lane holding j+32 needs "own <= partner" and records its complement
        // (own > partner, from the constant below), which is flipped back once per 8 steps before the store.
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
        if (glane == 0) R[0] = 0x0000003Fu; // state 0 (high half of register 0) starts at 0 (viterbi.cpp:71-78)
    }

    // One trellis step of phase PH (= t mod 6).  R: path metrics; w: the step's branch-metric word;
    // D: raw decision words (bit 8 / 24 valid, bits 9-15 / 25-31 zero).
    template <int PH>
    static __device__ __forceinline__ void step(uint32_t (&R)[NR], uint32_t (&D)[NR], uint32_t w, const Lane &L)
    {
        constexpr int axis = 5 - PH; // position bit separating the butterfly partners
        if constexpr (axis >= 1) {
            // metric pairs (m per half) and their complements (63 - m, psubusb 63, m), one per class offset in use
            uint32_t Mv[4], Miv[4];
#pragma unroll
            for (uint32_t x = 0; x < 4; x++) {
                Mv[x] = Miv[x] = 0;
                if (xused(PH, x)) {
                    Mv[x] = acs2_prmt(w, w, L.sel[PH][x]);
                    Miv[x] = acs2_fma(Mv[x], L.neg1, 0x003F003Fu);
                }
            }
            if constexpr (axis >= 6 - LB) {
                // ---- partners in another lane ----
                constexpr int lbit = axis - (6 - LB);
                constexpr int k = LB - 1 - lbit; // index into ck[]
                static_assert(k == PH, "lane phases come first");
#pragma unroll
                for (int i = 0; i < NR; i++) {
                    const uint32_t M = Mv[xreg(PH, i)], Mi = Miv[xreg(PH, i)];
                    const uint32_t G = __viaddmin_u16x2(R[i], Mi, ACS2_CAP);       // my candidate for the partner's new state
                    const uint32_t S = __shfl_xor_sync(0xFFFFFFFFu, G, 1 << lbit); // partner's candidate for mine
                    const uint32_t V = __viaddmin_u16x2(R[i], M, ACS2_CAP);        // my own candidate
                    R[i] = __vminu2(V, S);
                    D[i] = acs2_fma(V, L.one, acs2_fma(S, L.neg1, L.ck[k]));       // bit 8/24: own >= partner (or >, see ck)
                }
            } else {
                // ---- partners in another register of this lane ----
                constexpr int q = axis - 1;
#pragma unroll
                for (int a = 0; a < NR; a++) {
                    if ((a >> q) & 1) continue;
                    const int b = a | (1 << q);
                    const uint32_t M = Mv[xreg(PH, a)], Mi = Miv[xreg(PH, a)];
                    const uint32_t B = __viaddmin_u16x2(R[b], Mi, ACS2_CAP); // via j+32 -> new state 2j
                    const uint32_t E = __viaddmin_u16x2(R[b], M, ACS2_CAP);  // via j+32 -> new state 2j+1
                    const uint32_t Ya = __viaddmin_u16x2(R[a], M, B);        // min(X[j] + m, B); ties keep B's value
                    const uint32_t Yb = __viaddmin_u16x2(R[a], Mi, E);
                    D[a] = acs2_fma(Ya, L.one, acs2_fma(B, L.neg1, ACS2_C));
                    D[b] = acs2_fma(Yb, L.one, acs2_fma(E, L.neg1, ACS2_C));
                    R[a] = Ya;
                    R[b] = Yb;
                }
            }
        } else {
            // ---- partners are the two halves of one register: high = state j, low = state j+32 ----
            const uint32_t wc = w ^ 0x3F3F3F3Fu; // 63 - m per byte
            uint32_t MAv[4], MBv[4];
#pragma unroll
            for (uint32_t x = 0; x < 4; x++) {
                MAv[x] = MBv[x] = 0;
                if (xused(PH, x)) {
                    MAv[x] = acs2_prmt(w, wc, L.sel[PH][x]);              // (m, 63-m)
                    MBv[x] = acs2_fma(MAv[x], L.neg1, 0x003F003Fu);       // (63-m, m)
                }
            }
#pragma unroll
            for (int i = 0; i < NR; i++) {
                const uint32_t W = __byte_perm(R[i], R[i], 0x3232u); // (X[j], X[j])
                const uint32_t Z = __byte_perm(R[i], R[i], 0x1010u); // (X[j+32], X[j+32])
                const uint32_t V = __viaddmin_u16x2(Z, MBv[xreg(PH, i)], ACS2_CAP); // via j+32: (-> 2j, -> 2j+1)
                const uint32_t Y = __viaddmin_u16x2(W, MAv[xreg(PH, i)], V);
                D[i] = acs2_fma(Y, L.one, acs2_fma(V, L.neg1, ACS2_C));
                R[i] = Y;
            }
        }
    }

    // Reference renormalisation (viterbi.cpp:314-332): if metric of state 0 > 210, subtract the minimum of
    // all 64 metrics.  State 0 always sits in the high half of register 0 of the frame's lane 0.
    // It fires on ~11 % of a frame's steps (measured on the reference, all rates and SNRs), i.e. on a third of the steps
    // of a warp that carries four frames, and it sits on the step-to-step dependency chain, so the cross-lane minimum is
    // organised for latency.  RN selects how: 0 = one shuffle round per lane bit; 1 (default) = two lane bits per round
    // (three independent shuffles, three-input DPX minima): ACS kernel alone 0.969 -> 0.944 ms per 4096-frame batch.
    // (__reduce_min_sync over the frame's lanes was tried too: the compiler serialises the four groups of a warp around
    // CREDUX, 2.1 ms.)
    template <int RN>
    static __device__ __forceinline__ uint32_t group_min(uint32_t m, int group)
    {
        (void)group;
        if constexpr (RN == 1) {
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
            return __vminu2(m, __byte_perm(m, m, 0x1032u));
        } else {
#pragma unroll
            for (int s = 1; s < T; s <<= 1) m = __vminu2(m, __shfl_xor_sync(0xFFFFFFFFu, m, s));
            return __vminu2(m, __byte_perm(m, m, 0x1032u)); // both halves = min over all 64 states
        }
    }

    template <int RN>
    static __device__ __forceinline__ void renorm(uint32_t (&R)[NR], const Lane &L, int group)
    {
        const bool hot = R[0] > L.thr;
        if (__any_sync(0xFFFFFFFFu, hot)) {
            const uint32_t any = __ballot_sync(0xFFFFFFFFu, hot);
            uint32_t m = R[0];
#pragma unroll
            for (int i = 1; i < NR; i++) m = __vminu2(m, R[i]);
            m = group_min<RN>(m, group);
            const uint32_t sub = ((any >> (group * T)) & 1u) ? m : 0u;
#pragma unroll
            for (int i = 0; i < NR; i++) R[i] -= sub;
        }
    }

    // step + decision history + renormalisation; acc[j] collects registers 2j, 2j+1
    template <int PH, int RN>
    static __device__ __forceinline__ void one(uint32_t (&R)[NR], uint32_t (&acc)[NA], uint32_t w, const Lane &L,
                                               int glane, int group)
    {
        uint32_t D[NR];
        step<PH>(R, D, w, L);
        // bytes 1 and 3 of each raw decision word are 0/1: gather 4 of them, shift into the 8-step history
        if constexpr (NR >= 2) {
#pragma unroll
            for (int j = 0; j < NR / 2; j++) acc[j] = acc[j] * 2u + __byte_perm(D[2 * j], D[2 * j + 1], 0x7531u);
        } else {
            acc[0] = acc[0] * 2u + __byte_perm(D[0], 0u, 0x4431u); // one register: two bytes, the other two come from lane ^ 1
        }
        renorm<RN>(R, L, group);
    }
};

} // namespace b200rx
