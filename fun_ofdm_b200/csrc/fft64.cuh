// 64-point DFT of one OFDM symbol inside a warp, two points per lane, in registers (shared by the receive front end
// and the device-side frame generator).
#pragma once

#include <cuda_runtime.h>

namespace b200rx {

constexpr unsigned FULL = 0xFFFFFFFFu;

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 shfl_xor2(double2 v, int m)
{
    return make_double2(__shfl_xor_sync(FULL, v.x, m), __shfl_xor_sync(FULL, v.y, m));
}

// 64-point forward DFT of (v0, v1) = (x[lane], x[lane + 32]).  On return lane holds
// X[k0] in v0 and X[k1] in v1 with k = bitrev6((lane << 1) | slot); in the reference's shifted
// storage (fft.cpp:20-24: data[s] = X[(s + 32) % 64]) that is s = ((slot ^ 1) << 5) | bitrev5(lane).
__device__ __forceinline__ void warp_fft64(double2 &v0, double2 &v1, const double2 *tw, int lane)
{
    double2 a = v0, b = v1;
    v0 = cadd(a, b);
    v1 = cmul(csub(a, b), tw[lane]);
#pragma unroll
    for (int bb = 4; bb >= 0; --bb) {
        const int S = 1 << bb;
        const bool hi = (lane & S) != 0;
        const double2 send = hi ? v0 : v1;
        const double2 recv = shfl_xor2(send, S);
        a = hi ? recv : v0;
        b = hi ? v1 : recv;
        v0 = cadd(a, b);
        const double2 d = csub(a, b);
        v1 = (bb > 0) ? cmul(d, tw[(lane & (S - 1)) << (5 - bb)]) : d;
    }
}


// data carrier c (0..47) -> bin in the reference's shifted order (index s <-> subcarrier s - 32):
// 6..58 without the pilots 11, 25, 39, 53 and DC 32 (phase_tracker.cpp:46-50, symbol_mapper.cpp:24-29)
__device__ __forceinline__ int data_bin(int c)
{
    return c + 6 + (c >= 5) + (c >= 18) + (c >= 24) + (c >= 30) + (c >= 43);
}

} // namespace b200rx
