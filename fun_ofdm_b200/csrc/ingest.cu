// Ingest of host-resident samples for the host-buffer entry point (b200rx_decode_batch).
//
// The hot path never looks at a cyclic prefix: fft_symbols.cpp:59-62 copies a sample only while m_offset > 15, so of
// every 80-sample symbol slot 64 samples are used, and of the 128 + 80 * (1 + nsym) samples of a frame 128 + 64 * (1 + nsym).
// When the caller's buffer is pinned (b200rx_host_alloc / cudaHostRegister) the GPU can read it over PCIe itself, so
// instead of a DMA copy of the whole buffer this kernel pulls exactly the samples the front end will read - 20 % fewer
// bytes over the link that bounds the end-to-end rate - and stores them at the same indices of the device staging
// buffer (the skipped ranges are never read).  Few registers, one CTA per SM: it runs beside the decode kernels of the
// previous chunk.
#include "rx_internal.cuh"

#include <stdlib.h>

namespace b200rx {

namespace {

constexpr int PULL_THREADS = 256;
constexpr int PULL_BATCH = 8; // loads in flight per thread: PCIe round trips are long, keep many outstanding

// index of the n-th useful sample of a frame relative to its LTS1 tag
__device__ __forceinline__ uint32_t useful_index(uint32_t n)
{
    return n < 128u ? n : 128u + 80u * ((n - 128u) >> 6) + 16u + ((n - 128u) & 63u);
}

template <typename T>
__global__ void __launch_bounds__(PULL_THREADS) pull_kernel(const T *__restrict__ src, T *__restrict__ dst, uint64_t n_samples,
                                                            const uint64_t *__restrict__ lts1, const uint32_t *__restrict__ avail,
                                                            uint32_t n_frames)
{
    for (uint32_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const uint64_t p = lts1[f];
        if (p >= n_samples) continue;
        uint64_t av = avail[f];
        if (av > n_samples - p) av = n_samples - p;
        // LTS1 + LTS2 windows [0, 128), then per 80-sample slot the last 64 (fft_symbols.cpp:53-71)
        const uint32_t useful = av < 128 ? (uint32_t)av : 128u + 64u * (uint32_t)((av - 128) / 80);
        const T *s = src + p;
        T *d = dst + p;
        for (uint32_t base = 0; base < useful; base += PULL_THREADS * PULL_BATCH) {
            T v[PULL_BATCH];
#pragma unroll
            for (int u = 0; u < PULL_BATCH; u++) {
                const uint32_t n = base + u * PULL_THREADS + threadIdx.x;
                if (n < useful) v[u] = s[useful_index(n)];
            }
#pragma unroll
            for (int u = 0; u < PULL_BATCH; u++) {
                const uint32_t n = base + u * PULL_THREADS + threadIdx.x;
                if (n < useful) d[useful_index(n)] = v[u];
            }
        }
    }
}

} // namespace

// src: device-accessible address of the pinned host buffer; dst: device staging; both in format fmt.
cudaError_t launch_pull(const void *src, void *dst, int fmt, uint64_t n_samples, const uint64_t *lts1, const uint32_t *avail,
                        uint32_t n_frames, int sm_count, cudaStream_t s)
{
    if (n_frames == 0) return cudaSuccess;
    const uint32_t cap = (uint32_t)sm_count; // one CTA per SM keeps enough loads in flight to fill the PCIe link
    const unsigned grid = (unsigned)(n_frames < cap ? n_frames : cap);
    switch (fmt) {
        case FMT_FC64:
            pull_kernel<uint4><<<grid, PULL_THREADS, 0, s>>>((const uint4 *)src, (uint4 *)dst, n_samples, lts1, avail, n_frames);
            break;
        case FMT_FC32:
            pull_kernel<uint2><<<grid, PULL_THREADS, 0, s>>>((const uint2 *)src, (uint2 *)dst, n_samples, lts1, avail, n_frames);
            break;
        case FMT_SC16:
            pull_kernel<uint32_t><<<grid, PULL_THREADS, 0, s>>>((const uint32_t *)src, (uint32_t *)dst, n_samples, lts1, avail, n_frames);
            break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

} // namespace b200rx
