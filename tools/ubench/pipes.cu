// Micro-benchmarks behind the design of the ACS kernel: issue rate of the instructions the
// kernel is made of, alone and mixed, on one SM sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// Prints warp-instructions per cycle per SM sub-partition (4 warps resident on each, 8 independent chains per thread).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITER 2048
#define CH 8

__device__ __forceinline__ uint32_t hset_eq(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm volatile("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hmin2(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm volatile("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}

template <int MODE>
__global__ void bench(uint32_t *out, long long *cyc, uint32_t seed, uint32_t neg1)
{
    uint32_t r[CH], q[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { r[i] = seed * (i + 1) + threadIdx.x; q[i] = seed ^ (i * 77 + threadIdx.x); r[i] &= 0x00FF00FFu; q[i] &= 0x00FF00FFu; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) r[i] = __viaddmin_u16x2(r[i], q[i], 0x00FF00FFu);                  // DPX
            if (MODE == 1) r[i] = imad(r[i], neg1, q[i]);                                       // IMAD
            if (MODE == 2) r[i] = hset_eq(r[i], q[i]);                                          // HSET2
            if (MODE == 3) r[i] = prmt(r[i], q[i], 0x7531);                                     // PRMT
            if (MODE == 4) r[i] = __shfl_xor_sync(0xFFFFFFFFu, r[i], 1 + (i & 3));              // SHFL
            if (MODE == 5) { r[i] = __viaddmin_u16x2(r[i], q[i], 0x00FF00FFu); q[i] = imad(q[i], neg1, r[i]); }   // DPX + IMAD
            if (MODE == 6) { r[i] = __viaddmin_u16x2(r[i], q[i], 0x00FF00FFu); q[i] = hset_eq(q[i], r[i]); }      // DPX + HSET2
            if (MODE == 7) { r[i] = imad(r[i], neg1, q[i]); q[i] = hset_eq(q[i], r[i]); }                         // IMAD + HSET2
            if (MODE == 8) { r[i] = __viaddmin_u16x2(r[i], q[i], 0x00FF00FFu); q[i] = __shfl_xor_sync(0xFFFFFFFFu, q[i], 1); } // DPX + SHFL
            if (MODE == 9) { r[i] = __viaddmin_u16x2(r[i], q[i], 0x00FF00FFu); q[i] = imad(q[i], neg1, r[i]); r[i] ^= hset_eq(q[i], r[i]) & 1; } // DPX+IMAD+HSET2+LOP
            if (MODE == 10) r[i] = __vminu2(r[i], q[i]);                                        // VIMNMX
            if (MODE == 11) r[i] = __vimin3_u16x2(r[i], q[i], r[(i + 1) % CH]);                 // VIMNMX3
            if (MODE == 12) { r[i] = __viaddmin_u16x2(r[i], q[i], 0x00FF00FFu); q[i] = __vminu2(q[i], r[i]); }                         // DPX + VIMNMX
            if (MODE == 13) { r[i] = imad(r[i], neg1, q[i]); q[i] = __vminu2(q[i], r[i]); }                                            // IMAD + VIMNMX
            if (MODE == 14) { r[i] = __viaddmin_u16x2(r[i], q[i], 0x00FF00FFu); q[i] = imad(q[i], neg1, r[i]); r[i] = __vminu2(q[i], r[i]); } // DPX + IMAD + VIMNMX
            if (MODE == 15) r[i] = r[i] + q[i] + seed;                                           // IADD3
            if (MODE == 16) r[i] = (r[i] & q[i]) ^ seed;                                          // LOP3
            if (MODE == 17) { r[i] = __viaddmin_u16x2(r[i], q[i], 0x00FF00FFu); q[i] = (q[i] & r[i]) ^ seed; }                         // DPX + LOP3
            if (MODE == 18) { r[i] = imad(r[i], neg1, q[i]); q[i] = (q[i] & r[i]) ^ seed; }                                            // IMAD + LOP3
            if (MODE == 19) { r[i] = __viaddmin_u16x2(r[i], q[i], 0x00FF00FFu); q[i] = imad(q[i], neg1, r[i]); q[i] = imad(q[i], neg1, r[i]); } // DPX + 2 IMAD
            if (MODE == 20) { uint32_t t = __viaddmin_u16x2(q[i], r[i], 0x00FF00FFu); r[i] = __viaddmin_u16x2(r[i], q[i], t); q[i] = imad(r[i], seed, imad(t, neg1, 0x01000100u)); } // acs2 register phase
            if (MODE == 21) { uint32_t t = __viaddmin_u16x2(q[i], r[i], 0x00FF00FFu); r[i] = __viaddmin_u16x2(r[i], q[i], t); q[i] = imad(q[i], 2u, hset_eq(r[i], t)); } // acs3 register phase
            if (MODE == 22) { uint32_t t = __viaddmin_u16x2(q[i], r[i], 0x00FF00FFu); r[i] = __viaddmin_u16x2(r[i], q[i], t); q[i] = imad(q[i], 2u, r[i] + 0x01000100u - t); } // IADD3 decisions
            if (MODE == 23) r[i] = __vmaxu2(r[i], q[i]) + (r[i] > q[i] ? 1 : 0);                   // VIMNMX + ISETP/SEL
            if (MODE == 25) r[i] = hmin2(r[i], q[i]);                                             // HMNMX2
            if (MODE == 26) { r[i] = __viaddmin_u16x2(r[i], q[i], 0x00FF00FFu); q[i] = hmin2(q[i], r[i]); }                            // DPX + HMNMX2
            if (MODE == 27) { r[i] = imad(r[i], neg1, q[i]); q[i] = hmin2(q[i], r[i]); }                                               // IMAD + HMNMX2
            if (MODE == 24) r[i] = __hmin2(*(__half2*)&r[i], *(__half2*)&q[i]).x > (__half)0 ? r[i] : q[i]; // filler
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += r[i] + q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// correctness of the half2 compare on integer bit patterns 0 .. 0x3FF (fp16 subnormals): must equal the integer compare
__global__ void hset_check(uint32_t *bad)
{
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x; // 0 .. 1023*1024
    const uint32_t x = a & 1023u, y = a >> 10;
    uint32_t eq, ge;
    const uint32_t px = x | (y << 16), py = y | (x << 16);
    asm volatile("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(eq) : "r"(px), "r"(py));
    asm volatile("set.ge.u32.f16x2 %0, %1, %2;" : "=r"(ge) : "r"(px), "r"(py));
    const uint32_t want_eq = (x == y ? 0xFFFFu : 0u) | (y == x ? 0xFFFF0000u : 0u);
    const uint32_t want_ge = (x >= y ? 0xFFFFu : 0u) | (y >= x ? 0xFFFF0000u : 0u);
    if (eq != want_eq || ge != want_ge) atomicAdd(bad, 1u);
}

template <int MODE>
void run(const char *name, int per_iter)
{
    uint32_t *out; long long *cyc;
    const int warps = 16; // 4 per sub-partition
    cudaMalloc(&out, 148 * warps * 32 * 4);
    cudaMalloc(&cyc, 148 * 8);
    bench<MODE><<<148, warps * 32>>>(out, cyc, 12345u, 0xFFFFFFFFu);
    bench<MODE><<<148, warps * 32>>>(out, cyc, 12345u, 0xFFFFFFFFu);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += (double)h[i];
    avg /= 148;
    const double instr = (double)ITER * CH * per_iter * (warps / 4);
    printf("%-28s %8.0f cycles  %.3f warp-instr/cycle/SMSP (listed ops only)\n", name, avg, instr / avg);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    uint32_t *bad, hbad = 0;
    cudaMalloc(&bad, 4);
    cudaMemset(bad, 0, 4);
    hset_check<<<1024 * 1024 / 256, 256>>>(bad);
    cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
    printf("set.{eq,ge}.u32.f16x2 on bit patterns 0..1023 vs integer compare: %u mismatches\n", hbad);
    run<0>("VIADDMNMX.U16x2", 1);
    run<1>("IMAD", 1);
    run<2>("HSET2", 1);
    run<3>("PRMT", 1);
    run<4>("SHFL.BFLY", 1);
    run<10>("VIMNMX.U16x2", 1);
    run<11>("VIMNMX3.U16x2", 1);
    run<5>("VIADDMNMX + IMAD", 2);
    run<6>("VIADDMNMX + HSET2", 2);
    run<7>("IMAD + HSET2", 2);
    run<8>("VIADDMNMX + SHFL", 2);
    run<9>("VIADDMNMX+IMAD+HSET2+LOP3", 4);
    run<12>("VIADDMNMX + VIMNMX", 2);
    run<13>("IMAD + VIMNMX", 2);
    run<14>("VIADDMNMX + IMAD + VIMNMX", 3);
    run<15>("IADD3", 1);
    run<16>("LOP3", 1);
    run<17>("VIADDMNMX + LOP3", 2);
    run<18>("IMAD + LOP3", 2);
    run<19>("VIADDMNMX + 2 IMAD", 3);
    run<25>("HMNMX2 (min.f16x2)", 1);
    run<26>("VIADDMNMX + HMNMX2", 2);
    run<27>("IMAD + HMNMX2", 2);
    run<20>("reg phase: 2 DPX + 2 IMAD", 4);
    run<21>("reg phase: 2 DPX + HSET2 + IMAD", 4);
    run<22>("reg phase: 2 DPX + IADD3 + IMAD", 4);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
