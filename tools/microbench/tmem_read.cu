// Microbenchmark: TMEM read bandwidth of tcgen05.ld.32x32b as seen by 4 or 8 epilogue warps of one CTA
// (one CTA per SM).  Decides whether the int8 contraction's epilogue (which must read every accumulator)
// can go faster with a second epilogue warpgroup.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[X]);
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// WARPS epilogue warps; warp w reads TMEM lanes (w%4)*32.., columns [(w/4)*COLS_PER_GROUP, +COLS_PER_GROUP)
template <int WARPS, int X, bool PIPE>
__global__ void __launch_bounds__(WARPS * 32, 1) tmem_read_kernel(int iters, unsigned long long* cycles_out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    constexpr int GROUPS = WARPS / 4;
    constexpr int COLS = 256 / GROUPS;               // columns of one 256-column accumulator this warp reads per "tile"
    const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * COLS;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t buf = (it & 1) * 256;
        if constexpr (!PIPE) {
#pragma unroll
            for (int c = 0; c < COLS; c += X) {
                uint32_t r[X];
                tmem_ld<X>(taddr + buf + c, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t m = r[0];
#pragma unroll
                for (int i = 1; i < X; ++i) m = max(m, r[i]);   // a max tree like the real epilogue (cheap)
                acc = max(acc, m);
            }
        } else {                                     // the real epilogue's pattern: next slab in flight while this one is scanned
            uint32_t ra[X], rb[X];
            tmem_ld<X>(taddr + buf, ra);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < COLS; c += 2 * X) {
                tmem_ld<X>(taddr + buf + c + X, rb);
                uint32_t m = ra[0];
#pragma unroll
                for (int i = 1; i < X; ++i) m = max(m, ra[i]);
                acc = max(acc, m);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c + 2 * X < COLS) tmem_ld<X>(taddr + buf + c + 2 * X, ra);
                m = rb[0];
#pragma unroll
                for (int i = 1; i < X; ++i) m = max(m, rb[i]);
                acc = max(acc, m);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

template <int WARPS, int X, bool PIPE>
void run(const char* name, int grid) {
    unsigned long long* d_c; uint32_t* d_s;
    cudaMalloc(&d_c, grid * sizeof(unsigned long long)); cudaMalloc(&d_s, 4);
    const int iters = 20000;
    tmem_read_kernel<WARPS, X, PIPE><<<grid, WARPS * 32>>>(100, d_c, d_s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    tmem_read_kernel<WARPS, X, PIPE><<<grid, WARPS * 32>>>(iters, d_c, d_s);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long c0 = 0; cudaMemcpy(&c0, d_c, sizeof(c0), cudaMemcpyDeviceToHost);
    const double bytes_per_tile = 128.0 * 256.0 * 4.0;      // one 128 x 256 fp32/int32 accumulator
    printf("%-28s grid=%d: %s  %.1f cycles per 128x256 accumulator  (%.1f B/cycle/SM), %.3f us per tile\n", name, grid,
           cudaGetErrorString(e), (double)c0 / iters, bytes_per_tile * iters / (double)c0, ms * 1e3 / iters);
    cudaFree(d_c); cudaFree(d_s);
}

int main() {
    run<4, 32, false>("4 warps, x32, ld+wait", 148);
    run<8, 32, false>("8 warps, x32, ld+wait", 148);
    run<4, 32, true>("4 warps, x32, pipelined", 148);
    run<8, 32, true>("8 warps, x32, pipelined", 148);
    run<4, 16, true>("4 warps, x16, pipelined", 148);
    run<8, 16, true>("8 warps, x16, pipelined", 148);
    run<4, 32, true>("4 warps, x32, pipelined", 1);
    run<8, 32, true>("8 warps, x32, pipelined", 1);
    return 0;
}
