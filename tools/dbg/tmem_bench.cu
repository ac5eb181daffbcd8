// Microbenchmark: tensor-memory read throughput (tcgen05.ld) per warp and per SM, alone and next to a stream of
// tcgen05.mma, for the accumulator-drain patterns of the fused conv kernel (hp_tc_conv2.cu): how many bytes per clock
// can 4 / 8 / 12 / 16 warps pull out of TMEM, and what does a concurrent MMA stream do to it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _bin/tmem_bench tmem_bench.cu && _bin/tmem_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "../../hand_tracking_samples_b200/csrc/hp_ptx.cuh"
using namespace hp;

// WARPS_LD warps (warp w reads lane quadrant w % 4) each issue `reps` loads of X columns; MMA_MODE: 0 none,
// 1 = SS M128xN128 un-swizzled (conv1), 2 = TS M128xN192 (conv2 v2: A in TMEM columns 448..), issued by warp 16.
template <int X, int MMA_MODE, int INFLIGHT>
__global__ void __launch_bounds__(544, 1) k(long long *out, int reps, int warps_ld)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    __shared__ long long tmax[17];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (warp == 0) ptx::tmem_alloc<512>(&tptr);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tb = tptr;
    if (warp < warps_ld) {
        uint32_t acc = 0;
        const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16);
        const long long t0 = clock64();
        for (int r = 0; r < reps; r += INFLIGHT) {
#pragma unroll
            for (int u = 0; u < INFLIGHT; u++) {
                if (X == 32) { uint32_t v[32]; ptx::tmem_ld32(base + ((r + u) & 3) * 32, v); acc ^= v[0] ^ v[31]; }
                if (X == 16) { uint32_t v[16]; ptx::tmem_ld16(base + ((r + u) & 7) * 16, v); acc ^= v[0] ^ v[15]; }
                if (X == 64) { uint32_t v[64]; ptx::tmem_ld64(base + ((r + u) & 1) * 64, v); acc ^= v[0] ^ v[63]; }
            }
            ptx::tmem_ld_wait();
        }
        const long long t1 = clock64();
        if (acc == 0x12345) out[31] = acc;
        if (lane == 0) tmax[warp] = t1 - t0;
    } else if (warp == 16 && MMA_MODE) {
        const uint32_t sa = ptx::smem_u32(smem), sb = ptx::smem_u32(smem + 49152);
        const long long t0 = clock64();
        if (ptx::elect_one()) {
            if (MMA_MODE == 1) {
                const uint64_t ad = ptx::make_desc_nosw(sa, 128, 512), bd = ptx::make_desc_sw128(sb);
                constexpr uint32_t idesc = ptx::make_idesc_f16(128, 128);
                for (int r = 0; r < reps; r += 4) {
                    ptx::umma_f16_c<true>(tb + 256, ad, bd, idesc);
                    ptx::umma_f16_c<true>(tb + 256, ad + 16, bd + 2, idesc);
                    ptx::umma_f16_c<true>(tb + 384, ad + 32, bd + 4, idesc);
                    ptx::umma_f16_c<true>(tb + 384, ad + 48, bd + 6, idesc);
                }
            } else {
                const uint64_t bd = ptx::make_desc_nosw(sa, 4096, 128);
                constexpr uint32_t idesc = ptx::make_idesc_f16(128, 192);
                for (int r = 0; r < reps; r += 4) {
                    ptx::umma_f16_ts_c<true>(tb + 256, tb + 448, bd, idesc);
                    ptx::umma_f16_ts_c<true>(tb + 256, tb + 456, bd + 1, idesc);
                    ptx::umma_f16_ts_c<true>(tb + 256, tb + 464, bd + 16, idesc);
                    ptx::umma_f16_ts_c<true>(tb + 256, tb + 472, bd + 17, idesc);
                }
            }
            ptx::umma_commit(&bar);
        }
        __syncwarp();
        ptx::mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (lane == 0) tmax[16] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        long long m = 0;
        for (int w = 0; w < warps_ld; w++) m = tmax[w] > m ? tmax[w] : m;
        out[0] = m;
        out[1] = MMA_MODE ? tmax[16] : 0;
    }
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tb); }
}

template <int X, int MODE, int INF>
void run(long long *d, int warps_ld)
{
    const int reps = 4096;
    cudaFuncSetAttribute(k<X, MODE, INF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int i = 0; i < 2; i++) k<X, MODE, INF><<<148, 544, 100 * 1024>>>(d, reps, warps_ld);
    cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double bytes = (double)warps_ld * reps * 32.0 * X * 4.0;
    printf("ld.x%-3d in flight %d, %2d warps, mma %d : %7.1f cyc/ld/warp, %6.1f B/clk/SM", X, INF, warps_ld, MODE, (double)h[0] / reps,
           warps_ld ? bytes / (double)h[0] : 0.0);
    if (MODE) printf("   | %6.1f cyc/MMA", (double)h[1] / reps);
    printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    long long *d;
    cudaMalloc(&d, 256);
    for (int w : {1, 4, 8, 12, 16}) run<32, 0, 1>(d, w);
    for (int w : {4, 8, 16}) run<32, 0, 2>(d, w);
    for (int w : {4, 8}) run<64, 0, 1>(d, w);
    for (int w : {4, 8}) run<16, 0, 2>(d, w);
    run<32, 1, 1>(d, 0);
    run<32, 2, 1>(d, 0);
    for (int w : {4, 8, 12}) run<32, 1, 1>(d, w);
    for (int w : {4, 8, 12}) run<32, 2, 1>(d, w);
    for (int w : {8}) run<32, 1, 2>(d, w);
    return 0;
}
