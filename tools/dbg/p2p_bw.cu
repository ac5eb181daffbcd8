// p2p_bw.cu -- single-process 2-GPU microbenchmark of NVLink peer loads/stores from SM code, to size the exchange kernel
// (csrc/hp_peer.cu).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dbg/_bin/p2p_bw tools/dbg/p2p_bw.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE> __device__ __forceinline__ float4 ld(const float4 *p)
{
    float4 v;
    if (MODE == 0) asm volatile("ld.relaxed.sys.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else v = *p;
    return v;
}
// read n4 float4 from src (remote), sum into a local sink
template <int MODE, int U> __global__ void __launch_bounds__(512) rd(const float4 *src, float4 *sink, int n4)
{
    float4 a = make_float4(0, 0, 0, 0);
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) if (i + u * stride < n4) v[u] = ld<MODE>(src + i + u * stride); else v[u] = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; u++) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
    }
    if (a.x == 123.456f) sink[0] = a;
}
template <int U> __global__ void __launch_bounds__(512) wr(float4 *dst, int n4)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) dst[i] = make_float4(1.f, 2.f, 3.f, (float)i);
}
// exchange-like: read remote + local, write remote + local
template <int MODE, int U> __global__ void __launch_bounds__(512) rw(const float4 *rsrc, const float4 *lsrc, float4 *rdst, float4 *ldst, int n4)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride * U) {
        float4 v[U], w[U];
#pragma unroll
        for (int u = 0; u < U; u++) if (i + u * stride < n4) { v[u] = ld<MODE>(rsrc + i + u * stride); w[u] = ld<MODE>(lsrc + i + u * stride); }
#pragma unroll
        for (int u = 0; u < U; u++) if (i + u * stride < n4) {
            float4 s = make_float4(v[u].x + w[u].x, v[u].y + w[u].y, v[u].z + w[u].z, v[u].w + w[u].w);
            rdst[i + u * stride] = s; ldst[i + u * stride] = s;
        }
    }
}

int main()
{
    int nd = 0; CK(cudaGetDeviceCount(&nd)); if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    const size_t bytes = 18883584 / 2;   // what one rank moves each way per FC bucket at 2 GPUs
    const int n4 = (int)(bytes / 16);
    float4 *a[2], *b[2], *c[2];
    for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaDeviceEnablePeerAccess(1 - d, 0)); CK(cudaMalloc(&a[d], bytes)); CK(cudaMalloc(&b[d], bytes)); CK(cudaMalloc(&c[d], bytes)); CK(cudaMemset(a[d], 0, bytes)); CK(cudaMemset(b[d], 0, bytes)); }
    cudaStream_t st[2]; cudaEvent_t e0[2], e1[2];
    for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaStreamCreate(&st[d])); CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d])); }
    auto run = [&](const char *name, int bidir, auto launch) {
        for (int rep = 0; rep < 2; rep++) {
            for (int d = 0; d < (bidir ? 2 : 1); d++) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d])); for (int it = 0; it < 20; it++) launch(d); CK(cudaEventRecord(e1[d], st[d])); }
            for (int d = 0; d < (bidir ? 2 : 1); d++) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); }
        }
        float ms = 0; CK(cudaSetDevice(0)); CK(cudaEventElapsedTime(&ms, e0[0], e1[0]));
        printf("%-44s %s: %7.1f us  %6.1f GB/s per GPU per direction\n", name, bidir ? "both GPUs" : "one GPU  ", ms * 1e3 / 20, bytes / (ms * 1e-3 / 20) / 1e9);
    };
    for (int bidir = 0; bidir < 2; bidir++) {
        run("cudaMemcpyPeerAsync (copy engine)", bidir, [&](int d) { CK(cudaMemcpyPeerAsync(b[d], d, a[1 - d], 1 - d, bytes, st[d])); });
        for (int blocks : {16, 32, 64, 128, 296}) {
            char nm[96];
            snprintf(nm, 96, "read relaxed.sys U4 blocks %d", blocks); run(nm, bidir, [&](int d) { rd<0, 4><<<blocks, 512, 0, st[d]>>>(a[1 - d], b[d], n4); });
            snprintf(nm, 96, "read ld.nc U4 blocks %d", blocks); run(nm, bidir, [&](int d) { rd<1, 4><<<blocks, 512, 0, st[d]>>>(a[1 - d], b[d], n4); });
            snprintf(nm, 96, "read ld.nc U8 blocks %d", blocks); run(nm, bidir, [&](int d) { rd<1, 8><<<blocks, 512, 0, st[d]>>>(a[1 - d], b[d], n4); });
            snprintf(nm, 96, "write blocks %d", blocks); run(nm, bidir, [&](int d) { wr<1><<<blocks, 512, 0, st[d]>>>(a[1 - d], n4); });
            snprintf(nm, 96, "read+write (exchange-like) nc U4 blocks %d", blocks); run(nm, bidir, [&](int d) { rw<1, 4><<<blocks, 512, 0, st[d]>>>(a[1 - d], b[d], c[1 - d], c[d], n4); });
            snprintf(nm, 96, "read+write (exchange-like) sys U4 blocks %d", blocks); run(nm, bidir, [&](int d) { rw<0, 4><<<blocks, 512, 0, st[d]>>>(a[1 - d], b[d], c[1 - d], c[d], n4); });
        }
    }
    return 0;
}
