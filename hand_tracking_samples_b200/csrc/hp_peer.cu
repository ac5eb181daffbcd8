// hp_peer.cu -- the data-parallel exchange step of a training step as ONE kernel over NVLink peer memory:
// reduce-scatter of the gradient sums + the SGD update of CNN::Train (cnn.h:574-575, LFull::update cnn.h:438-445,
// LConv::update cnn.h:269-279) + all-gather of the UPDATED WEIGHTS, replacing "NCCL all-reduce, then sgd_kernel".
//
// One process per GPU.  Every rank exports its weight store, its gradient store and a small flag array as CUDA IPC
// handles (hp_dp_peer_export); hp_dp_peer_init maps the peers' stores into this process.  For a gradient bucket
// [off, off+count) of the flat .cnnb-ordered stores, rank r owns the r-th contiguous slice and
//     1. waits until every rank's gradient sums of the bucket are complete          (flag barrier, per CTA)
//     2. reads its slice of all G gradient stores (G-1 of them through NVLink), adds them in rank order
//     3. w <- fma(-alpha, sum, w) on its slice of its own master weights
//     4. stores the new w into ALL G weight stores (G-1 of them through NVLink)
//     5. signals / waits for "every rank's stores have landed"                       (flag barrier, per CTA)
// so every rank ends with bit-identical weights (they are the owner's bits), the wire volume is that of a ring
// all-reduce ((G-1)/G of the bucket each way), and the separate 113 MB read-modify-write pass of the SGD kernel is
// gone.  NVSwitch gives every peer full bandwidth, so the slice loops just keep many 16-byte loads in flight.
//
// Barriers: flags[b][p] on rank r is written only by CTA b of rank p and holds the last epoch that CTA reached;
// epochs increase monotonically (2 per launch), so no reset and no double buffering.  A CTA spins only on remote
// CTAs of the same kernel launch, which become resident as soon as their own rank's stream reaches the launch: no
// CTA waits on a CTA of its own grid, hence no co-residency requirement.
//
// Failure handling: a spin longer than the timeout (HP_PEER_TIMEOUT_S, default 30 s) ABORTS the exchange instead of
// hanging the GPU or carrying on with incomplete sums: the CTA latches the sticky error word (device + mapped host
// copy), overwrites its own flags on every rank with PEER_POISON so that a peer arriving late aborts too instead of
// passing the barrier on stale epochs, and the kernel returns before the reduce / update / store phases -- the master
// weights are left untouched.  Every later exchange kernel of the rank is a no-op, and the next hp_train_batch* /
// hp_save_cnnb call on the host returns HP_ERR_PEER.
#include "hp_common.cuh"
#include "hp_peer.cuh"

namespace hp {

#define LAUNCH_CHECK(net)                                   \
    do {                                                    \
        (net).launches++;                                   \
        HP_CUDA_TRY(cudaGetLastError());                    \
    } while (0)

constexpr unsigned long long PEER_TIMEOUT_DEFAULT_NS = 30000000000ull;  // 30 s

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// All CTAs with this blockIdx on all ranks meet here.  Everything the CTA's threads wrote before the call is
// visible to the peers' CTAs after they return (bar.sync, then a system-scope release by the signalling threads).
// Returns false (CTA-uniform) when the exchange must be abandoned: a peer did not arrive in time, a peer gave up
// (poisoned flag), or this rank's error word is already set.
template <int WORLD>
__device__ __forceinline__ bool peer_barrier(const PeerPtrs &P, int rank, uint32_t epoch)
{
    __syncthreads();
    int bad = 0;
    if (threadIdx.x < WORLD) {
        const int p = threadIdx.x;
        __threadfence_system();
        if (*reinterpret_cast<volatile uint32_t *>(P.error) != 0) bad = 1;   // sticky: an earlier exchange of this rank failed
        if (!bad) {
            st_release_sys(P.flags[p] + blockIdx.x * PEER_MAX_WORLD + rank, epoch);
            const uint32_t *mine = P.flags[rank] + blockIdx.x * PEER_MAX_WORLD + p;
            const unsigned long long t0 = globaltimer_ns();
            for (;;) {
                const uint32_t v = ld_acquire_sys(mine);
                if (v == PEER_POISON) { bad = 1; break; }
                if ((int32_t)(v - epoch) >= 0) break;
                if (globaltimer_ns() - t0 > P.timeout_ns) { bad = 1; break; }
            }
        }
        if (bad) {
            atomicCAS(P.error, 0u, 1u + (uint32_t)p);
            *P.error_host = 1u + (uint32_t)p;
            for (int q = 0; q < WORLD; q++) st_release_sys(P.flags[q] + blockIdx.x * PEER_MAX_WORLD + rank, PEER_POISON);
            __threadfence_system();
        }
    }
    return __syncthreads_or(bad) == 0;
}

__device__ __forceinline__ float4 ld_peer(const float4 *p)
{
    // peer gradient sums change every step and are read exactly once: bypass L1, do not allocate
    float4 v;
    asm volatile("ld.relaxed.sys.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// PEER_THREADS = 1024 with at most 64 registers: the CTAs are meant to sit on the SMs the persistent tensor-core kernels
// leave free (hp_dp_peer_init reserves them), ONE per SM.  A tcgen05 GEMM CTA owns ~61 k registers and ~226 KB of
// shared memory of its SM, so an exchange CTA that lands beside one keeps the GEMM CTA of that SM from launching until
// the exchange ends -- measured on 8xB200 with 48 x 512-thread CTAs: the fc2 dX GEMM took 52 us instead of 17 and the
// step 362 us instead of ~280.
template <int WORLD, int U>
__global__ void __launch_bounds__(PEER_THREADS, 1) peer_sgd_kernel(PeerPtrs P, int rank, int off, int count4, float alpha, uint32_t epoch)
{
    if (!peer_barrier<WORLD>(P, rank, epoch + 1)) return;   // incomplete sums: leave the weights alone
    const int lo = (int)((int64_t)count4 * rank / WORLD), hi = (int)((int64_t)count4 * (rank + 1) / WORLD);
    const int stride = gridDim.x * PEER_THREADS;
    for (int base = lo + blockIdx.x * PEER_THREADS + threadIdx.x; base < hi; base += stride * U) {
        float4 g[U][WORLD];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = base + u * stride;
#pragma unroll
            for (int p = 0; p < WORLD; p++)
                if (i < hi) g[u][p] = ld_peer(reinterpret_cast<const float4 *>(P.grads[p] + off) + i);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = base + u * stride;
            if (i < hi) {
                float4 s = g[u][0];     // rank order 0..G-1 on every owner: deterministic
#pragma unroll
                for (int p = 1; p < WORLD; p++) { s.x += g[u][p].x; s.y += g[u][p].y; s.z += g[u][p].z; s.w += g[u][p].w; }
                float4 w = reinterpret_cast<const float4 *>(P.params[rank] + off)[i];
                w.x = fmaf(-alpha, s.x, w.x); w.y = fmaf(-alpha, s.y, w.y);
                w.z = fmaf(-alpha, s.z, w.z); w.w = fmaf(-alpha, s.w, w.w);
#pragma unroll
                for (int p = 0; p < WORLD; p++) reinterpret_cast<float4 *>(P.params[p] + off)[i] = w;
            }
        }
    }
    peer_barrier<WORLD>(P, rank, epoch + 2);
}

// The conv bucket (16,864 floats) finishes last and its exchange is exposed at the end of the step, so it is done in
// ONE barrier instead of two: every rank pushes its gradient sums into slot [rank] of every peer's inbox, the CTAs meet
// once, and every rank then adds the G slots of its own inbox in rank order and updates its own weights -- all ranks
// compute the same bits from the same numbers in the same order.  The inbox is double-buffered by step parity: a peer
// can only overwrite the buffer read here after it has passed the NEXT step's barrier, which this rank joins after this
// kernel has finished (stream order), so no closing barrier is needed.
template <int WORLD>
__global__ void __launch_bounds__(PEER_THREADS, 1) peer_small_kernel(PeerPtrs P, int rank, int off, int count4, float alpha, uint32_t epoch, int parity)
{
    const int per = (count4 + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = (lo + per < count4) ? lo + per : count4;
    const size_t slot = (size_t)PEER_SMALL_FLOATS / 4;   // float4 per (parity, rank) slot
    for (int i = lo + threadIdx.x; i < hi; i += PEER_THREADS) {
        const float4 g = reinterpret_cast<const float4 *>(P.grads[rank] + off)[i];
#pragma unroll
        for (int p = 0; p < WORLD; p++) reinterpret_cast<float4 *>(P.inbox[p])[((size_t)parity * PEER_MAX_WORLD + rank) * slot + i] = g;
    }
    if (!peer_barrier<WORLD>(P, rank, epoch + 1)) return;
    for (int i = lo + threadIdx.x; i < hi; i += PEER_THREADS) {
        const float4 *in = reinterpret_cast<const float4 *>(P.inbox[rank]) + (size_t)parity * PEER_MAX_WORLD * slot + i;
        float4 s = ld_peer(in);
#pragma unroll
        for (int p = 1; p < WORLD; p++) {
            const float4 v = ld_peer(in + p * slot);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        float4 w = reinterpret_cast<const float4 *>(P.params[rank] + off)[i];
        w.x = fmaf(-alpha, s.x, w.x); w.y = fmaf(-alpha, s.y, w.y);
        w.z = fmaf(-alpha, s.z, w.z); w.w = fmaf(-alpha, s.w, w.w);
        reinterpret_cast<float4 *>(P.params[rank] + off)[i] = w;
    }
}

int peer_sgd_bucket(Net &net, float alpha, int off, int count, cudaStream_t s)
{
    PeerState *ps = net.peer;
    if (!ps) { set_error("peer path not initialised"); return 2; }
    const int count4 = count / 4;
    const uint32_t epoch = ps->epoch;
    if (count <= PEER_SMALL_FLOATS) {
        int blocks = (count4 + PEER_THREADS - 1) / PEER_THREADS;
        if (blocks > ps->max_blocks) blocks = ps->max_blocks;
        const int parity = (int)(ps->small_steps++ & 1);
        ps->epoch += 1;
        switch (ps->world) {
#define HP_CASE(W) case W: peer_small_kernel<W><<<blocks, PEER_THREADS, 0, s>>>(ps->ptrs, ps->rank, off, count4, alpha, epoch, parity); break;
        HP_CASE(2) HP_CASE(3) HP_CASE(4) HP_CASE(5) HP_CASE(6) HP_CASE(7) HP_CASE(8)
#undef HP_CASE
        default: set_error("peer path supports 2..8 ranks, got %d", ps->world); return 2;
        }
        LAUNCH_CHECK(net);
        net.tc_dirty = true;
        return 0;
    }
    // one CTA per reserved SM on the big buckets
    int blocks = (count4 / ps->world + PEER_THREADS - 1) / PEER_THREADS;
    if (blocks > ps->max_blocks) blocks = ps->max_blocks;
    if (blocks < 1) blocks = 1;
    ps->epoch += 2;
    switch (ps->world) {
#define HP_CASE(W, U) case W: peer_sgd_kernel<W, U><<<blocks, PEER_THREADS, 0, s>>>(ps->ptrs, ps->rank, off, count4, alpha, epoch); break;
    HP_CASE(2, 4) HP_CASE(3, 2) HP_CASE(4, 2) HP_CASE(5, 1) HP_CASE(6, 1) HP_CASE(7, 1) HP_CASE(8, 1)
#undef HP_CASE
    default: set_error("peer path supports 2..8 ranks, got %d", ps->world); return 2;
    }
    LAUNCH_CHECK(net);
    net.tc_dirty = true;
    return 0;
}

int peer_export(Net &net, void *out)
{
    if (!net.peer) {
        PeerState *ps = new PeerState;
        HP_CUDA_TRY(cudaMalloc((void **)&ps->my_flags, PEER_FLAG_WORDS * sizeof(uint32_t)));
        HP_CUDA_TRY(cudaMemset(ps->my_flags, 0, PEER_FLAG_WORDS * sizeof(uint32_t)));
        HP_CUDA_TRY(cudaMalloc((void **)&ps->my_inbox, PEER_INBOX_FLOATS * sizeof(float)));
        HP_CUDA_TRY(cudaMemset(ps->my_inbox, 0, PEER_INBOX_FLOATS * sizeof(float)));
        HP_CUDA_TRY(cudaHostAlloc((void **)&ps->host_err, sizeof(uint32_t), cudaHostAllocMapped));
        *ps->host_err = 0;
        HP_CUDA_TRY(cudaDeviceSynchronize());
        net.peer = ps;
    }
    cudaIpcMemHandle_t h[4];
    HP_CUDA_TRY(cudaIpcGetMemHandle(&h[0], net.params));
    HP_CUDA_TRY(cudaIpcGetMemHandle(&h[1], net.grads));
    HP_CUDA_TRY(cudaIpcGetMemHandle(&h[2], net.peer->my_flags));
    HP_CUDA_TRY(cudaIpcGetMemHandle(&h[3], net.peer->my_inbox));
    static_assert(sizeof(h) == 256, "HP_PEER_HANDLE_BYTES");
    memcpy(out, h, sizeof(h));
    return 0;
}

int peer_init(Net &net, const void *handles, int rank, int world, int reserved_sms)
{
    PeerState *ps = net.peer;
    if (!ps) { set_error("call hp_dp_peer_export first"); return 2; }
    if (world < 2 || world > PEER_MAX_WORLD) { set_error("peer path supports 2..%d ranks, got %d", PEER_MAX_WORLD, world); return 2; }
    ps->rank = rank;
    ps->world = world;
    for (int p = 0; p < world; p++) {
        if (p == rank) {
            ps->ptrs.params[p] = net.params;
            ps->ptrs.grads[p] = net.grads;
            ps->ptrs.flags[p] = ps->my_flags;
            ps->ptrs.inbox[p] = ps->my_inbox;
            continue;
        }
        cudaIpcMemHandle_t h[4];
        memcpy(h, (const char *)handles + (size_t)p * sizeof(h), sizeof(h));
        void *m[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int k = 0; k < 4; k++) {
            HP_CUDA_TRY(cudaIpcOpenMemHandle(&m[k], h[k], cudaIpcMemLazyEnablePeerAccess));
            ps->mapped[ps->n_mapped++] = m[k];
        }
        ps->ptrs.params[p] = (float *)m[0];
        ps->ptrs.grads[p] = (float *)m[1];
        ps->ptrs.flags[p] = (uint32_t *)m[2];
        ps->ptrs.inbox[p] = (float *)m[3];
    }
    ps->ptrs.error = ps->my_flags + PEER_MAX_BLOCKS * PEER_MAX_WORLD;
    {
        uint32_t *dptr = nullptr;
        HP_CUDA_TRY(cudaHostGetDevicePointer((void **)&dptr, ps->host_err, 0));
        ps->ptrs.error_host = dptr;
    }
    ps->ptrs.timeout_ns = PEER_TIMEOUT_DEFAULT_NS;
    if (const char *e = getenv("HP_PEER_TIMEOUT_S")) {
        const double sec = atof(e);
        if (sec > 0) ps->ptrs.timeout_ns = (unsigned long long)(sec * 1e9);
    }
    ps->max_blocks = reserved_sms > 0 ? (reserved_sms < PEER_MAX_BLOCKS ? reserved_sms : PEER_MAX_BLOCKS) : 16;
    if (const char *e = getenv("HP_PEER_BLOCKS")) {
        int b = atoi(e);
        if (b >= 1 && b <= PEER_MAX_BLOCKS) ps->max_blocks = b;
    }
    ps->epoch = 0;
    ps->small_steps = 0;
    ps->ready = true;
    return 0;
}

// host-visible latch, no device sync: non-zero once any exchange kernel of this rank has given up
int peer_failed(const Net &net)
{
    return (net.peer && net.peer->host_err) ? (int)*reinterpret_cast<volatile uint32_t *>(net.peer->host_err) : 0;
}

int peer_status(Net &net, int *err)
{
    *err = 0;
    if (!net.peer || !net.peer->ready) return 0;
    uint32_t v = 0;
    HP_CUDA_TRY(cudaMemcpy(&v, net.peer->ptrs.error, sizeof(v), cudaMemcpyDeviceToHost));
    *err = (int)v;
    return 0;
}

void peer_shutdown(Net &net)
{
    PeerState *ps = net.peer;
    if (!ps) return;
    cudaDeviceSynchronize();
    for (int i = 0; i < ps->n_mapped; i++) cudaIpcCloseMemHandle(ps->mapped[i]);
    if (ps->my_flags) cudaFree(ps->my_flags);
    if (ps->my_inbox) cudaFree(ps->my_inbox);
    if (ps->host_err) cudaFreeHost(ps->host_err);
    delete ps;
    net.peer = nullptr;
}

}  // namespace hp
