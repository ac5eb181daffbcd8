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
#include "hp_ptx.cuh"

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

// This CTA's barrier epoch so far (identical for CTA b of every rank: all ranks make the same sequence of launches with the
// same grids).  The caller advances it with peer_epoch_advance once the launch's barriers are behind it.
__device__ __forceinline__ uint32_t peer_epoch_load(const PeerPtrs &P)
{
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) s_epoch = P.counters[blockIdx.x];
    __syncthreads();
    return s_epoch;
}
__device__ __forceinline__ void peer_epoch_advance(const PeerPtrs &P, uint32_t epoch, uint32_t by)
{
    if (threadIdx.x == 0) P.counters[blockIdx.x] = epoch + by;
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
__global__ void __launch_bounds__(PEER_THREADS, 1) peer_sgd_kernel(PeerPtrs P, int rank, int off, int count4, float alpha)
{
    const uint32_t epoch = peer_epoch_load(P);
    peer_epoch_advance(P, epoch, 2);
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

// ---- the same exchange with the NVLink traffic moved by the TMA engine -----------------------------------------------
// peer_sgd_kernel keeps (G-1) x 16 bytes per thread in flight: 16 CTAs x 1024 threads reach ~220 GB/s of remote loads at
// two GPUs (tools/dbg/p2p_bw.cu) -- a 16.5 MB reduce-scatter read takes ~75 us, longer than the backward kernels it is
// meant to hide behind.  Here one thread per CTA issues cp.async.bulk copies of CHUNK bytes from every peer's gradient
// store into a shared-memory ring (mbarrier complete_tx), STAGES chunks ahead, so each of the 16 CTAs keeps
// STAGES x (G-1) x CHUNK bytes (~170 KB) in flight without occupying registers; the other warps add the chunks in rank
// order (same order, same bits as peer_sgd_kernel), apply w <- fma(-alpha, sum, w), store the owner's copy and stage
// the new weights in a double-buffered shared tile that the same thread bulk-stores into every peer's weight store.
constexpr int PT_CONS = PEER_THREADS - 32;   // consumer threads (warps 0..30); warp 31 lane 0 drives the TMA engine
constexpr int PT_CWARPS = PT_CONS / 32;
__device__ __forceinline__ void pt_wait(uint64_t *bar, uint32_t parity, uint32_t *err)
{
    const unsigned long long t0 = globaltimer_ns();
    while (!ptx::mbar_try_wait(bar, parity)) {
        if (globaltimer_ns() - t0 > 20000000000ull) {   // a bulk copy that never completes: fail loudly instead of hanging the GPU
            atomicExch(err, 0xdeadu);
            asm volatile("trap;");
        }
    }
}
// Stage layout: [WORLD gradient chunks in rank order][this rank's weight chunk]; everything the consumers touch comes
// out of shared memory (the first version read the local gradient and weight chunks with LDG inside the consumer loop:
// a DRAM round trip per chunk, slower than the LDG kernel it was meant to replace).  All hand-overs are mbarriers, so the
// 31 consumer warps run independently of one another:
//   full[s]      TMA -> consumers   (complete_tx)          empty[s]     consumer warps -> TMA thread (count 31)
//   out_full[b]  consumer warps -> TMA thread (count 31)    out_free[b]  TMA thread -> consumers (the bulk stores have read it)
template <int WORLD, int CHUNK_F4, int STAGES>
__global__ void __launch_bounds__(PEER_THREADS, 1) peer_sgd_tma_kernel(PeerPtrs P, int rank, int off, int count4, float alpha)
{
    const uint32_t epoch = peer_epoch_load(P);
    peer_epoch_advance(P, epoch, 2);
    constexpr int CHUNK_B = CHUNK_F4 * 16;
    constexpr int SLOTS = WORLD + 1;
    extern __shared__ uint8_t pt_raw[];
    uint8_t *sm = pt_raw + ((128u - (ptx::smem_u32(pt_raw) & 127u)) & 127u);
    uint8_t *stage0 = sm;                                     // [STAGES][SLOTS][CHUNK_B]
    uint8_t *out0 = sm + (size_t)STAGES * SLOTS * CHUNK_B;    // [2][CHUNK_B]
    __shared__ uint64_t full[STAGES], empty[STAGES], out_full[2], out_free[2];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], PT_CWARPS);
        }
        for (int b = 0; b < 2; b++) {
            ptx::mbar_init(&out_full[b], PT_CWARPS);
            ptx::mbar_init(&out_free[b], 1);
        }
        ptx::fence_barrier_init();
    }
    if (!peer_barrier<WORLD>(P, rank, epoch + 1)) return;   // incomplete sums: leave the weights alone (also a CTA-wide sync)
    ptx::fence_proxy_async_all();                           // the bulk (async-proxy) reads below come after the barrier's acquire
    const int lo = (int)((int64_t)count4 * rank / WORLD), hi = (int)((int64_t)count4 * (rank + 1) / WORLD);
    const int nchunks = (hi - lo + CHUNK_F4 - 1) / CHUNK_F4;
    const int my_chunks = (nchunks > (int)blockIdx.x) ? (nchunks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto chunk_range = [&](int k, int &f0, int &nf) {
        f0 = lo + ((int)blockIdx.x + k * (int)gridDim.x) * CHUNK_F4;
        nf = (hi - f0 < CHUNK_F4) ? hi - f0 : CHUNK_F4;
    };
    if (warp == 31) {
        if (lane == 0) {
            auto load = [&](int k) {
                const int sg = k % STAGES;
                int f0, nf;
                chunk_range(k, f0, nf);
                pt_wait(&empty[sg], ((k / STAGES) & 1) ^ 1, P.error);
                ptx::mbar_expect_tx(&full[sg], (uint32_t)(nf * 16 * SLOTS));
                uint8_t *st = stage0 + (size_t)sg * SLOTS * CHUNK_B;
#pragma unroll
                for (int p = 0; p < WORLD; p++)
                    ptx::bulk_load_1d(st + (size_t)p * CHUNK_B, reinterpret_cast<const float4 *>(P.grads[p] + off) + f0, (uint32_t)(nf * 16), &full[sg]);
                ptx::bulk_load_1d(st + (size_t)WORLD * CHUNK_B, reinterpret_cast<const float4 *>(P.params[rank] + off) + f0, (uint32_t)(nf * 16), &full[sg]);
            };
            for (int k = 0; k < STAGES - 1 && k < my_chunks; k++) load(k);
            for (int k = 0; k < my_chunks; k++) {
                if (k + STAGES - 1 < my_chunks) load(k + STAGES - 1);
                const int ob = k & 1;
                int f0, nf;
                chunk_range(k, f0, nf);
                pt_wait(&out_full[ob], (k >> 1) & 1, P.error);   // every consumer warp has written (and proxy-fenced) its part
                const uint8_t *ot = out0 + (size_t)ob * CHUNK_B;
#pragma unroll
                for (int p = 0; p < WORLD; p++)   // the owner's own store goes through the same engine
                    ptx::bulk_store_1d(reinterpret_cast<float4 *>(P.params[p] + off) + f0, ot, (uint32_t)(nf * 16));
                ptx::bulk_commit();
                if (k >= 1) {   // the stores of chunk k-1 have read their out buffer: hand it back
                    ptx::bulk_wait_read<1>();
                    ptx::mbar_arrive(&out_free[(k - 1) & 1]);
                }
            }
            ptx::bulk_wait<0>();      // every store of this CTA has been performed
            __threadfence_system();
        }
    } else {
        const int t = threadIdx.x;   // 0 .. PT_CONS-1
        for (int k = 0; k < my_chunks; k++) {
            const int sg = k % STAGES, ob = k & 1;
            int f0, nf;
            chunk_range(k, f0, nf);
            pt_wait(&full[sg], (k / STAGES) & 1, P.error);
            if (k >= 2) pt_wait(&out_free[ob], ((k >> 1) - 1) & 1, P.error);   // chunk k-2's stores no longer read this buffer
            const float4 *st = reinterpret_cast<const float4 *>(stage0 + (size_t)sg * SLOTS * CHUNK_B);
            float4 *ot = reinterpret_cast<float4 *>(out0 + (size_t)ob * CHUNK_B);
            for (int i = t; i < nf; i += PT_CONS) {
                float4 s = st[i];   // rank order 0..G-1 on every owner: deterministic, same bits as peer_sgd_kernel
#pragma unroll
                for (int p = 1; p < WORLD; p++) {
                    const float4 v = st[p * CHUNK_F4 + i];
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
                float4 w = st[WORLD * CHUNK_F4 + i];
                w.x = fmaf(-alpha, s.x, w.x); w.y = fmaf(-alpha, s.y, w.y);
                w.z = fmaf(-alpha, s.z, w.z); w.w = fmaf(-alpha, s.w, w.w);
                ot[i] = w;
            }
            ptx::fence_proxy_async();   // this thread's out-buffer writes -> visible to the bulk stores
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(&empty[sg]);
                ptx::mbar_arrive(&out_full[ob]);
            }
        }
    }
    peer_barrier<WORLD>(P, rank, epoch + 2);
}

template <int WORLD, int CHUNK_F4, int STAGES>
static int launch_peer_tma(PeerState *ps, int blocks, int off, int count4, float alpha, cudaStream_t s)
{
    constexpr int SMEM = (STAGES * (WORLD + 1) + 2) * CHUNK_F4 * 16 + 128;
    static_assert(SMEM <= 227 * 1024, "peer exchange tile too large");
    static bool attr = false;
    if (!attr) {
        HP_CUDA_TRY(cudaFuncSetAttribute(peer_sgd_tma_kernel<WORLD, CHUNK_F4, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr = true;
    }
    peer_sgd_tma_kernel<WORLD, CHUNK_F4, STAGES><<<blocks, PEER_THREADS, SMEM, s>>>(ps->ptrs, ps->rank, off, count4, alpha);
    return 0;
}

// The conv bucket (16,864 floats) finishes last and its exchange is exposed at the end of the step, so it is done in
// ONE barrier instead of two: every rank pushes its gradient sums into slot [rank] of every peer's inbox, the CTAs meet
// once, and every rank then adds the G slots of its own inbox in rank order and updates its own weights -- all ranks
// compute the same bits from the same numbers in the same order.  The inbox is double-buffered by step parity: a peer
// can only overwrite the buffer read here after it has passed the NEXT step's barrier, which this rank joins after this
// kernel has finished (stream order), so no closing barrier is needed.
template <int WORLD>
__global__ void __launch_bounds__(PEER_THREADS, 1) peer_small_kernel(PeerPtrs P, int rank, int off, int count4, float alpha)
{
    const uint32_t epoch = peer_epoch_load(P);
    const int parity = (int)(P.counters[PEER_MAX_BLOCKS + blockIdx.x] & 1u);   // launches of this kernel so far: inbox double buffer
    __syncthreads();
    if (threadIdx.x == 0) P.counters[PEER_MAX_BLOCKS + blockIdx.x] += 1u;
    peer_epoch_advance(P, epoch, 1);
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
    if (count <= PEER_SMALL_FLOATS) {
        int blocks = (count4 + PEER_THREADS - 1) / PEER_THREADS;
        if (blocks > ps->max_blocks) blocks = ps->max_blocks;
        switch (ps->world) {
#define HP_CASE(W) case W: peer_small_kernel<W><<<blocks, PEER_THREADS, 0, s>>>(ps->ptrs, ps->rank, off, count4, alpha); break;
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
    // HP_PEER_TMA=1 selects the TMA variant.  Measured at 2 GPUs (profiles/r2_dp_exchange.md) the two are equal (278-288 us
    // per step) and flat in the number of exchange CTAs beyond 16: the step is then bounded by the structure around the
    // exchange (exposed conv bucket + shadow refresh, SM reservation), not by the bytes in flight; the LDG kernel, which
    // round 1 validated on 8 GPUs, stays the default.
    static const bool use_tma = getenv("HP_PEER_TMA") && getenv("HP_PEER_TMA")[0] == '1';
    if (use_tma) {
        int rc = 2;
        switch (ps->world) {   // stage = (G + 1) chunks; STAGES x stage + 2 out buffers <= ~190 KB
        case 2: rc = launch_peer_tma<2, 1024, 3>(ps, blocks, off, count4, alpha, s); break;
        case 3: rc = launch_peer_tma<3, 1024, 2>(ps, blocks, off, count4, alpha, s); break;
        case 4: rc = launch_peer_tma<4, 512, 4>(ps, blocks, off, count4, alpha, s); break;
        case 5: rc = launch_peer_tma<5, 512, 3>(ps, blocks, off, count4, alpha, s); break;
        case 6: rc = launch_peer_tma<6, 512, 3>(ps, blocks, off, count4, alpha, s); break;
        case 7: rc = launch_peer_tma<7, 512, 2>(ps, blocks, off, count4, alpha, s); break;
        case 8: rc = launch_peer_tma<8, 512, 2>(ps, blocks, off, count4, alpha, s); break;
        default: set_error("peer path supports 2..8 ranks, got %d", ps->world); return 2;
        }
        if (rc) return rc;
        LAUNCH_CHECK(net);
        net.tc_dirty = true;
        return 0;
    }
    switch (ps->world) {
#define HP_CASE(W, U) case W: peer_sgd_kernel<W, U><<<blocks, PEER_THREADS, 0, s>>>(ps->ptrs, ps->rank, off, count4, alpha); break;
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
    ps->ptrs.counters = ps->my_flags + PEER_MAX_BLOCKS * PEER_MAX_WORLD + 32;   // zeroed with the flags at export
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
