// hp_peer.cuh -- state of the NVLink peer-memory data-parallel path (hp_peer.cu).
#pragma once
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

namespace hp {

constexpr int PEER_MAX_WORLD = 8;
constexpr int PEER_MAX_BLOCKS = 128;
constexpr int PEER_THREADS = 1024;
constexpr int PEER_SMALL_FLOATS = 16896;                                // buckets up to this size (the conv bucket: 16,864) go through the inbox
constexpr int PEER_INBOX_FLOATS = 2 * PEER_MAX_WORLD * PEER_SMALL_FLOATS;   // [parity][source rank][PEER_SMALL_FLOATS]
constexpr int PEER_FLAG_WORDS = PEER_MAX_BLOCKS * PEER_MAX_WORLD + 32 + 2 * PEER_MAX_BLOCKS;   // barrier flags + error word + per-CTA counters

struct PeerPtrs {                        // passed to the kernel by value
    float *params[PEER_MAX_WORLD];       // every rank's FP32 master weights (.cnnb order); [rank] is local
    float *grads[PEER_MAX_WORLD];        // every rank's gradient sums
    uint32_t *flags[PEER_MAX_WORLD];     // every rank's flag array [PEER_MAX_BLOCKS][PEER_MAX_WORLD]
    float *inbox[PEER_MAX_WORLD];        // every rank's small-bucket inbox
    uint32_t *counters;                  // local, per CTA: [b] barrier epoch reached so far, [PEER_MAX_BLOCKS + b] small-bucket launches so far.
                                         // Kept on the device (not passed as launch arguments) so that a captured CUDA graph of the
                                         // training step replays correctly: every launch advances its own CTAs' counters.
    uint32_t *error;                     // local error word (barrier timeout): sticky, every later exchange kernel of this rank is a no-op
    volatile uint32_t *error_host;       // the same word in mapped pinned host memory, so that the host sees it without a device sync
    unsigned long long timeout_ns;       // barrier spin limit (HP_PEER_TIMEOUT_S, default 30 s)
};
constexpr uint32_t PEER_POISON = 0x7ffffff0u;   // flag value a rank that gave up writes in place of its epoch

struct PeerState {
    PeerPtrs ptrs;
    uint32_t *my_flags = nullptr;
    float *my_inbox = nullptr;
    uint32_t *host_err = nullptr;        // cudaHostAlloc'ed (mapped) mirror of the error word
    void *mapped[4 * PEER_MAX_WORLD];
    int n_mapped = 0;
    int rank = 0, world = 1;
    int max_blocks = PEER_MAX_BLOCKS;

    bool ready = false;
};

struct Net;
int peer_export(Net &net, void *out192);
int peer_init(Net &net, const void *handles, int rank, int world, int reserved_sms);
int peer_sgd_bucket(Net &net, float alpha, int off, int count, cudaStream_t s);
int peer_status(Net &net, int *err);
int peer_failed(const Net &net);
void peer_shutdown(Net &net);

}  // namespace hp
