// qil_peer.cu -- the two collectives of the row-sharded encode over NVLink peer memory (no NCCL on the data path).
//
// One process per GPU.  Every rank owns a "symmetric" exchange buffer (cudaMalloc + CUDA IPC): a small header of
// arrival flags followed by a payload area.  The host exchanges the 64-byte IPC handles once (any transport); after
// qil_peer_connect each rank holds a device pointer to every peer's buffer, through which kernels load and store
// directly over NVLink / NVSwitch.
//
// A collective is two short kernels on the context's stream:
//   publish : copy the local contribution into the own payload area, then the last CTA to finish raises
//             flag[PUB][me] = epoch in EVERY rank's header (st.release.sys through the peer mapping);
//   collect : wait (ld.acquire.sys) until flag[PUB][g] == epoch for every g in the own header, then read the G
//             payloads -- peers' through the mapping -- and either sum them in rank order (all-reduce: every rank
//             computes the same bits) or concatenate them (all-gather) into the local destination; the last CTA
//             raises flag[ACK][me] = epoch everywhere ("I no longer read your payload"), which the next publish
//             waits for before it overwrites the payload.
// The reduction therefore runs inside the exchange kernel, tile by tile over the peers' memory, instead of behind a
// library call.  A rank that never arrives makes the waiters trap after ~4 s instead of hanging the device.
#include "qil_common.cuh"

#include <cstring>

namespace qil {

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerHeaderBytes = 1024;   // flags[2][kPeerMaxWorld] (u64) + done counters
constexpr int kPeerThreads = 256;

struct PeerDev {
    int rank, world;
    unsigned long long* my_flags;                       // own header: [2][kPeerMaxWorld]
    unsigned int* my_counter;                           // own header: last-CTA tickets (2 counters)
    unsigned long long* peer_flags[kPeerMaxWorld];      // every rank's header (own entry = local pointer)
    const double* peer_data[kPeerMaxWorld];             // every rank's payload
    double* my_data;
};

}  // namespace qil

struct qil_peer {
    qil_ctx* ctx = nullptr;
    int rank = 0, world = 1;
    size_t payload_bytes = 0;
    unsigned char* base = nullptr;                      // own buffer (cudaMalloc)
    void* mapped[qil::kPeerMaxWorld] = {nullptr};       // cudaIpcOpenMemHandle results (nullptr for self)
    qil::PeerDev dev;
    unsigned long long epoch = 0;
    bool connected = false;
};

namespace qil {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_peer(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long epoch) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
        __nanosleep(64);
        if (clock64() - t0 > 8000000000ll) {             // ~4 s at 1.9 GHz: a peer never arrived
            printf("qil_peer: timed out waiting for a peer (flag %llu < epoch %llu)\n", ld_acquire_sys(flag), epoch);
            __trap();
        }
    }
}

// true in every thread of exactly one CTA: the last one to arrive
__device__ __forceinline__ bool last_cta(unsigned int* counter) {
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(counter, 1u);
        s_last = (ticket == gridDim.x - 1);
        if (s_last) *counter = 0;
    }
    __syncthreads();
    return s_last != 0;
}

__global__ void __launch_bounds__(kPeerThreads)
peer_publish_kernel(const PeerDev pd, const double* __restrict__ src, long long count, unsigned long long epoch) {
    // every peer has finished reading the previous payload
    if (threadIdx.x < pd.world && epoch > 1) wait_flag(pd.my_flags + kPeerMaxWorld + threadIdx.x, epoch - 1);
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        pd.my_data[i] = src[i];
    if (last_cta(pd.my_counter)) {
        __threadfence_system();
        if (threadIdx.x < pd.world) st_release_sys(pd.peer_flags[threadIdx.x] + pd.rank, epoch);
    }
}

// gather == 0: dst[i] = sum_g payload_g[i] (rank order);  gather == 1: dst[g * count + i] = payload_g[i]
__global__ void __launch_bounds__(kPeerThreads)
peer_collect_kernel(const PeerDev pd, double* __restrict__ dst, long long count, unsigned long long epoch, int gather) {
    if (threadIdx.x < pd.world) wait_flag(pd.my_flags + threadIdx.x, epoch);
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        if (gather) {
            for (int g = 0; g < pd.world; ++g) dst[(long long)g * count + i] = ld_peer(pd.peer_data[g] + i);
        } else {
            double s = 0.0;
            for (int g = 0; g < pd.world; ++g) s += ld_peer(pd.peer_data[g] + i);
            dst[i] = s;
        }
    }
    if (last_cta(pd.my_counter + 1)) {
        if (threadIdx.x < pd.world) st_release_sys(pd.peer_flags[threadIdx.x] + kPeerMaxWorld + pd.rank, epoch);
    }
}

static int peer_exchange(qil_peer* p, const void* d_src, void* d_dst, int64_t count, int gather) {
    try {
        QIL_REQUIRE(p->connected, QIL_ERR_RUNTIME, "peer comm: not connected");
        QIL_REQUIRE(count >= 0 && (size_t)count * sizeof(double) <= p->payload_bytes, QIL_ERR_UNSUPPORTED,
                    "peer comm: %lld doubles exceed the %zu-byte exchange buffer", (long long)count, p->payload_bytes);
        if (count == 0) return 0;
        qil_ctx* ctx = p->ctx;
        const unsigned long long epoch = ++p->epoch;
        const int grid = (int)std::max<long long>(1, std::min<long long>((count + kPeerThreads - 1) / kPeerThreads,
                                                                         (long long)ctx->sm_count));
        peer_publish_kernel<<<grid, kPeerThreads, 0, ctx->stream>>>(p->dev, reinterpret_cast<const double*>(d_src),
                                                                   (long long)count, epoch);
        QIL_LAUNCH_CHECK(ctx);
        peer_collect_kernel<<<grid, kPeerThreads, 0, ctx->stream>>>(p->dev, reinterpret_cast<double*>(d_dst),
                                                                   (long long)count, epoch, gather);
        QIL_LAUNCH_CHECK(ctx);
        return 0;
    } catch (const Error& e) {
        callback_error_note() = e.msg;
        return 1;
    }
}

static int peer_allreduce_cb(void* user, void* d_buf, int64_t count) {
    return peer_exchange(reinterpret_cast<qil_peer*>(user), d_buf, d_buf, count, 0);
}
static int peer_allgather_cb(void* user, const void* d_send, void* d_recv, int64_t count) {
    return peer_exchange(reinterpret_cast<qil_peer*>(user), d_send, d_recv, count, 1);
}

qil_peer* peer_create(qil_ctx* ctx, int rank, int world, int64_t bytes, unsigned char* handle64) {
    QIL_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, QIL_ERR_ARGUMENT,
                "peer comm: rank %d of %d (at most %d ranks)", rank, world, kPeerMaxWorld);
    QIL_REQUIRE(bytes >= 8, QIL_ERR_ARGUMENT, "peer comm: empty exchange buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    qil_peer* p = new qil_peer();
    p->ctx = ctx;
    p->rank = rank;
    p->world = world;
    p->payload_bytes = ((size_t)bytes + 255) & ~(size_t)255;
    try {
        QIL_CUDA(cudaMalloc(&p->base, kPeerHeaderBytes + p->payload_bytes));
        QIL_CUDA(cudaMemset(p->base, 0, kPeerHeaderBytes));
        QIL_CUDA(cudaDeviceSynchronize());
        cudaIpcMemHandle_t h;
        QIL_CUDA(cudaIpcGetMemHandle(&h, p->base));
        std::memcpy(handle64, &h, 64);
    } catch (...) {
        if (p->base) cudaFree(p->base);
        delete p;
        throw;
    }
    return p;
}

void peer_connect(qil_peer* p, const unsigned char* all_handles) {
    QIL_REQUIRE(!p->connected, QIL_ERR_RUNTIME, "peer comm: already connected");
    PeerDev& d = p->dev;
    d.rank = p->rank;
    d.world = p->world;
    for (int g = 0; g < p->world; ++g) {
        unsigned char* b = p->base;
        if (g != p->rank) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, all_handles + (size_t)g * 64, 64);
            void* ptr = nullptr;
            QIL_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            p->mapped[g] = ptr;
            b = reinterpret_cast<unsigned char*>(ptr);
        }
        d.peer_flags[g] = reinterpret_cast<unsigned long long*>(b);
        d.peer_data[g] = reinterpret_cast<const double*>(b + kPeerHeaderBytes);
    }
    d.my_flags = reinterpret_cast<unsigned long long*>(p->base);
    d.my_counter = reinterpret_cast<unsigned int*>(p->base + 2 * kPeerMaxWorld * sizeof(unsigned long long));
    d.my_data = reinterpret_cast<double*>(p->base + kPeerHeaderBytes);
    p->connected = true;
}

void peer_fill_comm(qil_peer* p, qil_comm* out) {
    out->rank = p->rank;
    out->world = p->world;
    out->user = p;
    out->allreduce_sum_f64 = peer_allreduce_cb;
    out->allgather_f64 = peer_allgather_cb;
}

void peer_destroy(qil_peer* p) {
    if (!p) return;
    cudaStreamSynchronize(p->ctx->stream);
    for (int g = 0; g < p->world; ++g)
        if (p->mapped[g]) cudaIpcCloseMemHandle(p->mapped[g]);
    if (p->base) cudaFree(p->base);
    delete p;
}

}  // namespace qil
