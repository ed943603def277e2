// qil_upload.cu -- host -> device staging behind the C ABI: a ring of device buffers filled on a private copy stream,
// so that the PCIe transfer of signal i+1 overlaps the encode of signal i (the encode is ~5 ms at n = 28, the 2 GiB
// upload ~39 ms: end to end the path is PCIe bound, and only the overlap keeps the GPU from idling behind the link).
// Round 1 had this in the Python host on torch streams; a Julia (or any) caller of libqilcuda gets the same pipeline
// from four calls:  submit(host) -> acquire(&d_ptr) -> [qil_encode_*_dev on d_ptr] -> release.
#include "qil_common.cuh"

struct qil_uploader {
    qil_ctx* ctx = nullptr;
    int depth = 0;
    size_t bytes = 0;
    std::vector<void*> buf;
    std::vector<cudaEvent_t> ready, freed;     // upload of buffer i finished / consumer of buffer i finished
    std::vector<char> used;
    cudaStream_t copy_stream = nullptr;
    int head = 0, tail = 0, inflight = 0;
};

namespace qil {

qil_uploader* uploader_create(qil_ctx* ctx, int64_t bytes, int depth) {
    QIL_REQUIRE(bytes >= 1 && depth >= 1 && depth <= 16, QIL_ERR_ARGUMENT, "uploader: bytes >= 1, 1 <= depth <= 16");
    qil_uploader* u = new qil_uploader();
    u->ctx = ctx; u->depth = depth; u->bytes = (size_t)bytes;
    try {
        QIL_CUDA(cudaStreamCreateWithFlags(&u->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < depth; ++i) {
            void* p = nullptr;
            QIL_CUDA(cudaMalloc(&p, (size_t)bytes));
            u->buf.push_back(p);
            cudaEvent_t e;
            QIL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            u->ready.push_back(e);
            QIL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            u->freed.push_back(e);
            u->used.push_back(0);
        }
    } catch (...) {
        for (void* p : u->buf) cudaFree(p);
        for (cudaEvent_t e : u->ready) cudaEventDestroy(e);
        for (cudaEvent_t e : u->freed) cudaEventDestroy(e);
        if (u->copy_stream) cudaStreamDestroy(u->copy_stream);
        delete u;
        throw;
    }
    return u;
}

void uploader_submit(qil_uploader* u, const void* host, int64_t bytes) {
    QIL_REQUIRE(bytes >= 0 && (size_t)bytes <= u->bytes, QIL_ERR_ARGUMENT, "uploader: %lld bytes exceed the buffer size %zu",
                (long long)bytes, u->bytes);
    QIL_REQUIRE(u->inflight < u->depth, QIL_ERR_RUNTIME, "uploader: every buffer is in flight; acquire/release one first");
    const int i = u->head;
    if (u->used[i]) QIL_CUDA(cudaStreamWaitEvent(u->copy_stream, u->freed[i], 0));   // previous consumer of this buffer done
    QIL_CUDA(cudaMemcpyAsync(u->buf[i], host, (size_t)bytes, cudaMemcpyHostToDevice, u->copy_stream));
    QIL_CUDA(cudaEventRecord(u->ready[i], u->copy_stream));
    u->head = (i + 1) % u->depth;
    u->inflight++;
}

void* uploader_acquire(qil_uploader* u) {
    QIL_REQUIRE(u->inflight > 0, QIL_ERR_RUNTIME, "uploader: nothing submitted");
    const int i = u->tail;
    QIL_CUDA(cudaStreamWaitEvent(u->ctx->stream, u->ready[i], 0));    // the context's stream continues after the upload
    return u->buf[i];
}

void uploader_release(qil_uploader* u) {
    QIL_REQUIRE(u->inflight > 0, QIL_ERR_RUNTIME, "uploader: nothing acquired");
    const int i = u->tail;
    QIL_CUDA(cudaEventRecord(u->freed[i], u->ctx->stream));
    u->used[i] = 1;
    u->tail = (i + 1) % u->depth;
    u->inflight--;
}

void uploader_destroy(qil_uploader* u) {
    if (!u) return;
    cudaStreamSynchronize(u->copy_stream);
    cudaStreamSynchronize(u->ctx->stream);
    for (void* p : u->buf) cudaFree(p);
    for (cudaEvent_t e : u->ready) cudaEventDestroy(e);
    for (cudaEvent_t e : u->freed) cudaEventDestroy(e);
    cudaStreamDestroy(u->copy_stream);
    delete u;
}

}  // namespace qil
