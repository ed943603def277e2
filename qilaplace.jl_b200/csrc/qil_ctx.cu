// qil_ctx.cu -- context, device memory and MPS/MPO handle management.
#include "qil_common.cuh"

#include <cstring>

void* qil_ctx::alloc(size_t bytes) {
    if (bytes == 0) bytes = 16;
    void* p = nullptr;
    QIL_CUDA(cudaMallocAsync(&p, bytes, stream));
    return p;
}

void qil_ctx::free(void* p) {
    if (p) cudaFreeAsync(p, stream);
}

void* qil_ctx::get_scratch(size_t bytes) {
    if (bytes > scratch_bytes) {
        if (scratch) cudaFreeAsync(scratch, stream);
        scratch = nullptr;
        size_t nb = bytes + (bytes >> 2) + 256;
        QIL_CUDA(cudaMallocAsync(&scratch, nb, stream));
        scratch_bytes = nb;
    }
    return scratch;
}

unsigned int* qil_ctx::get_grid_sync() {
    if (!grid_sync) {
        QIL_CUDA(cudaMalloc(&grid_sync, 4 * sizeof(unsigned int)));
        QIL_CUDA(cudaMemset(grid_sync, 0, 4 * sizeof(unsigned int)));
    }
    return grid_sync;
}

qil_ctx::~qil_ctx() {
    if (grid_sync) cudaFree(grid_sync);
}

void qil_ctx::prof_begin(int id, double bytes, double flops) {
    if (!prof_on) return;
    if (prof_depth++ > 0) return;
    ProfRegion r;
    r.id = id;
    r.bytes = bytes;
    r.flops = flops;
    QIL_CUDA(cudaEventCreate(&r.e0));
    QIL_CUDA(cudaEventCreate(&r.e1));
    QIL_CUDA(cudaEventRecord(r.e0, stream));
    prof.push_back(r);
}

void qil_ctx::prof_end() {
    if (!prof_on || prof.empty() || prof_depth == 0) return;
    if (--prof_depth > 0) return;
    QIL_CUDA(cudaEventRecord(prof.back().e1, stream));
}

qil_ctx* qil_ctx::aux_ctx(int w) {
    while ((int)aux.size() <= w) {
        qil_ctx* a = new qil_ctx();
        a->device = device;
        a->sm_count = sm_count;
        a->smem_optin = smem_optin;
        a->own_stream = true;
        a->is_aux = true;
        a->d_margin = d_margin;
        QIL_CUDA(cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking));
        aux.push_back(a);
    }
    return aux[w];
}

void qil_ctx::release_aux() {
    for (qil_ctx* a : aux) {
        cudaStreamSynchronize(a->stream);
        if (a->scratch) cudaFreeAsync(a->scratch, a->stream);
        cudaStreamSynchronize(a->stream);
        cudaStreamDestroy(a->stream);
        delete a;
    }
    aux.clear();
}

void qil_ctx::sync() { QIL_CUDA(cudaStreamSynchronize(stream)); }

namespace qil {

void ensure_dynamic_smem_impl(const void* func, size_t bytes) {
    // cudaFuncSetAttribute applies to the CURRENT device: the running maximum is kept per (device, function)
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> current;
    int dev = 0;
    QIL_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = current[std::make_pair(dev, func)];
    if (bytes > cur) {
        QIL_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur = bytes;
    }
}

static void check_bonds(int n, const int64_t* bond) {
    QIL_REQUIRE(n >= 1 && n <= kMaxSites, QIL_ERR_ARGUMENT, "number of sites %d outside [1,%d]", n, kMaxSites);
    QIL_REQUIRE(bond[0] == 1 && bond[n] == 1, QIL_ERR_ARGUMENT, "boundary bonds must have dimension 1");
    for (int i = 0; i <= n; ++i)
        QIL_REQUIRE(bond[i] >= 1 && bond[i] < (1ll << 30), QIL_ERR_ARGUMENT, "bond %d has invalid dimension", i);
}

qil_mps* new_mps(qil_ctx* ctx, int n, int is_complex, const int64_t* bond, bool allocate) {
    check_bonds(n, bond);
    qil_mps* m = new qil_mps();
    m->ctx = ctx;
    m->n = n;
    m->is_complex = is_complex ? 1 : 0;
    m->bond.assign(bond, bond + n + 1);
    m->core.assign(n, nullptr);
    if (allocate)
        for (int i = 0; i < n; ++i) m->core[i] = ctx->alloc(m->core_elems(i) * elem_size(is_complex));
    return m;
}

qil_mpo* new_mpo(qil_ctx* ctx, int n, int is_complex, const int64_t* bond, bool allocate) {
    check_bonds(n, bond);
    qil_mpo* m = new qil_mpo();
    m->ctx = ctx;
    m->n = n;
    m->is_complex = is_complex ? 1 : 0;
    m->bond.assign(bond, bond + n + 1);
    m->core.assign(n, nullptr);
    if (allocate)
        for (int i = 0; i < n; ++i) m->core[i] = ctx->alloc(m->core_elems(i) * elem_size(is_complex));
    return m;
}

void destroy(qil_mps* m) {
    if (!m) return;
    if (!m->pool)
        for (void* p : m->core) m->ctx->free(p);
    delete m;
}

void unpool(qil_mps* m) {
    if (!m || !m->pool) return;
    for (int i = 0; i < m->n; ++i) {
        const size_t bytes = m->core_elems(i) * elem_size(m->is_complex);
        void* p = m->ctx->alloc(bytes);
        QIL_CUDA(cudaMemcpyAsync(p, m->core[i], bytes, cudaMemcpyDeviceToDevice, m->ctx->stream));
        m->core[i] = p;
    }
    QIL_CUDA(cudaStreamSynchronize(m->ctx->stream));     // the pool may be released right after
    m->pool.reset();
}

void destroy(qil_mpo* m) {
    if (!m) return;
    for (void* p : m->core) m->ctx->free(p);
    delete m;
}

ChainDesc make_desc(const qil_mps* m) {
    ChainDesc d;
    d.n = m->n;
    for (int i = 0; i <= m->n; ++i) d.bond[i] = (int)m->bond[i];
    for (int i = 0; i < m->n; ++i) d.core[i] = m->core[i];
    return d;
}

ChainDesc make_desc(const qil_mpo* m) {
    ChainDesc d;
    d.n = m->n;
    for (int i = 0; i <= m->n; ++i) d.bond[i] = (int)m->bond[i];
    for (int i = 0; i < m->n; ++i) d.core[i] = m->core[i];
    return d;
}

}  // namespace qil
