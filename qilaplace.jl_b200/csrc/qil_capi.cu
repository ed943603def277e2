// qil_capi.cu -- extern "C" boundary of libqilcuda.so (see include/qilcuda.h).
#include "qil_mpsops.cuh"

#include <cstring>

static thread_local std::string g_last_error;

#define QIL_API_BEGIN try {
#define QIL_API_END                                         \
    }                                                       \
    catch (const qil::Error& e) {                           \
        g_last_error = e.msg;                               \
        return e.code;                                      \
    }                                                       \
    catch (const std::bad_alloc&) {                         \
        g_last_error = "host allocation failed";            \
        return QIL_ERR_RUNTIME;                             \
    }                                                       \
    catch (const std::exception& e) {                       \
        g_last_error = e.what();                            \
        return QIL_ERR_RUNTIME;                             \
    }                                                       \
    return QIL_OK;

#define QIL_NONNULL(p) QIL_REQUIRE((p) != nullptr, QIL_ERR_ARGUMENT, "null pointer: %s", #p)

using namespace qil;

static inline int64_t fix_maxdim(int64_t maxdim) { return maxdim <= 0 ? ((int64_t)1 << 62) : maxdim; }

template <typename T>
static void qr_host(qil_ctx* ctx, int64_t m, int64_t n, const void* A, int positive, void* Qh, void* Rh) {
    Mat<T> dA(ctx, m, n), Q, R;
    QIL_CUDA(cudaMemcpyAsync(dA.p, A, (size_t)m * n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    qr_thin<T>(ctx, m, n, dA.p, n, positive != 0, Q, R);
    const int64_t k = std::min(m, n);
    QIL_CUDA(cudaMemcpyAsync(Qh, Q.p, (size_t)m * k * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    QIL_CUDA(cudaMemcpyAsync(Rh, R.p, (size_t)k * n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
}

template <typename T>
static int svd_host(qil_ctx* ctx, int64_t m, int64_t n, const void* A, double cutoff, int64_t maxdim, int64_t mindim,
                    void* Uh, double* Sh, void* Vhh) {
    Mat<T> dA(ctx, m, n), U, Vh;
    Mat<double> S;
    QIL_CUDA(cudaMemcpyAsync(dA.p, A, (size_t)m * n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    const int r = svd_trunc<T>(ctx, m, n, dA.p, n, cutoff, maxdim, mindim, Uh ? &U : nullptr, nullptr,
                               Vhh ? &Vh : nullptr, nullptr, Sh ? &S : nullptr);
    if (Uh) QIL_CUDA(cudaMemcpyAsync(Uh, U.p, (size_t)m * r * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    if (Sh) QIL_CUDA(cudaMemcpyAsync(Sh, S.p, (size_t)r * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (Vhh) QIL_CUDA(cudaMemcpyAsync(Vhh, Vh.p, (size_t)r * n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    return r;
}

template <typename T>
static int rsvd_host(qil_ctx* ctx, int64_t m, int64_t n, const void* A, const RsvdOpts& oin, const void* stream,
                     int64_t stream_len, void* Uh, double* Sh, void* Vhh) {
    Mat<T> dA(ctx, m, n), U, Vh, dS;
    Mat<double> S;
    RsvdOpts o = oin;
    QIL_CUDA(cudaMemcpyAsync(dA.p, A, (size_t)m * n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    if (stream && stream_len > 0) {
        dS = Mat<T>(ctx, stream_len, 1);
        QIL_CUDA(cudaMemcpyAsync(dS.p, stream, (size_t)stream_len * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        o.omega = dS.p; o.omega_rows = stream_len; o.omega_cols = 1;
    }
    const int r = rsvd_matrix<T>(ctx, dA.p, m, n, o, U, S, Vh);
    if (Uh) QIL_CUDA(cudaMemcpyAsync(Uh, U.p, (size_t)m * r * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    if (Sh) QIL_CUDA(cudaMemcpyAsync(Sh, S.p, (size_t)r * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (Vhh) QIL_CUDA(cudaMemcpyAsync(Vhh, Vh.p, (size_t)r * n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    return r;
}


extern "C" {


const char* qil_last_error(void) { return g_last_error.c_str(); }
const char* qil_version(void) { return "qilcuda 0.1 (sm_100a)"; }

static int create_impl(int device, void* stream, bool have_stream, qil_ctx** out) {
    QIL_API_BEGIN
    QIL_NONNULL(out);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        QIL_THROW(QIL_ERR_CUDA, "no CUDA device available (%s): libqilcuda has no CPU fallback",
                  cudaGetErrorString(e));
    QIL_REQUIRE(device >= 0 && device < count, QIL_ERR_ARGUMENT, "device %d out of range [0,%d)", device, count);
    QIL_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    QIL_CUDA(cudaGetDeviceProperties(&prop, device));
    QIL_REQUIRE(prop.major >= 10, QIL_ERR_UNSUPPORTED,
                "device %d is sm_%d%d; libqilcuda is built for sm_100a only", device, prop.major, prop.minor);
    qil_ctx* ctx = new qil_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    if (have_stream) {
        ctx->stream = (cudaStream_t)stream;
        ctx->own_stream = false;
    } else {
        QIL_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    {
        const double inf = 1e300;
        QIL_CUDA(cudaMalloc(&ctx->d_margin, sizeof(double)));
        QIL_CUDA(cudaMemcpy(ctx->d_margin, &inf, sizeof(double), cudaMemcpyHostToDevice));
    }
    // keep freed blocks in the pool instead of returning them to the driver at every sync
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = ctx;
    QIL_API_END
}

int qil_create(int device, qil_ctx** out) { return create_impl(device, nullptr, false, out); }
int qil_create_on_stream(int device, void* cuda_stream, qil_ctx** out) {
    return create_impl(device, cuda_stream, true, out);
}

int qil_destroy(qil_ctx* ctx) {
    QIL_API_BEGIN
    if (!ctx) return QIL_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->release_aux();
    if (ctx->d_margin) cudaFree(ctx->d_margin);
    if (ctx->scratch) cudaFreeAsync(ctx->scratch, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    QIL_API_END
}

int qil_sync(qil_ctx* ctx) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx);
    ctx->sync();
    QIL_API_END
}

int qil_launch_count(qil_ctx* ctx, uint64_t* out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx);
    QIL_NONNULL(out);
    *out = ctx->launches;
    QIL_API_END
}

int qil_truncation_margin(qil_ctx* ctx, int reset, double* out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    ctx->sync();
    for (qil_ctx* a : ctx->aux) a->sync();
    QIL_CUDA(cudaMemcpy(out, ctx->d_margin, sizeof(double), cudaMemcpyDeviceToHost));
    if (reset) {
        const double inf = 1e300;
        QIL_CUDA(cudaMemcpy(ctx->d_margin, &inf, sizeof(double), cudaMemcpyHostToDevice));
    }
    QIL_API_END
}

int qil_profile_enable(qil_ctx* ctx, int on) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx);
    ctx->prof_on = on != 0;
    ctx->prof_depth = 0;
    QIL_API_END
}

int qil_profile_reset(qil_ctx* ctx) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx);
    ctx->sync();
    for (auto& r : ctx->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    ctx->prof.clear();
    QIL_API_END
}

static void profile_sum(qil_ctx* ctx, int kernel_class, double* total_ms, int64_t* launches, double* bytes,
                        double* flops) {
    ctx->sync();
    double t = 0.0, b = 0.0, f = 0.0;
    int64_t c = 0;
    for (auto& r : ctx->prof) {
        if (r.id != kernel_class) continue;
        float ms = 0.f;
        QIL_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
        t += ms;
        b += r.bytes;
        f += r.flops;
        ++c;
    }
    *total_ms = t;
    *launches = c;
    if (bytes) *bytes = b;
    if (flops) *flops = f;
}

int qil_profile_read(qil_ctx* ctx, int kernel_class, double* total_ms, int64_t* launches) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(total_ms); QIL_NONNULL(launches);
    profile_sum(ctx, kernel_class, total_ms, launches, nullptr, nullptr);
    QIL_API_END
}

int qil_profile_read_work(qil_ctx* ctx, int kernel_class, double* total_ms, int64_t* launches, double* bytes,
                          double* flops) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(total_ms); QIL_NONNULL(launches); QIL_NONNULL(bytes); QIL_NONNULL(flops);
    profile_sum(ctx, kernel_class, total_ms, launches, bytes, flops);
    QIL_API_END
}

// ---- containers ----------------------------------------------------------------------------
int qil_mps_from_host(qil_ctx* ctx, int n, int is_complex, const int64_t* bond, const void* const* cores,
                      double amplitude, qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(bond); QIL_NONNULL(cores); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    chain_owner<qil_mps> m(new_mps(ctx, n, is_complex, bond, true));
    m->amplitude = amplitude;
    for (int i = 0; i < n; ++i)
        QIL_CUDA(cudaMemcpyAsync(m->core[i], cores[i], m->core_elems(i) * elem_size(is_complex),
                                 cudaMemcpyHostToDevice, ctx->stream));
    ctx->sync();
    *out = m.release();
    QIL_API_END
}

int qil_mps_info(const qil_mps* m, int* n, int* is_complex, double* amplitude) {
    QIL_API_BEGIN
    QIL_NONNULL(m);
    if (n) *n = m->n;
    if (is_complex) *is_complex = m->is_complex;
    if (amplitude) *amplitude = m->amplitude;
    QIL_API_END
}

int qil_mps_dims(const qil_mps* m, int64_t* bond) {
    QIL_API_BEGIN
    QIL_NONNULL(m); QIL_NONNULL(bond);
    for (int i = 0; i <= m->n; ++i) bond[i] = m->bond[i];
    QIL_API_END
}

int qil_mps_get_core(const qil_mps* m, int site, void* host_buf) {
    QIL_API_BEGIN
    QIL_NONNULL(m); QIL_NONNULL(host_buf);
    QIL_REQUIRE(site >= 0 && site < m->n, QIL_ERR_ARGUMENT, "site %d out of range [0,%d)", site, m->n);
    QIL_CUDA(cudaSetDevice(m->ctx->device));
    QIL_CUDA(cudaMemcpyAsync(host_buf, m->core[site], m->core_elems(site) * elem_size(m->is_complex),
                             cudaMemcpyDeviceToHost, m->ctx->stream));
    m->ctx->sync();
    QIL_API_END
}

int qil_mps_get_cores(const qil_mps* m, void* host_buf, int64_t host_bytes) {
    QIL_API_BEGIN
    QIL_NONNULL(m); QIL_NONNULL(host_buf);
    QIL_CUDA(cudaSetDevice(m->ctx->device));
    const size_t es = elem_size(m->is_complex);
    size_t need = 0;
    for (int i = 0; i < m->n; ++i) need += m->core_elems(i) * es;
    QIL_REQUIRE((size_t)host_bytes >= need, QIL_ERR_ARGUMENT, "get_cores: %lld bytes given, %zu needed",
                (long long)host_bytes, need);
    size_t off = 0;
    for (int i = 0; i < m->n; ++i) {
        const size_t nb = m->core_elems(i) * es;
        QIL_CUDA(cudaMemcpyAsync((char*)host_buf + off, m->core[i], nb, cudaMemcpyDeviceToHost, m->ctx->stream));
        off += nb;
    }
    m->ctx->sync();
    QIL_API_END
}

int qil_mps_set_amplitude(qil_mps* m, double amplitude) {
    QIL_API_BEGIN
    QIL_NONNULL(m);
    m->amplitude = amplitude;
    QIL_API_END
}

int qil_mps_clone(const qil_mps* m, qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(m); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(m->ctx->device));
    qil_mps* c = new_mps(m->ctx, m->n, m->is_complex, m->bond.data(), true);
    c->amplitude = m->amplitude;
    for (int i = 0; i < m->n; ++i)
        QIL_CUDA(cudaMemcpyAsync(c->core[i], m->core[i], m->core_elems(i) * elem_size(m->is_complex),
                                 cudaMemcpyDeviceToDevice, m->ctx->stream));
    *out = c;
    QIL_API_END
}

int qil_mps_free(qil_mps* m) {
    QIL_API_BEGIN
    if (m) {
        cudaSetDevice(m->ctx->device);
        destroy(m);
    }
    QIL_API_END
}

int qil_mpo_from_host(qil_ctx* ctx, int n, int is_complex, const int64_t* bond, const void* const* cores,
                      qil_mpo** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(bond); QIL_NONNULL(cores); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    chain_owner<qil_mpo> m(new_mpo(ctx, n, is_complex, bond, true));
    for (int i = 0; i < n; ++i)
        QIL_CUDA(cudaMemcpyAsync(m->core[i], cores[i], m->core_elems(i) * elem_size(is_complex),
                                 cudaMemcpyHostToDevice, ctx->stream));
    ctx->sync();
    *out = m.release();
    QIL_API_END
}

int qil_mpo_info(const qil_mpo* m, int* n, int* is_complex) {
    QIL_API_BEGIN
    QIL_NONNULL(m);
    if (n) *n = m->n;
    if (is_complex) *is_complex = m->is_complex;
    QIL_API_END
}

int qil_mpo_dims(const qil_mpo* m, int64_t* bond) {
    QIL_API_BEGIN
    QIL_NONNULL(m); QIL_NONNULL(bond);
    for (int i = 0; i <= m->n; ++i) bond[i] = m->bond[i];
    QIL_API_END
}

int qil_mpo_get_core(const qil_mpo* m, int site, void* host_buf) {
    QIL_API_BEGIN
    QIL_NONNULL(m); QIL_NONNULL(host_buf);
    QIL_REQUIRE(site >= 0 && site < m->n, QIL_ERR_ARGUMENT, "site %d out of range [0,%d)", site, m->n);
    QIL_CUDA(cudaSetDevice(m->ctx->device));
    QIL_CUDA(cudaMemcpyAsync(host_buf, m->core[site], m->core_elems(site) * elem_size(m->is_complex),
                             cudaMemcpyDeviceToHost, m->ctx->stream));
    m->ctx->sync();
    QIL_API_END
}

int qil_mpo_free(qil_mpo* m) {
    QIL_API_BEGIN
    if (m) {
        cudaSetDevice(m->ctx->device);
        destroy(m);
    }
    QIL_API_END
}

// ---- coefficient ---------------------------------------------------------------------------
int qil_coefficient_batch_dev(qil_ctx* ctx, const qil_mps* psi, const uint8_t* d_bits, int64_t B, void* d_out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(psi);
    QIL_REQUIRE(B >= 0, QIL_ERR_ARGUMENT, "coefficient: negative batch size");
    if (B == 0) return QIL_OK;
    QIL_NONNULL(d_bits); QIL_NONNULL(d_out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    coefficient_batch_dev(ctx, psi, d_bits, B, d_out);
    QIL_API_END
}

int qil_coefficient_batch(qil_ctx* ctx, const qil_mps* psi, const uint8_t* bits, int64_t B, void* out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(psi);
    QIL_REQUIRE(B >= 0, QIL_ERR_ARGUMENT, "coefficient: negative batch size");
    if (B == 0) return QIL_OK;
    QIL_NONNULL(bits); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    const size_t nb = (size_t)B * psi->n;
    for (size_t i = 0; i < nb; ++i)
        QIL_REQUIRE(bits[i] <= 1, QIL_ERR_ARGUMENT, "coefficient: bit value %d outside [0,1]", (int)bits[i]);
    const size_t ob = (size_t)B * elem_size(psi->is_complex);
    uint8_t* d_bits = (uint8_t*)ctx->alloc(nb);
    void* d_out = ctx->alloc(ob);
    QIL_CUDA(cudaMemcpyAsync(d_bits, bits, nb, cudaMemcpyHostToDevice, ctx->stream));
    coefficient_batch_dev(ctx, psi, d_bits, B, d_out);
    QIL_CUDA(cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    ctx->free(d_bits);
    ctx->free(d_out);
    QIL_API_END
}

static int grid_free_sites(const qil_mps* psi, const uint8_t* site_mode) {
    int F = 0;
    for (int i = 0; i < psi->n; ++i) {
        QIL_REQUIRE(site_mode[i] <= 2, QIL_ERR_ARGUMENT, "coefficient grid: site mode %d outside {0,1,2}", (int)site_mode[i]);
        F += (site_mode[i] == 2);
    }
    return F;
}

int qil_coefficient_grid_dev(qil_ctx* ctx, const qil_mps* psi, const uint8_t* site_mode, const int32_t* out_bit,
                             void* d_out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(psi); QIL_NONNULL(site_mode); QIL_NONNULL(d_out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    grid_free_sites(psi, site_mode);
    coefficient_grid_dev(ctx, psi, site_mode, out_bit, d_out);
    QIL_API_END
}

int qil_coefficient_grid(qil_ctx* ctx, const qil_mps* psi, const uint8_t* site_mode, const int32_t* out_bit, void* out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(psi); QIL_NONNULL(site_mode); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    const int F = grid_free_sites(psi, site_mode);
    QIL_REQUIRE(F <= 34, QIL_ERR_UNSUPPORTED, "coefficient grid: 2^%d results do not fit a host buffer here", F);
    const size_t ob = ((size_t)1 << F) * elem_size(psi->is_complex);
    void* d_out = ctx->alloc(ob);
    coefficient_grid_dev(ctx, psi, site_mode, out_bit, d_out);
    QIL_CUDA(cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    ctx->free(d_out);
    QIL_API_END
}

int qil_argmax_abs_dev(qil_ctx* ctx, int is_complex, const void* d_values, int64_t count, int64_t* index, double* absval,
                       void* value) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(d_values); QIL_NONNULL(index); QIL_NONNULL(absval);
    QIL_CUDA(cudaSetDevice(ctx->device));
    if (is_complex) argmax_abs<cplx>(ctx, (const cplx*)d_values, count, index, absval, (cplx*)value);
    else argmax_abs<double>(ctx, (const double*)d_values, count, index, absval, (double*)value);
    QIL_API_END
}

int qil_coefficient_grid_argmax(qil_ctx* ctx, const qil_mps* psi, const uint8_t* site_mode, const int32_t* out_bit,
                                int64_t* index, double* absval, void* value) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(psi); QIL_NONNULL(site_mode); QIL_NONNULL(index); QIL_NONNULL(absval);
    QIL_CUDA(cudaSetDevice(ctx->device));
    const int F = grid_free_sites(psi, site_mode);
    QIL_REQUIRE(F <= 34, QIL_ERR_UNSUPPORTED, "coefficient grid: 2^%d results do not fit device scratch here", F);
    const int64_t count = (int64_t)1 << F;
    void* d_out = ctx->alloc((size_t)count * elem_size(psi->is_complex));
    try {
        coefficient_grid_dev(ctx, psi, site_mode, out_bit, d_out);
        if (psi->is_complex) argmax_abs<cplx>(ctx, (const cplx*)d_out, count, index, absval, (cplx*)value);
        else argmax_abs<double>(ctx, (const double*)d_out, count, index, absval, (double*)value);
    } catch (...) {
        ctx->free(d_out);
        throw;
    }
    ctx->free(d_out);
    QIL_API_END
}

int qil_coefficient_batch_argmax(qil_ctx* ctx, const qil_mps* psi, const uint8_t* bits, int64_t B, int64_t* index,
                                 double* absval, void* value) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(psi); QIL_NONNULL(bits); QIL_NONNULL(index); QIL_NONNULL(absval);
    QIL_REQUIRE(B >= 1, QIL_ERR_ARGUMENT, "coefficient: empty batch");
    QIL_CUDA(cudaSetDevice(ctx->device));
    const size_t nb = (size_t)B * psi->n;
    for (size_t i = 0; i < nb; ++i)
        QIL_REQUIRE(bits[i] <= 1, QIL_ERR_ARGUMENT, "coefficient: bit value %d outside [0,1]", (int)bits[i]);
    uint8_t* d_bits = (uint8_t*)ctx->alloc(nb);
    void* d_out = ctx->alloc((size_t)B * elem_size(psi->is_complex));
    try {
        QIL_CUDA(cudaMemcpyAsync(d_bits, bits, nb, cudaMemcpyHostToDevice, ctx->stream));
        coefficient_batch_dev(ctx, psi, d_bits, B, d_out);
        if (psi->is_complex) argmax_abs<cplx>(ctx, (const cplx*)d_out, B, index, absval, (cplx*)value);
        else argmax_abs<double>(ctx, (const double*)d_out, B, index, absval, (double*)value);
    } catch (...) {
        ctx->free(d_bits); ctx->free(d_out);
        throw;
    }
    ctx->free(d_bits);
    ctx->free(d_out);
    QIL_API_END
}

int qil_mps_sum_sites(qil_ctx* ctx, const qil_mps* psi, const uint8_t* sum_mask, qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(psi); QIL_NONNULL(sum_mask); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    *out = mps_sum_sites(ctx, psi, sum_mask);
    QIL_API_END
}

int qil_mps_alloc(qil_ctx* ctx, int n, int is_complex, const int64_t* bond, double amplitude, qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(bond); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    qil_mps* m = new_mps(ctx, n, is_complex, bond, true);
    m->amplitude = amplitude;
    *out = m;
    QIL_API_END
}

int qil_mps_core_ptr(const qil_mps* m, int site, void** d_ptr, int64_t* elems) {
    QIL_API_BEGIN
    QIL_NONNULL(m); QIL_NONNULL(d_ptr);
    QIL_REQUIRE(site >= 0 && site < m->n, QIL_ERR_ARGUMENT, "site %d outside [0,%d)", site, m->n);
    *d_ptr = m->core[site];
    if (elems) *elems = (int64_t)m->core_elems(site);
    QIL_API_END
}


// ---- on-disk container: dims + raw cores, shared by the library, the numpy oracle and a Julia reader
//   bytes 0..7   "QILTN001"
//   uint32 kind (0 = MPS, 1 = MPO), uint32 is_complex, uint32 n, uint32 0
//   float64 amplitude (MPO: 1.0)
//   int64 bond[n+1]
//   cores back to back, C order ([l][s][r] / [l][p][s][r]), float64 or interleaved complex128, little endian
namespace {
struct FileCloser { FILE* f; ~FileCloser() { if (f) fclose(f); } };

void container_save(const char* path, int kind, int is_complex, int n, double amplitude, const std::vector<int64_t>& bond,
                    const std::vector<void*>& core, qil_ctx* ctx) {
    FileCloser fc{fopen(path, "wb")};
    QIL_REQUIRE(fc.f != nullptr, QIL_ERR_RUNTIME, "cannot open %s for writing", path);
    const char magic[8] = {'Q', 'I', 'L', 'T', 'N', '0', '0', '1'};
    const uint32_t hdr[4] = {(uint32_t)kind, (uint32_t)is_complex, (uint32_t)n, 0u};
    bool ok = fwrite(magic, 1, 8, fc.f) == 8 && fwrite(hdr, 4, 4, fc.f) == 4 && fwrite(&amplitude, 8, 1, fc.f) == 1 &&
              fwrite(bond.data(), 8, (size_t)n + 1, fc.f) == (size_t)n + 1;
    const size_t es = elem_size(is_complex), legs = kind == 0 ? 2 : 4;
    std::vector<char> buf;
    for (int i = 0; i < n && ok; ++i) {
        const size_t bytes = (size_t)bond[i] * legs * (size_t)bond[i + 1] * es;
        buf.resize(bytes);
        QIL_CUDA(cudaMemcpyAsync(buf.data(), core[i], bytes, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->sync();
        ok = fwrite(buf.data(), 1, bytes, fc.f) == bytes;
    }
    QIL_REQUIRE(ok, QIL_ERR_RUNTIME, "short write to %s", path);
}

struct ContainerData {
    int kind, is_complex, n;
    double amplitude;
    std::vector<int64_t> bond;
    std::vector<std::vector<char>> cores;
};
ContainerData container_load(const char* path) {
    FileCloser fc{fopen(path, "rb")};
    QIL_REQUIRE(fc.f != nullptr, QIL_ERR_RUNTIME, "cannot open %s", path);
    char magic[8];
    uint32_t hdr[4];
    ContainerData d;
    QIL_REQUIRE(fread(magic, 1, 8, fc.f) == 8 && memcmp(magic, "QILTN001", 8) == 0, QIL_ERR_ARGUMENT,
                "%s is not a QILTN001 container", path);
    QIL_REQUIRE(fread(hdr, 4, 4, fc.f) == 4 && fread(&d.amplitude, 8, 1, fc.f) == 1, QIL_ERR_ARGUMENT, "%s: truncated header", path);
    d.kind = (int)hdr[0]; d.is_complex = (int)hdr[1]; d.n = (int)hdr[2];
    QIL_REQUIRE(d.kind <= 1 && d.is_complex <= 1 && d.n >= 1 && d.n <= kMaxSites, QIL_ERR_ARGUMENT, "%s: bad header", path);
    d.bond.resize(d.n + 1);
    QIL_REQUIRE(fread(d.bond.data(), 8, (size_t)d.n + 1, fc.f) == (size_t)d.n + 1, QIL_ERR_ARGUMENT, "%s: truncated bonds", path);
    const size_t es = elem_size(d.is_complex), legs = d.kind == 0 ? 2 : 4;
    d.cores.resize(d.n);
    for (int i = 0; i < d.n; ++i) {
        QIL_REQUIRE(d.bond[i] >= 1 && d.bond[i + 1] >= 1 && d.bond[i] < (1ll << 30), QIL_ERR_ARGUMENT, "%s: bad bond", path);
        const size_t bytes = (size_t)d.bond[i] * legs * (size_t)d.bond[i + 1] * es;
        d.cores[i].resize(bytes);
        QIL_REQUIRE(fread(d.cores[i].data(), 1, bytes, fc.f) == bytes, QIL_ERR_ARGUMENT, "%s: truncated core %d", path, i);
    }
    return d;
}
}  // namespace

int qil_mps_save(const qil_mps* m, const char* path) {
    QIL_API_BEGIN
    QIL_NONNULL(m); QIL_NONNULL(path);
    QIL_CUDA(cudaSetDevice(m->ctx->device));
    container_save(path, 0, m->is_complex, m->n, m->amplitude, m->bond, m->core, m->ctx);
    QIL_API_END
}
int qil_mpo_save(const qil_mpo* m, const char* path) {
    QIL_API_BEGIN
    QIL_NONNULL(m); QIL_NONNULL(path);
    QIL_CUDA(cudaSetDevice(m->ctx->device));
    container_save(path, 1, m->is_complex, m->n, 1.0, m->bond, m->core, m->ctx);
    QIL_API_END
}
int qil_mps_load(qil_ctx* ctx, const char* path, qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(path); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    ContainerData d = container_load(path);
    QIL_REQUIRE(d.kind == 0, QIL_ERR_ARGUMENT, "%s holds an MPO, not an MPS", path);
    qil_mps* m = new_mps(ctx, d.n, d.is_complex, d.bond.data(), true);
    m->amplitude = d.amplitude;
    for (int i = 0; i < d.n; ++i)
        QIL_CUDA(cudaMemcpyAsync(m->core[i], d.cores[i].data(), d.cores[i].size(), cudaMemcpyHostToDevice, ctx->stream));
    ctx->sync();
    *out = m;
    QIL_API_END
}
int qil_mpo_load(qil_ctx* ctx, const char* path, qil_mpo** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(path); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    ContainerData d = container_load(path);
    QIL_REQUIRE(d.kind == 1, QIL_ERR_ARGUMENT, "%s holds an MPS, not an MPO", path);
    qil_mpo* m = new_mpo(ctx, d.n, d.is_complex, d.bond.data(), true);
    for (int i = 0; i < d.n; ++i)
        QIL_CUDA(cudaMemcpyAsync(m->core[i], d.cores[i].data(), d.cores[i].size(), cudaMemcpyHostToDevice, ctx->stream));
    ctx->sync();
    *out = m;
    QIL_API_END
}

int qil_uploader_create(qil_ctx* ctx, int64_t bytes, int depth, qil_uploader** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    *out = uploader_create(ctx, bytes, depth);
    QIL_API_END
}
int qil_uploader_submit(qil_uploader* u, const void* host, int64_t bytes) {
    QIL_API_BEGIN
    QIL_NONNULL(u); QIL_NONNULL(host);
    uploader_submit(u, host, bytes);
    QIL_API_END
}
int qil_uploader_acquire(qil_uploader* u, void** d_ptr) {
    QIL_API_BEGIN
    QIL_NONNULL(u); QIL_NONNULL(d_ptr);
    *d_ptr = uploader_acquire(u);
    QIL_API_END
}
int qil_uploader_release(qil_uploader* u) {
    QIL_API_BEGIN
    QIL_NONNULL(u);
    uploader_release(u);
    QIL_API_END
}
int qil_uploader_destroy(qil_uploader* u) {
    QIL_API_BEGIN
    uploader_destroy(u);
    QIL_API_END
}
int qil_host_register(void* host, int64_t bytes) {
    QIL_API_BEGIN
    QIL_NONNULL(host);
    QIL_CUDA(cudaHostRegister(host, (size_t)bytes, cudaHostRegisterDefault));
    QIL_API_END
}
int qil_host_unregister(void* host) {
    QIL_API_BEGIN
    QIL_NONNULL(host);
    QIL_CUDA(cudaHostUnregister(host));
    QIL_API_END
}

// ---- apply ---------------------------------------------------------------------------------
int qil_apply_mpo_mps(qil_ctx* ctx, const qil_mpo* W, const qil_mps* psi, qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(W); QIL_NONNULL(psi); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    *out = apply_mpo_mps(ctx, W, psi);
    QIL_API_END
}

int qil_apply_mpo_mps_zipup(qil_ctx* ctx, const qil_mpo* W, const qil_mps* psi, double cutoff, int64_t maxdim, qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(W); QIL_NONNULL(psi); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    *out = apply_mpo_mps_zipup(ctx, W, psi, cutoff, fix_maxdim(maxdim));
    QIL_API_END
}

int qil_apply_mpo_mps_batch(qil_ctx* ctx, const qil_mpo* W, const qil_mps* const* psis, int64_t count, qil_mps** outs) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(W);
    QIL_REQUIRE(count >= 0, QIL_ERR_ARGUMENT, "apply: negative batch size");
    if (count == 0) return QIL_OK;
    QIL_NONNULL(psis); QIL_NONNULL(outs);
    QIL_CUDA(cudaSetDevice(ctx->device));
    for (int64_t i = 0; i < count; ++i) outs[i] = nullptr;
    apply_mpo_mps_many(ctx, W, psis, count, outs);
    QIL_API_END
}

int qil_apply_mpo_mpo(qil_ctx* ctx, const qil_mpo* W1, const qil_mpo* W2, int start1, int start2,
                      qil_mpo** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(W1); QIL_NONNULL(W2); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    *out = apply_mpo_mpo(ctx, W1, W2, start1, start2);
    QIL_API_END
}

// ---- encode ----------------------------------------------------------------------------------

int qil_encode_svd_dev(qil_ctx* ctx, int is_complex, const void* d_x, int64_t N, double cutoff, int64_t maxdim,
                       qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(d_x); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    *out = is_complex ? encode_svd<cplx>(ctx, (const cplx*)d_x, N, cutoff, fix_maxdim(maxdim))
                      : encode_svd<double>(ctx, (const double*)d_x, N, cutoff, fix_maxdim(maxdim));
    QIL_API_END
}

int qil_encode_svd(qil_ctx* ctx, int is_complex, const void* x, int64_t N, double cutoff, int64_t maxdim,
                   qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(x); QIL_NONNULL(out);
    QIL_REQUIRE(N >= 1, QIL_ERR_ARGUMENT, "signal_mps: empty signal");
    QIL_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)N * elem_size(is_complex);
    void* d_x = ctx->alloc(bytes);
    QIL_CUDA(cudaMemcpyAsync(d_x, x, bytes, cudaMemcpyHostToDevice, ctx->stream));
    qil_mps* m = nullptr;
    try {
        m = is_complex ? encode_svd<cplx>(ctx, (const cplx*)d_x, N, cutoff, fix_maxdim(maxdim))
                       : encode_svd<double>(ctx, (const double*)d_x, N, cutoff, fix_maxdim(maxdim));
    } catch (...) {
        ctx->free(d_x);
        throw;
    }
    ctx->free(d_x);
    ctx->sync();
    *out = m;
    QIL_API_END
}

int qil_encode_rsvd_dev(qil_ctx* ctx, int is_complex, const void* d_x, int64_t N, int k, int p, int q, int64_t seed,
                        double cutoff, int64_t maxdim, int64_t mindim, const void* d_normal_stream,
                        int64_t stream_len, int64_t flags, qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(d_x); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    RsvdOpts o;
    o.k = k; o.p = p; o.q = q; o.seed = seed; o.cutoff = cutoff; o.maxdim = fix_maxdim(maxdim);
    o.mindim = mindim < 1 ? 1 : mindim;
    o.omega = d_normal_stream; o.omega_rows = d_normal_stream ? stream_len : 0; o.omega_cols = 1;
    o.adaptive = (flags & QIL_RSVD_ADAPTIVE) != 0;
    *out = is_complex ? encode_rsvd<cplx>(ctx, (const cplx*)d_x, N, o) : encode_rsvd<double>(ctx, (const double*)d_x, N, o);
    QIL_API_END
}

int qil_encode_rsvd(qil_ctx* ctx, int is_complex, const void* x, int64_t N, int k, int p, int q, int64_t seed,
                    double cutoff, int64_t maxdim, int64_t mindim, const void* normal_stream, int64_t stream_len,
                    int64_t flags, qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(x); QIL_NONNULL(out);
    QIL_REQUIRE(N >= 1, QIL_ERR_ARGUMENT, "signal_mps: empty signal");
    QIL_CUDA(cudaSetDevice(ctx->device));
    const size_t es = elem_size(is_complex);
    void* d_x = ctx->alloc((size_t)N * es);
    void* d_s = nullptr;
    QIL_CUDA(cudaMemcpyAsync(d_x, x, (size_t)N * es, cudaMemcpyHostToDevice, ctx->stream));
    if (normal_stream && stream_len > 0) {
        d_s = ctx->alloc((size_t)stream_len * es);
        QIL_CUDA(cudaMemcpyAsync(d_s, normal_stream, (size_t)stream_len * es, cudaMemcpyHostToDevice, ctx->stream));
    }
    int rc = qil_encode_rsvd_dev(ctx, is_complex, d_x, N, k, p, q, seed, cutoff, maxdim, mindim, d_s, stream_len,
                                 flags, out);
    ctx->free(d_x);
    if (d_s) ctx->free(d_s);
    cudaStreamSynchronize(ctx->stream);
    return rc;
    QIL_API_END
}

int qil_encode_rsvd_batch_dev(qil_ctx* ctx, int is_complex, const void* d_x, int64_t N, int64_t count, int k, int p,
                              int q, int64_t seed, double cutoff, int64_t maxdim, int64_t mindim, int workers,
                              const void* d_normal_stream, int64_t stream_len, int64_t flags, qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(out);
    QIL_REQUIRE(count >= 0, QIL_ERR_ARGUMENT, "signal batch: negative count");
    if (count == 0) return QIL_OK;
    QIL_NONNULL(d_x);
    QIL_REQUIRE(N >= 1, QIL_ERR_ARGUMENT, "signal_mps: empty signal");
    QIL_CUDA(cudaSetDevice(ctx->device));
    RsvdOpts o;
    o.k = k; o.p = p; o.q = q; o.seed = seed; o.cutoff = cutoff; o.maxdim = fix_maxdim(maxdim);
    o.mindim = mindim < 1 ? 1 : mindim;
    o.adaptive = (flags & QIL_RSVD_ADAPTIVE) != 0;
    o.omega = d_normal_stream; o.omega_rows = d_normal_stream ? stream_len : 0; o.omega_cols = 1;
    if (is_complex) encode_rsvd_batch<cplx>(ctx, (const cplx*)d_x, N, count, o, workers, out);
    else encode_rsvd_batch<double>(ctx, (const double*)d_x, N, count, o, workers, out);
    QIL_API_END
}

int qil_get_stream(qil_ctx* ctx, void** cuda_stream) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(cuda_stream);
    *cuda_stream = (void*)ctx->stream;
    QIL_API_END
}

int qil_encode_rsvd_sharded_dev(qil_ctx* ctx, const qil_comm* comm, int is_complex, const void* d_x_local,
                                int64_t N_total, int k, int p, int q, int64_t seed, double cutoff, int64_t maxdim,
                                int64_t mindim, const void* d_normal_stream, int64_t stream_len, int64_t flags,
                                qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(comm); QIL_NONNULL(d_x_local); QIL_NONNULL(out);
    QIL_NONNULL(comm->allreduce_sum_f64); QIL_NONNULL(comm->allgather_f64);
    QIL_CUDA(cudaSetDevice(ctx->device));
    RsvdOpts o;
    o.k = k; o.p = p; o.q = q; o.seed = seed; o.cutoff = cutoff; o.maxdim = fix_maxdim(maxdim);
    o.mindim = mindim < 1 ? 1 : mindim;
    o.omega = d_normal_stream; o.omega_rows = d_normal_stream ? stream_len : 0; o.omega_cols = 1;
    o.adaptive = (flags & QIL_RSVD_ADAPTIVE) != 0;
    *out = is_complex ? encode_rsvd_sharded<cplx>(ctx, comm, (const cplx*)d_x_local, N_total, o)
                      : encode_rsvd_sharded<double>(ctx, comm, (const double*)d_x_local, N_total, o);
    QIL_API_END
}

int qil_peer_create(qil_ctx* ctx, int rank, int world, int64_t bytes, qil_peer** out, unsigned char* handle64) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(out); QIL_NONNULL(handle64);
    QIL_CUDA(cudaSetDevice(ctx->device));
    *out = peer_create(ctx, rank, world, bytes, handle64);
    QIL_API_END
}

int qil_peer_connect(qil_peer* peer, const unsigned char* all_handles) {
    QIL_API_BEGIN
    QIL_NONNULL(peer); QIL_NONNULL(all_handles);
    peer_connect(peer, all_handles);
    QIL_API_END
}

int qil_peer_comm(qil_peer* peer, qil_comm* out) {
    QIL_API_BEGIN
    QIL_NONNULL(peer); QIL_NONNULL(out);
    peer_fill_comm(peer, out);
    QIL_API_END
}

int qil_peer_destroy(qil_peer* peer) {
    QIL_API_BEGIN
    peer_destroy(peer);
    QIL_API_END
}

int qil_ztmps_split(qil_ctx* ctx, const qil_mps* psi, double cutoff, int64_t maxdim, qil_mps** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(psi); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    *out = ztmps_split(ctx, psi, cutoff, fix_maxdim(maxdim));
    ctx->sync();
    QIL_API_END
}

// ---- gauge / compression -----------------------------------------------------------------------
int qil_canonicalize(qil_ctx* ctx, qil_mps* psi, int direction_right, int center, double cutoff, int64_t maxdim) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(psi);
    QIL_CUDA(cudaSetDevice(ctx->device));
    canonicalize(ctx, psi, direction_right ? 1 : 0, center, cutoff, fix_maxdim(maxdim));
    ctx->sync();
    QIL_API_END
}

int qil_compress(qil_ctx* ctx, qil_mps* psi, int64_t maxdim, double tol, int sweeps) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(psi);
    QIL_CUDA(cudaSetDevice(ctx->device));
    compress(ctx, psi, fix_maxdim(maxdim), tol, sweeps);
    ctx->sync();
    QIL_API_END
}

int qil_norm(qil_ctx* ctx, const qil_mps* psi, double* out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(psi); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    *out = mps_norm(ctx, psi);
    QIL_API_END
}

// ---- builders --------------------------------------------------------------------------------------
int qil_build_qft_mpo(qil_ctx* ctx, int n, double cutoff, int64_t maxdim, qil_mpo** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    *out = build_qft_mpo(ctx, n, cutoff, fix_maxdim(maxdim));
    ctx->sync();
    QIL_API_END
}

int qil_build_dt_mpo(qil_ctx* ctx, int n, double omega_r, double cutoff, int64_t maxdim, qil_mpo** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    *out = build_dt_mpo(ctx, n, omega_r, cutoff, fix_maxdim(maxdim));
    ctx->sync();
    QIL_API_END
}

int qil_build_zt_mpo(qil_ctx* ctx, int n, double omega_r, double cutoff, int64_t maxdim, qil_mpo** out) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(out);
    QIL_CUDA(cudaSetDevice(ctx->device));
    *out = build_zt_mpo(ctx, n, omega_r, cutoff, fix_maxdim(maxdim));
    ctx->sync();
    QIL_API_END
}

// ---- dense factorizations -----------------------------------------------------------------------
int qil_qr(qil_ctx* ctx, int is_complex, int64_t m, int64_t n, const void* A, int positive, void* Q, void* R) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(A); QIL_NONNULL(Q); QIL_NONNULL(R);
    QIL_REQUIRE(m >= 1 && n >= 1, QIL_ERR_ARGUMENT, "qr: empty matrix");
    QIL_CUDA(cudaSetDevice(ctx->device));
    if (is_complex) qr_host<cplx>(ctx, m, n, A, positive, Q, R);
    else qr_host<double>(ctx, m, n, A, positive, Q, R);
    QIL_API_END
}

int qil_svd_trunc(qil_ctx* ctx, int is_complex, int64_t m, int64_t n, const void* A, double cutoff, int64_t maxdim,
                  int64_t mindim, int64_t* rank, void* U, double* S, void* Vh) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(A); QIL_NONNULL(rank);
    QIL_REQUIRE(m >= 1 && n >= 1, QIL_ERR_ARGUMENT, "svd: empty matrix");
    QIL_CUDA(cudaSetDevice(ctx->device));
    *rank = is_complex ? svd_host<cplx>(ctx, m, n, A, cutoff, fix_maxdim(maxdim), mindim, U, S, Vh)
                       : svd_host<double>(ctx, m, n, A, cutoff, fix_maxdim(maxdim), mindim, U, S, Vh);
    QIL_API_END
}

int qil_rsvd(qil_ctx* ctx, int is_complex, int64_t m, int64_t n, const void* A, int k, int p, int q, int64_t seed,
             double cutoff, int64_t maxdim, int64_t mindim, const void* normal_stream, int64_t stream_len,
             int64_t* rank, void* U, double* S, void* Vh) {
    QIL_API_BEGIN
    QIL_NONNULL(ctx); QIL_NONNULL(rank);
    QIL_REQUIRE(m >= 1 && n >= 1, QIL_ERR_RUNTIME, "In `rsvd`, left or right index set is empty.");
    QIL_NONNULL(A);
    QIL_REQUIRE(k >= 1 && p >= 0 && q >= 0, QIL_ERR_ARGUMENT, "rsvd: k >= 1, p >= 0, q >= 0 required");
    QIL_CUDA(cudaSetDevice(ctx->device));
    RsvdOpts o;
    o.k = k; o.p = p; o.q = q; o.seed = seed; o.cutoff = cutoff;
    o.maxdim = maxdim <= 0 ? k : maxdim;   // rsvd.jl:48: maxdim defaults to k
    o.mindim = mindim < 1 ? 1 : mindim;
    *rank = is_complex ? rsvd_host<cplx>(ctx, m, n, A, o, normal_stream, stream_len, U, S, Vh)
                       : rsvd_host<double>(ctx, m, n, A, o, normal_stream, stream_len, U, S, Vh);
    QIL_API_END
}

}  // extern "C"
