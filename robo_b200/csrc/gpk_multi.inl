// gpk_multi.inl — included at the end of gpk_api.cu (after its extern "C" block): entry points that span several handles
// (GP-MCMC sub-models) or several GPUs (candidate shards + the arg-max exchange).
//
// NCCL is bound at run time (dlopen), so libgpk.so itself links cudart only and loads on machines without NCCL;
// gpk_comm_* return GPK_CUDA_ERROR with a message there.  If the process already holds a libnccl.so.2 (e.g. the one
// bundled with PyTorch) that copy is used, otherwise the system library.

// ---------------------------------------------------------------------------------------
// several fitted models, one candidate batch (second half)
// ---------------------------------------------------------------------------------------
__global__ void gpk_values_argmax_kernel(const double* __restrict__ v, long m, BestPair* __restrict__ block_best)
{
    const long c = (long)blockIdx.x * 256 + threadIdx.x;
    double val = 0.0;
    long long idx = -1;
    if (c < m) { val = v[c]; idx = c; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, val, off);
        long long oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (gpk_better(ov, oi, val, idx)) { val = ov; idx = oi; }
    }
    __shared__ double sv[8];
    __shared__ long long si[8];
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = val; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (gpk_better(sv[w], si[w], val, idx)) { val = sv[w]; idx = si[w]; }
        block_best[blockIdx.x].val = val;
        block_best[blockIdx.x].idx = idx;
    }
}

extern "C" int gpk_acq_multi(gpk_handle* const* hs, int n_models, const double* Xs, long m, int mode, int kind,
                  const double* eta, double par, double* out1, double* out2, long* n_negative, double* best_val,
                  long* best_idx) {
    if (!hs || n_models <= 0 || !hs[0]) return GPK_BAD_ARG;
    gpk_handle* h = hs[0];                      // owner of the shared buffers; errors are reported on it
    if (!Xs || m <= 0 || !out1 || (mode != 0 && mode != 1)) BAD("gpk_acq_multi: bad arguments");
    if (mode == 1 && !out2) BAD("gpk_acq_multi: mode 1 needs out2");
    if (mode == 0 && (kind < GPK_ACQ_EI || kind > GPK_ACQ_LCB || !eta)) BAD("gpk_acq_multi: mode 0 needs an acquisition and eta[n_models]");
    for (int i = 0; i < n_models; ++i) {
        gpk_handle* g = hs[i];
        int rc = require(g, true, true, true);
        if (rc) { set_err(h, "gpk_acq_multi: model %d: %s", i, g ? g->err : "null handle"); return rc ? rc : GPK_BAD_ARG; }
        if (g->device != h->device || g->d != h->d) BAD("gpk_acq_multi: model %d lives on another device or has another input dimension", i);
        if (g->fit_pending) BAD("gpk_acq_multi: model %d has a pending gpk_fit_begin", i);
        for (int j = 0; j < i; ++j)
            if (hs[j] == g) BAD("gpk_acq_multi: handle %d listed twice", i);
    }
    CK(cudaSetDevice(h->device));
    int rc;
    const size_t bytes = (size_t)n_models * m * 8;
    if ((rc = ensure(h, h->multi_cand, (size_t)m * h->d * 8))) return rc;
    if ((rc = ensure(h, h->multi_A, bytes))) return rc;
    if ((rc = ensure(h, h->multi_B, mode == 1 ? bytes : 8))) return rc;
    if ((rc = ensure(h, h->multi_out, (size_t)m * 16))) return rc;
    if ((rc = ensure(h, h->multi_bb, (size_t)((m + 255) / 256 + 1) * sizeof(BestPair)))) return rc;
    if ((rc = ensure(h, h->nneg, 8))) return rc;
    if ((rc = ensure(h, h->best, sizeof(BestPair)))) return rc;
    if (!h->ev_multi) CK(cudaEventCreateWithFlags(&h->ev_multi, cudaEventDisableTiming));
    double* A = ptr<double>(h->multi_A);
    double* B = ptr<double>(h->multi_B);
    unsigned long long* d_nneg = ptr<unsigned long long>(h->nneg);
    // candidates: one H2D for all models (pageable or pinned; m * d * 8 bytes)
    CK(cudaMemcpyAsync(h->multi_cand.p, Xs, (size_t)m * h->d * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(d_nneg, 0, 8, h->stream));
    CK(cudaEventRecord(h->ev_multi, h->stream));
    // every model scores the batch on its own stream: the small launches of the sub-models overlap on the GPU
    for (int i = 0; i < n_models; ++i) {
        gpk_handle* g = hs[i];
        if (g != h) CK(cudaStreamWaitEvent(g->stream, h->ev_multi, 0));
        if (mode == 0)
            rc = score_dev(g, ptr<double>(h->multi_cand), m, kind, eta[i], par, A + (size_t)i * m, nullptr, nullptr, nullptr,
                           d_nneg, 0, /*reset=*/false);
        else
            rc = score_dev(g, ptr<double>(h->multi_cand), m, GPK_ACQ_NONE, 0.0, 0.0, nullptr, A + (size_t)i * m,
                           B + (size_t)i * m, nullptr, d_nneg, 0, /*reset=*/false);
        if (rc) { if (g != h) set_err(h, "gpk_acq_multi: model %d: %s", i, g->err); return rc; }
        if (g != h) {
            cudaError_t e_ = cudaEventRecord(g->ev_order, g->stream);
            if (e_ == cudaSuccess) e_ = cudaStreamWaitEvent(h->stream, g->ev_order, 0);
            if (e_ != cudaSuccess) { set_err(h, "gpk_acq_multi: event hand-off -> %s", cudaGetErrorString(e_)); return GPK_CUDA_ERROR; }
        }
    }
    double* o1 = ptr<double>(h->multi_out);
    double* o2 = o1 + m;
    gpk_reduce_models_kernel<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(A, mode == 1 ? B : nullptr, n_models, m, mode,
                                                                                 o1, o2);
    CKL();
    BestPair bp;
    bp.val = 0.0; bp.idx = -1;
    if (mode == 0 && (best_val || best_idx)) {
        const int fb = (int)((m + 255) / 256);
        CK(cudaMemsetAsync(h->best.p, 0xFF, sizeof(BestPair), h->stream));
        gpk_values_argmax_kernel<<<fb, 256, 0, h->stream>>>(o1, m, ptr<BestPair>(h->multi_bb));
        CKL();
        gpk_argmax_final_kernel<<<1, 256, 0, h->stream>>>(ptr<BestPair>(h->multi_bb), fb, ptr<BestPair>(h->best));
        CKL();
        CK(cudaMemcpyAsync(&bp, h->best.p, sizeof(bp), cudaMemcpyDeviceToHost, h->stream));
    }
    unsigned long long nn = 0;
    CK(cudaMemcpyAsync(out1, o1, (size_t)m * 8, cudaMemcpyDeviceToHost, h->stream));
    if (mode == 1) CK(cudaMemcpyAsync(out2, o2, (size_t)m * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&nn, d_nneg, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (n_negative) *n_negative = (long)nn;
    if (best_val) *best_val = bp.val;
    if (best_idx) *best_idx = (long)bp.idx;
    return GPK_OK;
}

// ---------------------------------------------------------------------------------------
// multi-GPU: candidate shards + one 16-byte all-gather per arg-max
// ---------------------------------------------------------------------------------------
namespace {

typedef struct { char internal[128]; } gpk_nccl_id;
typedef int (*nccl_get_unique_id_fn)(gpk_nccl_id*);
typedef int (*nccl_comm_init_rank_fn)(void**, int, gpk_nccl_id, int);
typedef int (*nccl_comm_destroy_fn)(void*);
typedef int (*nccl_all_gather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*nccl_get_error_string_fn)(int);
typedef int (*nccl_get_version_fn)(int*);

struct NcclApi {
    void* lib = nullptr;
    nccl_get_unique_id_fn get_unique_id = nullptr;
    nccl_comm_init_rank_fn comm_init_rank = nullptr;
    nccl_comm_destroy_fn comm_destroy = nullptr;
    nccl_all_gather_fn all_gather = nullptr;
    nccl_get_error_string_fn get_error_string = nullptr;
    nccl_get_version_fn get_version = nullptr;
    char why[256] = {0};
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char* env = getenv("GPK_NCCL_LIB");
    void* lib = nullptr;
    if (env && *env) lib = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy this process already uses
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!lib) {
        snprintf(api.why, sizeof(api.why), "libnccl.so.2 not found (%s)", dlerror());
        return &api;
    }
    api.get_unique_id = (nccl_get_unique_id_fn)dlsym(lib, "ncclGetUniqueId");
    api.comm_init_rank = (nccl_comm_init_rank_fn)dlsym(lib, "ncclCommInitRank");
    api.comm_destroy = (nccl_comm_destroy_fn)dlsym(lib, "ncclCommDestroy");
    api.all_gather = (nccl_all_gather_fn)dlsym(lib, "ncclAllGather");
    api.get_error_string = (nccl_get_error_string_fn)dlsym(lib, "ncclGetErrorString");
    api.get_version = (nccl_get_version_fn)dlsym(lib, "ncclGetVersion");
    if (!api.get_unique_id || !api.comm_init_rank || !api.comm_destroy || !api.all_gather) {
        snprintf(api.why, sizeof(api.why), "libnccl.so.2 lacks ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllGather");
        return &api;
    }
    api.lib = lib;
    return &api;
}

#define CKN(call)                                                                                      \
    do {                                                                                               \
        int r_ = (call);                                                                               \
        if (r_ != 0) {                                                                                 \
            NcclApi* a_ = nccl_api();                                                                  \
            set_err(h, "%s -> NCCL error %d (%s)", #call, r_,                                          \
                    a_->get_error_string ? a_->get_error_string(r_) : "?");                            \
            return GPK_CUDA_ERROR;                                                                     \
        }                                                                                              \
    } while (0)

// deterministic merge of the gathered {value, global index} pairs: numpy.argmax ordering (NaN first, larger value,
// lowest index); one warp, lane r holds rank r's pair (world <= 32 per step, looped otherwise)
__global__ void gpk_merge_best_kernel(const BestPair* __restrict__ pairs, int world, BestPair* __restrict__ out)
{
    double val = 0.0;
    long long idx = -1;
    for (int r = threadIdx.x; r < world; r += 32)
        if (gpk_better(pairs[r].val, pairs[r].idx, val, idx)) { val = pairs[r].val; idx = pairs[r].idx; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, val, off);
        long long oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (gpk_better(ov, oi, val, idx)) { val = ov; idx = oi; }
    }
    if (threadIdx.x == 0) { out->val = val; out->idx = idx; }
}

__global__ void gpk_shift_index_kernel(BestPair* p, long long shift) {
    if (p->idx >= 0) p->idx += shift;
}

// rank r's contiguous slice [lo, hi) of m units (sizes differ by at most one; robo_b200/distributed.py:shard_bounds)
inline void shard_range(long m, int rank, int world, long* lo, long* hi) {
    const long base = m / world, rem = m % world;
    *lo = rank * base + std::min<long>(rank, rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
}

// local best (h->best, global index) -> all ranks -> merged pair in d_out (device, 16 bytes); asynchronous
int exchange_best(gpk_handle* h, const BestPair* d_local, BestPair* d_out) {
    if (h->world <= 1 || h->comm == nullptr) {
        if (d_out != d_local) CK(cudaMemcpyAsync(d_out, d_local, sizeof(BestPair), cudaMemcpyDeviceToDevice, h->stream));
        return GPK_OK;
    }
    NcclApi* api = nccl_api();
    int rc = ensure(h, h->gather, (size_t)h->world * sizeof(BestPair));
    if (rc) return rc;
    CKN(api->all_gather(d_local, h->gather.p, 2, /*ncclUint64*/ 5, h->comm, h->stream));
    gpk_merge_best_kernel<<<1, 32, 0, h->stream>>>(ptr<BestPair>(h->gather), h->world, d_out);
    CKL();
    return GPK_OK;
}

}  // namespace

extern "C" int gpk_comm_unique_id(void* id128) {
    if (!id128) return GPK_BAD_ARG;
    NcclApi* api = nccl_api();
    if (!api->lib) { fprintf(stderr, "gpk_comm_unique_id: %s\n", api->why); return GPK_CUDA_ERROR; }
    gpk_nccl_id id;
    memset(&id, 0, sizeof(id));
    if (api->get_unique_id(&id) != 0) return GPK_CUDA_ERROR;
    memcpy(id128, &id, sizeof(id));
    return GPK_OK;
}

extern "C" int gpk_comm_init(gpk_handle* h, int rank, int world, const void* id128) {
    if (!h) return GPK_BAD_ARG;
    if (world < 1 || rank < 0 || rank >= world) BAD("gpk_comm_init: need 0 <= rank < world");
    if (h->comm) BAD("gpk_comm_init: this handle already has a communicator (gpk_comm_destroy first)");
    CK(cudaSetDevice(h->device));
    h->rank = rank;
    h->world = world;
    if (world == 1) return GPK_OK;                      // nothing to exchange
    if (!id128) BAD("gpk_comm_init: world > 1 needs the 128-byte id from gpk_comm_unique_id (rank 0)");
    NcclApi* api = nccl_api();
    if (!api->lib) { set_err(h, "gpk_comm_init: %s", api->why); h->world = 1; h->rank = 0; return GPK_CUDA_ERROR; }
    gpk_nccl_id id;
    memcpy(&id, id128, sizeof(id));
    void* comm = nullptr;
    int r = api->comm_init_rank(&comm, world, id, rank);
    if (r != 0) {
        set_err(h, "ncclCommInitRank -> NCCL error %d (%s)", r, api->get_error_string ? api->get_error_string(r) : "?");
        h->world = 1; h->rank = 0;
        return GPK_CUDA_ERROR;
    }
    h->comm = comm;
    int rc = ensure(h, h->gather, (size_t)world * sizeof(BestPair));
    if (rc) return rc;
    if ((rc = ensure(h, h->best, sizeof(BestPair)))) return rc;
    if ((rc = ensure(h, h->best_global, sizeof(BestPair)))) return rc;
    return GPK_OK;
}

extern "C" int gpk_comm_destroy(gpk_handle* h) {
    if (!h) return GPK_BAD_ARG;
    if (h->comm) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        NcclApi* api = nccl_api();
        if (api->lib) api->comm_destroy(h->comm);
        h->comm = nullptr;
    }
    h->rank = 0;
    h->world = 1;
    return GPK_OK;
}

extern "C" int gpk_comm_info(gpk_handle* h, int* rank, int* world, int* nccl_version) {
    if (!h) return GPK_BAD_ARG;
    if (rank) *rank = h->rank;
    if (world) *world = h->world;
    if (nccl_version) {
        *nccl_version = 0;
        NcclApi* api = nccl_api();
        if (api->lib && api->get_version) api->get_version(nccl_version);
    }
    return GPK_OK;
}

extern "C" int gpk_shard_bounds(long m, int rank, int world, long* lo, long* hi) {
    if (m < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return GPK_BAD_ARG;
    shard_range(m, rank, world, lo, hi);
    return GPK_OK;
}

extern "C" int gpk_comm_argmax_pair(gpk_handle* h, double val, long idx, double* best_val, long* best_idx) {
    if (!h) return GPK_BAD_ARG;
    CK(cudaSetDevice(h->device));
    int rc;
    if ((rc = ensure(h, h->best, sizeof(BestPair)))) return rc;
    if ((rc = ensure(h, h->best_global, sizeof(BestPair)))) return rc;
    BestPair in;
    in.val = val; in.idx = idx;
    CK(cudaMemcpyAsync(h->best.p, &in, sizeof(in), cudaMemcpyHostToDevice, h->stream));
    if ((rc = exchange_best(h, ptr<BestPair>(h->best), ptr<BestPair>(h->best_global)))) return rc;
    BestPair bp;
    CK(cudaMemcpyAsync(&bp, h->best_global.p, sizeof(bp), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (best_val) *best_val = bp.val;
    if (best_idx) *best_idx = (long)bp.idx;
    return GPK_OK;
}

extern "C" int gpk_acq_argmax_sharded_dev(gpk_handle* h, const void* d_Xs_shard, long m_shard, long first_global, int kind,
                               double eta, double par, void* d_best) {
    int rc = require(h, true, true, true);
    if (rc) return rc;
    if (m_shard < 0 || first_global < 0 || (m_shard > 0 && !d_Xs_shard)) BAD("gpk_acq_argmax_sharded_dev: bad shard");
    if (kind < GPK_ACQ_EI || kind > GPK_ACQ_LCB) BAD("gpk_acq_argmax_sharded_dev: unknown acquisition %d", kind);
    CK(cudaSetDevice(h->device));
    if ((rc = ensure(h, h->best, sizeof(BestPair)))) return rc;
    if ((rc = ensure(h, h->best_global, sizeof(BestPair)))) return rc;
    if (m_shard > 0) {
        if ((rc = score_dev(h, (const double*)d_Xs_shard, m_shard, kind, eta, par, nullptr, nullptr, nullptr, nullptr, nullptr,
                            0, true, first_global)))
            return rc;
    } else {
        CK(cudaMemsetAsync(h->best.p, 0xFF, sizeof(BestPair), h->stream));     // empty shard: index -1 loses every merge
    }
    return exchange_best(h, ptr<BestPair>(h->best), d_best ? (BestPair*)d_best : ptr<BestPair>(h->best_global));
}

extern "C" int gpk_acq_argmax_sharded(gpk_handle* h, const double* Xs, long m_total, int kind, double eta, double par,
                           double* best_val, long* best_idx) {
    int rc = require(h, true, true, true);
    if (rc) return rc;
    if (!Xs || m_total <= 0) BAD("gpk_acq_argmax_sharded: need the full candidate batch (identical on every rank)");
    if (kind < GPK_ACQ_EI || kind > GPK_ACQ_LCB) BAD("gpk_acq_argmax_sharded: unknown acquisition %d", kind);
    CK(cudaSetDevice(h->device));
    long lo, hi;
    shard_range(m_total, h->rank, h->world, &lo, &hi);
    const long m = hi - lo;
    if ((rc = ensure(h, h->best, sizeof(BestPair)))) return rc;
    if ((rc = ensure(h, h->best_global, sizeof(BestPair)))) return rc;
    if (m > 0) {
        // this rank's slice goes through the host-batch path (staging of pageable memory, piecewise overlap with the
        // scoring); the arg-max it leaves in h->best is shard-local, so it is shifted to the global index on the device
        double bv;
        long bi, nn;
        if ((rc = gpk_acq(h, Xs + lo * h->d, m, kind, eta, par, nullptr, nullptr, nullptr, &bv, &bi, &nn))) return rc;
        gpk_shift_index_kernel<<<1, 1, 0, h->stream>>>(ptr<BestPair>(h->best), (long long)lo);
        CKL();
    } else {
        CK(cudaMemsetAsync(h->best.p, 0xFF, sizeof(BestPair), h->stream));
    }
    if ((rc = exchange_best(h, ptr<BestPair>(h->best), ptr<BestPair>(h->best_global)))) return rc;
    BestPair bp;
    CK(cudaMemcpyAsync(&bp, h->best_global.p, sizeof(bp), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (best_val) *best_val = bp.val;
    if (best_idx) *best_idx = (long)bp.idx;
    return GPK_OK;
}

extern "C" int gpk_maximize_random_sharded(gpk_handle* h, unsigned long long seed, long n_total, long n_uniform, const double* lower,
                                const double* upper, const double* incumbent, double scale, int kind, double eta,
                                double par, double* best_x, double* best_val, long* best_idx) {
    int rc = require(h, true, true, true);
    if (rc) return rc;
    if (!lower || !upper || !incumbent || n_total <= 0) BAD("gpk_maximize_random_sharded: bad arguments");
    if (kind < GPK_ACQ_EI || kind > GPK_ACQ_LCB) BAD("gpk_maximize_random_sharded: unknown acquisition %d", kind);
    CK(cudaSetDevice(h->device));
    long lo, hi;
    shard_range(n_total, h->rank, h->world, &lo, &hi);
    const long m = hi - lo;
    if ((rc = ensure(h, h->best, sizeof(BestPair)))) return rc;
    if ((rc = ensure(h, h->best_global, sizeof(BestPair)))) return rc;
    if (m > 0) {
        if ((rc = generate_candidates(h, seed, lo, m, n_uniform, h->d, lower, upper, incumbent, scale))) return rc;
        if ((rc = score_dev(h, ptr<double>(h->cand), m, kind, eta, par, nullptr, nullptr, nullptr, nullptr, nullptr, 0, true, lo)))
            return rc;
    } else {
        CK(cudaMemsetAsync(h->best.p, 0xFF, sizeof(BestPair), h->stream));
    }
    if ((rc = exchange_best(h, ptr<BestPair>(h->best), ptr<BestPair>(h->best_global)))) return rc;
    BestPair bp;
    CK(cudaMemcpyAsync(&bp, h->best_global.p, sizeof(bp), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (bp.idx >= 0 && best_x) {
        // the winner is re-created from its global index (Philox is keyed by it): identical on every rank, no broadcast
        if (bp.idx >= lo && bp.idx < hi) {
            CK(cudaMemcpy(best_x, ptr<double>(h->cand) + (bp.idx - lo) * h->d, (size_t)h->d * 8, cudaMemcpyDeviceToHost));
        } else {
            if ((rc = generate_candidates(h, seed, (long)bp.idx, 1, n_uniform, h->d, lower, upper, incumbent, scale))) return rc;
            CK(cudaMemcpyAsync(best_x, h->cand.p, (size_t)h->d * 8, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
    }
    if (best_val) *best_val = bp.val;
    if (best_idx) *best_idx = (long)bp.idx;
    return GPK_OK;
}
