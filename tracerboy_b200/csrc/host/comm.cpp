// comm.cpp — multi-GPU inside the product (SURVEY §8b "one CUDA stream per device + one NCCL communicator", §8e).
//
// One process per GPU; every process owns one TbHandle and attaches it to a communicator with tb_comm_init. The frame
// shards with no data-path collective (rank r renders its frames or its row bands of the replicated scene); the only
// exchange step is the reduction of the float4 accumulation buffers at readback, implemented here as an NCCL
// all-gather over NVLink followed by this library's own deterministic combine kernels (csrc/cuda/reduce.cu).
//
// NCCL is bound at run time (dlopen of libnccl.so.2, preferring a copy the process has already loaded, e.g. the one
// bundled with PyTorch) so that the library has no link-time dependency on it: single-GPU users never touch it.
#include <dlfcn.h>
#include <cstring>
#include <nccl.h>
#include "../cuda/reduce.h"
#include "handle.h"

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string error;
    bool ok = false;
};

NcclApi load_nccl() {
    NcclApi api;
    const char* override_ = getenv("TB_NCCL_LIB");
    void* lib = nullptr;
    if (override_) lib = dlopen(override_, RTLD_NOW | RTLD_LOCAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // the copy this process already uses (PyTorch's)
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!lib) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return api; }
    auto sym = [&](const char* name) { void* p = dlsym(lib, name); if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + name; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.ok = api.error.empty();
    return api;
}
NcclApi& nccl() {
    static NcclApi api = load_nccl(); // bound once, by whichever thread gets here first
    return api;
}

#define NCCL_OK(h, call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return fail(h, TB_ERR_NCCL, std::string(#call) + ": " + nccl().GetErrorString(r__)); } while (0)

void release_buffers(tbh::Comm* c) {
    for (void* p : {(void*)c->gather, (void*)c->pack, (void*)c->reducedAccum, (void*)c->reducedJittered}) if (p) cudaFree(p);
    c->gather = c->pack = c->reducedAccum = c->reducedJittered = nullptr;
    c->pixels = 0;
    c->valid = false;
}

} // namespace

namespace tbh {
void comm_release_buffers(TbHandle* h) { if (h->comm) release_buffers(h->comm); }
void comm_destroy(TbHandle* h) {
    if (!h->comm) return;
    release_buffers(h->comm);
    if (h->comm->nccl && nccl().ok) nccl().CommDestroy((ncclComm_t)h->comm->nccl);
    if (h->comm->ev0) cudaEventDestroy(h->comm->ev0);
    if (h->comm->ev1) cudaEventDestroy(h->comm->ev1);
    delete h->comm;
    h->comm = nullptr;
}
} // namespace tbh

extern "C" {

TB_API int tb_comm_get_unique_id(void* id, uint64_t bytes) {
    static_assert(sizeof(ncclUniqueId) == TB_COMM_ID_BYTES, "TB_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
    if (!id || bytes < sizeof(ncclUniqueId)) return fail(nullptr, TB_ERR_INVALID_ARG, "id buffer must hold TB_COMM_ID_BYTES bytes");
    if (!nccl().ok) return fail(nullptr, TB_ERR_NCCL, nccl().error);
    ncclUniqueId u;
    NCCL_OK(nullptr, nccl().GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return TB_OK;
}

TB_API int tb_comm_init(TbHandle* h, const void* id, int rank, int nranks, uint32_t shardMode) {
    if (!h || !id) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(h, TB_ERR_INVALID_ARG, "need 0 <= rank < nranks");
    if (shardMode != TB_SHARD_SAMPLES && shardMode != TB_SHARD_ROWS) return fail(h, TB_ERR_INVALID_ARG, "shard mode must be TB_SHARD_SAMPLES or TB_SHARD_ROWS");
    if (h->comm) return fail(h, TB_ERR_STATE, "the handle already has a communicator (tb_comm_destroy first)");
    if (!nccl().ok) return fail(h, TB_ERR_NCCL, nccl().error);
    CUDA_OK(h, cudaSetDevice(h->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t comm = nullptr;
    NCCL_OK(h, nccl().CommInitRank(&comm, nranks, u, rank));
    tbh::Comm* c = new tbh::Comm();
    c->nccl = comm; c->rank = rank; c->nranks = nranks; c->mode = shardMode;
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
    h->comm = c;
    int rc = shardMode == TB_SHARD_ROWS ? tb_set_row_shard(h, (uint32_t)rank, (uint32_t)nranks) : tb_set_frame_shard(h, (uint32_t)rank, (uint32_t)nranks);
    if (rc == TB_OK) rc = shardMode == TB_SHARD_ROWS ? tb_set_frame_shard(h, 0, 1) : tb_set_row_shard(h, 0, 1);
    if (rc != TB_OK) tbh::comm_destroy(h);
    return rc;
}

TB_API int tb_comm_destroy(TbHandle* h) {
    if (!h) return TB_ERR_INVALID_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    tbh::comm_destroy(h);
    return TB_OK;
}

TB_API int tb_comm_info(TbHandle* h, TbCommInfo* out) {
    if (!h || !out) return fail(h, TB_ERR_INVALID_ARG, "null argument");
    memset(out, 0, sizeof(*out));
    out->NumRanks = 1;
    if (!h->comm) return TB_OK;
    out->Rank = (uint32_t)h->comm->rank; out->NumRanks = (uint32_t)h->comm->nranks; out->ShardMode = h->comm->mode;
    out->Reductions = h->comm->reductions; out->BytesReceivedPerReduction = h->comm->bytesPerReduction;
    out->LastReductionMilliseconds = h->comm->lastMs; out->TotalReductionMilliseconds = h->comm->totalMs;
    if (nccl().ok) { int v = 0; if (nccl().GetVersion(&v) == ncclSuccess) out->NcclVersion = (uint32_t)v; }
    return TB_OK;
}

// The collective: every rank of the communicator calls it (tb_readback of an accumulation-derived buffer does so
// implicitly when the reduced image is stale). Enqueued on the handle's stream behind the frames rendered so far.
TB_API int tb_comm_reduce(TbHandle* h) {
    if (!h) return TB_ERR_INVALID_ARG;
    tbh::Comm* c = h->comm;
    if (!c) return fail(h, TB_ERR_STATE, "no communicator (tb_comm_init)");
    if (!h->width) return fail(h, TB_ERR_STATE, "no frame buffers (tb_resize)");
    CUDA_OK(h, cudaSetDevice(h->device));
    const size_t n = (size_t)h->width * h->height;
    const uint32_t N = (uint32_t)c->nranks;
    const size_t chunk = band_chunk_pixels(h->width, h->height, N);
    if (c->pixels != n) { // (re)allocate for this resolution
        release_buffers(c);
        const size_t gatherPixels = c->mode == TB_SHARD_ROWS ? 2 * chunk * N : 2 * n * N;
        CUDA_OK(h, cudaMalloc((void**)&c->gather, 16 * gatherPixels));
        if (c->mode == TB_SHARD_ROWS) CUDA_OK(h, cudaMalloc((void**)&c->pack, 16 * 2 * chunk));
        CUDA_OK(h, cudaMalloc((void**)&c->reducedAccum, 16 * n));
        CUDA_OK(h, cudaMalloc((void**)&c->reducedJittered, 16 * n));
        c->pixels = n;
    }
    ncclComm_t comm = (ncclComm_t)c->nccl;
    CUDA_OK(h, cudaEventRecord(c->ev0, h->stream));
    if (c->mode == TB_SHARD_ROWS) {
        // my bands of both buffers -> one packed chunk -> all-gather -> scatter every rank's bands back into place
        CUDA_OK(h, pack_bands(h->st.accum, h->width, h->height, (uint32_t)c->rank, N, c->pack, h->numSMs, h->stream, h->lc));
        CUDA_OK(h, pack_bands(h->st.jittered, h->width, h->height, (uint32_t)c->rank, N, c->pack + chunk, h->numSMs, h->stream, h->lc));
        NCCL_OK(h, nccl().AllGather(c->pack, c->gather, 2 * chunk * 4, ncclFloat, comm, h->stream));
        CUDA_OK(h, unpack_bands(c->gather, 2 * chunk, h->width, h->height, N, c->reducedAccum, h->numSMs, h->stream, h->lc));
        CUDA_OK(h, unpack_bands(c->gather + chunk, 2 * chunk, h->width, h->height, N, c->reducedJittered, h->numSMs, h->stream, h->lc));
        c->bytesPerReduction = 16ull * 2 * chunk * (N - 1);
    } else {
        // whole buffers of every rank, then the fixed-order sum (rank 0 + rank 1 + ...): identical bits on every rank
        NCCL_OK(h, nccl().GroupStart());
        ncclResult_t r1 = nccl().AllGather(h->st.accum, c->gather, n * 4, ncclFloat, comm, h->stream);
        ncclResult_t r2 = nccl().AllGather(h->st.jittered, c->gather + n * N, n * 4, ncclFloat, comm, h->stream);
        NCCL_OK(h, nccl().GroupEnd());
        NCCL_OK(h, r1); NCCL_OK(h, r2);
        CUDA_OK(h, sum_ranks(c->gather, N, n, c->reducedAccum, h->numSMs, h->stream, h->lc));
        CUDA_OK(h, sum_ranks(c->gather + n * N, N, n, c->reducedJittered, h->numSMs, h->stream, h->lc));
        c->bytesPerReduction = 16ull * 2 * n * (N - 1);
    }
    CUDA_OK(h, cudaEventRecord(c->ev1, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->lastMs = ms;
    c->totalMs += ms;
    c->reductions++;
    c->valid = true;
    return TB_OK;
}

} // extern "C"
