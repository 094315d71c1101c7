// comm.cpp — multi-GPU inside the product (SURVEY §8b "one CUDA stream per device + one NCCL communicator", §8e).
//
// One process per GPU; every process owns one TbHandle and attaches it to a communicator with tb_comm_init. The frame
// shards with no data-path collective (rank r renders its frames or its row bands of the replicated scene); the only
// exchange step is the reduction of the float4 accumulation buffers at readback, implemented here as an NCCL
// all-gather over NVLink followed by this library's own deterministic combine kernels (csrc/cuda/reduce.cu).
//
// Two transports for that step (TbCommInfo::Transport):
//   TB_COMM_TRANSPORT_PEER   every rank maps the other ranks' buffers with CUDA IPC (same node, NVLink / NVSwitch); one
//                            kernel per rank then reads its slice from all peers, sums in fixed rank order and stores
//                            the result into all peers (reduce.cu: k_sum_peers / k_scatter_bands_peers). NCCL only
//                            carries the 64-byte IPC handles at setup and a 4-byte all-gather before and after the
//                            kernel that orders the ranks' streams.
//   TB_COMM_TRANSPORT_NCCL   ncclAllGather of whole buffers (or packed bands) + local combine kernels; used when a
//                            mapping cannot be opened (different nodes, IPC not permitted) or TB_COMM_TRANSPORT=nccl.
// Both produce the same bits.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, preferring a copy the process has already loaded, e.g. the one
// bundled with PyTorch) so that the library has no link-time dependency on it: single-GPU users never touch it.
#include <dlfcn.h>
#include <cstring>
#include <ctime>
#include <vector>
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

// orders the ranks' streams: when it completes here, every rank's stream has reached it
ncclResult_t stream_barrier(tbh::Comm* c, cudaStream_t stream) {
    return nccl().AllGather(c->barrier + c->rank, c->barrier, 1, ncclUint32, (ncclComm_t)c->nccl, stream);
}

// COLLECTIVE when peers are mapped: nobody frees a buffer another rank still has open (undefined behaviour per the
// CUDA IPC contract). If the other ranks do not show up within 10 s the exported buffers are leaked instead of freed.
void release_buffers(TbHandle* h, tbh::Comm* c) {
    bool freeExported = true;
    if (c->peer) {
        cudaStreamSynchronize(h->stream);
        for (void* p : c->imported) cudaIpcCloseMemHandle(p);
        c->imported.clear();
        freeExported = false;
        if (nccl().ok && c->nccl && stream_barrier(c, h->stream) == ncclSuccess) {
            for (int ms = 0; ms < 10000 && !freeExported; ms++) {
                if (cudaStreamQuery(h->stream) == cudaSuccess) freeExported = true;
                else { struct timespec ts = {0, 1000000}; nanosleep(&ts, nullptr); }
            }
        }
        cudaGetLastError();
        c->peer = false;
    }
    for (void* p : {(void*)c->gather, (void*)c->pack, (void*)c->peerTable, (void*)c->barrier}) if (p) cudaFree(p);
    if (freeExported) { if (c->reducedAccum) cudaFree(c->reducedAccum); if (c->reducedJittered) cudaFree(c->reducedJittered); }
    c->gather = c->pack = c->reducedAccum = c->reducedJittered = nullptr;
    c->peerTable = nullptr; c->barrier = nullptr;
    c->pixels = 0;
    c->valid = false;
}

// all-gather of a few host bytes per rank (IPC handles, flags) through a device staging buffer
int exchange_bytes(TbHandle* h, tbh::Comm* c, const void* mine, size_t bytes, std::vector<uint8_t>& all) {
    const size_t N = (size_t)c->nranks;
    all.assign(bytes * N, 0);
    uint8_t* d = nullptr;
    CUDA_OK(h, cudaMalloc((void**)&d, bytes * N));
    struct Guard { uint8_t* p; ~Guard() { cudaFree(p); } } guard{d};
    CUDA_OK(h, cudaMemcpyAsync(d + bytes * c->rank, mine, bytes, cudaMemcpyHostToDevice, h->stream));
    NCCL_OK(h, nccl().AllGather(d + bytes * c->rank, d, bytes, ncclChar, (ncclComm_t)c->nccl, h->stream));
    CUDA_OK(h, cudaMemcpyAsync(all.data(), d, bytes * N, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream));
    return TB_OK;
}

// Maps every rank's four buffers into this process. All ranks reach the same verdict (c->peer): one rank that cannot
// export or open a mapping sends everybody to the NCCL transport.
int setup_peers(TbHandle* h, tbh::Comm* c) {
    c->peer = false;
    const uint32_t N = (uint32_t)c->nranks;
    if (N < 2) return TB_OK;
    const char* t = getenv("TB_COMM_TRANSPORT");
    struct Record { cudaIpcMemHandle_t handle[4]; uint32_t ok; uint32_t pad[15]; } mine;
    memset(&mine, 0, sizeof(mine));
    mine.ok = !(t && strcmp(t, "nccl") == 0);
    void* bufs[4] = {h->st.accum, h->st.jittered, c->reducedAccum, c->reducedJittered};
    for (int b = 0; b < 4 && mine.ok; b++)
        if (cudaIpcGetMemHandle(&mine.handle[b], bufs[b]) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
    std::vector<uint8_t> all;
    int rc = exchange_bytes(h, c, &mine, sizeof(mine), all);
    if (rc != TB_OK) return rc;
    const Record* recs = (const Record*)all.data();
    uint32_t opened = 1;
    for (uint32_t r = 0; r < N; r++) opened &= recs[r].ok;
    std::vector<float4*> table(4 * (size_t)N, nullptr);
    if (opened)
        for (uint32_t r = 0; r < N; r++)
            for (int b = 0; b < 4; b++) {
                void* p = bufs[b];
                if (r != (uint32_t)c->rank) {
                    p = nullptr;
                    if (cudaIpcOpenMemHandle(&p, recs[r].handle[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = 0; p = nullptr; }
                    else c->imported.push_back(p);
                }
                table[(size_t)b * N + r] = (float4*)p;
            }
    rc = exchange_bytes(h, c, &opened, sizeof(opened), all);
    if (rc != TB_OK) return rc;
    for (uint32_t r = 0; r < N; r++) opened &= ((const uint32_t*)all.data())[r];
    CUDA_OK(h, cudaMalloc((void**)&c->barrier, 4 * N));
    CUDA_OK(h, cudaMemsetAsync(c->barrier, 0, 4 * N, h->stream));
    if (!opened) { // somebody could not: close what was opened here, once everybody is past the exchange above
        for (void* p : c->imported) cudaIpcCloseMemHandle(p);
        c->imported.clear();
        return TB_OK;
    }
    CUDA_OK(h, cudaMalloc((void**)&c->peerTable, sizeof(float4*) * table.size()));
    CUDA_OK(h, cudaMemcpyAsync(c->peerTable, table.data(), sizeof(float4*) * table.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(h, cudaStreamSynchronize(h->stream)); // `table` is a local
    c->peer = true;
    return TB_OK;
}

} // namespace

namespace tbh {
void comm_release_buffers(TbHandle* h) { if (h->comm) release_buffers(h, h->comm); }
void comm_destroy(TbHandle* h) {
    if (!h->comm) return;
    release_buffers(h, h->comm);
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
    out->Transport = h->comm->peer ? TB_COMM_TRANSPORT_PEER : TB_COMM_TRANSPORT_NCCL;
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
    if (c->pixels != n) { // (re)allocate for this resolution; every rank takes the same branch (tb_resize is per job)
        release_buffers(h, c);
        CUDA_OK(h, cudaMalloc((void**)&c->reducedAccum, 16 * n));
        CUDA_OK(h, cudaMalloc((void**)&c->reducedJittered, 16 * n));
        int rc = setup_peers(h, c);
        if (rc != TB_OK) return rc;
        if (!c->peer) { // the all-gather transport's staging
            const size_t gatherPixels = c->mode == TB_SHARD_ROWS ? 2 * chunk * N : 2 * n * N;
            CUDA_OK(h, cudaMalloc((void**)&c->gather, 16 * gatherPixels));
            if (c->mode == TB_SHARD_ROWS) CUDA_OK(h, cudaMalloc((void**)&c->pack, 16 * 2 * chunk));
        }
        c->pixels = n;
    }
    ncclComm_t comm = (ncclComm_t)c->nccl;
    CUDA_OK(h, cudaEventRecord(c->ev0, h->stream));
    if (c->peer) {
        // barrier: every rank's frames are complete and nobody still reads the previous job-wide image; one kernel that
        // loads from / stores to all peers; barrier: every rank's part has landed everywhere
        NCCL_OK(h, stream_barrier(c, h->stream));
        if (c->mode == TB_SHARD_ROWS) CUDA_OK(h, scatter_bands_peers(c->peerTable, N, (uint32_t)c->rank, h->width, h->height, h->numSMs, h->stream, h->lc));
        else CUDA_OK(h, sum_peers(c->peerTable, N, (uint32_t)c->rank, n, h->numSMs, h->stream, h->lc));
        NCCL_OK(h, stream_barrier(c, h->stream));
        // arriving over NVLink per rank: rows: the other ranks' bands; samples: the other ranks' parts of this rank's
        // slice (loads) + the other ranks' finished slices (their stores)
        c->bytesPerReduction = c->mode == TB_SHARD_ROWS ? 16ull * 2 * (n - (chunk < n ? chunk : n)) : 16ull * 2 * 2 * ((n + N - 1) / N) * (N - 1);
    } else if (c->mode == TB_SHARD_ROWS) {
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
