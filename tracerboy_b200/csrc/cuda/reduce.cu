// reduce.cu — the path's only multi-GPU exchange step: after an NCCL all-gather of the per-rank float4
// accumulation buffers, combine them into the job-wide image, deterministically.
//
//   sample-index sharding (rank r rendered frames r, r+N, ...): out = ((g0 + g1) + g2) + ... in fixed rank order, so
//     the result depends on N but never on timing, and every rank holds the same bits (ncclAllReduce's ring / tree
//     order is an implementation detail of the library version and topology).
//   row-band sharding (bands of 8 rows, band b on rank b mod N): every pixel was computed entirely by one rank; each
//     rank packs its own bands, the packed chunks are all-gathered (1/N of the image per rank instead of N whole
//     images) and scattered back: bit-identical to the single-GPU render.
//
// All three kernels are pure HBM streams of 128-bit loads / stores: 16 B in + 16 B out per pixel (pack / unpack),
// 16 N B in + 16 B out per pixel (sum).
//
// Peer-memory transport (comm.cpp maps every rank's buffers into every process with CUDA IPC): the exchange and the
// combine are ONE kernel per rank that loads from and stores to the other GPUs' memory over NVLink.
//   k_sum_peers      rank r owns pixels [r n/N, (r+1) n/N): it reads that slice of every rank's accumulation buffers,
//                    adds them in the same fixed rank order and stores the sum into every rank's result buffers
//                    (a reduce-scatter and an all-gather in one pass: 2 (N-1)/N of a buffer crosses NVLink per rank
//                    and direction, instead of N-1 whole buffers received by the all-gather transport).
//   k_scatter_bands_peers  row bands: a rank stores its own bands straight into every rank's result buffers.
#include "reduce.h"

namespace tbd {
namespace {

__global__ void __launch_bounds__(256) k_sum_ranks(const float4* __restrict__ gathered, uint32_t nranks, size_t n, float4* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = gathered[i];
        for (uint32_t r = 1; r < nranks; r++) {
            const float4 b = gathered[(size_t)r * n + i];
            a.x = a.x + b.x; a.y = a.y + b.y; a.z = a.z + b.z; a.w = a.w + b.w;
        }
        out[i] = a;
    }
}

// chunk layout of one rank: its j-th band (image band j * stride + offset) occupies rows [8 j, 8 j + 8)
__global__ void __launch_bounds__(256) k_pack_bands(const float4* __restrict__ src, uint32_t width, uint32_t height, uint32_t offset, uint32_t stride,
                                                    size_t chunkPixels, float4* __restrict__ dst) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < chunkPixels; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t row = (uint32_t)(i / width), x = (uint32_t)(i % width);
        const uint32_t y = ((row >> 3) * stride + offset) * 8u + (row & 7u);
        dst[i] = y < height ? src[(size_t)y * width + x] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

__global__ void __launch_bounds__(256) k_unpack_bands(const float4* __restrict__ gathered, uint32_t width, uint32_t height, uint32_t stride,
                                                      size_t rankStridePixels, float4* __restrict__ out) {
    const size_t n = (size_t)width * height;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / width), x = (uint32_t)(i % width);
        const uint32_t band = y >> 3, r = band % stride, j = band / stride;
        out[i] = gathered[(size_t)r * rankStridePixels + (size_t)(j * 8u + (y & 7u)) * width + x];
    }
}

// table[b * nranks + r]: buffer b of rank r as mapped into this process; b = 0 accumulation (OutputTexture), 1 jittered
// accumulation, 2 job-wide accumulation, 3 job-wide jittered accumulation
__global__ void __launch_bounds__(256) k_sum_peers(float4* const* __restrict__ table, uint32_t nranks, size_t first, size_t last) {
    for (size_t i = first + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < last; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (uint32_t b = 0; b < 2; b++) {
            float4 a = table[b * nranks][i];
            for (uint32_t r = 1; r < nranks; r++) { // the order of k_sum_ranks: ((r0 + r1) + r2) + ...
                const float4 v = table[b * nranks + r][i];
                a.x = a.x + v.x; a.y = a.y + v.y; a.z = a.z + v.z; a.w = a.w + v.w;
            }
            for (uint32_t r = 0; r < nranks; r++) table[(2 + b) * nranks + r][i] = a;
        }
    }
}

__global__ void __launch_bounds__(256) k_scatter_bands_peers(float4* const* __restrict__ table, uint32_t nranks, uint32_t rank, uint32_t width, uint32_t height,
                                                             size_t chunkPixels) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < chunkPixels; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t row = (uint32_t)(i / width), x = (uint32_t)(i % width);
        const uint32_t y = ((row >> 3) * nranks + rank) * 8u + (row & 7u);
        if (y >= height) continue;
        const size_t p = (size_t)y * width + x;
#pragma unroll
        for (uint32_t b = 0; b < 2; b++) {
            const float4 v = table[b * nranks + rank][p];
            for (uint32_t r = 0; r < nranks; r++) table[(2 + b) * nranks + r][p] = v;
        }
    }
}

inline uint32_t stream_grid(size_t items, int numSMs) {
    size_t blocks = (items + 255) / 256, cap = (size_t)(numSMs > 0 ? numSMs : 148) * 8; // persistent: 8 blocks of 256 per SM
    return (uint32_t)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

} // namespace

size_t band_chunk_pixels(uint32_t width, uint32_t height, uint32_t stride) {
    const uint32_t bands = (height + 7) / 8;
    return (size_t)((bands + stride - 1) / stride) * 8u * width;
}

cudaError_t sum_ranks(const float4* gathered, uint32_t nranks, size_t pixels, float4* out, int numSMs, cudaStream_t stream, LaunchCounter& lc) {
    k_sum_ranks<<<stream_grid(pixels, numSMs), 256, 0, stream>>>(gathered, nranks, pixels, out); lc.count++;
    return cudaGetLastError();
}
cudaError_t pack_bands(const float4* src, uint32_t width, uint32_t height, uint32_t offset, uint32_t stride, float4* dst, int numSMs,
                       cudaStream_t stream, LaunchCounter& lc) {
    const size_t chunk = band_chunk_pixels(width, height, stride);
    k_pack_bands<<<stream_grid(chunk, numSMs), 256, 0, stream>>>(src, width, height, offset, stride, chunk, dst); lc.count++;
    return cudaGetLastError();
}
cudaError_t unpack_bands(const float4* gathered, size_t rankStridePixels, uint32_t width, uint32_t height, uint32_t stride, float4* out, int numSMs,
                         cudaStream_t stream, LaunchCounter& lc) {
    k_unpack_bands<<<stream_grid((size_t)width * height, numSMs), 256, 0, stream>>>(gathered, width, height, stride, rankStridePixels, out); lc.count++;
    return cudaGetLastError();
}

cudaError_t sum_peers(float4* const* table, uint32_t nranks, uint32_t rank, size_t pixels, int numSMs, cudaStream_t stream, LaunchCounter& lc) {
    const size_t slice = (pixels + nranks - 1) / nranks, first = slice * rank < pixels ? slice * rank : pixels, last = first + slice < pixels ? first + slice : pixels;
    if (last > first) { k_sum_peers<<<stream_grid(last - first, numSMs), 256, 0, stream>>>(table, nranks, first, last); lc.count++; }
    return cudaGetLastError();
}
cudaError_t scatter_bands_peers(float4* const* table, uint32_t nranks, uint32_t rank, uint32_t width, uint32_t height, int numSMs, cudaStream_t stream,
                                LaunchCounter& lc) {
    const size_t chunk = band_chunk_pixels(width, height, nranks);
    k_scatter_bands_peers<<<stream_grid(chunk, numSMs), 256, 0, stream>>>(table, nranks, rank, width, height, chunk); lc.count++;
    return cudaGetLastError();
}

} // namespace tbd
