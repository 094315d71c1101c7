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

} // namespace tbd
