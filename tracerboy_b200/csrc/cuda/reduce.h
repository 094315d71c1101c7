// reduce.h — combining the per-rank accumulation buffers after the all-gather (reduce.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "launch.h"

namespace tbd {
// pixels one rank contributes under row-band sharding: ceil(bands / stride) bands of 8 rows (the last one zero-padded)
size_t band_chunk_pixels(uint32_t width, uint32_t height, uint32_t stride);
// out[i] = ((g[0][i] + g[1][i]) + g[2][i]) + ... : fixed rank order
cudaError_t sum_ranks(const float4* gathered, uint32_t nranks, size_t pixels, float4* out, int numSMs, cudaStream_t stream, LaunchCounter& lc);
cudaError_t pack_bands(const float4* src, uint32_t width, uint32_t height, uint32_t offset, uint32_t stride, float4* dst, int numSMs,
                       cudaStream_t stream, LaunchCounter& lc);
// gathered: rank r's chunk starts at r * rankStridePixels (several buffers may travel in one all-gather)
cudaError_t unpack_bands(const float4* gathered, size_t rankStridePixels, uint32_t width, uint32_t height, uint32_t stride, float4* out, int numSMs,
                         cudaStream_t stream, LaunchCounter& lc);
// Peer-memory transport: table[b * nranks + r] = buffer b (0 accumulation, 1 jittered, 2 job-wide accumulation, 3 job-wide
// jittered) of rank r, mapped into this process. sum_peers: this rank's 1/N slice of the fixed-order sum, stored into every
// rank's job-wide buffers. scatter_bands_peers: this rank's row bands, stored into every rank's job-wide buffers.
cudaError_t sum_peers(float4* const* table, uint32_t nranks, uint32_t rank, size_t pixels, int numSMs, cudaStream_t stream, LaunchCounter& lc);
cudaError_t scatter_bands_peers(float4* const* table, uint32_t nranks, uint32_t rank, uint32_t width, uint32_t height, int numSMs, cudaStream_t stream,
                                LaunchCounter& lc);
} // namespace tbd
