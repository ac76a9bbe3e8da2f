// scan.cuh -- stable compaction of the active mask.
//
// The reference tests activeMask(j) inside the pair loop (e.g.
// src/SphereBVESolver.f90:402).  Here the mask is scanned once per evaluation:
//   scan[i]   = number of active particles in [0, i)      (length n+1)
//   active[c] = index of the c-th active particle         (length scan[n])
// active == Fortran pack([(j,j=1,n)], mask) (0-based): the order is stable, so
// the source list is bit-exact with the reference's visiting order.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace lpm {

constexpr int kScanBlock = 1024;    // elements per block (one per thread)

// inclusive scan of one int per thread across a 1024-thread block
__device__ __forceinline__ int block_inclusive_scan(int v, int* warp_sums)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    if (lane == 31) warp_sums[wid] = v;
    __syncthreads();
    if (wid == 0) {
        int w = warp_sums[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += o;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    if (wid > 0) v += warp_sums[wid - 1];
    return v;
}

__global__ void __launch_bounds__(kScanBlock) scan_count(int64_t n, const int32_t* __restrict__ mask, int32_t* __restrict__ blocksums)
{
    __shared__ int ws[32];
    int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    int v = (i < n && mask[i] != 0) ? 1 : 0;
    int s = block_inclusive_scan(v, ws);
    if (threadIdx.x == kScanBlock - 1) blocksums[blockIdx.x] = s;
}

// exclusive scan of the block sums by a single block (nblocks <= ~31k for n = 31M)
__global__ void __launch_bounds__(kScanBlock) scan_blocksums(int nblocks, int32_t* __restrict__ blocksums)
{
    __shared__ int ws[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += kScanBlock) {
        int i = base + threadIdx.x;
        int v = (i < nblocks) ? blocksums[i] : 0;
        int s = block_inclusive_scan(v, ws);
        int c = carry;
        if (i < nblocks) blocksums[i] = c + s - v;
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) carry = c + s;
        __syncthreads();
    }
    if (threadIdx.x == 0) blocksums[nblocks] = carry;    // total
}

__global__ void __launch_bounds__(kScanBlock) scan_scatter(int64_t n, const int32_t* __restrict__ mask,
                                                           const int32_t* __restrict__ blocksums, int nblocks,
                                                           int32_t* __restrict__ scan, int32_t* __restrict__ active)
{
    __shared__ int ws[32];
    int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    int v = (i < n && mask[i] != 0) ? 1 : 0;
    int s = block_inclusive_scan(v, ws);
    int ex = blocksums[blockIdx.x] + s - v;
    if (i < n) {
        scan[i] = ex;
        if (v && active) active[ex] = (int32_t)i;
    }
    if (i == n - 1) scan[n] = ex + v;
}

}  // namespace lpm
