// radix_sort.cuh -- hand-written LSD radix sort of (key, value) pairs, 8 bits per pass, stable.
// Used for the Morton-code sort of the LBVH builder (the reference's CPU RadixSort, cpu/aggregates.cpp:42-82,
// is the nearest counterpart: 6-bit digits, serial) and for the optional ray sort by material/direction key
// (no reference counterpart).
//
// Per pass: (1) per-tile digit histogram, (2) exclusive scan of the digit-major [256][nTiles] table,
// (3) stable scatter: each warp owns a contiguous sub-tile and ranks its keys 32 at a time with
// __match_any_sync, so equal digits keep their input order.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace bpt {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_STEPS = 16;                       // keys per lane
constexpr int RS_TILE = RS_THREADS * RS_STEPS;     // 4096 keys per block

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rs_histogram(const KeyT *__restrict__ keys, int n, int shift,
                                                           unsigned *__restrict__ hist, int nTiles) {
    __shared__ unsigned counts[256];
    counts[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int s = 0; s < RS_STEPS; ++s) {
        int i = base + s * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&counts[(unsigned)(keys[i] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nTiles + blockIdx.x] = counts[threadIdx.x];
}

// single-block exclusive scan over `len` entries
static __global__ void __launch_bounds__(1024) rs_scan(unsigned *__restrict__ data, int len) {
    __shared__ unsigned warpSums[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < len; base += 1024 * 4) {
        int i0 = base + threadIdx.x * 4;
        unsigned v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i0 + k < len) ? data[i0 + k] : 0u;
        unsigned tsum = v[0] + v[1] + v[2] + v[3];
        unsigned incl = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warpSums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned w = warpSums[lane];
            unsigned wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warpSums[lane] = wi - w;   // exclusive
        }
        __syncthreads();
        unsigned excl = carry + warpSums[warp] + (incl - tsum);
        unsigned run = excl;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < len) data[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = run;
        __syncthreads();
    }
}

template <typename KeyT, typename ValT>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter(const KeyT *__restrict__ keysIn, const ValT *__restrict__ valsIn,
                                                         KeyT *__restrict__ keysOut, ValT *__restrict__ valsOut, int n, int shift,
                                                         const unsigned *__restrict__ histScanned, int nTiles) {
    __shared__ unsigned warpCount[RS_WARPS][256];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < RS_WARPS * 256; d += RS_THREADS) (&warpCount[0][0])[d] = 0;
    __syncthreads();
    int warpBase = blockIdx.x * RS_TILE + warp * (32 * RS_STEPS);
    KeyT k[RS_STEPS];
#pragma unroll
    for (int s = 0; s < RS_STEPS; ++s) {
        int i = warpBase + s * 32 + lane;
        bool valid = i < n;
        k[s] = valid ? keysIn[i] : (KeyT)0;
        unsigned digit = (unsigned)(k[s] >> shift) & 0xffu;
        unsigned active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            unsigned peers = __match_any_sync(active, digit);
            if (lane == (__ffs(peers) - 1)) warpCount[warp][digit] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over warps, plus the global base for this tile
    {
        int d = threadIdx.x;   // RS_THREADS == 256 digits
        unsigned run = histScanned[(size_t)d * nTiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            unsigned c = warpCount[w][d];
            warpCount[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < RS_STEPS; ++s) {
        int i = warpBase + s * 32 + lane;
        bool valid = i < n;
        unsigned digit = (unsigned)(k[s] >> shift) & 0xffu;
        unsigned active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            unsigned peers = __match_any_sync(active, digit);
            int leader = __ffs(peers) - 1;
            unsigned base = 0;
            if (lane == leader) {
                base = warpCount[warp][digit];
                warpCount[warp][digit] = base + __popc(peers);
            }
            base = __shfl_sync(peers, base, leader);
            unsigned dst = base + __popc(peers & ((1u << lane) - 1));
            keysOut[dst] = k[s];
            valsOut[dst] = valsIn[i];
        }
        __syncwarp();
    }
}

// Sorts bits [beginBit, endBit) of the keys.  Returns true if the result is in (keysTmp, valsTmp), false if it is
// back in (keys, vals).  `hist` must hold 256 * ceil(n / RS_TILE) unsigneds.
template <typename KeyT, typename ValT>
inline bool RadixSortPairs(KeyT *keys, KeyT *keysTmp, ValT *vals, ValT *valsTmp, int n, int beginBit, int endBit, unsigned *hist,
                           cudaStream_t stream) {
    int nTiles = (n + RS_TILE - 1) / RS_TILE;
    bool flipped = false;
    if (n <= 0) return false;
    for (int shift = beginBit; shift < endBit; shift += 8) {
        rs_histogram<KeyT><<<nTiles, RS_THREADS, 0, stream>>>(keys, n, shift, hist, nTiles);
        rs_scan<<<1, 1024, 0, stream>>>(hist, 256 * nTiles);
        rs_scatter<KeyT, ValT><<<nTiles, RS_THREADS, 0, stream>>>(keys, vals, keysTmp, valsTmp, n, shift, hist, nTiles);
        KeyT *tk = keys; keys = keysTmp; keysTmp = tk;
        ValT *tv = vals; vals = valsTmp; valsTmp = tv;
        flipped = !flipped;
    }
    return flipped;
}

}  // namespace bpt
