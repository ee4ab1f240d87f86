// scan.cuh — single-pass device-wide exclusive prefix sum (uint32) with decoupled look-back, usable INSIDE other kernels.
//
// One launch instead of reduce / scan-the-sums / scan-down: a CTA takes a tile by ticket (so a tile only ever waits for tiles whose
// CTAs are already running), scans it in registers, publishes its aggregate in a 64-bit status word (flag + value in one store)
// and derives its exclusive prefix by looking back over its predecessors' status words. The element values come from a functor,
// so a caller's kernel can compute them on the fly (the kd-tree build scans split flags it never stores).
//
// Control block (`ctl`, uint64 words, zeroed by the caller before the launch): [0] ticket counter, [1..] one status word per tile.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pccb200 {

constexpr int      kScanThreads = 256;
constexpr int      kScanItems   = 8;  // per thread, consecutive
constexpr int      kScanTile    = kScanThreads * kScanItems;
constexpr uint64_t kScanAggregate = 1ull << 62, kScanPrefix = 2ull << 62, kScanFlagMask = 3ull << 62;

static inline size_t scanTiles( size_t n ) { return ( n + kScanTile - 1 ) / kScanTile; }
static inline size_t scanCtlWords( size_t n ) { return scanTiles( n ) + 2; }  // uint64 words

__device__ __forceinline__ uint32_t scanWarpInclusive( uint32_t v, int lane ) {
#pragma unroll
  for ( int o = 1; o < 32; o <<= 1 ) {
    const uint32_t t = __shfl_up_sync( 0xffffffffu, v, o );
    if ( lane >= o ) v += t;
  }
  return v;
}

// Scans one tile. Must be called by all kScanThreads threads of the CTA exactly once per launch (one tile per CTA).
// load(i) -> value of element i (i < n). out[i] = sum of elements before i; out[n] = grand total.
// Returns the tile index this CTA processed (uniform).
template <class Load>
__device__ __forceinline__ unsigned scanLookbackTile( Load load, uint32_t* __restrict__ out, size_t n, unsigned long long* __restrict__ ctl ) {
  __shared__ unsigned sTile;
  __shared__ uint32_t sWarp[kScanThreads / 32];
  __shared__ uint32_t sExclusive;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if ( threadIdx.x == 0 ) sTile = unsigned( atomicAdd( &ctl[0], 1ull ) );
  __syncthreads();
  const unsigned tile = sTile;
  const size_t   base = size_t( tile ) * kScanTile + size_t( threadIdx.x ) * kScanItems;
  uint32_t       v[kScanItems];
  uint32_t       sum = 0;
#pragma unroll
  for ( int i = 0; i < kScanItems; ++i ) {
    v[i] = ( base + i < n ) ? load( base + i ) : 0u;
    sum += v[i];
  }
  const uint32_t inc = scanWarpInclusive( sum, lane );
  if ( lane == 31 ) sWarp[w] = inc;
  __syncthreads();
  if ( w == 0 ) {
    const uint32_t s  = lane < kScanThreads / 32 ? sWarp[lane] : 0;
    const uint32_t in = scanWarpInclusive( s, lane );
    if ( lane < kScanThreads / 32 ) sWarp[lane] = in - s;
    const uint32_t aggregate = __shfl_sync( 0xffffffffu, in, kScanThreads / 32 - 1 );
    volatile unsigned long long* status = ctl + 1;
    uint32_t exclusive = 0;
    if ( tile == 0 ) {
      if ( lane == 0 ) status[0] = kScanPrefix | aggregate;
    } else {
      if ( lane == 0 ) status[tile] = kScanAggregate | aggregate;
      long long look = (long long)tile - 1;
      unsigned  spins = 0;
      for ( ;; ) {
        const long long    idx = look - lane;
        unsigned long long st  = idx >= 0 ? status[idx] : ( kScanPrefix | 0ull );
        while ( __any_sync( 0xffffffffu, ( st & kScanFlagMask ) == 0 ) ) {
          if ( ( st & kScanFlagMask ) == 0 ) st = status[idx];
          if ( ++spins > ( 1u << 26 ) ) __trap();  // a predecessor never published (corrupt control block): fail, do not hang
        }
        const unsigned isPrefix = __ballot_sync( 0xffffffffu, ( st & kScanFlagMask ) == kScanPrefix );
        const int      firstP   = isPrefix ? __ffs( isPrefix ) - 1 : 32;
        exclusive += __reduce_add_sync( 0xffffffffu, lane <= firstP ? uint32_t( st ) : 0u );
        if ( isPrefix ) break;
        look -= 32;
      }
      if ( lane == 0 ) status[tile] = kScanPrefix | uint64_t( exclusive + aggregate );
    }
    if ( lane == 0 ) sExclusive = exclusive;
  }
  __syncthreads();
  uint32_t ex = sExclusive + sWarp[w] + ( inc - sum );
#pragma unroll
  for ( int i = 0; i < kScanItems; ++i ) {
    if ( base + i < n ) out[base + i] = ex;
    ex += v[i];
  }
  if ( n > 0 && base <= n - 1 && n - 1 < base + kScanItems ) out[n] = ex;  // grand total, by the owner of the last element
  return tile;
}

}  // namespace pccb200
