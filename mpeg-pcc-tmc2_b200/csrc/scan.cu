// scan.cu — device-wide exclusive prefix sum (uint32), reduce-then-scan, recursion on the block sums.
// Used by the kd-tree build (Hoare-partition ranks), compaction of patch points and point emission order.
#include "common.cuh"

namespace pccb200 {

static constexpr int kScanThreads = 256;
static constexpr int kScanItems   = 8;  // per thread
static constexpr int kScanTile    = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t warpInclusive( uint32_t v, int lane ) {
#pragma unroll
  for ( int o = 1; o < 32; o <<= 1 ) {
    uint32_t t = __shfl_up_sync( 0xffffffffu, v, o );
    if ( lane >= o ) v += t;
  }
  return v;
}

// exclusive scan across the block of one value per thread; returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t blockExclusive( uint32_t v, uint32_t* total ) {
  __shared__ uint32_t warpSums[kScanThreads / 32];
  const int           lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t            inc = warpInclusive( v, lane );
  if ( lane == 31 ) warpSums[w] = inc;
  __syncthreads();
  if ( w == 0 ) {
    uint32_t s = lane < kScanThreads / 32 ? warpSums[lane] : 0;
    uint32_t i = warpInclusive( s, lane );
    if ( lane < kScanThreads / 32 ) warpSums[lane] = i - s;
    if ( lane == kScanThreads / 32 - 1 ) *total = i;
  }
  __syncthreads();
  return inc - v + warpSums[w];
}

__global__ void __launch_bounds__( kScanThreads ) scanReduceKernel( const uint32_t* __restrict__ in, size_t n,
                                                                    uint32_t* __restrict__ blockSums ) {
  const size_t base = size_t( blockIdx.x ) * kScanTile;
  uint32_t     s    = 0;
#pragma unroll
  for ( int i = 0; i < kScanItems; ++i ) {
    size_t j = base + size_t( i ) * kScanThreads + threadIdx.x;
    if ( j < n ) s += in[j];
  }
  __shared__ uint32_t total;
  blockExclusive( s, &total );
  if ( threadIdx.x == 0 ) blockSums[blockIdx.x] = total;
}

__global__ void __launch_bounds__( kScanThreads ) scanDownKernel( const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                  size_t n, const uint32_t* __restrict__ blockOffsets ) {
  const size_t base = size_t( blockIdx.x ) * kScanTile + size_t( threadIdx.x ) * kScanItems;
  uint32_t     v[kScanItems];
  uint32_t     s = 0;
#pragma unroll
  for ( int i = 0; i < kScanItems; ++i ) {
    v[i] = ( base + i < n ) ? in[base + i] : 0;
    s += v[i];
  }
  __shared__ uint32_t total;
  uint32_t            ex = blockExclusive( s, &total ) + ( blockOffsets ? blockOffsets[blockIdx.x] : 0 );
#pragma unroll
  for ( int i = 0; i < kScanItems; ++i ) {
    if ( base + i < n ) out[base + i] = ex;
    ex += v[i];
  }
  // out[n] = grand total, written by the thread that owns the last element
  if ( n > 0 && base <= n - 1 && n - 1 < base + kScanItems ) out[n] = ex;
}

size_t scanTmpElems( size_t n ) {
  size_t total = 0;
  while ( n > 1 ) {
    size_t blocks = ( n + kScanTile - 1 ) / kScanTile;
    total += blocks + 1;  // block sums + their scan (n+1 layout, shares storage)
    if ( blocks <= 1 ) break;
    n = blocks;
  }
  return total + 8;
}

void exclusiveScanU32( const uint32_t* in, uint32_t* out, size_t n, uint32_t* tmp, cudaStream_t s ) {
  if ( n == 0 ) {
    PCC_CUDA( cudaMemsetAsync( out, 0, sizeof( uint32_t ), s ) );
    return;
  }
  const size_t blocks = ( n + kScanTile - 1 ) / kScanTile;
  if ( blocks == 1 ) {
    scanDownKernel<<<1, kScanThreads, 0, s>>>( in, out, n, nullptr );
    PCC_LAUNCH_CHECK();
    return;
  }
  uint32_t* sums = tmp;  // blocks+1 entries, scanned in place
  scanReduceKernel<<<unsigned( blocks ), kScanThreads, 0, s>>>( in, n, sums );
  PCC_LAUNCH_CHECK();
  exclusiveScanU32( sums, sums, blocks, tmp + blocks + 1, s );  // in-place is safe: each tile is read before written
  scanDownKernel<<<unsigned( blocks ), kScanThreads, 0, s>>>( in, out, n, sums );
  PCC_LAUNCH_CHECK();
}

}  // namespace pccb200
