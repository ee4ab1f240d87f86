// orient.cu — spanning-tree normal orientation (PCCNormalsGenerator3::orientNormals, SPANNING_TREE:
// PccLibEncoder/source/PCCNormalsGenerator.cpp:198-242, addNeighbors :521-548, edge order
// PccLibEncoder/include/PCCNormalsGenerator.h:64-72).
//
// The reference grows a tree with a global std::priority_queue of directed k-NN edges keyed by
// (|n_i.n_j|, start, end): always the largest key leaving the visited set.  That greedy walk is inherently
// sequential (its result depends on the visiting order wherever the sign field is frustrated), so the exact
// formulation here splits it into
//   (1) data-parallel precomputation: every edge key and relative sign is independent of the walk
//       (|(-a).b| == |a.b| bit for bit), so all 16 N keys are computed and radix-sorted once; an edge's
//       RANK in that order replaces the (double,uint,uint) key;
//   (2) one warp per frame replays the walk on integer ranks with a 64-ary bit-tree priority queue
//       (3 upper levels in shared memory, leaf words in L2-resident global memory) that holds ONE entry per
//       frontier point (its best incoming edge) — identical pops to the reference's lazy-deletion heap,
//       ~5 dependent L2 round trips per visited point;
//   (3) data-parallel sign application and the global majority vote.
// Frames of a GOF run this stage concurrently on separate streams (one resident warp each).
#include <cub/device/device_radix_sort.cuh>

#include <mutex>

#include "stages.cuh"

namespace pccb200 {

namespace {

constexpr uint32_t kInvalid  = 0xFFFFFFFFu;
constexpr uint32_t kNone     = 0xFFFFFFFFu;  // best[]: not on the frontier
constexpr uint32_t kVisited  = 0xFFFFFFFEu;  // best[]: already in the tree
constexpr uint32_t kRankMask = 0x3FFFFFFFu;
constexpr uint32_t kRelNeg   = 0x40000000u;  // original dot(n_start, n_end) < 0
constexpr uint32_t kRelPos   = 0x80000000u;  // original dot(n_start, n_end) > 0

// Per point: sort its k neighbour slots by neighbour index so that slot order == (start, end) order, and
// compute each edge's key bits |n_i . n_j| and relative sign.
__global__ void __launch_bounds__( 128 )
    kEdgeKeys( const uint32_t* __restrict__ nbr, const double* __restrict__ normals, int n, int k, uint32_t* __restrict__ nbrSorted,
               uint32_t* __restrict__ relBits, uint64_t* __restrict__ keys, uint32_t* __restrict__ ids ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  uint32_t row[16];
#pragma unroll
  for ( int t = 0; t < 16; ++t ) row[t] = t < k ? nbr[size_t( i ) * k + t] : kInvalid;
  // insertion sort of 16 indices (invalid = 0xFFFFFFFF sorts last)
  for ( int a = 1; a < 16; ++a ) {
    const uint32_t v = row[a];
    int            b = a;
    while ( b > 0 && row[b - 1] > v ) {
      row[b] = row[b - 1];
      --b;
    }
    row[b] = v;
  }
  const double x = normals[3 * size_t( i )], y = normals[3 * size_t( i ) + 1], z = normals[3 * size_t( i ) + 2];
#pragma unroll
  for ( int t = 0; t < 16; ++t ) {
    const size_t   e = size_t( i ) * 16 + t;
    const uint32_t j = row[t];
    uint64_t       key = 0;
    uint32_t       rel = 0;
    if ( j != kInvalid ) {
      const double d = x * normals[3 * size_t( j )] + y * normals[3 * size_t( j ) + 1] + z * normals[3 * size_t( j ) + 2];
      key            = uint64_t( __double_as_longlong( fabs( d ) ) );
      rel            = d < 0.0 ? kRelNeg : ( d > 0.0 ? kRelPos : 0u );
    }
    nbrSorted[e] = j;
    relBits[e]   = rel;
    keys[e]      = key;
    ids[e]       = uint32_t( e );
  }
}

// after the sort: rank of every edge and the (start, end|rel) record of every rank
__global__ void kRanks( const uint32_t* __restrict__ sortedIds, const uint32_t* __restrict__ nbrSorted,
                        const uint32_t* __restrict__ relBits, size_t E, uint32_t* __restrict__ rankRel, uint2* __restrict__ byRank ) {
  const size_t r = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( r >= E ) return;
  const uint32_t e = sortedIds[r];
  rankRel[e]       = uint32_t( r ) | relBits[e];
  byRank[r]        = make_uint2( e >> 4, nbrSorted[e] | relBits[e] );  // end < 2^30 (E < 2^30)
}

struct WalkArgs {
  const uint32_t* nbr;        // n x k, original k-NN order (seed sign only)
  const uint32_t* nbrSorted;  // n x 16
  const uint32_t* rankRel;    // n x 16
  const uint2*    byRank;     // E
  const double*   normals;    // original (unoriented) normals
  const short4*   pts;
  uint64_t*       L0;    // E/64 words, zeroed
  uint32_t*       best;  // n, kNone
  uint8_t*        flip;  // n, 0
  int             n, k;
  int             nL0, nL1, nL2, nL3;
};

__device__ __forceinline__ int topBit( uint64_t w ) { return 63 - __clzll( (long long)w ); }

// 1 warp. Lane 0 drives; lanes 0..15 expand the neighbours of the newly visited point.
__global__ void __launch_bounds__( 32, 1 ) kWalk( WalkArgs a ) {
  extern __shared__ uint64_t smem[];
  uint64_t*                  L1   = smem;
  uint64_t*                  L2   = L1 + a.nL1;
  uint64_t*                  L3   = L2 + a.nL2;
  const int                  lane = threadIdx.x;
  for ( int i = lane; i < a.nL1 + a.nL2 + a.nL3; i += 32 ) smem[i] = 0;
  __syncwarp();

  auto finalNormal = [&]( uint32_t i, double out[3] ) {
    const double s = a.flip[i] ? -1.0 : 1.0;
    out[0] = s * a.normals[3 * size_t( i )], out[1] = s * a.normals[3 * size_t( i ) + 1], out[2] = s * a.normals[3 * size_t( i ) + 2];
  };

  // push / improve the frontier entries of `cur`'s unvisited neighbours (lanes 0..15)
  auto expand = [&]( uint32_t cur ) {
    uint32_t setRank = kNone, clrRank = kNone;
    if ( lane < 16 ) {
      const size_t   e = size_t( cur ) * 16 + lane;
      const uint32_t j = a.nbrSorted[e];
      if ( j != kInvalid ) {
        const uint32_t old = __ldcg( &a.best[j] );
        if ( old != kVisited ) {
          const uint32_t r = a.rankRel[e] & kRankMask;
          if ( old == kNone || r > old ) {
            a.best[j] = r;
            setRank   = r;
            clrRank   = old;
          }
        }
      }
    }
    __syncwarp();
    if ( clrRank != kNone ) {  // remove the superseded entry, emptying upper levels when a word drains
      uint64_t bit = 1ull << ( clrRank & 63 );
      uint64_t o   = atomicAnd( (unsigned long long*)&a.L0[clrRank >> 6], ~bit );
      if ( ( o & ~bit ) == 0 ) {
        uint32_t w = clrRank >> 6;
        bit        = 1ull << ( w & 63 );
        o          = atomicAnd( (unsigned long long*)&L1[w >> 6], ~bit );
        if ( ( o & ~bit ) == 0 ) {
          w >>= 6;
          bit = 1ull << ( w & 63 );
          o   = atomicAnd( (unsigned long long*)&L2[w >> 6], ~bit );
          if ( ( o & ~bit ) == 0 ) {
            w >>= 6;
            atomicAnd( (unsigned long long*)&L3[w >> 6], ~( 1ull << ( w & 63 ) ) );
          }
        }
      }
    }
    __syncwarp();
    if ( setRank != kNone ) {
      uint32_t w = setRank;
      atomicOr( (unsigned long long*)&a.L0[w >> 6], 1ull << ( w & 63 ) );
      w >>= 6;
      atomicOr( (unsigned long long*)&L1[w >> 6], 1ull << ( w & 63 ) );
      w >>= 6;
      atomicOr( (unsigned long long*)&L2[w >> 6], 1ull << ( w & 63 ) );
      w >>= 6;
      atomicOr( (unsigned long long*)&L3[w >> 6], 1ull << ( w & 63 ) );
    }
    __threadfence_block();
    __syncwarp();
  };

  // seeds in ascending index: scan 32 candidates per coalesced load, re-check each before use (a tree grown from
  // an earlier seed of the chunk may have swallowed a later one)
  for ( uint32_t base = 0; base < uint32_t( a.n ); base += 32 ) {
   unsigned pending = __ballot_sync( 0xffffffffu, base + lane < uint32_t( a.n ) && __ldcg( &a.best[base + lane] ) != kVisited );
   while ( pending ) {
    const uint32_t seed = base + ( __ffs( pending ) - 1 );
    pending &= pending - 1;
    if ( __ldcg( &a.best[seed] ) == kVisited ) continue;  // warp-uniform
    // ---- new tree: orient the seed from its already-visited neighbours (reference order of the k-NN list)
    if ( lane == 0 ) {
      double acc[3] = {0.0, 0.0, 0.0};
      int    cnt    = 0;
      for ( int t = 0; t < a.k; ++t ) {
        const uint32_t j = a.nbr[size_t( seed ) * a.k + t];
        if ( j == kInvalid ) break;
        if ( j != seed && __ldcg( &a.best[j] ) == kVisited ) {
          double nj[3];
          finalNormal( j, nj );
          acc[0] = acc[0] + nj[0], acc[1] = acc[1] + nj[1], acc[2] = acc[2] + nj[2];
          ++cnt;
        }
      }
      if ( cnt == 0 ) {
        if ( seed != 0 ) {
          finalNormal( seed - 1, acc );
        } else {
          const short4 p = a.pts[0];
          acc[0] = 0.0 - double( p.x ), acc[1] = 0.0 - double( p.y ), acc[2] = 0.0 - double( p.z );
        }
      }
      const double* ns = a.normals + 3 * size_t( seed );
      if ( ns[0] * acc[0] + ns[1] * acc[1] + ns[2] * acc[2] < 0.0 ) a.flip[seed] = 1;
      // remove the seed's own frontier entry, if any, then mark it visited
      const uint32_t old = a.best[seed];
      a.best[seed]       = kVisited;
      (void)old;  // the queue is empty between trees, so the seed has no frontier entry to remove
    }
    __threadfence_block();
    __syncwarp();
    expand( seed );
    // ---- grow: pop the largest rank until the queue is empty
    for ( ;; ) {
      uint32_t r = kNone;
      if ( lane == 0 ) {
        int w3 = a.nL3 - 1;
        while ( w3 >= 0 && L3[w3] == 0 ) --w3;
        if ( w3 >= 0 ) {
          const uint32_t i2 = uint32_t( w3 ) * 64 + topBit( L3[w3] );
          const uint32_t i1 = i2 * 64 + topBit( L2[i2] );
          const uint32_t i0 = i1 * 64 + topBit( L1[i1] );
          const uint64_t w0 = __ldcg( (const unsigned long long*)&a.L0[i0] );
          const int      b  = topBit( w0 );
          r                 = i0 * 64 + b;
          const uint64_t nw = w0 & ~( 1ull << b );
          atomicAnd( (unsigned long long*)&a.L0[i0], ~( 1ull << b ) );
          if ( nw == 0 ) {
            L1[i1] &= ~( 1ull << ( i0 & 63 ) );
            if ( L1[i1] == 0 ) {
              L2[i2] &= ~( 1ull << ( i1 & 63 ) );
              if ( L2[i2] == 0 ) L3[w3] &= ~( 1ull << ( i2 & 63 ) );
            }
          }
        }
      }
      r = __shfl_sync( 0xffffffffu, r, 0 );
      if ( r == kNone ) break;
      const uint2    se    = a.byRank[r];
      const uint32_t start = se.x, end = se.y & kRankMask, relHere = se.y & ( kRelNeg | kRelPos );
      if ( lane == 0 ) {
        const bool startFlipped = a.flip[start] != 0;
        // normals_[start] (final) . normals_[end] (original) < 0 ?
        const bool flipEnd = startFlipped ? ( relHere & kRelPos ) != 0 : ( relHere & kRelNeg ) != 0;
        a.flip[end]        = flipEnd ? 1 : 0;
        a.best[end]        = kVisited;
      }
      __threadfence_block();
      __syncwarp();
      expand( end );
    }
   }
  }
}

__global__ void kApplyFlip( double* __restrict__ normals, const uint8_t* __restrict__ flip, const short4* __restrict__ pts, int n,
                            unsigned int* __restrict__ negCount ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool      neg = false;
  if ( i < n ) {
    double x = normals[3 * size_t( i )], y = normals[3 * size_t( i ) + 1], z = normals[3 * size_t( i ) + 2];
    if ( flip[i] ) {
      x = -x, y = -y, z = -z;
      normals[3 * size_t( i )] = x, normals[3 * size_t( i ) + 1] = y, normals[3 * size_t( i ) + 2] = z;
    }
    const short4 p = pts[i];
    neg            = x * ( 0.0 - double( p.x ) ) + y * ( 0.0 - double( p.y ) ) + z * ( 0.0 - double( p.z ) ) < 0.0;
  }
  const unsigned m = __ballot_sync( 0xffffffffu, neg );
  if ( ( threadIdx.x & 31 ) == 0 && m ) atomicAdd( negCount, __popc( m ) );
}

__global__ void kNegateIfMajority( double* __restrict__ normals, int n, const unsigned int* __restrict__ negCount ) {
  if ( *negCount <= ( unsigned( n ) + 1u ) / 2u ) return;
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i < 3 * size_t( n ) ) normals[i] = -normals[i];
}

__global__ void kFillU32( uint32_t* p, size_t n, uint32_t v ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i < n ) p[i] = v;
}

}  // namespace

void orientNormals( OrientScratch& sc, const short4* pts, const uint32_t* nbr, int k, size_t n, double* normals, cudaStream_t s ) {
  if ( n == 0 ) return;
  if ( k > 16 ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
  const size_t E = n * 16;
  if ( E > size_t( kRankMask ) ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
  sc.nbrSorted.reserve( E ), sc.relBits.reserve( E ), sc.keysA.reserve( E ), sc.keysB.reserve( E );
  sc.idsA.reserve( E ), sc.idsB.reserve( E ), sc.rankRel.reserve( E ), sc.byRank.reserve( E );
  kEdgeKeys<<<divUp( n, 128 ), 128, 0, s>>>( nbr, normals, int( n ), k, sc.nbrSorted, sc.relBits, sc.keysA, sc.idsA );
  PCC_LAUNCH_CHECK();
  size_t tmpBytes = 0;
  PCC_CUDA( cub::DeviceRadixSort::SortPairs( nullptr, tmpBytes, sc.keysA.p, sc.keysB.p, sc.idsA.p, sc.idsB.p, E, 0, 62, s ) );
  sc.cubTmp.reserve( tmpBytes + 16 );
  PCC_CUDA( cub::DeviceRadixSort::SortPairs( sc.cubTmp.p, tmpBytes, sc.keysA.p, sc.keysB.p, sc.idsA.p, sc.idsB.p, E, 0, 62, s ) );
  kRanks<<<divUp( E, 256 ), 256, 0, s>>>( sc.idsB, sc.nbrSorted, sc.relBits, E, sc.rankRel, sc.byRank );
  PCC_LAUNCH_CHECK();

  WalkArgs a;
  a.nbr = nbr, a.nbrSorted = sc.nbrSorted, a.rankRel = sc.rankRel, a.byRank = sc.byRank, a.normals = normals, a.pts = pts;
  a.n = int( n ), a.k = k;
  a.nL0 = int( ( E + 63 ) / 64 ), a.nL1 = ( a.nL0 + 63 ) / 64, a.nL2 = ( a.nL1 + 63 ) / 64, a.nL3 = ( a.nL2 + 63 ) / 64;
  sc.L0.reserve( a.nL0 + 1 ), sc.best.reserve( n ), sc.flip.reserve( n ), sc.counter.reserve( 4 );
  PCC_CUDA( cudaMemsetAsync( sc.L0, 0, size_t( a.nL0 ) * 8, s ) );
  PCC_CUDA( cudaMemsetAsync( sc.flip, 0, n, s ) );
  PCC_CUDA( cudaMemsetAsync( sc.counter, 0, sizeof( unsigned ), s ) );
  kFillU32<<<divUp( n, 256 ), 256, 0, s>>>( sc.best, n, kNone );
  a.L0 = sc.L0, a.best = sc.best, a.flip = sc.flip;
  const size_t smemBytes = size_t( a.nL1 + a.nL2 + a.nL3 ) * 8;
  if ( smemBytes > 200 * 1024 ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
  {  // the limit is a per-function global: raise it once to the maximum any frame may need (frames run on concurrent host threads)
    static std::once_flag once;
    cudaError_t           err = cudaSuccess;
    std::call_once( once, [&]() { err = cudaFuncSetAttribute( kWalk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 ); } );
    PCC_CUDA( err );
  }
  {
    ProfScope t( sc.prof, "orient_walk", s );
    kWalk<<<1, 32, smemBytes, s>>>( a );
    PCC_LAUNCH_CHECK();
  }
  // The walk runs for seconds on one warp. Nothing that depends on it is enqueued until it has finished: a dependent
  // launch waiting at the head of a hardware queue would stall unrelated kernels of other frames' streams that share the
  // queue (CUDA_DEVICE_MAX_CONNECTIONS queues for all streams). The per-frame host thread simply waits here.
  PCC_CUDA( cudaStreamSynchronize( s ) );
  kApplyFlip<<<divUp( n, 256 ), 256, 0, s>>>( normals, sc.flip, pts, int( n ), sc.counter );
  kNegateIfMajority<<<divUp( 3 * n, 256 ), 256, 0, s>>>( normals, int( n ), sc.counter );
  PCC_LAUNCH_CHECK();
}

}  // namespace pccb200
