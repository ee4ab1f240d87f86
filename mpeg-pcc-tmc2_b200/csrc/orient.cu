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

#include <algorithm>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

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

// after the sort: rank (+ relative-sign bits) of every edge slot
__global__ void kRanks( const uint32_t* __restrict__ sortedIds, const uint32_t* __restrict__ relBits, size_t E, uint32_t* __restrict__ rankRel ) {
  const size_t r = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( r >= E ) return;
  const uint32_t e = sortedIds[r];
  rankRel[e]       = uint32_t( r ) | relBits[e];
}

__global__ void kInvert( const uint32_t* __restrict__ vind, int n, uint32_t* __restrict__ pos ) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if ( p < n ) pos[vind[p]] = p;
}

// The walk works on TREE POSITIONS (kd-tree leaf order): neighbours in space are neighbours in memory, so the frontier the
// walk is working on stays in L2. One 128-byte row per point: 16 x (neighbour position, rank | relative sign), slots in
// ascending ORIGINAL neighbour index (the tie-break order of the keys).
__global__ void kBuildRows( const uint32_t* __restrict__ vind, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ nbrSorted,
                            const uint32_t* __restrict__ rankRel, int n, uint2* __restrict__ rows ) {
  const size_t t = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( t >= size_t( n ) * 16 ) return;
  const uint32_t p = uint32_t( t >> 4 ), slot = uint32_t( t & 15 );
  const size_t   e = size_t( vind[p] ) * 16 + slot;
  const uint32_t j = nbrSorted[e];
  rows[t]          = make_uint2( j == kInvalid ? kInvalid : pos[j], rankRel[e] );
}

struct WalkArgs {
  const uint32_t* nbr;      // n x k, original k-NN order and indices (seed sign only)
  const uint32_t* pos;      // original index -> tree position
  const uint2*    rows;     // n x 16 by tree position: (neighbour position, rank | rel)
  uint32_t*       rankEnd;  // E: end position | (flip-the-end bit << 31) of the edge holding a rank; written when the rank is queued
  const double*   normals;  // original (unoriented) normals, original order
  const short4*   pts;      // original order
  uint64_t*       L0;       // E/64 words, zeroed: leaf level of the priority queue (one bit per rank)
  uint32_t*       best;     // by position: kNone / rank of the queued incoming edge / kVisited
  uint8_t*        flip;     // by position
  int             n, k;
  int             nL1, nL2, nL3, nL4;  // 32-ary upper levels in shared memory
};

__device__ __forceinline__ int  topBit64( uint64_t w ) { return 63 - __clzll( (long long)w ); }
__device__ __forceinline__ int  topBit32( uint32_t w ) { return 31 - __clz( int( w ) ); }
__device__ __forceinline__ void prefetchL2( const void* p ) { asm volatile( "prefetch.global.L2 [%0];" ::"l"( p ) ); }
// 8-byte asynchronous copy global -> shared (LDGSTS): no destination register, so nothing can stall on it before cpAsyncWait()
__device__ __forceinline__ void cpAsync8( void* smemDst, const void* g ) {
  asm volatile( "cp.async.ca.shared.global [%0], [%1], 8;" ::"r"( uint32_t( __cvta_generic_to_shared( smemDst ) ) ), "l"( g ) : "memory" );
}
__device__ __forceinline__ void cpAsync16( void* smemDst, const void* g ) {  // (.cg: straight from L2, never a stale L1 line)
  asm volatile( "cp.async.cg.shared.global [%0], [%1], 16;" ::"r"( uint32_t( __cvta_generic_to_shared( smemDst ) ) ), "l"( g ) : "memory" );
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile( "cp.async.commit_group;" ::: "memory" ); }
__device__ __forceinline__ void cpAsyncWait() { asm volatile( "cp.async.wait_all;" ::: "memory" ); }
__device__ __forceinline__ void cpAsyncWaitAllButLatest() { asm volatile( "cp.async.wait_group 1;" ::: "memory" ); }

// Upper levels of the bit-tree: level l has one bit per word of level l-1. L1 indexes the 64-bit leaf words.
struct Levels {
  uint32_t *L1, *L2, *L3, *L4;
  __device__ __forceinline__ void set( uint32_t rank ) const {
    uint32_t w = rank >> 6;  // leaf word
    atomicOr( &L1[w >> 5], 1u << ( w & 31 ) );
    w >>= 5;
    atomicOr( &L2[w >> 5], 1u << ( w & 31 ) );
    w >>= 5;
    atomicOr( &L3[w >> 5], 1u << ( w & 31 ) );
    w >>= 5;
    atomicOr( &L4[w >> 5], 1u << ( w & 31 ) );
  }
  // leaf word `w` has just become empty: clear its bit and cascade while words drain
  __device__ __forceinline__ void leafEmptied( uint32_t w ) const {
    uint32_t bit = 1u << ( w & 31 );
    uint32_t o   = atomicAnd( &L1[w >> 5], ~bit );
    if ( o & ~bit ) return;
    w >>= 5, bit = 1u << ( w & 31 );
    o = atomicAnd( &L2[w >> 5], ~bit );
    if ( o & ~bit ) return;
    w >>= 5, bit = 1u << ( w & 31 );
    o = atomicAnd( &L3[w >> 5], ~bit );
    if ( o & ~bit ) return;
    w >>= 5;
    atomicAnd( &L4[w >> 5], ~( 1u << ( w & 31 ) ) );
  }
  // index of the highest non-empty leaf word, or -1
  __device__ __forceinline__ int topLeaf( int nL4 ) const {
    int w4 = nL4 - 1;
    while ( w4 >= 0 && L4[w4] == 0 ) --w4;
    if ( w4 < 0 ) return -1;
    const uint32_t i3 = uint32_t( w4 ) * 32 + topBit32( L4[w4] );
    const uint32_t i2 = i3 * 32 + topBit32( L3[i3] );
    const uint32_t i1 = i2 * 32 + topBit32( L2[i2] );
    return int( i1 * 32 + topBit32( L1[i1] ) );
  }
};

// One warp per frame. Priority queue = a 32-entry HOT SET held in registers (one entry per lane: the most recently queued
// edges, which the greedy walk pops next most of the time) + the global bit-tree (leaf words in global memory, 32-ary upper
// levels in shared memory) for everything that spills. The largest queued rank is max(hot-set maximum, cached top of the
// bit-tree); the cached top only has to be re-read (two dependent loads, done by lane 16 while lanes 0..15 expand) after it was
// popped or superseded. Clears of the bit-tree are fire-and-forget: the touched leaf word is remembered and re-read before the
// next descent, which keeps the upper levels exact without waiting for an atomic's return value.
// Dependent chain per visited point in the common case: neighbour row -> best[] of the 16 neighbours.
// One CTA (one warp) per frame of the batch: every frame of a GOF walks inside the SAME launch, so the walks hold one hardware
// queue instead of one each (a long-running kernel blocks whatever is queued behind it on its connection).
__device__ __forceinline__ unsigned long long globalTimerNs() {
  unsigned long long t;
  asm volatile( "mov.u64 %0, %%globaltimer;" : "=l"( t ) );
  return t;
}

__global__ void __launch_bounds__( 32, 1 ) kWalk( const WalkArgs* __restrict__ batch, unsigned long long* __restrict__ stamps, int earlyPrefetch ) {
  const WalkArgs a = batch[blockIdx.x];
  if ( threadIdx.x == 0 ) stamps[2 * blockIdx.x] = globalTimerNs();
  extern __shared__ __align__( 16 ) uint32_t smemAll[];
  uint2* const    rowBuf = reinterpret_cast<uint2*>( smemAll );  // 16 x 8 bytes: landing zone of the next point's neighbour row
  uint64_t* const pendBuf = reinterpret_cast<uint64_t*>( smemAll + 32 );  // 32 x 16 bytes: re-read leaf words (one pair per lane)
  uint32_t* const stage  = smemAll + 160;                        // 2 x 32 words: hand-over of new entries to free hot slots
  uint32_t* const smem32 = smemAll + 224;                        // upper levels of the bit-tree
  Levels          lv{ smem32, smem32 + a.nL1, smem32 + a.nL1 + a.nL2, smem32 + a.nL1 + a.nL2 + a.nL3 };
  const int                  lane = threadIdx.x;
  const unsigned             FULL = 0xffffffffu;
  for ( int i = lane; i < a.nL1 + a.nL2 + a.nL3 + a.nL4; i += 32 ) smem32[i] = 0;
  __syncwarp();

  auto finalNormal = [&]( uint32_t orig, double out[3] ) {
    const double s = a.flip[a.pos[orig]] ? -1.0 : 1.0;
    out[0] = s * a.normals[3 * size_t( orig )], out[1] = s * a.normals[3 * size_t( orig ) + 1], out[2] = s * a.normals[3 * size_t( orig ) + 2];
  };

  uint32_t hotRank = 0, hotInfo = 0;  // per lane: rank + 1 (0 = free slot), end position | flip << 31
  uint32_t pendWord = kNone;          // per lane: leaf word this lane cleared a bit in during the previous step, not re-checked yet
  uint32_t gTop = 0;                  // warp-uniform: cached maximum of the bit-tree (rank + 1, 0 = tree empty)
  uint32_t gTopInfo = 0;              // LANE 16 only: end | flip of that entry (loaded behind the scenes, read when the tree top is popped)
  bool     gTopValid = true;          // warp-uniform
#ifdef PCC_WALK_STATS
  unsigned stNew = 0, stHot = 0, stTree = 0, stRefresh = 0, stStale = 0, stSpill = 0, stFlush = 0, stTreeClr = 0;
#endif

  // lane-parallel removal of ranks from the bit-tree (active lanes pass doIt = true; clearing a rank that is not in the tree is
  // a no-op). Fire-and-forget: the touched leaf word is re-read during the NEXT step by an asynchronous copy, off the critical
  // path; if it was drained it is unhooked from the upper levels then (nothing is inserted in between).
  auto treeClear = [&]( bool doIt, uint32_t rank ) {
    if ( doIt ) {
      pendWord = rank >> 6;
      atomicAnd( (unsigned long long*)&a.L0[pendWord], ~( 1ull << ( rank & 63 ) ) );
    }
    if ( __ballot_sync( FULL, doIt && rank + 1 == gTop ) ) gTopValid = false;
  };
  // lane-parallel insertion into the bit-tree; keeps the cached top current
  auto treeInsert = [&]( bool doIt, uint32_t rank, uint32_t info ) {
    if ( doIt ) {
      a.rankEnd[rank] = info;
      atomicOr( (unsigned long long*)&a.L0[rank >> 6], 1ull << ( rank & 63 ) );
      lv.set( rank );
    }
    const uint32_t mx = __reduce_max_sync( FULL, doIt ? rank + 1 : 0u );
    if ( mx > gTop && gTopValid ) {  // (an invalid cache is re-read anyway)
      const int src = __ffs( __ballot_sync( FULL, doIt && rank + 1 == mx ) ) - 1;
      gTop          = mx;
      gTopInfo      = __shfl_sync( FULL, info, src );
    }
  };

  // re-reads the largest rank held by the bit-tree (lane 16 descends the shared-memory levels, then reads the leaf word)
  auto refreshTop = [&]() {
#ifdef PCC_WALK_STATS
    ++stRefresh;
#endif
    uint32_t found = 0;
    for ( ;; ) {
      int      leaf = -1;
      uint64_t w0   = 1;
      if ( lane == 16 ) {
        leaf = lv.topLeaf( a.nL4 );
        if ( leaf >= 0 ) w0 = __ldcg( (const unsigned long long*)&a.L0[leaf] );
      }
      const bool stale = __shfl_sync( FULL, int( leaf >= 0 && w0 == 0 ), 16 ) != 0;  // an upper bit still pointing at a drained word
      if ( stale ) {  // unhook it and look again
        if ( lane == 16 ) lv.leafEmptied( uint32_t( leaf ) );
        __syncwarp();
#ifdef PCC_WALK_STATS
        ++stStale;
#endif
        continue;
      }
      if ( lane == 16 && leaf >= 0 ) {
        found    = uint32_t( leaf ) * 64 + topBit64( w0 ) + 1;
        gTopInfo = __ldcg( &a.rankEnd[found - 1] );
      }
      break;
    }
    gTop      = __shfl_sync( FULL, found, 16 );
    gTopValid = true;
  };

  // grows one tree from position `cur` (already marked visited, flip decided) until the queue is empty.
  // Priority queue = a 32-entry HOT SET in registers (one entry per lane: the most recently queued edges, which the greedy walk
  // pops next most of the time) + the bit-tree for everything that spills. A hot entry superseded by a better edge to the same
  // point is dropped lazily: every lane re-checks best[] of its entry's end at the top of each step (an entry superseded in the
  // current step is smaller than its successor, so it cannot win before that check); superseded ranks are always cleared in the
  // bit-tree, where they are if the hot set does not hold them.
  // Software-pipelined: the point visited next needs only best[] of the 16 neighbours, the hot-set maximum and the cached tree
  // top; its neighbour row is requested as soon as it is known and the queue upkeep runs while that row is on its way.
  auto grow = [&]( uint32_t cur, bool curFlip ) {
    // (the row travels global -> shared by an asynchronous copy: a register load would be waited for at the first copy of its
    // destination register, which the compiler places right behind the load)
    if ( lane < 16 ) cpAsync8( rowBuf + lane, a.rows + size_t( cur ) * 16 + lane );
    for ( ;; ) {
      cpAsyncWait();
      __syncwarp();
      const uint2 slot = lane < 16 ? rowBuf[lane] : make_uint2( kInvalid, 0 );
      __syncwarp();
      // the rows of ALL neighbours are pulled into L2 now, one round trip before the walk knows (from best[]) which of them it may
      // visit next: when the next point is one of them (about half of the steps) its row fetch then hits L2 instead of DRAM
      if ( earlyPrefetch && lane < 16 && slot.x != kInvalid ) prefetchL2( a.rows + size_t( slot.x ) * 16 );
      if ( pendWord != kNone ) cpAsync16( pendBuf + 2 * lane, a.L0 + ( pendWord & ~1u ) );
      cpAsyncCommit();
      // ---- A: state of the neighbours and of the hot entries' ends (best[] is private to this warp: L1-cached loads)
      uint32_t old = kVisited, hv = 0;
      if ( lane < 16 && slot.x != kInvalid ) old = a.best[slot.x];
      if ( hotRank != 0 ) hv = a.best[hotInfo & 0x7fffffffu];
      const uint32_t r        = slot.y & kRankMask;
      const bool     improved = old != kVisited && ( old == kNone || r > old );
      const bool     flipEnd  = curFlip ? ( slot.y & kRelPos ) != 0 : ( slot.y & kRelNeg ) != 0;  // n_cur(final) . n_j(original) < 0
      const uint32_t info     = slot.x | ( flipEnd ? 0x80000000u : 0u );
      const uint32_t myCand   = improved ? r + 1 : 0;
      if ( hv != hotRank - 1 ) hotRank = 0;  // superseded or visited through another edge (a free slot stays free: hv = 0 != -1)
      const uint32_t A      = __reduce_max_sync( FULL, myCand );
      const uint32_t hotMax = __reduce_max_sync( FULL, hotRank );
      if ( A == 0 && hotMax == 0 && gTop == 0 ) return;  // nothing queued anywhere: this tree is complete
      const bool newWins = A > hotMax && A > gTop;
      const bool hotWins = !newWins && hotMax > gTop;
      // ---- B: the point visited next
      int      src = 16;
      uint32_t nextInfo;
      if ( newWins ) {
        src      = __ffs( __ballot_sync( FULL, myCand == A ) ) - 1;
        nextInfo = __shfl_sync( FULL, info, src );
#ifdef PCC_WALK_STATS
        ++stNew;
#endif
      } else if ( hotWins ) {
        src      = __ffs( __ballot_sync( FULL, hotRank == hotMax ) ) - 1;
        nextInfo = __shfl_sync( FULL, hotInfo, src );
#ifdef PCC_WALK_STATS
        ++stHot;
#endif
      } else {
        nextInfo = __shfl_sync( FULL, gTopInfo, 16 );
#ifdef PCC_WALK_STATS
        ++stTree;
#endif
      }
      const uint32_t next     = nextInfo & 0x7fffffffu;
      const bool     nextFlip = ( nextInfo >> 31 ) != 0;
      // ---- C: frontier keys of the improved neighbours, the visited mark (one writer per address); then the next row is requested
      const bool toQueue = improved && !( newWins && lane == src );
      if ( toQueue ) {
        a.best[slot.x] = r;
        if ( !earlyPrefetch ) prefetchL2( a.rows + size_t( slot.x ) * 16 );  // the row the visit of slot.x will read
      }
      if ( newWins ? lane == src : lane == 0 ) {
        a.flip[next] = nextFlip ? 1 : 0;
        a.best[next] = kVisited;
      }
      if ( lane < 16 ) cpAsync8( rowBuf + lane, a.rows + size_t( next ) * 16 + lane );
      cpAsyncCommit();
      // ---- D: queue upkeep. Leaf words drained in the previous step are unhooked before anything is inserted
      cpAsyncWaitAllButLatest();
      if ( pendWord != kNone ) {
        if ( pendBuf[2 * lane + ( pendWord & 1u )] == 0 ) lv.leafEmptied( pendWord );
        pendWord = kNone;
      }
      // the popped entry leaves the queue; after a pop from the bit-tree its new top is usually in the same leaf word
      if ( hotWins ) {
        if ( lane == src ) hotRank = 0;
      } else if ( !newWins ) {
        uint32_t found   = 0;
        bool     descend = false;
        if ( lane == 16 ) {
          const uint32_t rank = gTop - 1;
          const uint64_t bit  = 1ull << ( rank & 63 );
          const uint64_t left = atomicAnd( (unsigned long long*)&a.L0[rank >> 6], ~bit ) & ~bit;
          if ( left != 0 ) {
            found    = ( rank & ~63u ) + topBit64( left ) + 1;
            gTopInfo = __ldcg( &a.rankEnd[found - 1] );
          } else {
            lv.leafEmptied( rank >> 6 );
            descend = true;
          }
        }
        gTop = __shfl_sync( FULL, found, 16 );
        if ( __shfl_sync( FULL, int( descend ), 16 ) ) {
          __syncwarp();
          refreshTop();
        }
      }
      __syncwarp();
      // new entries go to free hot slots (the i-th free slot takes the i-th new entry), the overflow to the bit-tree
      {
        const unsigned newMask = __ballot_sync( FULL, toQueue );
        if ( newMask ) {
          const unsigned freeMask = __ballot_sync( FULL, hotRank == 0 );
          const int      nFree = __popc( freeMask ), nNew = __popc( newMask );
          const unsigned below     = ( 1u << lane ) - 1u;
          const int      myNewIdx  = __popc( newMask & below );
          const int      myFreeIdx = __popc( freeMask & below );
          if ( toQueue ) stage[myNewIdx] = myCand, stage[32 + myNewIdx] = info;
          __syncwarp();
          if ( hotRank == 0 && myFreeIdx < nNew ) hotRank = stage[myFreeIdx], hotInfo = stage[32 + myFreeIdx];
          if ( nNew > nFree ) {
            treeInsert( toQueue && myNewIdx >= nFree, r, info );
#ifdef PCC_WALK_STATS
            stSpill += nNew - nFree;
#endif
          }
        }
      }
      // keep room in the hot set: when it is nearly full, the lower-ranked half moves to the bit-tree
      {
        const unsigned used = __ballot_sync( FULL, hotRank != 0 );
        if ( __popc( used ) > 24 ) {
          const uint32_t mean = __reduce_add_sync( FULL, hotRank >> 5 ) / __popc( used ) << 5;
          const bool     out  = hotRank != 0 && hotRank <= mean;
          treeInsert( out, hotRank - 1, hotInfo );
#ifdef PCC_WALK_STATS
          stFlush += __popc( __ballot_sync( FULL, out ) );
#endif
          if ( out ) hotRank = 0;
        }
      }
      // superseded ranks leave the bit-tree (after the inserts: an entry flushed above may be one of them)
      {
        const bool sup = improved && old != kNone;
        if ( __ballot_sync( FULL, sup ) ) {
          __syncwarp();
          treeClear( sup, old );
#ifdef PCC_WALK_STATS
          stTreeClr += __popc( __ballot_sync( FULL, sup ) );
#endif
        }
      }
      // (__syncwarp orders the lanes' memory accesses; a __threadfence_block here would also wait for the row requested in C)
      __syncwarp();
      if ( !gTopValid ) refreshTop();
      curFlip = nextFlip;
    }
  };

  // seeds in ascending ORIGINAL index: scan 32 candidates per load, re-check each before use (a tree grown from an earlier
  // seed of the chunk may have swallowed a later one)
  for ( uint32_t base = 0; base < uint32_t( a.n ); base += 32 ) {
    const bool     inRange = base + lane < uint32_t( a.n );
    const uint32_t myPos   = inRange ? a.pos[base + lane] : 0;
    unsigned       pending = __ballot_sync( 0xffffffffu, inRange && __ldcg( &a.best[myPos] ) != kVisited );
    while ( pending ) {
      const int      sl   = __ffs( pending ) - 1;
      const uint32_t seed = base + sl;  // original index
      pending &= pending - 1;
      const uint32_t seedPos = __shfl_sync( 0xffffffffu, myPos, sl );
      if ( __ldcg( &a.best[seedPos] ) == kVisited ) continue;  // warp-uniform
      // ---- new tree: orient the seed from its already-visited neighbours (reference order of the k-NN list)
      int seedFlip = 0;
      if ( lane == 0 ) {
        double acc[3] = {0.0, 0.0, 0.0};
        int    cnt    = 0;
        for ( int t = 0; t < a.k; ++t ) {
          const uint32_t j = a.nbr[size_t( seed ) * a.k + t];
          if ( j == kInvalid ) break;
          if ( j != seed && __ldcg( &a.best[a.pos[j]] ) == kVisited ) {
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
        seedFlip         = ns[0] * acc[0] + ns[1] * acc[1] + ns[2] * acc[2] < 0.0 ? 1 : 0;
        a.flip[seedPos]  = uint8_t( seedFlip );
        a.best[seedPos]  = kVisited;  // the queue is empty between trees: the seed has no entry to remove
      }
      seedFlip = __shfl_sync( 0xffffffffu, seedFlip, 0 );
      __threadfence_block();
      __syncwarp();
      grow( seedPos, seedFlip != 0 );
    }
  }
  if ( lane == 0 ) stamps[2 * blockIdx.x + 1] = globalTimerNs();
#ifdef PCC_WALK_STATS
  if ( lane == 0 ) printf( "walk n=%d popNew=%u popHot=%u popTree=%u refresh=%u stale=%u spill=%u flush=%u treeClear=%u\n", a.n, stNew, stHot, stTree, stRefresh, stStale, stSpill, stFlush, stTreeClr );
#endif
}

__global__ void kApplyFlip( double* __restrict__ normals, const uint8_t* __restrict__ flip, const uint32_t* __restrict__ pos, const short4* __restrict__ pts, int n,
                            unsigned int* __restrict__ negCount ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool      neg = false;
  if ( i < n ) {
    double x = normals[3 * size_t( i )], y = normals[3 * size_t( i ) + 1], z = normals[3 * size_t( i ) + 2];
    if ( flip[pos[i]] ) {
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

static_assert( sizeof( WalkArgs ) <= sizeof( OrientScratch::walkArgs ), "OrientScratch::walkArgs too small" );

void orientPrepare( OrientScratch& sc, OrientTemp& tmp, const short4* pts, const uint32_t* nbr, int k, const uint32_t* vind, size_t n,
                    double* normals, cudaStream_t s ) {
  sc.walkSmem = 0;
  if ( n == 0 ) return;
  if ( k > 16 ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
  const size_t E = n * 16;
  if ( E > size_t( kRankMask ) ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
  tmp.nbrSorted.reserve( E ), tmp.relBits.reserve( E ), tmp.keysA.reserve( E ), tmp.keysB.reserve( E );
  tmp.idsA.reserve( E ), tmp.idsB.reserve( E ), tmp.rankRel.reserve( E ), sc.rankEnd.reserve( E ), sc.rows.reserve( E ), sc.pos.reserve( n );
  kEdgeKeys<<<divUp( n, 128 ), 128, 0, s>>>( nbr, normals, int( n ), k, tmp.nbrSorted, tmp.relBits, tmp.keysA, tmp.idsA );
  PCC_LAUNCH_CHECK();
  size_t tmpBytes = 0;
  PCC_CUDA( cub::DeviceRadixSort::SortPairs( nullptr, tmpBytes, tmp.keysA.p, tmp.keysB.p, tmp.idsA.p, tmp.idsB.p, E, 0, 62, s ) );
  tmp.cubTmp.reserve( tmpBytes + 16 );
  PCC_CUDA( cub::DeviceRadixSort::SortPairs( tmp.cubTmp.p, tmpBytes, tmp.keysA.p, tmp.keysB.p, tmp.idsA.p, tmp.idsB.p, E, 0, 62, s ) );
  kRanks<<<divUp( E, 256 ), 256, 0, s>>>( tmp.idsB, tmp.relBits, E, tmp.rankRel );
  kInvert<<<divUp( n, 256 ), 256, 0, s>>>( vind, int( n ), sc.pos );
  kBuildRows<<<divUp( E, 256 ), 256, 0, s>>>( vind, sc.pos, tmp.nbrSorted, tmp.rankRel, int( n ), sc.rows );
  PCC_LAUNCH_CHECK();

  WalkArgs a;
  a.nbr = nbr, a.pos = sc.pos, a.rows = sc.rows, a.rankEnd = sc.rankEnd, a.normals = normals, a.pts = pts;
  a.n = int( n ), a.k = k;
  const int nL0 = int( ( E + 63 ) / 64 );
  a.nL1 = ( nL0 + 31 ) / 32, a.nL2 = ( a.nL1 + 31 ) / 32, a.nL3 = ( a.nL2 + 31 ) / 32, a.nL4 = ( a.nL3 + 31 ) / 32;
  sc.L0.reserve( nL0 + 1 ), sc.best.reserve( n ), sc.flip.reserve( n ), sc.counter.reserve( 4 );
  PCC_CUDA( cudaMemsetAsync( sc.L0, 0, size_t( nL0 ) * 8, s ) );
  PCC_CUDA( cudaMemsetAsync( sc.flip, 0, n, s ) );
  PCC_CUDA( cudaMemsetAsync( sc.counter, 0, sizeof( unsigned ), s ) );
  kFillU32<<<divUp( n, 256 ), 256, 0, s>>>( sc.best, n, kNone );
  PCC_LAUNCH_CHECK();
  a.L0 = sc.L0, a.best = sc.best, a.flip = sc.flip;
  const size_t smemBytes = size_t( a.nL1 + a.nL2 + a.nL3 + a.nL4 + 224 ) * 4;
  if ( smemBytes > 200 * 1024 ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
  memcpy( sc.walkArgs, &a, sizeof( a ) );
  sc.walkSmem = smemBytes;
}

void orientWalkBatch( OrientScratch* const* frames, int count, DevBuf<unsigned char>& devArgs, Profiler* prof, cudaStream_t s ) {
  std::vector<WalkArgs> args;
  size_t                smem = 0;
  for ( int i = 0; i < count; ++i )
    if ( frames[i] && frames[i]->walkSmem ) {
      WalkArgs a;
      memcpy( &a, frames[i]->walkArgs, sizeof( a ) );
      args.push_back( a );
      smem = std::max( smem, frames[i]->walkSmem );
    }
  if ( args.empty() ) return;
  {  // the limit is per function AND per device: raise it once per device to the maximum any frame may need
    static std::mutex m;
    static bool       raised[64] = { false };
    int               dev        = 0;
    PCC_CUDA( cudaGetDevice( &dev ) );
    std::lock_guard<std::mutex> lk( m );
    if ( dev < 0 || dev >= 64 || !raised[dev] ) {
      PCC_CUDA( cudaFuncSetAttribute( kWalk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 ) );
      if ( dev >= 0 && dev < 64 ) raised[dev] = true;
    }
  }
  // device block: the frames' arguments, then one (start, end) %globaltimer pair per frame
  const size_t argBytes = ( args.size() * sizeof( WalkArgs ) + 15 ) & ~size_t( 15 );
  devArgs.reserve( argBytes + args.size() * 16 );
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>( devArgs.p + argBytes );
  PCC_CUDA( cudaMemcpyAsync( devArgs.p, args.data(), args.size() * sizeof( WalkArgs ), cudaMemcpyHostToDevice, s ) );
  // (A/B switch: prefetch the neighbours' rows to L2 before best[] is known instead of after, for the improved ones only. Measured on
  //  the B200, profiles/r02f_*: 1231 vs 1210 ms per 0.83 Mpts walk alone, 52.8 vs 51.4 Mpts/s with 8 GOFs in flight - within the
  //  run-to-run noise, so the variant with less traffic stays the default)
  static const int earlyPrefetch = [] {
    const char* e = getenv( "PCCB200_WALK_EARLY_PREFETCH" );
    return e && e[0] == '1' ? 1 : 0;
  }();
  kWalk<<<unsigned( args.size() ), 32, smem, s>>>( reinterpret_cast<const WalkArgs*>( devArgs.p ), stamps, earlyPrefetch );
  PCC_LAUNCH_CHECK();
  // The walks run for a second or more on one warp each. NOTHING is enqueued behind them - an event record or a dependent
  // kernel would sit at the head of the stream's hardware queue until they finish and hold up every other stream that shares
  // the queue (there are 32 queues and hundreds of frame streams when several GOFs are in flight). The host polls instead.
  for ( ;; ) {
    const cudaError_t q = cudaStreamQuery( s );
    if ( q == cudaSuccess ) break;
    if ( q != cudaErrorNotReady ) PCC_CUDA( q );
    std::this_thread::sleep_for( std::chrono::microseconds( 200 ) );
  }
  if ( prof && prof->enabled ) {  // the span of the launch, timed on the device by the kernel itself
    std::vector<unsigned long long> h( 2 * args.size() );
    PCC_CUDA( cudaMemcpyAsync( h.data(), stamps, h.size() * 8, cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    unsigned long long t0 = ~0ull, t1 = 0;
    for ( size_t i = 0; i < args.size(); ++i ) t0 = std::min( t0, h[2 * i] ), t1 = std::max( t1, h[2 * i + 1] );
    prof->results.push_back( Profiler::Result{ "orient_walk", float( double( t1 - t0 ) * 1e-6 ), -1.f } );
  }
}

void orientFinish( OrientScratch& sc, const short4* pts, size_t n, double* normals, cudaStream_t s ) {
  if ( n == 0 ) return;
  kApplyFlip<<<divUp( n, 256 ), 256, 0, s>>>( normals, sc.flip, sc.pos, pts, int( n ), sc.counter );
  kNegateIfMajority<<<divUp( 3 * n, 256 ), 256, 0, s>>>( normals, int( n ), sc.counter );
  PCC_LAUNCH_CHECK();
}

void orientNormals( OrientScratch& sc, OrientTemp& tmp, DevBuf<unsigned char>& devArgs, const short4* pts, const uint32_t* nbr, int k,
                    const uint32_t* vind, size_t n, double* normals, cudaStream_t s ) {
  orientPrepare( sc, tmp, pts, nbr, k, vind, n, normals, s );
  OrientScratch* one = &sc;
  orientWalkBatch( &one, 1, devArgs, sc.prof, s );
  orientFinish( sc, pts, n, normals, s );
}

}  // namespace pccb200
