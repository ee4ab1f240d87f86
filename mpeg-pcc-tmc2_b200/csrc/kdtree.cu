// kdtree.cu — GPU construction of nanoflann's kd-tree and exact k-NN by per-thread traversal.
//
// Build (restates nanoflann.hpp:1041-1181). Per node with more than 10 points:
//     1. per-dimension min/max of its points                (computeMinMax)
//     2. cut dimension / cut value                          (middleSplit_)
//     3. Hoare sweep #1  [<cut | >=cut], sweep #2 [==cut | >cut]   (planeSplit)    -> flags + prefix sum:
//        the sequential two-pointer sweep swaps the i-th misplaced element from the left with the i-th
//        misplaced element from the right, so ranks from one prefix sum reproduce its permutation exactly.
//     4. split position (lim1/lim2/half rule), children, tight divlow/divhigh (max of left / min of right)
// Two phases:
//   TOP: while a node holds more than kLocalMax points the tree is built level-synchronously over all points, FIVE launches
//        per level: scan of the sweep-1 flags (single-pass look-back scan, flags computed on the fly: scan.cuh) | compaction of
//        the misplaced elements | sweep-1 swap fused into the scan of the sweep-2 flags | compaction + children of the level |
//        sweep-2 swap fused with the tight split bounds, the next level's slot assignment and its min/max. The level loop
//        runs without host round trips: slot counts live on the device, the grids are sized for the worst case, and a level
//        without slots costs five empty launches; the host reads back ONCE per tree (leftover slots, node count, depth).
//   LOCAL: a subtree of at most kLocalMax points is finished by ONE CTA in shared memory (points, permutation, scans, node
//        lists all on chip), persistent CTAs taking subtrees by ticket: the bottom ~8 levels cost one launch.
// Search (nanoflann.hpp:1207-1254): explicit-stack depth-first traversal, near child first, far child tested
// against the current worst distance when it is popped — the same moment the recursion would test it.
#include <limits.h>

#include <mutex>

#include "kdtree.cuh"
#include "scan.cuh"

namespace pccb200 {

namespace {

constexpr int kLocalMax     = 2048;  // subtrees up to this size are finished inside one CTA
constexpr int kLocalThreads = 256;
constexpr int kLocalItems   = kLocalMax / kLocalThreads;  // consecutive positions per thread in the block scans
constexpr int kLocalNodes   = kLocalMax / ( kLeafMaxSize + 1 ) + 2;  // nodes with > 10 points alive at one local level
constexpr int kMaxTopLevels = 64;    // control blocks of the look-back scans are laid out for this many top levels
constexpr int kLocalRecInts = 12;    // node, lo, hi, box[6], depth, pad, pad

// device counters (ints)
enum Counter { C_NODES = 0, C_LOCAL, C_DEPTH, C_TICKET, C_ERROR, C_TOPDEPTH, C_ROOTBOX = 8 /*6*/, C_LEVEL = 16 /* kMaxTopLevels + 2 */, C_COUNT = C_LEVEL + kMaxTopLevels + 2 };

enum SlotField {
  F_NODE = 0, F_LO, F_HI, F_BOX /*6*/, F_MM = F_BOX + 6 /*6*/, F_FEAT = F_MM + 6, F_CUT, F_LIM1, F_IDX, F_CHILD0, F_CHILD1, F_COUNT
};

struct SlotView {
  int* base;
  int  stride;
  __device__ __forceinline__ int& at( int field, int s ) const { return base[size_t( field ) * stride + s]; }
};

__device__ __forceinline__ int coord( const short4& p, int d ) { return d == 0 ? p.x : ( d == 1 ? p.y : p.z ); }

// same-address extremes: read (L2) before the atomic; a stale value can only cause a redundant atomic (see patches.cu relaxMin)
__device__ __forceinline__ void relaxMin( int* addr, int v ) {
  if ( v < __ldcg( addr ) ) atomicMin( addr, v );
}
__device__ __forceinline__ void relaxMax( int* addr, int v ) {
  if ( v > __ldcg( addr ) ) atomicMax( addr, v );
}

// middleSplit_ (nanoflann.hpp:1103-1131). Spans are int16 in the reference (ElementType), compared in double.
__device__ __forceinline__ void decideCut( const int lo[3], const int hi[3], const int mn[3], const int mx[3], int& feat, int& cut ) {
  short maxSpan = short( hi[0] - lo[0] );
  for ( int d = 1; d < 3; ++d ) {
    const short sp = short( hi[d] - lo[d] );
    if ( sp > maxSpan ) maxSpan = sp;
  }
  int best = -1;
  feat     = 0;
  for ( int d = 0; d < 3; ++d ) {
    const short sp = short( hi[d] - lo[d] );
    if ( double( sp ) > ( 1.0 - 0.00001 ) * double( maxSpan ) ) {
      const short spread = short( mx[d] - mn[d] );
      if ( spread > best ) feat = d, best = spread;
    }
  }
  const int split = ( lo[feat] + hi[feat] ) / 2;
  cut             = split < mn[feat] ? mn[feat] : ( split > mx[feat] ? mx[feat] : split );
}

// the root's cell is the tight bounding box of the cloud (computeBoundingBox): its slot carries it in the min/max fields
__device__ __forceinline__ void slotCut( const SlotView& sv, int s, bool rootLevel, int& feat, int& cut ) {
  int lo[3], hi[3], mn[3], mx[3];
#pragma unroll
  for ( int d = 0; d < 3; ++d ) {
    mn[d] = sv.at( F_MM + d, s ), mx[d] = sv.at( F_MM + 3 + d, s );
    lo[d] = rootLevel ? mn[d] : sv.at( F_BOX + d, s ), hi[d] = rootLevel ? mx[d] : sv.at( F_BOX + 3 + d, s );
  }
  decideCut( lo, hi, mn, mx, feat, cut );
}

// ptsT is kept in tree order THROUGHOUT the build (ptsT[p] == pts[vind[p]]): every swap moves the point along with its index, so
// the per-level passes read coordinates coalesced instead of gathering pts[vind[p]].
__global__ void kInit( uint32_t* vind, int* slotOf, int n, SlotView sv, int* cnt, const short4* __restrict__ pts, short4* __restrict__ ptsT ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i < n ) vind[i] = i, slotOf[i] = 0, ptsT[i] = pts[i];
  if ( i == 0 ) {
    sv.at( F_NODE, 0 ) = 0, sv.at( F_LO, 0 ) = 0, sv.at( F_HI, 0 ) = n;
    for ( int d = 0; d < 3; ++d ) sv.at( F_MM + d, 0 ) = INT_MAX, sv.at( F_MM + 3 + d, 0 ) = INT_MIN;
    for ( int c = 0; c < C_COUNT; ++c ) cnt[c] = 0;
    cnt[C_NODES]     = 1;  // node 0 is the root
    cnt[C_LEVEL + 0] = 1;  // one slot at level 0 (unused when the root goes straight to the local phase)
  }
}

// min/max of the root. Consecutive positions share the slot, so a warp reduces with redux.sync first.
__global__ void kRootMinMax( SlotView sv, const short4* __restrict__ pts, int n ) {
  const int p   = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned  act = __ballot_sync( 0xffffffffu, p < n );
  if ( p >= n ) return;
  const short4 q    = pts[p];
  const int    v[3] = {q.x, q.y, q.z};
#pragma unroll
  for ( int d = 0; d < 3; ++d ) {
    const int mn = __reduce_min_sync( act, v[d] ), mx = __reduce_max_sync( act, v[d] );
    if ( ( threadIdx.x & 31 ) == __ffs( act ) - 1 ) {
      relaxMin( &sv.at( F_MM + d, 0 ), mn );
      relaxMax( &sv.at( F_MM + 3 + d, 0 ), mx );
    }
  }
}

// the root box for the searches; mode 1: a root of at most kLocalMax points is the one local subtree; mode 2: the root is a leaf
__global__ void kRootBox( SlotView sv, int* cnt, int* localRec, int4* nodes, int n, int mode ) {
  for ( int d = 0; d < 6; ++d ) cnt[C_ROOTBOX + d] = sv.at( F_MM + d, 0 );
  if ( mode != 0 ) {
    localRec[0] = mode == 2 ? -1 : 0, localRec[1] = 0, localRec[2] = n;
    for ( int d = 0; d < 6; ++d ) localRec[3 + d] = sv.at( F_MM + d, 0 );
    localRec[9]      = 0;
    cnt[C_LOCAL]     = 1;
    cnt[C_LEVEL + 0] = 0;
    if ( mode == 2 ) nodes[0] = make_int4( 0, n | kLeafBit, 0, 0 );
  }
}

// ---- TOP phase, launch 1 of 5: scan of the sweep-1 flags (v < cut over [lo, hi)); the decision is stored for the later launches
struct Sweep1Flags {
  SlotView        sv;
  const short4*   ptsT;
  const int*      slotOf;
  bool            rootLevel;
  int             lastS, feat, cut;
  __device__ __forceinline__ uint32_t operator()( size_t p ) {
    const int s = slotOf[p];
    if ( s < 0 ) return 0u;
    if ( s != lastS ) {
      slotCut( sv, s, rootLevel, feat, cut );
      lastS = s;
    }
    if ( int( p ) == sv.at( F_LO, s ) ) sv.at( F_FEAT, s ) = feat, sv.at( F_CUT, s ) = cut;
    return coord( ptsT[p], feat ) < cut ? 1u : 0u;
  }
};
__global__ void __launch_bounds__( kScanThreads )
    kScanSweep1( SlotView sv, const short4* __restrict__ ptsT, const int* __restrict__ slotOf, int n, int level,
                 const int* __restrict__ cnt, uint32_t* __restrict__ S, unsigned long long* __restrict__ ctl ) {
  if ( cnt[C_LEVEL + level] == 0 ) return;
  Sweep1Flags f{ sv, ptsT, slotOf, level == 0, -1, 0, 0 };
  scanLookbackTile( f, S, size_t( n ), ctl );
}

// ---- launches 2 and 4: misplaced elements write their POSITION'S INDEX to rank-indexed side lists.
//   left-misplaced  (inside the first `cnt` positions of the sweep range, flag 0): rank = #flag0 before it   -> tmpA[subLo + rank]
//   right-misplaced (beyond the first `cnt` positions, flag 1):                    rank = #flag1 after it    -> tmpB[subLo + rank]
// PASS 2 also creates the children of the level (threads 0 .. numSlots-1; independent of the compaction: lim2 follows from the scan)
template <int PASS>
__global__ void kCompact( SlotView sv, SlotView nx, const uint32_t* __restrict__ vind, const int* __restrict__ slotOf, const uint32_t* __restrict__ S,
                          uint32_t* __restrict__ tmpA, uint32_t* __restrict__ tmpB, int n, int level, int* __restrict__ cnt,
                          int4* __restrict__ nodes, int* __restrict__ localRec, int maxLocal ) {
  const int numSlots = cnt[C_LEVEL + level];
  if ( numSlots == 0 ) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if ( PASS == 2 && p == 0 ) atomicMax( &cnt[C_TOPDEPTH], level + 1 );
  if ( PASS == 2 && p < numSlots ) {
    // children (nanoflann.hpp:1138-1140 + divideTree :1061-1085)
    const int s  = p;
    const int lo = sv.at( F_LO, s ), hi = sv.at( F_HI, s ), count = hi - lo;
    const int lim1 = sv.at( F_LIM1, s ), lim2 = lim1 + int( S[hi] - S[lo + lim1] );
    const int idx  = lim1 > count / 2 ? lim1 : ( lim2 < count / 2 ? lim2 : count / 2 );
    sv.at( F_IDX, s ) = idx;
    const int  feat = sv.at( F_FEAT, s ), cut = sv.at( F_CUT, s );
    const bool rootLevel = level == 0;
    const int  child0    = atomicAdd( &cnt[C_NODES], 2 );
    nodes[sv.at( F_NODE, s )] = make_int4( child0, feat, INT_MIN, INT_MAX );  // divlow / divhigh follow by atomics (kApply2Assign)
    for ( int c = 0; c < 2; ++c ) {
      const int clo = c == 0 ? lo : lo + idx, chi = c == 0 ? lo + idx : hi;
      int       box[6];
#pragma unroll
      for ( int d = 0; d < 6; ++d ) box[d] = rootLevel ? sv.at( F_MM + d, s ) : sv.at( F_BOX + d, s );
      if ( c == 0 )
        box[3 + feat] = cut;  // left cell: high = cutval
      else
        box[feat] = cut;  // right cell: low = cutval
      int childSlot = -1;
      if ( chi - clo <= kLeafMaxSize ) nodes[child0 + c] = make_int4( clo, chi | kLeafBit, 0, 0 );
      if ( chi - clo <= kLocalMax ) {  // finished by the local phase (a leaf only has its points gathered there)
        const int at = atomicAdd( &cnt[C_LOCAL], 1 );
        if ( at < maxLocal ) {
          int* r = localRec + size_t( at ) * kLocalRecInts;
          r[0] = chi - clo <= kLeafMaxSize ? -1 : child0 + c, r[1] = clo, r[2] = chi;
#pragma unroll
          for ( int d = 0; d < 6; ++d ) r[3 + d] = box[d];
          r[9] = level + 1;
        } else {
          cnt[C_ERROR] = 1;
        }
      } else {
        childSlot = atomicAdd( &cnt[C_LEVEL + level + 1], 1 );
        nx.at( F_NODE, childSlot ) = child0 + c, nx.at( F_LO, childSlot ) = clo, nx.at( F_HI, childSlot ) = chi;
#pragma unroll
        for ( int d = 0; d < 6; ++d ) nx.at( F_BOX + d, childSlot ) = box[d];
#pragma unroll
        for ( int d = 0; d < 3; ++d ) nx.at( F_MM + d, childSlot ) = INT_MAX, nx.at( F_MM + 3 + d, childSlot ) = INT_MIN;
      }
      sv.at( F_CHILD0 + c, s ) = childSlot;
    }
  }
  if ( p >= n ) return;
  const int s = slotOf[p];
  if ( s < 0 ) return;
  const int subLo = PASS == 1 ? sv.at( F_LO, s ) : sv.at( F_LO, s ) + sv.at( F_LIM1, s );
  if ( p < subLo ) return;
  const int      hi  = sv.at( F_HI, s );
  const uint32_t cnt1 = S[hi] - S[subLo];
  const uint32_t i   = p - subLo;
  const uint32_t f   = S[p + 1] - S[p];
  if ( i < cnt1 && !f ) {
    tmpA[subLo + ( i - ( S[p] - S[subLo] ) )] = vind[p];
  } else if ( i >= cnt1 && f ) {
    tmpB[subLo + ( S[hi] - S[p + 1] )] = vind[p];
  }
}

// the swap of one sweep at position p: the i-th left-misplaced and the i-th (from the right) right-misplaced trade places
__device__ __forceinline__ void applySwap( uint32_t* __restrict__ vind, const short4* __restrict__ pts, short4* __restrict__ ptsT,
                                           const uint32_t* __restrict__ S, const uint32_t* __restrict__ tmpA, const uint32_t* __restrict__ tmpB, int p,
                                           int subLo, int hi, uint32_t& cntOut ) {
  const uint32_t cnt = S[hi] - S[subLo];
  cntOut             = cnt;
  if ( p < subLo ) return;
  const uint32_t i = p - subLo;
  const uint32_t f = S[p + 1] - S[p];
  uint32_t id = 0xFFFFFFFFu;
  if ( i < cnt && !f ) {
    id = tmpB[subLo + ( i - ( S[p] - S[subLo] ) )];
  } else if ( i >= cnt && f ) {
    id = tmpA[subLo + ( S[hi] - S[p + 1] )];
  }
  if ( id != 0xFFFFFFFFu ) vind[p] = id, ptsT[p] = pts[id];
}

// ---- launch 3: sweep-1 swap fused into the scan of the sweep-2 flags (v <= cut over [lo + lim1, hi))
struct Sweep2Flags {
  SlotView        sv;
  const short4*   pts;
  short4*         ptsT;
  uint32_t*       vind;
  const int*      slotOf;
  const uint32_t *S1, *tmpA, *tmpB;
  __device__ __forceinline__ uint32_t operator()( size_t p ) {
    const int s = slotOf[p];
    if ( s < 0 ) return 0u;
    const int lo = sv.at( F_LO, s ), hi = sv.at( F_HI, s );
    uint32_t  lim1;
    applySwap( vind, pts, ptsT, S1, tmpA, tmpB, int( p ), lo, hi, lim1 );
    if ( int( p ) == lo ) sv.at( F_LIM1, s ) = int( lim1 );
    if ( int( p ) < lo + int( lim1 ) ) return 0u;
    return coord( ptsT[p], sv.at( F_FEAT, s ) ) <= sv.at( F_CUT, s ) ? 1u : 0u;
  }
};
__global__ void __launch_bounds__( kScanThreads )
    kScanSweep2( SlotView sv, const short4* __restrict__ pts, short4* __restrict__ ptsT, uint32_t* __restrict__ vind, const int* __restrict__ slotOf, int n, int level,
                 const int* __restrict__ cnt, const uint32_t* __restrict__ S1, const uint32_t* __restrict__ tmpA, const uint32_t* __restrict__ tmpB,
                 uint32_t* __restrict__ S2, unsigned long long* __restrict__ ctl ) {
  if ( cnt[C_LEVEL + level] == 0 ) return;
  Sweep2Flags f{ sv, pts, ptsT, vind, slotOf, S1, tmpA, tmpB };
  scanLookbackTile( f, S2, size_t( n ), ctl );
}

// ---- launch 5: sweep-2 swap, tight split bounds (divlow = max of the left child along feat, divhigh = min of the right child,
// straight into the node record), the slot of every position for the next level and that level's min/max
__global__ void kApply2Assign( SlotView sv, SlotView nx, const short4* __restrict__ pts, short4* __restrict__ ptsT, uint32_t* __restrict__ vind, int* __restrict__ slotOf,
                               const uint32_t* __restrict__ S2, const uint32_t* __restrict__ tmpA, const uint32_t* __restrict__ tmpB, int n, int level,
                               const int* __restrict__ cnt, int4* __restrict__ nodes ) {
  if ( cnt[C_LEVEL + level] == 0 ) return;
  const int p   = blockIdx.x * blockDim.x + threadIdx.x;
  const int s   = p < n ? slotOf[p] : -1;
  unsigned  act = __ballot_sync( 0xffffffffu, s >= 0 );
  if ( s < 0 ) return;
  const int lo = sv.at( F_LO, s ), hi = sv.at( F_HI, s );
  uint32_t  unused;
  applySwap( vind, pts, ptsT, S2, tmpA, tmpB, p, lo + sv.at( F_LIM1, s ), hi, unused );
  const int    idx  = sv.at( F_IDX, s );
  const int    side = ( p - lo ) < idx ? 0 : 1;
  const short4 q    = ptsT[p];
  const int    v    = coord( q, sv.at( F_FEAT, s ) );
  const int    child = sv.at( F_CHILD0 + side, s );
  slotOf[p]         = child;
  int* const   node = reinterpret_cast<int*>( nodes + sv.at( F_NODE, s ) );
  const int    key    = s * 2 + side;
  const int    leader = __ffs( act ) - 1;
  const int    k0     = __shfl_sync( act, key, leader );
  const bool   same   = __all_sync( act, key == k0 );
  const int    c[3]   = {q.x, q.y, q.z};
  if ( same ) {
    const int r = side == 0 ? __reduce_max_sync( act, v ) : __reduce_min_sync( act, v );
    int       mn[3], mx[3];
#pragma unroll
    for ( int d = 0; d < 3; ++d ) mn[d] = __reduce_min_sync( act, c[d] ), mx[d] = __reduce_max_sync( act, c[d] );
    if ( ( threadIdx.x & 31 ) == leader ) {
      if ( side == 0 )
        relaxMax( node + 2, r );
      else
        relaxMin( node + 3, r );
      if ( child >= 0 ) {
#pragma unroll
        for ( int d = 0; d < 3; ++d ) relaxMin( &nx.at( F_MM + d, child ), mn[d] ), relaxMax( &nx.at( F_MM + 3 + d, child ), mx[d] );
      }
    }
  } else {
    if ( side == 0 )
      relaxMax( node + 2, v );
    else
      relaxMin( node + 3, v );
    if ( child >= 0 ) {
#pragma unroll
      for ( int d = 0; d < 3; ++d ) relaxMin( &nx.at( F_MM + d, child ), c[d] ), relaxMax( &nx.at( F_MM + 3 + d, child ), c[d] );
    }
  }
}

// ---- LOCAL phase ------------------------------------------------------------------------------------------------------------
struct LNode {
  int node, lo, hi, box[6], mm[6], feat, cut, lim1, idx, child[2], divlow, divhigh;
};
struct LocalSmem {
  short4   sp[kLocalMax];      // points of the subtree, in tree order as it evolves
  uint32_t sid[kLocalMax];     // their caller indices
  uint16_t S[kLocalMax + 2];   // exclusive scan of the sweep flags
  uint16_t A[kLocalMax], B[kLocalMax];  // misplaced positions by rank
  uint16_t nodeOf[kLocalMax];  // live node (index into cur) of every position, 0xFFFF = finished
  LNode    list[2][kLocalNodes];
  uint32_t warpSum[kLocalThreads / 32];
  int      nNext, ticket;
};

// exclusive scan of one flag per position over [0, m) (m <= kLocalMax), blocked arrangement, result in sm.S[0..m]
template <class Flag>
__device__ __forceinline__ void localScan( LocalSmem& sm, int m, Flag flag ) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int base = threadIdx.x * kLocalItems;
  uint32_t  f[kLocalItems], sum = 0;
#pragma unroll
  for ( int i = 0; i < kLocalItems; ++i ) {
    f[i] = ( base + i < m ) ? ( flag( base + i ) ? 1u : 0u ) : 0u;
    sum += f[i];
  }
  const uint32_t inc = scanWarpInclusive( sum, lane );
  if ( lane == 31 ) sm.warpSum[w] = inc;
  __syncthreads();
  if ( w == 0 ) {
    const uint32_t s  = lane < kLocalThreads / 32 ? sm.warpSum[lane] : 0;
    const uint32_t in = scanWarpInclusive( s, lane );
    if ( lane < kLocalThreads / 32 ) sm.warpSum[lane] = in - s;
  }
  __syncthreads();
  uint32_t ex = sm.warpSum[w] + ( inc - sum );
#pragma unroll
  for ( int i = 0; i < kLocalItems; ++i ) {
    if ( base + i <= m ) sm.S[base + i] = uint16_t( ex );
    ex += f[i];
  }
  if ( base + kLocalItems == m ) sm.S[m] = uint16_t( ex );  // (m a multiple of the items per thread: the next thread's base is m itself, covered above otherwise)
  __syncthreads();
}

// one Hoare sweep over every live node: positions [subLo(k), hi(k)) with the given flags already scanned into sm.S
template <class SubLo>
__device__ __forceinline__ void localSwap( LocalSmem& sm, const LNode* cur, int m, SubLo subLoOf, bool firstSweep, LNode* curW ) {
  // compaction of the misplaced positions
  for ( int i = threadIdx.x; i < m; i += kLocalThreads ) {
    const int k = sm.nodeOf[i];
    if ( k == 0xFFFF ) continue;
    const int subLo = subLoOf( k ), hi = cur[k].hi;
    if ( i < subLo ) continue;
    const int cnt = sm.S[hi] - sm.S[subLo], j = i - subLo, f = sm.S[i + 1] - sm.S[i];
    if ( j < cnt && !f )
      sm.A[subLo + ( j - ( sm.S[i] - sm.S[subLo] ) )] = uint16_t( i );
    else if ( j >= cnt && f )
      sm.B[subLo + ( sm.S[hi] - sm.S[i + 1] )] = uint16_t( i );
  }
  __syncthreads();
  // swap in two steps: every misplaced position reads its partner's point, then all write
  short4   np[kLocalItems];
  uint32_t ni[kLocalItems];
  bool     mv[kLocalItems];
#pragma unroll
  for ( int t = 0; t < kLocalItems; ++t ) {
    const int i = threadIdx.x + t * kLocalThreads;
    mv[t]       = false;
    if ( i >= m ) continue;
    const int k = sm.nodeOf[i];
    if ( k == 0xFFFF ) continue;
    const int subLo = subLoOf( k ), hi = cur[k].hi;
    if ( i < subLo ) continue;
    const int cnt = sm.S[hi] - sm.S[subLo], j = i - subLo, f = sm.S[i + 1] - sm.S[i];
    int       q = -1;
    if ( j < cnt && !f )
      q = sm.B[subLo + ( j - ( sm.S[i] - sm.S[subLo] ) )];
    else if ( j >= cnt && f )
      q = sm.A[subLo + ( sm.S[hi] - sm.S[i + 1] )];
    if ( q >= 0 ) mv[t] = true, np[t] = sm.sp[q], ni[t] = sm.sid[q];
    if ( j == 0 ) {
      if ( firstSweep )
        curW[k].lim1 = cnt;
      else
        curW[k].idx = cnt;  // (second sweep: its count, turned into lim2 / idx by the children step)
    }
  }
  __syncthreads();
#pragma unroll
  for ( int t = 0; t < kLocalItems; ++t ) {
    const int i = threadIdx.x + t * kLocalThreads;
    if ( mv[t] ) sm.sp[i] = np[t], sm.sid[i] = ni[t];
  }
  __syncthreads();
}

__global__ void __launch_bounds__( kLocalThreads )
    kLocalSubtrees( uint32_t* __restrict__ vind, short4* __restrict__ ptsT, int4* __restrict__ nodes,
                    const int* __restrict__ localRec, int* __restrict__ cnt, int maxLocal ) {
  extern __shared__ __align__( 16 ) unsigned char smemRaw[];
  LocalSmem& sm = *reinterpret_cast<LocalSmem*>( smemRaw );
  const int  lane = threadIdx.x & 31;
  for ( ;; ) {
    __syncthreads();
    if ( threadIdx.x == 0 ) sm.ticket = atomicAdd( &cnt[C_TICKET], 1 );
    __syncthreads();
    const int rec = sm.ticket;
    if ( rec >= min( *reinterpret_cast<volatile int*>( &cnt[C_LOCAL] ), maxLocal ) ) return;
    const int* r   = localRec + size_t( rec ) * kLocalRecInts;
    const int  glo = r[1], m = r[2] - r[1];
    for ( int i = threadIdx.x; i < m; i += kLocalThreads ) {
      sm.sid[i] = vind[glo + i], sm.sp[i] = ptsT[glo + i];
      sm.nodeOf[i] = 0;
    }
    int nCur = 0, depth = r[9], which = 0;
    if ( r[0] >= 0 ) {  // (a leaf created by the top phase only has its points gathered)
      nCur = 1;
      if ( threadIdx.x == 0 ) {
        LNode& nd = sm.list[0][0];
        nd.node = r[0], nd.lo = 0, nd.hi = m;
        for ( int d = 0; d < 6; ++d ) nd.box[d] = r[3 + d];
      }
    }
    __syncthreads();
    while ( nCur > 0 ) {
      LNode* cur = sm.list[which];
      LNode* nxt = sm.list[which ^ 1];
      // 1. min/max per node
      for ( int k = threadIdx.x; k < nCur; k += kLocalThreads )
        for ( int d = 0; d < 3; ++d ) cur[k].mm[d] = INT_MAX, cur[k].mm[3 + d] = INT_MIN;
      if ( threadIdx.x == 0 ) sm.nNext = 0;
      __syncthreads();
      for ( int i0 = 0; i0 < m; i0 += kLocalThreads ) {  // (whole warps stay in the loop: the reductions below are warp-wide)
        const int      i   = i0 + threadIdx.x;
        const int      k   = i < m ? sm.nodeOf[i] : 0xFFFF;
        const unsigned act = __ballot_sync( 0xffffffffu, k != 0xFFFF );
        if ( k == 0xFFFF ) continue;
        const short4 q      = sm.sp[i];
        const int    v[3]   = {q.x, q.y, q.z};
        const int    leader = __ffs( act ) - 1;
        const bool   same   = __all_sync( act, k == __shfl_sync( act, k, leader ) );
#pragma unroll
        for ( int d = 0; d < 3; ++d ) {
          if ( same ) {
            const int mn = __reduce_min_sync( act, v[d] ), mx = __reduce_max_sync( act, v[d] );
            if ( lane == leader ) atomicMin( &cur[k].mm[d], mn ), atomicMax( &cur[k].mm[3 + d], mx );
          } else {
            atomicMin( &cur[k].mm[d], v[d] ), atomicMax( &cur[k].mm[3 + d], v[d] );
          }
        }
      }
      __syncthreads();
      // 2. cut per node
      for ( int k = threadIdx.x; k < nCur; k += kLocalThreads ) decideCut( cur[k].box, cur[k].box + 3, cur[k].mm, cur[k].mm + 3, cur[k].feat, cur[k].cut );
      __syncthreads();
      // 3. sweep 1: [ < cut | >= cut ] over [lo, hi)
      localScan( sm, m, [&]( int i ) {
        const int k = sm.nodeOf[i];
        return k != 0xFFFF && coord( sm.sp[i], cur[k].feat ) < cur[k].cut;
      } );
      localSwap( sm, cur, m, [&]( int k ) { return cur[k].lo; }, true, cur );
      // 4. sweep 2: [ == cut | > cut ] over [lo + lim1, hi)
      localScan( sm, m, [&]( int i ) {
        const int k = sm.nodeOf[i];
        return k != 0xFFFF && i >= cur[k].lo + cur[k].lim1 && coord( sm.sp[i], cur[k].feat ) <= cur[k].cut;
      } );
      for ( int k = threadIdx.x; k < nCur; k += kLocalThreads ) cur[k].idx = 0;  // (set by the sweep when its range is not empty)
      __syncthreads();
      localSwap( sm, cur, m, [&]( int k ) { return cur[k].lo + cur[k].lim1; }, false, cur );
      // 5. children
      for ( int k = threadIdx.x; k < nCur; k += kLocalThreads ) {
        LNode&    nd    = cur[k];
        const int count = nd.hi - nd.lo, lim1 = nd.lim1, lim2 = lim1 + nd.idx;
        const int idx   = lim1 > count / 2 ? lim1 : ( lim2 < count / 2 ? lim2 : count / 2 );
        nd.idx          = idx;
        const int child0 = atomicAdd( &cnt[C_NODES], 2 );
        nd.divlow = INT_MIN, nd.divhigh = INT_MAX;
        for ( int c = 0; c < 2; ++c ) {
          const int clo = c == 0 ? nd.lo : nd.lo + idx, chi = c == 0 ? nd.lo + idx : nd.hi;
          int       slot = -1;
          if ( chi - clo <= kLeafMaxSize ) {
            nodes[child0 + c] = make_int4( glo + clo, ( glo + chi ) | kLeafBit, 0, 0 );
          } else {
            slot      = atomicAdd( &sm.nNext, 1 );
            LNode& ch = nxt[slot];
            ch.node = child0 + c, ch.lo = clo, ch.hi = chi;
            for ( int d = 0; d < 6; ++d ) ch.box[d] = nd.box[d];
            if ( c == 0 )
              ch.box[3 + nd.feat] = nd.cut;
            else
              ch.box[nd.feat] = nd.cut;
          }
          nd.child[c] = slot;
        }
        nd.lim1 = child0;  // (lim1 is dead: keeps the children's node id for step 7)
      }
      __syncthreads();
      // 6. tight split bounds + next node of every position
      for ( int i0 = 0; i0 < m; i0 += kLocalThreads ) {
        const int      i   = i0 + threadIdx.x;
        const int      k   = i < m ? sm.nodeOf[i] : 0xFFFF;
        const unsigned act = __ballot_sync( 0xffffffffu, k != 0xFFFF );
        if ( k == 0xFFFF ) continue;
        const int  side   = ( i - cur[k].lo ) < cur[k].idx ? 0 : 1;
        const int  v      = coord( sm.sp[i], cur[k].feat );
        const int  key    = k * 2 + side;
        const int  leader = __ffs( act ) - 1;
        const bool same   = __all_sync( act, key == __shfl_sync( act, key, leader ) );
        if ( same ) {
          const int rr = side == 0 ? __reduce_max_sync( act, v ) : __reduce_min_sync( act, v );
          if ( lane == leader ) {
            if ( side == 0 )
              atomicMax( &cur[k].divlow, rr );
            else
              atomicMin( &cur[k].divhigh, rr );
          }
        } else {
          if ( side == 0 )
            atomicMax( &cur[k].divlow, v );
          else
            atomicMin( &cur[k].divhigh, v );
        }
        const int ch  = cur[k].child[side];
        sm.nodeOf[i] = ch < 0 ? uint16_t( 0xFFFF ) : uint16_t( ch );
      }
      __syncthreads();
      // 7. node records of this level
      for ( int k = threadIdx.x; k < nCur; k += kLocalThreads ) nodes[cur[k].node] = make_int4( cur[k].lim1, cur[k].feat, cur[k].divlow, cur[k].divhigh );
      nCur = sm.nNext;
      which ^= 1;
      ++depth;
      __syncthreads();
    }
    for ( int i = threadIdx.x; i < m; i += kLocalThreads ) vind[glo + i] = sm.sid[i], ptsT[glo + i] = sm.sp[i];
    if ( threadIdx.x == 0 ) atomicMax( &cnt[C_DEPTH], depth );
  }
}

// ------------------------------------------------------------------------------------------- k-NN search
constexpr int kMaxStack = 96;

template <int K>
__global__ void __launch_bounds__( 128 )
    kKnn( const int4* __restrict__ nodes, const short4* __restrict__ ptsT, const uint32_t* __restrict__ vind, const int* __restrict__ rootBox,
          const short4* __restrict__ queries, const uint32_t* __restrict__ order, int nq, int treeSize,
          uint32_t* __restrict__ outIdx, float* __restrict__ outDist ) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if ( t >= nq ) return;
  const uint32_t qi = order ? order[t] : uint32_t( t );
  const short4   q  = queries[qi];
  const float    qf[3] = {float( q.x ), float( q.y ), float( q.z )};

  float    bd[K];  // ascending; FLT_MAX = empty
  uint32_t bi[K];
#pragma unroll
  for ( int j = 0; j < K; ++j ) bd[j] = 3.402823466e+38f, bi[j] = 0xFFFFFFFFu;

  // explicit stack: node, lower bound of the cell distance, per-dimension contributions to that bound
  int   stNode[kMaxStack];
  float stMin[kMaxStack];
  float stSide[kMaxStack][3];
  int   sp = 0;
  {
    float side[3] = {0.f, 0.f, 0.f}, mind = 0.f;
#pragma unroll
    for ( int d = 0; d < 3; ++d ) {
      if ( qf[d] < float( rootBox[d] ) ) {
        float e = qf[d] - float( rootBox[d] );
        side[d] = e * e;
        mind += side[d];
      }
      if ( qf[d] > float( rootBox[3 + d] ) ) {
        float e = qf[d] - float( rootBox[3 + d] );
        side[d] = e * e;
        mind += side[d];
      }
    }
    stNode[0] = 0, stMin[0] = mind;
    stSide[0][0] = side[0], stSide[0][1] = side[1], stSide[0][2] = side[2];
    sp = treeSize > 0 ? 1 : 0;
  }
  while ( sp > 0 ) {
    --sp;
    int   node = stNode[sp];
    float mind = stMin[sp];
    if ( !( mind <= bd[K - 1] ) ) continue;  // the far-child test of searchLevel, evaluated when the recursion would
    float side[3] = {stSide[sp][0], stSide[sp][1], stSide[sp][2]};
    int4  nd      = nodes[node];
    while ( !( nd.y & kLeafBit ) ) {
      const int   f     = nd.y;
      const float v     = qf[f];
      const float diff1 = v - float( nd.z ), diff2 = v - float( nd.w );
      int         nearC, farC;
      float       cut;
      if ( diff1 + diff2 < 0.f ) {
        nearC = nd.x, farC = nd.x + 1, cut = diff2 * diff2;
      } else {
        nearC = nd.x + 1, farC = nd.x, cut = diff1 * diff1;
      }
      // push the far child with its own bound
      stNode[sp]    = farC;
      stMin[sp]     = mind + cut - side[f];
      stSide[sp][0] = f == 0 ? cut : side[0];
      stSide[sp][1] = f == 1 ? cut : side[1];
      stSide[sp][2] = f == 2 ? cut : side[2];
      ++sp;
      node = nearC;
      nd   = nodes[node];
    }
    const int   lo = nd.x, hi = nd.y & 0x7fffffff;
    const float worstAtEntry = bd[K - 1];
    for ( int p = lo; p < hi; ++p ) {
      const short4 c  = ptsT[p];
      const float  dx = qf[0] - float( c.x ), dy = qf[1] - float( c.y ), dz = qf[2] - float( c.z );
      float        d  = dx * dx;
      d += dy * dy;
      d += dz * dz;
      if ( d < worstAtEntry && d < bd[K - 1] ) {
        // insert after all entries <= d (equal distances keep arrival order)
        const uint32_t id = vind[p];
#pragma unroll
        for ( int j = K - 1; j > 0; --j ) {
          if ( bd[j - 1] > d ) {
            bd[j] = bd[j - 1], bi[j] = bi[j - 1];
          } else if ( bd[j] > d ) {
            bd[j] = d, bi[j] = id;
          }
        }
        if ( bd[0] > d ) bd[0] = d, bi[0] = id;
      }
    }
  }
#pragma unroll
  for ( int j = 0; j < K; ++j ) {
    outIdx[size_t( qi ) * K + j] = bi[j];
    if ( outDist ) outDist[size_t( qi ) * K + j] = bi[j] == 0xFFFFFFFFu ? -1.0f : bd[j];
  }
}

}  // namespace

// ======================================================================================================
void kdBuild( KdTree& t, const short4* xyz4, size_t n, cudaStream_t s ) {
  t.n         = n;
  t.numNodes  = 0;
  t.numLevels = 0;
  if ( n == 0 ) return;
  const int    N        = int( n );
  const int    maxSlots = N / ( kLocalMax + 1 ) + 2;       // nodes of more than kLocalMax points alive at one level
  const size_t maxLocal = size_t( N ) / 32 + 64;           // local subtrees + leaves split off by the top phase
  const size_t ctlWords = scanCtlWords( n );               // uint64 words of one look-back scan
  t.pts.reserve( n );
  PCC_CUDA( cudaMemcpyAsync( t.pts, xyz4, n * sizeof( short4 ), cudaMemcpyDeviceToDevice, s ) );
  t.ptsT.reserve( n ), t.vind.reserve( n ), t.nodes.reserve( 2 * n + 2 );
  t.tmpA.reserve( n ), t.tmpB.reserve( n ), t.flags.reserve( n + 1 ), t.scanOut.reserve( n + 1 );
  t.scanTmp.reserve( 2 * ( 2 * kMaxTopLevels * ctlWords ) + 4 );
  t.slotOf.reserve( n ), t.slotOfNext.reserve( maxLocal * kLocalRecInts );
  t.slotI[0].reserve( size_t( F_COUNT ) * maxSlots ), t.slotI[1].reserve( size_t( F_COUNT ) * maxSlots );
  t.counters.reserve( C_COUNT + 8 );
  t.hostInts.reserve( 32 );
  {  // the local phase needs more than the default 48 KB of dynamic shared memory: per function AND per device
    static std::mutex m;
    static bool       raised[64] = { false };
    int               dev        = 0;
    PCC_CUDA( cudaGetDevice( &dev ) );
    std::lock_guard<std::mutex> lk( m );
    if ( dev < 0 || dev >= 64 || !raised[dev] ) {
      PCC_CUDA( cudaFuncSetAttribute( kLocalSubtrees, cudaFuncAttributeMaxDynamicSharedMemorySize, int( sizeof( LocalSmem ) ) ) );
      if ( dev >= 0 && dev < 64 ) raised[dev] = true;
    }
  }
  const int TB = 256, gridN = divUp( n, TB ), gridScan = int( scanTiles( n ) );
  int* const      cnt      = t.counters;
  int* const      localRec = t.slotOfNext;
  uint32_t* const S1       = t.scanOut;
  uint32_t* const S2       = t.flags;
  unsigned long long* const ctl = reinterpret_cast<unsigned long long*>( ( reinterpret_cast<uintptr_t>( t.scanTmp.p ) + 7 ) & ~uintptr_t( 7 ) );
  const int mode = N <= kLeafMaxSize ? 2 : ( N <= kLocalMax ? 1 : 0 );  // root: leaf / one local subtree / top phase

  kInit<<<gridN, TB, 0, s>>>( t.vind, t.slotOf, N, SlotView{ t.slotI[0], maxSlots }, cnt, t.pts, t.ptsT );
  kRootMinMax<<<gridN, TB, 0, s>>>( SlotView{ t.slotI[0], maxSlots }, t.pts, N );
  kRootBox<<<1, 1, 0, s>>>( SlotView{ t.slotI[0], maxSlots }, cnt, localRec, t.nodes, N, mode );
  PCC_LAUNCH_CHECK();

  auto launchLevel = [&]( int level ) {
    SlotView sv{ t.slotI[level & 1], maxSlots }, nx{ t.slotI[( level + 1 ) & 1], maxSlots };
    unsigned long long* c1 = ctl + size_t( 2 * level ) * ctlWords;
    unsigned long long* c2 = c1 + ctlWords;
    kScanSweep1<<<gridScan, kScanThreads, 0, s>>>( sv, t.ptsT, t.slotOf, N, level, cnt, S1, c1 );
    kCompact<1><<<gridN, TB, 0, s>>>( sv, nx, t.vind, t.slotOf, S1, t.tmpA, t.tmpB, N, level, cnt, t.nodes, localRec, int( maxLocal ) );
    kScanSweep2<<<gridScan, kScanThreads, 0, s>>>( sv, t.pts, t.ptsT, t.vind, t.slotOf, N, level, cnt, S1, t.tmpA, t.tmpB, S2, c2 );
    kCompact<2><<<gridN, TB, 0, s>>>( sv, nx, t.vind, t.slotOf, S2, t.tmpA, t.tmpB, N, level, cnt, t.nodes, localRec, int( maxLocal ) );
    kApply2Assign<<<gridN, TB, 0, s>>>( sv, nx, t.pts, t.ptsT, t.vind, t.slotOf, S2, t.tmpA, t.tmpB, N, level, cnt, t.nodes );
    PCC_LAUNCH_CHECK();
  };
  auto launchLocal = [&]() {
    const int grid = std::max( 1, std::min( 148 * 2, divUp( n, 1024 ) ) );
    kLocalSubtrees<<<grid, kLocalThreads, sizeof( LocalSmem ), s>>>( t.vind, t.ptsT, t.nodes, localRec, cnt, int( maxLocal ) );
    PCC_LAUNCH_CHECK();
  };
  // read-back: the first 16 counters (nodes, local records, depth, ticket, overflow flag, top depth, ..., root box at 8) + slots left at `level`
  auto readBack = [&]( int level ) {
    PCC_CUDA( cudaMemcpyAsync( t.hostInts.p, cnt, 16 * sizeof( int ), cudaMemcpyDeviceToHost, s ) );
    PCC_CUDA( cudaMemcpyAsync( t.hostInts.p + 16, cnt + C_LEVEL + level, sizeof( int ), cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    if ( t.hostInts.p[C_ERROR] ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };  // (more subtrees than a kd-tree over integer coordinates can have)
  };

  int level = 0;
  if ( mode == 0 ) {
    // levels launched before the first look at the device: the depth the previous tree built here needed (+1), else an estimate
    int want = t.lastTopLevels > 0 ? t.lastTopLevels + 1 : 2;
    if ( t.lastTopLevels <= 0 )
      for ( size_t c = n; c > size_t( kLocalMax ); c = c * 2 / 3 ) ++want;  // (spatial middle splits: deeper than a median tree)
    want = std::min( want, kMaxTopLevels );
    PCC_CUDA( cudaMemsetAsync( ctl, 0, size_t( 2 * want ) * ctlWords * sizeof( unsigned long long ), s ) );
    for ( ; level < want; ++level ) launchLevel( level );
  }
  launchLocal();
  readBack( level );
  while ( mode == 0 && t.hostInts.p[16] > 0 ) {  // deeper than expected: one more level at a time
    if ( level >= kMaxTopLevels ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
    const int doneLocal = t.hostInts.p[C_LOCAL];
    PCC_CUDA( cudaMemsetAsync( ctl + size_t( 2 * level ) * ctlWords, 0, 2 * ctlWords * sizeof( unsigned long long ), s ) );
    launchLevel( level );
    ++level;
    PCC_CUDA( cudaMemcpyAsync( cnt + C_TICKET, &doneLocal, sizeof( int ), cudaMemcpyHostToDevice, s ) );  // (pageable source: copied before the call returns)
    launchLocal();
    readBack( level );
  }
  if ( mode == 0 ) t.lastTopLevels = t.hostInts.p[C_TOPDEPTH];  // levels that actually held slots (for the next build of a similar cloud)
  t.numNodes  = t.hostInts.p[C_NODES];
  t.numLevels = t.hostInts.p[C_DEPTH] + 1;
  for ( int d = 0; d < 6; ++d ) t.rootBox[d] = t.hostInts.p[C_ROOTBOX + d];
}

void kdKnn( const KdTree& t, const short4* queries, size_t nq, const uint32_t* queryOrder, int k, uint32_t* outIdx,
            float* outDist, cudaStream_t s ) {
  if ( nq == 0 ) return;
  // the traversal stack holds one far child per level: a deeper tree (degenerate input) would overflow it silently
  if ( t.numLevels + 2 > kMaxStack ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
  const int* rb = t.counters.p + 8;  // the root box stays on the device (C_ROOTBOX)
  const int TB = 128, grid = divUp( nq, TB );
  switch ( k ) {
    case 16: kKnn<16><<<grid, TB, 0, s>>>( t.nodes, t.ptsT, t.vind, rb, queries, queryOrder, int( nq ), int( t.n ), outIdx, outDist ); break;
    case 8: kKnn<8><<<grid, TB, 0, s>>>( t.nodes, t.ptsT, t.vind, rb, queries, queryOrder, int( nq ), int( t.n ), outIdx, outDist ); break;
    case 1: kKnn<1><<<grid, TB, 0, s>>>( t.nodes, t.ptsT, t.vind, rb, queries, queryOrder, int( nq ), int( t.n ), outIdx, outDist ); break;
    default: throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
  }
  PCC_LAUNCH_CHECK();
}

}  // namespace pccb200
