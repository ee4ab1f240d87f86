// kdtree.cu — level-synchronous GPU construction of nanoflann's kd-tree and exact k-NN by per-thread traversal.
//
// Build (restates nanoflann.hpp:1041-1181 as data-parallel passes, one tree level at a time):
//   for every node of the level ("slot") with more than 10 points
//     1. per-dimension min/max of its points                (computeMinMax)        -> warp-aggregated atomics
//     2. cut dimension / cut value                          (middleSplit_)         -> one thread per slot
//     3. Hoare sweep #1  [<cut | >=cut], sweep #2 [==cut | >cut]   (planeSplit)    -> flags + device-wide scan:
//        the sequential two-pointer sweep swaps the i-th misplaced element from the left with the i-th
//        misplaced element from the right, so ranks from one prefix sum reproduce its permutation exactly.
//     4. split position (lim1/lim2/half rule), children, tight divlow/divhigh (max of left / min of right)
// Search (nanoflann.hpp:1207-1254): explicit-stack depth-first traversal, near child first, far child tested
// against the current worst distance when it is popped — the same moment the recursion would test it.
#include <limits.h>

#include "kdtree.cuh"

namespace pccb200 {

namespace {

enum SlotField {
  F_NODE = 0, F_LO, F_HI, F_BOX /*6*/, F_MM = F_BOX + 6 /*6*/, F_FEAT = F_MM + 6, F_CUT, F_SUBLO, F_LIM1, F_LIM2, F_IDX,
  F_DIVLOW, F_DIVHIGH, F_CHILD0, F_CHILD1, F_COUNT
};

struct SlotView {
  int* base;
  int  stride;
  __device__ __forceinline__ int& at( int field, int s ) const { return base[size_t( field ) * stride + s]; }
};

__device__ __forceinline__ int coord( const short4& p, int d ) { return d == 0 ? p.x : ( d == 1 ? p.y : p.z ); }

__global__ void kInitRoot( SlotView sv, int n ) {
  sv.at( F_NODE, 0 ) = 0;
  sv.at( F_LO, 0 )   = 0;
  sv.at( F_HI, 0 )   = n;
}

__global__ void kIota( uint32_t* v, int* slotOf, int n, int slot ) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i < n ) {
    v[i]      = i;
    slotOf[i] = slot;
  }
}

__global__ void kResetMinMax( SlotView sv, int numSlots ) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if ( s >= numSlots ) return;
#pragma unroll
  for ( int d = 0; d < 3; ++d ) {
    sv.at( F_MM + d, s )     = INT_MAX;
    sv.at( F_MM + 3 + d, s ) = INT_MIN;
  }
  sv.at( F_DIVLOW, s )  = INT_MIN;
  sv.at( F_DIVHIGH, s ) = INT_MAX;
}

// same-address extremes: read (L2) before the atomic; a stale value can only cause a redundant atomic (see patches.cu relaxMin)
__device__ __forceinline__ void relaxMin( int* addr, int v ) {
  if ( v < __ldcg( addr ) ) atomicMin( addr, v );
}
__device__ __forceinline__ void relaxMax( int* addr, int v ) {
  if ( v > __ldcg( addr ) ) atomicMax( addr, v );
}

// step 1: min/max per slot and dimension. Consecutive positions almost always share a slot, so a warp first
// agrees on one slot and reduces with redux.sync; mixed warps fall back to per-lane atomics.
__global__ void kMinMax( SlotView sv, const short4* __restrict__ pts, const uint32_t* __restrict__ vind,
                         const int* __restrict__ slotOf, int n ) {
  int       p    = blockIdx.x * blockDim.x + threadIdx.x;
  const int s    = p < n ? slotOf[p] : -1;
  unsigned  act  = __ballot_sync( 0xffffffffu, s >= 0 );
  if ( s < 0 ) return;
  const short4 q      = pts[vind[p]];
  const int    leader = __ffs( act ) - 1;
  const int    s0     = __shfl_sync( act, s, leader );
  const bool   same   = __all_sync( act, s == s0 );
  const int    v[3]   = {q.x, q.y, q.z};
  if ( same ) {
#pragma unroll
    for ( int d = 0; d < 3; ++d ) {
      int mn = __reduce_min_sync( act, v[d] ), mx = __reduce_max_sync( act, v[d] );
      if ( ( threadIdx.x & 31 ) == leader ) {
        relaxMin( &sv.at( F_MM + d, s ), mn );
        relaxMax( &sv.at( F_MM + 3 + d, s ), mx );
      }
    }
  } else {
#pragma unroll
    for ( int d = 0; d < 3; ++d ) {
      relaxMin( &sv.at( F_MM + d, s ), v[d] );
      relaxMax( &sv.at( F_MM + 3 + d, s ), v[d] );
    }
  }
}

__global__ void kRootBox( SlotView sv, int* rootBox ) {
  for ( int d = 0; d < 6; ++d ) {
    sv.at( F_BOX + d, 0 ) = sv.at( F_MM + d, 0 );
    rootBox[d]            = sv.at( F_MM + d, 0 );
  }
}

// step 2 (middleSplit_, nanoflann.hpp:1103-1131). Spans are int16 in the reference (ElementType), compared in double.
__global__ void kDecide( SlotView sv, int numSlots ) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if ( s >= numSlots ) return;
  int lo[3], hi[3], mn[3], mx[3];
#pragma unroll
  for ( int d = 0; d < 3; ++d ) {
    lo[d] = sv.at( F_BOX + d, s ), hi[d] = sv.at( F_BOX + 3 + d, s );
    mn[d] = sv.at( F_MM + d, s ), mx[d] = sv.at( F_MM + 3 + d, s );
  }
  short maxSpan = short( hi[0] - lo[0] );
  for ( int d = 1; d < 3; ++d ) {
    short sp = short( hi[d] - lo[d] );
    if ( sp > maxSpan ) maxSpan = sp;
  }
  int feat = 0, best = -1;
  for ( int d = 0; d < 3; ++d ) {
    short sp = short( hi[d] - lo[d] );
    if ( double( sp ) > ( 1.0 - 0.00001 ) * double( maxSpan ) ) {
      short spread = short( mx[d] - mn[d] );
      if ( spread > best ) feat = d, best = spread;
    }
  }
  int split = ( lo[feat] + hi[feat] ) / 2;
  int cut   = split < mn[feat] ? mn[feat] : ( split > mx[feat] ? mx[feat] : split );
  sv.at( F_FEAT, s )  = feat;
  sv.at( F_CUT, s )   = cut;
  sv.at( F_SUBLO, s ) = sv.at( F_LO, s );
}

// step 3a: flags for sweep PASS (1: v < cut over [lo,hi) ; 2: v <= cut over [lo+lim1,hi))
template <int PASS>
__global__ void kFlags( SlotView sv, const short4* __restrict__ pts, const uint32_t* __restrict__ vind,
                        const int* __restrict__ slotOf, uint32_t* __restrict__ flags, int n ) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if ( p >= n ) return;
  const int s = slotOf[p];
  uint32_t  f = 0;
  if ( s >= 0 && p >= sv.at( F_SUBLO, s ) ) {
    const int v = coord( pts[vind[p]], sv.at( F_FEAT, s ) ), cut = sv.at( F_CUT, s );
    f           = PASS == 1 ? ( v < cut ) : ( v <= cut );
  }
  flags[p] = f;
}

// step 3b: misplaced elements write themselves to rank-indexed side lists.
//   left-misplaced  (inside the first `cnt` positions, flag 0): rank = #flag0 before it        -> tmpA[subLo + rank]
//   right-misplaced (beyond the first `cnt` positions, flag 1): rank = #flag1 after it         -> tmpB[subLo + rank]
__global__ void kCompact( SlotView sv, const uint32_t* __restrict__ vind, const int* __restrict__ slotOf,
                          const uint32_t* __restrict__ flags, const uint32_t* __restrict__ S, uint32_t* __restrict__ tmpA,
                          uint32_t* __restrict__ tmpB, int n ) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if ( p >= n ) return;
  const int s = slotOf[p];
  if ( s < 0 ) return;
  const int subLo = sv.at( F_SUBLO, s );
  if ( p < subLo ) return;
  const int      hi  = sv.at( F_HI, s );
  const uint32_t cnt = S[hi] - S[subLo];
  const uint32_t i   = p - subLo;
  const uint32_t f   = flags[p];
  if ( i < cnt && !f ) {
    tmpA[subLo + ( i - ( S[p] - S[subLo] ) )] = vind[p];
  } else if ( i >= cnt && f ) {
    tmpB[subLo + ( S[hi] - S[p + 1] )] = vind[p];
  }
}

// step 3c: the i-th left-misplaced and the i-th (from the right) right-misplaced trade places.
template <int PASS>
__global__ void kApply( SlotView sv, uint32_t* __restrict__ vind, const int* __restrict__ slotOf,
                        const uint32_t* __restrict__ flags, const uint32_t* __restrict__ S, const uint32_t* __restrict__ tmpA,
                        const uint32_t* __restrict__ tmpB, int n ) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if ( p >= n ) return;
  const int s = slotOf[p];
  if ( s < 0 ) return;
  const int subLo = sv.at( F_SUBLO, s );
  if ( p < subLo ) return;
  const int      hi  = sv.at( F_HI, s );
  const uint32_t cnt = S[hi] - S[subLo];
  const uint32_t i   = p - subLo;
  const uint32_t f   = flags[p];
  if ( i < cnt && !f ) {
    vind[p] = tmpB[subLo + ( i - ( S[p] - S[subLo] ) )];
  } else if ( i >= cnt && f ) {
    vind[p] = tmpA[subLo + ( S[hi] - S[p + 1] )];
  }
  if ( i == 0 ) {
    if ( PASS == 1 ) {
      sv.at( F_LIM1, s ) = int( cnt );
    } else {
      sv.at( F_LIM2, s ) = sv.at( F_LIM1, s ) + int( cnt );
    }
  }
}

// between the sweeps: the second sweep only looks at [lo+lim1, hi)
__global__ void kAdvanceSubLo( SlotView sv, int numSlots ) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if ( s >= numSlots ) return;
  sv.at( F_SUBLO, s ) = sv.at( F_LO, s ) + sv.at( F_LIM1, s );
}

// step 4 (nanoflann.hpp:1138-1140 + divideTree :1061-1085): children, next level's slots.
__global__ void kChildren( SlotView sv, SlotView next, int numSlots, int nodeBase, int4* __restrict__ nodes, int* counters ) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if ( s >= numSlots ) return;
  const int lo = sv.at( F_LO, s ), hi = sv.at( F_HI, s ), count = hi - lo;
  const int lim1 = sv.at( F_LIM1, s ), lim2 = sv.at( F_LIM2, s );
  int       idx;
  if ( lim1 > count / 2 )
    idx = lim1;
  else if ( lim2 < count / 2 )
    idx = lim2;
  else
    idx = count / 2;
  sv.at( F_IDX, s ) = idx;
  const int feat = sv.at( F_FEAT, s ), cut = sv.at( F_CUT, s );
  const int child0 = nodeBase + 2 * s;
  for ( int c = 0; c < 2; ++c ) {
    const int clo = c == 0 ? lo : lo + idx, chi = c == 0 ? lo + idx : hi;
    int       childSlot = -1;
    if ( chi - clo <= kLeafMaxSize ) {
      nodes[child0 + c] = make_int4( clo, chi | kLeafBit, 0, 0 );
    } else {
      childSlot                    = atomicAdd( &counters[0], 1 );
      next.at( F_NODE, childSlot ) = child0 + c;
      next.at( F_LO, childSlot )   = clo;
      next.at( F_HI, childSlot )   = chi;
#pragma unroll
      for ( int d = 0; d < 6; ++d ) next.at( F_BOX + d, childSlot ) = sv.at( F_BOX + d, s );
      if ( c == 0 )
        next.at( F_BOX + 3 + feat, childSlot ) = cut;  // left cell: high = cutval
      else
        next.at( F_BOX + feat, childSlot ) = cut;  // right cell: low = cutval
    }
    sv.at( F_CHILD0 + c, s ) = childSlot;
  }
}

// step 4b: tight split bounds (divlow = max of the left child along feat, divhigh = min of the right child)
// and the slot of every position for the next level.
__global__ void kDivsAndAssign( SlotView sv, const short4* __restrict__ pts, const uint32_t* __restrict__ vind,
                                const int* __restrict__ slotOf, int* __restrict__ slotOfNext, int n ) {
  int        p   = blockIdx.x * blockDim.x + threadIdx.x;
  const int  s   = p < n ? slotOf[p] : -1;
  unsigned   act = __ballot_sync( 0xffffffffu, s >= 0 );
  if ( p < n && s < 0 ) slotOfNext[p] = -1;
  if ( s < 0 ) return;
  const int  lo = sv.at( F_LO, s ), idx = sv.at( F_IDX, s );
  const int  side = ( p - lo ) < idx ? 0 : 1;
  const int  v    = coord( pts[vind[p]], sv.at( F_FEAT, s ) );
  slotOfNext[p]   = sv.at( F_CHILD0 + side, s );
  const int  key    = s * 2 + side;
  const int  leader = __ffs( act ) - 1;
  const int  k0     = __shfl_sync( act, key, leader );
  const bool same   = __all_sync( act, key == k0 );
  if ( same ) {
    int r = side == 0 ? __reduce_max_sync( act, v ) : __reduce_min_sync( act, v );
    if ( ( threadIdx.x & 31 ) == leader ) {
      if ( side == 0 )
        relaxMax( &sv.at( F_DIVLOW, s ), r );
      else
        relaxMin( &sv.at( F_DIVHIGH, s ), r );
    }
  } else {
    if ( side == 0 )
      relaxMax( &sv.at( F_DIVLOW, s ), v );
    else
      relaxMin( &sv.at( F_DIVHIGH, s ), v );
  }
}

__global__ void kWriteInternal( SlotView sv, int numSlots, int nodeBase, int4* __restrict__ nodes ) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if ( s >= numSlots ) return;
  nodes[sv.at( F_NODE, s )] = make_int4( nodeBase + 2 * s, sv.at( F_FEAT, s ), sv.at( F_DIVLOW, s ), sv.at( F_DIVHIGH, s ) );
}

__global__ void kGatherPts( const short4* __restrict__ pts, const uint32_t* __restrict__ vind, short4* __restrict__ ptsT, int n ) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if ( p < n ) ptsT[p] = pts[vind[p]];
}

// ------------------------------------------------------------------------------------------- k-NN search
constexpr int kMaxStack = 96;

struct Box6 {
  int v[6];
};

template <int K>
__global__ void __launch_bounds__( 128 )
    kKnn( const int4* __restrict__ nodes, const short4* __restrict__ ptsT, const uint32_t* __restrict__ vind, Box6 rootBox,
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
      if ( qf[d] < float( rootBox.v[d] ) ) {
        float e = qf[d] - float( rootBox.v[d] );
        side[d] = e * e;
        mind += side[d];
      }
      if ( qf[d] > float( rootBox.v[3 + d] ) ) {
        float e = qf[d] - float( rootBox.v[3 + d] );
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
  const int N        = int( n );
  const int maxSlots = N / ( kLeafMaxSize + 1 ) + 2;
  t.pts.reserve( n );
  PCC_CUDA( cudaMemcpyAsync( t.pts, xyz4, n * sizeof( short4 ), cudaMemcpyDeviceToDevice, s ) );
  t.ptsT.reserve( n );
  t.vind.reserve( n );
  t.nodes.reserve( 2 * n + 2 );
  t.tmpA.reserve( n ), t.tmpB.reserve( n ), t.flags.reserve( n ), t.scanOut.reserve( n + 1 );
  t.scanTmp.reserve( scanTmpElems( n ) );
  t.slotOf.reserve( n ), t.slotOfNext.reserve( n );
  t.slotI[0].reserve( size_t( F_COUNT ) * maxSlots ), t.slotI[1].reserve( size_t( F_COUNT ) * maxSlots );
  t.counters.reserve( 16 );
  t.hostInts.reserve( 16 );
  const int TB = 256, gridN = divUp( n, TB );

  if ( N <= kLeafMaxSize ) {  // the root is a leaf
    kIota<<<gridN, TB, 0, s>>>( t.vind, t.slotOf, N, 0 );
    SlotView sv{ t.slotI[0], maxSlots };
    kResetMinMax<<<1, 32, 0, s>>>( sv, 1 );
    kMinMax<<<gridN, TB, 0, s>>>( sv, t.pts, t.vind, t.slotOf, N );
    kRootBox<<<1, 1, 0, s>>>( sv, t.counters.p + 8 );
    int4 leaf = make_int4( 0, N | kLeafBit, 0, 0 );
    PCC_CUDA( cudaMemcpyAsync( t.nodes, &leaf, sizeof( leaf ), cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( t.rootBox, t.counters.p + 8, 6 * sizeof( int ), cudaMemcpyDeviceToHost, s ) );
    kGatherPts<<<gridN, TB, 0, s>>>( t.pts, t.vind, t.ptsT, N );
    PCC_LAUNCH_CHECK();
    streamWait( s );
    t.numNodes = 1;
    return;
  }

  int cur = 0;
  {
    SlotView sv{ t.slotI[cur], maxSlots };
    kIota<<<gridN, TB, 0, s>>>( t.vind, t.slotOf, N, 0 );
    kInitRoot<<<1, 1, 0, s>>>( sv, N );
    kResetMinMax<<<1, 32, 0, s>>>( sv, 1 );
    kMinMax<<<gridN, TB, 0, s>>>( sv, t.pts, t.vind, t.slotOf, N );
    kRootBox<<<1, 1, 0, s>>>( sv, t.counters.p + 8 );
    PCC_CUDA( cudaMemcpyAsync( t.rootBox, t.counters.p + 8, 6 * sizeof( int ), cudaMemcpyDeviceToHost, s ) );
    PCC_LAUNCH_CHECK();
  }
  int numSlots = 1, nodeCount = 1;
  int* slotOf = t.slotOf;
  int* slotOfNext = t.slotOfNext;
  bool first = true;
  while ( numSlots > 0 ) {
    SlotView  sv{ t.slotI[cur], maxSlots }, nx{ t.slotI[cur ^ 1], maxSlots };
    const int gridS = divUp( numSlots, 128 );
    if ( !first ) {
      kResetMinMax<<<gridS, 128, 0, s>>>( sv, numSlots );
      kMinMax<<<gridN, TB, 0, s>>>( sv, t.pts, t.vind, slotOf, N );
    }
    first = false;
    kDecide<<<gridS, 128, 0, s>>>( sv, numSlots );
    // sweep 1
    kFlags<1><<<gridN, TB, 0, s>>>( sv, t.pts, t.vind, slotOf, t.flags, N );
    exclusiveScanU32( t.flags, t.scanOut, n, t.scanTmp, s );
    kCompact<<<gridN, TB, 0, s>>>( sv, t.vind, slotOf, t.flags, t.scanOut, t.tmpA, t.tmpB, N );
    kApply<1><<<gridN, TB, 0, s>>>( sv, t.vind, slotOf, t.flags, t.scanOut, t.tmpA, t.tmpB, N );
    kAdvanceSubLo<<<gridS, 128, 0, s>>>( sv, numSlots );
    // sweep 2
    kFlags<2><<<gridN, TB, 0, s>>>( sv, t.pts, t.vind, slotOf, t.flags, N );
    exclusiveScanU32( t.flags, t.scanOut, n, t.scanTmp, s );
    kCompact<<<gridN, TB, 0, s>>>( sv, t.vind, slotOf, t.flags, t.scanOut, t.tmpA, t.tmpB, N );
    kApply<2><<<gridN, TB, 0, s>>>( sv, t.vind, slotOf, t.flags, t.scanOut, t.tmpA, t.tmpB, N );
    // children
    PCC_CUDA( cudaMemsetAsync( t.counters, 0, sizeof( int ), s ) );
    kChildren<<<gridS, 128, 0, s>>>( sv, nx, numSlots, nodeCount, t.nodes, t.counters );
    kDivsAndAssign<<<gridN, TB, 0, s>>>( sv, t.pts, t.vind, slotOf, slotOfNext, N );
    kWriteInternal<<<gridS, 128, 0, s>>>( sv, numSlots, nodeCount, t.nodes );
    PCC_LAUNCH_CHECK();
    PCC_CUDA( cudaMemcpyAsync( t.hostInts.p, t.counters, sizeof( int ), cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    const int nextSlots = t.hostInts.p[0];
    nodeCount += 2 * numSlots;
    numSlots = nextSlots;
    cur ^= 1;
    std::swap( slotOf, slotOfNext );
    ++t.numLevels;
  }
  t.numNodes = nodeCount;
  kGatherPts<<<gridN, TB, 0, s>>>( t.pts, t.vind, t.ptsT, N );
  PCC_LAUNCH_CHECK();
}

void kdKnn( const KdTree& t, const short4* queries, size_t nq, const uint32_t* queryOrder, int k, uint32_t* outIdx,
            float* outDist, cudaStream_t s ) {
  if ( nq == 0 ) return;
  // the traversal stack holds one far child per level: a deeper tree (degenerate input) would overflow it silently
  if ( t.numLevels + 2 > kMaxStack ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
  Box6 rb;
  for ( int d = 0; d < 6; ++d ) rb.v[d] = t.rootBox[d];
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
