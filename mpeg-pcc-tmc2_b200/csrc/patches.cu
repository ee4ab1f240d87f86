// patches.cu — patch segmentation: connected components, projection (D0/D1), occupancy, residual loop.
// Replaces PCCPatchSegmenter3::segmentPatches (PccLibEncoder/source/PCCPatchSegmenter.cpp:537-1320, CTC path)
// and resampledPointcloud (:362-470).
//
// Exact data-parallel forms used here:
//  * The reference flood-fills the DIRECTED k-NN graph from seeds in ascending point index, removing each found
//    set before the next seed. A forward-reachable set is closed under out-edges, so every point's component is
//    the smallest seed index that reaches it. We contract mutually-linked points (both directions present: the
//    bulk of a surface) with a lock-free union-find — all members of such a set share one reachable set — and
//    then propagate the minimum seed along the remaining one-way edges between contracted sets to a fixed point.
//  * depth0 is a strict min (max) per pixel: one 64-bit atomicMin of (depth<<32 | point) per point gives both
//    the depth and the owning point whose colour gates D1.  D1, the per-block peak filter, occupancy and the
//    resampled cloud are per-pixel / per-point and order-free.
//  * The residual test only needs min squared distance to the resampled cloud up to the detection threshold (9):
//    resampled points are kept as a 3-D bitmap over the cloud's bounding box and probed within radius 3.
#include <limits.h>

#include <algorithm>
#include <cmath>

#include "stages.cuh"

namespace pccb200 {

namespace {

constexpr uint32_t kInf      = 0xFFFFFFFFu;
constexpr int16_t  kInfDepth = 32767;

// view id -> normal, tangent, bitangent axis, projection mode (PCCPatch::setViewId, PCCPatch.cpp:111-138)
__constant__ int cViewAxes[6][4] = {{0, 2, 1, 0}, {1, 2, 0, 0}, {2, 0, 1, 0}, {0, 2, 1, 1}, {1, 2, 0, 1}, {2, 0, 1, 1}};
const int        hViewAxes[6][4] = {{0, 2, 1, 0}, {1, 2, 0, 0}, {2, 0, 1, 0}, {0, 2, 1, 1}, {1, 2, 0, 1}, {2, 0, 1, 1}};

__device__ __forceinline__ int axisOf( const short4& p, int a ) { return a == 0 ? p.x : ( a == 1 ? p.y : p.z ); }

// device-side description of the patches created in the current outer iteration
struct DevPatch {
  int       viewId;
  int       u1, v1, d1, sizeU, sizeV, sizeU0, sizeV0;
  long long pixOff;  // into the per-iteration pixel arrays
  long long blkOff;  // into the per-iteration block arrays
};

// bit t of mutual[i] is set when i appears in the neighbour list of nbr[i][t]
__global__ void __launch_bounds__( 128 ) kMutual( const uint32_t* __restrict__ nbr, int n, int k, uint16_t* __restrict__ mutual ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  uint32_t m = 0;
  for ( int t = 0; t < k; ++t ) {
    const uint32_t j = nbr[size_t( i ) * k + t];
    if ( j == kInf ) break;
    const uint4* row = reinterpret_cast<const uint4*>( nbr + size_t( j ) * k );
    bool         hit = false;
    if ( k == 16 ) {
#pragma unroll
      for ( int q = 0; q < 4; ++q ) {
        const uint4 r = row[q];
        hit |= r.x == uint32_t( i ) || r.y == uint32_t( i ) || r.z == uint32_t( i ) || r.w == uint32_t( i );
      }
    } else {
      for ( int q = 0; q < k; ++q ) hit |= nbr[size_t( j ) * k + q] == uint32_t( i );
    }
    if ( hit ) m |= 1u << t;
  }
  mutual[i] = uint16_t( m );
}

__global__ void kIterInit( const uint8_t* __restrict__ raw, const uint8_t* __restrict__ minD2, int n, int detectThreshold,
                           uint32_t* __restrict__ parent, uint32_t* __restrict__ compLabel, uint32_t* __restrict__ compSize ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  parent[i]    = i;
  compSize[i]  = 0;
  compLabel[i] = ( raw[i] && int( minD2[i] ) > detectThreshold ) ? uint32_t( i ) : kInf;
}

__device__ __forceinline__ uint32_t ufFind( uint32_t* parent, uint32_t x ) {
  uint32_t p = parent[x];
  while ( p != x ) {
    const uint32_t gp = parent[p];
    if ( gp != p ) parent[x] = gp;  // path halving (benign race: always points to an ancestor)
    x = p;
    p = gp;
  }
  return x;
}

// union of mutually-linked, same-orientation, still-raw points; roots are hooked larger-under-smaller
__global__ void __launch_bounds__( 128 )
    kHook( const uint32_t* __restrict__ nbr, const uint16_t* __restrict__ mutual, const uint8_t* __restrict__ raw,
           const uint8_t* __restrict__ partition, int n, int k, uint32_t* parent ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n || !raw[i] ) return;
  const uint32_t m  = mutual[i];
  const uint8_t  pi = partition[i];
  for ( int t = 0; t < k; ++t ) {
    if ( !( m >> t & 1u ) ) continue;
    const uint32_t j = nbr[size_t( i ) * k + t];
    if ( j <= uint32_t( i ) || !raw[j] || partition[j] != pi ) continue;  // each mutual pair once
    uint32_t a = uint32_t( i ), b = j;
    for ( ;; ) {
      a = ufFind( parent, a ), b = ufFind( parent, b );
      if ( a == b ) break;
      if ( a < b ) {
        const uint32_t t2 = a;
        a                 = b;
        b                 = t2;
      }
      if ( atomicCAS( &parent[a], a, b ) == a ) break;  // a (larger root) now hangs under b
    }
  }
}

__global__ void kCompressAndSeed( uint32_t* parent, const uint8_t* __restrict__ raw, int n, uint32_t* compLabel ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n || !raw[i] ) return;
  const uint32_t r = ufFind( parent, uint32_t( i ) );
  parent[i]        = r;
  const uint32_t l = compLabel[i];
  if ( l != kInf && r != uint32_t( i ) ) atomicMin( &compLabel[r], l );
}

// one round of min-seed propagation along directed edges between different contracted sets
__global__ void __launch_bounds__( 128 )
    kPropagate( const uint32_t* __restrict__ nbr, const uint8_t* __restrict__ raw, const uint8_t* __restrict__ partition,
                const uint32_t* __restrict__ parent, int n, int k, uint32_t* compLabel, int* changed ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n || !raw[i] ) return;
  const uint32_t ri = parent[i];
  const uint32_t li = compLabel[ri];
  if ( li == kInf ) return;
  const uint8_t pi = partition[i];
  bool          ch = false;
  for ( int t = 0; t < k; ++t ) {
    const uint32_t j = nbr[size_t( i ) * k + t];
    if ( j == kInf ) break;
    if ( !raw[j] || partition[j] != pi ) continue;
    const uint32_t rj = parent[j];
    if ( rj != ri && compLabel[rj] > li ) {
      if ( atomicMin( &compLabel[rj], li ) > li ) ch = true;
    }
  }
  if ( ch ) *changed = 1;
}

// The same propagation on a compact list of the edges that can still carry a label: after the contraction only the one-way
// links between DIFFERENT sets matter (about a tenth of the 16 n neighbour slots). Built once per outer iteration, then every
// round is a pass over that list instead of over all neighbour rows.
__global__ void __launch_bounds__( 128 )
    kCrossEdges( const uint32_t* __restrict__ nbr, const uint8_t* __restrict__ raw, const uint8_t* __restrict__ partition,
                 const uint32_t* __restrict__ parent, int n, int k, uint2* __restrict__ edges, unsigned* __restrict__ count, unsigned capacity ) {
  const int i    = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t  to[16];
  int       c = 0;
  uint32_t  ri = 0;
  if ( i < n && raw[i] ) {
    ri               = parent[i];
    const uint8_t pi = partition[i];
    for ( int t = 0; t < k && t < 16; ++t ) {
      const uint32_t j = nbr[size_t( i ) * k + t];
      if ( j == kInf ) break;
      if ( !raw[j] || partition[j] != pi ) continue;
      const uint32_t rj = parent[j];
      if ( rj != ri ) to[c++] = rj;
    }
  }
  // warp-aggregated reservation
  int inc = c;
#pragma unroll
  for ( int o = 1; o < 32; o <<= 1 ) {
    const int t = __shfl_up_sync( 0xffffffffu, inc, o );
    if ( lane >= o ) inc += t;
  }
  const int total = __shfl_sync( 0xffffffffu, inc, 31 );
  unsigned  base  = 0;
  if ( lane == 31 && total ) base = atomicAdd( count, unsigned( total ) );
  base = __shfl_sync( 0xffffffffu, base, 31 ) + unsigned( inc - c );
  for ( int e = 0; e < c; ++e )
    if ( base + e < capacity ) edges[base + e] = make_uint2( ri, to[e] );
}
// `rounds` relaxation passes per launch (asynchronous updates are fine: min-label propagation is monotone, every schedule reaches
// the same fixed point). *changed = 1 when a label moved, 2 when the list did not fit (the caller falls back to kPropagate).
__global__ void __launch_bounds__( 256 )
    kPropagateEdges( const uint2* __restrict__ edges, const unsigned* __restrict__ count, unsigned capacity, uint32_t* compLabel, int* changed, int rounds ) {
  const unsigned E = *count;
  if ( E > capacity ) {
    if ( blockIdx.x == 0 && threadIdx.x == 0 ) *changed = 2;
    return;
  }
  bool ch = false;
  for ( int r = 0; r < rounds; ++r )
    for ( unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x ) {
      const uint2    ed = edges[e];
      const uint32_t li = *reinterpret_cast<volatile uint32_t*>( &compLabel[ed.x] );
      if ( li != kInf && *reinterpret_cast<volatile uint32_t*>( &compLabel[ed.y] ) > li && atomicMin( &compLabel[ed.y], li ) > li ) ch = true;
    }
  if ( ch ) *changed = 1;
}

__global__ void kLabelAndCount( const uint8_t* __restrict__ raw, const uint32_t* __restrict__ parent, const uint32_t* __restrict__ compLabel,
                                int n, uint32_t* __restrict__ label, uint32_t* compSize ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  uint32_t l = kInf;
  if ( raw[i] ) l = compLabel[parent[i]];
  label[i] = l;
  if ( l != kInf ) atomicAdd( &compSize[l], 1u );
}

__global__ void kKeptFlags( const uint32_t* __restrict__ compSize, int n, uint32_t minPoints, uint32_t* __restrict__ kept ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i < n ) kept[i] = compSize[i] >= minPoints ? 1u : 0u;
}

// member[i] = index (within this iteration) of the patch the point belongs to, or -1
struct PatchStats {  // per new patch
  int seed;
  int minU, minV;
  int bbMin[3], bbMax[3];
};

__global__ void kMembers( const uint32_t* __restrict__ label, const uint32_t* __restrict__ kept, const uint32_t* __restrict__ keptScan,
                          int n, int* __restrict__ member, PatchStats* __restrict__ stats ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  const uint32_t l = label[i];
  int            m = -1;
  if ( l != kInf && kept[l] ) m = int( keptScan[l] );
  member[i] = m;
  if ( m >= 0 && l == uint32_t( i ) ) {
    PatchStats st;
    st.seed = i, st.minU = st.minV = INT_MAX;
    st.bbMin[0] = st.bbMin[1] = st.bbMin[2] = INT_MAX;
    st.bbMax[0] = st.bbMax[1] = st.bbMax[2] = 0;  // the reference starts the max at 0
    stats[m] = st;
  }
}

// atomicMin / atomicMax on a per-patch extreme that ~10^4 points share: a plain (L2) read first - the stored value only ever
// moves towards the extreme, so a stale read can at worst cost a redundant atomic, never skip a needed one. Almost every point
// is inside the box the first few have spanned, which takes the serialised same-address atomics off the critical path.
__device__ __forceinline__ void relaxMin( int* addr, int v ) {
  if ( v < __ldcg( addr ) ) atomicMin( addr, v );
}
__device__ __forceinline__ void relaxMax( int* addr, int v ) {
  if ( v > __ldcg( addr ) ) atomicMax( addr, v );
}

// Points of a warp almost always belong to the same patch (the input order is spatially coherent): the warp reduces with
// redux.sync and one lane issues the atomics; mixed warps fall back to per-lane relaxed atomics.
__global__ void kMinUV( const short4* __restrict__ pts, const uint8_t* __restrict__ partition, const int* __restrict__ member, int n,
                        PatchStats* stats ) {
  const int      i   = blockIdx.x * blockDim.x + threadIdx.x;
  const int      m   = i < n ? member[i] : -1;
  const unsigned act = __ballot_sync( 0xffffffffu, m >= 0 );
  if ( m < 0 ) return;
  const int    view = partition[stats[m].seed];
  const short4 p    = pts[i];
  const int    u = axisOf( p, cViewAxes[view][1] ), v = axisOf( p, cViewAxes[view][2] );
  const int    leader = __ffs( act ) - 1;
  if ( __all_sync( act, m == __shfl_sync( act, m, leader ) ) ) {
    const int mu = __reduce_min_sync( act, u ), mv = __reduce_min_sync( act, v );
    if ( ( threadIdx.x & 31 ) == leader ) relaxMin( &stats[m].minU, mu ), relaxMin( &stats[m].minV, mv );
  } else {
    relaxMin( &stats[m].minU, u ), relaxMin( &stats[m].minV, v );
  }
}

__global__ void kSplitAndBounds( const short4* __restrict__ pts, const uint8_t* __restrict__ partition, int* __restrict__ member, int n,
                                 int maxPatchSize, int splitting, PatchStats* stats ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int       m = i < n ? member[i] : -1;
  short4    p = make_short4( 0, 0, 0, 0 );
  if ( m >= 0 ) {
    const int view = partition[stats[m].seed];
    p              = pts[i];
    if ( splitting ) {
      const int u = axisOf( p, cViewAxes[view][1] ), v = axisOf( p, cViewAxes[view][2] );
      if ( !( u - stats[m].minU < maxPatchSize && v - stats[m].minV < maxPatchSize ) ) {
        member[i] = -1;  // consumed by the component, not part of the patch
        m         = -1;
      }
    }
  }
  const unsigned act = __ballot_sync( 0xffffffffu, m >= 0 );
  if ( m < 0 ) return;
  const int c[3]   = {p.x, p.y, p.z};
  const int leader = __ffs( act ) - 1;
  if ( __all_sync( act, m == __shfl_sync( act, m, leader ) ) ) {
    int lo[3], hi[3];
#pragma unroll
    for ( int d = 0; d < 3; ++d ) lo[d] = __reduce_min_sync( act, c[d] ), hi[d] = __reduce_max_sync( act, c[d] );
    if ( ( threadIdx.x & 31 ) == leader ) {
#pragma unroll
      for ( int d = 0; d < 3; ++d ) relaxMin( &stats[m].bbMin[d], lo[d] ), relaxMax( &stats[m].bbMax[d], hi[d] );
    }
  } else {
#pragma unroll
    for ( int d = 0; d < 3; ++d ) relaxMin( &stats[m].bbMin[d], c[d] ), relaxMax( &stats[m].bbMax[d], c[d] );
  }
}

__global__ void kFillU64( unsigned long long* p, size_t n, unsigned long long v ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i < n ) p[i] = v;
}

// depth0: nearest (mode 0) / farthest (mode 1) point per pixel, with its index
__global__ void kDepth0( const short4* __restrict__ pts, const int* __restrict__ member, int n, const DevPatch* __restrict__ patches,
                         unsigned long long* __restrict__ keys ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  const int m = member[i];
  if ( m < 0 ) return;
  const DevPatch P    = patches[m];
  const short4   p    = pts[i];
  const int      d    = axisOf( p, cViewAxes[P.viewId][0] );
  const int      u    = axisOf( p, cViewAxes[P.viewId][1] ) - P.u1, v = axisOf( p, cViewAxes[P.viewId][2] ) - P.v1;
  const int      mode = cViewAxes[P.viewId][3];
  const unsigned long long key = ( (unsigned long long)( mode == 0 ? d : 0x7FFF - d ) << 32 ) | uint32_t( i );
  atomicMin( &keys[P.pixOff + (long long)v * P.sizeU + u], key );
}

// per 16x16 block: the peak depth0 (min for mode 0, max for mode 1)
__global__ void kPeak( const DevPatch* __restrict__ patches, const unsigned long long* __restrict__ keys, int occRes, int* __restrict__ peak ) {
  const DevPatch P  = patches[blockIdx.y];
  const int      px = P.sizeU * P.sizeV;
  const int      mode = cViewAxes[P.viewId][3];
  for ( int p = blockIdx.x * blockDim.x + threadIdx.x; p < px; p += gridDim.x * blockDim.x ) {
    const unsigned long long key = keys[P.pixOff + p];
    if ( key == ~0ull ) continue;
    const int hi = int( key >> 32 );
    const int d  = mode == 0 ? hi : 0x7FFF - hi;
    const int u = p % P.sizeU, v = p / P.sizeU;
    int*      pk = &peak[P.blkOff + ( v / occRes ) * P.sizeU0 + u / occRes];
    if ( mode == 0 )
      atomicMin( pk, d );
    else
      atomicMax( pk, d );
  }
}

// filter depth0 against the block peak and the coding range; write depth0/owner, initialise depth1
__global__ void kFilter( const DevPatch* __restrict__ patches, const unsigned long long* __restrict__ keys, const int* __restrict__ peak,
                         int occRes, int thickness, int maxAllowedDepth, int* __restrict__ d0, int* __restrict__ d1,
                         uint32_t* __restrict__ owner ) {
  const DevPatch P    = patches[blockIdx.y];
  const int      px   = P.sizeU * P.sizeV;
  const int      mode = cViewAxes[P.viewId][3], dir = 1 - 2 * mode;
  for ( int p = blockIdx.x * blockDim.x + threadIdx.x; p < px; p += gridDim.x * blockDim.x ) {
    const unsigned long long key = keys[P.pixOff + p];
    int                      d   = kInfDepth;
    uint32_t                 own = kInf;
    if ( key != ~0ull ) {
      const int hi = int( key >> 32 );
      d            = mode == 0 ? hi : 0x7FFF - hi;
      own          = uint32_t( key );
      const int u = p % P.sizeU, v = p / P.sizeU;
      const int pk = peak[P.blkOff + ( v / occRes ) * P.sizeU0 + u / occRes];
      const short a = short( abs( d - pk ) ), b = short( thickness + dir * d ), c = short( dir * P.d1 + maxAllowedDepth );
      if ( a > 32 || b > c ) d = kInfDepth, own = kInf;
    }
    d0[P.pixOff + p] = d, d1[P.pixOff + p] = d, owner[P.pixOff + p] = own;
  }
}

// depth1: farthest point within surfaceThickness of depth0 whose colour is close to the depth0 point's colour
__global__ void kDepth1( const short4* __restrict__ pts, const uchar4* __restrict__ rgb, const int* __restrict__ member, int n,
                         const DevPatch* __restrict__ patches, const int* __restrict__ d0, const uint32_t* __restrict__ owner,
                         int thickness, int* d1 ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  const int m = member[i];
  if ( m < 0 ) return;
  const DevPatch  P = patches[m];
  const short4    p = pts[i];
  const int       d = axisOf( p, cViewAxes[P.viewId][0] );
  const int       u = axisOf( p, cViewAxes[P.viewId][1] ) - P.u1, v = axisOf( p, cViewAxes[P.viewId][2] ) - P.v1;
  const long long q = P.pixOff + (long long)v * P.sizeU + u;
  const int       depth0 = d0[q];
  if ( !( depth0 < kInfDepth ) ) return;
  const int   dir   = 1 - 2 * cViewAxes[P.viewId][3];
  const short delta = short( dir * ( d - depth0 ) );
  if ( !( delta <= thickness && delta >= 0 ) ) return;
  const uchar4 ci = rgb[i], c0 = rgb[owner[q]];
  if ( !( abs( int( c0.x ) - ci.x ) < 128 && abs( int( c0.y ) - ci.y ) < 128 && abs( int( c0.z ) - ci.z ) < 128 ) ) return;
  if ( dir > 0 )
    atomicMax( &d1[q], d );
  else
    atomicMin( &d1[q], d );
}

struct Bitmap3 {
  uint32_t* bits;
  int       mn[3], dim[3];
  int       wordsX;
};

__device__ __forceinline__ void bitmapSet( const Bitmap3& b, int x, int y, int z ) {
  x -= b.mn[0], y -= b.mn[1], z -= b.mn[2];
  if ( x < 0 || y < 0 || z < 0 || x >= b.dim[0] || y >= b.dim[1] || z >= b.dim[2] ) return;
  atomicOr( &b.bits[( size_t( z ) * b.dim[1] + y ) * b.wordsX + ( x >> 5 )], 1u << ( x & 31 ) );
}
__device__ __forceinline__ bool bitmapGet( const Bitmap3& b, int x, int y, int z ) {
  x -= b.mn[0], y -= b.mn[1], z -= b.mn[2];
  if ( x < 0 || y < 0 || z < 0 || x >= b.dim[0] || y >= b.dim[1] || z >= b.dim[2] ) return false;
  return ( b.bits[( size_t( z ) * b.dim[1] + y ) * b.wordsX + ( x >> 5 )] >> ( x & 31 ) ) & 1u;
}

struct PatchCounters {
  int d0Count, sizeD;
};

// occupancy, re-based int16 depth maps, counters, resampled-cloud bitmap (resampledPointcloud, :362-470)
__global__ void kFinalize( const DevPatch* __restrict__ patches, const int* __restrict__ d0, const int* __restrict__ d1, int occRes,
                           int16_t* __restrict__ depthOut, uint8_t* __restrict__ occOut, const long long* __restrict__ depthOutOff,
                           const long long* __restrict__ occOutOff, PatchCounters* counters, Bitmap3 bm ) {
  const DevPatch  P    = patches[blockIdx.y];
  const int       px   = P.sizeU * P.sizeV;
  const int       na = cViewAxes[P.viewId][0], ta = cViewAxes[P.viewId][1], ba = cViewAxes[P.viewId][2];
  const int       dir  = 1 - 2 * cViewAxes[P.viewId][3];
  const long long dOff = depthOutOff[blockIdx.y], oOff = occOutOff[blockIdx.y];
  int             cnt = 0, sizeD = 0;
  for ( int p = blockIdx.x * blockDim.x + threadIdx.x; p < px; p += gridDim.x * blockDim.x ) {
    const int a = d0[P.pixOff + p], b = d1[P.pixOff + p];
    if ( a < kInfDepth ) {
      const int u = p % P.sizeU, v = p / P.sizeU;
      occOut[oOff + ( v / occRes ) * P.sizeU0 + u / occRes] = 1;
      int q[3];
      q[ta] = u + P.u1, q[ba] = v + P.v1;
      q[na] = a;
      bitmapSet( bm, q[0], q[1], q[2] );
      q[na] = b;
      bitmapSet( bm, q[0], q[1], q[2] );
      const int ra = dir * ( a - P.d1 ), rb = dir * ( b - P.d1 );
      depthOut[dOff + p]      = int16_t( ra );
      depthOut[dOff + px + p] = int16_t( rb );
      ++cnt;
      sizeD = max( sizeD, max( ra, rb ) );
    } else {
      depthOut[dOff + p]      = kInfDepth;
      depthOut[dOff + px + p] = kInfDepth;
    }
  }
  if ( cnt ) atomicAdd( &counters[blockIdx.y].d0Count, cnt );
  if ( sizeD ) atomicMax( &counters[blockIdx.y].sizeD, sizeD );
}

// residual: squared distance of every still-raw point to the resampled cloud, exact up to radius^2
__global__ void __launch_bounds__( 128 )
    kResidual( const short4* __restrict__ pts, int n, Bitmap3 bm, int radius, int selectThreshold, uint8_t* __restrict__ raw,
               uint8_t* __restrict__ minD2, unsigned* rawCount ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool      still = false;
  if ( i < n && raw[i] ) {
    const short4 p = pts[i];
    int          best = 255;
    for ( int dz = -radius; dz <= radius; ++dz )
      for ( int dy = -radius; dy <= radius; ++dy ) {
        const int rem = radius * radius - dz * dz - dy * dy;
        if ( rem < 0 ) continue;
        for ( int dx = -radius; dx <= radius; ++dx ) {
          const int d2 = dx * dx + dy * dy + dz * dz;
          if ( dx * dx > rem || d2 >= best ) continue;
          if ( bitmapGet( bm, p.x + dx, p.y + dy, p.z + dz ) ) best = d2;
        }
      }
    minD2[i] = uint8_t( best );
    still    = best > selectThreshold;
    raw[i]   = still ? 1 : 0;
  }
  const unsigned m = __ballot_sync( 0xffffffffu, still );
  if ( ( threadIdx.x & 31 ) == 0 && m ) atomicAdd( rawCount, __popc( m ) );
}

__global__ void kPackRgb( const uint8_t* __restrict__ in, int n, uchar4* __restrict__ out ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i < n ) out[i] = make_uchar4( in[3 * size_t( i )], in[3 * size_t( i ) + 1], in[3 * size_t( i ) + 2], 0 );
}

__global__ void kBounds( const short4* __restrict__ pts, int n, int* __restrict__ mm ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int       lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  if ( i < n ) {
    const short4 c = pts[i];
    lo[0] = hi[0] = c.x, lo[1] = hi[1] = c.y, lo[2] = hi[2] = c.z;
  }
#pragma unroll
  for ( int d = 0; d < 3; ++d ) {
    const int a = __reduce_min_sync( 0xffffffffu, lo[d] ), b = __reduce_max_sync( 0xffffffffu, hi[d] );
    if ( ( threadIdx.x & 31 ) == 0 ) atomicMin( &mm[d], a ), atomicMax( &mm[3 + d], b );
  }
}

}  // namespace

void packRgb( const uint8_t* rgb3, size_t n, uchar4* out, cudaStream_t s ) {
  if ( n == 0 ) return;
  kPackRgb<<<divUp( n, 256 ), 256, 0, s>>>( rgb3, int( n ), out );
  PCC_LAUNCH_CHECK();
}

void segmentPatches( PatchScratch& sc, PatchResult& out, const short4* pts, const uchar4* rgb, const uint32_t* nbr, int k,
                     const uint8_t* partition, size_t n, const pccb200_seg_params& prm, cudaStream_t s ) {
  out.patches.clear();
  out.depthElems = out.occElems = 0;
  out.outerIterations           = 0;
  if ( n == 0 ) return;
  const int N = int( n ), TB = 256, gridN = divUp( n, TB );
  const int detect = int( std::floor( prm.max_allowed_dist2_raw_detection ) ), select = int( std::floor( prm.max_allowed_dist2_raw_selection ) );
  int       radius = 0;
  while ( ( radius + 1 ) * ( radius + 1 ) <= std::max( detect, select ) ) ++radius;
  if ( radius > 6 || detect > 200 || prm.occupancy_resolution <= 0 ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };

  sc.raw.reserve( n ), sc.minD2.reserve( n ), sc.mutual.reserve( n ), sc.parent.reserve( n ), sc.compLabel.reserve( n );
  sc.compSize.reserve( n ), sc.label.reserve( n ), sc.kept.reserve( n + 1 ), sc.keptScan.reserve( n + 1 ), sc.member.reserve( n );
  sc.scanTmp.reserve( scanTmpElems( n ) ), sc.ints.reserve( 16 ), sc.keys.reserve( 4 * n );
  PCC_CUDA( cudaMemsetAsync( sc.raw, 1, n, s ) );
  PCC_CUDA( cudaMemsetAsync( sc.minD2, 255, n, s ) );
  kMutual<<<divUp( n, 128 ), 128, 0, s>>>( nbr, N, k, sc.mutual );
  // resampled-cloud bitmap over the cloud's bounding box
  int init[8] = {0, 0, INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
  PCC_CUDA( cudaMemcpyAsync( sc.ints, init, sizeof( init ), cudaMemcpyHostToDevice, s ) );
  kBounds<<<gridN, TB, 0, s>>>( pts, N, sc.ints.p + 2 );
  int mm[6];
  PCC_CUDA( cudaMemcpyAsync( mm, sc.ints.p + 2, sizeof( mm ), cudaMemcpyDeviceToHost, s ) );
  streamWait( s );
  Bitmap3 bm;
  for ( int d = 0; d < 3; ++d ) bm.mn[d] = mm[d], bm.dim[d] = mm[3 + d] - mm[d] + 1;
  bm.wordsX             = ( bm.dim[0] + 31 ) / 32;
  const size_t bmWords  = size_t( bm.wordsX ) * bm.dim[1] * bm.dim[2];
  sc.bitmap.reserve( bmWords );
  PCC_CUDA( cudaMemsetAsync( sc.bitmap, 0, bmWords * sizeof( uint32_t ), s ) );
  bm.bits = sc.bitmap;

  const int occRes = prm.occupancy_resolution, minLevel = prm.min_level;
  for ( ;; ) {
    // ---- connected components (exact directed reachability, see file header)
    kIterInit<<<gridN, TB, 0, s>>>( sc.raw, sc.minD2, N, detect, sc.parent, sc.compLabel, sc.compSize );
    kHook<<<divUp( n, 128 ), 128, 0, s>>>( nbr, sc.mutual, sc.raw, partition, N, k, sc.parent );
    kCompressAndSeed<<<gridN, TB, 0, s>>>( sc.parent, sc.raw, N, sc.compLabel );
    {
      const unsigned capacity = unsigned( 4 * n );
      sc.keys.reserve( capacity );  // (the projection's pixel keys are not alive yet: 8 bytes per entry, as an edge)
      uint2* const edges = reinterpret_cast<uint2*>( sc.keys.p );
      PCC_CUDA( cudaMemsetAsync( sc.ints.p + 8, 0, sizeof( int ), s ) );
      kCrossEdges<<<divUp( n, 128 ), 128, 0, s>>>( nbr, sc.raw, partition, sc.parent, N, k, edges, reinterpret_cast<unsigned*>( sc.ints.p + 8 ), capacity );
      bool fallback = false;
      for ( ;; ) {
        PCC_CUDA( cudaMemsetAsync( sc.ints, 0, sizeof( int ), s ) );
        if ( !fallback )
          kPropagateEdges<<<148 * 4, 256, 0, s>>>( edges, reinterpret_cast<unsigned*>( sc.ints.p + 8 ), capacity, sc.compLabel, sc.ints, 8 );
        else
          for ( int r = 0; r < 4; ++r ) kPropagate<<<divUp( n, 128 ), 128, 0, s>>>( nbr, sc.raw, partition, sc.parent, N, k, sc.compLabel, sc.ints );
        int changed = 0;
        PCC_CUDA( cudaMemcpyAsync( &changed, sc.ints, sizeof( int ), cudaMemcpyDeviceToHost, s ) );
        streamWait( s );
        if ( changed == 2 ) {
          fallback = true;
          continue;
        }
        if ( !changed ) break;
      }
    }
    kLabelAndCount<<<gridN, TB, 0, s>>>( sc.raw, sc.parent, sc.compLabel, N, sc.label, sc.compSize );
    kKeptFlags<<<gridN, TB, 0, s>>>( sc.compSize, N, uint32_t( prm.min_point_count_per_cc ), sc.kept );
    exclusiveScanU32( sc.kept, sc.keptScan, n, sc.scanTmp, s );
    uint32_t numNew = 0;
    PCC_CUDA( cudaMemcpyAsync( &numNew, sc.keptScan.p + n, sizeof( uint32_t ), cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    if ( numNew == 0 ) break;
    ++out.outerIterations;
    // ---- per-patch extents
    sc.stats.reserve( numNew * sizeof( PatchStats ) ), sc.devPatches.reserve( numNew * sizeof( DevPatch ) ), sc.counters.reserve( numNew * sizeof( PatchCounters ) );
    PatchStats*    dStats    = reinterpret_cast<PatchStats*>( sc.stats.p );
    DevPatch*      dPatches  = reinterpret_cast<DevPatch*>( sc.devPatches.p );
    PatchCounters* dCounters = reinterpret_cast<PatchCounters*>( sc.counters.p );
    sc.depthOff.reserve( numNew ), sc.occOff.reserve( numNew );
    kMembers<<<gridN, TB, 0, s>>>( sc.label, sc.kept, sc.keptScan, N, sc.member, dStats );
    if ( prm.enable_patch_splitting ) kMinUV<<<gridN, TB, 0, s>>>( pts, partition, sc.member, N, dStats );
    kSplitAndBounds<<<gridN, TB, 0, s>>>( pts, partition, sc.member, N, prm.max_patch_size, prm.enable_patch_splitting, dStats );
    std::vector<PatchStats> hStats( numNew );
    std::vector<uint8_t>    hView( numNew );
    PCC_CUDA( cudaMemcpyAsync( hStats.data(), dStats, numNew * sizeof( PatchStats ), cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    // view ids: partition[seed] — gather on the host side from a small device read
    sc.seedIdx.reserve( numNew ), sc.seedView.reserve( numNew );
    {
      std::vector<uint32_t> seeds( numNew );
      for ( uint32_t j = 0; j < numNew; ++j ) seeds[j] = uint32_t( hStats[j].seed );
      PCC_CUDA( cudaMemcpyAsync( sc.seedIdx, seeds.data(), numNew * sizeof( uint32_t ), cudaMemcpyHostToDevice, s ) );
      gatherU8( partition, sc.seedIdx, numNew, sc.seedView, s );
      PCC_CUDA( cudaMemcpyAsync( hView.data(), sc.seedView, numNew, cudaMemcpyDeviceToHost, s ) );
      streamWait( s );
    }
    std::vector<DevPatch>  hPatches( numNew );
    std::vector<long long> hDepthOff( numNew ), hOccOff( numNew );
    long long              pix = 0, blk = 0;
    const size_t           firstNew = out.patches.size();
    for ( uint32_t j = 0; j < numNew; ++j ) {
      const PatchStats& st = hStats[j];
      pccb200_patch     m{};
      m.best_match_idx = -1;
      m.index   = int32_t( out.patches.size() );
      m.view_id = hView[j];
      m.normal_axis = hViewAxes[m.view_id][0], m.tangent_axis = hViewAxes[m.view_id][1];
      m.bitangent_axis = hViewAxes[m.view_id][2], m.projection_mode = hViewAxes[m.view_id][3];
      DevPatch& P = hPatches[j];
      P.viewId    = m.view_id;
      if ( st.bbMin[0] == INT_MAX ) {
        // every point was split away: the reference keeps an empty patch in its list
        P.u1 = P.v1 = P.d1 = P.sizeU = P.sizeV = P.sizeU0 = P.sizeV0 = 0;
      } else {
        const int ta = m.tangent_axis, ba = m.bitangent_axis, na = m.normal_axis;
        m.size_u = 1 + st.bbMax[ta] - st.bbMin[ta], m.size_v = 1 + st.bbMax[ba] - st.bbMin[ba];
        m.u1 = st.bbMin[ta], m.v1 = st.bbMin[ba];
        const int maxU = st.bbMax[ta] - m.u1, maxV = st.bbMax[ba] - m.v1;
        const int ext  = m.projection_mode == 0 ? st.bbMin[na] : st.bbMax[na];
        m.d1 = m.projection_mode == 0 ? ( ext / minLevel ) * minLevel : int( std::ceil( double( ext ) / double( minLevel ) ) ) * minLevel;
        m.size_u0 = maxU / occRes + 1, m.size_v0 = maxV / occRes + 1;
        m.size_2d_x = int( std::ceil( double( maxU + 1 ) / double( prm.quantizer_size_x ) ) * prm.quantizer_size_x );
        m.size_2d_y = int( std::ceil( double( maxV + 1 ) / double( prm.quantizer_size_y ) ) * prm.quantizer_size_y );
        P.u1 = m.u1, P.v1 = m.v1, P.d1 = m.d1, P.sizeU = m.size_u, P.sizeV = m.size_v, P.sizeU0 = m.size_u0, P.sizeV0 = m.size_v0;
      }
      P.pixOff = pix, P.blkOff = blk;
      m.depth_offset = int64_t( out.depthElems ), m.occ_offset = int64_t( out.occElems );
      hDepthOff[j] = m.depth_offset, hOccOff[j] = m.occ_offset;
      pix += (long long)m.size_u * m.size_v, blk += (long long)m.size_u0 * m.size_v0;
      out.depthElems += 2 * size_t( m.size_u ) * m.size_v, out.occElems += size_t( m.size_u0 ) * m.size_v0;
      out.patches.push_back( m );
    }
    PCC_CUDA( cudaMemcpyAsync( dPatches, hPatches.data(), numNew * sizeof( DevPatch ), cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( sc.depthOff, hDepthOff.data(), numNew * sizeof( long long ), cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( sc.occOff, hOccOff.data(), numNew * sizeof( long long ), cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemsetAsync( dCounters, 0, numNew * sizeof( PatchCounters ), s ) );
    // ---- projection
    sc.keys.reserve( pix + 1 ), sc.d0.reserve( pix + 1 ), sc.d1.reserve( pix + 1 ), sc.owner.reserve( pix + 1 ), sc.peak.reserve( blk + 1 );
    out.depth.grow( out.depthElems + 1, s ), out.occ.grow( out.occElems + 1, s );
    kFillU64<<<divUp( pix + 1, 256 ), 256, 0, s>>>( sc.keys, size_t( pix ), ~0ull );
    // peak init: +inf for mode 0, 0 for mode 1 -> fill per patch on the host-built table
    {
      std::vector<int> hPeak( blk );
      for ( uint32_t j = 0; j < numNew; ++j ) {
        const int v = hViewAxes[hPatches[j].viewId][3] == 0 ? int( kInfDepth ) : 0;
        std::fill( hPeak.begin() + hPatches[j].blkOff, hPeak.begin() + hPatches[j].blkOff + (long long)hPatches[j].sizeU0 * hPatches[j].sizeV0, v );
      }
      if ( blk ) PCC_CUDA( cudaMemcpyAsync( sc.peak, hPeak.data(), blk * sizeof( int ), cudaMemcpyHostToDevice, s ) );
      streamWait( s );  // hPeak goes out of scope
    }
    PCC_CUDA( cudaMemsetAsync( out.occ.p + hOccOff[0], 0, out.occElems - size_t( hOccOff[0] ), s ) );
    kDepth0<<<gridN, TB, 0, s>>>( pts, sc.member, N, dPatches, sc.keys );
    int maxPix = 1;
    for ( uint32_t j = 0; j < numNew; ++j ) maxPix = std::max( maxPix, hPatches[j].sizeU * hPatches[j].sizeV );
    const dim3 pg( std::min( divUp( maxPix, 256 ), 64 ), numNew );
    kPeak<<<pg, 256, 0, s>>>( dPatches, sc.keys, occRes, sc.peak );
    kFilter<<<pg, 256, 0, s>>>( dPatches, sc.keys, sc.peak, occRes, prm.surface_thickness, prm.max_allowed_depth, sc.d0, sc.d1, sc.owner );
    if ( prm.surface_thickness > 0 )
      kDepth1<<<gridN, TB, 0, s>>>( pts, rgb, sc.member, N, dPatches, sc.d0, sc.owner, prm.surface_thickness, sc.d1 );
    kFinalize<<<pg, 256, 0, s>>>( dPatches, sc.d0, sc.d1, occRes, out.depth, out.occ, sc.depthOff, sc.occOff, dCounters, bm );
    // ---- residual
    PCC_CUDA( cudaMemsetAsync( sc.ints.p + 1, 0, sizeof( int ), s ) );
    kResidual<<<divUp( n, 128 ), 128, 0, s>>>( pts, N, bm, radius, select, sc.raw, sc.minD2, reinterpret_cast<unsigned*>( sc.ints.p + 1 ) );
    PCC_LAUNCH_CHECK();
    std::vector<PatchCounters> hCnt( numNew );
    unsigned                   rawCount = 0;
    PCC_CUDA( cudaMemcpyAsync( hCnt.data(), dCounters, numNew * sizeof( PatchCounters ), cudaMemcpyDeviceToHost, s ) );
    PCC_CUDA( cudaMemcpyAsync( &rawCount, sc.ints.p + 1, sizeof( unsigned ), cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    for ( uint32_t j = 0; j < numNew; ++j ) {
      pccb200_patch& m = out.patches[firstNew + j];
      m.d0_count       = hCnt[j].d0Count;
      int sizeD        = hCnt[j].sizeD;
      m.size_d_pixel   = sizeD;
      const int bd     = std::min( prm.geometry_bitdepth_3d, prm.geometry_bitdepth_2d );
      sizeD            = std::min( ( 1 << bd ) - 1, sizeD );
      const int lv     = int( std::log2( double( minLevel ) ) );
      int       qd     = sizeD == 0 ? 0 : ( ( sizeD - 1 ) / minLevel + 1 );
      qd               = std::min( qd, ( 1 << ( bd - lv ) ) - 1 );
      m.size_d         = qd == 0 ? 0 : qd * minLevel - 1;
    }
    if ( rawCount == 0 ) break;
  }
}

}  // namespace pccb200
