// refine.cu — grid-based refinement of the initial segmentation
// (PCCPatchSegmenter3::refineSegmentationGridBased, PccLibEncoder/source/PCCPatchSegmenter.cpp:1386-1561;
//  voxel classes AttributeOfGridCell / PointIndicesOfGridCell, PCCPatchSegmenter.h:430-510).
//
// Exact data-parallel form of the reference's sequential voxel sweep:
//  * voxels are numbered by first appearance in the point list            -> stable sort by key + sort by first index
//  * neighbour voxels = centres with dist^2 < R, ordered by (dist^2, index), cut when >= maxNN points are
//    gathered (nanoflann radiusSearch + IndexDist_Sorter is a total order) -> one warp per voxel probes a dense
//    centre grid with a distance-sorted offset table and sorts the hits in shared memory
//  * inside one sweep the only sequential coupling is the voxel class: voxel j is processed iff it is an edge
//    voxel at sweep start or some processed voxel i<j marks it INDIRECT_EDGE; scores/PPIs are frozen during the
//    sweep, so this is a monotone propagation in index order, iterated to its fixed point
//  * per-point relabelling  argmax_k( n.o_k + w_v * smooth_k )  is independent per point (fp64, reference order)
#include <cub/device/device_radix_sort.cuh>

#include <limits.h>

#include <algorithm>

#include "stages.cuh"

namespace pccb200 {

namespace {

enum : uint8_t { NO_EDGE = 0x00, INDIRECT_EDGE = 0x01, M_DIRECT_EDGE = 0x10, S_DIRECT_EDGE = 0x11 };

struct GridGeom {
  int voxShift, gridShift, half;
  int gmin[3], gdim[3];  // bounding box of voxel centres (dense lookup grid)
};

__global__ void kMaxCoord( const short4* __restrict__ pts, int n, int* __restrict__ out ) {
  int       i = blockIdx.x * blockDim.x + threadIdx.x;
  int       m = INT_MIN;
  if ( i < n ) {
    const short4 p = pts[i];
    m              = max( int( p.x ), max( int( p.y ), int( p.z ) ) );
  }
  m = __reduce_max_sync( 0xffffffffu, m );
  if ( ( threadIdx.x & 31 ) == 0 ) atomicMax( out, m );
}

__global__ void kVoxelKeys( const short4* __restrict__ pts, int n, int voxShift, int gridShift, int half, uint32_t* __restrict__ keys,
                            uint32_t* __restrict__ ids ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  const short4   p = pts[i];
  const uint32_t x = ( uint32_t( p.x ) + half ) >> voxShift, y = ( uint32_t( p.y ) + half ) >> voxShift,
                 z = ( uint32_t( p.z ) + half ) >> voxShift;
  keys[i] = x + ( y << gridShift ) + ( z << ( 2 * gridShift ) );
  ids[i]  = i;
}

__global__ void kHeads( const uint32_t* __restrict__ keys, int n, uint32_t* __restrict__ head ) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if ( p < n ) head[p] = ( p == 0 || keys[p] != keys[p - 1] ) ? 1u : 0u;
}

// per key-run: first position and first (smallest) point index
__global__ void kRunStarts( const uint32_t* __restrict__ head, const uint32_t* __restrict__ headScan, const uint32_t* __restrict__ idsSorted,
                            int n, uint32_t* __restrict__ runStart, uint32_t* __restrict__ runFirstIdx, uint32_t* __restrict__ runId ) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if ( p >= n || !head[p] ) return;
  const uint32_t r = headScan[p];
  runStart[r]      = p;
  runFirstIdx[r]   = idsSorted[p];
  runId[r]         = r;
}

// voxel v (in first-appearance order) <- run order[v]
__global__ void kVoxelSetup( const uint32_t* __restrict__ order, const uint32_t* __restrict__ runStart, int V, int n,
                             const uint32_t* __restrict__ idsSorted, const short4* __restrict__ pts, int voxShift, int half,
                             uint32_t* __restrict__ voxStart, uint32_t* __restrict__ voxCount, short4* __restrict__ centers,
                             uint32_t* __restrict__ runToVox ) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if ( v >= V ) return;
  const uint32_t r = order[v];
  const uint32_t s = runStart[r];
  // run r ends where the next run (in key order) starts
  const uint32_t e = ( r + 1 < uint32_t( V ) ) ? runStart[r + 1] : uint32_t( n );
  voxStart[v]      = s;
  voxCount[v]      = e - s;
  const short4 p   = pts[idsSorted[s]];
  centers[v] = make_short4( short( ( uint32_t( p.x ) + half ) >> voxShift ), short( ( uint32_t( p.y ) + half ) >> voxShift ),
                            short( ( uint32_t( p.z ) + half ) >> voxShift ), 0 );
  runToVox[r] = v;
}

__global__ void kCenterBounds( const short4* __restrict__ centers, int V, int* __restrict__ mm ) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  int       lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  if ( v < V ) {
    const short4 c = centers[v];
    lo[0] = hi[0] = c.x, lo[1] = hi[1] = c.y, lo[2] = hi[2] = c.z;
  }
#pragma unroll
  for ( int d = 0; d < 3; ++d ) {
    const int a = __reduce_min_sync( 0xffffffffu, lo[d] ), b = __reduce_max_sync( 0xffffffffu, hi[d] );
    if ( ( threadIdx.x & 31 ) == 0 ) atomicMin( &mm[d], a ), atomicMax( &mm[3 + d], b );
  }
}

__global__ void kFillGrid( const short4* __restrict__ centers, int V, GridGeom g, int* __restrict__ grid ) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if ( v >= V ) return;
  const short4 c = centers[v];
  grid[( size_t( c.z - g.gmin[2] ) * g.gdim[1] + ( c.y - g.gmin[1] ) ) * g.gdim[0] + ( c.x - g.gmin[0] )] = v + 1;
}

// ---- adjacency --------------------------------------------------------------------------------------
constexpr int kMaxHits      = 2048;  // >= number of lattice offsets with d^2 < 48 (1393), power of two for the sort
constexpr int kAdjWarps     = 4;
constexpr int kMaxNear      = 27;

struct AdjOut {
  uint32_t* adjOff;    // V
  uint32_t* adjLen;    // V
  uint32_t* adjData;   // capacity
  uint32_t* nearData;  // V x kMaxNear
  uint8_t*  nearLen;   // V
  double*   weight;    // V
  unsigned long long* allocCursor;
  unsigned long long  capacity;
  int*      overflow;
};

__global__ void __launch_bounds__( 32 * kAdjWarps )
    kAdjacency( const short4* __restrict__ centers, const uint32_t* __restrict__ voxCount, int V, GridGeom g, const int* __restrict__ grid,
                const int* __restrict__ offsets, int numOffsets, int maxNN, int nearRange, double lambda, AdjOut out ) {
  __shared__ uint32_t hits[kAdjWarps][kMaxHits];
  const int           warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int           v = blockIdx.x * kAdjWarps + warp;
  if ( v >= V ) return;
  uint32_t*    h = hits[warp];
  const short4 c = centers[v];
  int          nh = 0;
  for ( int base = 0; base < numOffsets; base += 32 ) {
    uint32_t key = 0xFFFFFFFFu;
    if ( base + lane < numOffsets ) {
      const int o  = offsets[base + lane];  // dx,dy,dz in signed bytes, d2 in the top byte
      const int dx = (signed char)( o & 0xff ), dy = (signed char)( ( o >> 8 ) & 0xff ), dz = (signed char)( ( o >> 16 ) & 0xff );
      const int x = c.x + dx - g.gmin[0], y = c.y + dy - g.gmin[1], z = c.z + dz - g.gmin[2];
      if ( x >= 0 && y >= 0 && z >= 0 && x < g.gdim[0] && y < g.gdim[1] && z < g.gdim[2] ) {
        const int id = grid[( size_t( z ) * g.gdim[1] + y ) * g.gdim[0] + x];
        if ( id > 0 ) key = ( uint32_t( o >> 24 ) << 24 ) | uint32_t( id - 1 );
      }
    }
    const unsigned m = __ballot_sync( 0xffffffffu, key != 0xFFFFFFFFu );
    if ( key != 0xFFFFFFFFu ) h[nh + __popc( m & ( ( 1u << lane ) - 1 ) )] = key;
    nh += __popc( m );
  }
  // pad to a power of two and bitonic-sort ascending by (d2, index)
  int P = 32;
  while ( P < nh ) P <<= 1;
  for ( int i = nh + lane; i < P; i += 32 ) h[i] = 0xFFFFFFFFu;
  __syncwarp();
  for ( int k = 2; k <= P; k <<= 1 )
    for ( int j = k >> 1; j > 0; j >>= 1 ) {
      for ( int i = lane; i < P; i += 32 ) {
        const int l = i ^ j;
        if ( l > i ) {
          const uint32_t a = h[i], b = h[l];
          const bool     up = ( i & k ) == 0;
          if ( ( a > b ) == up ) h[i] = b, h[l] = a;
        }
      }
      __syncwarp();
    }
  // cut where the running point count first reaches maxNN (that entry included)
  uint32_t running = 0;
  int      len     = nh;
  for ( int base = 0; base < nh; base += 32 ) {
    uint32_t cnt = 0;
    if ( base + lane < nh ) cnt = voxCount[h[base + lane] & 0xFFFFFFu] & 0xffu;  // getPointCount() is uint8_t
    uint32_t inc = cnt;
#pragma unroll
    for ( int o = 1; o < 32; o <<= 1 ) {
      const uint32_t t = __shfl_up_sync( 0xffffffffu, inc, o );
      if ( lane >= o ) inc += t;
    }
    const unsigned reach = __ballot_sync( 0xffffffffu, base + lane < nh && running + inc >= uint32_t( maxNN ) );
    if ( reach ) {
      const int first = __ffs( reach ) - 1;
      len             = base + first + 1;
      running += __shfl_sync( 0xffffffffu, inc, first );
      break;
    }
    running += __shfl_sync( 0xffffffffu, inc, 31 );
  }
  // allocate + write
  unsigned long long off = 0;
  if ( lane == 0 ) off = atomicAdd( out.allocCursor, (unsigned long long)len );
  off = __shfl_sync( 0xffffffffu, off, 0 );
  if ( off + len > out.capacity ) {
    if ( lane == 0 ) *out.overflow = 1;
    return;
  }
  int nearCount = 0;
  for ( int base = 0; base < len; base += 32 ) {
    bool     isNear = false;
    uint32_t o      = 0;
    if ( base + lane < len ) {
      o                        = h[base + lane] & 0xFFFFFFu;
      out.adjData[off + base + lane] = o;
      const short4 q           = centers[o];
      isNear = abs( int( c.x ) - q.x ) <= nearRange && abs( int( c.y ) - q.y ) <= nearRange && abs( int( c.z ) - q.z ) <= nearRange;
    }
    const unsigned m = __ballot_sync( 0xffffffffu, isNear );
    if ( isNear ) {
      const int slot = nearCount + __popc( m & ( ( 1u << lane ) - 1 ) );
      if ( slot < kMaxNear ) out.nearData[size_t( v ) * kMaxNear + slot] = o;
    }
    nearCount += __popc( m );
  }
  if ( lane == 0 ) {
    out.adjOff[v]  = uint32_t( off );
    out.adjLen[v]  = uint32_t( len );
    out.nearLen[v] = uint8_t( min( nearCount, kMaxNear ) );
    out.weight[v]  = lambda / double( running );
  }
}

// ---- voxel state -------------------------------------------------------------------------------------
constexpr uint32_t kNoEntry = 0xFFFFFFFFu;  // work-list slot reserved but not written yet
struct VoxState {
  uint16_t* score;  // V x 6 (stored as 8 x u16 = 16 B rows)
  uint8_t*  edge;
  uint8_t*  ppi;
  uint8_t*  dirty;
  uint8_t*  mark;
  uint8_t*  active;
};

// recount histogram; apply INDIRECT marks; re-derive class/PPI for dirty voxels (updateScores); then the work list of the NEXT
// sweep: the voxels that are edge voxels at its start. `list` / `ctl` are the next sweep's (counter 0: kSweepStatic cleared it).
__global__ void kRecountAndActivate( VoxState st, const uint32_t* __restrict__ voxStart, const uint32_t* __restrict__ voxCount,
                                     const uint32_t* __restrict__ idsSorted, const uint8_t* __restrict__ partition, int V, int initialise,
                                     uint32_t* __restrict__ list, unsigned* __restrict__ ctl, uint32_t* __restrict__ tail, unsigned* __restrict__ tailCtl ) {
  const int v  = blockIdx.x * blockDim.x + threadIdx.x;
  bool      on = false;
  if ( v < 4 ) tailCtl[v] = 0;
  if ( v < V ) {
    tail[v]          = kNoEntry;  // (the tail list of the sweep that just ended is dead: reset for the next one)
    const uint32_t s = voxStart[v], c = voxCount[v];
    uint8_t edge = st.edge[v];
    bool    dirty = st.dirty[v] != 0;
    if ( initialise ) {
      edge  = ( uint8_t( c ) == 1 ) ? S_DIRECT_EDGE : M_DIRECT_EDGE;
      dirty = true;
    } else if ( st.mark[v] && edge == NO_EDGE ) {
      edge = INDIRECT_EDGE;
    }
    if ( dirty ) {  // (only a relabelled voxel has a new histogram: the others keep their row)
      uint16_t sc[6] = {0, 0, 0, 0, 0, 0};
      for ( uint32_t j = 0; j < c; ++j ) ++sc[partition[idsSorted[s + j]]];
#pragma unroll
      for ( int k = 0; k < 6; ++k ) st.score[size_t( v ) * 8 + k] = sc[k];
      if ( edge != S_DIRECT_EDGE ) {
        int used = 0;
#pragma unroll
        for ( int k = 0; k < 6; ++k ) used += sc[k] != 0;
        edge = used == 1 ? NO_EDGE : M_DIRECT_EDGE;
      }
      int best = 0;
#pragma unroll
      for ( int k = 1; k < 6; ++k )
        if ( sc[k] > sc[best] ) best = k;
      st.ppi[v] = uint8_t( best );
    }
    st.edge[v]   = edge;
    st.dirty[v]  = 0;
    st.mark[v]   = 0;
    on           = edge != NO_EDGE;
    st.active[v] = on ? 1 : 0;
  }
  const unsigned m    = __ballot_sync( 0xffffffffu, on );
  unsigned       base = 0;
  if ( ( threadIdx.x & 31 ) == 0 && m ) base = atomicAdd( ctl, __popc( m ) );
  base = __shfl_sync( 0xffffffffu, base, 0 );
  if ( on ) list[base + __popc( m & ( ( 1u << ( threadIdx.x & 31 ) ) - 1 ) )] = v;
}

// One voxel of a sweep (one warp): smooth = sum of the neighbours' histograms (uint16 wrap-around like ScoresVector_t), top = first
// arg-max; the 2nd voxel classification: NO_EDGE neighbours whose PPI differs are marked, and those with a larger index join this
// very sweep (appended to the TAIL list); then the voxel's points are relabelled (argmax of n.o_k + w_v * smooth_k) - scores,
// classes and PPIs are frozen during a sweep (the recount launch updates them), so relabelling inside the sweep changes nothing
// another warp reads, and the result does not depend on the processing order (a voxel is only ever activated by a smaller index).
struct SweepData {
  VoxState        st;
  const uint32_t *adjOff, *adjLen, *adjData, *nearData;
  const uint8_t*  nearLen;
  const double*   weight;
  const uint32_t *voxStart, *voxCount, *idsSorted;
  const double*   normals;
  uint8_t*        partition;
  uint32_t*       tail;     // voxels activated during the sweep
  unsigned*       tailCtl;  // [0] entries reserved, [1] next ticket, [2] entries finished
};
// Loads are issued in three dependency levels, everything that only needs v first (the structure's pointers may alias as far as
// the compiler knows, so program order is what lets the loads of the later steps overlap the adjacency gather):
//   1. adjOff/adjLen, nearLen/nearData, edge/ppi/weight/voxStart/voxCount of v      2. adjData, edge/ppi of the near voxels, point ids
//   3. the neighbours' score rows, the points' normals
__device__ __forceinline__ void sweepVoxel( const SweepData& d, uint32_t v, int lane ) {
  const uint32_t off = d.adjOff[v], len = d.adjLen[v];
  const int      nl  = d.nearLen[v];
  const uint32_t o   = lane < kMaxNear ? d.nearData[size_t( v ) * kMaxNear + lane] : 0u;
  uint8_t        edgeHere = d.st.edge[v];
  const uint8_t  ppiHere  = d.st.ppi[v];
  const double   wv  = d.weight[v];
  const uint32_t st0 = d.voxStart[v], c = d.voxCount[v];
  uint8_t        edgeNear = 0xff, ppiNear = 0;
  if ( lane < nl ) edgeNear = d.st.edge[o], ppiNear = d.st.ppi[o];
  // (a voxel holds at most 64 points: 4 x 4 x 4 positions -> two per lane)
  const bool     has0 = uint32_t( lane ) < c, has1 = uint32_t( lane ) + 32 < c;
  const uint32_t p0 = has0 ? d.idsSorted[st0 + lane] : 0u, p1 = has1 ? d.idsSorted[st0 + lane + 32] : 0u;
  uint32_t       s[6] = {0, 0, 0, 0, 0, 0};
  for ( uint32_t i = lane; i < len; i += 32 ) {
    const uint4 r = *reinterpret_cast<const uint4*>( d.st.score + size_t( d.adjData[off + i] ) * 8 );
    s[0] += r.x & 0xffff, s[1] += r.x >> 16, s[2] += r.y & 0xffff, s[3] += r.y >> 16, s[4] += r.z & 0xffff, s[5] += r.z >> 16;
  }
  double n0[3] = {0.0, 0.0, 0.0}, n1[3] = {0.0, 0.0, 0.0};
  if ( has0 ) n0[0] = d.normals[3 * size_t( p0 )], n0[1] = d.normals[3 * size_t( p0 ) + 1], n0[2] = d.normals[3 * size_t( p0 ) + 2];
  if ( has1 ) n1[0] = d.normals[3 * size_t( p1 )], n1[1] = d.normals[3 * size_t( p1 ) + 1], n1[2] = d.normals[3 * size_t( p1 ) + 2];
#pragma unroll
  for ( int k = 0; k < 6; ++k ) s[k] = __reduce_add_sync( 0xffffffffu, s[k] ) & 0xffffu;
  int top = 0;
#pragma unroll
  for ( int k = 1; k < 6; ++k )
    if ( s[k] > s[top] ) top = k;
  if ( lane < nl && edgeNear == NO_EDGE && ppiNear != top ) {
    d.st.mark[o] = 1;
    if ( o > v ) {
      // byte-granular test-and-set on the active flags
      unsigned*      word = reinterpret_cast<unsigned*>( d.st.active ) + ( o >> 2 );
      const unsigned bit  = 1u << ( 8 * ( o & 3 ) );
      const unsigned old  = atomicOr( word, bit );
      if ( !( old & bit ) ) {
        const unsigned at = atomicAdd( &d.tailCtl[0], 1u );
        *reinterpret_cast<volatile uint32_t*>( &d.tail[at] ) = o;
      }
    }
  }
  // relabel the voxel's points (one lane per point)
  if ( edgeHere == NO_EDGE ) edgeHere = INDIRECT_EDGE;  // activated during this sweep
  if ( edgeHere != M_DIRECT_EDGE ) {
    int used = 0;
#pragma unroll
    for ( int k = 0; k < 6; ++k ) used += s[k] != 0;
    if ( used == 1 && s[ppiHere] > 0 ) return;
  }
  auto label = [&]( const double n[3] ) {
    const double x = n[0], y = n[1], z = n[2];
    const double dd[6] = {x * 1.0 + y * 0.0 + z * 0.0,  x * 0.0 + y * 1.0 + z * 0.0,  x * 0.0 + y * 0.0 + z * 1.0,
                          x * -1.0 + y * 0.0 + z * 0.0, x * 0.0 + y * -1.0 + z * 0.0, x * 0.0 + y * 0.0 + z * -1.0};
    int          best = 0;
    double       bs   = dd[0] + wv * double( uint16_t( s[0] ) );
#pragma unroll
    for ( int k = 1; k < 6; ++k ) {
      const double sc = dd[k] + wv * double( uint16_t( s[k] ) );
      if ( sc > bs ) bs = sc, best = k;
    }
    return uint8_t( best );
  };
  if ( has0 ) d.partition[p0] = label( n0 );
  if ( has1 ) d.partition[p1] = label( n1 );
  for ( uint32_t jj = lane + 64; jj < c; jj += 32 ) {  // (not reached with voxel dimension 4; kept for generality)
    const uint32_t p  = d.idsSorted[st0 + jj];
    const double   nn[3] = {d.normals[3 * size_t( p )], d.normals[3 * size_t( p ) + 1], d.normals[3 * size_t( p ) + 2]};
    d.partition[p]       = label( nn );
  }
  if ( lane == 0 ) d.st.dirty[v] = 1;
}

// Sweep, launch 1 of 3: the voxels that are edge voxels at sweep start (the list kRecountAndActivate left, complete before this
// launch) are dealt out by warp index - no tickets, no polling. Also clears the OTHER initial list's counter (the next sweep's).
__device__ __forceinline__ unsigned ldVolatile( const unsigned* p ) { return *reinterpret_cast<const volatile unsigned*>( p ); }
__global__ void __launch_bounds__( 128 ) kSweepStatic( SweepData d, const uint32_t* __restrict__ list, const unsigned* __restrict__ ctl, unsigned* __restrict__ nextCtl ) {
  const int      lane  = threadIdx.x & 31;
  const unsigned count = ctl[0], nWarps = gridDim.x * ( blockDim.x / 32 );
  if ( blockIdx.x == 0 && threadIdx.x == 0 ) nextCtl[0] = 0;
  for ( unsigned w = ( blockIdx.x * blockDim.x + threadIdx.x ) / 32; w < count; w += nWarps ) sweepVoxel( d, list[w], lane );
}

// Sweep, launch 2 of 3: the voxels activated during the sweep - a persistent kernel over the growing tail list. A warp takes a
// ticket, waits until that list slot is written (or until every reserved entry is finished: nothing can be appended any more) and
// processes it. Only warps that hold a ticket wait, and an entry is reserved by a RUNNING warp that writes it at once: no
// dependence on CTAs that are not resident yet. With an empty tail the first ticket sees "0 reserved, 0 finished" and the kernel ends.
__global__ void __launch_bounds__( 128 ) kSweepTail( SweepData d, unsigned maxSleepNs ) {
  const int lane = threadIdx.x & 31;
  unsigned* ctl  = d.tailCtl;
  for ( ;; ) {
    uint32_t v = kNoEntry;
    if ( lane == 0 ) {
      const unsigned w   = atomicAdd( &ctl[1], 1u );
      unsigned       nap = 128;  // (the waiters poll L2: back off, the other frames' kernels share it)
      for ( ;; ) {
        const unsigned c0 = ldVolatile( &ctl[0] );
        if ( w < c0 ) {
          while ( ( v = ldVolatile( &d.tail[w] ) ) == kNoEntry ) __nanosleep( 20 );
          break;
        }
        const unsigned done = ldVolatile( &ctl[2] );
        if ( done == c0 && ldVolatile( &ctl[0] ) == c0 ) break;  // quiescent: all reserved entries finished, none in flight
        __nanosleep( nap );
        nap = min( nap * 2, maxSleepNs );
      }
    }
    v = __shfl_sync( 0xffffffffu, v, 0 );
    if ( v == kNoEntry ) return;
    sweepVoxel( d, v, lane );
    __threadfence();  // the entries appended above are reserved (and written) before this one counts as finished
    __syncwarp();
    if ( lane == 0 ) atomicAdd( &ctl[2], 1u );
  }
}

std::vector<int> makeOffsets( int r2 ) {
  std::vector<int> o;
  int              r = 0;
  while ( ( r + 1 ) * ( r + 1 ) < r2 ) ++r;
  for ( int dz = -r; dz <= r; ++dz )
    for ( int dy = -r; dy <= r; ++dy )
      for ( int dx = -r; dx <= r; ++dx ) {
        const int d2 = dx * dx + dy * dy + dz * dz;
        if ( d2 < r2 ) o.push_back( ( dx & 0xff ) | ( ( dy & 0xff ) << 8 ) | ( ( dz & 0xff ) << 16 ) | ( d2 << 24 ) );
      }
  std::sort( o.begin(), o.end(), []( int a, int b ) { return ( unsigned( a ) >> 24 ) < ( unsigned( b ) >> 24 ); } );
  return o;
}

}  // namespace

void refineSegmentation( RefineScratch& sc, const short4* pts, const double* normals, size_t n, const pccb200_seg_params& prm,
                         uint8_t* partition, cudaStream_t s ) {
  if ( n == 0 ) return;
  const int N = int( n );
  // ---- geometry of the voxel grid (PCCPatchSegmenter.cpp:1397-1413)
  sc.ints.reserve( 64 );
  int init[8] = {INT_MIN, 0, INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
  PCC_CUDA( cudaMemcpyAsync( sc.ints, init, sizeof( init ), cudaMemcpyHostToDevice, s ) );
  kMaxCoord<<<divUp( n, 256 ), 256, 0, s>>>( pts, N, sc.ints );
  int geoMax = 0;
  PCC_CUDA( cudaMemcpyAsync( &geoMax, sc.ints, sizeof( int ), cudaMemcpyDeviceToHost, s ) );
  streamWait( s );
  size_t geoRange = 1;
  for ( size_t i = size_t( int16_t( geoMax ) - 1 ); i != 0u; i >>= 1, geoRange <<= 1 ) {}
  int voxShift = 0, gridShift = 0;
  for ( size_t i = size_t( prm.voxel_dim_refine ); i > 1; ++voxShift, i >>= 1 ) {}
  const size_t gridDim = geoRange >> voxShift;
  for ( size_t i = gridDim; i > 1; ++gridShift, i >>= 1 ) {}
  if ( 3 * gridShift + 2 > 31 || prm.voxel_dim_refine > 4 ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
  const int half = prm.voxel_dim_refine >> 1;

  // ---- voxels in first-appearance order
  sc.keysA.reserve( n ), sc.keysB.reserve( n ), sc.idsA.reserve( n ), sc.idsSorted.reserve( n ), sc.head.reserve( n + 1 );
  sc.headScan.reserve( n + 1 ), sc.scanTmp.reserve( scanTmpElems( n ) );
  kVoxelKeys<<<divUp( n, 256 ), 256, 0, s>>>( pts, N, voxShift, gridShift, half, sc.keysA, sc.idsA );
  size_t tmpBytes = 0;
  const int keyBits = std::min( 32, 3 * gridShift + 2 );
  PCC_CUDA( cub::DeviceRadixSort::SortPairs( nullptr, tmpBytes, sc.keysA.p, sc.keysB.p, sc.idsA.p, sc.idsSorted.p, N, 0, keyBits, s ) );
  sc.cubTmp.reserve( tmpBytes + 16 );
  PCC_CUDA( cub::DeviceRadixSort::SortPairs( sc.cubTmp.p, tmpBytes, sc.keysA.p, sc.keysB.p, sc.idsA.p, sc.idsSorted.p, N, 0, keyBits, s ) );
  kHeads<<<divUp( n, 256 ), 256, 0, s>>>( sc.keysB, N, sc.head );
  exclusiveScanU32( sc.head, sc.headScan, n, sc.scanTmp, s );
  uint32_t Vu = 0;
  PCC_CUDA( cudaMemcpyAsync( &Vu, sc.headScan.p + n, sizeof( uint32_t ), cudaMemcpyDeviceToHost, s ) );
  streamWait( s );
  const int V = int( Vu );
  if ( V >= ( 1 << 24 ) ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
  sc.runStart.reserve( V + 1 ), sc.runFirst.reserve( V ), sc.runId.reserve( V ), sc.runFirstSorted.reserve( V ), sc.order.reserve( V );
  sc.voxStart.reserve( V ), sc.voxCount.reserve( V ), sc.centers.reserve( V ), sc.runToVox.reserve( V );
  kRunStarts<<<divUp( n, 256 ), 256, 0, s>>>( sc.head, sc.headScan, sc.idsSorted, N, sc.runStart, sc.runFirst, sc.runId );
  int idxBits = 1;
  while ( ( size_t( 1 ) << idxBits ) < n ) ++idxBits;
  PCC_CUDA( cub::DeviceRadixSort::SortPairs( nullptr, tmpBytes, sc.runFirst.p, sc.runFirstSorted.p, sc.runId.p, sc.order.p, V, 0, idxBits, s ) );
  sc.cubTmp.reserve( tmpBytes + 16 );
  PCC_CUDA( cub::DeviceRadixSort::SortPairs( sc.cubTmp.p, tmpBytes, sc.runFirst.p, sc.runFirstSorted.p, sc.runId.p, sc.order.p, V, 0, idxBits, s ) );
  kVoxelSetup<<<divUp( V, 256 ), 256, 0, s>>>( sc.order, sc.runStart, V, N, sc.idsSorted, pts, voxShift, half, sc.voxStart, sc.voxCount,
                                               sc.centers, sc.runToVox );
  // ---- dense lookup grid over the centres' bounding box
  kCenterBounds<<<divUp( V, 256 ), 256, 0, s>>>( sc.centers, V, sc.ints.p + 2 );
  int mm[6];
  PCC_CUDA( cudaMemcpyAsync( mm, sc.ints.p + 2, sizeof( mm ), cudaMemcpyDeviceToHost, s ) );
  streamWait( s );
  GridGeom g;
  g.voxShift = voxShift, g.gridShift = gridShift, g.half = half;
  size_t cells = 1;
  for ( int d = 0; d < 3; ++d ) g.gmin[d] = mm[d], g.gdim[d] = mm[3 + d] - mm[d] + 1, cells *= size_t( g.gdim[d] );
  sc.grid.reserve( cells );
  PCC_CUDA( cudaMemsetAsync( sc.grid, 0, cells * sizeof( int ), s ) );
  kFillGrid<<<divUp( V, 256 ), 256, 0, s>>>( sc.centers, V, g, sc.grid );
  // ---- adjacency within the search radius, cut at maxNN points
  const int r2 = prm.search_radius_refine >> voxShift;
  if ( sc.offsetsR2 != r2 ) {
    std::vector<int> off = makeOffsets( r2 );
    if ( off.size() > size_t( kMaxHits ) ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
    sc.offsets.reserve( off.size() );
    PCC_CUDA( cudaMemcpyAsync( sc.offsets, off.data(), off.size() * sizeof( int ), cudaMemcpyHostToDevice, s ) );
    streamWait( s );
    sc.offsetsR2 = r2, sc.numOffsets = int( off.size() );
  }
  sc.adjOff.reserve( V ), sc.adjLen.reserve( V ), sc.nearData.reserve( size_t( V ) * kMaxNear ), sc.nearLen.reserve( V ), sc.weight.reserve( V );
  sc.cursor.reserve( 2 );
  size_t capacity = std::max( sc.adjData.cap, size_t( V ) * 192 );
  for ( ;; ) {
    sc.adjData.reserve( capacity );
    PCC_CUDA( cudaMemsetAsync( sc.cursor, 0, 2 * sizeof( unsigned long long ), s ) );
    AdjOut out{ sc.adjOff, sc.adjLen, sc.adjData, sc.nearData, sc.nearLen, sc.weight, sc.cursor.p, (unsigned long long)sc.adjData.cap,
                reinterpret_cast<int*>( sc.cursor.p + 1 ) };
    kAdjacency<<<divUp( V, kAdjWarps ), 32 * kAdjWarps, 0, s>>>( sc.centers, sc.voxCount, V, g, sc.grid, sc.offsets, sc.numOffsets,
                                                                prm.max_nn_count_refine, prm.voxel_dim_refine >= 4 ? 1 : 2,
                                                                prm.lambda_refine, out );
    PCC_LAUNCH_CHECK();
    unsigned long long res[2];
    PCC_CUDA( cudaMemcpyAsync( res, sc.cursor, sizeof( res ), cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    if ( !( res[1] & 0xffffffffull ) ) break;
    capacity = size_t( res[0] ) + size_t( V );  // cursor counted every request, so this is the exact need
  }
  // ---- voxel state + sweeps
  sc.score.reserve( size_t( V ) * 8 ), sc.smooth.reserve( size_t( V ) * 8 ), sc.edge.reserve( V + 4 ), sc.ppi.reserve( V + 4 );
  sc.dirty.reserve( V + 4 ), sc.mark.reserve( V + 4 ), sc.active.reserve( V + 8 );
  PCC_CUDA( cudaMemsetAsync( sc.score, 0, size_t( V ) * 8 * sizeof( uint16_t ), s ) );
  PCC_CUDA( cudaMemsetAsync( sc.active, 0, V + 8, s ) );
  PCC_CUDA( cudaMemsetAsync( sc.edge, 0, V, s ) );
  PCC_CUDA( cudaMemsetAsync( sc.mark, 0, V, s ) );
  PCC_CUDA( cudaMemsetAsync( sc.dirty, 0, V, s ) );
  VoxState st{ sc.score, sc.edge, sc.ppi, sc.dirty, sc.mark, sc.active };
  // Three launches per sweep and no host round trip: the initial work list of a sweep is dealt out statically (kSweepStatic), the
  // voxels activated during the sweep go through a device-side growing list (kSweepTail), the recount fills the next initial list
  // (two initial lists / counters alternate).
  sc.list.reserve( 3 * size_t( V ) ), sc.count.reserve( 12 );
  uint32_t* lists[2] = { sc.list.p, sc.list.p + V };
  unsigned* ctls[2]  = { sc.count.p, sc.count.p + 4 };
  uint32_t* tail     = sc.list.p + 2 * size_t( V );
  unsigned* tailCtl  = sc.count.p + 8;
  PCC_CUDA( cudaMemsetAsync( sc.count, 0, 12 * sizeof( unsigned ), s ) );
  kRecountAndActivate<<<divUp( V, 128 ), 128, 0, s>>>( st, sc.voxStart, sc.voxCount, sc.idsSorted, partition, V, 1, lists[0], ctls[0], tail, tailCtl );
  const int iterations = std::max( 1, prm.iteration_count_refine );
  static const int ctasPerSm = [] {  // (tuning knob of the static part; 8 x 128 threads per SM)
    const char* e = getenv( "PCCB200_SWEEP_CTAS_PER_SM" );
    return e && atoi( e ) > 0 ? atoi( e ) : 8;
  }();
  // (tuning knob of the tail: its warps mostly WAIT - for tickets, for cascades - so few of them is better when other frames' kernels
  //  share the SMs: 2 / 4 / 8 CTAs per SM gave 56.1 / 52.5 / 48.5 Mpoints/s in the default bench, profiles/r02f, r02h, r02k)
  static const int tailCtasPerSm = [] {
    const char* e = getenv( "PCCB200_SWEEP_TAIL_CTAS_PER_SM" );
    return e && atoi( e ) > 0 ? atoi( e ) : 2;
  }();
  const int staticCtas = int( std::min<size_t>( divUp( V, 4 ), size_t( 148 ) * ctasPerSm ) );
  const int tailCtas   = int( std::min<size_t>( divUp( V, 4 ), size_t( 148 ) * tailCtasPerSm ) );
  static const unsigned maxSleepNs = [] {
    const char* e = getenv( "PCCB200_SWEEP_MAX_SLEEP_NS" );
    return unsigned( e && atoi( e ) > 0 ? atoi( e ) : 2048 );
  }();
  SweepData d{ st, sc.adjOff, sc.adjLen, sc.adjData, sc.nearData, sc.nearLen, sc.weight, sc.voxStart, sc.voxCount, sc.idsSorted, normals, partition, tail, tailCtl };
  for ( int it = 0; it < iterations; ++it ) {
    const int a = it & 1, b = a ^ 1;
    kSweepStatic<<<staticCtas, 128, 0, s>>>( d, lists[a], ctls[a], ctls[b] );
    kSweepTail<<<tailCtas, 128, 0, s>>>( d, maxSleepNs );
    kRecountAndActivate<<<divUp( V, 128 ), 128, 0, s>>>( st, sc.voxStart, sc.voxCount, sc.idsSorted, partition, V, 0, lists[b], ctls[b], tail, tailCtl );
    PCC_LAUNCH_CHECK();
  }
}

}  // namespace pccb200
