// color.cu — colour transfer from the source cloud to the reconstructed cloud
// (PCCPointSet3::transferColors, PccLibCommon/source/PCCPointSet.cpp:807-1124) with the CTC parameter values:
// 8 forward / 1 backward neighbours, distance-weighted means with offset 4, every distance / colour gate disabled
// (thresholds 1000 >= 512 become DBL_MAX), bestColorSearchRange 0, fixWeight (w = 0).  Then the colour of a
// reconstructed point is
//    * the weighted mean ( w = 1/(sqrt(d2)+4) ) of the source points whose nearest reconstructed point it is
//      ("backward votes", ordered by distance; a coinciding source point wins outright), or
//    * without votes: the forward value = weighted mean ( w = 1/(d2+4) ) of its 8 nearest source points in
//      nanoflann result order (a coinciding source point wins outright).
// Both neighbour searches run on nanoflann-identical trees (kdtree.cu), so ties resolve as in the reference.
// Votes are grouped per target with one stable radix sort on (target, distance) — equal distances keep ascending source
// order, which is what the reference's insertion sort (std::sort on <= 16 elements) produces.
#include <cub/device/device_radix_sort.cuh>

#include "stages.cuh"

namespace pccb200 {

namespace {

__device__ __forceinline__ uint8_t roundClip( double v ) {
  const double r = round( v );
  return uint8_t( r < 0.0 ? 0.0 : ( r > 255.0 ? 255.0 : r ) );
}

__global__ void __launch_bounds__( 128 )
    kForward( const uint32_t* __restrict__ idx, const float* __restrict__ dist, const uchar4* __restrict__ srcRgb, int R, uchar4* __restrict__ out ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= R ) return;
  const uint32_t* id = idx + size_t( i ) * 8;
  const float*    d  = dist + size_t( i ) * 8;
  int             cnt = 0;
  while ( cnt < 8 && id[cnt] != 0xFFFFFFFFu ) ++cnt;
  if ( cnt == 0 ) {
    out[i] = make_uchar4( 0, 0, 0, 0 );
    return;
  }
  const uchar4 c0 = srcRgb[id[0]];
  if ( double( d[0] ) < 0.0001 || cnt == 1 ) {
    out[i] = c0;
    return;
  }
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, sw = 0.0;
  for ( int j = 0; j < cnt; ++j ) {
    const double w = 1 / ( double( d[j] ) + 4.0 );
    const uchar4 c = srcRgb[id[j]];
    a0 += double( c.x ) * w, a1 += double( c.y ) * w, a2 += double( c.z ) * w;
    sw += w;
  }
  out[i] = make_uchar4( roundClip( a0 / sw ), roundClip( a1 / sw ), roundClip( a2 / sw ), 0 );
}

__global__ void kVoteKeys( const uint32_t* __restrict__ target, const float* __restrict__ dist, int n, uint64_t* __restrict__ keys,
                           uint32_t* __restrict__ ids ) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if ( s >= n ) return;
  keys[s] = ( uint64_t( target[s] ) << 24 ) | uint64_t( uint32_t( dist[s] ) & 0xFFFFFFu );
  ids[s]  = s;
}

// one thread per vote position; the head of each target's run accumulates the whole run in order
__global__ void __launch_bounds__( 128 )
    kBackward( const uint64_t* __restrict__ keys, const uint32_t* __restrict__ srcIds, int n, const uchar4* __restrict__ srcRgb,
               uchar4* __restrict__ recRgb ) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if ( p >= n ) return;
  const uint32_t tgt = uint32_t( keys[p] >> 24 );
  if ( p > 0 && uint32_t( keys[p - 1] >> 24 ) == tgt ) return;
  int e = p + 1;
  while ( e < n && uint32_t( keys[e] >> 24 ) == tgt ) ++e;
  const uchar4 first = srcRgb[srcIds[p]];
  if ( ( keys[p] & 0xFFFFFFu ) == 0 || e - p == 1 ) {
    recRgb[tgt] = first;  // a source point at this very position, or a single vote: round(1.0 * c) == c
    return;
  }
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, sw = 0.0;
  for ( int j = p; j < e; ++j ) {
    const double w = 1 / ( sqrt( double( uint32_t( keys[j] & 0xFFFFFFu ) ) ) + 4.0 );
    const uchar4 c = srcRgb[srcIds[j]];
    a0 += ( double( c.x ) * w ), a1 += ( double( c.y ) * w ), a2 += ( double( c.z ) * w );
    sw += w;
  }
  // color0 = clip( round( w*centroid1 + (1-w)*centroid2 ) ) with w = 0
  recRgb[tgt] = make_uchar4( roundClip( 0.0 * 0.0 + 1.0 * ( a0 / sw ) ), roundClip( 0.0 * 0.0 + 1.0 * ( a1 / sw ) ),
                             roundClip( 0.0 * 0.0 + 1.0 * ( a2 / sw ) ), 0 );
}

}  // namespace

void transferColors( ColorScratch& sc, const KdTree& srcTree, const short4* srcPts, const uchar4* srcRgb, size_t n, const short4* recPts, size_t R,
                     uchar4* recRgb, cudaStream_t s ) {
  if ( R == 0 || n == 0 ) return;
  // forward: 8-NN of every reconstructed point in the source cloud
  sc.fwdIdx.reserve( R * 8 ), sc.fwdDist.reserve( R * 8 );
  kdKnn( srcTree, recPts, R, nullptr, 8, sc.fwdIdx, sc.fwdDist, s );
  kForward<<<divUp( R, 128 ), 128, 0, s>>>( sc.fwdIdx, sc.fwdDist, srcRgb, int( R ), recRgb );
  // backward: nearest reconstructed point of every source point
  kdBuild( sc.recTree, recPts, R, s );
  sc.bwdIdx.reserve( n ), sc.bwdDist.reserve( n ), sc.keysA.reserve( n ), sc.keysB.reserve( n ), sc.idsA.reserve( n ), sc.idsB.reserve( n );
  kdKnn( sc.recTree, srcPts, n, srcTree.vind, 1, sc.bwdIdx, sc.bwdDist, s );
  kVoteKeys<<<divUp( n, 256 ), 256, 0, s>>>( sc.bwdIdx, sc.bwdDist, int( n ), sc.keysA, sc.idsA );
  int tbits = 1;
  while ( ( size_t( 1 ) << tbits ) < R ) ++tbits;
  size_t tmpBytes = 0;
  PCC_CUDA( cub::DeviceRadixSort::SortPairs( nullptr, tmpBytes, sc.keysA.p, sc.keysB.p, sc.idsA.p, sc.idsB.p, int( n ), 0, 24 + tbits, s ) );
  sc.cubTmp.reserve( tmpBytes + 16 );
  PCC_CUDA( cub::DeviceRadixSort::SortPairs( sc.cubTmp.p, tmpBytes, sc.keysA.p, sc.keysB.p, sc.idsA.p, sc.idsB.p, int( n ), 0, 24 + tbits, s ) );
  kBackward<<<divUp( n, 128 ), 128, 0, s>>>( sc.keysB, sc.idsB, int( n ), srcRgb, recRgb );
  PCC_LAUNCH_CHECK();
}

}  // namespace pccb200
