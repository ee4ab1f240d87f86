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
// order, which is what the reference's insertion sort (std::sort on <= 16 elements) produces. A target with MORE than 16 votes
// goes through libstdc++'s introsort in the reference (PCCPointSet.cpp:955), which permutes ties: those runs are re-ordered by
// the step-by-step emulation of stdsort.cuh before the fp64 accumulation, so the sums stay bit-exact for any run length.
#include <cub/device/device_radix_sort.cuh>

#include "stages.cuh"
#include "stdsort.cuh"

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

constexpr int      kDistBits = 26;
constexpr uint32_t kDistMask = ( 1u << kDistBits ) - 1u;
struct VoteLessByDistance {  // the reference's comparator: distance only (ties are where std::sort's algorithm shows)
  __host__ __device__ bool operator()( const uint64_t& a, const uint64_t& b ) const { return ( a >> 32 ) < ( b >> 32 ); }
};

__global__ void kVoteKeys( const uint32_t* __restrict__ target, const float* __restrict__ dist, int n, uint64_t* __restrict__ keys,
                           uint32_t* __restrict__ ids ) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if ( s >= n ) return;
  // squared distances are exact integers below 2^26 (12-bit coordinates: 3 * 4095^2)
  keys[s] = ( uint64_t( target[s] ) << kDistBits ) | uint64_t( uint32_t( dist[s] ) & kDistMask );
  ids[s]  = s;
}

// one thread per vote position; the head of each target's run accumulates the whole run in order.
// `scratch` (the radix sort's input keys, dead after the sort) gives a run of more than 16 votes a private work area at its own
// positions [p, e): the votes are put back into arrival order (ascending source index, as the reference pushes them) and sorted
// by the std::sort emulation.
__global__ void __launch_bounds__( 128 )
    kBackward( const uint64_t* __restrict__ keys, const uint32_t* __restrict__ srcIds, int n, const uchar4* __restrict__ srcRgb,
               uchar4* __restrict__ recRgb, uint64_t* __restrict__ scratch ) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if ( p >= n ) return;
  const uint32_t tgt = uint32_t( keys[p] >> kDistBits );
  if ( p > 0 && uint32_t( keys[p - 1] >> kDistBits ) == tgt ) return;
  int e = p + 1;
  while ( e < n && uint32_t( keys[e] >> kDistBits ) == tgt ) ++e;
  const uchar4 first = srcRgb[srcIds[p]];
  if ( ( keys[p] & kDistMask ) == 0 || e - p == 1 ) {
    recRgb[tgt] = first;  // a source point at this very position, or a single vote: round(1.0 * c) == c
    return;
  }
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, sw = 0.0;
  if ( e - p <= 16 ) {
    for ( int j = p; j < e; ++j ) {
      const double w = 1 / ( sqrt( double( uint32_t( keys[j] & kDistMask ) ) ) + 4.0 );
      const uchar4 c = srcRgb[srcIds[j]];
      a0 += ( double( c.x ) * w ), a1 += ( double( c.y ) * w ), a2 += ( double( c.z ) * w );
      sw += w;
    }
  } else {
    uint64_t* v   = scratch + p;
    const int len = e - p;
    for ( int j = 0; j < len; ++j ) {  // (distance << 32 | source index), inserted in ascending source index
      const uint64_t item = ( uint64_t( uint32_t( keys[p + j] & kDistMask ) ) << 32 ) | srcIds[p + j];
      int            b    = j;
      while ( b > 0 && uint32_t( v[b - 1] ) > uint32_t( item ) ) v[b] = v[b - 1], --b;
      v[b] = item;
    }
    stdsort::sort( v, v + len, VoteLessByDistance() );
    for ( int j = 0; j < len; ++j ) {
      const double w = 1 / ( sqrt( double( uint32_t( v[j] >> 32 ) ) ) + 4.0 );
      const uchar4 c = srcRgb[uint32_t( v[j] )];
      a0 += ( double( c.x ) * w ), a1 += ( double( c.y ) * w ), a2 += ( double( c.z ) * w );
      sw += w;
    }
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
  PCC_CUDA( cub::DeviceRadixSort::SortPairs( nullptr, tmpBytes, sc.keysA.p, sc.keysB.p, sc.idsA.p, sc.idsB.p, int( n ), 0, kDistBits + tbits, s ) );
  sc.cubTmp.reserve( tmpBytes + 16 );
  PCC_CUDA( cub::DeviceRadixSort::SortPairs( sc.cubTmp.p, tmpBytes, sc.keysA.p, sc.keysB.p, sc.idsA.p, sc.idsB.p, int( n ), 0, kDistBits + tbits, s ) );
  kBackward<<<divUp( n, 128 ), 128, 0, s>>>( sc.keysB, sc.idsB, int( n ), srcRgb, recRgb, sc.keysA );
  PCC_LAUNCH_CHECK();
}

// ---- §8f-2: RGB444 -> YUV 4:2:0 (8 bit) of a padded attribute frame, as PCCVideoEncoder::compress converts it before the
// codec (PccLibEncoder/source/PCCVideoEncoder.cpp:326-353 -> PCCInternalColorConverter "RGB444ToYUV420_8_4":
// PccLibColorConverter/source/PCCInternalColorConverter.cpp:407-424, 553-593, 642-666, filters
// PccLibColorConverter/include/PCCInternalColorConverter.h:153-183). Doing it here keeps the frame on the device until it has
// its final 1.5 bytes per pixel instead of 6. Arithmetic as the reference: float samples, double accumulation in tap order, one
// rounding to float per stage (no fused multiply-add: the library is built with -fmad=false).
namespace {
// down-sampling filter 4 (DF_GS), the default of PCCVideoEncoder::compress: the reference's constants, float( c * 512 ), shift 9
__constant__ float cGsHor[15] = {float( -0.01716352771649 * 512 ), float( 0.0 ), float( +0.04066666714886 * 512 ), float( 0.0 ),
                                 float( -0.09154810319329 * 512 ), float( 0.0 ), float( 0.31577823859943 * 512 ), float( 0.50453345032298 * 512 ),
                                 float( 0.31577823859943 * 512 ), float( 0.0 ), float( -0.09154810319329 * 512 ), float( 0.0 ),
                                 float( 0.04066666714886 * 512 ), float( 0.0 ), float( -0.01716352771649 * 512 )};
__constant__ float cGsVer[16] = {float( -0.00945406160902 * 512 ), float( -0.01539537217249 * 512 ), float( 0.02360533018213 * 512 ),
                                 float( 0.03519540819902 * 512 ), float( -0.05254456550808 * 512 ), float( -0.08189331229717 * 512 ),
                                 float( 0.14630826357715 * 512 ), float( 0.45417830962846 * 512 ), float( 0.45417830962846 * 512 ),
                                 float( 0.14630826357715 * 512 ), float( -0.08189331229717 * 512 ), float( -0.05254456550808 * 512 ),
                                 float( 0.03519540819902 * 512 ), float( 0.02360533018213 * 512 ), float( -0.01539537217249 * 512 ),
                                 float( -0.00945406160902 * 512 )};

__device__ __forceinline__ uint8_t quantise8( float v, bool chroma ) {  // floatYUVToYUV, one byte per sample
  float r = roundf( float( 255. * double( v ) + ( chroma ? 128. : 0. ) ) );
  r       = fminf( fmaxf( r, 0.f ), 255.f );
  return uint8_t( r );
}

__global__ void kRgbToYuv444( const uint16_t* __restrict__ rgb, size_t Q, uint8_t* __restrict__ outY, float* __restrict__ U, float* __restrict__ V ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i >= Q ) return;
  const float  r = float( rgb[i] ) / 255.f, g = float( rgb[Q + i] ) / 255.f, b = float( rgb[2 * Q + i] ) / 255.f;
  const double y = 0.212600 * r + 0.715200 * g + 0.072200 * b, u = -0.114572 * r - 0.385428 * g + 0.500000 * b,
               v = 0.500000 * r - 0.454153 * g - 0.045847 * b;
  outY[i] = quantise8( float( y < 0.0 ? 0.0 : ( y > 1.0 ? 1.0 : y ) ), false );
  U[i]    = float( u < -0.5 ? -0.5 : ( u > 0.5 ? 0.5 : u ) );
  V[i]    = float( v < -0.5 ? -0.5 : ( v > 0.5 ? 0.5 : v ) );
}
// horizontal pass of both chroma planes (blockIdx.z selects the plane): W x H -> W/2 x H
__global__ void kChromaDownH( const float* __restrict__ U, const float* __restrict__ V, int W, int H, float* __restrict__ tU, float* __restrict__ tV ) {
  const int    w2 = W / 2;
  const size_t t  = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( t >= size_t( w2 ) * H ) return;
  const float* in  = blockIdx.z ? V : U;
  float*       out = blockIdx.z ? tV : tU;
  const int    i = int( t / w2 ), j = int( t % w2 );
  double       acc = 0;
#pragma unroll
  for ( int k = 0; k < 15; ++k ) {
    const int x = min( max( 2 * j + k - 7, 0 ), W - 1 );
    acc         = acc + double( cGsHor[k] ) * double( in[size_t( i ) * W + x] );
  }
  out[t] = float( ( acc + 0.0 ) * double( 1.0f / 512.f ) );
}
// vertical pass + quantisation: W/2 x H -> W/2 x H/2 bytes
__global__ void kChromaDownV( const float* __restrict__ tU, const float* __restrict__ tV, int w2, int H, uint8_t* __restrict__ outU, uint8_t* __restrict__ outV ) {
  const int    h2 = H / 2;
  const size_t t  = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( t >= size_t( w2 ) * h2 ) return;
  const float* in  = blockIdx.z ? tV : tU;
  uint8_t*     out = blockIdx.z ? outV : outU;
  const int    i = int( t / w2 ), j = int( t % w2 );
  double       acc = 0;
#pragma unroll
  for ( int k = 0; k < 16; ++k ) {
    const int y = min( max( 2 * i + k - 7, 0 ), H - 1 );
    acc         = acc + double( cGsVer[k] ) * double( in[size_t( y ) * w2 + j] );
  }
  out[t] = quantise8( float( ( acc + 0.0 ) * double( 1.0f / 512.f ) ), true );
}
}  // namespace

void rgbPlanesToYuv420( const uint16_t* rgbPlanes, int W, int H, YuvScratch& sc, uint8_t* out, cudaStream_t s ) {
  const size_t Q = size_t( W ) * H, q4 = size_t( W / 2 ) * ( H / 2 );
  sc.U.reserve( Q ), sc.V.reserve( Q ), sc.tU.reserve( Q / 2 + 1 ), sc.tV.reserve( Q / 2 + 1 );
  kRgbToYuv444<<<divUp( Q, 256 ), 256, 0, s>>>( rgbPlanes, Q, out, sc.U, sc.V );
  kChromaDownH<<<dim3( divUp( size_t( W / 2 ) * H, 256 ), 1, 2 ), 256, 0, s>>>( sc.U, sc.V, W, H, sc.tU, sc.tV );
  kChromaDownV<<<dim3( divUp( q4, 256 ), 1, 2 ), 256, 0, s>>>( sc.tU, sc.tV, W / 2, H, out + Q, out + Q + q4 );
  PCC_LAUNCH_CHECK();
}

}  // namespace pccb200
