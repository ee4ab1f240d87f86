// postrecon.cuh — per-element arithmetic of the post-reconstruction chain, written once for host and device:
// the same functions are compiled by g++ into a CPU test (tests/test_postrecon_functions.py: against the reference itself) and
// by nvcc into the kernels of postrecon.cu. EXPERIMENTAL this round: the kernels around these functions have not run on a GPU
// yet, so nothing here is exported through include/pccb200.h.
//   * grid-based geometry smoothing, PCCCodec::smoothPointCloudPostprocess (PccLibCommon/source/PCCCodec.cpp:54-150),
//     gridFiltering (:1002-1065), smoothPointCloudGrid (:1067-1106)
//   * YUV 4:2:0 (8 bit) -> YUV 4:4:4 (16 bit), PCCInternalColorConverter "YUV420ToYUV444_8_0"
//     (PccLibColorConverter/source/PCCInternalColorConverter.cpp:467-485, 595-610, 669-695)
//   * PCCPointSet3::convertYUV16ToRGB8 (PccLibCommon/include/PCCPointSet.h:133-166)
#pragma once
#include <stdint.h>

#include <cmath>

#include "stdsort.cuh"

#if defined( __CUDACC__ )
#define PCC_HD __host__ __device__
#else
#define PCC_HD
#endif

namespace pccb200 {
namespace postrecon {

// ---- geometry smoothing -------------------------------------------------------------------------------------------------
// Dense cell grid of side w (cell = gridSize^3 voxels). Per cell: number of points, integer coordinate sums (the reference
// accumulates floats; the sums are integers below 2^24, so they are exact and order-free), smallest and largest patch index + 1
// (the reference's "another patch has points here" flag is: not all equal), and whether a boundary point touches it.
struct CellGrid {
  int             w, gridSize;
  const uint32_t* count;   // points per cell (the reference keeps uint16: wraps at 65536 like it)
  const int*      sum;     // 3 per cell
  const uint32_t* minPatch;
  const uint32_t* maxPatch;
  const uint8_t*  used;    // a boundary point's 2x2x2 neighbourhood covers this cell (cellIndex != -1 in the reference)
};
PCC_HD inline bool nearBorder( const int16_t* p, int gridSize, int w ) {
  const int disth = gridSize / 2 > 1 ? gridSize / 2 : 1, th = gridSize * w;
  return p[0] < disth || p[1] < disth || p[2] < disth || th <= p[0] + disth || th <= p[1] + disth || th <= p[2] + disth;
}
// the lower corner of the 2x2x2 cell neighbourhood of a point
PCC_HD inline void cornerCell( const int16_t* p, int gridSize, int S[3] ) {
  const int half = gridSize / 2;
  for ( int k = 0; k < 3; ++k ) S[k] = p[k] / gridSize + ( ( p[k] - ( p[k] / gridSize ) * gridSize ) < half ? -1 : 0 );
}
// gridFiltering + the decision of smoothPointCloudGrid for one boundary point (type 1, not near the border).
// Returns true and the new position when the point moves (its boundary type becomes 3).
PCC_HD inline bool smoothPoint( const int16_t* p, const CellGrid& g, double threshold, int16_t out[3] ) {
  const int w = g.w, gridSize = g.gridSize, half = gridSize / 2, g2 = gridSize * 2, w3 = w * w * w;
  int       S[3], idx[2][2][2];
  cornerCell( p, gridSize, S );
  bool other = false;
  for ( int dz = 0; dz < 2; ++dz )
    for ( int dy = 0; dy < 2; ++dy )
      for ( int dx = 0; dx < 2; ++dx ) {
        const int t     = ( S[0] + dx ) + ( S[1] + dy ) * w + ( S[2] + dz ) * w * w;
        idx[dz][dy][dx] = t;
        const uint16_t c = uint16_t( g.count[t] );
        if ( c != 0 && g.minPatch[t] != g.maxPatch[t] ) other = true;
      }
  if ( !other ) return false;
  const double cur[3] = {double( p[0] ), double( p[1] ), double( p[2] )};
  int          W[3], Q[3];
  for ( int k = 0; k < 3; ++k ) W[k] = ( p[k] - S[k] * gridSize - half ) * 2 + 1, Q[k] = g2 - W[k];
  double sum[3] = {0.0, 0.0, 0.0};
  int    cnt    = 0;
  for ( int dz = 0; dz < 2; ++dz )
    for ( int dy = 0; dy < 2; ++dy )
      for ( int dx = 0; dx < 2; ++dx ) {
        const int      t = idx[dz][dy][dx];
        const uint16_t c = uint16_t( g.count[t] );
        double         v[3] = {cur[0], cur[1], cur[2]};
        if ( ( ( dx == 0 && dy == 0 && dz == 0 ) || t < w3 ) && c > 0 )
          for ( int k = 0; k < 3; ++k ) v[k] = double( float( g.sum[3 * t + k] ) / float( c ) );  // the cell's float centroid
        const int wgt = ( dx ? W[0] : Q[0] ) * ( dy ? W[1] : Q[1] ) * ( dz ? W[2] : Q[2] );
        for ( int k = 0; k < 3; ++k ) {
          v[k] *= double( wgt );
          sum[k] += v[k];
        }
        cnt += wgt * int( c );
      }
  for ( int k = 0; k < 3; ++k ) sum[k] /= double( g2 * g2 * g2 );
  cnt /= g2 * g2 * g2;
  double centroid[3], d[3];
  for ( int k = 0; k < 3; ++k ) centroid[k] = sum[k] * double( cnt ), d[k] = cur[k] * double( cnt ) - centroid[k];
  const double dist2 = ( d[0] * d[0] + d[1] * d[1] + d[2] * d[2] ) / double( cnt ) + 0.5;
  const int    thr   = int( threshold ) > cnt ? int( threshold ) : cnt;
  if ( !( dist2 >= double( thr * 2 ) ) ) return false;  // (also false for the NaN of an empty neighbourhood)
  for ( int k = 0; k < 3; ++k ) out[k] = int16_t( double( int64_t( centroid[k] / double( cnt ) + 0.5 ) ) );
  return true;
}

// ---- colour conversions -------------------------------------------------------------------------------------------------
PCC_HD inline float yuv8ToFloat( uint8_t v, bool chroma ) {  // YUVtoFloatYUV, one byte per sample
  const float f = float( ( 1.0 / 255. ) * double( int( v ) - ( chroma ? 128 : 0 ) ) );
  const float lo = chroma ? -0.5f : 0.f, hi = chroma ? 0.5f : 1.f;
  return f < lo ? lo : ( f > hi ? hi : f );
}
PCC_HD inline uint16_t floatToYuv16( float v, bool chroma ) {  // floatYUVToYUV, two bytes per sample
  float r = roundf( float( 65535. * double( v ) + ( chroma ? 32768. : 0. ) ) );
  r       = r < 0.f ? 0.f : ( r > 65535.f ? 65535.f : r );
  return uint16_t( r );
}
PCC_HD inline int clampIndex( int v, int hi ) { return v < 0 ? 0 : ( v > hi ? hi : v ); }
// up-sampling filter 0 (UF_F0): vertical pass of one chroma plane (w2 x h2 floats in) -> rows 2i and 2i+1 of a w2 x 2*h2 plane
PCC_HD inline void upsampleVertical( const float* in, int w2, int h2, int i, int j, float& even, float& odd ) {
  const float ver0[4] = {-8.0f, +64.0f, +216.0f, -16.0f}, ver1[4] = {-16.0f, +216.0f, +64.0f, -8.0f};
  float       a = 0, b = 0;
  for ( int t = 0; t < 4; ++t ) a += ver0[t] * in[clampIndex( i + t - 2, h2 - 1 ) * w2 + j];
  for ( int t = 0; t < 4; ++t ) b += ver1[t] * in[clampIndex( i + 1 + t - 2, h2 - 1 ) * w2 + j];
  even = ( a + 0.f ) * ( 1.0f / 256.f ), odd = ( b + 0.f ) * ( 1.0f / 256.f );
}
// horizontal pass: samples 2j and 2j+1 of row i (tmp: w2 floats per row)
PCC_HD inline void upsampleHorizontal( const float* tmpRow, int w2, int j, float& even, float& odd ) {
  const float hor1[4] = {-16.0f, +144.0f, +144.0f, -16.0f};
  float       a = 0, b = 0;
  a += 0.0f * tmpRow[clampIndex( j - 1, w2 - 1 )];
  a += 256.0f * tmpRow[j];
  for ( int t = 0; t < 4; ++t ) b += hor1[t] * tmpRow[clampIndex( j + 1 + t - 2, w2 - 1 )];
  even = ( a + 0.f ) * ( 1.0f / 256.f ), odd = ( b + 0.f ) * ( 1.0f / 256.f );
}
PCC_HD inline void yuv16ToRgb8( const uint16_t yuv[3], uint8_t rgb[3] ) {
  const double wgt = 1.0 / 65535.0;
  double       y = wgt * double( yuv[0] ), u = wgt * ( double( yuv[1] ) - 32768.0 ), v = wgt * ( double( yuv[2] ) - 32768.0 );
  y = y < 0.0 ? 0.0 : ( y > 1.0 ? 1.0 : y ), u = u < -0.5 ? -0.5 : ( u > 0.5 ? 0.5 : u ), v = v < -0.5 ? -0.5 : ( v > 0.5 ? 0.5 : v );
  const double c[3] = {y + 1.57480 * v, y - 0.18733 * u - 0.46813 * v, y + 1.85563 * u};
  for ( int k = 0; k < 3; ++k ) {
    const double r = round( c[k] * 255 );
    rgb[k]         = uint8_t( r < 0.0 ? 0.0 : ( r > 255.0 ? 255.0 : r ) );
  }
}

// ---- colour transfer onto the smoothed cloud: PCCPointSet3::transferColors16bitBP with the arguments of encode / decode
// (PccLibCommon/source/PCCPointSet.cpp:1126-1485; PccLibEncoder/source/PCCEncoder.cpp:656-672), per target point of boundary type 3
// forward: the (up to 8) nearest source points of the target, nanoflann order, squared distances
PCC_HD inline void forwardColour( const uint32_t* id, const float* dist2, int cnt, const uint16_t* srcCol, uint16_t out[3] ) {
  if ( double( dist2[0] ) < 0.0001 || cnt == 1 ) {
    for ( int k = 0; k < 3; ++k ) out[k] = srcCol[3 * size_t( id[0] ) + k];
    return;
  }
  double acc[3] = {0.0, 0.0, 0.0}, sum = 0.0;
  for ( int i = 0; i < cnt; ++i ) {
    const double wgt = 1 / ( double( dist2[i] ) + 4.0 );
    for ( int k = 0; k < 3; ++k ) acc[k] += srcCol[3 * size_t( id[i] ) + k] * wgt;
    sum += wgt;
  }
  for ( int k = 0; k < 3; ++k ) {
    const double r = round( acc[k] / sum );
    out[k]         = uint16_t( r < 0.0 ? 0.0 : ( r > 65535.0 ? 65535.0 : r ) );
  }
}
// backward: the votes a target received (in sampling order), sorted as std::sort sorts them, distance-weighted mean
struct Vote {
  double   dist;  // squared distance source sample -> target
  uint16_t c[3];
};
struct VoteByDistance {
  PCC_HD bool operator()( const Vote& a, const Vote& b ) const { return a.dist < b.dist; }
};
PCC_HD inline void backwardColour( Vote* votes, int n, const uint16_t refined[3], uint16_t out[3] ) {
  if ( n == 0 ) {
    for ( int k = 0; k < 3; ++k ) out[k] = refined[k];
    return;
  }
  stdsort::sort( votes, votes + n, VoteByDistance() );
  double c2[3] = {0.0, 0.0, 0.0};
  if ( n == 1 ) {
    for ( int k = 0; k < 3; ++k ) c2[k] = votes[0].c[k];
  } else {
    double sum = 0.0;
    for ( int i = 0; i < n; ++i ) {
      const double wgt = 1 / ( sqrt( votes[i].dist ) + 4.0 );
      for ( int k = 0; k < 3; ++k ) c2[k] += ( votes[i].c[k] * wgt );
      sum += wgt;
    }
    for ( int k = 0; k < 3; ++k ) c2[k] /= sum;
  }
  for ( int k = 0; k < 3; ++k ) {
    double v = round( 0.0 * double( refined[k] ) + 1.0 * c2[k] );  // fixWeight: w = 0
    v        = v < 0.0 ? 0.0 : ( v > 65535.0 ? 65535.0 : v );
    out[k]   = uint16_t( v );
  }
}

}  // namespace postrecon
}  // namespace pccb200
