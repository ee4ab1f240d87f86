// TEST-ONLY: drives the host/device functions of mpeg-pcc-tmc2_b200/csrc/postrecon.cuh sequentially on the CPU, with the grid
// accumulation the kernels do with atomics written as a plain loop, so that tests/test_postrecon_functions.py can compare the
// arithmetic with the reference itself before the kernels ever run.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../mpeg-pcc-tmc2_b200/csrc/postrecon.cuh"

using namespace pccb200::postrecon;

extern "C" {

void prx_smooth_geometry( int16_t* xyz, uint16_t* boundary, const uint32_t* partition, size_t n, int gridSize, double threshold ) {
  if ( n == 0 ) return;
  int maxSize = 0;
  for ( size_t i = 0; i < 3 * n; ++i ) maxSize = std::max( maxSize, int( xyz[i] ) );
  maxSize = std::max( maxSize, std::max( int( xyz[0] ), std::max( int( xyz[1] ), int( xyz[2] ) ) ) );
  const int    w = ( maxSize + gridSize - 1 ) / gridSize;
  const size_t cells = size_t( w ) * w * w;
  std::vector<uint32_t> count( cells, 0 ), minPatch( cells, 0xFFFFFFFFu ), maxPatch( cells, 0 );
  std::vector<int>      sum( 3 * cells, 0 );
  std::vector<uint8_t>  used( cells, 0 );
  for ( size_t i = 0; i < n; ++i ) {  // kernel 1: mark the cells around boundary points
    const int16_t* p = xyz + 3 * i;
    if ( boundary[i] != 1 || nearBorder( p, gridSize, w ) ) continue;
    int S[3];
    cornerCell( p, gridSize, S );
    for ( int d = 0; d < 8; ++d ) used[size_t( S[0] + ( d & 1 ) ) + size_t( S[1] + ( ( d >> 1 ) & 1 ) ) * w + size_t( S[2] + ( d >> 2 ) ) * w * w] = 1;
  }
  for ( size_t i = 0; i < n; ++i ) {  // kernel 2: accumulate every point into its cell (atomics on the device)
    const int16_t* p = xyz + 3 * i;
    if ( nearBorder( p, gridSize, w ) ) continue;
    const size_t c = size_t( p[0] / gridSize ) + size_t( p[1] / gridSize ) * w + size_t( p[2] / gridSize ) * w * w;
    if ( !used[c] ) continue;
    ++count[c];
    for ( int k = 0; k < 3; ++k ) sum[3 * c + k] += p[k];
    minPatch[c] = std::min( minPatch[c], partition[i] + 1 ), maxPatch[c] = std::max( maxPatch[c], partition[i] + 1 );
  }
  CellGrid g{w, gridSize, count.data(), sum.data(), minPatch.data(), maxPatch.data(), used.data()};
  std::vector<int16_t>  moved( 3 * n );
  std::vector<uint8_t>  flag( n, 0 );
  for ( size_t i = 0; i < n; ++i ) {  // kernel 3: one thread per point (reads the frozen grid and the ORIGINAL position only)
    const int16_t* p = xyz + 3 * i;
    if ( boundary[i] != 1 || nearBorder( p, gridSize, w ) ) continue;
    flag[i] = smoothPoint( p, g, threshold, &moved[3 * i] ) ? 1 : 0;
  }
  for ( size_t i = 0; i < n; ++i )
    if ( flag[i] ) {
      std::memcpy( xyz + 3 * i, &moved[3 * i], 6 );
      boundary[i] = 3;
    }
}

void prx_yuv420_to_yuv444_16( const uint8_t* yuv420, size_t W, size_t H, uint16_t* yuv444 ) {
  const size_t Q = W * H;
  const int    w2 = int( W / 2 ), h2 = int( H / 2 );
  for ( size_t i = 0; i < Q; ++i ) yuv444[i] = floatToYuv16( yuv8ToFloat( yuv420[i], false ), false );
  for ( int c = 0; c < 2; ++c ) {
    const uint8_t*     src = yuv420 + Q + size_t( c ) * w2 * h2;
    std::vector<float> in( size_t( w2 ) * h2 ), tmp( size_t( w2 ) * H );
    for ( size_t i = 0; i < in.size(); ++i ) in[i] = yuv8ToFloat( src[i], true );
    for ( int i = 0; i < h2; ++i )
      for ( int j = 0; j < w2; ++j ) upsampleVertical( in.data(), w2, h2, i, j, tmp[size_t( 2 * i ) * w2 + j], tmp[size_t( 2 * i + 1 ) * w2 + j] );
    uint16_t* dst = yuv444 + Q * ( 1 + c );
    for ( size_t i = 0; i < H; ++i )
      for ( int j = 0; j < w2; ++j ) {
        float e, o;
        upsampleHorizontal( tmp.data() + i * w2, w2, j, e, o );
        dst[i * W + 2 * j] = floatToYuv16( e, true ), dst[i * W + 2 * j + 1] = floatToYuv16( o, true );
      }
  }
}

void prx_yuv16_to_rgb8( const uint16_t* yuv, size_t n, uint8_t* rgb ) {
  for ( size_t i = 0; i < n; ++i ) yuv16ToRgb8( yuv + 3 * i, rgb + 3 * i );
}

// per target t of `count` targets: forward colour from its neighbour row (k entries, 0xFFFFFFFF = none)
void prx_forward( const uint32_t* idx, const float* dist2, int k, size_t count, const uint16_t* srcCol, uint16_t* out ) {
  for ( size_t t = 0; t < count; ++t ) {
    int cnt = 0;
    while ( cnt < k && idx[t * k + cnt] != 0xFFFFFFFFu ) ++cnt;
    forwardColour( idx + t * k, dist2 + t * k, cnt, srcCol, out + 3 * t );
  }
}
// votes in CSR form (offsets per target, entries in sampling order): (dist, colour); refined = forward colours; out = final colours
void prx_backward( const uint64_t* offsets, const double* voteDist, const uint16_t* voteCol, size_t count, const uint16_t* refined, uint16_t* out ) {
  std::vector<Vote> v;
  for ( size_t t = 0; t < count; ++t ) {
    v.clear();
    for ( uint64_t e = offsets[t]; e < offsets[t + 1]; ++e ) v.push_back( Vote{voteDist[e], {voteCol[3 * e], voteCol[3 * e + 1], voteCol[3 * e + 2]}} );
    backwardColour( v.data(), int( v.size() ), refined + 3 * t, out + 3 * t );
  }
}
}
