// postrecon.cu — the post-reconstruction chain that follows generatePointCloud in PCCEncoder::encode (PCCEncoder.cpp:556-719) and
// PCCDecoder::decode (PCCDecoder.cpp:356-475) under the CTC: grid-based geometry smoothing
// (PCCCodec::smoothPointCloudPostprocess, PCCCodec.cpp:54-150, 982-1168), the YUV420 -> YUV444(16) -> RGB8 conversions around
// colorPointCloud (PCCCodec.cpp:1319-1460) and the colour transfer onto the smoothed cloud (PCCPointSet3::transferColors16bitBP,
// PCCPointSet.cpp:1126-1485). Kernels around the host/device functions of postrecon.cuh (whose arithmetic is pinned against the
// reference on the CPU: tests/test_postrecon_functions.py); GPU parity vs the oracle: tests/test_gpu_postrecon.py.
// Entry points are declared in include/pccb200.h.
#include <cub/device/device_radix_sort.cuh>

#include "postrecon.cuh"
#include "stages.cuh"

using namespace pccb200;

#include "ctx.cuh"

namespace {

using namespace pccb200::postrecon;

__global__ void kMaxI16( const int16_t* __restrict__ v, size_t n, int* __restrict__ out ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  int          m = i < n ? int( v[i] ) : INT_MIN;
  m              = __reduce_max_sync( 0xffffffffu, m );
  if ( ( threadIdx.x & 31 ) == 0 ) atomicMax( out, m );
}
__global__ void kMarkCells( const int16_t* __restrict__ xyz, const uint16_t* __restrict__ boundary, size_t n, int gridSize, int w, uint8_t* __restrict__ used ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i >= n || boundary[i] != 1 ) return;
  const int16_t* p = xyz + 3 * i;
  if ( nearBorder( p, gridSize, w ) ) return;
  int S[3];
  cornerCell( p, gridSize, S );
  for ( int d = 0; d < 8; ++d ) used[size_t( S[0] + ( d & 1 ) ) + size_t( S[1] + ( ( d >> 1 ) & 1 ) ) * w + size_t( S[2] + ( d >> 2 ) ) * w * w] = 1;
}
__global__ void kAccumulateCells( const int16_t* __restrict__ xyz, const uint32_t* __restrict__ partition, size_t n, int gridSize, int w,
                                  const uint8_t* __restrict__ used, uint32_t* __restrict__ count, int* __restrict__ sum, uint32_t* __restrict__ minPatch,
                                  uint32_t* __restrict__ maxPatch ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  const int16_t* p = xyz + 3 * i;
  if ( nearBorder( p, gridSize, w ) ) return;
  const size_t c = size_t( p[0] / gridSize ) + size_t( p[1] / gridSize ) * w + size_t( p[2] / gridSize ) * w * w;
  if ( !used[c] ) return;
  atomicAdd( &count[c], 1u );
  atomicAdd( &sum[3 * c], int( p[0] ) ), atomicAdd( &sum[3 * c + 1], int( p[1] ) ), atomicAdd( &sum[3 * c + 2], int( p[2] ) );
  atomicMin( &minPatch[c], partition[i] + 1 ), atomicMax( &maxPatch[c], partition[i] + 1 );
}
__global__ void kSmoothPoints( const int16_t* __restrict__ xyz, const uint16_t* __restrict__ boundary, size_t n, CellGrid g, double threshold,
                               int16_t* __restrict__ outXyz, uint16_t* __restrict__ outBoundary ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  const int16_t* p = xyz + 3 * i;
  int16_t        q[3] = {p[0], p[1], p[2]};
  uint16_t       b = boundary[i];
  if ( b == 1 && !nearBorder( p, g.gridSize, g.w ) && smoothPoint( p, g, threshold, q ) ) b = 3;
  outXyz[3 * i] = q[0], outXyz[3 * i + 1] = q[1], outXyz[3 * i + 2] = q[2];
  outBoundary[i] = b;
}

__global__ void kLuma16( const uint8_t* __restrict__ y, size_t Q, uint16_t* __restrict__ out ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i < Q ) out[i] = floatToYuv16( yuv8ToFloat( y[i], false ), false );
}
__global__ void kChromaToFloat( const uint8_t* __restrict__ c, size_t n, float* __restrict__ out ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i < n ) out[i] = yuv8ToFloat( c[i], true );
}
__global__ void kChromaUpV( const float* __restrict__ in, int w2, int h2, float* __restrict__ tmp ) {
  const size_t t = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( t >= size_t( w2 ) * h2 ) return;
  const int i = int( t / w2 ), j = int( t % w2 );
  upsampleVertical( in, w2, h2, i, j, tmp[size_t( 2 * i ) * w2 + j], tmp[size_t( 2 * i + 1 ) * w2 + j] );
}
__global__ void kChromaUpH( const float* __restrict__ tmp, int w2, int H, uint16_t* __restrict__ out ) {
  const size_t t = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( t >= size_t( w2 ) * H ) return;
  const int i = int( t / w2 ), j = int( t % w2 );
  float     e, o;
  upsampleHorizontal( tmp + size_t( i ) * w2, w2, j, e, o );
  out[size_t( i ) * ( 2 * w2 ) + 2 * j] = floatToYuv16( e, true ), out[size_t( i ) * ( 2 * w2 ) + 2 * j + 1] = floatToYuv16( o, true );
}
__global__ void kYuv16ToRgb8( const uint16_t* __restrict__ yuv, size_t n, uint8_t* __restrict__ rgb ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  const uint16_t in[3] = {yuv[3 * i], yuv[3 * i + 1], yuv[3 * i + 2]};
  uint8_t        out[3];
  yuv16ToRgb8( in, out );
  rgb[3 * i] = out[0], rgb[3 * i + 1] = out[1], rgb[3 * i + 2] = out[2];
}

// ---- colour transfer onto the smoothed cloud (transferColors16bitBP as encode / decode call it) ---------------------------------
__global__ void kFlagMoved( const uint16_t* __restrict__ boundary, size_t n, uint32_t* __restrict__ flag ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i < n ) flag[i] = boundary[i] == 3 ? 1u : 0u;
}
__global__ void kListMoved( const uint32_t* __restrict__ flag, const uint32_t* __restrict__ scan, size_t n, const short4* __restrict__ tgt4,
                            uint32_t* __restrict__ moved, short4* __restrict__ queries ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i < n && flag[i] ) moved[scan[i]] = uint32_t( i ), queries[scan[i]] = tgt4[i];
}
__global__ void kForwardColours( const uint32_t* __restrict__ fidx, const float* __restrict__ fdist, size_t M, const uint16_t* __restrict__ srcCol,
                                 uint16_t* __restrict__ refined ) {
  const size_t m = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( m >= M ) return;
  uint32_t id[8];
  float    d[8];
  int      cnt = 0;
  for ( int i = 0; i < 8; ++i ) {
    id[i] = fidx[8 * m + i], d[i] = fdist[8 * m + i];
    if ( id[i] != 0xFFFFFFFFu && cnt == i ) ++cnt;
  }
  uint16_t out[3];
  forwardColour( id, d, cnt, srcCol, out );
  refined[3 * m] = out[0], refined[3 * m + 1] = out[1], refined[3 * m + 2] = out[2];
}
// the sampled source points (row-major over the moved targets' neighbour rows: the reference's sampling order) as 1-NN queries
__global__ void kSampleQueries( const uint32_t* __restrict__ fidx, size_t E, const short4* __restrict__ src4, short4* __restrict__ q ) {
  const size_t e = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( e < E ) q[e] = src4[fidx[e] != 0xFFFFFFFFu ? fidx[e] : 0u];
}
__global__ void kVoteKeys16( const uint32_t* __restrict__ fidx, const uint32_t* __restrict__ bidx, size_t E, const uint16_t* __restrict__ srcCol,
                             const uint16_t* __restrict__ tgtCol, const uint16_t* __restrict__ tgtBoundary, uint32_t* __restrict__ keys,
                             uint32_t* __restrict__ vals ) {
  const size_t e = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( e >= E ) return;
  uint32_t key = 0xFFFFFFFFu;
  const uint32_t si = fidx[e];
  if ( si != 0xFFFFFFFFu ) {
    const uint32_t t = bidx[e];
    if ( t != 0xFFFFFFFFu && tgtBoundary[t] == 3 ) {  // (votes for targets that keep their colour are never read)
      bool close = true;
      for ( int k = 0; k < 3; ++k ) close = close && abs( int( srcCol[3 * size_t( si ) + k] ) - int( tgtCol[3 * size_t( t ) + k] ) ) < 40;
      if ( close ) key = t;
    }
  }
  keys[e] = key, vals[e] = uint32_t( e );
}
__device__ __forceinline__ size_t lowerBound( const uint32_t* a, size_t n, uint32_t v ) {
  size_t lo = 0, hi = n;
  while ( lo < hi ) {
    const size_t mid = ( lo + hi ) / 2;
    if ( a[mid] < v ) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
__global__ void kBackwardColours( const uint32_t* __restrict__ moved, size_t M, const uint32_t* __restrict__ sortedKeys, const uint32_t* __restrict__ sortedVals,
                                  size_t E, const uint32_t* __restrict__ fidx, const float* __restrict__ bdist, const uint16_t* __restrict__ srcCol,
                                  const uint16_t* __restrict__ refined, Vote* __restrict__ votes, uint16_t* __restrict__ outCol ) {
  const size_t m = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( m >= M ) return;
  const uint32_t t = moved[m];
  const size_t   a = lowerBound( sortedKeys, E, t ), b = lowerBound( sortedKeys, E, t + 1 );
  for ( size_t j = a; j < b; ++j ) {  // radix sort is stable: the votes of a target are in sampling order
    const uint32_t e = sortedVals[j], si = fidx[e];
    votes[j]         = Vote{double( bdist[e] ), {srcCol[3 * size_t( si )], srcCol[3 * size_t( si ) + 1], srcCol[3 * size_t( si ) + 2]}};
  }
  uint16_t out[3];
  const uint16_t ref[3] = {refined[3 * m], refined[3 * m + 1], refined[3 * m + 2]};
  backwardColour( votes + a, int( b - a ), ref, out );
  outCol[3 * size_t( t )] = out[0], outCol[3 * size_t( t ) + 1] = out[1], outCol[3 * size_t( t ) + 2] = out[2];
}

}  // namespace

extern "C" {

// PCCPointSet3::transferColors16bitBP as encode / decode call it: new 16-bit colours (in place) for the target points of boundary type 3
int pccb200_transfer_colors16_smoothed( pccb200_ctx* ctx, const int16_t* srcXyz, const uint16_t* srcCol, size_t S, const int16_t* tgtXyz,
                                         uint16_t* tgtCol, const uint16_t* tgtBoundary, size_t T ) {
  return guarded( ctx, [&]() -> int {
    if ( !srcXyz || !srcCol || !tgtXyz || !tgtCol || !tgtBoundary ) return PCCB200_ERR_BAD_ARG;
    if ( S == 0 || T == 0 ) return PCCB200_OK;
    cudaStream_t     s = ctx->stream;
    DevBuf<int16_t>  raw;
    DevBuf<short4>   src4, tgt4, q4, sq4;
    DevBuf<uint16_t> dSrcCol, dTgtCol, dOutCol, dB, dRefined;
    DevBuf<uint32_t> flag, scan, scanTmp, moved, fidx, bidx, keysA, keysB, valsA, valsB;
    DevBuf<float>    fdist, bdist;
    DevBuf<uint8_t>  cubTmp;
    DevBuf<Vote>     votes;
    KdTree           treeS, treeT;
    raw.reserve( 3 * std::max( S, T ) ), src4.reserve( S + 1 ), tgt4.reserve( T + 1 );
    PCC_CUDA( cudaMemcpyAsync( raw, srcXyz, 3 * S * 2, cudaMemcpyHostToDevice, s ) );
    packXyz( raw, S, src4, s );
    streamWait( s );
    PCC_CUDA( cudaMemcpyAsync( raw, tgtXyz, 3 * T * 2, cudaMemcpyHostToDevice, s ) );
    packXyz( raw, T, tgt4, s );
    dSrcCol.reserve( 3 * S ), dTgtCol.reserve( 3 * T ), dOutCol.reserve( 3 * T ), dB.reserve( T );
    PCC_CUDA( cudaMemcpyAsync( dSrcCol, srcCol, 3 * S * 2, cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( dTgtCol, tgtCol, 3 * T * 2, cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( dOutCol, tgtCol, 3 * T * 2, cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( dB, tgtBoundary, T * 2, cudaMemcpyHostToDevice, s ) );
    kdBuild( treeS, src4, S, s );
    kdBuild( treeT, tgt4, T, s );
    // the targets the smoothing moved, in index order
    flag.reserve( T + 1 ), scan.reserve( T + 2 ), scanTmp.reserve( scanTmpElems( T ) );
    kFlagMoved<<<divUp( T, 256 ), 256, 0, s>>>( dB, T, flag );
    exclusiveScanU32( flag, scan, T, scanTmp, s );
    uint32_t M = 0;
    PCC_CUDA( cudaMemcpyAsync( &M, scan.p + T, sizeof( uint32_t ), cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    if ( M == 0 ) return PCCB200_OK;
    const size_t E = size_t( M ) * 8;
    moved.reserve( M ), q4.reserve( M ), fidx.reserve( E ), fdist.reserve( E ), dRefined.reserve( 3 * size_t( M ) );
    kListMoved<<<divUp( T, 256 ), 256, 0, s>>>( flag, scan, T, tgt4, moved, q4 );
    kdKnn( treeS, q4, M, nullptr, 8, fidx, fdist, s );
    kForwardColours<<<divUp( M, 128 ), 128, 0, s>>>( fidx, fdist, M, dSrcCol, dRefined );
    // backward votes
    sq4.reserve( E ), bidx.reserve( E ), bdist.reserve( E ), keysA.reserve( E ), keysB.reserve( E ), valsA.reserve( E ), valsB.reserve( E ), votes.reserve( E );
    kSampleQueries<<<divUp( E, 256 ), 256, 0, s>>>( fidx, E, src4, sq4 );
    kdKnn( treeT, sq4, E, nullptr, 1, bidx, bdist, s );
    kVoteKeys16<<<divUp( E, 256 ), 256, 0, s>>>( fidx, bidx, E, dSrcCol, dTgtCol, dB, keysA, valsA );
    size_t tmpBytes = 0;
    PCC_CUDA( cub::DeviceRadixSort::SortPairs( nullptr, tmpBytes, keysA.p, keysB.p, valsA.p, valsB.p, int( E ), 0, 32, s ) );
    cubTmp.reserve( tmpBytes + 16 );
    PCC_CUDA( cub::DeviceRadixSort::SortPairs( cubTmp.p, tmpBytes, keysA.p, keysB.p, valsA.p, valsB.p, int( E ), 0, 32, s ) );
    kBackwardColours<<<divUp( M, 128 ), 128, 0, s>>>( moved, M, keysB, valsB, E, fidx, bdist, dSrcCol, dRefined, votes, dOutCol );
    PCC_LAUNCH_CHECK();
    PCC_CUDA( cudaMemcpyAsync( tgtCol, dOutCol, 3 * T * 2, cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    return PCCB200_OK;
  } );
}

// PCCCodec::smoothPointCloudPostprocess (grid smoothing), in place on host arrays: positions n x 3, boundary point types, patch index
int pccb200_smooth_geometry( pccb200_ctx* ctx, int16_t* xyz, uint16_t* boundary, const uint32_t* partition, size_t n, int gridSize, double threshold ) {
  return guarded( ctx, [&]() -> int {
    if ( !xyz || !boundary || !partition || gridSize < 2 ) return PCCB200_ERR_BAD_ARG;
    if ( n == 0 ) return PCCB200_OK;
    cudaStream_t     s = ctx->stream;
    DevBuf<int16_t>  dXyz, dOutXyz;
    DevBuf<uint16_t> dB, dOutB;
    DevBuf<uint32_t> dPart, dCount, dMin, dMax;
    DevBuf<int>      dSum, dMaxCoord;
    DevBuf<uint8_t>  dUsed;
    dXyz.reserve( 3 * n ), dOutXyz.reserve( 3 * n ), dB.reserve( n ), dOutB.reserve( n ), dPart.reserve( n ), dMaxCoord.reserve( 1 );
    PCC_CUDA( cudaMemcpyAsync( dXyz, xyz, 3 * n * 2, cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( dB, boundary, n * 2, cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( dPart, partition, n * 4, cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemsetAsync( dMaxCoord, 0, sizeof( int ), s ) );
    kMaxI16<<<divUp( 3 * n, 256 ), 256, 0, s>>>( dXyz, 3 * n, dMaxCoord );
    int maxSize = 0;
    PCC_CUDA( cudaMemcpyAsync( &maxSize, dMaxCoord, sizeof( int ), cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    const int    w     = ( maxSize + gridSize - 1 ) / gridSize;
    const size_t cells = size_t( w ) * w * w;
    if ( cells == 0 ) return PCCB200_OK;
    dCount.reserve( cells ), dMin.reserve( cells ), dMax.reserve( cells ), dSum.reserve( 3 * cells ), dUsed.reserve( cells );
    PCC_CUDA( cudaMemsetAsync( dCount, 0, cells * 4, s ) );
    PCC_CUDA( cudaMemsetAsync( dMin, 0xff, cells * 4, s ) );
    PCC_CUDA( cudaMemsetAsync( dMax, 0, cells * 4, s ) );
    PCC_CUDA( cudaMemsetAsync( dSum, 0, 3 * cells * 4, s ) );
    PCC_CUDA( cudaMemsetAsync( dUsed, 0, cells, s ) );
    kMarkCells<<<divUp( n, 256 ), 256, 0, s>>>( dXyz, dB, n, gridSize, w, dUsed );
    kAccumulateCells<<<divUp( n, 256 ), 256, 0, s>>>( dXyz, dPart, n, gridSize, w, dUsed, dCount, dSum, dMin, dMax );
    CellGrid g{w, gridSize, dCount, dSum, dMin, dMax, dUsed};
    kSmoothPoints<<<divUp( n, 256 ), 256, 0, s>>>( dXyz, dB, n, g, threshold, dOutXyz, dOutB );
    PCC_LAUNCH_CHECK();
    PCC_CUDA( cudaMemcpyAsync( xyz, dOutXyz, 3 * n * 2, cudaMemcpyDeviceToHost, s ) );
    PCC_CUDA( cudaMemcpyAsync( boundary, dOutB, n * 2, cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    return PCCB200_OK;
  } );
}

// PCCInternalColorConverter "YUV420ToYUV444_8_0": Y (W*H), U, V ((W/2)*(H/2)) bytes -> three W*H planes of uint16
int pccb200_yuv420_to_yuv444_16( pccb200_ctx* ctx, const uint8_t* yuv420, size_t W, size_t H, uint16_t* yuv444 ) {
  return guarded( ctx, [&]() -> int {
    if ( !yuv420 || !yuv444 || W % 2 || H % 2 || W == 0 || H == 0 ) return PCCB200_ERR_BAD_ARG;
    cudaStream_t     s = ctx->stream;
    const size_t     Q = W * H, q4 = ( W / 2 ) * ( H / 2 );
    const int        w2 = int( W / 2 ), h2 = int( H / 2 );
    DevBuf<uint8_t>  dIn;
    DevBuf<uint16_t> dOut;
    DevBuf<float>    dC, dTmp;
    dIn.reserve( Q + 2 * q4 ), dOut.reserve( 3 * Q ), dC.reserve( q4 ), dTmp.reserve( 2 * q4 );
    PCC_CUDA( cudaMemcpyAsync( dIn, yuv420, Q + 2 * q4, cudaMemcpyHostToDevice, s ) );
    kLuma16<<<divUp( Q, 256 ), 256, 0, s>>>( dIn, Q, dOut );
    for ( int c = 0; c < 2; ++c ) {
      kChromaToFloat<<<divUp( q4, 256 ), 256, 0, s>>>( dIn.p + Q + size_t( c ) * q4, q4, dC );
      kChromaUpV<<<divUp( q4, 256 ), 256, 0, s>>>( dC, w2, h2, dTmp );
      kChromaUpH<<<divUp( size_t( w2 ) * H, 256 ), 256, 0, s>>>( dTmp, w2, int( H ), dOut.p + Q * ( 1 + c ) );
    }
    PCC_LAUNCH_CHECK();
    PCC_CUDA( cudaMemcpyAsync( yuv444, dOut, 3 * Q * 2, cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    return PCCB200_OK;
  } );
}

// PCCPointSet3::convertYUV16ToRGB8 for n points (n x 3 each)
int pccb200_yuv16_to_rgb8( pccb200_ctx* ctx, const uint16_t* yuv, size_t n, uint8_t* rgb ) {
  return guarded( ctx, [&]() -> int {
    if ( !yuv || !rgb ) return PCCB200_ERR_BAD_ARG;
    if ( n == 0 ) return PCCB200_OK;
    cudaStream_t     s = ctx->stream;
    DevBuf<uint16_t> dIn;
    DevBuf<uint8_t>  dOut;
    dIn.reserve( 3 * n ), dOut.reserve( 3 * n );
    PCC_CUDA( cudaMemcpyAsync( dIn, yuv, 3 * n * 2, cudaMemcpyHostToDevice, s ) );
    kYuv16ToRgb8<<<divUp( n, 256 ), 256, 0, s>>>( dIn, n, dOut );
    PCC_LAUNCH_CHECK();
    PCC_CUDA( cudaMemcpyAsync( rgb, dOut, 3 * n, cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
    return PCCB200_OK;
  } );
}
}
