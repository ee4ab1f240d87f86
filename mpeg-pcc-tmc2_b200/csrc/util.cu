// util.cu — small layout kernels.
#include "stages.cuh"

namespace pccb200 {

namespace {
// PCCPointSet3::positions_ is std::vector<PCCVector3<int16_t>> (6 B/point AoS, PCCPointSet.h:520-528);
// on the device every point is one aligned 8-byte short4 so a point is a single load.
__global__ void kPackXyz( const int16_t* __restrict__ in, int n, short4* __restrict__ out ) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i < n ) out[i] = make_short4( in[3 * size_t( i )], in[3 * size_t( i ) + 1], in[3 * size_t( i ) + 2], 0 );
}
__global__ void kGatherU8( const uint8_t* __restrict__ src, const uint32_t* __restrict__ idx, int n, uint8_t* __restrict__ dst ) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i < n ) dst[i] = src[idx[i]];
}
}  // namespace

void gatherU8( const uint8_t* src, const uint32_t* idx, size_t n, uint8_t* dst, cudaStream_t s ) {
  if ( n == 0 ) return;
  kGatherU8<<<divUp( n, 256 ), 256, 0, s>>>( src, idx, int( n ), dst );
  PCC_LAUNCH_CHECK();
}

void packXyz( const int16_t* xyz3, size_t n, short4* out, cudaStream_t s ) {
  if ( n == 0 ) return;
  kPackXyz<<<divUp( n, 256 ), 256, 0, s>>>( xyz3, int( n ), out );
  PCC_LAUNCH_CHECK();
}

}  // namespace pccb200

// ---- a4: PCCEncoder::calculateWeightNormal (PccLibEncoder/source/PCCEncoder.cpp:3569-3626) -------------
// Projected-area of the cloud on the three axis planes: three 2^b x 2^b bitmaps filled with atomicOr, then popcounts.
namespace pccb200 {
namespace {
__global__ void kProjectFaces( const short4* __restrict__ pts, int n, int side, uint32_t* __restrict__ faces ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  const short4 p  = pts[i];
  const int    p0 = max( 0, min( side - 1, int( p.x ) ) ), p1 = max( 0, min( side - 1, int( p.y ) ) ), p2 = max( 0, min( side - 1, int( p.z ) ) );
  const size_t plane = size_t( side ) * side;
  const size_t b0 = size_t( p2 ) * side + p1, b1 = plane + size_t( p0 ) * side + p2, b2 = 2 * plane + size_t( p1 ) * side + p0;
  atomicOr( &faces[b0 >> 5], 1u << ( b0 & 31 ) );
  atomicOr( &faces[b1 >> 5], 1u << ( b1 & 31 ) );
  atomicOr( &faces[b2 >> 5], 1u << ( b2 & 31 ) );
}
__global__ void kCountFaces( const uint32_t* __restrict__ faces, size_t wordsPerPlane, unsigned* __restrict__ counts ) {
  const size_t w = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  const int    a = blockIdx.y;
  unsigned     c = w < wordsPerPlane ? __popc( faces[a * wordsPerPlane + w] ) : 0;
  c              = __reduce_add_sync( 0xffffffffu, c );
  if ( ( threadIdx.x & 31 ) == 0 && c ) atomicAdd( &counts[a], c );
}
}  // namespace

void projectedAreas( const short4* pts, size_t n, int bits, uint32_t* faces, unsigned* counts, cudaStream_t s ) {
  const int    side  = 1 << bits;
  const size_t words = size_t( side ) * side / 32;
  PCC_CUDA( cudaMemsetAsync( faces, 0, 3 * words * sizeof( uint32_t ), s ) );
  PCC_CUDA( cudaMemsetAsync( counts, 0, 3 * sizeof( unsigned ), s ) );
  if ( n ) kProjectFaces<<<divUp( n, 256 ), 256, 0, s>>>( pts, int( n ), side, faces );
  kCountFaces<<<dim3( divUp( words, 256 ), 3 ), 256, 0, s>>>( faces, words, counts );
  PCC_LAUNCH_CHECK();
}
}  // namespace pccb200
