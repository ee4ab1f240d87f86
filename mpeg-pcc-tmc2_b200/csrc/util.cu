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
