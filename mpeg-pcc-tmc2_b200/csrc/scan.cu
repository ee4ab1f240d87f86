// scan.cu — device-wide exclusive prefix sum (uint32): ONE kernel launch (single-pass, decoupled look-back: scan.cuh) plus
// the memset of its control block. Used for compaction of patch points, point emission order, voxel run numbering; the kd-tree
// build scans with the same tile routine inside its own kernels (flags computed on the fly).
#include "common.cuh"
#include "scan.cuh"

namespace pccb200 {

namespace {
struct LoadU32 {
  const uint32_t* in;
  __device__ __forceinline__ uint32_t operator()( size_t i ) const { return in[i]; }
};
__global__ void __launch_bounds__( kScanThreads ) kScanU32( const uint32_t* in, uint32_t* out, size_t n, unsigned long long* ctl ) {
  scanLookbackTile( LoadU32{ in }, out, n, ctl );  // (in == out is fine: a thread reads all its elements before it writes any)
}
}  // namespace

size_t scanTmpElems( size_t n ) { return 2 * scanCtlWords( n ) + 4; }  // uint32 elements (8-byte alignment slack included)

void exclusiveScanU32( const uint32_t* in, uint32_t* out, size_t n, uint32_t* tmp, cudaStream_t s ) {
  if ( n == 0 ) {
    PCC_CUDA( cudaMemsetAsync( out, 0, sizeof( uint32_t ), s ) );
    return;
  }
  unsigned long long* ctl = reinterpret_cast<unsigned long long*>( ( reinterpret_cast<uintptr_t>( tmp ) + 7 ) & ~uintptr_t( 7 ) );
  PCC_CUDA( cudaMemsetAsync( ctl, 0, scanCtlWords( n ) * sizeof( unsigned long long ), s ) );
  kScanU32<<<unsigned( scanTiles( n ) ), kScanThreads, 0, s>>>( in, out, n, ctl );
  PCC_LAUNCH_CHECK();
}

}  // namespace pccb200
