// pack_ra.cu — device side of the random-access packing (ra_pack.hpp): the placement kernel and its host driver.
// One CTA works through the jobs of a batch one after the other (a job = a canvas + an ordered list of patches; the frames
// of a sub-context depend on each other through their matched patches, so a batch is a sequential chain by nature). The
// canvas is a bit matrix in shared memory; for every item all threads test candidate positions in raster order and the
// first fit wins - the order-dependent first fit of PCCPatch::checkFitPatchCanvas (PccLibCommon/source/PCCPatch.cpp:310-335,
// ...ForGPA :667-692), whose every probe copies the whole std::vector<bool> canvas in the reference.
#include <climits>
#include <stdexcept>

#include "ra_pack.hpp"
#include "stages.cuh"

namespace pccb200 {

namespace {

constexpr int kWordsPerRow = 6;     // 192 blocks: 2560-px canvases (vox11) plus slack
constexpr int kMaxRows     = 1024;  // 16384 px

__device__ __forceinline__ bool spanFree( const uint32_t* row, int x0, int w ) {
  int x = x0, left = w;
  while ( left > 0 ) {
    const int      word = x >> 5, bit = x & 31, take = min( left, 32 - bit );
    const uint32_t mask = ( take == 32 ? 0xFFFFFFFFu : ( ( 1u << take ) - 1u ) ) << bit;
    if ( row[word] & mask ) return false;
    x += take, left -= take;
  }
  return true;
}
// the bounding box of the patch (orientation 0 = default, 1 = swapped axes) must lie on the canvas and be free
__device__ __forceinline__ bool boxFits( const uint32_t* canvas, int sizeU, int sizeV, int u, int v, int orient, int sU0, int sV0 ) {
  const int bw = orient == 0 ? sU0 : sV0, bh = orient == 0 ? sV0 : sU0;
  if ( u < 0 || v < 0 || u + bw > sizeU || v + bh > sizeV ) return false;
  for ( int r = 0; r < bh; ++r )
    if ( !spanFree( canvas + ( v + r ) * kWordsPerRow, u, bw ) ) return false;
  return true;
}

__global__ void __launch_bounds__( 512, 1 ) kPlace( ra::PlaceItem* __restrict__ items, ra::PlaceJob* __restrict__ jobs, int numJobs, const uint8_t* __restrict__ occ, int occRes ) {
  extern __shared__ uint32_t canvas[];  // kMaxRows x kWordsPerRow
  __shared__ int             best;
  for ( int ji = 0; ji < numJobs; ++ji ) {
    const ra::PlaceJob job = jobs[ji];
    __syncthreads();
    for ( int i = threadIdx.x; i < kMaxRows * kWordsPerRow; i += blockDim.x ) canvas[i] = 0;
    __syncthreads();
    const int sizeU = job.sizeU;
    int       sizeV = job.sizeV, width = job.widthPx, height = job.heightPx, error = 0;
    if ( sizeU > kWordsPerRow * 32 || sizeV > kMaxRows ) error = 1;
    for ( int k = 0; k < job.numItems && !error; ++k ) {
      ra::PlaceItem it = items[job.firstItem + k];
      if ( it.refItem >= 0 ) {  // (written by this CTA earlier in the launch; ordered by the barrier that ended that item)
        const ra::PlaceItem r = items[it.refItem];
        it.u0 = r.u0, it.v0 = r.v0, it.orient = r.orient;
      }
      const int sU0 = it.sizeU0, sV0 = it.sizeV0;
      const int o0 = it.aspU0 > it.aspV0 ? 1 : 0, o1 = o0 ^ 1;  // g_orientationHorizontal = {SWAP, DEFAULT}; g_orientationVertical = {DEFAULT, SWAP}
      int       u0 = it.u0, v0 = it.v0, orient = it.orient;
      bool      found = it.mode == ra::PLACE_FIXED;
      int       mode  = it.mode;
      if ( mode == ra::PLACE_STICKY ) {  // position (0,0) with both orientations; afterwards only the one tried last
        if ( boxFits( canvas, sizeU, sizeV, 0, 0, o0, sU0, sV0 ) ) {
          u0 = 0, v0 = 0, orient = o0, found = true;
        } else if ( boxFits( canvas, sizeU, sizeV, 0, 0, o1, sU0, sV0 ) ) {
          u0 = 0, v0 = 0, orient = o1, found = true;
        } else {
          orient = o1, mode = ra::PLACE_KNOWN;
        }
      }
      while ( !found ) {
        if ( mode == ra::PLACE_MATCHED && boxFits( canvas, sizeU, sizeV, u0, v0, orient, sU0, sV0 ) ) break;
        const int per   = mode == ra::PLACE_BEST_EFFORT ? 2 : 1;
        const int total = sizeV * sizeU * per;
        int       hit   = -1;
        for ( int base = 0; base < total && hit < 0; base += blockDim.x ) {
          if ( threadIdx.x == 0 ) best = INT_MAX;
          __syncthreads();
          const int c = base + threadIdx.x;
          if ( c < total ) {
            const int pos = c / per, o = per == 2 ? ( ( c & 1 ) ? o1 : o0 ) : orient;
            if ( boxFits( canvas, sizeU, sizeV, pos % sizeU, pos / sizeU, o, sU0, sV0 ) ) atomicMin( &best, c );
          }
          __syncthreads();
          if ( best != INT_MAX ) hit = best;
          __syncthreads();
        }
        if ( hit >= 0 ) {
          const int pos = hit / per;
          u0 = pos % sizeU, v0 = pos / sizeU;
          if ( per == 2 ) orient = ( hit & 1 ) ? o1 : o0;
          found = true;
        } else {
          sizeV *= 2;  // (the new rows are already zero)
          if ( sizeV > kMaxRows ) {
            error = 1;
            break;
          }
        }
      }
      if ( error ) break;
      // take the occupied blocks only (PCCEncoder.cpp:1383-1391)
      for ( int b = threadIdx.x; b < sU0 * sV0; b += blockDim.x ) {
        const int ub = b % sU0, vb = b / sU0;
        if ( !occ[it.occOff + vb * it.occStride + ub] ) continue;
        const int x = orient == 0 ? ub + u0 : vb + u0, y = orient == 0 ? vb + v0 : ub + v0;
        if ( x < sizeU && y < sizeV ) atomicOr( &canvas[y * kWordsPerRow + ( x >> 5 )], 1u << ( x & 31 ) );
      }
      height = max( height, ( v0 + ( orient == 0 ? sV0 : sU0 ) ) * occRes );
      width  = max( width, ( u0 + ( orient == 0 ? sU0 : sV0 ) ) * occRes );
      if ( threadIdx.x == 0 ) {
        ra::PlaceItem& out = items[job.firstItem + k];
        out.u0 = u0, out.v0 = v0, out.orient = orient;
      }
      __syncthreads();
    }
    if ( threadIdx.x == 0 ) jobs[ji].widthPx = width, jobs[ji].heightPx = height, jobs[ji].error = error;
  }
}

struct CudaPlacer : ra::Placer {
  cudaStream_t              s;
  int                       occRes;
  DevBuf<ra::PlaceItem>&    dItems;
  DevBuf<ra::PlaceJob>&     dJobs;
  DevBuf<uint8_t>&          dOcc;
  Profiler*                 prof;
  CudaPlacer( cudaStream_t st, int res, DevBuf<ra::PlaceItem>& a, DevBuf<ra::PlaceJob>& b, DevBuf<uint8_t>& c, Profiler* p ) :
      s( st ), occRes( res ), dItems( a ), dJobs( b ), dOcc( c ), prof( p ) {}
  void run( std::vector<ra::PlaceItem>& items, std::vector<ra::PlaceJob>& jobs, const std::vector<uint8_t>& occ ) override {
    if ( jobs.empty() ) return;
    dItems.reserve( items.size() + 1 ), dJobs.reserve( jobs.size() ), dOcc.reserve( occ.size() + 1 );
    if ( !items.empty() ) PCC_CUDA( cudaMemcpyAsync( dItems, items.data(), items.size() * sizeof( ra::PlaceItem ), cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( dJobs, jobs.data(), jobs.size() * sizeof( ra::PlaceJob ), cudaMemcpyHostToDevice, s ) );
    if ( !occ.empty() ) PCC_CUDA( cudaMemcpyAsync( dOcc, occ.data(), occ.size(), cudaMemcpyHostToDevice, s ) );
    const size_t smem = size_t( kMaxRows ) * kWordsPerRow * sizeof( uint32_t );
    kPlace<<<1, 512, smem, s>>>( dItems, dJobs, int( jobs.size() ), dOcc, occRes );
    PCC_LAUNCH_CHECK();
    if ( !items.empty() ) PCC_CUDA( cudaMemcpyAsync( items.data(), dItems, items.size() * sizeof( ra::PlaceItem ), cudaMemcpyDeviceToHost, s ) );
    PCC_CUDA( cudaMemcpyAsync( jobs.data(), dJobs, jobs.size() * sizeof( ra::PlaceJob ), cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
  }
};

}  // namespace

bool packGofRandomAccess( std::vector<ra::Frame>& frames, int occRes, size_t minWidth, size_t minHeight, RaPackScratch& sc, Profiler* prof, cudaStream_t s ) {
  ProfScope  t( prof, "pack_ra", s );
  CudaPlacer placer( s, occRes, sc.items, sc.jobs, sc.occ, prof );
  ra::GofPacker packer( frames, occRes, minWidth, minHeight, placer );
  return packer.run();
}

}  // namespace pccb200
