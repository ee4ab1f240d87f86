// ctx.cuh — the library context and per-frame state shared by api.cu and gof.cu.
#pragma once
#include <chrono>
#include <memory>
#include <mutex>
#include <new>

#include "stages.cuh"

// per-frame device state of a GOF (gof.cu); pooled in the context and reused by successive GOFs
struct FrameState {
  cudaStream_t     stream = nullptr;
  const int16_t*   hXyz   = nullptr;
  const uint8_t*   hRgb   = nullptr;
  size_t           n      = 0;
  DevBuf<int16_t>  xyzRaw;
  DevBuf<uint8_t>  rgbRaw, partition;
  DevBuf<short4>   xyz4;
  DevBuf<uchar4>   rgb4, recRgb;
  KdTree           tree;
  DevBuf<uint32_t> nbr;
  DevBuf<double>   normals;
  OrientScratch    orient;
  FrameScratch*    fat = nullptr;  // leased for the duration of one stage group (ScratchLease)
  PatchResult      seg;  // patches in creation order + device arenas
  // canvas
  std::vector<pccb200_patch> packed;  // packed (sorted) order, u0/v0/orientation filled
  DevBuf<CanvasPatch>        dPatches;
  DevBuf<long long>          elemBase;
  DevBuf<int>                packResult;
  long long                  totalElems = 0;
  int                        heightPx = 0, widthPx = 0, maxPatchPixels = 1, maxPatchBlocks = 1;
  CanvasImages               im;
  ReconScratch               rc;
  AttrImages                 attr;
  YuvScratch                 yuv;
  DevBuf<uint8_t>            geoLuma8;  // GEO0 / GEO1 narrowed to bytes on request (hand-off to an 8-bit codec)
  Profiler                   prof;
  bool                       decodedSet = false;  // the caller replaced om/geo0/geo1 by decoded planes
  int                        status = 0;
  std::string                error;
  ~FrameState() {
    if ( stream ) cudaStreamDestroy( stream );
  }
};

struct pccb200_ctx {
  int              device = 0;
  cudaStream_t     stream = nullptr;
  std::string      lastError;
  Profiler         prof;
  // scratch for the stage-level entry points
  KdTree           tree;
  DevBuf<int16_t>  xyzRaw;
  DevBuf<short4>   xyz4, q4;
  DevBuf<uint32_t> nbr;
  DevBuf<float>    nbrDist;
  DevBuf<double>   normals;
  DevBuf<uint8_t>  rgbRaw, partition;
  DevBuf<uint32_t> faces;
  DevBuf<unsigned> faceCounts;
  DevBuf<uchar4>   rgb4;
  OrientScratch    orient;
  FrameScratch     own;  // scratch of the stage-level entry points (the GOF path leases sets from the device pool)
  DevBuf<unsigned char> walkArgs;  // per-frame arguments of a batched orientation walk
  RaPackScratch         raPack;    // batches of the random-access packer
  std::vector<std::unique_ptr<FrameState>> framePool;  // reused by successive GOFs (gof.cu)
  ~pccb200_ctx();
};

struct pccb200_patchlist {
  std::vector<pccb200_patch> patches;
  std::vector<int16_t>       depth;
  std::vector<uint8_t>       occ;
};


namespace pccb200 {

// Holds one FrameScratch set of the device for a frame while a stage group runs on its stream; the stream is drained before the
// set goes back to the pool (the next holder may be another frame on another stream).
struct ScratchLease {
  FrameState& fs;
  int         device;
  const char* holdName;
  std::chrono::steady_clock::time_point t1;
  // (with the profiler on, the host-side wait for a set and the wall time it was held are reported as spans without a start)
  ScratchLease( FrameState& f, int dev, const char* hold = "hold" ) : fs( f ), device( dev ), holdName( hold ) {
    const auto t0 = std::chrono::steady_clock::now();
    fs.fat        = acquireFrameScratch( dev );
    t1            = std::chrono::steady_clock::now();
    if ( fs.prof.enabled ) fs.prof.results.push_back( Profiler::Result{ "lease_wait", std::chrono::duration<float, std::milli>( t1 - t0 ).count(), -1.f } );
  }
  ~ScratchLease() {
    try {
      streamWait( fs.stream );
    } catch ( const CudaError& ) {
      cudaGetLastError();  // (the failure is reported by the next checked call on this stream)
    }
    releaseFrameScratch( device, fs.fat );
    fs.fat = nullptr;
    if ( fs.prof.enabled )
      fs.prof.results.push_back( Profiler::Result{ holdName, std::chrono::duration<float, std::milli>( std::chrono::steady_clock::now() - t1 ).count(), -1.f } );
  }
  ScratchLease( const ScratchLease& )            = delete;
  ScratchLease& operator=( const ScratchLease& ) = delete;
};

// Parameter combinations the CUDA path implements (the CTC values and the variations the tests cover). Anything
// else is PCCB200_ERR_UNSUPPORTED at the entry point - never a silently different result.
//  * voxel_dim_refine: only 4. Dimensions 1 / 2 make the reference classify edge voxels over a 5x5x5 neighbourhood
//    (idvSearchRange 2, PCCPatchSegmenter.cpp:1470: up to 125 near voxels); refine.cu keeps 27 (the 3x3x3 of dimension >= 4).
//  * search_radius_refine: the lattice offsets with d^2 < (radius >> 2) must fit the adjacency kernel's hit buffer (2048).
inline bool segParamsSupported( const pccb200_seg_params& p ) {
  return p.nn_normal_estimation == 16 && p.max_nn_count_patch_seg == 16 && p.geometry_bitdepth_3d >= 1 && p.geometry_bitdepth_3d <= 12 &&
         ( p.normal_orientation == 0 || p.normal_orientation == 1 ) && p.voxel_dim_refine == 4 && p.search_radius_refine >= 4 &&
         ( p.search_radius_refine >> 2 ) <= 56 && p.max_nn_count_refine >= 1 && p.iteration_count_refine >= 0;
}

template <class F>
inline int guarded( pccb200_ctx* ctx, F&& f ) {
  if ( !ctx ) return PCCB200_ERR_BAD_ARG;
  try {
    PCC_CUDA( cudaSetDevice( ctx->device ) );
    return f();
  } catch ( const CudaError& e ) {
    char buf[512];
    snprintf( buf, sizeof( buf ), "CUDA error %d (%s) at %s:%d", int( e.code ), cudaGetErrorString( e.code ), e.file, e.line );
    ctx->lastError = buf;
    cudaGetLastError();
    return PCCB200_ERR_CUDA;
  } catch ( const std::bad_alloc& ) {
    ctx->lastError = "host allocation failed";
    return PCCB200_ERR_CUDA;
  }
}

inline void uploadXyz( pccb200_ctx* c, const int16_t* xyz, size_t n, DevBuf<short4>& dst ) {
  c->xyzRaw.reserve( 3 * n + 3 );
  dst.reserve( n + 1 );
  if ( n == 0 ) return;
  PCC_CUDA( cudaMemcpyAsync( c->xyzRaw, xyz, 3 * n * sizeof( int16_t ), cudaMemcpyHostToDevice, c->stream ) );
  packXyz( c->xyzRaw, n, dst, c->stream );
}


}  // namespace pccb200
