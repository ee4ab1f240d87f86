// ctx.cuh — the library context and per-frame state shared by api.cu and gof.cu.
#pragma once
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
  RefineScratch    refine;
  PatchScratch     patch;
  PatchResult      seg;  // patches in creation order + device arenas
  // canvas
  std::vector<pccb200_patch> packed;  // packed (sorted) order, u0/v0/orientation filled
  DevBuf<CanvasPatch>        dPatches;
  DevBuf<long long>          elemBase;
  DevBuf<int>                packResult;
  long long                  totalElems = 0;
  int                        heightPx = 0, maxPatchPixels = 1, maxPatchBlocks = 1;
  CanvasImages               im;
  ReconScratch               rc;
  ColorScratch               color;
  AttrImages                 attr;
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
  RefineScratch    refine;
  PatchScratch     patch;
  DevBuf<unsigned char> walkArgs;  // per-frame arguments of a batched orientation walk
  std::vector<std::unique_ptr<FrameState>> framePool;  // reused by successive GOFs (gof.cu)
  ~pccb200_ctx();
};

struct pccb200_patchlist {
  std::vector<pccb200_patch> patches;
  std::vector<int16_t>       depth;
  std::vector<uint8_t>       occ;
};


namespace pccb200 {

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
