// ctx.cuh — the library context and per-frame state shared by api.cu and gof.cu.
#pragma once
#include <memory>
#include <mutex>
#include <new>

#include "stages.cuh"

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
  std::vector<std::unique_ptr<struct FrameState>> framePool;  // reused by successive GOFs (gof.cu)
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
