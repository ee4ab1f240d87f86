// api.cu — the C ABI (include/pccb200.h). Host-side orchestration only; all point/pixel work is in kernels.
#include <mutex>
#include <new>

#include "stages.cuh"

using namespace pccb200;

struct pccb200_ctx {
  int              device = 0;
  cudaStream_t     stream = nullptr;
  std::string      lastError;
  // scratch for the stage-level entry points
  KdTree           tree;
  DevBuf<int16_t>  xyzRaw;
  DevBuf<short4>   xyz4, q4;
  DevBuf<uint32_t> nbr;
  DevBuf<float>    nbrDist;
  DevBuf<double>   normals;
};

namespace {

template <class F>
int guarded( pccb200_ctx* ctx, F&& f ) {
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

void uploadXyz( pccb200_ctx* c, const int16_t* xyz, size_t n, DevBuf<short4>& dst ) {
  c->xyzRaw.reserve( 3 * n + 3 );
  dst.reserve( n + 1 );
  if ( n == 0 ) return;
  PCC_CUDA( cudaMemcpyAsync( c->xyzRaw, xyz, 3 * n * sizeof( int16_t ), cudaMemcpyHostToDevice, c->stream ) );
  packXyz( c->xyzRaw, n, dst, c->stream );
}

}  // namespace

extern "C" {

const char* pccb200_version( void ) { return "pccb200 0.1 (sm_100a)"; }

int pccb200_create( int device, pccb200_ctx** out ) {
  if ( !out ) return PCCB200_ERR_BAD_ARG;
  *out      = nullptr;
  int count = 0;
  if ( cudaGetDeviceCount( &count ) != cudaSuccess || count <= 0 || device < 0 || device >= count ) {
    cudaGetLastError();
    return PCCB200_ERR_NO_DEVICE;  // no CPU fallback by design
  }
  pccb200_ctx* c = new ( std::nothrow ) pccb200_ctx();
  if ( !c ) return PCCB200_ERR_CUDA;
  c->device = device;
  if ( cudaSetDevice( device ) != cudaSuccess || cudaStreamCreateWithFlags( &c->stream, cudaStreamNonBlocking ) != cudaSuccess ) {
    cudaGetLastError();
    delete c;
    return PCCB200_ERR_CUDA;
  }
  *out = c;
  return PCCB200_OK;
}

void pccb200_destroy( pccb200_ctx* ctx ) {
  if ( !ctx ) return;
  cudaSetDevice( ctx->device );
  if ( ctx->stream ) {
    cudaStreamSynchronize( ctx->stream );
    cudaStreamDestroy( ctx->stream );
  }
  delete ctx;
}

const char* pccb200_last_error( const pccb200_ctx* ctx ) { return ctx ? ctx->lastError.c_str() : "null context"; }

int pccb200_knn( pccb200_ctx* ctx, const int16_t* xyz, size_t n, const int16_t* q, size_t nq, int k, uint32_t* idx, float* dist2 ) {
  return guarded( ctx, [&]() -> int {
    if ( !xyz || !idx || ( k != 1 && k != 8 && k != 16 ) ) return PCCB200_ERR_BAD_ARG;
    uploadXyz( ctx, xyz, n, ctx->xyz4 );
    kdBuild( ctx->tree, ctx->xyz4, n, ctx->stream );
    const short4*   queries = ctx->xyz4;
    const uint32_t* order   = ctx->tree.vind;
    if ( q ) {
      uploadXyz( ctx, q, nq, ctx->q4 );
      queries = ctx->q4, order = nullptr;
    } else {
      nq = n;
    }
    ctx->nbr.reserve( nq * k + 1 );
    ctx->nbrDist.reserve( nq * k + 1 );
    if ( n == 0 ) {
      PCC_CUDA( cudaMemsetAsync( ctx->nbr, 0xff, nq * k * sizeof( uint32_t ), ctx->stream ) );
    } else {
      kdKnn( ctx->tree, queries, nq, order, k, ctx->nbr, ctx->nbrDist, ctx->stream );
    }
    PCC_CUDA( cudaMemcpyAsync( idx, ctx->nbr, nq * k * sizeof( uint32_t ), cudaMemcpyDeviceToHost, ctx->stream ) );
    if ( dist2 && n ) PCC_CUDA( cudaMemcpyAsync( dist2, ctx->nbrDist, nq * k * sizeof( float ), cudaMemcpyDeviceToHost, ctx->stream ) );
    PCC_CUDA( cudaStreamSynchronize( ctx->stream ) );
    return PCCB200_OK;
  } );
}

int pccb200_kdtree_order( pccb200_ctx* ctx, const int16_t* xyz, size_t n, uint32_t* vind ) {
  return guarded( ctx, [&]() -> int {
    if ( !xyz || !vind ) return PCCB200_ERR_BAD_ARG;
    uploadXyz( ctx, xyz, n, ctx->xyz4 );
    kdBuild( ctx->tree, ctx->xyz4, n, ctx->stream );
    if ( n ) PCC_CUDA( cudaMemcpyAsync( vind, ctx->tree.vind, n * sizeof( uint32_t ), cudaMemcpyDeviceToHost, ctx->stream ) );
    PCC_CUDA( cudaStreamSynchronize( ctx->stream ) );
    return PCCB200_OK;
  } );
}

int pccb200_normals( pccb200_ctx* ctx, const int16_t* xyz, size_t n, int k, int orientation, double* normals ) {
  return guarded( ctx, [&]() -> int {
    if ( !xyz || !normals || k != 16 ) return PCCB200_ERR_BAD_ARG;
    if ( orientation != 0 ) return PCCB200_ERR_UNSUPPORTED;
    if ( n == 0 ) return PCCB200_OK;
    uploadXyz( ctx, xyz, n, ctx->xyz4 );
    kdBuild( ctx->tree, ctx->xyz4, n, ctx->stream );
    ctx->nbr.reserve( n * k + 1 );
    ctx->normals.reserve( 3 * n );
    kdKnn( ctx->tree, ctx->xyz4, n, ctx->tree.vind, k, ctx->nbr, nullptr, ctx->stream );
    computeNormals( ctx->xyz4, ctx->nbr, k, n, ctx->normals, ctx->stream );
    PCC_CUDA( cudaMemcpyAsync( normals, ctx->normals, 3 * n * sizeof( double ), cudaMemcpyDeviceToHost, ctx->stream ) );
    PCC_CUDA( cudaStreamSynchronize( ctx->stream ) );
    return PCCB200_OK;
  } );
}

}  // extern "C"
