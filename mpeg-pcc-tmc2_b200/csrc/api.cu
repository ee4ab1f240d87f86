// api.cu — the C ABI (include/pccb200.h). Host-side orchestration only; all point/pixel work is in kernels.
#include <stdlib.h>

#include <mutex>
#include <algorithm>
#include <new>

#include "stages.cuh"

using namespace pccb200;

#include "ctx.cuh"

extern "C" {

const char* pccb200_version( void ) { return "pccb200 0.1 (sm_100a)"; }

int pccb200_create( int device, pccb200_ctx** out ) {
  if ( !out ) return PCCB200_ERR_BAD_ARG;
  *out      = nullptr;
  // one hardware work queue per frame stream (must be set before the CUDA context exists; harmless otherwise)
  setenv( "CUDA_DEVICE_MAX_CONNECTIONS", "32", 0 );
  int count = 0;
  if ( cudaGetDeviceCount( &count ) != cudaSuccess || count <= 0 || device < 0 || device >= count ) {
    cudaGetLastError();
    return PCCB200_ERR_NO_DEVICE;  // no CPU fallback by design
  }
  // host threads wait for the device sleeping, not spinning (one thread per frame in flight: far more threads than cores);
  // best effort - the explicit waits of the library (streamWait) block regardless
  if ( !spinWaits() && cudaSetDevice( device ) == cudaSuccess ) cudaSetDeviceFlags( cudaDeviceScheduleBlockingSync );
  cudaGetLastError();
  pccb200_ctx* c = new ( std::nothrow ) pccb200_ctx();
  if ( !c ) return PCCB200_ERR_CUDA;
  c->device = device;
  // The context's own stream carries the GOF-level work that is latency-critical and tiny: the batched orientation walks (one warp
  // per frame) and the single-CTA placement searches of the random-access packer, which the host waits for launch by launch. It gets
  // the highest priority, so that these CTAs take the next free SM slot instead of queueing behind the frames' N-sized grids.
  int prioLow = 0, prioHigh = 0;
  if ( cudaSetDevice( device ) == cudaSuccess ) cudaDeviceGetStreamPriorityRange( &prioLow, &prioHigh );
  if ( cudaSetDevice( device ) != cudaSuccess || cudaStreamCreateWithPriority( &c->stream, cudaStreamNonBlocking, prioHigh ) != cudaSuccess ) {
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

int pccb200_profile_enable( pccb200_ctx* ctx, int on ) {
  if ( !ctx ) return PCCB200_ERR_BAD_ARG;
  ctx->prof.enabled = on != 0;
  ctx->prof.results.clear();
  return PCCB200_OK;
}

int pccb200_profile_read( pccb200_ctx* ctx, char* names, float* ms, float* start_ms, int capacity, int* count ) {
  if ( !ctx || !count ) return PCCB200_ERR_BAD_ARG;
  const int n = int( ctx->prof.results.size() );
  *count      = n;
  if ( names && ms ) {
    for ( int i = 0; i < n && i < capacity; ++i ) {
      snprintf( names + 32 * i, 32, "%s", ctx->prof.results[i].first );
      ms[i] = ctx->prof.results[i].second;
      if ( start_ms ) start_ms[i] = ctx->prof.results[i].start;
    }
    ctx->prof.results.clear();
  }
  return PCCB200_OK;
}

int pccb200_set_scratch_sets( int device, int count ) {
  if ( device < 0 || count < 1 ) return PCCB200_ERR_BAD_ARG;
  return setFrameScratchSets( device, count );
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
    streamWait( ctx->stream );
    return PCCB200_OK;
  } );
}

int pccb200_kdtree_order( pccb200_ctx* ctx, const int16_t* xyz, size_t n, uint32_t* vind ) {
  return guarded( ctx, [&]() -> int {
    if ( !xyz || !vind ) return PCCB200_ERR_BAD_ARG;
    uploadXyz( ctx, xyz, n, ctx->xyz4 );
    kdBuild( ctx->tree, ctx->xyz4, n, ctx->stream );
    if ( n ) PCC_CUDA( cudaMemcpyAsync( vind, ctx->tree.vind, n * sizeof( uint32_t ), cudaMemcpyDeviceToHost, ctx->stream ) );
    streamWait( ctx->stream );
    return PCCB200_OK;
  } );
}

int pccb200_normals( pccb200_ctx* ctx, const int16_t* xyz, size_t n, int k, int orientation, double* normals ) {
  return guarded( ctx, [&]() -> int {
    if ( !xyz || !normals || k != 16 ) return PCCB200_ERR_BAD_ARG;
    if ( orientation != 0 && orientation != 1 ) return PCCB200_ERR_UNSUPPORTED;
    if ( n == 0 ) return PCCB200_OK;
    uploadXyz( ctx, xyz, n, ctx->xyz4 );
    kdBuild( ctx->tree, ctx->xyz4, n, ctx->stream );
    ctx->nbr.reserve( n * k + 1 );
    ctx->normals.reserve( 3 * n );
    kdKnn( ctx->tree, ctx->xyz4, n, ctx->tree.vind, k, ctx->nbr, nullptr, ctx->stream );
    computeNormals( ctx->xyz4, ctx->nbr, k, n, ctx->normals, ctx->stream );
    if ( orientation == 1 ) orientNormals( ctx->orient, ctx->own.orientTmp, ctx->walkArgs, ctx->xyz4, ctx->nbr, k, ctx->tree.vind, n, ctx->normals, ctx->stream );
    PCC_CUDA( cudaMemcpyAsync( normals, ctx->normals, 3 * n * sizeof( double ), cudaMemcpyDeviceToHost, ctx->stream ) );
    streamWait( ctx->stream );
    return PCCB200_OK;
  } );
}

int pccb200_weight_normal( pccb200_ctx* ctx, const int16_t* xyz, size_t n, int bits, double minWeight, double w[3] ) {
  return guarded( ctx, [&]() -> int {
    if ( !xyz || !w || bits < 1 || bits > 12 ) return PCCB200_ERR_BAD_ARG;
    uploadXyz( ctx, xyz, n, ctx->xyz4 );
    const size_t words = ( size_t( 1 ) << ( 2 * bits ) ) / 32;
    ctx->faces.reserve( 3 * words + 1 ), ctx->faceCounts.reserve( 4 );
    projectedAreas( ctx->xyz4, n, bits, ctx->faces, ctx->faceCounts, ctx->stream );
    unsigned cnt[3];
    PCC_CUDA( cudaMemcpyAsync( cnt, ctx->faceCounts, sizeof( cnt ), cudaMemcpyDeviceToHost, ctx->stream ) );
    streamWait( ctx->stream );
    // three numbers on the host: order the planes by area (stable, ascending) and derive the weights
    int order[3] = { 0, 1, 2 };
    for ( int a = 1; a < 3; ++a )
      for ( int b = a; b > 0 && cnt[order[b]] < cnt[order[b - 1]]; --b ) std::swap( order[b], order[b - 1] );
    const double c0 = double( cnt[order[0]] ), c1 = double( cnt[order[1]] ), c2 = double( cnt[order[2]] );
    double       ax[3];
    if ( c0 / c2 >= minWeight ) {
      ax[order[0]] = c0 / c2, ax[order[1]] = c1 / c2, ax[order[2]] = 1.0;
    } else {
      const double tb = c1 / c2, ta = c0 / c2;
      ax[order[0]] = minWeight, ax[order[2]] = 1.0;
      ax[order[1]] = minWeight + ( tb - ta ) / ( 1.0 - ta ) * ( 1 - minWeight );
    }
    w[0] = ax[0], w[1] = ax[1], w[2] = ax[2];
    return PCCB200_OK;
  } );
}

int pccb200_segment_frame( pccb200_ctx* ctx, const int16_t* xyz, const uint8_t* rgb, size_t n, const pccb200_seg_params* prm,
                           double* normals, uint8_t* part0, uint8_t* part1, pccb200_patchlist** out ) {
  return guarded( ctx, [&]() -> int {
    if ( !xyz || !rgb || !prm || !out ) return PCCB200_ERR_BAD_ARG;
    if ( !segParamsSupported( *prm ) ) return PCCB200_ERR_UNSUPPORTED;
    *out = nullptr;
    pccb200_patchlist* pl = new pccb200_patchlist();
    if ( n == 0 ) {
      *out = pl;
      return PCCB200_OK;
    }
    cudaStream_t s = ctx->stream;
    const int    k = 16;
    Profiler*    pf = &ctx->prof;
    {
      ProfScope t( pf, "h2d", s );
      uploadXyz( ctx, xyz, n, ctx->xyz4 );
      ctx->rgbRaw.reserve( 3 * n ), ctx->rgb4.reserve( n ), ctx->partition.reserve( n );
      PCC_CUDA( cudaMemcpyAsync( ctx->rgbRaw, rgb, 3 * n, cudaMemcpyHostToDevice, s ) );
      packRgb( ctx->rgbRaw, n, ctx->rgb4, s );
    }
    {
      ProfScope t( pf, "kdtree_build", s );
      kdBuild( ctx->tree, ctx->xyz4, n, s );
    }
    ctx->nbr.reserve( n * k + 1 ), ctx->normals.reserve( 3 * n );
    {
      ProfScope t( pf, "knn16", s );
      kdKnn( ctx->tree, ctx->xyz4, n, ctx->tree.vind, k, ctx->nbr, nullptr, s );
    }
    {
      ProfScope t( pf, "normals", s );
      computeNormals( ctx->xyz4, ctx->nbr, k, n, ctx->normals, s );
    }
    if ( prm->normal_orientation == 1 ) {
      ProfScope t( pf, "orient", s );
      ctx->orient.prof = pf;
      orientNormals( ctx->orient, ctx->own.orientTmp, ctx->walkArgs, ctx->xyz4, ctx->nbr, k, ctx->tree.vind, n, ctx->normals, s );
    }
    if ( normals ) PCC_CUDA( cudaMemcpyAsync( normals, ctx->normals, 3 * n * sizeof( double ), cudaMemcpyDeviceToHost, s ) );
    {
      ProfScope t( pf, "initial_seg", s );
      initialSegmentation( ctx->normals, n, prm->weight_normal, ctx->partition, s );
    }
    if ( part0 ) PCC_CUDA( cudaMemcpyAsync( part0, ctx->partition, n, cudaMemcpyDeviceToHost, s ) );
    {
      ProfScope t( pf, "refine", s );
      refineSegmentation( ctx->own.refine, ctx->xyz4, ctx->normals, n, *prm, ctx->partition, s );
    }
    if ( part1 ) PCC_CUDA( cudaMemcpyAsync( part1, ctx->partition, n, cudaMemcpyDeviceToHost, s ) );
    PatchResult res;
    {
      ProfScope t( pf, "patches", s );
      segmentPatches( ctx->own.patch, res, ctx->xyz4, ctx->rgb4, ctx->nbr, k, ctx->partition, n, *prm, s );
    }
    {
      ProfScope t( pf, "d2h", s );
      pl->patches = res.patches;
      pl->depth.resize( res.depthElems ), pl->occ.resize( res.occElems );
      if ( res.depthElems ) PCC_CUDA( cudaMemcpyAsync( pl->depth.data(), res.depth, res.depthElems * sizeof( int16_t ), cudaMemcpyDeviceToHost, s ) );
      if ( res.occElems ) PCC_CUDA( cudaMemcpyAsync( pl->occ.data(), res.occ, res.occElems, cudaMemcpyDeviceToHost, s ) );
    }
    streamWait( s );
    ctx->prof.collect( s );
    *out = pl;
    return PCCB200_OK;
  } );
}

int    pccb200_patches_count( const pccb200_patchlist* pl ) { return pl ? int( pl->patches.size() ) : 0; }
size_t pccb200_patches_depth_elems( const pccb200_patchlist* pl ) { return pl ? pl->depth.size() : 0; }
size_t pccb200_patches_occ_elems( const pccb200_patchlist* pl ) { return pl ? pl->occ.size() : 0; }
int    pccb200_patches_get( const pccb200_patchlist* pl, pccb200_patch* patches, int16_t* depth, uint8_t* occ ) {
  if ( !pl ) return PCCB200_ERR_BAD_ARG;
  if ( patches ) std::copy( pl->patches.begin(), pl->patches.end(), patches );
  if ( depth ) std::copy( pl->depth.begin(), pl->depth.end(), depth );
  if ( occ ) std::copy( pl->occ.begin(), pl->occ.end(), occ );
  return PCCB200_OK;
}
void pccb200_patches_free( pccb200_patchlist* pl ) { delete pl; }

}  // extern "C"
