// gof.cu — GOF-level entry points: the stages of PCCEncoder::encode between `sources` and the three
// videoEncoder.compress calls (PccLibEncoder/source/PCCEncoder.cpp:103-424), for all frames of a GOF.
//
// Frames of a GOF are independent up to the canvas size (the reference itself runs them under tbb::parallel_for,
// PCCEncoder.cpp:4729-4747, 6670-6729, 344-420), so every frame gets its own CUDA stream, device buffers and host worker
// thread; the GPU overlaps the frames' kernels — in particular the latency-bound orientation walk (orient.cu), which
// occupies one warp per frame, overlaps with the bandwidth-bound stages of the other frames.
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <condition_variable>
#include <exception>
#include <mutex>
#include <thread>

#include "stages.cuh"

using namespace pccb200;

#include "ctx.cuh"

pccb200_ctx::~pccb200_ctx() {}

struct pccb200_gof {
  pccb200_ctx*                   ctx = nullptr;
  int                            nframes = 0, occPrec = 4, stage = 0;
  pccb200_seg_params             prm;
  size_t                         W = 0, H = 0;
  std::vector<FrameState*>       frames;
  std::vector<pccb200_patchlist> lists;  // host copies of the packed patch lists
};

namespace {

template <class F>
int forEachFrame( pccb200_gof* g, F&& fn ) {
  std::vector<std::thread> workers;
  workers.reserve( g->nframes );
  for ( int f = 0; f < g->nframes; ++f ) {
    workers.emplace_back( [g, f, &fn]() {
      FrameState& fs = *g->frames[f];
      try {
        PCC_CUDA( cudaSetDevice( g->ctx->device ) );
        fn( fs, f );
        streamWait( fs.stream );
      } catch ( const CudaError& e ) {
        char buf[512];
        snprintf( buf, sizeof( buf ), "frame %d: CUDA error %d (%s) at %s:%d", f, int( e.code ), cudaGetErrorString( e.code ), e.file, e.line );
        fs.error  = buf;
        fs.status = PCCB200_ERR_CUDA;
        cudaGetLastError();
      } catch ( const std::exception& e ) {
        fs.error  = e.what();
        fs.status = PCCB200_ERR_CUDA;
      }
    } );
  }
  for ( auto& w : workers ) w.join();
  for ( int f = 0; f < g->nframes; ++f )
    if ( g->frames[f]->status != 0 ) {
      g->ctx->lastError = g->frames[f]->error;
      return g->frames[f]->status;
    }
  return PCCB200_OK;
}

// Meeting point of the frame threads of a GOF between the data-parallel part of the orientation and the walks: the last frame to
// arrive launches the walks of ALL frames as one grid on the context's stream and waits for them; then everybody goes on.
// (One launch = one hardware queue held for the ~1 s the walks take, instead of one per frame - with every queue occupied by a
// running walk nothing else could start on the device, e.g. the stages of the next GOF.)
struct WalkGate {
  std::mutex                  m;
  std::condition_variable     cv;
  int                         expected = 0, arrived = 0;
  bool                        done = false, failed = false;
  std::string                 error;
  std::vector<OrientScratch*> items;
  void arrive( OrientScratch* item, pccb200_ctx* ctx ) {
    std::unique_lock<std::mutex> lk( m );
    if ( item ) items.push_back( item );
    if ( ++arrived < expected ) {
      cv.wait( lk, [&]() { return done; } );
    } else {
      try {
        orientWalkBatch( items.data(), int( items.size() ), ctx->walkArgs, &ctx->prof, ctx->stream );
      } catch ( const CudaError& e ) {
        char buf[256];
        snprintf( buf, sizeof( buf ), "orientation walk: CUDA error %d (%s) at %s:%d", int( e.code ), cudaGetErrorString( e.code ), e.file, e.line );
        failed = true, error = buf;
        cudaGetLastError();
      } catch ( const std::exception& e ) {  // (e.g. bad_alloc: the other frame threads must still be released)
        failed = true, error = std::string( "orientation walk: " ) + e.what();
      } catch ( ... ) {
        failed = true, error = "orientation walk: unknown failure";
      }
      done = true;
      cv.notify_all();
    }
    if ( failed ) throw std::runtime_error( error );
  }
};

// a1..a4 + the data-parallel part of a5 for one frame (PCCPatchSegmenter3::compute)
void segmentFrameBeforeWalk( FrameState& fs, const pccb200_seg_params& prm ) {
  cudaStream_t s = fs.stream;
  const size_t n = fs.n;
  const int    k = 16;
  Profiler*    pf = &fs.prof;
  fs.seg.patches.clear();
  fs.seg.depthElems = fs.seg.occElems = 0;
  fs.orient.walkSmem = 0;  // (frame states are pooled: an empty frame must not re-submit the previous GOF's walk)
  if ( n == 0 ) return;
  {
    ProfScope t( pf, "h2d", s );
    fs.xyzRaw.reserve( 3 * n ), fs.rgbRaw.reserve( 3 * n ), fs.xyz4.reserve( n ), fs.rgb4.reserve( n ), fs.partition.reserve( n );
    PCC_CUDA( cudaMemcpyAsync( fs.xyzRaw, fs.hXyz, 3 * n * sizeof( int16_t ), cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( fs.rgbRaw, fs.hRgb, 3 * n, cudaMemcpyHostToDevice, s ) );
    packXyz( fs.xyzRaw, n, fs.xyz4, s );
    packRgb( fs.rgbRaw, n, fs.rgb4, s );
  }
  {
    ProfScope t( pf, "kdtree_build", s );
    kdBuild( fs.tree, fs.xyz4, n, s );
  }
  fs.nbr.reserve( n * k + 1 ), fs.normals.reserve( 3 * n );
  {
    ProfScope t( pf, "knn16", s );
    kdKnn( fs.tree, fs.xyz4, n, fs.tree.vind, k, fs.nbr, nullptr, s );
  }
  {
    ProfScope t( pf, "normals", s );
    computeNormals( fs.xyz4, fs.nbr, k, n, fs.normals, s );
  }
  if ( prm.normal_orientation == 1 ) {
    ProfScope t( pf, "orient", s );
    fs.orient.prof = pf;
    orientPrepare( fs.orient, fs.fat->orientTmp, fs.xyz4, fs.nbr, k, fs.tree.vind, n, fs.normals, s );
  }
}

// ... a5 (the orientation walks of all frames, one launch: WalkGate below) ... then a5's sign application and a6..a11
void segmentFrameAfterWalk( FrameState& fs, const pccb200_seg_params& prm ) {
  cudaStream_t s = fs.stream;
  const size_t n = fs.n;
  const int    k = 16;
  Profiler*    pf = &fs.prof;
  if ( n == 0 ) return;
  if ( prm.normal_orientation == 1 ) orientFinish( fs.orient, fs.xyz4, n, fs.normals, s );
  {
    ProfScope t( pf, "initial_seg", s );
    initialSegmentation( fs.normals, n, prm.weight_normal, fs.partition, s );
  }
  {
    ProfScope t( pf, "refine", s );
    refineSegmentation( fs.fat->refine, fs.xyz4, fs.normals, n, prm, fs.partition, s );
  }
  {
    ProfScope t( pf, "patches", s );
    segmentPatches( fs.fat->patch, fs.seg, fs.xyz4, fs.rgb4, fs.nbr, k, fs.partition, n, prm, s );
  }
}

// PCCPatch::gt (PccLibCommon/source/PCCPatch.cpp:349-371): a strict total order, so any sort reproduces std::sort
bool patchBefore( const pccb200_patch& a, const pccb200_patch& b ) {
  const int amax = std::max( a.size_u0, a.size_v0 ), amin = std::min( a.size_u0, a.size_v0 );
  const int bmax = std::max( b.size_u0, b.size_v0 ), bmin = std::min( b.size_u0, b.size_v0 );
  return amax != bmax ? amax > bmax : ( amin != bmin ? amin > bmin : a.index < b.index );
}

// a13 for one frame: sort (KB of metadata, host) + first-fit placement (device)
void packFrame( FrameState& fs, const pccb200_seg_params& prm, int presetWidth, int presetHeight, int numTilesHor, double tileRatio ) {
  cudaStream_t s = fs.stream;
  ProfScope    t( &fs.prof, "pack", s );
  fs.packed   = fs.seg.patches;
  fs.heightPx = presetHeight, fs.widthPx = presetWidth;
  const int P = int( fs.packed.size() );
  fs.totalElems = 0, fs.maxPatchPixels = 1, fs.maxPatchBlocks = 1;
  if ( P == 0 ) return;
  std::sort( fs.packed.begin(), fs.packed.end(), patchBefore );
  const int occRes = prm.occupancy_resolution;
  int       sizeU = presetWidth / occRes, sizeV = std::max( fs.packed[0].size_v0, fs.packed[0].size_u0 );
  for ( auto& p : fs.packed ) sizeU = std::max( sizeU, p.size_u0 + 1 );
  const int tileW = sizeU / numTilesHor, tileH = int( tileW * tileRatio );
  sizeV           = sizeV >= tileH ? sizeV : tileH;
  std::vector<CanvasPatch> cp( P );
  std::vector<long long>   base( P + 1 );
  for ( int i = 0; i < P; ++i ) {
    const pccb200_patch& m = fs.packed[i];
    CanvasPatch&         c = cp[i];
    c.viewId = m.view_id, c.u1 = m.u1, c.v1 = m.v1, c.d1 = m.d1, c.sizeU = m.size_u, c.sizeV = m.size_v;
    c.sizeU0 = m.size_u0, c.sizeV0 = m.size_v0, c.u0 = c.v0 = c.orientation = 0, c.pad = 0;
    c.depthOff = m.depth_offset, c.occOff = m.occ_offset;
    base[i]    = fs.totalElems;
    fs.totalElems += (long long)m.size_u0 * m.size_v0 * occRes * occRes;
    fs.maxPatchPixels = std::max( fs.maxPatchPixels, m.size_u * m.size_v );
    fs.maxPatchBlocks = std::max( fs.maxPatchBlocks, m.size_u0 * m.size_v0 );
  }
  base[P] = fs.totalElems;
  fs.dPatches.reserve( P ), fs.elemBase.reserve( P + 1 ), fs.packResult.reserve( 4 );
  PCC_CUDA( cudaMemcpyAsync( fs.dPatches, cp.data(), P * sizeof( CanvasPatch ), cudaMemcpyHostToDevice, s ) );
  PCC_CUDA( cudaMemcpyAsync( fs.elemBase, base.data(), ( P + 1 ) * sizeof( long long ), cudaMemcpyHostToDevice, s ) );
  const int rc = packPatches( fs.dPatches, P, fs.seg.occ, sizeU, sizeV, occRes, fs.packResult, s );
  if ( rc != PCCB200_OK ) throw std::runtime_error( "canvas wider than the packer supports" );
  int res[2] = { 0, 0 };
  PCC_CUDA( cudaMemcpyAsync( cp.data(), fs.dPatches, P * sizeof( CanvasPatch ), cudaMemcpyDeviceToHost, s ) );
  PCC_CUDA( cudaMemcpyAsync( res, fs.packResult, sizeof( res ), cudaMemcpyDeviceToHost, s ) );
  streamWait( s );
  if ( res[1] ) throw std::runtime_error( "patch packing exceeded the maximum canvas height" );
  for ( int i = 0; i < P; ++i ) fs.packed[i].u0 = cp[i].u0, fs.packed[i].v0 = cp[i].v0, fs.packed[i].orientation = cp[i].orientation;
  fs.heightPx = res[0];
}

// device-side records of a finished packing (fs.packed with u0 / v0 / orientation set): what a16..a22 read
void installPacking( FrameState& fs, int occRes ) {
  cudaStream_t s = fs.stream;
  const int    P = int( fs.packed.size() );
  fs.totalElems = 0, fs.maxPatchPixels = 1, fs.maxPatchBlocks = 1;
  if ( P == 0 ) return;
  std::vector<CanvasPatch> cp( P );
  std::vector<long long>   base( P + 1 );
  for ( int i = 0; i < P; ++i ) {
    const pccb200_patch& m = fs.packed[i];
    CanvasPatch&         c = cp[i];
    c.viewId = m.view_id, c.u1 = m.u1, c.v1 = m.v1, c.d1 = m.d1, c.sizeU = m.size_u, c.sizeV = m.size_v;
    c.sizeU0 = m.size_u0, c.sizeV0 = m.size_v0, c.u0 = m.u0, c.v0 = m.v0, c.orientation = m.orientation, c.pad = 0;
    c.depthOff = m.depth_offset, c.occOff = m.occ_offset;
    base[i]    = fs.totalElems;
    fs.totalElems += (long long)m.size_u0 * m.size_v0 * occRes * occRes;
    fs.maxPatchPixels = std::max( fs.maxPatchPixels, m.size_u * m.size_v );
    fs.maxPatchBlocks = std::max( fs.maxPatchBlocks, m.size_u0 * m.size_v0 );
  }
  base[P] = fs.totalElems;
  fs.dPatches.reserve( P ), fs.elemBase.reserve( P + 1 );
  PCC_CUDA( cudaMemcpyAsync( fs.dPatches, cp.data(), P * sizeof( CanvasPatch ), cudaMemcpyHostToDevice, s ) );
  PCC_CUDA( cudaMemcpyAsync( fs.elemBase, base.data(), ( P + 1 ) * sizeof( long long ), cudaMemcpyHostToDevice, s ) );
  streamWait( s );  // (cp / base are locals)
}

// a15: random-access packing of the whole GOF (constrainedPack + global patch allocation): every frame is placed against the
// previous one, so this runs once per GOF after all frames are segmented - metadata on the host, placement searches on the device.
// `all` = the patch records + block occupancies of ALL frames of the GOF in frame order; localOf[f] = index of frame f among the
// frames this GOF object holds, or -1 (sharded GOF: the frame lives on another rank). The packing is deterministic, so every rank
// that runs it on the same records gets the same result and installs the placements of its own frames.
void packRaAndInstall( pccb200_gof* g, std::vector<ra::Frame>& all, const std::vector<int>& localOf, int minW, int minH ) {
  const int occRes = g->prm.occupancy_resolution;
  if ( !packGofRandomAccess( all, occRes, size_t( minW ), size_t( minH ), g->ctx->raPack, &g->ctx->prof, g->ctx->stream ) )
    throw std::runtime_error( "random-access packing exceeded the packer's canvas limits" );
  for ( size_t f = 0; f < all.size(); ++f ) {
    if ( localOf[f] < 0 ) continue;
    FrameState& fs = *g->frames[localOf[f]];
    fs.packed.clear();
    std::vector<uint8_t> arena;
    for ( auto& p : all[f].patches ) {
      p.m.occ_offset = int64_t( arena.size() );
      arena.insert( arena.end(), p.occ.begin(), p.occ.end() );
      fs.packed.push_back( p.m );
    }
    fs.seg.occ.reserve( arena.size() + 1 );
    fs.seg.occElems = arena.size();
    if ( !arena.empty() ) {
      PCC_CUDA( cudaMemcpyAsync( fs.seg.occ, arena.data(), arena.size(), cudaMemcpyHostToDevice, fs.stream ) );
      streamWait( fs.stream );
    }
    fs.heightPx = int( all[f].height );
    fs.widthPx  = int( all[f].width );
    installPacking( fs, occRes );
  }
}

// the gof's own frames as packer input (patch records in creation order + their block occupancies from the device arena)
std::vector<ra::Frame> ownRaFrames( pccb200_gof* g ) {
  std::vector<ra::Frame> frames( g->nframes );
  for ( int f = 0; f < g->nframes; ++f ) {
    FrameState&          fs = *g->frames[f];
    std::vector<uint8_t> occ( fs.seg.occElems );
    if ( fs.seg.occElems ) {
      PCC_CUDA( cudaMemcpyAsync( occ.data(), fs.seg.occ, fs.seg.occElems, cudaMemcpyDeviceToHost, fs.stream ) );
      streamWait( fs.stream );
    }
    frames[f].patches.resize( fs.seg.patches.size() );
    for ( size_t i = 0; i < fs.seg.patches.size(); ++i ) {
      ra::Patch& p = frames[f].patches[i];
      p.m          = fs.seg.patches[i];
      p.m.best_match_idx = -1, p.m.is_global = 0;
      p.occ.assign( occ.begin() + p.m.occ_offset, occ.begin() + p.m.occ_offset + size_t( p.m.size_u0 ) * p.m.size_v0 );
    }
  }
  return frames;
}

void packGofRa( pccb200_gof* g, int minW, int minH ) {
  std::vector<ra::Frame> frames = ownRaFrames( g );
  std::vector<int>       localOf( g->nframes );
  for ( int f = 0; f < g->nframes; ++f ) localOf[f] = f;
  packRaAndInstall( g, frames, localOf, minW, minH );
}

// a14: one canvas size per GOF (PCCEncoder::resizeTileGeometryVideo + resizeGeometryVideo, PCCEncoder.cpp:5546-5634)
void setGofCanvas( pccb200_gof* g, size_t W, size_t H ) {
  g->W = size_t( std::ceil( double( W ) / 64.0 ) * 64 ), g->H = size_t( std::ceil( double( H ) / 64.0 ) * 64 );
}

// host copies of the patch lists (KBs of metadata + the per-patch maps downstream reference code reads). packedOrder: the records of
// fs.packed with the maps re-laid out in that order; otherwise (a GOF stopped after the segmentation) the records in creation order.
int hostPatchLists( pccb200_gof* g, bool packedOrder );

size_t copyOut( void* dst, const void* dev, size_t elems, size_t elemBytes, cudaStream_t s ) {
  if ( dst && elems ) {
    PCC_CUDA( cudaMemcpyAsync( dst, dev, elems * elemBytes, cudaMemcpyDeviceToHost, s ) );
    streamWait( s );
  }
  return elems;
}

__global__ void kUnpackXyz( const short4* __restrict__ in, int n, int16_t* __restrict__ out ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i < n ) out[3 * size_t( i )] = in[i].x, out[3 * size_t( i ) + 1] = in[i].y, out[3 * size_t( i ) + 2] = in[i].z;
}
__global__ void kNarrowU16( const uint16_t* __restrict__ in, size_t n, uint8_t* __restrict__ out ) {
  const size_t i = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( i < n ) out[i] = uint8_t( in[i] );  // (what PCCImage::write does for one byte per sample: a plain narrowing cast)
}
__global__ void kUnpackRgb( const uchar4* __restrict__ in, int n, uint8_t* __restrict__ out ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i < n ) out[3 * size_t( i )] = in[i].x, out[3 * size_t( i ) + 1] = in[i].y, out[3 * size_t( i ) + 2] = in[i].z;
}


// a16..a26 on a fixed canvas (g->W x g->H): occupancy + geometry images, reconstruction, colour transfer, attribute images
int runCanvasStages( pccb200_gof* g, int stopAfter ) {
  const int Wi = int( g->W ), Hi = int( g->H );
  int       rc = PCCB200_OK;
  const bool haveImages = g->stage >= 2;
  rc = forEachFrame( g, [&]( FrameState& fs, int ) {
        cudaStream_t s = fs.stream;
        ScratchLease lease( fs, g->ctx->device, "hold_canvas" );
        if ( !haveImages ) {
          {
            ProfScope t( &fs.prof, "images", s );
            formOccupancyAndGeometry( fs.dPatches, int( fs.packed.size() ), fs.maxPatchPixels, fs.maxPatchBlocks, fs.seg.depth, g->prm.occupancy_resolution,
                                      g->occPrec, Wi, Hi, fs.im, s );
          }
          int err = 0;
          PCC_CUDA( cudaMemcpyAsync( &err, fs.im.error, sizeof( int ), cudaMemcpyDeviceToHost, s ) );
          streamWait( s );
          if ( err ) throw std::runtime_error( "patch2Canvas out of the canvas" );
        } else if ( fs.decodedSet ) {
          // decoded occupancy video: a18 again (generateBlockToPatchFromOccupancyMapVideo runs on the decoded frame, :168)
          blockToPatchFromVideo( fs.dPatches, int( fs.packed.size() ), fs.maxPatchBlocks, g->prm.occupancy_resolution, g->occPrec, Wi, Hi, fs.im.om,
                                 fs.im.blockToPatch, s );
        }
        if ( stopAfter == 2 ) return;
        // the occupancy and geometry videos are coded losslessly / passed through here: decoded == source
        {
          ProfScope t( &fs.prof, "reconstruct", s );
          reconstructPoints( fs.dPatches, fs.elemBase, int( fs.packed.size() ), fs.totalElems, g->prm.occupancy_resolution, g->occPrec, Wi, Hi, fs.im.om,
                             fs.im.blockToPatch, fs.im.geo0, fs.im.geo1, fs.rc, fs.fat->recon, s );
        }
        if ( stopAfter == 3 ) return;
        const size_t R = fs.rc.numPoints;
        fs.recRgb.reserve( R + 1 );
        {
          ProfScope t( &fs.prof, "color_transfer", s );
          transferColors( fs.fat->color, fs.tree, fs.xyz4, fs.rgb4, fs.n, fs.rc.recXyz, R, fs.recRgb, s );
        }
        {
          ProfScope t( &fs.prof, "attribute_images", s );
          formAttributeImages( fs.rc.pointToPixel, fs.recRgb, R, fs.im.om, Wi, Hi, g->occPrec, fs.attr, fs.fat->attr, s );
        }
      } );
  if ( rc != PCCB200_OK ) return rc;
  g->stage = stopAfter == 2 ? 2 : ( stopAfter == 3 ? 3 : 4 );
  return PCCB200_OK;
}

void collectProfiles( pccb200_gof* g ) {
  g->ctx->prof.collect( g->ctx->stream );  // GOF-level spans (the batched orientation walk)
  for ( auto* fs : g->frames ) {
    fs->prof.collect( fs->stream );
    for ( auto& r : fs->prof.results ) g->ctx->prof.results.push_back( r );
    fs->prof.results.clear();
  }
}

int hostPatchLists( pccb200_gof* g, bool packedOrder ) {
  g->lists.assign( g->nframes, pccb200_patchlist() );
  return forEachFrame( g, [&]( FrameState& fs, int f ) {
    pccb200_patchlist& pl = g->lists[f];
    pl.patches            = packedOrder ? fs.packed : fs.seg.patches;
    std::vector<int16_t> depth( fs.seg.depthElems );
    std::vector<uint8_t> occ( fs.seg.occElems );
    {
      ProfScope t( &fs.prof, "d2h_patches", fs.stream );
      if ( fs.seg.depthElems ) PCC_CUDA( cudaMemcpyAsync( depth.data(), fs.seg.depth, fs.seg.depthElems * sizeof( int16_t ), cudaMemcpyDeviceToHost, fs.stream ) );
      if ( fs.seg.occElems ) PCC_CUDA( cudaMemcpyAsync( occ.data(), fs.seg.occ, fs.seg.occElems, cudaMemcpyDeviceToHost, fs.stream ) );
      streamWait( fs.stream );
    }
    // the device arenas are in creation order; hand the maps out in the order of the patch records
    pl.depth.resize( fs.seg.depthElems ), pl.occ.resize( fs.seg.occElems );
    size_t dOff = 0, oOff = 0;
    for ( auto& m : pl.patches ) {
      const size_t px = 2 * size_t( m.size_u ) * m.size_v, nb = size_t( m.size_u0 ) * m.size_v0;
      std::copy( depth.begin() + m.depth_offset, depth.begin() + m.depth_offset + px, pl.depth.begin() + dOff );
      std::copy( occ.begin() + m.occ_offset, occ.begin() + m.occ_offset + nb, pl.occ.begin() + oOff );
      m.depth_offset = int64_t( dOff ), m.occ_offset = int64_t( oOff );
      dOff += px, oOff += nb;
    }
  } );
}

}  // namespace

extern "C" {

// Sharded random access: the frames of a GOF live on several ranks, but every frame is packed against the previous
// one (PCCEncoder.cpp:4778-4805) and the global patch allocation iterates over the whole GOF (:6838-6970). Each rank segments its
// frames (pccb200_encode_gof with stop_after = 5), the ranks all-gather the patch records + block occupancies (KBs per frame:
// pccb200_gof_patches / pccb200_patches_get with a NULL depth pointer), and every rank calls this with the records of ALL frames
// in frame order: the packing is deterministic, so all ranks compute the same placements; each installs those of its own frames.
//   patch_counts[f]            patches of frame f (f = 0 .. total_frames-1)
//   patches                    the records of all frames, concatenated; occ_offset relative to the frame's own occupancy block
//   occ, occ_sizes[f]          the frames' block occupancies, concatenated / bytes per frame
//   local_frame[f]             index of frame f among the frames of `gof`, or -1 if another rank holds it
// Afterwards the GOF is in the state pccb200_encode_gof( stop_after = 1 ) leaves (pccb200_gof_dims = the GOF-wide canvas).
int pccb200_gof_pack_ra( pccb200_gof* g, int totalFrames, const int* patchCounts, const pccb200_patch* patches, const uint8_t* occ,
                         const size_t* occSizes, const int* localFrame ) {
  if ( !g || totalFrames < g->nframes || !patchCounts || !occSizes || !localFrame ) return PCCB200_ERR_BAD_ARG;
  if ( g->stage != 0 || g->prm.global_patch_allocation == 0 ) return PCCB200_ERR_STATE;
  return guarded( g->ctx, [&]() -> int {
    std::vector<ra::Frame> all( totalFrames );
    std::vector<int>       localOf( totalFrames, -1 ), seen( g->nframes, 0 );
    size_t                 pAt = 0, oAt = 0;
    for ( int f = 0; f < totalFrames; ++f ) {
      if ( patchCounts[f] < 0 || ( patchCounts[f] && ( !patches || !occ ) ) ) return PCCB200_ERR_BAD_ARG;
      if ( localFrame[f] >= g->nframes ) return PCCB200_ERR_BAD_ARG;
      if ( localFrame[f] >= 0 ) {
        if ( seen[localFrame[f]]++ || size_t( patchCounts[f] ) != g->frames[localFrame[f]]->seg.patches.size() ) return PCCB200_ERR_BAD_ARG;
        localOf[f] = localFrame[f];
      }
      all[f].patches.resize( patchCounts[f] );
      for ( int i = 0; i < patchCounts[f]; ++i ) {
        ra::Patch& p = all[f].patches[i];
        p.m          = patches[pAt + i];
        const size_t nb = size_t( p.m.size_u0 ) * p.m.size_v0;
        if ( p.m.size_u0 < 0 || p.m.size_v0 < 0 || p.m.occ_offset < 0 || size_t( p.m.occ_offset ) + nb > occSizes[f] ) return PCCB200_ERR_BAD_ARG;
        p.occ.assign( occ + oAt + p.m.occ_offset, occ + oAt + p.m.occ_offset + nb );
        p.m.best_match_idx = -1, p.m.is_global = 0;
        if ( localOf[f] >= 0 ) {  // the device arenas of a local frame are addressed by ITS records (the exchanged copies are re-based)
          const pccb200_patch& own = g->frames[localOf[f]]->seg.patches[i];
          if ( own.index != p.m.index || own.size_u != p.m.size_u || own.size_v != p.m.size_v ) return PCCB200_ERR_BAD_ARG;
          p.m.depth_offset = own.depth_offset;
        }
      }
      pAt += size_t( patchCounts[f] ), oAt += occSizes[f];
    }
    for ( int f = 0; f < g->nframes; ++f )
      if ( !seen[f] ) return PCCB200_ERR_BAD_ARG;
    const int minW = g->prm.geometry_bitdepth_3d > 11 ? 2560 : 1280, minH = 1280;
    try {
      packRaAndInstall( g, all, localOf, minW, minH );
    } catch ( const std::exception& e ) {
      g->ctx->lastError = e.what();
      return PCCB200_ERR_CUDA;
    }
    size_t W = minW, H = minH;
    for ( auto& fr : all ) H = std::max( H, fr.height ), W = std::max( W, fr.width );  // (all frames are here: the GOF-wide maximum)
    setGofCanvas( g, W, H );
    g->stage = 1;
    const int rc = hostPatchLists( g, true );
    collectProfiles( g );
    return rc;
  } );
}

int pccb200_encode_gof( pccb200_ctx* ctx, int nframes, const int16_t* const* xyz, const uint8_t* const* rgb, const size_t* n,
                        const pccb200_seg_params* prm, int occupancyPrecision, int stopAfter, pccb200_gof** out ) {
  if ( !ctx || !out || nframes < 0 || ( nframes > 0 && ( !xyz || !rgb || !n ) ) || !prm ) return PCCB200_ERR_BAD_ARG;
  if ( !segParamsSupported( *prm ) || prm->occupancy_resolution != 16 || occupancyPrecision < 1 || 16 % occupancyPrecision != 0 ||
       prm->map_count_minus1 != 1 || ( prm->global_patch_allocation != 0 && prm->global_patch_allocation != 1 ) )
    return PCCB200_ERR_UNSUPPORTED;
  if ( stopAfter < 0 || stopAfter == 4 || stopAfter > 5 || ( stopAfter == 5 && prm->global_patch_allocation == 0 ) ) return PCCB200_ERR_BAD_ARG;
  *out = nullptr;
  return guarded( ctx, [&]() -> int {
    pccb200_gof* g = new pccb200_gof();
    g->ctx = ctx, g->nframes = nframes, g->occPrec = occupancyPrecision, g->prm = *prm;
    while ( int( ctx->framePool.size() ) < nframes ) {
      ctx->framePool.emplace_back( new FrameState() );
      PCC_CUDA( cudaStreamCreateWithFlags( &ctx->framePool.back()->stream, cudaStreamNonBlocking ) );
    }
    if ( ctx->prof.enabled ) {  // common time origin of all frame streams
      if ( !ctx->prof.origin ) PCC_CUDA( cudaEventCreate( &ctx->prof.origin ) );
      PCC_CUDA( cudaEventRecord( ctx->prof.origin, ctx->stream ) );
      PCC_CUDA( cudaEventSynchronize( ctx->prof.origin ) );
    }
    for ( int f = 0; f < nframes; ++f ) {
      FrameState* fs = ctx->framePool[f].get();
      fs->hXyz = xyz[f], fs->hRgb = rgb[f], fs->n = n[f], fs->status = 0, fs->error.clear();
      fs->prof.enabled = ctx->prof.enabled;
      fs->decodedSet   = false;
      fs->prof.origin  = ctx->prof.enabled ? ctx->prof.origin : nullptr;
      g->frames.push_back( fs );
    }
    const int minW = prm->geometry_bitdepth_3d > 11 ? 2560 : 1280, minH = 1280;  // minimumImageWidth/Height (cfg/sequence/*_vox11.cfg: 2560)
    WalkGate gate;
    gate.expected  = nframes;
    int       rc   = forEachFrame( g, [&]( FrameState& fs, int ) {
      std::exception_ptr early;
      try {
        ScratchLease lease( fs, ctx->device, "hold_pre" );  // (drains the stream before the set is handed on)
        segmentFrameBeforeWalk( fs, g->prm );
        streamWait( fs.stream );
      } catch ( ... ) {
        early = std::current_exception();  // (this frame still has to show up at the gate, or the others would wait forever)
      }
      gate.arrive( early || fs.orient.walkSmem == 0 ? nullptr : &fs.orient, ctx );
      if ( early ) std::rethrow_exception( early );
      {
        ScratchLease lease( fs, ctx->device, "hold_post" );
        segmentFrameAfterWalk( fs, g->prm );
        streamWait( fs.stream );
      }
      if ( g->prm.global_patch_allocation == 0 && stopAfter != 5 ) packFrame( fs, g->prm, minW, minH, 2, 1.0 );
    } );
    if ( rc == PCCB200_OK && g->prm.global_patch_allocation != 0 && stopAfter != 5 ) {
      try {
        packGofRa( g, minW, minH );
      } catch ( const std::exception& e ) {
        ctx->lastError = e.what();
        rc             = PCCB200_ERR_CUDA;
      } catch ( const CudaError& e ) {
        char buf[256];
        snprintf( buf, sizeof( buf ), "random-access packing: CUDA error %d (%s) at %s:%d", int( e.code ), cudaGetErrorString( e.code ), e.file, e.line );
        ctx->lastError = buf;
        rc             = PCCB200_ERR_CUDA;
        cudaGetLastError();
      }
    }
    if ( rc != PCCB200_OK ) {
      delete g;
      return rc;
    }
    if ( stopAfter == 5 ) {  // sharded random access: the packing needs the other ranks' patch records first (pccb200_gof_pack_ra)
      g->stage = 0;
      rc       = hostPatchLists( g, false );
    } else {
      size_t W = minW, H = minH;
      for ( auto* fs : g->frames ) H = std::max( H, size_t( fs->heightPx ) ), W = std::max( W, size_t( fs->widthPx ) );
      setGofCanvas( g, W, H );
      g->stage = 1;
      if ( stopAfter != 1 ) {
        rc = runCanvasStages( g, stopAfter );
        if ( rc != PCCB200_OK ) {
          delete g;
          return rc;
        }
      }
      rc = hostPatchLists( g, true );
    }
    collectProfiles( g );
    if ( rc != PCCB200_OK ) {
      delete g;
      return rc;
    }
    *out = g;
    return PCCB200_OK;
  } );
}

int pccb200_gof_resume( pccb200_gof* g, size_t width, size_t height, int stopAfter ) {
  if ( !g || stopAfter == 1 ) return PCCB200_ERR_BAD_ARG;
  // A finished GOF (stage 4) may be formed AGAIN on a larger canvas: a rank that went ahead with its local canvas size while the
  // GOF-wide maximum was still being reduced (bench.py, sharded frames) repeats a16..a26 when another rank needed more rows. All
  // inputs of the canvas stages (packed patch records, depth arenas, source cloud + tree) are still resident.
  if ( g->stage == 4 && stopAfter == 0 && width >= g->W && height >= g->H && ( width != g->W || height != g->H ) && width % 64 == 0 && height % 64 == 0 ) {
    for ( auto* fs : g->frames ) fs->decodedSet = false;
    g->stage = 1;
  }
  if ( g->stage != 1 && g->stage != 2 ) return PCCB200_ERR_STATE;
  if ( width < g->W || height < g->H || width % 64 || height % 64 ) return PCCB200_ERR_BAD_ARG;
  if ( g->stage == 2 && ( width != g->W || height != g->H || stopAfter == 2 ) ) return PCCB200_ERR_BAD_ARG;
  return guarded( g->ctx, [&]() -> int {
    g->W = width, g->H = height;
    const int rc = runCanvasStages( g, stopAfter );
    collectProfiles( g );
    return rc;
  } );
}

int pccb200_gof_set_decoded( pccb200_gof* g, int f, const uint8_t* occVideo, const uint16_t* geo0, const uint16_t* geo1 ) {
  if ( !g || f < 0 || f >= g->nframes ) return PCCB200_ERR_BAD_ARG;
  if ( g->stage != 2 ) return PCCB200_ERR_STATE;
  return guarded( g->ctx, [&]() -> int {
    FrameState&  fs = *g->frames[f];
    const size_t Q = g->W * g->H, cells = ( g->W / g->occPrec ) * ( g->H / g->occPrec );
    if ( occVideo ) PCC_CUDA( cudaMemcpyAsync( fs.im.om, occVideo, cells, cudaMemcpyHostToDevice, fs.stream ) );
    if ( geo0 ) PCC_CUDA( cudaMemcpyAsync( fs.im.geo0, geo0, Q * 2, cudaMemcpyHostToDevice, fs.stream ) );
    if ( geo1 ) PCC_CUDA( cudaMemcpyAsync( fs.im.geo1, geo1, Q * 2, cudaMemcpyHostToDevice, fs.stream ) );
    streamWait( fs.stream );
    fs.decodedSet = fs.decodedSet || occVideo != nullptr;
    return PCCB200_OK;
  } );
}

// PCCCodec::generatePointCloud as the decoder calls it (PccLibDecoder/source/PCCDecoder.cpp:334-351): patches rebuilt from the
// atlas syntax, decoded occupancy + geometry frames in, reconstructed cloud out.
int pccb200_generate_point_cloud( pccb200_ctx* ctx, const pccb200_patch* patches, int numPatches, const uint8_t* occVideo, const uint16_t* geo0,
                                  const uint16_t* geo1, size_t width, size_t height, int occupancyPrecision, size_t capacity, int16_t* xyz,
                                  uint32_t* pointToPixel, uint32_t* partition, uint16_t* boundary, size_t* recPoints ) {
  if ( !ctx || numPatches < 0 || ( numPatches && !patches ) || !occVideo || !geo0 || !geo1 || !recPoints || width == 0 || height == 0 ||
       width % 16 || height % 16 || width > 65536 || height > 65536 || occupancyPrecision < 1 || 16 % occupancyPrecision )
    return PCCB200_ERR_BAD_ARG;
  // The patch records come from a bitstream (PCCDecoder.cpp:900-1040): nothing in them is trusted. Every patch must lie inside
  // the canvas in 16-pixel blocks, in its orientation (PCCPatch::patchBlock2CanvasBlock, PCCPatch.cpp:253-308).
  for ( int i = 0; i < numPatches; ++i ) {
    const pccb200_patch& m = patches[i];
    if ( m.view_id < 0 || m.view_id > 5 || ( m.orientation != 0 && m.orientation != 1 ) ) return PCCB200_ERR_UNSUPPORTED;
    const long long bw = (long long)( width / 16 ), bh = (long long)( height / 16 );
    const long long su = m.orientation == 0 ? m.size_u0 : m.size_v0, sv = m.orientation == 0 ? m.size_v0 : m.size_u0;
    if ( m.u0 < 0 || m.v0 < 0 || m.size_u0 <= 0 || m.size_v0 <= 0 || m.u0 + su > bw || m.v0 + sv > bh ) return PCCB200_ERR_BAD_ARG;
  }
  return guarded( ctx, [&]() -> int {
    if ( ctx->framePool.empty() ) {
      ctx->framePool.emplace_back( new FrameState() );
      PCC_CUDA( cudaStreamCreateWithFlags( &ctx->framePool.back()->stream, cudaStreamNonBlocking ) );
    }
    FrameState&  fs = *ctx->framePool[0];
    cudaStream_t s  = fs.stream;
    const int    W = int( width ), H = int( height ), occRes = 16, P = numPatches;
    const size_t Q = width * height, cells = ( width / occupancyPrecision ) * ( height / occupancyPrecision ), blocks = ( width / 16 ) * ( height / 16 );
    std::vector<CanvasPatch> cp( P );
    std::vector<long long>   base( P + 1 );
    long long                total = 0;
    int                      maxBlocks = 1;
    for ( int i = 0; i < P; ++i ) {
      const pccb200_patch& m = patches[i];
      CanvasPatch& c = cp[i];
      c.viewId = m.view_id, c.u1 = m.u1, c.v1 = m.v1, c.d1 = m.d1, c.sizeU = m.size_u, c.sizeV = m.size_v, c.sizeU0 = m.size_u0, c.sizeV0 = m.size_v0;
      c.u0 = m.u0, c.v0 = m.v0, c.orientation = m.orientation, c.pad = 0, c.depthOff = c.occOff = 0;
      base[i] = total;
      total += (long long)m.size_u0 * m.size_v0 * occRes * occRes;
      maxBlocks = std::max( maxBlocks, m.size_u0 * m.size_v0 );
    }
    base[P] = total;
    fs.dPatches.reserve( P + 1 ), fs.elemBase.reserve( P + 2 );
    fs.im.om.reserve( cells ), fs.im.geo0.reserve( Q ), fs.im.geo1.reserve( Q ), fs.im.blockToPatch.reserve( blocks );
    if ( P ) PCC_CUDA( cudaMemcpyAsync( fs.dPatches, cp.data(), P * sizeof( CanvasPatch ), cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( fs.elemBase, base.data(), ( P + 1 ) * sizeof( long long ), cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( fs.im.om, occVideo, cells, cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( fs.im.geo0, geo0, Q * 2, cudaMemcpyHostToDevice, s ) );
    PCC_CUDA( cudaMemcpyAsync( fs.im.geo1, geo1, Q * 2, cudaMemcpyHostToDevice, s ) );
    blockToPatchFromVideo( fs.dPatches, P, maxBlocks, occRes, occupancyPrecision, W, H, fs.im.om, fs.im.blockToPatch, s );
    const size_t R = reconstructPoints( fs.dPatches, fs.elemBase, P, total, occRes, occupancyPrecision, W, H, fs.im.om, fs.im.blockToPatch, fs.im.geo0,
                                        fs.im.geo1, fs.rc, ctx->own.recon, s );
    *recPoints = R;
    if ( R > capacity ) return ( xyz || pointToPixel || partition || boundary ) ? PCCB200_ERR_CAPACITY : PCCB200_OK;
    if ( R ) {
      if ( xyz ) {
        fs.xyzRaw.reserve( 3 * R );
        kUnpackXyz<<<divUp( R, 256 ), 256, 0, s>>>( fs.rc.recXyz, int( R ), fs.xyzRaw );
        PCC_CUDA( cudaMemcpyAsync( xyz, fs.xyzRaw, 3 * R * 2, cudaMemcpyDeviceToHost, s ) );
      }
      if ( pointToPixel ) PCC_CUDA( cudaMemcpyAsync( pointToPixel, fs.rc.pointToPixel, 3 * R * 4, cudaMemcpyDeviceToHost, s ) );
      if ( partition ) PCC_CUDA( cudaMemcpyAsync( partition, fs.rc.recPartition, R * 4, cudaMemcpyDeviceToHost, s ) );
      if ( boundary ) PCC_CUDA( cudaMemcpyAsync( boundary, fs.rc.boundary, R * 2, cudaMemcpyDeviceToHost, s ) );
    }
    streamWait( s );
    return PCCB200_OK;
  } );
}

void pccb200_gof_free( pccb200_gof* g ) { delete g; }

int pccb200_gof_dims( const pccb200_gof* g, int f, size_t* w, size_t* h, size_t* recPoints ) {
  if ( !g || f < 0 || f >= g->nframes ) return PCCB200_ERR_BAD_ARG;
  if ( w ) *w = g->W;
  if ( h ) *h = g->H;
  if ( recPoints ) *recPoints = g->stage >= 3 ? g->frames[f]->rc.numPoints : 0;
  return PCCB200_OK;
}

const pccb200_patchlist* pccb200_gof_patches( const pccb200_gof* g, int f ) {
  if ( !g || f < 0 || f >= g->nframes ) return nullptr;
  return &g->lists[f];
}

size_t pccb200_gof_get( pccb200_gof* g, int f, int what, void* dst ) {
  if ( !g || f < 0 || f >= g->nframes ) return 0;
  FrameState&  fs = *g->frames[f];
  cudaStream_t s  = fs.stream;
  size_t       result = 0;
  guarded( g->ctx, [&]() -> int {
    const size_t Q = g->W * g->H, cells = ( g->W / g->occPrec ) * ( g->H / g->occPrec ), blocks = ( g->W / 16 ) * ( g->H / 16 );
    const size_t R = g->stage >= 3 ? fs.rc.numPoints : 0;
    if ( ( ( what >= 1 && what <= 5 ) || what == PCCB200_GOF_GEO0_LUMA8 || what == PCCB200_GOF_GEO1_LUMA8 ) && g->stage < 2 ) return 0;
    if ( what >= 6 && what <= 9 && g->stage < 3 ) return 0;
    if ( what >= 10 && what <= 16 && g->stage < 4 ) return 0;
    switch ( what ) {
      case PCCB200_GOF_OCCUPANCY: result = copyOut( dst, fs.im.occ, Q, 1, s ); break;
      case PCCB200_GOF_OM_VIDEO: result = copyOut( dst, fs.im.om, cells, 1, s ); break;
      case PCCB200_GOF_BLOCK_TO_PATCH: result = copyOut( dst, fs.im.blockToPatch, blocks, 4, s ); break;
      case PCCB200_GOF_GEO0: result = copyOut( dst, fs.im.geo0, Q, 2, s ); break;
      case PCCB200_GOF_GEO1: result = copyOut( dst, fs.im.geo1, Q, 2, s ); break;
      case PCCB200_GOF_REC_XYZ:
        if ( dst && R ) {
          fs.xyzRaw.reserve( 3 * R );
          kUnpackXyz<<<divUp( R, 256 ), 256, 0, s>>>( fs.rc.recXyz, int( R ), fs.xyzRaw );
          copyOut( dst, fs.xyzRaw, 3 * R, 2, s );
        }
        result = 3 * R;
        break;
      case PCCB200_GOF_POINT_TO_PIXEL: result = copyOut( dst, fs.rc.pointToPixel, 3 * R, 4, s ); break;
      case PCCB200_GOF_REC_PARTITION: result = copyOut( dst, fs.rc.recPartition, R, 4, s ); break;
      case PCCB200_GOF_REC_BOUNDARY: result = copyOut( dst, fs.rc.boundary, R, 2, s ); break;
      case PCCB200_GOF_REC_RGB:
        if ( dst && R ) {
          fs.rgbRaw.reserve( 3 * R );
          kUnpackRgb<<<divUp( R, 256 ), 256, 0, s>>>( fs.recRgb, int( R ), fs.rgbRaw );
          copyOut( dst, fs.rgbRaw, 3 * R, 1, s );
        }
        result = 3 * R;
        break;
      case PCCB200_GOF_ATTR0_RAW: result = copyOut( dst, fs.attr.rawPlanes[0], 3 * Q, 2, s ); break;
      case PCCB200_GOF_ATTR1_RAW: result = copyOut( dst, fs.attr.rawPlanes[1], 3 * Q, 2, s ); break;
      case PCCB200_GOF_ATTR0: result = copyOut( dst, fs.attr.planes[0], 3 * Q, 2, s ); break;
      case PCCB200_GOF_ATTR1: result = copyOut( dst, fs.attr.planes[1], 3 * Q, 2, s ); break;
      case PCCB200_GOF_GEO0_LUMA8:
      case PCCB200_GOF_GEO1_LUMA8:
        result = Q;
        if ( dst ) {
          fs.geoLuma8.reserve( Q );
          kNarrowU16<<<divUp( Q, 256 ), 256, 0, s>>>( what == PCCB200_GOF_GEO0_LUMA8 ? fs.im.geo0.p : fs.im.geo1.p, Q, fs.geoLuma8 );
          copyOut( dst, fs.geoLuma8, Q, 1, s );
        }
        break;
      case PCCB200_GOF_ATTR0_YUV420:
      case PCCB200_GOF_ATTR1_YUV420: {
        const int m = what == PCCB200_GOF_ATTR0_YUV420 ? 0 : 1;
        result      = Q + 2 * ( ( g->W / 2 ) * ( g->H / 2 ) );
        if ( dst ) {  // converted on request: the frame leaves the device with 1.5 bytes per pixel
          fs.yuv.out[m].reserve( result );
          rgbPlanesToYuv420( fs.attr.planes[m], int( g->W ), int( g->H ), fs.yuv, fs.yuv.out[m], s );
          copyOut( dst, fs.yuv.out[m], result, 1, s );
        }
        break;
      }
      default: break;
    }
    return 0;
  } );
  return result;
}

}  // extern "C"
