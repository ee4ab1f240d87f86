// shim_harness.cpp — TEST INFRASTRUCTURE: runs the reference's encoder harness (ref_harness.cpp, libtmc2ref.so) with its hot-path
// stages replaced by the reference-side shim of libpccb200 (integration/pccb200_shim.cpp), so that a test can compare what the
// reference's own data structures (PCCPatch, PCCFrameContext, PCCImage, PCCPointSet3) hold after the shim filled them with what
// the unmodified reference stages leave there. Built by oracle/Makefile (target `shim`) into oracle/_ref/libtmc2shim.so, linked
// against libtmc2ref.so and libpccb200.so; needs a GPU to run
// (except shim_load_compare, the input side: host code).
#include "PCCCommon.h"
#include "PCCFrameContext.h"
#include "PCCContext.h"
#include "PCCGroupOfFrames.h"
#include "PCCEncoderParameters.h"
#include "PCCEncoder.h"

#include <fcntl.h>
#include <unistd.h>

#include <chrono>
#include <iostream>

#include "../integration/pccb200_shim.h"

using namespace pcc;

struct RefHotPathHooks {
  int ( *stageA )( PCCEncoder& enc, PCCGroupOfFrames& sources, PCCContext& context );
  int ( *stageB1 )( PCCEncoder& enc, PCCContext& context, PCCGroupOfFrames& reconstructs, std::vector<std::vector<uint32_t>>& partitions );
  int ( *stageB2 )( PCCEncoder& enc, PCCContext& context );
};
extern "C" void  ref_set_hot_path_hooks( const RefHotPathHooks* hooks );
extern "C" void* ref_encode_gof( int nframes, const int16_t* const* xyz, const uint8_t* const* rgb, const size_t* n, const pccb200_seg_params* p,
                                 int occupancyPrecision, int stopAfter );

namespace {
pccb200shim::Session* gSession  = nullptr;
int                   gLastCode = 0;
int hookA( PCCEncoder& enc, PCCGroupOfFrames& sources, PCCContext& context ) {
  return gLastCode = pccb200shim::stageA( *gSession, enc.params_, sources, context );
}
int hookB1( PCCEncoder& enc, PCCContext& context, PCCGroupOfFrames& reconstructs, std::vector<std::vector<uint32_t>>& partitions ) {
  return gLastCode = pccb200shim::stageB1( *gSession, enc.params_, context, reconstructs, partitions );
}
int hookB2( PCCEncoder&, PCCContext& context ) { return gLastCode = pccb200shim::stageB2( *gSession, context ); }
const RefHotPathHooks kHooks = {hookA, hookB1, hookB2};
// decoder-side binding: the reference's own encoder stages up to the geometry video, then every frame reconstructed by
// pccb200shim::decodeFrame (what PCCDecoder::decode would call in place of generatePointCloud)
int hookDecode( PCCEncoder& enc, PCCContext& context, PCCGroupOfFrames& reconstructs, std::vector<std::vector<uint32_t>>& partitions ) {
  for ( size_t f = 0; f < context.size(); ++f ) {
    PCCPointSet3 rec;
    gLastCode = pccb200shim::decodeFrame( *gSession, context, f, enc.params_.occupancyPrecision_, rec, partitions[f] );
    if ( gLastCode != 0 ) return gLastCode;
    reconstructs[f].clear();
    reconstructs[f].appendPointSet( rec );
  }
  return 0;
}
const RefHotPathHooks kDecodeHooks = {nullptr, hookDecode, nullptr};
}  // namespace

extern "C" {
// same contract as ref_encode_gof (the handle works with ref_gof_*), hot path through libpccb200; *code receives the last status
void* shim_encode_gof( int nframes, const int16_t* const* xyz, const uint8_t* const* rgb, const size_t* n, const pccb200_seg_params* p,
                       int occupancyPrecision, int stopAfter, int* code ) {
  pccb200shim::Session session;
  gSession  = &session;
  gLastCode = 0;
  ref_set_hot_path_hooks( &kHooks );
  void* h = ref_encode_gof( nframes, xyz, rgb, n, p, occupancyPrecision, stopAfter );
  ref_set_hot_path_hooks( nullptr );
  gSession = nullptr;
  if ( code ) *code = gLastCode;
  return h;
}
// reference stages up to the geometry images, reconstruction through the decoder-side binding (use stopAfter = 3)
void* shim_decode_gof( int nframes, const int16_t* const* xyz, const uint8_t* const* rgb, const size_t* n, const pccb200_seg_params* p,
                       int occupancyPrecision, int stopAfter, int* code ) {
  pccb200shim::Session session;
  gSession  = &session;
  gLastCode = 0;
  ref_set_hot_path_hooks( &kDecodeHooks );
  void* h = ref_encode_gof( nframes, xyz, rgb, n, p, occupancyPrecision, stopAfter );
  ref_set_hot_path_hooks( nullptr );
  gSession = nullptr;
  if ( code ) *code = gLastCode;
  return h;
}
// PCCGroupOfFrames::load by the reference and by pccb200shim::loadFrames on the same files. Returns 0 when both give the same
// return value, frame count and, frame by frame, point count / colour flag / positions / colours; otherwise 1000 * (frame + 1) +
// what differed (1 count, 2 flags, 3 positions, 4 colours), -1 return values, -2 frame counts. *frames / *points: what was loaded.
int shim_load_compare( const char* pattern, size_t start, size_t end, int colorTransform, size_t nbThread, size_t* frames, size_t* points,
                       double* secondsReference, double* secondsShim ) {
  PCCGroupOfFrames a, b;
  a.setFrameCount( 1 );  // (a reused group: load resizes it)
  b.setFrameCount( 1 );
  const auto t0 = std::chrono::steady_clock::now();
  bool       ra;
  {
    std::cout.setstate( std::ios_base::failbit );  // (the reference prints "Error: can't open ..." for a missing frame)
    ra = a.load( pattern, start, end, PCCColorTransform( colorTransform ), false, nbThread );
    std::cout.clear();
  }
  const auto t1 = std::chrono::steady_clock::now();
  fflush( stdout );
  const int keep = dup( 1 ), null = open( "/dev/null", O_WRONLY );
  dup2( null, 1 );
  const bool rb = pccb200shim::loadFrames( b, pattern, start, end, PCCColorTransform( colorTransform ), nbThread );
  fflush( stdout );
  dup2( keep, 1 );
  close( keep );
  close( null );
  const auto t2 = std::chrono::steady_clock::now();
  if ( secondsReference ) *secondsReference = std::chrono::duration<double>( t1 - t0 ).count();
  if ( secondsShim ) *secondsShim = std::chrono::duration<double>( t2 - t1 ).count();
  if ( frames ) *frames = b.getFrameCount();
  if ( points ) *points = 0;
  if ( ra != rb ) return -1;
  if ( a.getFrameCount() != b.getFrameCount() ) return -2;
  for ( size_t f = 0; f < a.getFrameCount(); ++f ) {
    const PCCPointSet3 &x = a[f], &y = b[f];
    const int           base = 1000 * int( f + 1 );
    if ( x.getPointCount() != y.getPointCount() ) return base + 1;
    if ( x.hasColors() != y.hasColors() || y.hasNormals() || y.hasReflectances() ) return base + 2;
    if ( points ) *points += y.getPointCount();
    for ( size_t i = 0; i < x.getPointCount(); ++i ) {
      if ( !( x[i] == y[i] ) ) return base + 3;
      if ( x.hasColors() && !( x.getColor( i ) == y.getColor( i ) ) ) return base + 4;
    }
  }
  return 0;
}
}
