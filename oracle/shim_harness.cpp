// shim_harness.cpp — TEST INFRASTRUCTURE: runs the reference's encoder harness (ref_harness.cpp, libtmc2ref.so) with its hot-path
// stages replaced by the reference-side shim of libpccb200 (integration/pccb200_shim.cpp), so that a test can compare what the
// reference's own data structures (PCCPatch, PCCFrameContext, PCCImage, PCCPointSet3) hold after the shim filled them with what
// the unmodified reference stages leave there. Built by oracle/Makefile (target `shim`) into oracle/_ref/libtmc2shim.so, linked
// against libtmc2ref.so and libpccb200.so; needs a GPU to run.
#include "PCCCommon.h"
#include "PCCFrameContext.h"
#include "PCCContext.h"
#include "PCCGroupOfFrames.h"
#include "PCCEncoderParameters.h"
#include "PCCEncoder.h"

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
}
