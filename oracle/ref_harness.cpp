// ref_harness.cpp — OUR thin C-ABI wrapper around the UNMODIFIED reference classes (TEST INFRASTRUCTURE ONLY).
//
// Compiled by oracle/Makefile together with the reference's own sources (taken where they lie under
// /root/reference) into oracle/_ref/libtmc2ref.so.  It lets tests pin the CPU restatement (pcc_oracle.cpp)
// and the CUDA path against the reference itself, stage by stage, and lets bench.py time the reference's own
// CPU implementation of the hot path (`--impl reference`, cpu_baseline.kind = "reference").
// Nothing here re-implements an algorithm: every function converts plain arrays to the reference's containers,
// calls the reference's own methods and converts back.
// (built with -fno-access-control so the harness can call the reference's private stage methods)
// The reference's PCCEncoder.cpp, unmodified, is part of this translation unit (see oracle/Makefile).
#include "PCCEncoder.cpp"
#include "PCCInternalColorConverter.h"
#include "PCCCommon.h"
#include "PCCHighLevelSyntax.h"
#include "PCCBitstream.h"
#include "PCCVideoBitstream.h"
#include "PCCContext.h"
#include "PCCFrameContext.h"
#include "PCCPatch.h"
#include "PCCPatchSegmenter.h"
#include "PCCVideoEncoder.h"
#include "PCCSystem.h"
#include "PCCGroupOfFrames.h"
#include "PCCPointSet.h"
#include "PCCEncoderParameters.h"
#include "PCCKdTree.h"
#include "PCCNormalsGenerator.h"
#include "PCCEncoder.h"

#include "../include/pccb200.h"

#include <chrono>
#include <cstring>
#include <iostream>
#include <sstream>
#include <unistd.h>

using namespace pcc;

namespace {

void toPointSet( const int16_t* xyz, const uint8_t* rgb, size_t n, PCCPointSet3& ps ) {
  ps.resize( n );
  if ( rgb ) ps.addColors();
  for ( size_t i = 0; i < n; ++i ) {
    ps[i] = PCCPoint3D( xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] );
    if ( rgb ) ps.setColor( i, PCCColor3B( rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2] ) );
  }
}

// The reference prints per-patch chatter on std::cout / printf; silence it while a harness call runs.
struct Quiet {
  std::streambuf*   old;
  std::stringstream sink;
  Quiet() : old( std::cout.rdbuf( sink.rdbuf() ) ) {}
  ~Quiet() { std::cout.rdbuf( old ); }
};

PCCPatchSegmenter3Parameters toSegParams( const pccb200_seg_params& p ) {
  PCCPatchSegmenter3Parameters s;
  s.gridBasedSegmentation_               = false;
  s.voxelDimensionGridBasedSegmentation_ = 2;
  s.nnNormalEstimation_                  = p.nn_normal_estimation;
  s.normalOrientation_                   = p.normal_orientation;
  s.gridBasedRefineSegmentation_         = true;
  s.maxNNCountRefineSegmentation_        = p.max_nn_count_refine;
  s.iterationCountRefineSegmentation_    = p.iteration_count_refine;
  s.voxelDimensionRefineSegmentation_    = p.voxel_dim_refine;
  s.searchRadiusRefineSegmentation_      = p.search_radius_refine;
  s.occupancyResolution_                 = p.occupancy_resolution;
  s.enablePatchSplitting_                = p.enable_patch_splitting != 0;
  s.maxPatchSize_                        = p.max_patch_size;
  s.quantizerSizeX_                      = p.quantizer_size_x;
  s.quantizerSizeY_                      = p.quantizer_size_y;
  s.minPointCountPerCCPatchSegmentation_ = p.min_point_count_per_cc;
  s.maxNNCountPatchSegmentation_         = p.max_nn_count_patch_seg;
  s.surfaceThickness_                    = p.surface_thickness;
  s.minLevel_                            = p.min_level;
  s.mapCountMinus1_                      = p.map_count_minus1;
  s.maxAllowedDist2RawPointsDetection_   = p.max_allowed_dist2_raw_detection;
  s.maxAllowedDist2RawPointsSelection_   = p.max_allowed_dist2_raw_selection;
  s.lambdaRefineSegmentation_            = p.lambda_refine;
  s.useEnhancedOccupancyMapCode_         = false;
  s.absoluteD1_                          = true;
  s.surfaceSeparation_                   = false;
  s.additionalProjectionPlaneMode_       = 0;
  s.partialAdditionalProjectionPlane_    = 0.0;
  s.maxAllowedDepth_                     = p.max_allowed_depth;
  s.geometryBitDepth2D_                  = p.geometry_bitdepth_2d;
  s.geometryBitDepth3D_                  = p.geometry_bitdepth_3d;
  s.EOMFixBitCount_                      = 2;
  s.EOMSingleLayerMode_                  = false;
  s.patchExpansion_                      = false;
  s.highGradientSeparation_              = false;
  s.minGradient_                         = 15.0;
  s.minNumHighGradientPoints_            = 256;
  s.enablePointCloudPartitioning_        = false;
  s.numTilesHor_                         = 2;
  s.tileHeightToWidthRatio_              = 1.0;
  s.numCutsAlong1stLongestAxis_          = 0;
  s.numCutsAlong2ndLongestAxis_          = 0;
  s.numCutsAlong3rdLongestAxis_          = 0;
  s.createSubPointCloud_                 = false;
  s.weightNormal_ = PCCVector3D( p.weight_normal[0], p.weight_normal[1], p.weight_normal[2] );
  return s;
}

struct PatchList {
  std::vector<PCCPatch> patches;
};

void fillPatch( const PCCPatch& s, pccb200_patch& d, int64_t depthOff, int64_t occOff ) {
  std::memset( &d, 0, sizeof( d ) );
  d.index            = int32_t( s.getIndex() );
  d.view_id          = int32_t( s.getViewId() );
  d.normal_axis      = int32_t( s.getNormalAxis() );
  d.tangent_axis     = int32_t( s.getTangentAxis() );
  d.bitangent_axis   = int32_t( s.getBitangentAxis() );
  d.projection_mode  = int32_t( s.getProjectionMode() );
  d.u1               = int32_t( s.getU1() );
  d.v1               = int32_t( s.getV1() );
  d.d1               = int32_t( s.getD1() );
  d.size_u           = int32_t( s.getSizeU() );
  d.size_v           = int32_t( s.getSizeV() );
  d.size_d           = int32_t( s.getSizeD() );
  d.size_d_pixel     = int32_t( s.getSizeDPixel() );
  d.size_u0          = int32_t( s.getSizeU0() );
  d.size_v0          = int32_t( s.getSizeV0() );
  d.size_2d_x        = int32_t( s.getPatchSize2DXInPixel() );
  d.size_2d_y        = int32_t( s.getPatchSize2DYInPixel() );
  d.u0               = int32_t( s.getU0() );
  d.v0               = int32_t( s.getV0() );
  d.orientation      = int32_t( s.getPatchOrientation() );
  d.d0_count         = int32_t( s.getD0Count() );
  d.eom_and_d1_count = int32_t( s.getEOMandD1Count() );
  d.best_match_idx   = int32_t( s.getBestMatchIdx() );
  d.is_global        = s.getIsGlobalPatch() ? 1 : 0;
  d.depth_offset     = depthOff;
  d.occ_offset       = occOff;
}

}  // namespace

extern "C" {

// ---- PCCKdTree (PccLibCommon/source/PCCKdTree.cpp:42-79) ---------------------------------------------
void ref_knn( const int16_t* xyz, size_t n, const int16_t* q, size_t nq, int k, uint32_t* idx, float* dist2 ) {
  PCCPointSet3 cloud;
  toPointSet( xyz, nullptr, n, cloud );
  PCCKdTree   tree( cloud );
  PCCNNResult res;
  for ( size_t i = 0; i < nq; ++i ) {
    PCCPoint3D p( q[3 * i], q[3 * i + 1], q[3 * i + 2] );
    tree.search( p, k, res );
    for ( int j = 0; j < k; ++j ) {
      // when the tree has fewer than k points the tail of the reference's arrays is garbage: mark it
      const bool valid = size_t( j ) < std::min<size_t>( n, k );
      idx[i * k + j]   = valid ? uint32_t( res.indices( j ) ) : 0xFFFFFFFFu;
      dist2[i * k + j] = valid ? float( res.dist( j ) ) : -1.0f;
    }
  }
}

size_t ref_radius( const int16_t* xyz, size_t n, const int16_t* q, size_t nq, double radius2, size_t maxResults,
                   uint64_t* offsets, uint32_t* idx, float* dist2 ) {
  PCCPointSet3 cloud;
  toPointSet( xyz, nullptr, n, cloud );
  PCCKdTree tree( cloud );
  size_t    total = 0;
  for ( size_t i = 0; i < nq; ++i ) {
    PCCNNResult res;
    PCCPoint3D  p( q[3 * i], q[3 * i + 1], q[3 * i + 2] );
    tree.searchRadius( p, maxResults, radius2, res );
    offsets[i] = total;
    if ( idx )
      for ( size_t j = 0; j < res.count(); ++j ) {
        idx[total + j] = uint32_t( res.indices( j ) );
        if ( dist2 ) dist2[total + j] = float( res.dist( j ) );
      }
    total += res.count();
  }
  offsets[nq] = total;
  return total;
}

// ---- PCCNormalsGenerator3 (PccLibEncoder/source/PCCNormalsGenerator.cpp:61-242) -----------------------
void ref_normals( const int16_t* xyz, size_t n, int k, int orientation, double* normals ) {
  Quiet        quiet;
  PCCPointSet3 cloud;
  toPointSet( xyz, nullptr, n, cloud );
  PCCKdTree                            tree( cloud );
  PCCNormalsGenerator3                 gen;
  const double                         mx = ( std::numeric_limits<double>::max )();
  const PCCNormalsGenerator3Parameters gp = {PCCVector3D( 0.0 ), mx,       mx,       mx,
                                             mx,                 size_t( k ), size_t( k ), size_t( k ),
                                             0,                  static_cast<PCCNormalsGeneratorOrientation>( orientation ),
                                             false,              false,    false};
  gen.compute( cloud, tree, gp, 1 );
  for ( size_t i = 0; i < n; ++i )
    for ( int d = 0; d < 3; ++d ) normals[3 * i + d] = gen.getNormal( i )[d];
}

// ---- PCCEncoder::calculateWeightNormal (PccLibEncoder/source/PCCEncoder.cpp:3569-3626) ----------------
// PCCPointSet3::read (PCCPointSet.cpp:464-757): positions and colours of a .ply file as the reference loads them
int ref_read_ply( const char* path, int16_t* xyz, uint8_t* rgb, size_t capacity, size_t* n, int* hasColours ) {
  Quiet        quiet;
  PCCPointSet3 ps;
  if ( !ps.read( path ) ) return -1;
  *n          = ps.getPointCount();
  *hasColours = ps.hasColors() ? 1 : 0;
  if ( !xyz ) return 0;
  if ( capacity < *n ) return -5;
  for ( size_t i = 0; i < *n; ++i ) {
    for ( int d = 0; d < 3; ++d ) xyz[3 * i + d] = ps[i][d];
    if ( rgb && ps.hasColors() )
      for ( int d = 0; d < 3; ++d ) rgb[3 * i + d] = ps.getColor( i )[d];
  }
  return 0;
}

// PCCCodec::smoothPointCloudPostprocess, grid smoothing (PCCCodec.cpp:54-150), in place on positions / boundary point types
void ref_smooth_geometry( int16_t* xyz, uint16_t* boundary, const uint32_t* partition, size_t n, int gridSize, double threshold ) {
  Quiet        quiet;
  PCCPointSet3 rec;
  rec.resize( n );
  for ( size_t i = 0; i < n; ++i ) {
    rec[i] = PCCPoint3D( xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] );
    rec.setBoundaryPointType( i, boundary[i] );
  }
  std::vector<uint32_t>        part( partition, partition + n );
  GeneratePointCloudParameters gp;
  gp.flagGeometrySmoothing_ = true, gp.gridSmoothing_ = true, gp.gridSize_ = size_t( gridSize ), gp.thresholdSmoothing_ = threshold;
  gp.pbfEnableFlag_         = false;
  PCCEncoder enc;
  enc.smoothPointCloudPostprocess( rec, COLOR_TRANSFORM_NONE, gp, part );
  for ( size_t i = 0; i < n; ++i ) {
    for ( int d = 0; d < 3; ++d ) xyz[3 * i + d] = rec[i][d];
    boundary[i] = rec.getBoundaryPointType( i );
  }
}

// PCCPointSet3::transferColors16bitBP with the arguments of PCCEncoder::encode (PCCEncoder.cpp:656-672); target colours in place
void ref_transfer_colors16_smoothed( const int16_t* srcXyz, const uint16_t* srcCol, size_t S, const int16_t* tgtXyz, uint16_t* tgtCol,
                                     const uint16_t* tgtBoundary, size_t T ) {
  Quiet        quiet;
  PCCPointSet3 source, target;
  source.resize( S ), source.addColors(), source.addColors16bit();
  target.resize( T ), target.addColors(), target.addColors16bit();
  for ( size_t i = 0; i < S; ++i ) {
    source[i] = PCCPoint3D( srcXyz[3 * i], srcXyz[3 * i + 1], srcXyz[3 * i + 2] );
    source.setColor16bit( i, PCCColor16bit( srcCol[3 * i], srcCol[3 * i + 1], srcCol[3 * i + 2] ) );
  }
  for ( size_t i = 0; i < T; ++i ) {
    target[i] = PCCPoint3D( tgtXyz[3 * i], tgtXyz[3 * i + 1], tgtXyz[3 * i + 2] );
    target.setColor16bit( i, PCCColor16bit( tgtCol[3 * i], tgtCol[3 * i + 1], tgtCol[3 * i + 2] ) );
    target.setBoundaryPointType( i, tgtBoundary[i] );
  }
  source.transferColors16bitBP( target, 1, int32_t( 0 ), false, 8, 1, true, true, true, false, 4, 4, 1000, 1000, 1000 * 256, 1000 * 256 );
  for ( size_t i = 0; i < T; ++i )
    for ( int d = 0; d < 3; ++d ) tgtCol[3 * i + d] = target.getColor16bit( i )[d];
}

// PCCInternalColorConverter "YUV420ToYUV444_8_0" on one frame: the inverse conversion after the video codec
void ref_yuv420_to_yuv444_16( const uint8_t* yuv420, size_t W, size_t H, uint16_t* yuv444 ) {
  Quiet             quiet;
  PCCVideoAttribute video;
  video.resize( 1 );
  auto& img = video.getFrame( 0 );
  img.resize( W, H, PCCCOLORFORMAT::YUV420 );
  const size_t Q = W * H, q4 = ( W / 2 ) * ( H / 2 );
  img.getChannel( 0 ).assign( yuv420, yuv420 + Q );
  img.getChannel( 1 ).assign( yuv420 + Q, yuv420 + Q + q4 );
  img.getChannel( 2 ).assign( yuv420 + Q + q4, yuv420 + Q + 2 * q4 );
  PCCInternalColorConverter<uint16_t> converter;
  converter.convert( "YUV420ToYUV444_8_0", video );
  auto& out = video.getFrame( 0 );
  for ( int c = 0; c < 3; ++c ) std::copy( out.getChannel( c ).begin(), out.getChannel( c ).end(), yuv444 + c * Q );
}
// PCCPointSet3::convertYUV16ToRGB8
void ref_yuv16_to_rgb8( const uint16_t* yuv, size_t n, uint8_t* rgb ) {
  PCCPointSet3 ps;
  ps.resize( n ), ps.addColors(), ps.addColors16bit();
  for ( size_t i = 0; i < n; ++i ) ps.setColor16bit( i, PCCColor16bit( yuv[3 * i], yuv[3 * i + 1], yuv[3 * i + 2] ) );
  ps.convertYUV16ToRGB8();
  for ( size_t i = 0; i < n; ++i )
    for ( int d = 0; d < 3; ++d ) rgb[3 * i + d] = ps.getColor( i )[d];
}

void ref_weight_normal( const int16_t* xyz, size_t n, int bits, double minW, double w[3] ) {
  Quiet        quiet;
  PCCPointSet3 cloud;
  toPointSet( xyz, nullptr, n, cloud );
  PCCEncoder           enc;
  PCCEncoderParameters ep;
  ep.enhancedPP_   = true;
  ep.minWeightEPP_ = minW;
  enc.setParameters( ep );
  PCCVector3D r = enc.calculateWeightNormal( size_t( bits ), cloud );
  w[0] = r[0], w[1] = r[1], w[2] = r[2];
}

// ---- PCCPatchSegmenter3 stage by stage (PccLibEncoder/source/PCCPatchSegmenter.cpp:53-150) ------------
// Runs the reference segmenter exactly as compute() does and exposes the intermediate results.
// normals (n x 3), partition0 (after initialSegmentation), partition1 (after refineSegmentationGridBased)
// may be NULL. Returns an opaque patch list (ref_patches_*).
void* ref_segment_frame( const int16_t*            xyz,
                         const uint8_t*            rgb,
                         size_t                    n,
                         const pccb200_seg_params* p,
                         double*                   normals,
                         uint8_t*                  partition0,
                         uint8_t*                  partition1,
                         double*                   seconds ) {
  Quiet        quiet;
  PCCPointSet3 cloud;
  toPointSet( xyz, rgb, n, cloud );
  PCCPatchSegmenter3Parameters sp = toSegParams( *p );
  PCCPatchSegmenter3           seg;
  seg.setNbThread( 1 );
  auto       t0 = std::chrono::steady_clock::now();
  PatchList* pl = new PatchList();
  if ( n == 0 ) return pl;
  PCCKdTree                            tree( cloud );
  PCCNormalsGenerator3                 gen;
  const double                         mx = ( std::numeric_limits<double>::max )();
  const PCCNormalsGenerator3Parameters gp = {PCCVector3D( 0.0 ),
                                             mx,
                                             mx,
                                             mx,
                                             mx,
                                             sp.nnNormalEstimation_,
                                             sp.nnNormalEstimation_,
                                             sp.nnNormalEstimation_,
                                             0,
                                             static_cast<PCCNormalsGeneratorOrientation>( sp.normalOrientation_ ),
                                             false,
                                             false,
                                             false};
  gen.compute( cloud, tree, gp, 1 );
  if ( normals )
    for ( size_t i = 0; i < n; ++i )
      for ( int d = 0; d < 3; ++d ) normals[3 * i + d] = gen.getNormal( i )[d];
  std::vector<size_t> partition;
  seg.initialSegmentation( cloud, gen, seg.orientations6, seg.orientationCount6, partition, sp.weightNormal_ );
  if ( partition0 )
    for ( size_t i = 0; i < n; ++i ) partition0[i] = uint8_t( partition[i] );
  seg.refineSegmentationGridBased( cloud, gen, seg.orientations6, seg.orientationCount6,
                                   sp.maxNNCountRefineSegmentation_, sp.lambdaRefineSegmentation_,
                                   sp.iterationCountRefineSegmentation_, sp.voxelDimensionRefineSegmentation_,
                                   sp.searchRadiusRefineSegmentation_, partition );
  if ( partition1 )
    for ( size_t i = 0; i < n; ++i ) partition1[i] = uint8_t( partition[i] );
  PCCPointSet3              resampled;
  std::vector<size_t>       patchPartition, resampledPatchPartition, rawPoints;
  std::vector<PCCPointSet3> sub;
  float                     dist = 0;
  seg.segmentPatches( cloud, 0, tree, sp, partition, pl->patches, patchPartition, resampledPatchPartition, rawPoints,
                      resampled, sub, dist, gen, seg.orientations6, seg.orientationCount6 );
  if ( seconds ) *seconds = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
  return pl;
}

int ref_patches_count( void* h ) { return int( static_cast<PatchList*>( h )->patches.size() ); }
size_t ref_patches_depth_elems( void* h ) {
  size_t s = 0;
  for ( auto& p : static_cast<PatchList*>( h )->patches ) s += 2 * p.getSizeU() * p.getSizeV();
  return s;
}
size_t ref_patches_occ_elems( void* h ) {
  size_t s = 0;
  for ( auto& p : static_cast<PatchList*>( h )->patches ) s += p.getSizeU0() * p.getSizeV0();
  return s;
}
void ref_patches_get( void* h, pccb200_patch* out, int16_t* depth, uint8_t* occ ) {
  int64_t dOff = 0, oOff = 0;
  size_t  i = 0;
  for ( auto& p : static_cast<PatchList*>( h )->patches ) {
    fillPatch( p, out[i++], dOff, oOff );
    const size_t px = p.getSizeU() * p.getSizeV();
    for ( int m = 0; m < 2; ++m ) {
      const auto& dm = p.getDepth( m );
      for ( size_t j = 0; j < px; ++j ) depth[dOff + m * px + j] = j < dm.size() ? dm[j] : int16_t( 0 );
    }
    dOff += 2 * px;
    const auto&  o  = p.getOccupancy();
    const size_t nb = p.getSizeU0() * p.getSizeV0();
    for ( size_t j = 0; j < nb; ++j ) occ[oOff + j] = j < o.size() ? uint8_t( o[j] ) : 0;
    oOff += nb;
  }
}
void ref_patches_free( void* h ) { delete static_cast<PatchList*>( h ); }

}  // extern "C"

// =====================================================================================================
// GOF-level harness: the reference's own encoder stages, in the order PCCEncoder::encode runs them
// (PccLibEncoder/source/PCCEncoder.cpp:71-424), with the three videoEncoder.compress calls replaced by
// identity (a lossless codec: decoded == source). Exposes every intermediate product the hot path hands over.
// =====================================================================================================
namespace {

struct RefFrame {
  PatchList            patches;  // packed order, with u0/v0/orientation
  size_t               width = 0, height = 0;
  std::vector<uint8_t> occupancy, omVideo;
  std::vector<uint32_t> blockToPatch;
  std::vector<uint16_t> geo[2];
  std::vector<int16_t>  recXyz;
  std::vector<uint32_t> pointToPixel, recPartition;
  std::vector<uint16_t> recBoundary;
  std::vector<uint8_t>  recRgb;
  std::vector<uint16_t> attrRaw[2], attr[2];
  std::vector<uint8_t>  attrYuv[2];
};
struct RefGof {
  std::vector<RefFrame> frames;
  double                seconds[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // segment, pack, occupancy, geometry, reconstruct, attribute, padding, total
};

void setCtcParams( PCCEncoderParameters& ep, const pccb200_seg_params& p, int occupancyPrecision ) {
  // cfg/common/ctc-common.cfg + cfg/condition/ctc-all-intra.cfg + per-sequence values carried in pccb200_seg_params
  ep.nnNormalEstimation_                  = p.nn_normal_estimation;
  ep.normalOrientation_                   = p.normal_orientation;
  ep.gridBasedRefineSegmentation_         = true;
  ep.maxNNCountRefineSegmentation_        = p.max_nn_count_refine;
  ep.iterationCountRefineSegmentation_    = p.iteration_count_refine;
  ep.voxelDimensionRefineSegmentation_    = p.voxel_dim_refine;
  ep.searchRadiusRefineSegmentation_      = p.search_radius_refine;
  ep.occupancyResolution_                 = p.occupancy_resolution;
  ep.enablePatchSplitting_                = p.enable_patch_splitting != 0;
  ep.maxPatchSize_                        = p.max_patch_size;
  ep.minPointCountPerCCPatchSegmentation_ = p.min_point_count_per_cc;
  ep.maxNNCountPatchSegmentation_         = p.max_nn_count_patch_seg;
  ep.surfaceThickness_                    = p.surface_thickness;
  ep.minLevel_                            = p.min_level;
  ep.maxAllowedDist2RawPointsDetection_   = p.max_allowed_dist2_raw_detection;
  ep.maxAllowedDist2RawPointsSelection_   = p.max_allowed_dist2_raw_selection;
  ep.lambdaRefineSegmentation_            = p.lambda_refine;
  ep.mapCountMinus1_                      = p.map_count_minus1;
  ep.geometry3dCoordinatesBitdepth_       = p.geometry_bitdepth_3d - 1;
  ep.geometryNominal2dBitdepth_           = p.geometry_bitdepth_2d;
  ep.minimumImageWidth_                   = p.geometry_bitdepth_3d > 11 ? 2560 : 1280;
  ep.minimumImageHeight_                  = 1280;
  ep.occupancyPrecision_                  = occupancyPrecision;
  ep.bestColorSearchRange_                = 0;
  ep.numNeighborsColorTransferFwd_        = 8;
  ep.numNeighborsColorTransferBwd_        = 1;
  ep.useDistWeightedAverageFwd_ = ep.useDistWeightedAverageBwd_ = true;
  ep.skipAvgIfIdenticalSourcePointPresentFwd_ = ep.skipAvgIfIdenticalSourcePointPresentBwd_ = true;
  ep.distOffsetFwd_ = ep.distOffsetBwd_ = 4;
  ep.maxGeometryDist2Fwd_ = ep.maxGeometryDist2Bwd_ = 1000;
  ep.maxColorDist2Fwd_ = ep.maxColorDist2Bwd_ = 1000;
  ep.flagGeometrySmoothing_               = true;
  ep.gridSmoothing_                       = true;
  ep.flagColorPreSmoothing_               = true;
  ep.enablePointCloudPartitioning_        = false;
  ep.enhancedOccupancyMapCode_            = false;
  ep.profileReconstructionIdc_            = 1;
  ep.constrainedPack_                     = p.global_patch_allocation != 0;  // (cfg/condition/ctc-random-access.cfg keeps the default 1)
  ep.globalPatchAllocation_               = p.global_patch_allocation;
  ep.nbThread_                            = 1;
  ep.groupOfFramesSize_                   = 32;
  ep.compressedStreamPath_                = "/tmp/pccb200_ref_harness.bin";
}

}  // namespace

// Optional replacement of the hot-path stages by an external implementation (oracle/shim_harness.cpp registers the reference-side
// shim of libpccb200 here): everything else of ref_encode_gof - set-up, the order of the steps, the extraction of the products from
// the reference's own data structures - is shared, so the two runs differ in nothing but who filled those structures.
struct RefHotPathHooks {
  int ( *stageA )( PCCEncoder& enc, PCCGroupOfFrames& sources, PCCContext& context );
  int ( *stageB1 )( PCCEncoder& enc, PCCContext& context, PCCGroupOfFrames& reconstructs, std::vector<std::vector<uint32_t>>& partitions );
  int ( *stageB2 )( PCCEncoder& enc, PCCContext& context );
};
static const RefHotPathHooks* gHooks = nullptr;
extern "C" void ref_set_hot_path_hooks( const RefHotPathHooks* hooks ) { gHooks = hooks; }

namespace {

template <typename T>
void planes( const PCCImage<T, 3>& img, std::vector<uint16_t>& out ) {
  const size_t n = img.getWidth() * img.getHeight();
  out.resize( 3 * n );
  for ( int c = 0; c < 3; ++c )
    for ( size_t i = 0; i < n; ++i ) out[c * n + i] = uint16_t( img.getChannel( c )[i] );
}

}  // namespace

extern "C" {

// frames: nframes clouds (xyz[f]: n[f] x 3 int16, rgb[f]: n[f] x 3 uint8). stopAfter: 0 = everything,
// 1 = after packing, 2 = after geometry images, 3 = after generatePointCloud (saves time in focused tests).
void* ref_encode_gof( int nframes, const int16_t* const* xyz, const uint8_t* const* rgb, const size_t* n,
                      const pccb200_seg_params* p, int occupancyPrecision, int stopAfter ) {
  Quiet  quiet;
  FILE*  devnull  = fopen( "/dev/null", "w" );
  int    savedOut = dup( 1 );
  fflush( stdout );
  dup2( fileno( devnull ), 1 );  // the reference also chats through printf
  RefGof* G = new RefGof();
  G->frames.resize( nframes );
  using clk = std::chrono::steady_clock;
  auto secs = []( clk::time_point a ) { return std::chrono::duration<double>( clk::now() - a ).count(); };
  {
    PCCEncoderParameters ep;
    setCtcParams( ep, *p, occupancyPrecision );
    ep.check();
    PCCGroupOfFrames sources, reconstructs;
    sources.setFrameCount( nframes );
    for ( int f = 0; f < nframes; ++f ) toPointSet( xyz[f], rgb[f], n[f], sources[f] );
    PCCLogger  logger;
    PCCEncoder enc;
    enc.setLogger( logger );
    enc.setParameters( ep );
    PCCContext context;
    context.addV3CParameterSet( 0 );
    context.setActiveVpsId( 0 );
    auto tAll = clk::now();
    // ---- PCCEncoder::encode, :80-100
    reconstructs.setFrameCount( sources.getFrameCount() );
    context.resizeAtlas( 1 );
    context.setAtlasIndex( 0 );
    context.resize( sources.getFrameCount() );
    auto& frames = context.getFrames();
    for ( size_t i = 0; i < frames.size(); i++ ) {
      auto& fc = frames[i].getTitleFrameContext();
      fc.setFrameIndex( i );
      fc.setRawPatchEnabledFlag( false );
      fc.setUseRawPointsSeparateVideo( false );
      fc.setGeometry3dCoordinatesBitdepth( enc.params_.geometry3dCoordinatesBitdepth_ + 1 );
      fc.setGeometry2dBitdepth( enc.params_.geometryNominal2dBitdepth_ );
      fc.setMaxDepth( ( 1 << enc.params_.geometryNominal2dBitdepth_ ) - 1 );
      fc.setLog2PatchQuantizerSizeX( enc.params_.log2QuantizerSizeX_ );
      fc.setLog2PatchQuantizerSizeY( enc.params_.log2QuantizerSizeY_ );
    }
    const RefHotPathHooks* hooks = gHooks;
    const bool             hookA = hooks && hooks->stageA, hookB1 = hooks && hooks->stageB1, hookB2 = hooks && hooks->stageB2;
    int                    hookError = 0;
    auto t0 = clk::now();
    if ( hookA ) {
      hookError = hooks->stageA( enc, sources, context );  // :103-172 through the external implementation
    } else {
      enc.generateSegments( sources, context );  // :103
      G->seconds[0] = secs( t0 );
      enc.params_.initializeContext( context );   // :107
      t0 = clk::now();
      enc.placeSegments( sources, context );  // :110
      G->seconds[1] = secs( t0 );
    }
    if ( hookError ) {
      G->frames.clear();
      nframes = 0;
      stopAfter = 1;
    }
    for ( int f = 0; f < nframes; ++f ) {
      auto& tile             = context[f].getTile( 0 );
      G->frames[f].patches.patches = tile.getPatches();
      G->frames[f].width     = context[f].getAtlasFrameWidth();
      G->frames[f].height    = context[f].getAtlasFrameHeight();
    }
    if ( stopAfter != 1 ) {
      t0 = clk::now();
      if ( !hookA ) {
        enc.generateOccupancyMap( context, true );         // :133
        enc.generateOccupancyMapVideo( sources, context );  // :139  (compress skipped: lossless)
        enc.generateBlockToPatchFromOccupancyMapVideo( context, enc.params_.occupancyResolution_, enc.params_.occupancyPrecision_ );  // :168
      }
      G->seconds[2] = secs( t0 );
      for ( int f = 0; f < nframes; ++f ) {
        auto& R   = G->frames[f];
        auto& occ = context[f].getTitleFrameContext().getOccupancyMap();
        R.occupancy.assign( occ.begin(), occ.end() );
        auto& om = context.getVideoOccupancyMap().getFrame( f );
        R.omVideo.assign( om.getChannel( 0 ).begin(), om.getChannel( 0 ).end() );
        auto& b2p = context[f].getTile( 0 ).getBlockToPatch();
        R.blockToPatch.assign( b2p.begin(), b2p.end() );
      }
      t0 = clk::now();
      if ( !hookA ) enc.generateGeometryVideo( sources, context );  // :172
      G->seconds[3] = secs( t0 );
      for ( int f = 0; f < nframes; ++f )
        for ( int m = 0; m < 2; ++m ) {
          auto& img = context.getVideoGeometryMultiple()[0].getFrame( 2 * f + m );
          G->frames[f].geo[m].assign( img.getChannel( 0 ).begin(), img.getChannel( 0 ).end() );
        }
    }
    if ( stopAfter == 0 || stopAfter == 3 ) {
      t0 = clk::now();
      GeneratePointCloudParameters gpc;
      enc.setGeneratePointCloudParameters( gpc, context );  // :315
      context.allocOneLayerData();
      std::vector<std::vector<uint32_t>> partitions( context.size() );
      if ( hookB1 ) {
        hooks->stageB1( enc, context, reconstructs, partitions );  // :319-341 through the external implementation
      } else {
        for ( size_t f = 0; f < context.size(); f++ ) {
          PCCPointSet3 rec;
          enc.generatePointCloud( rec, context, f, 0, gpc, partitions[f], false );  // :328
          reconstructs[f].appendPointSet( rec );
        }
      }
      G->seconds[4] = secs( t0 );
      for ( int f = 0; f < nframes; ++f ) {
        auto&        R   = G->frames[f];
        auto&        rec = reconstructs[f];
        const size_t m   = rec.getPointCount();
        R.recXyz.resize( 3 * m ), R.recBoundary.resize( m );
        for ( size_t i = 0; i < m; ++i ) {
          for ( int d = 0; d < 3; ++d ) R.recXyz[3 * i + d] = rec[i][d];
          R.recBoundary[i] = rec.getBoundaryPointType( i );
        }
        auto& p2p = context[f].getTitleFrameContext().getPointToPixel();
        R.pointToPixel.resize( 3 * p2p.size() );
        for ( size_t i = 0; i < p2p.size(); ++i )
          for ( int d = 0; d < 3; ++d ) R.pointToPixel[3 * i + d] = uint32_t( p2p[i][d] );
        R.recPartition = partitions[f];
      }
    }
    if ( stopAfter == 0 ) {
      t0 = clk::now();
      if ( !hookB1 ) enc.generateAttributeVideo( sources, reconstructs, context, enc.params_ );  // :341
      G->seconds[5] = secs( t0 );
      auto& video = context.getVideoAttributesMultiple()[0];
      for ( int f = 0; f < nframes; ++f ) {
        auto& R = G->frames[f];
        for ( int m = 0; m < 2; ++m ) planes( video.getFrame( 2 * f + m ), R.attrRaw[m] );
        auto&        rec = reconstructs[f];
        const size_t mP  = rec.getPointCount();
        R.recRgb.resize( 3 * mP );
        for ( size_t i = 0; i < mP; ++i )
          for ( int d = 0; d < 3; ++d ) R.recRgb[3 * i + d] = rec.getColor( i )[d];
      }
      t0 = clk::now();
      if ( hookB2 ) {
        hooks->stageB2( enc, context );  // :344-424 through the external implementation
      } else {
        for ( int f = 0; f < nframes; ++f )
          for ( int m = 0; m < 2; ++m ) enc.dilateSmoothedPushPull( frames[f].getTitleFrameContext(), video.getFrame( 2 * f + m ) );  // :367
        // Group dilation of the two attribute maps (PCCEncoder.cpp:391-413). The reference has this loop inline in encode(), not
        // as a member that could be called, so it is applied here to the reference's own images: where the (block-precision)
        // occupancy map is empty both maps take the rounded mean of their 8-bit values.
        if ( enc.params_.mapCountMinus1_ > 0 && !enc.params_.multipleStreams_ && enc.params_.groupDilation_ ) {
          for ( int f = 0; f < nframes; ++f ) {
            auto&        tile = frames[f].getTitleFrameContext();
            auto&        occ  = tile.getOccupancyMap();
            const size_t w = tile.getWidth(), h = tile.getHeight();
            auto &       t0img = video.getFrame( 2 * f ), &t1img = video.getFrame( 2 * f + 1 );
            for ( size_t y = 0; y < h; ++y )
              for ( size_t x = 0; x < w; ++x ) {
                if ( occ[y * w + x] != 0 ) continue;
                for ( size_t c = 0; c < 3; ++c ) {
                  const uint32_t mean = ( uint32_t( uint8_t( t0img.getValue( c, x, y ) ) ) + uint32_t( uint8_t( t1img.getValue( c, x, y ) ) ) + 1 ) >> 1;
                  t0img.setValue( c, x, y, uint8_t( mean ) ), t1img.setValue( c, x, y, uint8_t( mean ) );
                }
              }
          }
        }
      }
      G->seconds[6] = secs( t0 );
      for ( int f = 0; f < nframes; ++f )
        for ( int m = 0; m < 2; ++m ) planes( video.getFrame( 2 * f + m ), G->frames[f].attr[m] );
      // what PCCVideoEncoder::compress does to this video before it reaches the codec (PCCVideoEncoder.cpp:326-353, internal
      // colour converter, 8 bit, default down-sampling filter 4)
      PCCVideoAttribute                   yuv = video;
      PCCInternalColorConverter<uint16_t> converter;
      converter.convert( "RGB444ToYUV420_8_4", yuv );
      for ( int f = 0; f < nframes; ++f )
        for ( int m = 0; m < 2; ++m ) {
          auto& img = yuv.getFrame( 2 * f + m );
          auto& out = G->frames[f].attrYuv[m];
          out.clear();
          for ( int c = 0; c < 3; ++c )
            for ( auto v : img.getChannel( c ) ) out.push_back( uint8_t( v ) );
        }
    }
    G->seconds[7] = secs( tAll );
  }
  fflush( stdout );
  dup2( savedOut, 1 );
  close( savedOut );
  fclose( devnull );
  return G;
}

// PCCEncoder::placeSegments (:4762-4835) alone, on patch lists given by the caller (metadata + block occupancy; no clouds, no
// depth maps): packing tests that do not need a segmentation. `ra` != 0 = constrainedPack + globalPatchAllocation 1.
void* ref_pack_gof( int nframes, const int* counts, const pccb200_patch* patches, const uint8_t* occ, const int64_t* occBase, int ra, int bits ) {
  Quiet  quiet;
  FILE*  devnull  = fopen( "/dev/null", "w" );
  int    savedOut = dup( 1 );
  fflush( stdout );
  dup2( fileno( devnull ), 1 );
  RefGof* G = new RefGof();
  G->frames.resize( nframes );
  {
    pccb200_seg_params sp{};
    sp.nn_normal_estimation = 16, sp.normal_orientation = 1, sp.max_nn_count_refine = 1024, sp.iteration_count_refine = 1, sp.voxel_dim_refine = 4;
    sp.search_radius_refine = 192, sp.occupancy_resolution = 16, sp.enable_patch_splitting = 1, sp.max_patch_size = 1024, sp.quantizer_size_x = 16;
    sp.quantizer_size_y = 16, sp.min_point_count_per_cc = 16, sp.max_nn_count_patch_seg = 16, sp.surface_thickness = 4, sp.min_level = 64;
    sp.max_allowed_depth = 255, sp.geometry_bitdepth_2d = 8, sp.geometry_bitdepth_3d = bits + 1, sp.map_count_minus1 = 1;
    sp.global_patch_allocation = ra ? 1 : 0, sp.lambda_refine = 3.0, sp.max_allowed_dist2_raw_detection = 9.0, sp.max_allowed_dist2_raw_selection = 1.0;
    PCCEncoderParameters ep;
    setCtcParams( ep, sp, 4 );
    ep.check();
    PCCGroupOfFrames sources;
    sources.setFrameCount( nframes );
    PCCLogger  logger;
    PCCEncoder enc;
    enc.setLogger( logger );
    enc.setParameters( ep );
    PCCContext context;
    context.addV3CParameterSet( 0 );
    context.setActiveVpsId( 0 );
    context.resizeAtlas( 1 );
    context.setAtlasIndex( 0 );
    context.resize( nframes );
    auto&  frames = context.getFrames();
    size_t at     = 0;
    for ( int f = 0; f < nframes; ++f ) {
      auto& fc = frames[f].getTitleFrameContext();
      fc.setFrameIndex( f );
      fc.setRawPatchEnabledFlag( false );
      fc.setUseRawPointsSeparateVideo( false );
      sources[f].resize( counts[f] ? 1 : 0 );  // (placeSegments only looks at "empty or not")
      auto& list = fc.getPatches();
      list.resize( counts[f] );
      for ( int i = 0; i < counts[f]; ++i, ++at ) {
        const pccb200_patch& r = patches[at];
        PCCPatch&            p = list[i];
        p.setIndex( r.index );
        p.setViewId( r.view_id );
        p.setU1( r.u1 ), p.setV1( r.v1 ), p.setD1( r.d1 ), p.setSizeU( r.size_u ), p.setSizeV( r.size_v );
        p.setSizeU0( r.size_u0 ), p.setSizeV0( r.size_v0 ), p.setOccupancyResolution( 16 );
        p.getPreGPAPatchData().initialize(), p.getCurGPAPatchData().initialize();
        std::vector<bool> o( size_t( r.size_u0 ) * r.size_v0 );
        for ( size_t b = 0; b < o.size(); ++b ) o[b] = occ[occBase[f] + r.occ_offset + b] != 0;
        p.setOccupancy( o );
      }
    }
    enc.params_.initializeContext( context );
    enc.placeSegments( sources, context );
    for ( int f = 0; f < nframes; ++f ) {
      G->frames[f].patches.patches = context[f].getTile( 0 ).getPatches();
      G->frames[f].width           = context[f].getAtlasFrameWidth();
      G->frames[f].height          = context[f].getAtlasFrameHeight();
    }
  }
  fflush( stdout );
  dup2( savedOut, 1 );
  close( savedOut );
  fclose( devnull );
  return G;
}

void   ref_gof_free( void* h ) { delete static_cast<RefGof*>( h ); }
void   ref_gof_seconds( void* h, double* out ) { std::memcpy( out, static_cast<RefGof*>( h )->seconds, 8 * sizeof( double ) ); }
void   ref_gof_dims( void* h, int f, size_t* w, size_t* hgt, size_t* recPoints ) {
  auto& R = static_cast<RefGof*>( h )->frames[f];
  *w = R.width, *hgt = R.height, *recPoints = R.recXyz.size() / 3;
}
void* ref_gof_patches( void* h, int f ) {  // borrowed view usable with ref_patches_* (do NOT free)
  return &static_cast<RefGof*>( h )->frames[f].patches;
}
// what: see tests/bindings.py GOF_* ; returns the element count, copies when dst != NULL
size_t ref_gof_get( void* h, int f, int what, void* dst ) {
  auto&  R = static_cast<RefGof*>( h )->frames[f];
  auto   put = [&]( const void* src, size_t bytes ) {
    if ( dst && bytes ) std::memcpy( dst, src, bytes );
  };
  switch ( what ) {
    case 1: put( R.occupancy.data(), R.occupancy.size() ); return R.occupancy.size();
    case 2: put( R.omVideo.data(), R.omVideo.size() ); return R.omVideo.size();
    case 3: put( R.blockToPatch.data(), R.blockToPatch.size() * 4 ); return R.blockToPatch.size();
    case 4: put( R.geo[0].data(), R.geo[0].size() * 2 ); return R.geo[0].size();
    case 5: put( R.geo[1].data(), R.geo[1].size() * 2 ); return R.geo[1].size();
    case 6: put( R.recXyz.data(), R.recXyz.size() * 2 ); return R.recXyz.size();
    case 7: put( R.pointToPixel.data(), R.pointToPixel.size() * 4 ); return R.pointToPixel.size();
    case 8: put( R.recPartition.data(), R.recPartition.size() * 4 ); return R.recPartition.size();
    case 9: put( R.recBoundary.data(), R.recBoundary.size() * 2 ); return R.recBoundary.size();
    case 10: put( R.recRgb.data(), R.recRgb.size() ); return R.recRgb.size();
    case 11: put( R.attrRaw[0].data(), R.attrRaw[0].size() * 2 ); return R.attrRaw[0].size();
    case 12: put( R.attrRaw[1].data(), R.attrRaw[1].size() * 2 ); return R.attrRaw[1].size();
    case 13: put( R.attr[0].data(), R.attr[0].size() * 2 ); return R.attr[0].size();
    case 14: put( R.attr[1].data(), R.attr[1].size() * 2 ); return R.attr[1].size();
    case 15: put( R.attrYuv[0].data(), R.attrYuv[0].size() ); return R.attrYuv[0].size();
    case 16: put( R.attrYuv[1].data(), R.attrYuv[1].size() ); return R.attrYuv[1].size();
    default: return 0;
  }
}

}  // extern "C"
