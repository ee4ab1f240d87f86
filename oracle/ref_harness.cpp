// ref_harness.cpp — OUR thin C-ABI wrapper around the UNMODIFIED reference classes (TEST INFRASTRUCTURE ONLY).
//
// Compiled by oracle/Makefile together with the reference's own sources (taken where they lie under
// /root/reference) into oracle/_ref/libtmc2ref.so.  It lets tests pin the CPU restatement (pcc_oracle.cpp)
// and the CUDA path against the reference itself, stage by stage, and lets bench.py time the reference's own
// CPU implementation of the hot path (`--impl reference`, cpu_baseline.kind = "reference").
// Nothing here re-implements an algorithm: every function converts plain arrays to the reference's containers,
// calls the reference's own methods and converts back.
// (built with -fno-access-control so the harness can call the reference's private stage methods)
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
