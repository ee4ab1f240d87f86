// pccb200_shim.cpp — see pccb200_shim.h. Reference-side code: uses the reference's classes, calls only the C ABI.
#include "pccb200_shim.h"

#include <cstdio>

#include "PCCFrameContext.h"
#include "PCCImage.h"
#include "PCCPatch.h"
#include "PCCPointSet.h"
#include "PCCVideo.h"

using namespace pcc;

namespace pccb200shim {

Session::~Session() {
  if ( gof ) pccb200_gof_free( gof );
  if ( ctx ) pccb200_destroy( ctx );
}

pccb200_seg_params toSegParams( const PCCEncoderParameters& p ) {
  pccb200_seg_params s{};
  s.nn_normal_estimation            = int32_t( p.nnNormalEstimation_ );
  s.normal_orientation              = int32_t( p.normalOrientation_ );
  s.max_nn_count_refine             = int32_t( p.maxNNCountRefineSegmentation_ );
  s.iteration_count_refine          = int32_t( p.iterationCountRefineSegmentation_ );
  s.voxel_dim_refine                = int32_t( p.voxelDimensionRefineSegmentation_ );
  s.search_radius_refine            = int32_t( p.searchRadiusRefineSegmentation_ );
  s.occupancy_resolution            = int32_t( p.occupancyResolution_ );
  s.enable_patch_splitting          = p.enablePatchSplitting_ ? 1 : 0;
  s.max_patch_size                  = int32_t( p.maxPatchSize_ );
  s.quantizer_size_x                = 1 << p.log2QuantizerSizeX_;
  s.quantizer_size_y                = 1 << p.log2QuantizerSizeY_;
  s.min_point_count_per_cc          = int32_t( p.minPointCountPerCCPatchSegmentation_ );
  s.max_nn_count_patch_seg          = int32_t( p.maxNNCountPatchSegmentation_ );
  s.surface_thickness               = int32_t( p.surfaceThickness_ );
  s.min_level                       = int32_t( p.minLevel_ );
  s.max_allowed_depth               = ( 1 << p.geometryNominal2dBitdepth_ ) - 1;
  s.geometry_bitdepth_2d            = int32_t( p.geometryNominal2dBitdepth_ );
  s.geometry_bitdepth_3d            = int32_t( p.geometry3dCoordinatesBitdepth_ ) + 1;
  s.map_count_minus1                = int32_t( p.mapCountMinus1_ );
  s.global_patch_allocation         = p.constrainedPack_ ? int32_t( p.globalPatchAllocation_ ) : 0;
  s.lambda_refine                   = p.lambdaRefineSegmentation_;
  s.max_allowed_dist2_raw_detection = p.maxAllowedDist2RawPointsDetection_;
  s.max_allowed_dist2_raw_selection = p.maxAllowedDist2RawPointsSelection_;
  s.weight_normal[0] = s.weight_normal[1] = s.weight_normal[2] = 1.0;
  return s;
}

namespace {

template <typename T, typename U>
int fetch( pccb200_gof* gof, int f, int what, std::vector<T>& dst, std::vector<U>& staging ) {
  const size_t n = pccb200_gof_get( gof, f, what, nullptr );
  staging.resize( n );
  if ( n && pccb200_gof_get( gof, f, what, staging.data() ) != n ) return PCCB200_ERR_STATE;
  dst.assign( staging.begin(), staging.end() );  // widening copy into the reference's element type
  return PCCB200_OK;
}

// PCCPatch records from the ABI's
void fillPatches( std::vector<PCCPatch>& out, const pccb200_patchlist* pl ) {
  const int                  P = pccb200_patches_count( pl );
  std::vector<pccb200_patch> rec( P );
  std::vector<int16_t>       depth( pccb200_patches_depth_elems( pl ) );
  std::vector<uint8_t>       occ( pccb200_patches_occ_elems( pl ) );
  pccb200_patches_get( pl, rec.data(), depth.data(), occ.data() );
  out.clear();
  out.resize( P );
  for ( int i = 0; i < P; ++i ) {
    const pccb200_patch& r = rec[i];
    PCCPatch&            p = out[i];
    p.setIndex( r.index );
    p.setViewId( r.view_id );  // normal / tangent / bitangent axis + projection mode
    p.setU1( r.u1 ), p.setV1( r.v1 ), p.setD1( r.d1 );
    p.setSizeU( r.size_u ), p.setSizeV( r.size_v ), p.setSizeD( r.size_d ), p.setSizeDPixel( r.size_d_pixel );
    p.setSizeU0( r.size_u0 ), p.setSizeV0( r.size_v0 );
    p.setPatchSize2DXInPixel( r.size_2d_x ), p.setPatchSize2DYInPixel( r.size_2d_y );
    p.setU0( r.u0 ), p.setV0( r.v0 ), p.setPatchOrientation( r.orientation );
    p.setOccupancyResolution( 16 );
    p.setD0Count( r.d0_count ), p.setEOMandD1Count( r.eom_and_d1_count );
    p.setBestMatchIdx( r.best_match_idx ), p.setIsGlobalPatch( r.is_global != 0 );
    p.setPatchType( uint8_t( r.best_match_idx >= 0 ? P_INTER : P_INTRA ) );
    const size_t px = size_t( r.size_u ) * r.size_v;
    p.setDepth( 0, std::vector<int16_t>( depth.begin() + r.depth_offset, depth.begin() + r.depth_offset + px ) );
    p.setDepth( 1, std::vector<int16_t>( depth.begin() + r.depth_offset + px, depth.begin() + r.depth_offset + 2 * px ) );
    std::vector<bool> o( size_t( r.size_u0 ) * r.size_v0 );
    for ( size_t b = 0; b < o.size(); ++b ) o[b] = occ[r.occ_offset + b] != 0;
    p.setOccupancy( o );
  }
}

}  // namespace

int stageA( Session& s, PCCEncoderParameters& params, const PCCGroupOfFrames& sources, PCCContext& context ) {
  if ( !s.ctx ) {
    const int rc = pccb200_create( 0, &s.ctx );
    if ( rc != PCCB200_OK ) return rc;
  }
  if ( s.gof ) pccb200_gof_free( s.gof ), s.gof = nullptr;
  const size_t       frameCount = sources.getFrameCount();
  pccb200_seg_params sp         = toSegParams( params );
  if ( params.enhancedPP_ && frameCount > 0 && sources[0].getPointCount() > 0 ) {  // calculateWeightNormal (:3569), frame 0 of the GOF (:4726)
    const PCCPointSet3& f0 = sources[0];
    const int rc = pccb200_weight_normal( s.ctx, &const_cast<PCCPointSet3&>( f0 ).getPositions()[0][0], f0.getPointCount(),
                                          int( params.geometry3dCoordinatesBitdepth_ ) + 1, params.minWeightEPP_, sp.weight_normal );
    if ( rc != PCCB200_OK ) return rc;
  }
  // PCCPointSet3 keeps positions as vector<PCCVector3<int16_t>> and colours as vector<PCCVector3<uint8_t>>: n x 3 arrays
  static const int16_t        noXyz[3] = {0, 0, 0};
  static const uint8_t        noRgb[3] = {0, 0, 0};
  std::vector<const int16_t*> xyz( frameCount );
  std::vector<const uint8_t*> rgb( frameCount );
  std::vector<size_t>         n( frameCount );
  for ( size_t f = 0; f < frameCount; ++f ) {
    auto& ps = const_cast<PCCPointSet3&>( sources[f] );
    n[f]     = ps.getPointCount();
    xyz[f]   = n[f] ? &ps.getPositions()[0][0] : noXyz;
    rgb[f]   = n[f] ? &ps.getColors()[0][0] : noRgb;
  }
  int rc = pccb200_encode_gof( s.ctx, int( frameCount ), xyz.data(), rgb.data(), n.data(), &sp, int( params.occupancyPrecision_ ), 2, &s.gof );
  if ( rc != PCCB200_OK ) {
    fprintf( stderr, "pccb200: %s\n", pccb200_last_error( s.ctx ) );
    return rc;
  }
  params.initializeContext( context );  // PCCEncoder.cpp:107, unchanged
  size_t W = 0, H = 0;
  pccb200_gof_dims( s.gof, 0, &W, &H, nullptr );
  const size_t prec = params.occupancyPrecision_;
  auto&        videoOcc = context.getVideoOccupancyMap();
  videoOcc.resize( frameCount );
  if ( context.getVideoGeometryMultiple().empty() ) context.getVideoGeometryMultiple().resize( 1 );
  auto& videoGeo = context.getVideoGeometryMultiple()[0];
  videoGeo.resize( 2 * frameCount );
  std::vector<uint8_t>  b8;
  std::vector<uint16_t> b16;
  std::vector<uint32_t> b32;
  for ( size_t f = 0; f < frameCount; ++f ) {
    auto& tile       = context[f].getTile( 0 );
    tile.getWidth()  = W;
    tile.getHeight() = H;
    context[f].setAtlasFrameWidth( W );
    context[f].setAtlasFrameHeight( H );
    fillPatches( tile.getPatches(), pccb200_gof_patches( s.gof, int( f ) ) );
    if ( ( rc = fetch( s.gof, int( f ), PCCB200_GOF_OCCUPANCY, tile.getOccupancyMap(), b8 ) ) != PCCB200_OK ) return rc;
    if ( ( rc = fetch( s.gof, int( f ), PCCB200_GOF_BLOCK_TO_PATCH, tile.getBlockToPatch(), b32 ) ) != PCCB200_OK ) return rc;
    auto& om = videoOcc.getFrame( f );  // generateOccupancyMapVideo: luma = cell occupancy, chroma planes 0
    om.resize( W / prec, H / prec, PCCCOLORFORMAT::YUV420 );
    if ( ( rc = fetch( s.gof, int( f ), PCCB200_GOF_OM_VIDEO, om.getChannel( 0 ), b8 ) ) != PCCB200_OK ) return rc;
    for ( int m = 0; m < 2; ++m ) {     // generateGeometryVideo: D0 / D1 interleaved, luma only
      auto& img = videoGeo.getFrame( 2 * f + m );
      img.resize( W, H, PCCCOLORFORMAT::YUV444 );
      if ( ( rc = fetch( s.gof, int( f ), m ? PCCB200_GOF_GEO1 : PCCB200_GOF_GEO0, img.getChannel( 0 ), b16 ) ) != PCCB200_OK ) return rc;
    }
  }
  return PCCB200_OK;
}

namespace {
int fetchPlanes( pccb200_gof* gof, int f, int what, PCCImage<uint16_t, 3>& img, size_t W, size_t H, std::vector<uint16_t>& staging ) {
  const size_t n = pccb200_gof_get( gof, f, what, nullptr );
  if ( n != 3 * W * H ) return PCCB200_ERR_STATE;
  staging.resize( n );
  pccb200_gof_get( gof, f, what, staging.data() );
  img.resize( W, H, PCCCOLORFORMAT::RGB444 );
  for ( int c = 0; c < 3; ++c ) img.getChannel( c ).assign( staging.begin() + c * W * H, staging.begin() + ( c + 1 ) * W * H );
  return PCCB200_OK;
}
}  // namespace

int stageB1( Session& s, const PCCEncoderParameters& params, PCCContext& context, PCCGroupOfFrames& reconstructs,
             std::vector<std::vector<uint32_t>>& partitions ) {
  size_t W = 0, H = 0, R = 0;
  pccb200_gof_dims( s.gof, 0, &W, &H, nullptr );
  // lossless / pass-through geometry codec: reconstruct from the frames as produced. (With a lossy codec the decoded luma planes go
  // back first: pccb200_gof_set_decoded( gof, f, occVideo, geo0, geo1 ).)
  int rc = pccb200_gof_resume( s.gof, W, H, 0 );
  if ( rc != PCCB200_OK ) return rc;
  const size_t frameCount = context.size();
  partitions.assign( frameCount, std::vector<uint32_t>() );
  if ( context.getVideoAttributesMultiple().empty() ) context.getVideoAttributesMultiple().resize( 1 );
  auto& videoAttr = context.getVideoAttributesMultiple()[0];
  videoAttr.resize( 2 * frameCount );
  std::vector<int16_t>  xyz;
  std::vector<uint8_t>  rgb;
  std::vector<uint16_t> b16;
  std::vector<uint32_t> b32;
  const size_t          prec = params.occupancyPrecision_;
  for ( size_t f = 0; f < frameCount; ++f ) {
    pccb200_gof_dims( s.gof, int( f ), &W, &H, &R );
    auto&        tile = context[f].getTile( 0 );
    PCCPointSet3 rec;
    rec.resize( R );
    rec.addColors();
    xyz.resize( 3 * R ), rgb.resize( 3 * R ), b16.resize( R );
    if ( R ) {
      pccb200_gof_get( s.gof, int( f ), PCCB200_GOF_REC_XYZ, xyz.data() );
      pccb200_gof_get( s.gof, int( f ), PCCB200_GOF_REC_RGB, rgb.data() );
      pccb200_gof_get( s.gof, int( f ), PCCB200_GOF_REC_BOUNDARY, b16.data() );
    }
    for ( size_t i = 0; i < R; ++i ) {
      rec[i] = PCCPoint3D( xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] );
      rec.setColor( i, PCCColor3B( rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2] ) );
      rec.setBoundaryPointType( i, b16[i] );
    }
    reconstructs[f].clear();
    reconstructs[f].appendPointSet( rec );
    b32.resize( 3 * R );
    if ( R ) pccb200_gof_get( s.gof, int( f ), PCCB200_GOF_POINT_TO_PIXEL, b32.data() );
    auto& p2p = tile.getPointToPixel();
    p2p.resize( R );
    for ( size_t i = 0; i < R; ++i ) p2p[i] = PCCVector3<size_t>( b32[3 * i], b32[3 * i + 1], b32[3 * i + 2] );
    partitions[f].resize( R );
    if ( R ) pccb200_gof_get( s.gof, int( f ), PCCB200_GOF_REC_PARTITION, partitions[f].data() );
    // what generatePointCloud also leaves behind (PCCCodec.cpp:683, 843, 888-950): the patch of every point in the cloud itself and the
    // point counts of the tile, which colorPointCloud (:1348) and the conformance log (PCCEncoder.cpp:605) read
    for ( size_t i = 0; i < R; ++i ) reconstructs[f].setPointPatchIndex( i, 0, partitions[f][i] );
    tile.setTotalNumberOfRegularPoints( R );
    tile.setTotalNumberOfEOMPoints( 0 );
    tile.setTotalNumberOfRawPoints( 0 );
    // generatePointCloud leaves the block-precision occupancy in the tile (PCCCodec.cpp:559-572)
    auto&       occ = tile.getOccupancyMap();
    const auto& om  = context.getVideoOccupancyMap().getFrame( f ).getChannel( 0 );
    occ.assign( W * H, 0 );
    for ( size_t y = 0; y < H; ++y )
      for ( size_t x = 0; x < W; ++x ) occ[y * W + x] = om[( y / prec ) * ( W / prec ) + x / prec];
    for ( int m = 0; m < 2; ++m )
      if ( ( rc = fetchPlanes( s.gof, int( f ), m ? PCCB200_GOF_ATTR1_RAW : PCCB200_GOF_ATTR0_RAW, videoAttr.getFrame( 2 * f + m ), W, H, b16 ) ) != PCCB200_OK )
        return rc;
  }
  return PCCB200_OK;
}

int stageB2( Session& s, PCCContext& context ) {
  size_t W = 0, H = 0;
  pccb200_gof_dims( s.gof, 0, &W, &H, nullptr );
  auto&                 videoAttr = context.getVideoAttributesMultiple()[0];
  std::vector<uint16_t> b16;
  for ( size_t f = 0; f < context.size(); ++f )
    for ( int m = 0; m < 2; ++m ) {
      const int rc = fetchPlanes( s.gof, int( f ), m ? PCCB200_GOF_ATTR1 : PCCB200_GOF_ATTR0, videoAttr.getFrame( 2 * f + m ), W, H, b16 );
      if ( rc != PCCB200_OK ) return rc;
    }
  return PCCB200_OK;
}

int decodeFrame( Session& s, PCCContext& context, size_t frameIdx, size_t occupancyPrecision, PCCPointSet3& reconstruct,
                 std::vector<uint32_t>& partition ) {
  if ( !s.ctx ) {
    const int rc = pccb200_create( 0, &s.ctx );
    if ( rc != PCCB200_OK ) return rc;
  }
  auto&        tile = context[frameIdx].getTile( 0 );
  auto&        om   = context.getVideoOccupancyMap().getFrame( frameIdx );
  auto&        geo  = context.getVideoGeometryMultiple()[0];
  const size_t W = geo.getFrame( 2 * frameIdx ).getWidth(), H = geo.getFrame( 2 * frameIdx ).getHeight();
  // the fields the decoder rebuilds from the atlas syntax (PCCDecoder.cpp:900-1040) are all pccb200_generate_point_cloud reads
  std::vector<pccb200_patch> recs( tile.getPatches().size() );
  for ( size_t i = 0; i < recs.size(); ++i ) {
    const PCCPatch& p = tile.getPatches()[i];
    pccb200_patch&  r = recs[i];
    r                 = pccb200_patch{};
    r.index = int32_t( i ), r.view_id = int32_t( p.getViewId() );
    r.u1 = int32_t( p.getU1() ), r.v1 = int32_t( p.getV1() ), r.d1 = int32_t( p.getD1() );
    r.size_u = int32_t( p.getSizeU0() * 16 ), r.size_v = int32_t( p.getSizeV0() * 16 );
    r.size_u0 = int32_t( p.getSizeU0() ), r.size_v0 = int32_t( p.getSizeV0() );
    r.u0 = int32_t( p.getU0() ), r.v0 = int32_t( p.getV0() ), r.orientation = int32_t( p.getPatchOrientation() );
    r.best_match_idx = -1;
  }
  const std::vector<uint8_t>&  occ = om.getChannel( 0 );
  const std::vector<uint16_t>& g0  = geo.getFrame( 2 * frameIdx ).getChannel( 0 );
  const std::vector<uint16_t>& g1  = geo.getFrame( 2 * frameIdx + 1 ).getChannel( 0 );
  size_t                       R   = 0;
  int rc = pccb200_generate_point_cloud( s.ctx, recs.data(), int( recs.size() ), occ.data(), g0.data(), g1.data(), W, H, int( occupancyPrecision ), 0,
                                         nullptr, nullptr, nullptr, nullptr, &R );
  if ( rc != PCCB200_OK ) return rc;
  std::vector<int16_t>  xyz( 3 * R );
  std::vector<uint32_t> p2p( 3 * R );
  std::vector<uint16_t> bnd( R );
  partition.assign( R, 0 );
  rc = pccb200_generate_point_cloud( s.ctx, recs.data(), int( recs.size() ), occ.data(), g0.data(), g1.data(), W, H, int( occupancyPrecision ), R,
                                     xyz.data(), p2p.data(), partition.data(), bnd.data(), &R );
  if ( rc != PCCB200_OK ) return rc;
  reconstruct.clear();
  reconstruct.addColors();  // (generatePointCloud: PCCCodec.cpp:539)
  reconstruct.resize( R );
  auto& pointToPixel = tile.getPointToPixel();
  pointToPixel.resize( R );
  for ( size_t i = 0; i < R; ++i ) {
    reconstruct[i] = PCCPoint3D( xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2] );
    reconstruct.setBoundaryPointType( i, bnd[i] );
    pointToPixel[i] = PCCVector3<size_t>( p2p[3 * i], p2p[3 * i + 1], p2p[3 * i + 2] );
  }
  for ( size_t i = 0; i < R; ++i ) reconstruct.setPointPatchIndex( i, 0, partition[i] );
  tile.setTotalNumberOfRegularPoints( R );  // (PCCCodec.cpp:843, 888-950: read by colorPointCloud and the conformance log)
  tile.setTotalNumberOfEOMPoints( 0 );
  tile.setTotalNumberOfRawPoints( 0 );
  auto& map = tile.getOccupancyMap();  // block-precision occupancy, as generatePointCloud leaves it (PCCCodec.cpp:559-572)
  map.assign( W * H, 0 );
  for ( size_t y = 0; y < H; ++y )
    for ( size_t x = 0; x < W; ++x ) map[y * W + x] = occ[( y / occupancyPrecision ) * ( W / occupancyPrecision ) + x / occupancyPrecision];
  return PCCB200_OK;
}

namespace {
int needCtx( Session& s ) { return s.ctx ? PCCB200_OK : pccb200_create( 0, &s.ctx ); }
}  // namespace

int smoothGeometry( Session& s, PCCPointSet3& reconstruct, const std::vector<uint32_t>& partition, size_t gridSize, double threshold ) {
  int rc = needCtx( s );
  if ( rc != PCCB200_OK ) return rc;
  const size_t n = reconstruct.getPointCount();
  if ( n == 0 ) return PCCB200_OK;
  // positions_ is vector<PCCVector3<int16_t>>, boundaryPointTypes_ vector<uint16_t>: the arrays the ABI takes, updated in place
  return pccb200_smooth_geometry( s.ctx, &reconstruct.getPositions()[0][0], reconstruct.getBoundaryPointTypes().data(), partition.data(), n,
                                  int( gridSize ), threshold );
}

int transferColors16( Session& s, PCCPointSet3& source, PCCPointSet3& target ) {
  int rc = needCtx( s );
  if ( rc != PCCB200_OK ) return rc;
  if ( source.getPointCount() == 0 || target.getPointCount() == 0 ) return PCCB200_OK;
  return pccb200_transfer_colors16_smoothed( s.ctx, &source.getPositions()[0][0], &source.getColors16bit()[0][0], source.getPointCount(),
                                             &target.getPositions()[0][0], &target.getColors16bit()[0][0], target.getBoundaryPointTypes().data(),
                                             target.getPointCount() );
}

int yuv16ToRgb8( Session& s, PCCPointSet3& cloud ) {
  int rc = needCtx( s );
  if ( rc != PCCB200_OK ) return rc;
  const size_t n = cloud.getPointCount();
  if ( n == 0 ) return PCCB200_OK;
  return pccb200_yuv16_to_rgb8( s.ctx, &cloud.getColors16bit()[0][0], n, &cloud.getColors()[0][0] );
}

bool loadFrames( PCCGroupOfFrames& frames, const std::string& path, size_t startFrameNumber, size_t endFrameNumber,
                 PCCColorTransform colorTransform, size_t nbThread ) {
  static_assert( sizeof( PCCPoint3D ) == 3 * sizeof( int16_t ) && sizeof( PCCColor3B ) == 3, "point sets are read in place" );
  if ( endFrameNumber < startFrameNumber ) return false;
  const size_t count = endFrameNumber - startFrameNumber;
  frames.setFrameCount( count );
  if ( count == 0 ) return false;
  // pass 1, headers: point counts and colour flags; the group ends in front of the first unreadable file
  std::vector<size_t> n( count, 0 );
  std::vector<int>    colours( count, 0 );
  size_t              good = 0;
  pccb200_ply_read_frames( path.c_str(), startFrameNumber, endFrameNumber, nullptr, nullptr, nullptr, n.data(), colours.data(), int( nbThread ), &good );
  std::vector<int16_t*> xyz( count, nullptr );
  std::vector<uint8_t*> rgb( count, nullptr );
  for ( size_t k = 0; k < good; ++k ) {
    auto& ps = frames[k];
    ps.resize( 0 );
    ps.removeNormals();
    ps.removeReflectances();
    if ( colours[k] ) ps.addColors();
    else ps.removeColors();
    ps.resize( n[k] );
    xyz[k] = n[k] ? &ps.getPositions()[0][0] : nullptr;
    rgb[k] = ( n[k] && colours[k] ) ? &ps.getColors()[0][0] : nullptr;
  }
  // pass 2, bodies (a frame without points has nothing to parse: its buffers stay null, which asks for its header again)
  if ( good ) {
    std::vector<size_t> capacity( n );
    size_t              read = 0;
    pccb200_ply_read_frames( path.c_str(), startFrameNumber, startFrameNumber + good, xyz.data(), rgb.data(), capacity.data(), n.data(), nullptr,
                             int( nbThread ), &read );
    good = read;  // (smaller after a body line with too few values: the reference's read() fails there as well)
  }
  if ( good < count ) {
    char fileName[4096];
    snprintf( fileName, sizeof( fileName ), path.c_str(), startFrameNumber + good );
    printf( "Error: can't open %s\n", fileName );
    frames.setFrameCount( good );
  }
  if ( colorTransform == COLOR_TRANSFORM_RGB_TO_YCBCR )
    for ( size_t k = 0; k < good; ++k ) frames[k].convertRGBToYUV();
  return true;
}

}  // namespace pccb200shim
