/*
 * pccb200.h — C ABI of libpccb200.so: B200 (sm_100a) implementation of the V-PCC (TMC2 v24.0) encoder-side
 * patch-generation + packing + image-formation hot path and of PCCCodec::generatePointCloud.
 *
 * The reference has no FFI today (it is one C++ process); these are the entry points a maintainer binds from
 * PccLibEncoder / PccLibCommon in place of the CPU bodies.  Each entry point cites the reference interface it
 * replaces (paths relative to the reference checkout, source/lib/...).  INTEGRATION.md shows the C++ shim.
 *
 * Conventions: plain pointers + sizes, host buffers owned by the caller, device memory owned by the library.
 * Every function returns 0 on success or a negative pccb200_status; nothing calls exit().  The library needs
 * a CUDA device: there is no CPU fallback (pccb200_create fails with PCCB200_ERR_NO_DEVICE).
 */
#ifndef PCCB200_H
#define PCCB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pccb200_status {
  PCCB200_OK              = 0,
  PCCB200_ERR_NO_DEVICE   = -1, /* no CUDA device / driver: the product has no CPU path                 */
  PCCB200_ERR_CUDA        = -2, /* a CUDA call failed; see pccb200_last_error                           */
  PCCB200_ERR_BAD_ARG     = -3,
  PCCB200_ERR_STATE       = -4, /* stage called out of order                                            */
  PCCB200_ERR_CAPACITY    = -5, /* caller buffer too small                                              */
  PCCB200_ERR_UNSUPPORTED = -6, /* parameter combination outside the implemented (CTC) hot path         */
  PCCB200_ERR_CANVAS      = -180 /* patch2Canvas out of bounds (reference: exit(180), PCCPatch.cpp:238) */
} pccb200_status;

/* Parameters of PCCPatchSegmenter3Parameters that the CTC hot path reads
 * (PccLibEncoder/include/PCCPatchSegmenter.h:48-99, filled by PCCEncoder::generateSegments,
 *  PccLibEncoder/source/PCCEncoder.cpp:4672-4727). Defaults in comments = CTC all-intra r3 longdress. */
typedef struct pccb200_seg_params {
  int32_t nn_normal_estimation;            /* 16  */
  int32_t normal_orientation;              /* 1 = spanning tree (0 none) */
  int32_t max_nn_count_refine;             /* 1024 */
  int32_t iteration_count_refine;          /* 50 (longdress), 10 common */
  int32_t voxel_dim_refine;                /* 4   */
  int32_t search_radius_refine;            /* 192 */
  int32_t occupancy_resolution;            /* 16  */
  int32_t enable_patch_splitting;          /* 1   */
  int32_t max_patch_size;                  /* 1024 */
  int32_t quantizer_size_x;                /* 16 = 1<<log2QuantizerSizeX */
  int32_t quantizer_size_y;                /* 16  */
  int32_t min_point_count_per_cc;          /* 16  */
  int32_t max_nn_count_patch_seg;          /* 16  */
  int32_t surface_thickness;               /* 4   */
  int32_t min_level;                       /* 64  */
  int32_t max_allowed_depth;               /* 255 = (1<<geometryNominal2dBitdepth)-1 */
  int32_t geometry_bitdepth_2d;            /* 8   */
  int32_t geometry_bitdepth_3d;            /* 11 = geometry3dCoordinatesBitdepth+1 */
  int32_t map_count_minus1;                /* 1   */
  int32_t global_patch_allocation;         /* 0 = all-intra packing (constrainedPack 0, globalPatchAllocation 0: every frame packed on its
                                              own, PCCEncoder::packFlexible); 1 = random-access packing (cfg/condition/ctc-random-access.cfg:
                                              constrainedPack 1 + globalPatchAllocation 1: spatialConsistencyPackFlexible against the previous
                                              frame, then PCCEncoder::performDataAdaptiveGPAMethod over the GOF)                              */
  double  lambda_refine;                   /* 3.0 */
  double  max_allowed_dist2_raw_detection; /* 9.0 */
  double  max_allowed_dist2_raw_selection; /* 1.0 */
  double  weight_normal[3];                /* PCCEncoder::calculateWeightNormal (frame 0 of the GOF) */
} pccb200_seg_params;

/* One patch: the fields of PCCPatch (PccLibCommon/include/PCCPatch.h:353-409) that the hot path produces
 * and that downstream reference code (packing, createPatchFrameDataStructure, generatePointCloud) reads. */
typedef struct pccb200_patch {
  int32_t index;          /* creation order == PCCPatch::index_ */
  int32_t view_id;        /* 0..5, PCCPatch::setViewId (PCCPatch.cpp:111-138) */
  int32_t normal_axis, tangent_axis, bitangent_axis, projection_mode;
  int32_t u1, v1, d1;     /* 3-D offsets */
  int32_t size_u, size_v; /* depth-map size in pixels */
  int32_t size_d, size_d_pixel;
  int32_t size_u0, size_v0;   /* size in occupancy blocks */
  int32_t size_2d_x, size_2d_y; /* patchSize2D{X,Y}InPixel (quantised) */
  int32_t u0, v0, orientation;  /* filled by packing; -1 before */
  int32_t d0_count, eom_and_d1_count;
  int32_t best_match_idx; /* PCCPatch::bestMatchIdx_: index of the matched patch in the previous frame's list, -1 = none (always -1
                             in all-intra packing)                                                                             */
  int32_t is_global;      /* PCCPatch::isGlobalPatch_ (set by the global patch allocation)                                    */
  int64_t depth_offset;   /* into the int16 depth arena: depth_[0] then depth_[1], size_u*size_v each */
  int64_t occ_offset;     /* into the uint8 occupancy arena: size_u0*size_v0 */
} pccb200_patch;

typedef struct pccb200_ctx pccb200_ctx;

/* library / device lifetime */
int         pccb200_create( int device, pccb200_ctx** out );
void        pccb200_destroy( pccb200_ctx* ctx );
const char* pccb200_last_error( const pccb200_ctx* ctx );
const char* pccb200_version( void );
/* Device memory policy of the GOF entry points. A frame keeps ~0.3 GB (0.83 Mpts) while its orientation walk runs; the ~1.2 GB
 * of scratch the data-parallel stages before and after the walk need is leased from a per-device pool shared by all contexts
 * of the process. `count` (>= 1) bounds the number of scratch sets, i.e. of frames inside those stages at the same time
 * (default 24, or the environment variable PCCB200_SCRATCH_SETS). Returns the previous bound, or a negative pccb200_status. */
int         pccb200_set_scratch_sets( int device, int count );

/* Input side: PCCPointSet3::read (PccLibCommon/source/PCCPointSet.cpp:464-757) for what the hot path consumes - positions and
 * colours of one .ply frame (ascii or binary, any of the reference's property types for x/y/z, uchar red/green/blue, other
 * properties skipped), converted exactly as the reference's assignments convert them, written straight into caller buffers
 * (pin them: they are what pccb200_encode_gof uploads). Host code, no device needed. Call with xyz == NULL to get the point
 * count of the header in *n; capacity is in points; *has_colours (may be NULL) tells whether rgb was filled. */
int pccb200_ply_read( const char* path, int16_t* xyz, uint8_t* rgb, size_t capacity, size_t* n, int* has_colours );
/* PCCGroupOfFrames::load (PccLibCommon/source/PCCGroupOfFrames.cpp:46-83) on top of it: frame k (0 <= k < end_frame -
 * start_frame) is the file sprintf( path_pattern, start_frame + k ) - the pattern of --uncompressedDataPath, e.g.
 * "longdress_vox10_%04d.ply". xyz / rgb / capacity / n / has_colours are arrays with one entry per frame; xyz == NULL asks for
 * the point counts (n) and colour flags only. The frames are read by `threads` host threads (<= 0: all cores), several frames
 * at a time. Like the reference, the group ends at the first frame that cannot be read: *frames_read (may be NULL) receives
 * the number of good frames in front of it and the status of that frame is returned (PCCB200_OK when all were read; the
 * buffers of later frames may have been written). end_frame == start_frame is an empty group (OK, nothing read; the
 * reference's load returns false for it). Host code, no device needed. */
int pccb200_ply_read_frames( const char* path_pattern, size_t start_frame, size_t end_frame, int16_t* const* xyz, uint8_t* const* rgb,
                             const size_t* capacity, size_t* n, int* has_colours, int threads, size_t* frames_read );

/* Per-stage device timing (CUDA events on the launching stream). While enabled, every entry point appends one
 * (name, milliseconds) record per stage and frame; read returns and clears them. names: capacity x 32 chars; start_ms
 * (may be NULL) = start of the span relative to the beginning of the GOF call (-1 for the single-frame entry points). */
int pccb200_profile_enable( pccb200_ctx* ctx, int on );
int pccb200_profile_read( pccb200_ctx* ctx, char* names, float* ms, float* start_ms, int capacity, int* count );

/* ---- stage-level entry points -------------------------------------------------------------------------
 * Each replaces one public reference class used on its own by the tools (PccAppNormalGenerator uses PCCKdTree +
 * PCCNormalsGenerator3 directly) and is what the parity tests drive. Host buffers in, host buffers out. */

/* PCCKdTree::PCCKdTree + PCCKdTree::search (PccLibCommon/source/PCCKdTree.cpp:42-63): builds the nanoflann-
 * shaped tree over xyz (n x 3 int16, |coord| < 2048) and answers nq k-NN queries (k in {1,8,16}).
 * q == NULL means "the points themselves" (nq is then ignored). idx/dist2: row-major nq x k in nanoflann result
 * order; rows are padded with 0xFFFFFFFF / -1 when n < k. dist2 may be NULL. */
int pccb200_knn( pccb200_ctx* ctx, const int16_t* xyz, size_t n, const int16_t* q, size_t nq, int k, uint32_t* idx,
                 float* dist2 );
/* the tree's leaf order (nanoflann's vind after buildIndex); for tests of the builder */
int pccb200_kdtree_order( pccb200_ctx* ctx, const int16_t* xyz, size_t n, uint32_t* vind );

/* PCCNormalsGenerator3::compute (PccLibEncoder/source/PCCNormalsGenerator.cpp:61-70) with the parameters
 * PCCPatchSegmenter3::compute passes (PCCPatchSegmenter.cpp:88-107): k-NN PCA normals (view point = origin),
 * orientation 0 = none, 1 = spanning tree. normals: n x 3 doubles. */
int pccb200_normals( pccb200_ctx* ctx, const int16_t* xyz, size_t n, int k, int orientation, double* normals );

/* PCCEncoder::calculateWeightNormal (PccLibEncoder/source/PCCEncoder.cpp:3569-3626), enhancedPP on: axis weights
 * from the projected areas of frame 0 of the GOF. bits = geometry3dCoordinatesBitdepth + 1, min_weight = minWeightEPP. */
int pccb200_weight_normal( pccb200_ctx* ctx, const int16_t* xyz, size_t n, int bits, double min_weight, double w[3] );

/* PCCPatchSegmenter3::compute (PccLibEncoder/source/PCCPatchSegmenter.cpp:53-150), CTC path: k-d tree, normals +
 * orientation, initial segmentation, grid-based refinement, patch segmentation (a1..a11 of SURVEY.md §8a) for ONE
 * frame. rgb is n x 3 uint8 (the D1 colour gate reads it). The optional outputs expose the intermediate results
 * the reference keeps in PCCNormalsGenerator3::normals_ and `partition` (before / after refinement); pass NULL to
 * skip them. The patch list is returned as an opaque handle (pccb200_patches_*). */
typedef struct pccb200_patchlist pccb200_patchlist;
int    pccb200_segment_frame( pccb200_ctx* ctx, const int16_t* xyz, const uint8_t* rgb, size_t n,
                              const pccb200_seg_params* params, double* normals, uint8_t* partition_initial,
                              uint8_t* partition_refined, pccb200_patchlist** out );
int    pccb200_patches_count( const pccb200_patchlist* pl );
size_t pccb200_patches_depth_elems( const pccb200_patchlist* pl );
size_t pccb200_patches_occ_elems( const pccb200_patchlist* pl );
/* copies patch records, both depth maps of every patch (int16, re-based to d1, 32767 = empty) and the block
 * occupancy (uint8) into caller buffers sized by the three calls above */
int    pccb200_patches_get( const pccb200_patchlist* pl, pccb200_patch* patches, int16_t* depth_arena, uint8_t* occ_arena );
void   pccb200_patches_free( pccb200_patchlist* pl );

/* ---- GOF-level entry points ------------------------------------------------------------------------------
 * PCCEncoder::encode between the source clouds and the three videoEncoder.compress calls
 * (PccLibEncoder/source/PCCEncoder.cpp:103-424): generateSegments (:103) -> placeSegments (:110) -> generateOccupancyMap
 * + generateOccupancyMapVideo (:133,:139) -> generateBlockToPatchFromOccupancyMapVideo (:168) -> generateGeometryVideo
 * (:172) -> generatePointCloud (:328) -> generateAttributeVideo (:341) -> attribute padding (:344-424), all frames of the
 * GOF concurrently (one stream per frame). The occupancy/geometry videos are treated as losslessly coded (decoded ==
 * source), which is what the reference sees for occupancy in CTC and what a passthrough codec gives for geometry.
 * params->weight_normal must hold pccb200_weight_normal() of frame 0 (PCCEncoder.cpp:4726).
 * stop_after: 0 = all stages, 1 = after packing, 2 = after the geometry images, 3 = after generatePointCloud, 5 = after the segmentation
 * (no packing; random access only: see pccb200_gof_pack_ra).
 * One GOF is live per context: a later pccb200_encode_gof reuses the device buffers of the earlier one. */
typedef struct pccb200_gof pccb200_gof;
int  pccb200_encode_gof( pccb200_ctx* ctx, int nframes, const int16_t* const* xyz, const uint8_t* const* rgb, const size_t* n,
                         const pccb200_seg_params* params, int occupancy_precision, int stop_after, pccb200_gof** out );
/* Multi-GPU (frames of one GOF sharded over ranks): call pccb200_encode_gof with stop_after = 1 on every rank, all-reduce
 * (MAX) the canvas size from pccb200_gof_dims — the one cross-frame reduction of the all-intra path
 * (PCCEncoder::resizeGeometryVideo, PCCEncoder.cpp:5546-5591) — then resume the remaining stages on that canvas.
 * A rank need not wait for the reduction: it may resume on its local size at once and, should the reduced size turn out larger,
 * call pccb200_gof_resume( gof, W, H, 0 ) again on the finished GOF with the larger canvas (all products are formed anew). */
int  pccb200_gof_resume( pccb200_gof* gof, size_t width, size_t height, int stop_after );
/* Multi-GPU, random access (global_patch_allocation = 1): every frame is packed against the previous one
 * (PCCEncoder::placeSegments, PCCEncoder.cpp:4778-4805) and the global patch allocation iterates over the whole GOF
 * (performDataAdaptiveGPAMethod, :6838-6970), so a sharded GOF needs ONE exchange of patch metadata. Each rank
 * calls pccb200_encode_gof with stop_after = 5 (segmentation only), the ranks all-gather the records of pccb200_gof_patches
 * (pccb200_patches_get with depth = NULL: patch fields + 16-pixel block occupancies, KBs per frame), and every rank calls
 * pccb200_gof_pack_ra with the records of ALL frames of the GOF in frame order:
 *   patch_counts[f]     patches of frame f, f = 0 .. total_frames-1
 *   patches             the frames' records, concatenated; occ_offset relative to the frame's own occupancy block
 *   occ, occ_sizes[f]   the frames' block occupancies, concatenated / bytes per frame
 *   local_frame[f]      index of frame f among the frames of `gof`, or -1 when another rank holds it
 * The packing is deterministic: all ranks compute the same placements, each installs those of its own frames. Afterwards the GOF is
 * in the state stop_after = 1 leaves (pccb200_gof_dims returns the GOF-wide canvas); continue with pccb200_gof_resume. */
int  pccb200_gof_pack_ra( pccb200_gof* gof, int total_frames, const int* patch_counts, const pccb200_patch* patches, const uint8_t* occ,
                          const size_t* occ_sizes, const int* local_frame );
/* Lossy geometry codec: run pccb200_encode_gof / pccb200_gof_resume with stop_after = 2, hand GOF_OM_VIDEO / GOF_GEO0 / GOF_GEO1 to
 * the codec, give the DECODED luma planes back (any pointer may be NULL = keep the source), then pccb200_gof_resume(gof, W, H, 0)
 * reconstructs from them exactly as PCCEncoder::encode does after videoEncoder.compress replaced `video` by the reconstruction
 * (PCCVideoEncoder.cpp:403-416, PCCEncoder.cpp:168, 319-334). */
int  pccb200_gof_set_decoded( pccb200_gof* gof, int f, const uint8_t* occ_video, const uint16_t* geo0, const uint16_t* geo1 );
void pccb200_gof_free( pccb200_gof* gof );

/* PCCCodec::generatePointCloud (PccLibCommon/source/PCCCodec.cpp:519-980) as PCCDecoder::decode calls it
 * (PccLibDecoder/source/PCCDecoder.cpp:334-351), CTC reconstruction options (two maps, absolute D1, duplicate removal, no EOM/PLR/
 * raw patches): `patches` in atlas order with the fields the decoder rebuilds from the syntax (u0, v0, orientation, size_u0,
 * size_v0, u1, v1, d1, view_id; PCCDecoder.cpp:900-1040), decoded occupancy video luma ((W/p)*(H/p) uint8), decoded geometry luma
 * D0/D1 (W*H uint16). Outputs (each may be NULL) hold up to `capacity` points: positions (n x 3 int16), pointToPixel (n x 3
 * uint32: x, y, map), partition (patch index) and boundary point type. *rec_points receives the true count; when it exceeds
 * capacity and an output was requested the call returns PCCB200_ERR_CAPACITY (call once with capacity 0 / NULL outputs to size). */
int  pccb200_generate_point_cloud( pccb200_ctx* ctx, const pccb200_patch* patches, int num_patches, const uint8_t* occ_video,
                                   const uint16_t* geo0, const uint16_t* geo1, size_t width, size_t height, int occupancy_precision,
                                   size_t capacity, int16_t* xyz, uint32_t* point_to_pixel, uint32_t* partition, uint16_t* boundary,
                                   size_t* rec_points );
/* canvas size (identical for all frames of the GOF) and number of reconstructed points of frame f */
int  pccb200_gof_dims( const pccb200_gof* gof, int f, size_t* width, size_t* height, size_t* rec_points );
/* patches of frame f in PACKED order (the order of tile.getPatches() after packFlexible) with u0/v0/orientation filled;
 * borrowed: valid until pccb200_gof_free */
const pccb200_patchlist* pccb200_gof_patches( const pccb200_gof* gof, int f );

/* products of frame f; pccb200_gof_get returns the element count and, when dst != NULL, copies device -> host */
enum {
  PCCB200_GOF_OCCUPANCY      = 1,  /* uint8  W*H          PCCFrameContext::occupancyMap_ before generatePointCloud          */
  PCCB200_GOF_OM_VIDEO       = 2,  /* uint8  (W/p)*(H/p)  luma plane of the occupancy video frame (-> HM)                   */
  PCCB200_GOF_BLOCK_TO_PATCH = 3,  /* uint32 (W/16)*(H/16) PCCFrameContext::blockToPatch_ (patch index + 1, packed order)   */
  PCCB200_GOF_GEO0           = 4,  /* uint16 W*H          luma plane of geometry frame 2f   (D0, dilated) (-> HM)           */
  PCCB200_GOF_GEO1           = 5,  /* uint16 W*H          luma plane of geometry frame 2f+1 (D1, dilated) (-> HM)           */
  PCCB200_GOF_REC_XYZ        = 6,  /* int16  R*3          reconstructed positions, reference emission order                 */
  PCCB200_GOF_POINT_TO_PIXEL = 7,  /* uint32 R*3          PCCFrameContext::pointToPixel_ (x, y, map)                        */
  PCCB200_GOF_REC_PARTITION  = 8,  /* uint32 R            patch index (packed order) of every reconstructed point           */
  PCCB200_GOF_REC_BOUNDARY   = 9,  /* uint16 R            PCCPointSet3 boundary point type after identifyBoundaryPoints     */
  PCCB200_GOF_REC_RGB        = 10, /* uint8  R*3          colours after transferColors                                      */
  PCCB200_GOF_ATTR0_RAW      = 11, /* uint16 3*W*H planar R,G,B: attribute frame 2f   before padding                        */
  PCCB200_GOF_ATTR1_RAW      = 12, /* uint16 3*W*H        attribute frame 2f+1 before padding                               */
  PCCB200_GOF_ATTR0          = 13, /* uint16 3*W*H        attribute frame 2f   after push-pull padding (-> colour conversion -> HM) */
  PCCB200_GOF_ATTR1          = 14, /* uint16 3*W*H        attribute frame 2f+1 after push-pull padding                      */
  /* the padded attribute frames as PCCVideoEncoder::compress hands them to the codec (PccLibEncoder/source/PCCVideoEncoder.cpp:
   * 326-353): RGB444 -> YUV 4:2:0, 8 bit, PCCInternalColorConverter "RGB444ToYUV420_8_4" (the default down-sampling filter).
   * Planes Y (W*H), U, V ((W/2)*(H/2) each); converted on the device when requested: a quarter of the bytes of ATTR0/ATTR1. */
  PCCB200_GOF_ATTR0_YUV420   = 15, /* uint8  W*H*3/2 */
  PCCB200_GOF_ATTR1_YUV420   = 16, /* uint8  W*H*3/2 */
  /* the geometry luma planes as the 8-bit codec input holds them (PCCVideo::write with one byte per sample,
   * PccLibCommon/include/PCCImage.h / PCCVideo.h; geometryNominal2dBitdepth 8): the values of GEO0 / GEO1 narrowed on the device, half the
   * bytes over PCIe. The chroma planes of the 4:2:0 file are constant (the geometry frames carry luma only). */
  PCCB200_GOF_GEO0_LUMA8     = 17, /* uint8  W*H */
  PCCB200_GOF_GEO1_LUMA8     = 18  /* uint8  W*H */
};
size_t pccb200_gof_get( pccb200_gof* gof, int f, int what, void* dst );

/* ---- post-reconstruction chain: what follows generatePointCloud in PCCEncoder::encode
 * (PccLibEncoder/source/PCCEncoder.cpp:556-719) and PCCDecoder::decode (PccLibDecoder/source/PCCDecoder.cpp:356-475) under the CTC.
 * Host buffers in / out like every other entry point. */

/* PCCCodec::smoothPointCloudPostprocess, grid-based geometry smoothing (PccLibCommon/source/PCCCodec.cpp:54-150, 982-1168;
 * CTC: gridSmoothing on, gridSize 8, thresholdSmoothing 64). In place: xyz (n x 3) receives the smoothed positions, boundary[]
 * the updated boundary point types (a moved point becomes type 3, PCCCodec.cpp:1139); partition[] = patch index of every point. */
int pccb200_smooth_geometry( pccb200_ctx* ctx, int16_t* xyz, uint16_t* boundary, const uint32_t* partition, size_t n, int grid_size,
                             double threshold );

/* PCCPointSet3::transferColors16bitBP as encode / decode call it after the smoothing (PccLibCommon/source/PCCPointSet.cpp:1126-1485:
 * 8-NN forward, 1-NN backward votes with the colour gate, fp64 sums in std::sort order): new 16-bit colours, in place in tgt_col
 * (T x 3), for the target points of boundary type 3; src = the cloud before smoothing with its colours. */
int pccb200_transfer_colors16_smoothed( pccb200_ctx* ctx, const int16_t* src_xyz, const uint16_t* src_col, size_t S, const int16_t* tgt_xyz,
                                        uint16_t* tgt_col, const uint16_t* tgt_boundary, size_t T );

/* The two ends of colorPointCloud (PCCCodec.cpp:1319-1460): the decoded attribute frame YUV 4:2:0 (8 bit, planes Y, U, V) ->
 * YUV 4:4:4 16 bit (PCCInternalColorConverter "YUV420ToYUV444_16_...": PccLibColorConverter/source/PCCInternalColorConverter.cpp:
 * 425-470, up-sampling filter 0), planar 3 x W x H; and PCCPointSet3::convertYUV16ToRGB8 of the point colours (n x 3 -> n x 3). */
int pccb200_yuv420_to_yuv444_16( pccb200_ctx* ctx, const uint8_t* yuv420, size_t width, size_t height, uint16_t* yuv444 );
int pccb200_yuv16_to_rgb8( pccb200_ctx* ctx, const uint16_t* yuv, size_t n, uint8_t* rgb );

#ifdef __cplusplus
}
#endif
#endif /* PCCB200_H */
