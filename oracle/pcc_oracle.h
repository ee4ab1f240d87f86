/*
 * pcc_oracle.h — C ABI of the CPU *oracle* for the V-PCC patch-generation / image-formation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a from-scratch, single-threaded CPU restatement of the reference
 * algorithms (MPEGGroup/mpeg-pcc-tmc2 v24.0); each function cites the reference file:line it restates.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load it.
 * The product (mpeg-pcc-tmc2_b200/csrc, libpccb200.so) never links, imports or calls anything here.
 *
 * Parity status: PINNED — every entry point is checked against the reference itself compiled from
 * /root/reference (oracle/_ref/libtmc2ref.so, see oracle/Makefile + oracle/ref_harness.cpp) in
 * tests/test_oracle_vs_ref.py, and against committed fixtures in tests/golden/ generated from that build.
 */
#ifndef PCC_ORACLE_H
#define PCC_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/pccb200.h" /* shared POD parameter / patch structs (layout contract only) */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- a1: nanoflann-equivalent kd-tree (PCCKdTree.cpp:42-79, nanoflann.hpp:1041-1254) ---------------- */
void*  pcco_kdtree_build( const int16_t* xyz, size_t n );
void   pcco_kdtree_free( void* tree );
/* permutation of point indices in tree (leaf) order == nanoflann's vind after buildIndex */
void   pcco_kdtree_vind( void* tree, uint32_t* vind );
/* k-NN of nq queries; idx/dist2 are row-major nq x k, rows padded with 0xFFFFFFFF / -1 when the tree holds
 * fewer than k points. Order inside a row == nanoflann KNNResultSet order (distance, then first-visited). */
void   pcco_knn( void* tree, const int16_t* q, size_t nq, int k, uint32_t* idx, float* dist2 );
/* radius search (dist2 < radius2), results sorted by (dist2, index) as nanoflann's IndexDist_Sorter does,
 * truncated to max_results per query (PCCKdTree.cpp:65-79). CSR output; returns total entries written.
 * Call with idx==NULL to size. offsets has nq+1 entries. */
size_t pcco_radius( void* tree, const int16_t* q, size_t nq, double radius2, size_t max_results, uint64_t* offsets,
                    uint32_t* idx, float* dist2 );

/* ---- a2/a3: normals (PCCNormalsGenerator.cpp:71-185) and orientation (:198-242, :521-548) ---------- */
/* nbr: n x k neighbour lists from pcco_knn on the same cloud. normals: n x 3 doubles.
 * orientation: 0 none, 1 spanning tree. */
void pcco_normals( const int16_t* xyz, size_t n, const uint32_t* nbr, int k, double* normals );
void pcco_orient_normals( const int16_t* xyz, size_t n, const uint32_t* nbr, int k, double* normals );

/* ---- a4: axis weights (PCCEncoder.cpp:3569-3626) ----------------------------------------------------- */
void pcco_weight_normal( const int16_t* xyz, size_t n, int geometry_bitdepth_3d, double min_weight_epp, double w[3] );

/* ---- a5/a6: initial + grid-based refined segmentation (PCCPatchSegmenter.cpp:226-265, 1386-1561) ----- */
void pcco_initial_segmentation( const double* normals, size_t n, const double w[3], uint8_t* partition );
void pcco_refine_segmentation( const int16_t* xyz, const double* normals, size_t n, const pccb200_seg_params* p,
                               uint8_t* partition );

/* ---- a7–a11: patch segmentation (PCCPatchSegmenter.cpp:537-1320) ------------------------------------ */
/* Returns an opaque patch list; query with the accessors below. */
void*  pcco_segment_patches( const int16_t* xyz, const uint8_t* rgb, size_t n, const uint32_t* nbr, int k,
                             const uint8_t* partition, const pccb200_seg_params* p );
int    pcco_patches_count( void* pl );
size_t pcco_patches_depth_elems( void* pl );
size_t pcco_patches_occ_elems( void* pl );
void   pcco_patches_get( void* pl, pccb200_patch* patches, int16_t* depth_arena, uint8_t* occ_arena );
void   pcco_patches_free( void* pl );

/* ---- whole-frame convenience: a1..a11 (PCCPatchSegmenter3::compute, PCCPatchSegmenter.cpp:53-150) ---- */
void* pcco_segment_frame( const int16_t* xyz, const uint8_t* rgb, size_t n, const pccb200_seg_params* p );

/* ---- a13–a26 for one GOF (PCCEncoder::encode :103-424 with a lossless codec); products by id, see tests/bindings.py
 * stop_after: 0 all, 1 packing, 2 geometry images, 3 generatePointCloud */
void*  pcco_encode_gof( int nframes, const int16_t* const* xyz, const uint8_t* const* rgb, const size_t* n,
                        const pccb200_seg_params* p, int occupancy_precision, int stop_after );
/* same with a lower bound on the canvas (frames of one GOF sharded over ranks: pass the all-reduced size) */
void*  pcco_encode_gof_canvas( int nframes, const int16_t* const* xyz, const uint8_t* const* rgb, const size_t* n,
                               const pccb200_seg_params* p, int occupancy_precision, int stop_after, size_t force_w, size_t force_h );
/* a13-a15 alone: packing of caller-given patch lists (records + block occupancy, frame f's occupancy bytes start at occ_base[f]) */
void*  pcco_pack_gof( int nframes, const int* counts, const pccb200_patch* patches, const uint8_t* occ, const int64_t* occ_base, int ra, int bits );
void   pcco_gof_free( void* h );
void   pcco_gof_dims( void* h, int f, size_t* w, size_t* hgt, size_t* rec_points );
void*  pcco_gof_patches( void* h, int f ); /* borrowed patch list for pcco_patches_* */
size_t pcco_gof_get( void* h, int f, int what, void* dst );

/* ---- §8f-1, first stage: grid-based geometry smoothing of the reconstructed cloud (PCCCodec.cpp:54-150, 982-1106); in place */
void   pcco_smooth_geometry( int16_t* xyz, uint16_t* boundary, const uint32_t* partition, size_t n, int grid_size, double threshold );

/* ---- §8f-1, second stage: PCCPointSet3::transferColors16bitBP as encode / decode call it (PCCPointSet.cpp:1126-1485,
 * PCCEncoder.cpp:656-672): new 16-bit colours for the target points the geometry smoothing moved (boundary type 3); in place */
void   pcco_transfer_colors16_smoothed( const int16_t* src_xyz, const uint16_t* src_col, size_t ns, const int16_t* tgt_xyz, uint16_t* tgt_col,
                                        const uint16_t* tgt_boundary, size_t nt );

/* ---- §8f-1, the ends of the chain: decoded 8-bit YUV 4:2:0 frame -> 16-bit YUV 4:4:4 planes (PCCInternalColorConverter
 * "YUV420ToYUV444_8_0"), and PCCPointSet3::convertYUV16ToRGB8 per point */
void   pcco_yuv420_to_yuv444_16( const uint8_t* yuv420, size_t width, size_t height, uint16_t* yuv444 );
void   pcco_yuv16_to_rgb8( const uint16_t* yuv, size_t n, uint8_t* rgb );

#ifdef __cplusplus
}
#endif
#endif
