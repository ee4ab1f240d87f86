// pccb200_shim.h — the reference-side binding of libpccb200.so: what a maintainer adds to PccLibEncoder so that
// PCCEncoder::encode (PccLibEncoder/source/PCCEncoder.cpp:71-424) runs its hot path on the GPU and everything else
// (PccAppEncoder, bitstream writer, video-codec wrappers, post-processing, metrics) stays as it is.
//
// It is written against the reference's own headers and types (C++14) and only calls the C ABI of include/pccb200.h.
// The three stage functions are drop-in bodies for the calls PCCEncoder::encode makes:
//   stageA  : generateSegments (:103) + placeSegments (:110) + generateOccupancyMap (:133) + generateOccupancyMapVideo (:139)
//             + generateBlockToPatchFromOccupancyMapVideo (:168) + generateGeometryVideo (:172)
//             (params_.initializeContext( context ), :107, stays between "segments" and "placement": it is called from here)
//   stageB1 : the generatePointCloud loop (:319-334) + generateAttributeVideo (:341)
//   stageB2 : the attribute padding loop (:344-424)
// In-tree this file is compiled and linked by oracle/Makefile (target `shim`, against the reference sources where they lie)
// and exercised by tests/test_shim.py: on a GPU box the reference's own data structures, filled through this shim, must be
// identical to what the unmodified reference stages leave in them.
#pragma once
#include <vector>

#include "PCCContext.h"
#include "PCCEncoderParameters.h"
#include "PCCGroupOfFrames.h"
#include "pccb200.h"

namespace pccb200shim {

struct Session {  // one per PCCEncoder object
  pccb200_ctx* ctx = nullptr;
  pccb200_gof* gof = nullptr;
  ~Session();
};

// the fields PCCEncoder::generateSegments copies into PCCPatchSegmenter3Parameters (PCCEncoder.cpp:4672-4727), plus the packing mode
pccb200_seg_params toSegParams( const pcc::PCCEncoderParameters& p );

// each returns 0 or a negative pccb200_status (the caller keeps the reference's convention: print + exit)
int stageA( Session& s, pcc::PCCEncoderParameters& params, const pcc::PCCGroupOfFrames& sources, pcc::PCCContext& context );
int stageB1( Session& s, const pcc::PCCEncoderParameters& params, pcc::PCCContext& context, pcc::PCCGroupOfFrames& reconstructs,
             std::vector<std::vector<uint32_t>>& partitions );
int stageB2( Session& s, pcc::PCCContext& context );

// Decoder side (PccLibDecoder/source/PCCDecoder.cpp:320-351): drop-in body for the per-frame block
//   generateBlockToPatchFromOccupancyMapVideo( ... ) + generatePointCloud( tileReconstrct, context, frameIdx, tileIdx, gpcParams, partition, true )
// for the CTC reconstruction options (two maps, absolute D1, duplicate removal, no EOM / PLR / raw patches, one tile): the patches of
// the tile as the decoder rebuilt them from the atlas syntax, the decoded occupancy and geometry frames of `context`; fills
// `reconstruct` (positions + boundary point types), tile.getPointToPixel(), tile.getOccupancyMap() (the
// block-precision map generatePointCloud leaves there) and `partition`. Needs s.ctx only (no GOF).
int decodeFrame( Session& s, pcc::PCCContext& context, size_t frameIdx, size_t occupancyPrecision, pcc::PCCPointSet3& reconstruct,
                 std::vector<uint32_t>& partition );

// Post-reconstruction chain as PCCEncoder::encode (PccLibEncoder/source/PCCEncoder.cpp:647-702) and
// PCCDecoder::decode (PccLibDecoder/source/PCCDecoder.cpp:401-470) run it under the CTC - drop-in bodies for three calls:
//   smoothGeometry   : smoothPointCloudPostprocess( reconstruct, colorTransform, ppSEIParams, partition ) with gridSmoothing
//                      (positions and boundary point types of `reconstruct` are updated in place)
//   transferColors16 : tempFrameBuffer.transferColors16bitBP( reconstruct, filterType 1, 0, false, 8, 1, true, true, true, false, 4, 4,
//                      1000, 1000, 1000 * 256, 1000 * 256 ) - new 16-bit colours for the points the smoothing moved
//   yuv16ToRgb8      : reconstruct.convertYUV16ToRGB8()
int smoothGeometry( Session& s, pcc::PCCPointSet3& reconstruct, const std::vector<uint32_t>& partition, size_t gridSize, double threshold );
int transferColors16( Session& s, pcc::PCCPointSet3& source, pcc::PCCPointSet3& target );
int yuv16ToRgb8( Session& s, pcc::PCCPointSet3& cloud );

// Input side: drop-in body for PCCGroupOfFrames::load( path, start, end, colorTransform, false, nbThread )
// (PccLibCommon/source/PCCGroupOfFrames.cpp:46-83; the call of PccAppEncoder.cpp:1048) - positions and colours of every frame,
// parsed by pccb200_ply_read_frames straight into the point sets' own storage, several frames at a time. Same result and return
// value as the reference's: the group is cut at the first frame that cannot be read, false only for an empty range. Normals and
// reflectances, which the encoder does not read from its sources, are not loaded. Host code, no device needed.
bool loadFrames( pcc::PCCGroupOfFrames& frames, const std::string& path, size_t startFrameNumber, size_t endFrameNumber,
                 pcc::PCCColorTransform colorTransform, size_t nbThread );

}  // namespace pccb200shim
