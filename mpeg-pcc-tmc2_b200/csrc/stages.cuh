// stages.cuh — internal stage interfaces of libpccb200 (device pointers, one stream per frame).
#pragma once
#include "common.cuh"
#include "kdtree.cuh"
#include "ra_pack.hpp"

namespace pccb200 {

// normals.cu
void computeNormals( const short4* pts, const uint32_t* nbr, int k, size_t n, double* normals, cudaStream_t s );
void initialSegmentation( const double* normals, size_t n, const double w[3], uint8_t* partition, cudaStream_t s );

// orient.cu
// (temporaries of orientPrepare: dead once the rows are built - part of the leased FrameScratch below)
struct OrientTemp {
  DevBuf<uint32_t> nbrSorted, relBits, idsA, idsB, rankRel;
  DevBuf<uint64_t> keysA, keysB;
  DevBuf<uint8_t>  cubTmp;
};
// (what the walk itself and orientFinish read: stays with the frame)
struct OrientScratch {
  DevBuf<uint32_t> best, pos, rankEnd;
  DevBuf<uint2>    rows;
  DevBuf<uint64_t> L0;
  DevBuf<uint8_t>  flip;
  DevBuf<unsigned> counter;
  Profiler*        prof = nullptr;
  alignas( 8 ) unsigned char walkArgs[160];  // the frame's walk arguments, filled by orientPrepare (opaque here)
  size_t walkSmem = 0;                       // dynamic shared memory of its walk (0: nothing to walk)
};
// Three steps, so that the walks of all frames of a GOF can share one launch: data-parallel preparation on the frame's stream,
// the walks of a batch of prepared frames (blocks the calling host thread until they are done), sign application.
void orientPrepare( OrientScratch& sc, OrientTemp& tmp, const short4* pts, const uint32_t* nbr, int k, const uint32_t* vind, size_t n, double* normals,
                    cudaStream_t s );
void orientWalkBatch( OrientScratch* const* frames, int count, DevBuf<unsigned char>& devArgs, Profiler* prof, cudaStream_t s );
void orientFinish( OrientScratch& sc, const short4* pts, size_t n, double* normals, cudaStream_t s );
// all three for one frame
void orientNormals( OrientScratch& sc, OrientTemp& tmp, DevBuf<unsigned char>& devArgs, const short4* pts, const uint32_t* nbr, int k, const uint32_t* vind, size_t n,
                    double* normals, cudaStream_t s );

// refine.cu
struct RefineScratch {
  DevBuf<int>                ints, grid, offsets;
  DevBuf<uint32_t>           keysA, keysB, idsA, idsSorted, head, headScan, scanTmp;
  DevBuf<uint32_t>           runStart, runFirst, runId, runFirstSorted, order, voxStart, voxCount, runToVox;
  DevBuf<short4>             centers;
  DevBuf<uint32_t>           adjOff, adjLen, adjData, nearData, list;
  DevBuf<uint8_t>            nearLen, edge, ppi, dirty, mark, active, cubTmp;
  DevBuf<double>             weight;
  DevBuf<uint16_t>           score, smooth;
  DevBuf<unsigned long long> cursor;
  DevBuf<unsigned>           count;
  int                        offsetsR2 = -1, numOffsets = 0;
};
void refineSegmentation( RefineScratch& sc, const short4* pts, const double* normals, size_t n, const pccb200_seg_params& prm,
                         uint8_t* partition, cudaStream_t s );

// patches.cu
struct PatchScratch {
  DevBuf<uint8_t>            raw, minD2, seedView;
  DevBuf<uint16_t>           mutual;
  DevBuf<uint32_t>           parent, compLabel, compSize, label, kept, keptScan, scanTmp, bitmap, owner, seedIdx;
  DevBuf<int>                member, ints, d0, d1, peak;
  DevBuf<unsigned long long> keys;
  DevBuf<long long>          depthOff, occOff;
  DevBuf<char>               stats, devPatches, counters;  // typed inside patches.cu
};
struct PatchResult {
  std::vector<pccb200_patch> patches;
  DevBuf<int16_t>            depth;  // arena on the device
  DevBuf<uint8_t>            occ;
  size_t                     depthElems = 0, occElems = 0;
  int                        outerIterations = 0;
};
void packRgb( const uint8_t* rgb3, size_t n, uchar4* out, cudaStream_t s );
void segmentPatches( PatchScratch& sc, PatchResult& out, const short4* pts, const uchar4* rgb, const uint32_t* nbr, int k,
                     const uint8_t* partition, size_t n, const pccb200_seg_params& prm, cudaStream_t s );

// canvas.cu
struct CanvasPatch {  // device-side patch record in packed order
  int       viewId, u1, v1, d1, sizeU, sizeV, sizeU0, sizeV0, u0, v0, orientation, pad;
  long long depthOff, occOff;
};
struct CanvasImages {
  DevBuf<uint8_t>  occ, om, blockEmpty;
  DevBuf<uint16_t> geo0, geo1;
  DevBuf<uint32_t> blockToPatch;
  DevBuf<int>      error;
};
struct ReconTemp {
  DevBuf<uint32_t> counts, offsets, scanTmp;
};
struct ReconScratch {  // the reconstructed cloud (a product: stays with the frame)
  DevBuf<uint32_t> pointToPixel, recPartition;
  DevBuf<short4>   recXyz;
  DevBuf<uint16_t> boundary;
  size_t           numPoints = 0;
};
struct AttrTemp {
  DevBuf<ushort4>               T[2], tmp, tmp2;
  DevBuf<uint8_t>               occ;
  std::vector<DevBuf<ushort4>>  mip;
  std::vector<DevBuf<uint8_t>>  mipOcc;
};
struct AttrImages {  // products: attribute frames before / after padding
  DevBuf<uint16_t> rawPlanes[2], planes[2];
};
int    packPatches( CanvasPatch* dPatches, int numPatches, const uint8_t* occArena, int sizeU, int sizeV, int occRes, int* dResult, cudaStream_t s );
void   formOccupancyAndGeometry( const CanvasPatch* dPatches, int numPatches, int maxPatchPixels, int maxPatchBlocks, const int16_t* depthArena,
                                 int occRes, int prec, int W, int H, CanvasImages& im, cudaStream_t s );
void   blockToPatchFromVideo( const CanvasPatch* dPatches, int numPatches, int maxPatchBlocks, int occRes, int prec, int W, int H, const uint8_t* om,
                              uint32_t* blockToPatch, cudaStream_t s );
size_t reconstructPoints( const CanvasPatch* dPatches, const long long* dElemBase, int numPatches, long long totalElems, int occRes, int prec, int W,
                          int H, const uint8_t* om, const uint32_t* blockToPatch, const uint16_t* geo0, const uint16_t* geo1, ReconScratch& rc,
                          ReconTemp& tmp, cudaStream_t s );
void   formAttributeImages( const uint32_t* pointToPixel, const uchar4* recRgb, size_t R, const uint8_t* om, int W, int H, int prec, AttrImages& at,
                            AttrTemp& tmp, cudaStream_t s );

// pack_ra.cu: a15, random-access packing of a whole GOF (host logic in ra_pack.hpp, placement searches on the device)
struct RaPackScratch {
  DevBuf<ra::PlaceItem> items;
  DevBuf<ra::PlaceJob>  jobs;
  DevBuf<uint8_t>       occ;
};
bool packGofRandomAccess( std::vector<ra::Frame>& frames, int occRes, size_t minWidth, size_t minHeight, RaPackScratch& sc, Profiler* prof,
                          cudaStream_t s );

// color.cu
struct ColorScratch {
  KdTree           recTree;
  DevBuf<uint32_t> fwdIdx, bwdIdx, idsA, idsB;
  DevBuf<float>    fwdDist, bwdDist;
  DevBuf<uint64_t> keysA, keysB;
  DevBuf<uint8_t>  cubTmp;
};
void transferColors( ColorScratch& sc, const KdTree& srcTree, const short4* srcPts, const uchar4* srcRgb, size_t n, const short4* recPts, size_t R,
                     uchar4* recRgb, cudaStream_t s );

// Everything a frame needs only INSIDE one group of stages (before the orientation walk: sort buffers of the edge keys; after
// it: refinement grid, patch bitmaps, colour-transfer tree and vote buffers, push-pull pyramid ...): ~1.2 GB at 0.83 Mpts, five
// times what a frame has to keep while its walk runs. Frames lease one set from a per-device pool for the duration of a stage
// group (scratch.cu), so the number of frames in flight is bounded by what they KEEP, not by their peak.
struct FrameScratch {
  OrientTemp    orientTmp;
  RefineScratch refine;
  PatchScratch  patch;
  ColorScratch  color;
  ReconTemp     recon;
  AttrTemp      attr;
};
FrameScratch* acquireFrameScratch( int device );             // blocks while all sets of the device are leased
void          releaseFrameScratch( int device, FrameScratch* );
int           setFrameScratchSets( int device, int count );  // upper bound of sets (>= 1); returns the previous bound

// color.cu: RGB444 planes (uint16, 8-bit values) -> 8-bit YUV 4:2:0 (Y plane, then U, then V), W and H even
struct YuvScratch {
  DevBuf<float>   U, V, tU, tV;
  DevBuf<uint8_t> out[2];
};
void rgbPlanesToYuv420( const uint16_t* rgbPlanes, int W, int H, YuvScratch& sc, uint8_t* out, cudaStream_t s );

// util.cu
void projectedAreas( const short4* pts, size_t n, int bits, uint32_t* faces, unsigned* counts, cudaStream_t s );
void gatherU8( const uint8_t* src, const uint32_t* idx, size_t n, uint8_t* dst, cudaStream_t s );
void packXyz( const int16_t* xyz3, size_t n, short4* out, cudaStream_t s );  // device int16 AoS (n x 3) -> short4

}  // namespace pccb200
