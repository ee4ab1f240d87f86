// kdtree.cuh — device kd-tree with nanoflann-identical shape and traversal order.
// Replaces PCCKdTree (PccLibCommon/source/PCCKdTree.cpp:42-79) + nanoflann divideTree/searchLevel
// (dependencies/nanoflann/nanoflann.hpp:1041-1254).  The point order inside the leaves (vind) and the
// near-child-first visiting order decide which of several equidistant points a k-NN query keeps
// (KNNResultSet::addPoint, nanoflann.hpp:110-131), so both are reproduced exactly.
#pragma once
#include "common.cuh"

namespace pccb200 {

// 16-byte node record, one load per visit.
//   internal: x = id of child 0 (child 1 = x+1), y = split dimension (0..2), z = divlow, w = divhigh
//   leaf    : x = first position in tree order, y = (last+1) | kLeafBit, z = w = 0
static constexpr int kLeafBit     = int( 0x80000000u );
static constexpr int kLeafMaxSize = 10;  // PCCKdTree.cpp:58

struct KdTree {
  size_t            n = 0;
  DevBuf<short4>    pts;    // points in caller order (w unused)
  DevBuf<short4>    ptsT;   // points in tree order: ptsT[p] = pts[vind[p]]
  DevBuf<uint32_t>  vind;   // tree order -> caller index
  DevBuf<int4>      nodes;  // node 0 is the root
  int               rootBox[6] = {0, 0, 0, 0, 0, 0};  // tight lo[3], hi[3]
  int               numNodes   = 0;
  int               numLevels  = 0;  // depth of the tree (bounds the search stack)
  int               lastTopLevels = 0;  // level-synchronous levels the previous build needed (kdtree.cu)
  // build scratch (kept for reuse)
  DevBuf<uint32_t>  tmpA, tmpB, flags, scanOut, scanTmp;  // flags / scanOut: the scans of the two sweeps; scanTmp: look-back control blocks
  DevBuf<int>       slotOf, slotOfNext;                    // slotOfNext: records of the local subtrees
  DevBuf<int>       slotI[2];  // per-slot int records, double buffered (see kdtree.cu)
  DevBuf<int>       counters;
  PinBuf<int>       hostInts;  // page-locked landing zone of the one read-back per tree (a pageable copy would spin in the driver)
};

// xyz4: n points already on the device as short4. Builds nodes/vind/ptsT on stream s (one host synchronisation at the end).
void kdBuild( KdTree& t, const short4* xyz4, size_t n, cudaStream_t s );

// k-NN (k <= 16) of nq queries (short4, device). Outputs row-major nq x k, rows in query order:
//   outIdx  : caller indices of the neighbours, nanoflann result order; 0xFFFFFFFF padding if n < k
//   outDist : squared distances as float (may be null)
// If queryOrder != null, thread t answers query queryOrder[t] (use the tree order of a self-query for
// coherent traversals); rows are still written at the query's own index.
void kdKnn( const KdTree& t, const short4* queries, size_t nq, const uint32_t* queryOrder, int k, uint32_t* outIdx,
            float* outDist, cudaStream_t s );

}  // namespace pccb200
