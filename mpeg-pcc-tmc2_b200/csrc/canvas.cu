// canvas.cu — everything that happens on the 2-D atlas canvas:
//   a13  first-fit patch packing            PCCEncoder::packFlexible            (PCCEncoder.cpp:2306-2449)
//   a16  full-resolution occupancy map      PCCEncoder::generateOccupancyMap    (:3768-3784)
//   a17  occupancy video frame              PCCEncoder::generateOccupancyMapVideo (:806-861)
//   a18  block-to-patch map                 PCCCodec::generateBlockToPatchFromOccupancyMapVideo (PCCCodec.cpp:1736-1774)
//   a19  geometry images D0/D1              PCCEncoder::generateIntraImage      (:3929-3960)
//   a20  block dilation                     PCCEncoder::dilate3DPadding         (:5951-6130, geometryPadding = 0)
//   a21  group dilation                     PCCEncoder::dilateGroupGeometryVideo (:3717-3739)
//   a22  point reconstruction               PCCCodec::generatePointCloud        (PCCCodec.cpp:519-980) + identifyBoundaryPoints (:268-327)
//   a25  attribute images T0/T1             PCCEncoder::generateAttributeVideo(tile) (:6736-6819)
//   a26  push-pull background fill          PCCEncoder::dilateSmoothedPushPull  (:6542-6591)
// All integer work; every kernel is one pass over pixels / blocks / points (HBM-bound, no tensor cores).
#include <limits.h>

#include <algorithm>

#include <mutex>

#include "stages.cuh"

namespace pccb200 {

namespace {

constexpr int16_t kInfDepth = 32767;
enum { OR_DEFAULT = 0, OR_SWAP = 1 };  // PATCH_ORIENTATION_DEFAULT / _SWAP (PCCBitstreamCommon.h:112-122)

__constant__ int cAxes[6][4] = {{0, 2, 1, 0}, {1, 2, 0, 0}, {2, 0, 1, 0}, {0, 2, 1, 1}, {1, 2, 0, 1}, {2, 0, 1, 1}};

__device__ __forceinline__ void pixelToCanvas( const CanvasPatch& m, int occRes, int u, int v, int& x, int& y ) {
  if ( m.orientation == OR_DEFAULT ) {
    x = u + m.u0 * occRes, y = v + m.v0 * occRes;
  } else {
    x = v + m.u0 * occRes, y = u + m.v0 * occRes;
  }
}
__device__ __forceinline__ void blockToCanvas( const CanvasPatch& m, int ub, int vb, int& x, int& y ) {
  if ( m.orientation == OR_DEFAULT ) {
    x = ub + m.u0, y = vb + m.v0;
  } else {
    x = vb + m.u0, y = ub + m.v0;
  }
}

// ------------------------------------------------------------------------------------------------ a13
// One CTA per frame. The canvas of 16-px blocks is a bit matrix in shared memory (4 x 32-bit words per row:
// up to 128 blocks = 2048 px wide). Patches are placed one after the other (the reference's order-dependent
// first fit); for each patch all threads test candidate positions in raster order and the first fit wins.
constexpr int kPackWordsPerRow = 6;  // 192 blocks: 2560-px canvases (vox11) plus slack
constexpr int kPackMaxRows     = 1024;

__device__ __forceinline__ bool rowFree( const uint32_t* row, int x0, int w ) {
  // bits [x0, x0+w) of the row must be zero
  int x = x0, left = w;
  while ( left > 0 ) {
    const int      word = x >> 5, bit = x & 31, take = min( left, 32 - bit );
    const uint32_t mask = ( take == 32 ? 0xFFFFFFFFu : ( ( 1u << take ) - 1u ) ) << bit;
    if ( row[word] & mask ) return false;
    x += take, left -= take;
  }
  return true;
}

__global__ void __launch_bounds__( 512, 1 )
    kPack( CanvasPatch* __restrict__ patches, int numPatches, const uint8_t* __restrict__ occArena, int sizeU, int sizeV0, int occRes,
           int* __restrict__ result /* [0]=height in px, [1]=error */ ) {
  extern __shared__ uint32_t canvas[];  // kPackMaxRows x kPackWordsPerRow
  __shared__ int             best;
  int                        sizeV = sizeV0;
  for ( int i = threadIdx.x; i < kPackMaxRows * kPackWordsPerRow; i += blockDim.x ) canvas[i] = 0;
  __syncthreads();
  int height = sizeV * occRes;
  for ( int pi = 0; pi < numPatches; ++pi ) {
    const CanvasPatch m  = patches[pi];
    const int         wU = m.sizeU0, hV = m.sizeV0;
    int               found = -1;
    while ( found < 0 ) {
      // candidate index = ( v * sizeU + u ) * 2 + o ; orientation order depends on the aspect (PCCEncoder.cpp:2396-2402)
      const int total = sizeV * sizeU * 2;
      for ( int base = 0; base < total && found < 0; base += blockDim.x ) {
        if ( threadIdx.x == 0 ) best = INT_MAX;
        __syncthreads();
        const int c = base + threadIdx.x;
        if ( c < total ) {
          const int o = c & 1, pos = c >> 1, u = pos % sizeU, v = pos / sizeU;
          const int orient = ( wU > hV ) ? ( o == 0 ? OR_SWAP : OR_DEFAULT ) : ( o == 0 ? OR_DEFAULT : OR_SWAP );
          const int bw = orient == OR_DEFAULT ? wU : hV, bh = orient == OR_DEFAULT ? hV : wU;  // bounding box on the canvas
          bool      fits = ( u + bw <= sizeU ) && ( v + bh <= sizeV );
          for ( int r = 0; r < bh && fits; ++r ) fits = rowFree( canvas + ( v + r ) * kPackWordsPerRow, u, bw );
          if ( fits ) atomicMin( &best, c );
        }
        __syncthreads();
        if ( best != INT_MAX ) found = best;
        __syncthreads();
      }
      if ( found < 0 ) {
        sizeV *= 2;  // new rows are already zero
        if ( sizeV > kPackMaxRows ) {
          if ( threadIdx.x == 0 ) result[1] = 1;
          return;
        }
      }
    }
    const int o = found & 1, pos = found >> 1, u0 = pos % sizeU, v0 = pos / sizeU;
    const int orient = ( wU > hV ) ? ( o == 0 ? OR_SWAP : OR_DEFAULT ) : ( o == 0 ? OR_DEFAULT : OR_SWAP );
    // take the occupied blocks only
    for ( int b = threadIdx.x; b < wU * hV; b += blockDim.x ) {
      if ( !occArena[m.occOff + b] ) continue;
      const int ub = b % wU, vb = b / wU;
      const int x = orient == OR_DEFAULT ? ub + u0 : vb + u0, y = orient == OR_DEFAULT ? vb + v0 : ub + v0;
      atomicOr( &canvas[y * kPackWordsPerRow + ( x >> 5 )], 1u << ( x & 31 ) );
    }
    if ( threadIdx.x == 0 ) {
      patches[pi].u0 = u0, patches[pi].v0 = v0, patches[pi].orientation = orient;
    }
    height = max( height, ( v0 + ( orient == OR_DEFAULT ? hV : wU ) ) * occRes );
    __syncthreads();
  }
  if ( threadIdx.x == 0 ) result[0] = height;
}

// ------------------------------------------------------------------------------------------- a16 + a19
// grid.y = patch. Writes occupancy, D0 and D1 of every occupied patch pixel to the canvas.
__global__ void kScatterPatches( const CanvasPatch* __restrict__ patches, const int16_t* __restrict__ depthArena, int occRes, int W, int H,
                                 uint8_t* __restrict__ occ, uint16_t* __restrict__ geo0, uint16_t* __restrict__ geo1, int* __restrict__ error ) {
  const CanvasPatch m  = patches[blockIdx.y];
  const int         px = m.sizeU * m.sizeV;
  for ( int p = blockIdx.x * blockDim.x + threadIdx.x; p < px; p += gridDim.x * blockDim.x ) {
    const int16_t d0 = depthArena[m.depthOff + p];
    if ( !( d0 < kInfDepth ) ) continue;
    int x, y;
    pixelToCanvas( m, occRes, p % m.sizeU, p / m.sizeU, x, y );
    if ( x >= W || y >= H ) {  // PCCPatch::patch2Canvas would exit(180)
      *error = PCCB200_ERR_CANVAS;
      continue;
    }
    const size_t q = size_t( y ) * W + x;
    occ[q]         = 1;
    geo0[q]        = uint16_t( d0 );
    geo1[q]        = uint16_t( depthArena[m.depthOff + px + p] );
  }
}

// ------------------------------------------------------------------------------------------------ a17
__global__ void kOccupancyVideo( const uint8_t* __restrict__ occ, int W, int H, int prec, uint8_t* __restrict__ om ) {
  const int ow = W / prec, oh = H / prec;
  const int c  = blockIdx.x * blockDim.x + threadIdx.x;
  if ( c >= ow * oh ) return;
  const int cx = c % ow, cy = c / ow;
  uint8_t   any = 0;
  for ( int j = 0; j < prec; ++j )
    for ( int i = 0; i < prec; ++i ) any |= occ[size_t( cy * prec + j ) * W + cx * prec + i];
  om[c] = any ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ a18
// grid.y = patch; one thread per patch block: the highest patch index whose block sees an active cell wins.
__global__ void kBlockToPatch( const CanvasPatch* __restrict__ patches, const uint8_t* __restrict__ om, int occRes, int prec, int W, int H,
                               uint32_t* __restrict__ blockToPatch ) {
  const CanvasPatch m = patches[blockIdx.y];
  const int         b = blockIdx.x * blockDim.x + threadIdx.x;
  if ( b >= m.sizeU0 * m.sizeV0 ) return;
  const int ub = b % m.sizeU0, vb = b / m.sizeU0;
  int       bx, by;
  blockToCanvas( m, ub, vb, bx, by );
  const int bw = W / occRes, ow = W / prec, cells = occRes / prec;
  if ( bx < 0 || by < 0 || bx >= bw || by >= H / occRes ) return;
  bool any = false;
  for ( int j = 0; j < cells && !any; ++j )
    for ( int i = 0; i < cells && !any; ++i ) any = om[size_t( by * cells + j ) * ow + bx * cells + i] != 0;
  if ( any ) atomicMax( &blockToPatch[size_t( by ) * bw + bx], uint32_t( blockIdx.y ) + 1u );
}

// ------------------------------------------------------------------------------------------------ a20
// One CTA (occRes x occRes threads, occRes = 16) per canvas block. Partly filled blocks grow by 4-neighbour
// averaging waves inside the block; empty blocks are only flagged here and resolved by kFillEmptyBlocks.
__global__ void __launch_bounds__( 256 )
    kDilateBlocks( uint16_t* __restrict__ img, const uint8_t* __restrict__ occ, int W, uint8_t* __restrict__ blockEmpty ) {
  __shared__ int      val[16][16];
  __shared__ int      state[16][16];  // 0 = empty, k = filled in wave k
  __shared__ int      filled;
  const int           u = threadIdx.x & 15, v = threadIdx.x >> 4;
  const int           bw = W / 16;
  const size_t        q = size_t( blockIdx.y * 16 + v ) * W + blockIdx.x * 16 + u;
  if ( threadIdx.x == 0 ) filled = 0;
  __syncthreads();
  const int s0 = occ[q] ? 1 : 0;
  state[v][u]  = s0;
  val[v][u]    = img[q];
  if ( s0 ) atomicAdd( &filled, 1 );
  __syncthreads();
  if ( filled == 0 ) {
    if ( threadIdx.x == 0 ) blockEmpty[blockIdx.y * bw + blockIdx.x] = 1;
    return;
  }
  if ( threadIdx.x == 0 ) blockEmpty[blockIdx.y * bw + blockIdx.x] = 0;
  int iteration = 1;
  while ( filled < 256 ) {
    int sum = 0, cnt = 0;
    if ( state[v][u] == 0 ) {
      if ( v > 0 && state[v - 1][u] == iteration ) sum += val[v - 1][u], ++cnt;
      if ( u > 0 && state[v][u - 1] == iteration ) sum += val[v][u - 1], ++cnt;
      if ( u < 15 && state[v][u + 1] == iteration ) sum += val[v][u + 1], ++cnt;
      if ( v < 15 && state[v + 1][u] == iteration ) sum += val[v + 1][u], ++cnt;
    }
    __syncthreads();
    if ( cnt ) {
      val[v][u]   = ( sum + cnt / 2 ) / cnt;
      state[v][u] = iteration + 1;
      atomicAdd( &filled, 1 );
    }
    __syncthreads();
    ++iteration;
  }
  img[q] = uint16_t( val[v][u] );
}

// Empty blocks copy, in the reference's raster order, the border pixels of the block to their left (first column of
// blocks: of the block above). Unrolled: the source is the nearest non-empty block to the left in the block row, else the
// nearest non-empty block above in block column 0, else nothing (zeros).
__global__ void __launch_bounds__( 256 )
    kFillEmptyBlocks( uint16_t* __restrict__ img, const uint8_t* __restrict__ blockEmpty, int W ) {
  const int bw = W / 16, ub = blockIdx.x, vb = blockIdx.y;
  if ( !blockEmpty[vb * bw + ub] ) return;
  const int u = threadIdx.x & 15, v = threadIdx.x >> 4;
  const int x = ub * 16 + u, y = vb * 16 + v;
  int       src = ub - 1;
  while ( src >= 0 && blockEmpty[vb * bw + src] ) --src;
  uint16_t value = 0;
  if ( src >= 0 ) {
    value = img[size_t( y ) * W + src * 16 + 15];
  } else {
    int up = vb - 1;
    while ( up >= 0 && blockEmpty[up * bw + 0] ) --up;
    if ( up >= 0 ) value = img[size_t( up * 16 + 15 ) * W + ( ub == 0 ? x : 15 )];
  }
  img[size_t( y ) * W + x] = value;
}

// ------------------------------------------------------------------------------------------------ a21
__global__ void kGroupDilate( uint16_t* __restrict__ geo0, uint16_t* __restrict__ geo1, const uint8_t* __restrict__ om, int W, int H, int prec ) {
  const size_t q = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( q >= size_t( W ) * H ) return;
  const int x = int( q % W ), y = int( q / W );
  if ( om[size_t( y / prec ) * ( W / prec ) + x / prec ] ) return;
  const uint32_t avg = ( uint32_t( geo0[q] ) + uint32_t( geo1[q] ) + 1u ) >> 1;
  geo0[q] = geo1[q] = uint16_t( avg );
}

// ------------------------------------------------------------------------------------------------ a22
// Element e of patch p = ( local block , pixel in block ), in the reference's emission order (patch, block row-major,
// pixel row-major). Pass 1 counts the points an element emits (0, 1 or 2), a device-wide scan gives positions, pass 2 writes.
__device__ __forceinline__ int reconstructElement( const CanvasPatch& m, int patchIndex, long long local, int occRes, int prec, int W, int H,
                                                   const uint8_t* om, const uint32_t* blockToPatch, const uint16_t* geo0, const uint16_t* geo1,
                                                   short4 out[2], int& x, int& y ) {
  const int pixPerBlock = occRes * occRes;
  const int blk = int( local / pixPerBlock ), pix = int( local % pixPerBlock );
  const int ub = blk % m.sizeU0, vb = blk / m.sizeU0;
  int       bx, by;
  blockToCanvas( m, ub, vb, bx, by );
  if ( bx < 0 || by < 0 || bx >= W / occRes || by >= H / occRes ) return 0;
  if ( blockToPatch[size_t( by ) * ( W / occRes ) + bx] != uint32_t( patchIndex ) + 1u ) return 0;
  const int u = ub * occRes + pix % occRes, v = vb * occRes + pix / occRes;
  pixelToCanvas( m, occRes, u, v, x, y );
  if ( !om[size_t( y / prec ) * ( W / prec ) + x / prec] ) return 0;
  const int na = cAxes[m.viewId][0], ta = cAxes[m.viewId][1], ba = cAxes[m.viewId][2], mode = cAxes[m.viewId][3];
  int       n  = 0;
#pragma unroll
  for ( int map = 0; map < 2; ++map ) {
    const int depth = map == 0 ? geo0[size_t( y ) * W + x] : geo1[size_t( y ) * W + x];
    int       nc;
    if ( mode == 0 ) {
      nc = depth + m.d1;
    } else {
      nc = m.d1 - depth;
      if ( nc < 0 ) nc = 0;
    }
    short c[3];
    c[na] = short( nc ), c[ta] = short( u + m.u1 ), c[ba] = short( v + m.v1 );
    const short4 p = make_short4( c[0], c[1], c[2], 0 );
    if ( map == 1 && p.x == out[0].x && p.y == out[0].y && p.z == out[0].z ) break;  // removeDuplicatePoints
    out[n++] = p;
  }
  return n;
}

__global__ void kReconstructCount( const CanvasPatch* __restrict__ patches, const long long* __restrict__ elemBase, int numPatches,
                                   long long totalElems, int occRes, int prec, int W, int H, const uint8_t* __restrict__ om,
                                   const uint32_t* __restrict__ blockToPatch, const uint16_t* __restrict__ geo0, const uint16_t* __restrict__ geo1,
                                   uint32_t* __restrict__ counts ) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ( e >= totalElems ) return;
  int lo = 0, hi = numPatches - 1;  // last patch whose base <= e
  while ( lo < hi ) {
    const int mid = ( lo + hi + 1 ) >> 1;
    if ( elemBase[mid] <= e ) lo = mid;
    else hi = mid - 1;
  }
  short4 pts[2];
  int    x, y;
  counts[e] = reconstructElement( patches[lo], lo, e - elemBase[lo], occRes, prec, W, H, om, blockToPatch, geo0, geo1, pts, x, y );
}

__global__ void kReconstructEmit( const CanvasPatch* __restrict__ patches, const long long* __restrict__ elemBase, int numPatches,
                                  long long totalElems, int occRes, int prec, int W, int H, const uint8_t* __restrict__ om,
                                  const uint32_t* __restrict__ blockToPatch, const uint16_t* __restrict__ geo0, const uint16_t* __restrict__ geo1,
                                  const uint32_t* __restrict__ offsets, short4* __restrict__ recXyz, uint32_t* __restrict__ pointToPixel,
                                  uint32_t* __restrict__ recPartition ) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ( e >= totalElems ) return;
  if ( offsets[e + 1] == offsets[e] ) return;
  int lo = 0, hi = numPatches - 1;
  while ( lo < hi ) {
    const int mid = ( lo + hi + 1 ) >> 1;
    if ( elemBase[mid] <= e ) lo = mid;
    else hi = mid - 1;
  }
  short4    pts[2];
  int       x, y;
  const int n = reconstructElement( patches[lo], lo, e - elemBase[lo], occRes, prec, W, H, om, blockToPatch, geo0, geo1, pts, x, y );
  for ( int k = 0; k < n; ++k ) {
    const size_t o          = offsets[e] + k;
    recXyz[o]               = pts[k];
    pointToPixel[3 * o]     = x;
    pointToPixel[3 * o + 1] = y;
    pointToPixel[3 * o + 2] = k;
    recPartition[o]         = lo;
  }
}

__global__ void kBoundary( const uint32_t* __restrict__ pointToPixel, int R, const uint8_t* __restrict__ om, int W, int H, int prec,
                           uint16_t* __restrict__ boundary ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= R ) return;
  const int x = pointToPixel[3 * size_t( i )], y = pointToPixel[3 * size_t( i ) + 1];
  const int ow = W / prec;
  auto      at = [&]( int xx, int yy ) { return om[size_t( yy / prec ) * ow + xx / prec] != 0; };
  bool      b  = false;
  if ( at( x, y ) ) {
    const bool yIn = y > 0 && y < H - 1, xIn = x > 0 && x < W - 1;
    if ( yIn && ( !at( x, y - 1 ) || !at( x, y + 1 ) ) ) b = true;
    if ( !b && xIn && ( !at( x + 1, y ) || !at( x - 1, y ) ) ) b = true;
    if ( !b && yIn && x > 0 && ( !at( x - 1, y - 1 ) || !at( x - 1, y + 1 ) ) ) b = true;
    if ( !b && yIn && x < W - 1 && ( !at( x + 1, y - 1 ) || !at( x + 1, y + 1 ) ) ) b = true;
    if ( y == 0 || y == H - 1 || x == 0 || x == W - 1 ) b = true;
    if ( !b ) {
      for ( int ix = -2; ix <= 2 && !b; ++ix )
        for ( int iy = -2; iy <= 2 && !b; ++iy )
          if ( ( abs( ix ) > 1 || abs( iy ) > 1 ) && y + iy >= 0 && y + iy < H && x + ix >= 0 && x + ix < W && !at( x + ix, y + iy ) ) b = true;
      if ( y == 1 || y == H - 2 || x == 1 || x == W - 2 ) b = true;
    }
  }
  boundary[i] = b ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ a25
// Canvas pixels hold colours as ushort4 (r,g,b,flag). Pass 1: every point writes its own map, D1 points set the flag.
// Pass 2: D0 points whose pixel carries no D1 point copy their colour to T1.
__global__ void kAttrScatter( const uint32_t* __restrict__ pointToPixel, const uchar4* __restrict__ recRgb, int R, int W,
                              ushort4* __restrict__ T0, ushort4* __restrict__ T1 ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= R ) return;
  const size_t q = size_t( pointToPixel[3 * size_t( i ) + 1] ) * W + pointToPixel[3 * size_t( i )];
  const uchar4 c = recRgb[i];
  if ( pointToPixel[3 * size_t( i ) + 2] == 0 )
    T0[q] = make_ushort4( c.x, c.y, c.z, 0 );
  else
    T1[q] = make_ushort4( c.x, c.y, c.z, 1 );
}
__global__ void kAttrFallback( const uint32_t* __restrict__ pointToPixel, const uchar4* __restrict__ recRgb, int R, int W,
                               ushort4* __restrict__ T1 ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= R || pointToPixel[3 * size_t( i ) + 2] != 0 ) return;
  const size_t q = size_t( pointToPixel[3 * size_t( i ) + 1] ) * W + pointToPixel[3 * size_t( i )];
  if ( T1[q].w ) return;
  const uchar4 c = recRgb[i];
  T1[q]          = make_ushort4( c.x, c.y, c.z, 0 );
}

__global__ void kUpsampleOccupancy( const uint8_t* __restrict__ om, int W, int H, int prec, uint8_t* __restrict__ occ ) {
  const size_t q = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( q >= size_t( W ) * H ) return;
  occ[q] = om[size_t( ( q / W ) / prec ) * ( W / prec ) + ( q % W ) / prec];
}

// interleaved (r,g,b,*) -> three planes of uint16 (PCCImage<uint16_t,3>::channels_)
__global__ void kToPlanes( const ushort4* __restrict__ img, size_t n, uint16_t* __restrict__ planes ) {
  const size_t q = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( q >= n ) return;
  const ushort4 c   = img[q];
  planes[q]         = c.x;
  planes[n + q]     = c.y;
  planes[2 * n + q] = c.z;
}

// ------------------------------------------------------------------------------------------------ a26
__device__ __forceinline__ int mean4w( int p1, int w1, int p2, int w2, int p3, int w3, int p4, int w4 ) {
  return ( p1 * w1 + p2 * w2 + p3 * w3 + p4 * w4 ) / ( w1 + w2 + w3 + w4 );
}

// Both attribute frames (T0, T1) go through the pyramid in the SAME launches (blockIdx.y selects the frame; the occupancy pyramid
// is common to both), and the small levels (at most kTailMaxPixels pixels) are done by ONE CTA per frame in shared memory
// (kPushPullTail): pulls down to the coarsest level, pushes and all smoothing passes back up.
struct Pair {
  ushort4* p[2];
};
struct ConstPair {
  const ushort4* p[2];
};

// pull: occupancy-weighted 2x2 mean (values truncated to 8 bits as the reference's unsigned char locals do)
__device__ __forceinline__ void pullPixel( const ushort4* __restrict__ img, const uint8_t* __restrict__ occ, int W, int H, int x, int y, ushort4& out,
                                           uint8_t& o ) {
  const int  X = x << 1, Y = y << 1;
  const bool in2 = X + 1 < W, in3 = Y + 1 < H;
  const int  w1 = occ[size_t( Y ) * W + X] ? 255 : 0, w2 = ( in2 && occ[size_t( Y ) * W + X + 1] ) ? 255 : 0,
            w3 = ( in3 && occ[size_t( Y + 1 ) * W + X] ) ? 255 : 0, w4 = ( in2 && in3 && occ[size_t( Y + 1 ) * W + X + 1] ) ? 255 : 0;
  out = make_ushort4( 0, 0, 0, 0 );
  o   = 0;
  if ( w1 + w2 + w3 + w4 > 0 ) {
    const ushort4 z  = make_ushort4( 0, 0, 0, 0 );
    const ushort4 a  = img[size_t( Y ) * W + X], b = in2 ? img[size_t( Y ) * W + X + 1] : z, c = in3 ? img[size_t( Y + 1 ) * W + X] : z,
                  d  = ( in2 && in3 ) ? img[size_t( Y + 1 ) * W + X + 1] : z;
    out.x            = mean4w( a.x & 0xff, w1, b.x & 0xff, w2, c.x & 0xff, w3, d.x & 0xff, w4 );
    out.y            = mean4w( a.y & 0xff, w1, b.y & 0xff, w2, c.y & 0xff, w3, d.y & 0xff, w4 );
    out.z            = mean4w( a.z & 0xff, w1, b.z & 0xff, w2, c.z & 0xff, w3, d.z & 0xff, w4 );
    o                = 1;
  }
}
__global__ void kPull( ConstPair img, const uint8_t* __restrict__ occ, int W, int H, Pair mip, uint8_t* __restrict__ mipOcc, int nw, int nh ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= nw * nh ) return;
  ushort4 out;
  uint8_t o;
  pullPixel( img.p[blockIdx.y], occ, W, H, i % nw, i / nw, out, o );
  mip.p[blockIdx.y][i] = out;
  if ( blockIdx.y == 0 ) mipOcc[i] = o;
}

// push: unoccupied pixels take the bilinear-like blend (144,48,48,16) of the coarser level
__device__ __forceinline__ ushort4 pushPixel( const ushort4* __restrict__ mip, int w, int h, int X, int Y, unsigned short keepW ) {
  const int  x = X >> 1, y = Y >> 1;
  const int  dx = ( X & 1 ) ? 1 : -1, dy = ( Y & 1 ) ? 1 : -1;
  const bool okx = dx < 0 ? x > 0 : x < w - 1, oky = dy < 0 ? y > 0 : y < h - 1;
  const ushort4 z = make_ushort4( 0, 0, 0, 0 );
  const ushort4 v = mip[size_t( y ) * w + x], vx = okx ? mip[size_t( y ) * w + x + dx] : z, vy = oky ? mip[size_t( y + dy ) * w + x] : z,
                vd = ( okx && oky ) ? mip[size_t( y + dy ) * w + x + dx] : z;
  const int wx = okx ? 48 : 0, wy = oky ? 48 : 0, wd = ( okx && oky ) ? 16 : 0;
  return make_ushort4( mean4w( v.x, 144, vx.x, wx, vy.x, wy, vd.x, wd ), mean4w( v.y, 144, vx.y, wx, vy.y, wy, vd.y, wd ),
                       mean4w( v.z, 144, vx.z, wx, vy.z, wy, vd.z, wd ), keepW );
}
__global__ void kPushFill( Pair img, const uint8_t* __restrict__ occ, int W, int H, ConstPair mip, int w, int h ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= W * H || occ[i] ) return;
  ushort4* im = img.p[blockIdx.y];
  im[i]       = pushPixel( mip.p[blockIdx.y], w, h, i % W, i / W, im[i].w );
}

// one Jacobi pass of the 8-neighbour smoothing of unoccupied pixels (occupied ones are copied through)
__device__ __forceinline__ ushort4 smoothPixel( const ushort4* __restrict__ src, const uint8_t* __restrict__ occ, int W, int H, int i ) {
  if ( occ[i] ) return src[i];
  const int x = i % W, y = i / W;
  const int x1 = x > 0 ? x - 1 : x, y1 = y > 0 ? y - 1 : y, x2 = x < W - 1 ? x + 1 : x, y2 = y < H - 1 ? y + 1 : y;
  auto      at = [&]( int xx, int yy ) { return src[size_t( yy ) * W + xx]; };
  const ushort4 a = at( x1, y1 ), b = at( x2, y1 ), c = at( x1, y2 ), d = at( x2, y2 ), e = at( x1, y ), f = at( x2, y ), g = at( x, y1 ), h = at( x, y2 );
  return make_ushort4( ( a.x + b.x + c.x + d.x + e.x + f.x + g.x + h.x + 4 ) >> 3, ( a.y + b.y + c.y + d.y + e.y + f.y + g.y + h.y + 4 ) >> 3,
                       ( a.z + b.z + c.z + d.z + e.z + f.z + g.z + h.z + 4 ) >> 3, src[i].w );
}
__global__ void kSmooth8( ConstPair src, Pair dst, const uint8_t* __restrict__ occ, int W, int H ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= W * H ) return;
  dst.p[blockIdx.y][i] = smoothPixel( src.p[blockIdx.y], occ, W, H, i );
}

// group dilation of the two attribute maps (PCCEncoder.cpp:391-413): unoccupied pixels take the rounded mean of T0 and T1
__global__ void kAttrGroupDilate( ushort4* __restrict__ t0, ushort4* __restrict__ t1, const uint8_t* __restrict__ occ, size_t n ) {
  const size_t q = size_t( blockIdx.x ) * blockDim.x + threadIdx.x;
  if ( q >= n || occ[q] ) return;
  const ushort4 a = t0[q], b = t1[q];
  const ushort4 m = make_ushort4( ( ( a.x & 0xff ) + ( b.x & 0xff ) + 1 ) >> 1, ( ( a.y & 0xff ) + ( b.y & 0xff ) + 1 ) >> 1,
                                  ( ( a.z & 0xff ) + ( b.z & 0xff ) + 1 ) >> 1, 0 );
  t0[q] = make_ushort4( m.x, m.y, m.z, a.w ), t1[q] = make_ushort4( m.x, m.y, m.z, b.w );
}

// the small levels of the pyramid in one CTA per frame: level 0 of `a` is the finest level handled here (in global memory:
// base / baseOcc), the coarser ones live in shared memory only
constexpr int kTailMaxPixels = 6400;  // 80 x 80
constexpr int kTailLevels    = 12;
constexpr int kTailThreads   = 512;
struct TailArgs {
  int n;  // levels incl. the base level
  int w[kTailLevels], h[kTailLevels];
  int firstIters;  // smoothing passes after the coarsest push (grows by one per level, at most 16)
};
__global__ void __launch_bounds__( kTailThreads ) kPushPullTail( Pair base, const uint8_t* __restrict__ baseOcc, TailArgs a ) {
  extern __shared__ __align__( 16 ) unsigned char tailSmem[];
  ushort4* img[kTailLevels];
  uint8_t* occ[kTailLevels];
  size_t   px = 0;
  for ( int l = 0; l < a.n; ++l ) px += size_t( a.w[l] ) * a.h[l];
  {
    ushort4* ip = reinterpret_cast<ushort4*>( tailSmem );
    uint8_t* op = reinterpret_cast<uint8_t*>( ip + px + size_t( a.w[0] ) * a.h[0] );  // after the images and the ping-pong buffer
    for ( int l = 0; l < kTailLevels; ++l ) {
      img[l] = ip, occ[l] = op;
      if ( l < a.n ) ip += size_t( a.w[l] ) * a.h[l], op += size_t( a.w[l] ) * a.h[l];
    }
  }
  ushort4* const tmp = reinterpret_cast<ushort4*>( tailSmem ) + px;
  ushort4* const g   = base.p[blockIdx.x];
  const int      n0  = a.w[0] * a.h[0];
  for ( int i = threadIdx.x; i < n0; i += kTailThreads ) img[0][i] = g[i], occ[0][i] = baseOcc[i];
  __syncthreads();
  for ( int l = 1; l < a.n; ++l ) {
    const int nw = a.w[l], nh = a.h[l];
    for ( int i = threadIdx.x; i < nw * nh; i += kTailThreads ) pullPixel( img[l - 1], occ[l - 1], a.w[l - 1], a.h[l - 1], i % nw, i / nw, img[l][i], occ[l][i] );
    __syncthreads();
  }
  int iters = a.firstIters;
  for ( int l = a.n - 1; l >= 1; --l ) {
    const int dw = a.w[l - 1], dh = a.h[l - 1], n = dw * dh;
    for ( int i = threadIdx.x; i < n; i += kTailThreads )
      if ( !occ[l - 1][i] ) img[l - 1][i] = pushPixel( img[l], a.w[l], a.h[l], i % dw, i / dw, img[l - 1][i].w );
    __syncthreads();
    ushort4 *src = img[l - 1], *dst = tmp;
    for ( int it = 0; it < iters; ++it ) {
      for ( int i = threadIdx.x; i < n; i += kTailThreads ) dst[i] = smoothPixel( src, occ[l - 1], dw, dh, i );
      __syncthreads();
      ushort4* t = src;
      src = dst, dst = t;
    }
    if ( src != img[l - 1] ) {
      for ( int i = threadIdx.x; i < n; i += kTailThreads ) img[l - 1][i] = src[i];
      __syncthreads();
    }
    iters = min( iters + 1, 16 );
  }
  for ( int i = threadIdx.x; i < n0; i += kTailThreads ) g[i] = img[0][i];
}

}  // namespace

// ======================================================================================================
int packPatches( CanvasPatch* dPatches, int numPatches, const uint8_t* occArena, int sizeU, int sizeV, int occRes, int* dResult,
                 cudaStream_t s ) {
  if ( sizeU > kPackWordsPerRow * 32 ) return PCCB200_ERR_UNSUPPORTED;
  const size_t smem = size_t( kPackMaxRows ) * kPackWordsPerRow * sizeof( uint32_t );
  PCC_CUDA( cudaFuncSetAttribute( kPack, cudaFuncAttributeMaxDynamicSharedMemorySize, int( smem ) ) );
  PCC_CUDA( cudaMemsetAsync( dResult, 0, 2 * sizeof( int ), s ) );
  kPack<<<1, 512, smem, s>>>( dPatches, numPatches, occArena, sizeU, sizeV, occRes, dResult );
  PCC_LAUNCH_CHECK();
  return PCCB200_OK;
}

void formOccupancyAndGeometry( const CanvasPatch* dPatches, int numPatches, int maxPatchPixels, int maxPatchBlocks, const int16_t* depthArena,
                               int occRes, int prec, int W, int H, CanvasImages& im, cudaStream_t s ) {
  const size_t Q = size_t( W ) * H, cells = size_t( W / prec ) * ( H / prec ), blocks = size_t( W / occRes ) * ( H / occRes );
  im.occ.reserve( Q ), im.geo0.reserve( Q ), im.geo1.reserve( Q ), im.om.reserve( cells ), im.blockToPatch.reserve( blocks );
  im.blockEmpty.reserve( blocks ), im.error.reserve( 2 );
  PCC_CUDA( cudaMemsetAsync( im.occ, 0, Q, s ) );
  PCC_CUDA( cudaMemsetAsync( im.geo0, 0, Q * 2, s ) );
  PCC_CUDA( cudaMemsetAsync( im.geo1, 0, Q * 2, s ) );
  PCC_CUDA( cudaMemsetAsync( im.blockToPatch, 0, blocks * 4, s ) );
  PCC_CUDA( cudaMemsetAsync( im.error, 0, sizeof( int ), s ) );
  if ( numPatches ) {
    const dim3 g( std::max( 1, std::min( divUp( maxPatchPixels, 256 ), 64 ) ), numPatches );
    kScatterPatches<<<g, 256, 0, s>>>( dPatches, depthArena, occRes, W, H, im.occ, im.geo0, im.geo1, im.error );
  }
  kOccupancyVideo<<<divUp( cells, 256 ), 256, 0, s>>>( im.occ, W, H, prec, im.om );
  if ( numPatches ) {
    const dim3 g( std::max( 1, divUp( maxPatchBlocks, 128 ) ), numPatches );
    kBlockToPatch<<<g, 128, 0, s>>>( dPatches, im.om, occRes, prec, W, H, im.blockToPatch );
  }
  if ( occRes != 16 ) throw CudaError{ cudaErrorInvalidValue, __FILE__, __LINE__ };
  const dim3 bg( W / 16, H / 16 );
  for ( int map = 0; map < 2; ++map ) {
    uint16_t* img = map == 0 ? im.geo0.p : im.geo1.p;
    kDilateBlocks<<<bg, 256, 0, s>>>( img, im.occ, W, im.blockEmpty );
    kFillEmptyBlocks<<<bg, 256, 0, s>>>( img, im.blockEmpty, W );
  }
  kGroupDilate<<<divUp( Q, 256 ), 256, 0, s>>>( im.geo0, im.geo1, im.om, W, H, prec );
  PCC_LAUNCH_CHECK();
}

// a18 alone: block-to-patch from a (decoded) occupancy video
void blockToPatchFromVideo( const CanvasPatch* dPatches, int numPatches, int maxPatchBlocks, int occRes, int prec, int W, int H, const uint8_t* om,
                            uint32_t* blockToPatch, cudaStream_t s ) {
  const size_t blocks = size_t( W / occRes ) * ( H / occRes );
  PCC_CUDA( cudaMemsetAsync( blockToPatch, 0, blocks * 4, s ) );
  if ( numPatches ) {
    const dim3 g( std::max( 1, divUp( maxPatchBlocks, 128 ) ), numPatches );
    kBlockToPatch<<<g, 128, 0, s>>>( dPatches, om, occRes, prec, W, H, blockToPatch );
    PCC_LAUNCH_CHECK();
  }
}

size_t reconstructPoints( const CanvasPatch* dPatches, const long long* dElemBase, int numPatches, long long totalElems, int occRes, int prec, int W,
                          int H, const uint8_t* om, const uint32_t* blockToPatch, const uint16_t* geo0, const uint16_t* geo1, ReconScratch& rc,
                          ReconTemp& tmp, cudaStream_t s ) {
  rc.numPoints = 0;
  if ( totalElems == 0 || numPatches == 0 ) return 0;
  tmp.counts.reserve( totalElems + 1 ), tmp.offsets.reserve( totalElems + 2 ), tmp.scanTmp.reserve( scanTmpElems( totalElems ) );
  kReconstructCount<<<divUp( totalElems, 256 ), 256, 0, s>>>( dPatches, dElemBase, numPatches, totalElems, occRes, prec, W, H, om, blockToPatch, geo0,
                                                              geo1, tmp.counts );
  exclusiveScanU32( tmp.counts, tmp.offsets, totalElems, tmp.scanTmp, s );
  uint32_t R = 0;
  PCC_CUDA( cudaMemcpyAsync( &R, tmp.offsets.p + totalElems, sizeof( uint32_t ), cudaMemcpyDeviceToHost, s ) );
  streamWait( s );
  rc.numPoints = R;
  if ( R == 0 ) return 0;
  rc.recXyz.reserve( R ), rc.pointToPixel.reserve( 3 * size_t( R ) ), rc.recPartition.reserve( R ), rc.boundary.reserve( R );
  kReconstructEmit<<<divUp( totalElems, 256 ), 256, 0, s>>>( dPatches, dElemBase, numPatches, totalElems, occRes, prec, W, H, om, blockToPatch, geo0,
                                                             geo1, tmp.offsets, rc.recXyz, rc.pointToPixel, rc.recPartition );
  kBoundary<<<divUp( R, 256 ), 256, 0, s>>>( rc.pointToPixel, int( R ), om, W, H, prec, rc.boundary );
  PCC_LAUNCH_CHECK();
  return R;
}

void formAttributeImages( const uint32_t* pointToPixel, const uchar4* recRgb, size_t R, const uint8_t* om, int W, int H, int prec, AttrImages& out,
                          AttrTemp& at, cudaStream_t s ) {
  const size_t Q = size_t( W ) * H;
  at.T[0].reserve( Q ), at.T[1].reserve( Q ), at.tmp.reserve( Q ), at.occ.reserve( Q );
  for ( int m = 0; m < 2; ++m ) out.rawPlanes[m].reserve( 3 * Q ), out.planes[m].reserve( 3 * Q );
  PCC_CUDA( cudaMemsetAsync( at.T[0], 0, Q * sizeof( ushort4 ), s ) );
  PCC_CUDA( cudaMemsetAsync( at.T[1], 0, Q * sizeof( ushort4 ), s ) );
  if ( R ) {
    kAttrScatter<<<divUp( R, 256 ), 256, 0, s>>>( pointToPixel, recRgb, int( R ), W, at.T[0], at.T[1] );
    kAttrFallback<<<divUp( R, 256 ), 256, 0, s>>>( pointToPixel, recRgb, int( R ), W, at.T[1] );
  }
  kUpsampleOccupancy<<<divUp( Q, 256 ), 256, 0, s>>>( om, W, H, prec, at.occ );
  for ( int m = 0; m < 2; ++m ) kToPlanes<<<divUp( Q, 256 ), 256, 0, s>>>( at.T[m], Q, out.rawPlanes[m] );
  // push-pull pyramid, both frames per launch; the small levels in one CTA per frame (kPushPullTail)
  std::vector<int> lw, lh;
  {
    int w = W, h = H;
    for ( ;; ) {
      w = ( w + 1 ) >> 1, h = ( h + 1 ) >> 1;
      lw.push_back( w ), lh.push_back( h );
      if ( w <= 4 || h <= 4 ) break;
    }
  }
  const int L = int( lw.size() );
  int       T = 0;  // first level small enough for the tail kernel (its finest level, kept in global memory)
  while ( T < L - 1 && size_t( lw[T] ) * lh[T] > size_t( kTailMaxPixels ) ) ++T;
  const bool useTail = T < L - 1 && size_t( lw[T] ) * lh[T] <= size_t( kTailMaxPixels ) && L - T <= kTailLevels;
  if ( !useTail ) T = L - 1;  // (every level goes through the per-level launches)
  if ( int( at.mip.size() ) < 2 * L ) at.mip.resize( 2 * L );
  if ( int( at.mipOcc.size() ) < L ) at.mipOcc.resize( L );
  at.tmp2.reserve( Q );
  for ( int l = 0; l <= T; ++l ) {
    at.mip[l].reserve( size_t( lw[l] ) * lh[l] ), at.mip[L + l].reserve( size_t( lw[l] ) * lh[l] );
    at.mipOcc[l].reserve( size_t( lw[l] ) * lh[l] );
  }
  auto level = [&]( int l ) { return l < 0 ? Pair{ { at.T[0].p, at.T[1].p } } : Pair{ { at.mip[l].p, at.mip[L + l].p } }; };
  auto constant = []( Pair p ) { return ConstPair{ { p.p[0], p.p[1] } }; };
  for ( int l = 0; l <= T; ++l ) {  // pulls down to the tail's base level
    const uint8_t* socc = l == 0 ? at.occ.p : at.mipOcc[l - 1].p;
    const int      sw = l == 0 ? W : lw[l - 1], sh = l == 0 ? H : lh[l - 1];
    kPull<<<dim3( divUp( size_t( lw[l] ) * lh[l], 256 ), 2 ), 256, 0, s>>>( constant( level( l - 1 ) ), socc, sw, sh, level( l ), at.mipOcc[l], lw[l], lh[l] );
  }
  int iters = 4;
  if ( useTail ) {
    TailArgs a;
    a.n = L - T;
    size_t px = 0;
    for ( int l = 0; l < kTailLevels; ++l ) a.w[l] = a.h[l] = 0;
    for ( int l = T; l < L; ++l ) a.w[l - T] = lw[l], a.h[l - T] = lh[l], px += size_t( lw[l] ) * lh[l];
    a.firstIters      = 4;
    const size_t smem = ( px + size_t( lw[T] ) * lh[T] ) * sizeof( ushort4 ) + px + 16;
    {
      static std::mutex           m;
      static bool                 raised[64] = { false };
      int                         dev        = 0;
      PCC_CUDA( cudaGetDevice( &dev ) );
      std::lock_guard<std::mutex> lk( m );
      if ( dev < 0 || dev >= 64 || !raised[dev] ) {
        PCC_CUDA( cudaFuncSetAttribute( kPushPullTail, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 ) );
        if ( dev >= 0 && dev < 64 ) raised[dev] = true;
      }
    }
    kPushPullTail<<<2, kTailThreads, smem, s>>>( level( T ), at.mipOcc[T], a );
    iters = std::min( 4 + ( L - 1 - T ), 16 );
  }
  for ( int l = T; l >= 0; --l ) {  // pushes of the large levels: destination = level l-1 (the frames themselves for l = 0)
    const uint8_t* docc = l == 0 ? at.occ.p : at.mipOcc[l - 1].p;
    const int      dw = l == 0 ? W : lw[l - 1], dh = l == 0 ? H : lh[l - 1];
    const size_t   n  = size_t( dw ) * dh;
    const Pair     dst = level( l - 1 );
    kPushFill<<<dim3( divUp( n, 256 ), 2 ), 256, 0, s>>>( dst, docc, dw, dh, constant( level( l ) ), lw[l], lh[l] );
    Pair a = dst, b = Pair{ { at.tmp.p, at.tmp2.p } };
    for ( int it = 0; it < iters; ++it ) {
      kSmooth8<<<dim3( divUp( n, 256 ), 2 ), 256, 0, s>>>( constant( a ), b, docc, dw, dh );
      std::swap( a, b );
    }
    if ( a.p[0] != dst.p[0] )
      for ( int m = 0; m < 2; ++m ) PCC_CUDA( cudaMemcpyAsync( dst.p[m], a.p[m], n * sizeof( ushort4 ), cudaMemcpyDeviceToDevice, s ) );
    iters = std::min( iters + 1, 16 );
  }
  kAttrGroupDilate<<<divUp( Q, 256 ), 256, 0, s>>>( at.T[0], at.T[1], at.occ, Q );
  for ( int m = 0; m < 2; ++m ) kToPlanes<<<divUp( Q, 256 ), 256, 0, s>>>( at.T[m], Q, out.planes[m] );
  PCC_LAUNCH_CHECK();
}

}  // namespace pccb200
