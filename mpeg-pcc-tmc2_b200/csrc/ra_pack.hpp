// ra_pack.hpp — random-access packing of a GOF: PCCEncoder::spatialConsistencyPackFlexible
// (PccLibEncoder/source/PCCEncoder.cpp:1183-1412) for frames 1.., then PCCEncoder::performDataAdaptiveGPAMethod and its helpers
// (:6821-7860), as PCCEncoder::placeSegments runs them with constrainedPack 1 + globalPatchAllocation 1 (:4762-4835).
//
// What is sequential and tiny (patch matching by bounding-box IoU, global patch tracks, sub-context bookkeeping, re-ordering:
// KBs of metadata per frame) runs here on the host. What is a search - "first position in raster order where this patch fits
// the canvas" - runs on the device: the host describes every canvas session as a JOB (a canvas + an ordered list of ITEMS to
// place) and hands whole batches of jobs to a Placer; the CUDA placer (pack_ra.cu, kPlace) keeps the canvas as a bit matrix in
// shared memory and tests all candidate positions of an item in parallel. Items may take their preferred position from the
// result of an earlier item of the same batch, so the frames of a sub-context (each placed against the previous one) are
// one launch, not one launch per frame.
#pragma once
#include <algorithm>
#include <map>
#include <utility>
#include <vector>

#include "../../include/pccb200.h"

namespace pccb200 {
namespace ra {

enum PlaceMode {
  PLACE_FIXED = 0,        // no search: the position / orientation is given (a global patch at the position of its union patch)
  PLACE_BEST_EFFORT = 1,  // raster scan, two orientations per position (order by the aspect: PCCCommon.h:131-150)
  PLACE_MATCHED = 2,      // the given position first, then a raster scan with the given orientation
  PLACE_KNOWN = 3,        // raster scan with the given orientation
  PLACE_STICKY = 4        // position (0,0) with both orientations, then a raster scan with the second one (PCCEncoder.cpp:7171-7193:
                          // the orientation tried last overwrites the "unknown" marker of a union patch)
};

struct PlaceItem {
  int sizeU0, sizeV0;    // footprint in occupancy blocks
  int aspU0, aspV0;      // the sizes whose aspect selects the orientation order of a best-effort placement
  int mode;              // PlaceMode
  int refItem;           // >= 0: the given position / orientation is the RESULT of that (earlier) item of the batch
  int u0, v0, orient;    // in: given position / orientation (when refItem < 0); out: the placement
  int occOff, occStride; // occupancy flags (one byte per block, row stride occStride) in the batch's byte blob
  int pad;
};
struct PlaceJob {
  int firstItem, numItems;
  int sizeU, sizeV;         // canvas in blocks (sizeV doubles while an item does not fit)
  int widthPx, heightPx;    // out: extent of the packing in pixels (never below the initial canvas)
  int error, pad;           // out: 1 = the canvas outgrew the packer's limits
};

struct Placer {
  virtual ~Placer() {}
  virtual void run( std::vector<PlaceItem>& items, std::vector<PlaceJob>& jobs, const std::vector<uint8_t>& occ ) = 0;
};

// GPAPatchData (PccLibCommon/include/PCCPatch.h:42-71): one trial placement of a patch
struct Trial {
  bool                 matched = false, global = false;
  int                  track = -1, sizeU0 = 0, sizeV0 = 0, u0 = -1, v0 = -1, orient = -1;
  std::vector<uint8_t> occ;
};
struct Patch {
  pccb200_patch        m;
  std::vector<uint8_t> occ;  // size_u0 x size_v0 block flags
  Trial                cur, pre;
};
struct Frame {
  std::vector<Patch> patches;
  size_t             width = 0, height = 0;
};

// PCCPatch::gt (PccLibCommon/source/PCCPatch.cpp:349-371)
inline bool largerFirst( const Patch& a, const Patch& b ) {
  const int amax = std::max( a.m.size_u0, a.m.size_v0 ), amin = std::min( a.m.size_u0, a.m.size_v0 );
  const int bmax = std::max( b.m.size_u0, b.m.size_v0 ), bmin = std::min( b.m.size_u0, b.m.size_v0 );
  return amax != bmax ? amax > bmax : ( amin != bmin ? amin > bmin : a.m.index < b.m.index );
}
// pcc::computeIOU of the patches' (u1, v1, sizeU, sizeV) rectangles (PccLibEncoder/source/PCCPatchSegmenter.cpp:1563-1570)
inline float boxIou( const pccb200_patch& a, const pccb200_patch& b ) {
  const int x1 = std::max( a.u1, b.u1 ), y1 = std::max( a.v1, b.v1 );
  int       w = std::min( a.u1 + a.size_u, b.u1 + b.size_u ) - x1, h = std::min( a.v1 + a.size_v, b.v1 + b.size_v ) - y1;
  if ( w <= 0 || h <= 0 ) w = h = 0;
  const int inter = w * h, uni = a.size_u * a.size_v + b.size_u * b.size_v - inter;
  return static_cast<float>( inter ) / uni;
}

// A batch under construction
struct Batch {
  std::vector<PlaceItem> items;
  std::vector<PlaceJob>  jobs;
  std::vector<uint8_t>   occ;
  int openJob( int sizeU, int sizeV, int occRes ) {
    PlaceJob j{};
    j.firstItem = int( items.size() ), j.numItems = 0, j.sizeU = sizeU, j.sizeV = sizeV;
    j.widthPx = sizeU * occRes, j.heightPx = sizeV * occRes;
    jobs.push_back( j );
    return int( jobs.size() ) - 1;
  }
  int add( int sizeU0, int sizeV0, int aspU0, int aspV0, int mode, int refItem, int u0, int v0, int orient, const std::vector<uint8_t>& o, int stride ) {
    PlaceItem it{};
    it.sizeU0 = sizeU0, it.sizeV0 = sizeV0, it.aspU0 = aspU0, it.aspV0 = aspV0, it.mode = mode, it.refItem = refItem;
    it.u0 = u0, it.v0 = v0, it.orient = orient, it.occOff = int( occ.size() ), it.occStride = stride;
    occ.insert( occ.end(), o.begin(), o.end() );
    items.push_back( it );
    ++jobs.back().numItems;
    return int( items.size() ) - 1;
  }
};

class GofPacker {
 public:
  GofPacker( std::vector<Frame>& frames, int occRes, size_t minWidth, size_t minHeight, Placer& placer ) :
      F( frames ), occRes_( occRes ), minW_( minWidth ), minH_( minHeight ), placer_( placer ) {}

  // PCCEncoder::placeSegments for one tile, random-access condition; on return every frame holds its final patch order,
  // placements and (for global patches) union-sized occupancy, and width / height of its own packing
  bool run() {
    if ( F.empty() ) return true;
    if ( !packAgainstPrevious() ) return false;
    if ( F[0].patches.empty() ) return true;
    size_t tw = minW_, th = minH_;  // resizeTileGeometryVideo (:5593-5634)
    for ( auto& f : F ) tw = std::max( tw, f.width ), th = std::max( th, f.height );
    for ( auto& f : F ) f.width = tw, f.height = th;
    return allocateGlobalPatches();
  }

 private:
  std::vector<Frame>& F;
  int                 occRes_;
  size_t              minW_, minH_;
  Placer&             placer_;
  std::vector<size_t> curW_, curH_, preW_, preH_;
  typedef std::map<size_t, std::vector<std::pair<size_t, size_t>>> Tracks;
  struct UnionPatch {
    int                  sizeU0 = 0, sizeV0 = 0, u0 = 0, v0 = 0, orient = 0;
    std::vector<uint8_t> occ;
  };
  typedef std::map<size_t, UnionPatch> Unions;

  bool launch( Batch& b ) {
    if ( b.jobs.empty() ) return true;
    placer_.run( b.items, b.jobs, b.occ );
    for ( auto& j : b.jobs )
      if ( j.error ) return false;
    return true;
  }

  // frame 0: packFlexible (:2306-2449); frames 1..: spatialConsistencyPackFlexible (:1183-1412). The matching needs only the
  // ORDER of the previous frame's list, so all frames are matched first and placed in one batch (one job per frame; a matched
  // patch refers to the item of the patch it matches).
  bool packAgainstPrevious() {
    Batch            b;
    std::vector<int> firstItem( F.size(), -1 );
    for ( size_t f = 0; f < F.size(); ++f ) {
      auto& P    = F[f].patches;
      F[f].width = minW_;
      if ( P.empty() ) continue;
      std::sort( P.begin(), P.end(), largerFirst );
      int sizeU = int( minW_ ) / occRes_, sizeV = std::max( P[0].m.size_u0, P[0].m.size_v0 );
      if ( f > 0 ) {
        const auto&        Q = F[f - 1].patches;
        std::vector<Patch> order;
        for ( size_t qi = 0; qi < Q.size(); ++qi ) {
          float best = 0.0F;
          int   bestIdx = -1;
          for ( size_t ci = 0; ci < P.size(); ++ci )
            if ( P[ci].m.view_id == Q[qi].m.view_id && P[ci].m.best_match_idx == -1 ) {
              const float iou = boxIou( Q[qi].m, P[ci].m );
              if ( iou > best ) best = iou, bestIdx = int( ci );
            }
          if ( best > 0.2F ) {
            P[bestIdx].m.best_match_idx = int( qi );
            order.push_back( P[bestIdx] );
          }
        }
        for ( auto& p : P )
          if ( p.m.best_match_idx == -1 ) order.push_back( p );
        P.swap( order );
      }
      for ( auto& p : P ) sizeU = std::max( sizeU, p.m.size_u0 + 1 );
      const int tileH = int( ( sizeU / 2 ) * 1.0 );  // numTilesHor 2, tileHeightToWidthRatio 1
      sizeV           = std::max( sizeV, tileH );
      b.openJob( sizeU, sizeV, occRes_ );
      firstItem[f] = int( b.items.size() );
      for ( auto& p : P ) {
        const pccb200_patch& m = p.m;
        if ( f > 0 && m.best_match_idx != -1 && firstItem[f - 1] >= 0 )
          b.add( m.size_u0, m.size_v0, m.size_u0, m.size_v0, PLACE_MATCHED, firstItem[f - 1] + m.best_match_idx, 0, 0, 0, p.occ, m.size_u0 );
        else
          b.add( m.size_u0, m.size_v0, m.size_u0, m.size_v0, PLACE_BEST_EFFORT, -1, 0, 0, 0, p.occ, m.size_u0 );
      }
    }
    if ( !launch( b ) ) return false;
    size_t job = 0;
    for ( size_t f = 0; f < F.size(); ++f ) {
      if ( firstItem[f] < 0 ) continue;
      for ( size_t i = 0; i < F[f].patches.size(); ++i ) {
        const PlaceItem& it = b.items[firstItem[f] + i];
        pccb200_patch&   m  = F[f].patches[i].m;
        m.u0 = it.u0, m.v0 = it.v0, m.orientation = it.orient;
      }
      F[f].width = size_t( b.jobs[job].widthPx ), F[f].height = size_t( b.jobs[job].heightPx );
      ++job;
    }
    return true;
  }

  void clearTrials( size_t a, size_t b ) {
    for ( size_t j = a; j < b; ++j )
      for ( auto& p : F[j].patches ) p.cur = Trial();
  }

  // packingFirstFrame (:7226-7356)
  bool packFirstOfSubContext( size_t f, bool hasRef ) {
    auto& P = F[f].patches;
    int   sizeU = int( F[f].width ) / occRes_, sizeV = 0;
    for ( auto& p : P ) sizeV = std::max( sizeV, std::max( p.m.size_u0, p.m.size_v0 ) ), sizeU = std::max( sizeU, p.m.size_u0 + 1 );
    Batch b;
    b.openJob( sizeU, sizeV, occRes_ );
    for ( auto& p : P ) {
      Trial& g = p.cur;
      g.occ    = p.occ, g.sizeU0 = p.m.size_u0, g.sizeV0 = p.m.size_v0;
      if ( p.m.best_match_idx != -1 && hasRef ) {
        const pccb200_patch& q = F[f - 1].patches[p.m.best_match_idx].m;
        b.add( g.sizeU0, g.sizeV0, g.sizeU0, g.sizeV0, PLACE_MATCHED, -1, q.u0, q.v0, q.orientation, p.occ, p.m.size_u0 );
      } else {
        b.add( g.sizeU0, g.sizeV0, g.sizeU0, g.sizeV0, PLACE_BEST_EFFORT, -1, 0, 0, 0, p.occ, p.m.size_u0 );
      }
    }
    if ( !launch( b ) ) return false;
    for ( size_t i = 0; i < P.size(); ++i ) P[i].cur.u0 = b.items[i].u0, P[i].cur.v0 = b.items[i].v0, P[i].cur.orient = b.items[i].orient;
    curW_[f] = size_t( b.jobs[0].widthPx ), curH_[f] = size_t( b.jobs[0].heightPx );
    return true;
  }

  // generateGlobalPatches (:7003-7060)
  void extendTracks( size_t f, Tracks& tracks, size_t preIndex ) {
    auto& C = F[f].patches;
    for ( auto& t : tracks ) {
      auto& tp = t.second;
      if ( tp.empty() ) continue;
      const pccb200_patch& q = F[tp[preIndex].first].patches[tp[preIndex].second].m;
      float                best = 0.0F;
      int                  bestIdx = -1;
      for ( size_t ci = 0; ci < C.size(); ++ci )
        if ( q.view_id == C[ci].m.view_id && !C[ci].cur.matched ) {
          const float iou = boxIou( q, C[ci].m );
          if ( iou > best ) best = iou, bestIdx = int( ci );
        }
      if ( best > 0.2F ) {
        C[bestIdx].cur.matched = true;
        tp.emplace_back( f, size_t( bestIdx ) );
      } else {
        tp.clear();
      }
    }
    for ( auto& t : tracks )
      for ( auto& fp : t.second ) {
        Trial& g = F[fp.first].patches[fp.second].cur;
        g.global = true, g.track = int( t.first );
      }
  }

  // unionPatchGenerationAndPacking (:7062-7224); height of the union packing in pixels, or 0 with ok = false
  size_t packUnions( const Tracks& tracks, size_t frameWidth, Unions& unions, int refFrame, bool useRef, bool& ok ) {
    unions.clear();
    for ( auto& t : tracks ) {
      if ( t.second.empty() ) continue;
      UnionPatch U;
      for ( auto& fp : t.second ) {
        const pccb200_patch& m = F[fp.first].patches[fp.second].m;
        U.sizeU0 = std::max( U.sizeU0, m.size_u0 ), U.sizeV0 = std::max( U.sizeV0, m.size_v0 );
      }
      U.occ.assign( size_t( U.sizeU0 ) * U.sizeV0, 0 );
      if ( useRef ) {
        const int matched = F[t.second[0].first].patches[t.second[0].second].m.best_match_idx;
        U.orient          = matched == -1 ? -1 : F[refFrame].patches[matched].m.orientation;
      }
      for ( auto& fp : t.second ) {
        const Patch& p = F[fp.first].patches[fp.second];
        for ( int v = 0; v < p.m.size_v0; ++v )
          for ( int u = 0; u < p.m.size_u0; ++u )
            if ( p.occ[size_t( v ) * p.m.size_u0 + u] ) U.occ[size_t( v ) * U.sizeU0 + u] = 1;
      }
      unions[t.first] = std::move( U );
    }
    int sizeU = int( frameWidth ) / occRes_, sizeV = 0;
    for ( auto& u : unions ) sizeU = std::max( sizeU, u.second.sizeU0 + 1 ), sizeV = std::max( sizeV, u.second.sizeV0 + 1 );
    if ( unions.empty() ) return size_t( sizeV ) * occRes_;
    Batch b;
    b.openJob( sizeU, sizeV, occRes_ );
    for ( auto& it : unions ) {
      UnionPatch& U    = it.second;
      const int   mode = !useRef ? PLACE_BEST_EFFORT : ( U.orient != -1 ? PLACE_KNOWN : PLACE_STICKY );
      b.add( U.sizeU0, U.sizeV0, U.sizeU0, U.sizeV0, mode, -1, 0, 0, U.orient, U.occ, U.sizeU0 );
    }
    ok = launch( b );
    if ( !ok ) return 0;
    size_t i = 0;
    for ( auto& it : unions ) {
      it.second.u0 = b.items[i].u0, it.second.v0 = b.items[i].v0, it.second.orient = b.items[i].orient;
      ++i;
    }
    return size_t( b.jobs[0].heightPx );
  }

  // updateGPAPatchInformation (:7493-7529)
  void adoptUnionSizes( size_t first, size_t second, Unions& unions ) {
    for ( size_t f = first; f < second; ++f )
      for ( auto& p : F[f].patches ) {
        Trial& g = p.cur;
        if ( g.global ) {
          const UnionPatch& U = unions[size_t( g.track )];
          g.sizeU0 = U.sizeU0, g.sizeV0 = U.sizeV0;
          g.occ.assign( size_t( U.sizeU0 ) * U.sizeV0, 0 );
          for ( int v = 0; v < p.m.size_v0; ++v )
            for ( int u = 0; u < p.m.size_u0; ++u )
              if ( p.occ[size_t( v ) * p.m.size_u0 + u] ) g.occ[size_t( v ) * U.sizeU0 + u] = 1;
        } else {
          g.sizeU0 = p.m.size_u0, g.sizeV0 = p.m.size_v0, g.occ = p.occ;
        }
      }
  }

  // performGPAPacking (:7531-7660) for all frames of the trial sub-context in one batch; true = the trial is rejected
  bool packSubContext( size_t first, size_t second, Unions& unions, size_t unionsHeight, bool useRef, bool& ok ) {
    Batch                         b;
    std::vector<std::vector<int>> itemOf( second - first );  // per frame: item of every patch
    size_t                        frames = 0;
    for ( size_t f = first; f < second; ++f, ++frames ) {
      auto& P = F[f].patches;
      if ( P.empty() ) break;
      int       sizeU = int( minW_ ) / occRes_;
      const int sizeV = int( unionsHeight ) / occRes_;
      for ( auto& p : P ) sizeU = std::max( sizeU, p.cur.sizeU0 + 1 );
      b.openJob( sizeU, sizeV, occRes_ );
      auto& slot = itemOf[f - first];
      slot.assign( P.size(), -1 );
      for ( size_t i = 0; i < P.size(); ++i ) {
        Trial& g = P[i].cur;
        if ( !g.global ) continue;
        const UnionPatch& U = unions[size_t( g.track )];
        slot[i]             = b.add( g.sizeU0, g.sizeV0, g.sizeU0, g.sizeV0, PLACE_FIXED, -1, U.u0, U.v0, U.orient, g.occ, g.sizeU0 );
      }
      for ( size_t i = 0; i < P.size(); ++i ) {
        Trial& g = P[i].cur;
        if ( g.global ) continue;
        const pccb200_patch& m = P[i].m;
        if ( f == 0 || ( f == first && !useRef ) || m.best_match_idx == -1 ) {
          slot[i] = b.add( g.sizeU0, g.sizeV0, m.size_u0, m.size_v0, PLACE_BEST_EFFORT, -1, 0, 0, 0, P[i].occ, m.size_u0 );
        } else if ( f == first ) {  // the matched patch lies before the sub-context: its final placement
          const pccb200_patch& q = F[f - 1].patches[m.best_match_idx].m;
          slot[i] = b.add( g.sizeU0, g.sizeV0, g.sizeU0, g.sizeV0, PLACE_MATCHED, -1, q.u0, q.v0, q.orientation, P[i].occ, m.size_u0 );
        } else {  // inside the sub-context: its trial placement, i.e. the result of its item in this batch
          const int ref = itemOf[f - 1 - first][m.best_match_idx];
          slot[i]       = b.add( g.sizeU0, g.sizeV0, g.sizeU0, g.sizeV0, PLACE_MATCHED, ref, 0, 0, 0, P[i].occ, m.size_u0 );
        }
      }
    }
    ok = launch( b );
    if ( !ok ) return true;
    bool   tooHigh = false;
    size_t bad     = 0;
    for ( size_t k = 0; k < frames; ++k ) {
      const size_t f = first + k;
      auto&        P = F[f].patches;
      for ( size_t i = 0; i < P.size(); ++i ) {
        const PlaceItem& it = b.items[itemOf[k][i]];
        P[i].cur.u0 = it.u0, P[i].cur.v0 = it.v0, P[i].cur.orient = it.orient;
      }
      curW_[f] = size_t( b.jobs[k].widthPx ), curH_[f] = size_t( b.jobs[k].heightPx );
      if ( curH_[f] > minH_ ) {
        tooHigh = true;
        break;
      }
      if ( double( curH_[f] ) / double( F[f].height ) >= 1.10 ) ++bad;  // BAD_HEIGHT_THRESHOLD (PCCEncoder.h:111)
    }
    return tooHigh || bad > 2;  // BAD_CONDITION_THRESHOLD (PCCEncoder.h:112)
  }

  // updatePatchInformation (:7358-7491)
  void commit( size_t first, size_t second ) {
    for ( size_t f = first; f < second; ++f ) {
      F[f].width = preW_[f], F[f].height = preH_[f];
      for ( auto& p : F[f].patches ) {
        p.m.size_u0 = p.pre.sizeU0, p.m.size_v0 = p.pre.sizeV0, p.occ = p.pre.occ;
        p.m.u0 = p.pre.u0, p.m.v0 = p.pre.v0, p.m.orientation = p.pre.orient;
        p.m.is_global = p.pre.global ? 1 : 0;
      }
    }
    if ( second - first == 1 ) {
      for ( auto& p : F[first].patches ) p.m.best_match_idx = -1;
      return;
    }
    int globalCount = 0;
    for ( size_t f = first; f < second; ++f ) {
      auto& P = F[f].patches;
      for ( size_t i = 0; i < P.size(); ++i ) P[i].m.index = int( i );
      std::vector<Patch> old;
      old.swap( P );
      globalCount = 0;
      for ( auto& p : old ) globalCount += p.m.is_global;
      if ( f == first ) {
        for ( auto& p : old )
          if ( p.m.is_global ) P.push_back( p );
      } else {  // global patches in the order of the patches they match in the previous frame
        const int prevCount = int( F[f - 1].patches.size() );
        for ( int i = 0; i < prevCount; ++i )
          for ( auto& p : old )
            if ( p.m.best_match_idx == i && p.m.is_global ) {
              P.push_back( p );
              break;
            }
      }
      for ( auto& p : old )
        if ( !p.m.is_global ) P.push_back( p );
    }
    for ( size_t f = first; f < second; ++f ) {
      auto& P = F[f].patches;
      for ( int i = 0; i < globalCount && i < int( P.size() ); ++i ) {
        if ( f > first ) P[i].m.best_match_idx = i;
        P[i].m.index = i;
      }
      if ( f == second - 1 ) {
        for ( int i = globalCount; i < int( P.size() ); ++i ) P[i].m.index = i;
        continue;
      }
      auto&             N = F[f + 1].patches;
      std::vector<bool> updated( N.size(), false );
      for ( int i = globalCount; i < int( P.size() ); ++i ) {
        for ( int j = globalCount; j < int( N.size() ); ++j )
          if ( P[i].m.index == N[j].m.best_match_idx && !updated[j] ) {
            N[j].m.best_match_idx = i;
            updated[j]            = true;
            break;
          }
        P[i].m.index = i;
      }
    }
    for ( auto& p : F[first].patches ) p.m.best_match_idx = -1;
  }

  // performDataAdaptiveGPAMethod (:6821-6969): grow a sub-context frame by frame while the union packing stays good
  bool allocateGlobalPatches() {
    const size_t n = F.size();
    curW_.assign( n, 0 ), curH_.assign( n, 0 ), preW_.assign( n, 0 ), preH_.assign( n, 0 );
    size_t preFirst = 0, preSecond = 0;
    Tracks tracks;
    Unions unions;
    bool   start = true;
    for ( size_t f = 0; f < n; ++f ) {
      bool useRef = true;
      if ( start ) {  // initializeSubContext (:6971-6990) + packingFirstFrame
        preFirst = f, preSecond = f + 1;
        tracks.clear();
        auto& P = F[f].patches;
        for ( size_t i = 0; i < P.size(); ++i ) {
          tracks[i].emplace_back( f, i );
          P[i].cur.global = true, P[i].cur.track = int( i );
        }
        if ( preFirst == 0 ) useRef = false;
        if ( !packFirstOfSubContext( f, useRef ) ) return false;
        preW_[f] = curW_[f], preH_[f] = curH_[f], curW_[f] = curH_[f] = 0;
        for ( auto& p : P ) {
          p.pre = p.cur;
          p.cur = Trial();
        }
        if ( f == n - 1 ) {
          commit( preFirst, preSecond );
          break;
        }
        start = false;
        continue;
      }
      const size_t curFirst = preFirst, curSecond = f + 1;
      int          refFrame = int( curFirst ) - 1;
      if ( curFirst == 0 ) useRef = false, refFrame = -1;
      clearTrials( curFirst, curSecond );
      extendTracks( f, tracks, f - curFirst - 1 );
      bool         ok = true;
      const size_t unionsHeight = packUnions( tracks, F[f].width, unions, refFrame, useRef, ok );
      if ( !ok ) return false;
      bool       badCount  = double( unions.size() ) / double( tracks.size() ) < 0.15;
      const bool badHeight = unionsHeight > minH_;
      if ( unionsHeight == 0 ) badCount = true;
      bool badPacking = false;
      if ( !badCount && !badHeight ) {
        adoptUnionSizes( curFirst, curSecond, unions );
        badPacking = packSubContext( curFirst, curSecond, unions, unionsHeight, useRef, ok );
        if ( !ok ) return false;
      }
      if ( badCount || badHeight || badPacking ) {  // the previous trial stands; frame f opens the next sub-context
        clearTrials( curFirst, curSecond );
        unions.clear();
        tracks.clear();
        start = true;
        --f;
        commit( preFirst, preSecond );
      } else {
        for ( size_t j = curFirst; j < curSecond; ++j ) {
          preW_[j] = curW_[j], preH_[j] = curH_[j];
          for ( auto& p : F[j].patches ) p.pre = p.cur;
        }
        preFirst = curFirst, preSecond = curSecond;
        clearTrials( curFirst, curSecond );
        unions.clear();
        if ( f == n - 1 ) {
          commit( preFirst, preSecond );
          break;
        }
      }
    }
    return true;
  }
};

}  // namespace ra
}  // namespace pccb200
