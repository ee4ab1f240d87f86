// ra_host_shim.cpp — TEST-ONLY harness for the HOST logic of the product's random-access packer
// (mpeg-pcc-tmc2_b200/csrc/ra_pack.hpp) on machines without a GPU: the same GofPacker the library runs, driven through a
// sequential stand-in for the CUDA placement kernel (pack_ra.cu, kPlace). Built by tests/test_ra_pack_host.py with g++;
// never part of libpccb200.so - the product has no CPU path.
#include <cstring>

#include "../mpeg-pcc-tmc2_b200/csrc/ra_pack.hpp"

using namespace pccb200::ra;

namespace {

struct SequentialPlacer : Placer {
  int occRes;
  explicit SequentialPlacer( int r ) : occRes( r ) {}
  static bool fits( const std::vector<uint8_t>& c, int sizeU, int sizeV, int u, int v, int o, int sU0, int sV0 ) {
    const int bw = o == 0 ? sU0 : sV0, bh = o == 0 ? sV0 : sU0;
    if ( u < 0 || v < 0 || u + bw > sizeU || v + bh > sizeV ) return false;
    for ( int y = v; y < v + bh; ++y )
      for ( int x = u; x < u + bw; ++x )
        if ( c[size_t( y ) * sizeU + x] ) return false;
    return true;
  }
  void run( std::vector<PlaceItem>& items, std::vector<PlaceJob>& jobs, const std::vector<uint8_t>& occ ) override {
    for ( auto& job : jobs ) {
      const int            sizeU = job.sizeU;
      int                  sizeV = job.sizeV;
      std::vector<uint8_t> c( size_t( sizeU ) * sizeV, 0 );
      for ( int k = 0; k < job.numItems; ++k ) {
        PlaceItem& it = items[job.firstItem + k];
        if ( it.refItem >= 0 ) it.u0 = items[it.refItem].u0, it.v0 = items[it.refItem].v0, it.orient = items[it.refItem].orient;
        const int o0 = it.aspU0 > it.aspV0 ? 1 : 0, o1 = o0 ^ 1;
        int       mode = it.mode;
        bool      found = mode == PLACE_FIXED;
        if ( mode == PLACE_STICKY ) {
          if ( fits( c, sizeU, sizeV, 0, 0, o0, it.sizeU0, it.sizeV0 ) ) it.u0 = it.v0 = 0, it.orient = o0, found = true;
          else if ( fits( c, sizeU, sizeV, 0, 0, o1, it.sizeU0, it.sizeV0 ) ) it.u0 = it.v0 = 0, it.orient = o1, found = true;
          else it.orient = o1, mode = PLACE_KNOWN;
        }
        while ( !found ) {
          if ( mode == PLACE_MATCHED && fits( c, sizeU, sizeV, it.u0, it.v0, it.orient, it.sizeU0, it.sizeV0 ) ) break;
          for ( int v = 0; v < sizeV && !found; ++v )
            for ( int u = 0; u < sizeU && !found; ++u ) {
              if ( mode == PLACE_BEST_EFFORT ) {
                for ( int o : {o0, o1} )
                  if ( !found && fits( c, sizeU, sizeV, u, v, o, it.sizeU0, it.sizeV0 ) ) it.u0 = u, it.v0 = v, it.orient = o, found = true;
              } else if ( fits( c, sizeU, sizeV, u, v, it.orient, it.sizeU0, it.sizeV0 ) ) {
                it.u0 = u, it.v0 = v, found = true;
              }
            }
          if ( !found ) {
            sizeV *= 2;
            c.resize( size_t( sizeU ) * sizeV, 0 );
          }
        }
        for ( int vb = 0; vb < it.sizeV0; ++vb )
          for ( int ub = 0; ub < it.sizeU0; ++ub ) {
            if ( !occ[it.occOff + vb * it.occStride + ub] ) continue;
            const int x = it.orient == 0 ? ub + it.u0 : vb + it.u0, y = it.orient == 0 ? vb + it.v0 : ub + it.v0;
            if ( x < sizeU && y < sizeV ) c[size_t( y ) * sizeU + x] = 1;
          }
        job.heightPx = std::max( job.heightPx, ( it.v0 + ( it.orient == 0 ? it.sizeV0 : it.sizeU0 ) ) * occRes );
        job.widthPx  = std::max( job.widthPx, ( it.u0 + ( it.orient == 0 ? it.sizeU0 : it.sizeV0 ) ) * occRes );
      }
    }
  }
};

std::vector<Frame> gFrames;

}  // namespace

extern "C" {

// frames in: for frame f, counts[f] patches (creation order) starting at patches[first[f]], occupancy bytes in occ at occ_offset
int ra_shim_pack( int nframes, const int* counts, const pccb200_patch* patches, const uint8_t* occ, const int64_t* occBase, int occRes, int minW,
                  int minH ) {
  gFrames.assign( nframes, Frame() );
  size_t at = 0;
  for ( int f = 0; f < nframes; ++f ) {
    gFrames[f].patches.resize( counts[f] );
    for ( int i = 0; i < counts[f]; ++i, ++at ) {
      Patch& p = gFrames[f].patches[i];
      p.m      = patches[at];
      p.m.best_match_idx = -1, p.m.is_global = 0;
      const uint8_t* o = occ + occBase[f] + p.m.occ_offset;
      p.occ.assign( o, o + size_t( p.m.size_u0 ) * p.m.size_v0 );
    }
  }
  SequentialPlacer placer( occRes );
  GofPacker        packer( gFrames, occRes, size_t( minW ), size_t( minH ), placer );
  return packer.run() ? 0 : -1;
}
int    ra_shim_count( int f ) { return int( gFrames[f].patches.size() ); }
size_t ra_shim_occ_bytes( int f ) {
  size_t n = 0;
  for ( auto& p : gFrames[f].patches ) n += p.occ.size();
  return n;
}
void ra_shim_get( int f, pccb200_patch* out, uint8_t* occ, int64_t* wh ) {
  size_t off = 0;
  for ( size_t i = 0; i < gFrames[f].patches.size(); ++i ) {
    out[i]            = gFrames[f].patches[i].m;
    out[i].occ_offset = int64_t( off );
    std::memcpy( occ + off, gFrames[f].patches[i].occ.data(), gFrames[f].patches[i].occ.size() );
    off += gFrames[f].patches[i].occ.size();
  }
  wh[0] = int64_t( gFrames[f].width ), wh[1] = int64_t( gFrames[f].height );
}
}
