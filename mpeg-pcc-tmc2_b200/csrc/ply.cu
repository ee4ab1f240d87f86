// ply.cu — input side: PCCPointSet3::read (PccLibCommon/source/PCCPointSet.cpp:464-757) for what the hot path
// consumes, positions (int16 x 3) and colours (uint8 x 3), straight into caller (ideally pinned) buffers. Host code only: the
// reference parses an ascii frame token by token through std::vector<std::string> + atof (one thread, ~4 MB per frame), which
// becomes the wall-clock floor once the hot path runs at tens of frames per second; here the file is read once, the body is cut
// at line boundaries and parsed by a few threads, integers and plain decimals without strtod.
// Semantics kept: header grammar and property-type table of the reference, values converted like its assignments do
// (double -> int16 and int -> uint8 conversions of atof / atoi results; binary x/y/z of 2 bytes are read as uint16), blank body
// lines skipped, a body line with fewer tokens than properties ends the read with an error.
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pccb200.h"

namespace {

struct Property {
  std::string name;
  int         bytes;   // 0 = unknown type (the reference leaves such a property without a size: nothing is skipped for it)
  int         kind;    // 0 float, 1 unsigned, 2 signed
};
struct Header {
  bool                  ascii = false;
  double                version = 1.0;
  size_t                count = 0, bodyOffset = 0;
  std::vector<Property> props;
  int                   ix = -1, iy = -1, iz = -1, ir = -1, ig = -1, ib = -1;
  bool                  colours = false;
};

inline bool isSep( char c ) { return c == ' ' || c == '\t' || c == '\r'; }

void tokens( const char* b, const char* e, std::vector<std::string>& out ) {
  out.clear();
  while ( b < e ) {
    while ( b < e && isSep( *b ) ) ++b;
    const char* t = b;
    while ( b < e && !isSep( *b ) ) ++b;
    if ( b > t ) out.emplace_back( t, b );
  }
}

// header lines end at '\n'; the reference reads them with getline into a 4 KB buffer
int parseHeader( const char* data, size_t size, Header& h ) {
  size_t                   pos = 0;
  std::vector<std::string> tk;
  auto                     line = [&]( const char*& b, const char*& e ) -> bool {
    if ( pos >= size ) return false;
    const char* nl = static_cast<const char*>( memchr( data + pos, '\n', size - pos ) );
    b              = data + pos;
    e              = nl ? nl : data + size;
    pos            = nl ? size_t( nl - data ) + 1 : size;
    return true;
  };
  const char *b, *e;
  if ( !line( b, e ) ) return PCCB200_ERR_BAD_ARG;
  tokens( b, e, tk );
  if ( tk.empty() || tk[0] != "ply" ) return PCCB200_ERR_BAD_ARG;
  bool vertex = true;
  for ( ;; ) {
    if ( !line( b, e ) ) return PCCB200_ERR_BAD_ARG;  // corrupted header
    tokens( b, e, tk );
    if ( tk.empty() || tk[0] == "comment" ) continue;
    if ( tk[0] == "format" ) {
      if ( tk.size() != 3 ) return PCCB200_ERR_BAD_ARG;
      h.ascii   = tk[1] == "ascii";
      h.version = atof( tk[2].c_str() );
    } else if ( tk[0] == "element" ) {
      if ( tk.size() != 3 ) return PCCB200_ERR_BAD_ARG;
      if ( tk[1] == "vertex" ) h.count = size_t( atoi( tk[2].c_str() ) );
      else vertex = false;
    } else if ( tk[0] == "property" && vertex ) {
      if ( tk.size() != 3 ) return PCCB200_ERR_BAD_ARG;
      Property           p{tk[2], 0, 0};
      const std::string& t = tk[1];
      if ( t == "double" || t == "float64" ) p.bytes = 8, p.kind = 0;
      else if ( t == "float" || t == "float32" ) p.bytes = 4, p.kind = 0;
      else if ( t == "uint64" ) p.bytes = 8, p.kind = 1;
      else if ( t == "uint32" ) p.bytes = 4, p.kind = 1;
      else if ( t == "uint16" ) p.bytes = 2, p.kind = 1;
      else if ( t == "uchar" || t == "uint8" ) p.bytes = 1, p.kind = 1;
      else if ( t == "int64" ) p.bytes = 8, p.kind = 2;
      else if ( t == "int32" || t == "int" ) p.bytes = 4, p.kind = 2;
      else if ( t == "int16" ) p.bytes = 2, p.kind = 2;
      else if ( t == "char" || t == "int8" ) p.bytes = 1, p.kind = 2;
      h.props.push_back( p );
    } else if ( tk[0] == "end_header" ) {
      break;
    }
  }
  if ( h.version != 1.0 ) return PCCB200_ERR_UNSUPPORTED;
  h.bodyOffset = pos;
  for ( size_t a = 0; a < h.props.size(); ++a ) {
    const Property& p     = h.props[a];
    const bool      coord = p.bytes == 8 || p.bytes == 4 || p.bytes == 2;
    if ( p.name == "x" && coord ) h.ix = int( a );
    else if ( p.name == "y" && coord ) h.iy = int( a );
    else if ( p.name == "z" && coord ) h.iz = int( a );
    else if ( p.name == "red" && p.bytes == 1 ) h.ir = int( a );
    else if ( p.name == "green" && p.bytes == 1 ) h.ig = int( a );
    else if ( p.name == "blue" && p.bytes == 1 ) h.ib = int( a );
  }
  if ( h.ix < 0 || h.iy < 0 || h.iz < 0 ) return PCCB200_ERR_BAD_ARG;  // missing coordinates
  h.colours = h.ir >= 0 && h.ig >= 0 && h.ib >= 0;
  return PCCB200_OK;
}

// atof of a token, converted like `int16 = double` does; digits-only and "digits.digits" tokens (what voxelised clouds hold)
// are evaluated without strtod: only their integer part matters after the truncating conversion, unless the token has an exponent
inline int16_t coordinate( const char* b, const char* e ) {
  const char* p   = b;
  bool        neg = false;
  if ( p < e && ( *p == '-' || *p == '+' ) ) neg = *p == '-', ++p;
  long long   v = 0;
  const char* d = p;
  while ( p < e && *p >= '0' && *p <= '9' && p - d < 15 ) v = v * 10 + ( *p - '0' ), ++p;
  bool simple = p > d;
  if ( simple && p < e && *p == '.' ) {
    ++p;
    while ( p < e && *p >= '0' && *p <= '9' ) ++p;
  }
  if ( simple && p == e ) return int16_t( neg ? -v : v );  // (truncation toward zero == dropping the fraction)
  char buf[64];
  const size_t len = std::min<size_t>( size_t( e - b ), sizeof( buf ) - 1 );
  memcpy( buf, b, len );
  buf[len] = 0;
  return int16_t( atof( buf ) );
}
inline uint8_t colourByte( const char* b, const char* e ) {  // uint8 = atoi( token )
  char         buf[32];
  const size_t len = std::min<size_t>( size_t( e - b ), sizeof( buf ) - 1 );
  memcpy( buf, b, len );
  buf[len] = 0;
  return uint8_t( atoi( buf ) );
}

struct Chunk {
  size_t begin, end;      // byte range of the body, cut at line starts
  size_t lines = 0;       // non-blank lines in it
  size_t firstPoint = 0;  // index of its first point
  bool   shortLine = false;
};

// walks the lines of a chunk; emit( tokenStarts, tokenEnds, count ) per non-blank line; returns false when emit does
template <class F>
bool forLines( const char* data, const Chunk& c, size_t maxTokens, F&& emit ) {
  std::vector<const char*> tb( maxTokens ), te( maxTokens );
  const char*              p   = data + c.begin;
  const char* const        end = data + c.end;
  while ( p < end ) {
    const char* nl = static_cast<const char*>( memchr( p, '\n', size_t( end - p ) ) );
    const char* le = nl ? nl : end;
    size_t      nt = 0, total = 0;
    const char* q  = p;
    while ( q < le ) {
      while ( q < le && isSep( *q ) ) ++q;
      const char* t = q;
      while ( q < le && !isSep( *q ) ) ++q;
      if ( q > t ) {
        if ( nt < maxTokens ) tb[nt] = t, te[nt] = q, ++nt;
        ++total;
      }
    }
    if ( total && !emit( tb.data(), te.data(), total ) ) return false;
    p = nl ? nl + 1 : end;
  }
  return true;
}

}  // namespace

// one frame; an ascii body is parsed by up to maxThreads threads
static int readOne( const char* path, int16_t* xyz, uint8_t* rgb, size_t capacity, size_t* n, int* hasColours, size_t maxThreads ) {
  if ( !path || !n ) return PCCB200_ERR_BAD_ARG;
  *n = 0;
  FILE* f = fopen( path, "rb" );
  if ( !f ) return PCCB200_ERR_BAD_ARG;
  fseek( f, 0, SEEK_END );
  const long fileSize = ftell( f );
  fseek( f, 0, SEEK_SET );
  // (a size query needs the header only: the first 64 KB, the whole file if the header should be longer than that)
  const size_t      whole = fileSize > 0 ? size_t( fileSize ) : 0;
  std::vector<char> file( xyz ? whole : std::min<size_t>( whole, 65536 ) );
  size_t            got = file.empty() ? 0 : fread( file.data(), 1, file.size(), f );
  Header            h;
  int               rc = got == file.size() ? parseHeader( file.data(), file.size(), h ) : PCCB200_ERR_BAD_ARG;
  if ( rc != PCCB200_OK && got == file.size() && file.size() < whole ) {
    file.resize( whole );
    got += fread( file.data() + got, 1, whole - got, f );
    h  = Header();
    rc = got == whole ? parseHeader( file.data(), file.size(), h ) : PCCB200_ERR_BAD_ARG;
  }
  fclose( f );
  if ( rc != PCCB200_OK ) return rc;
  if ( hasColours ) *hasColours = h.colours ? 1 : 0;
  *n = h.count;
  if ( !xyz ) return PCCB200_OK;  // size query
  if ( capacity < h.count ) return PCCB200_ERR_CAPACITY;
  const char*  data = file.data();
  const size_t size = file.size(), np = h.props.size();
  memset( xyz, 0, h.count * 3 * sizeof( int16_t ) );  // (points a short file does not reach stay zero, as after resize())
  if ( rgb && h.colours ) memset( rgb, 0, h.count * 3 );
  if ( h.ascii ) {
    const size_t body    = size - h.bodyOffset;
    const int    threads = int( std::max<size_t>( 1, std::min<size_t>( std::min<size_t>( maxThreads, std::thread::hardware_concurrency() ), body >> 20 ) ) );
    std::vector<Chunk> chunks( threads );
    size_t             at = h.bodyOffset;
    for ( int t = 0; t < threads; ++t ) {
      chunks[t].begin = at;
      size_t cut      = t + 1 == threads ? size : h.bodyOffset + body * ( t + 1 ) / threads;
      if ( cut < at ) cut = at;
      if ( t + 1 < threads ) {  // move the cut to the next line start
        const char* nl = static_cast<const char*>( memchr( data + cut, '\n', size - cut ) );
        cut            = nl ? size_t( nl - data ) + 1 : size;
      }
      chunks[t].end = at = cut;
    }
    auto inParallel = [&]( auto&& fn ) {
      std::vector<std::thread> th;
      for ( int t = 1; t < threads; ++t ) th.emplace_back( [&, t]() { fn( chunks[t] ); } );
      fn( chunks[0] );
      for ( auto& x : th ) x.join();
    };
    inParallel( [&]( Chunk& c ) {  // pass 1: count the points of every chunk
      forLines( data, c, 1, [&]( const char**, const char**, size_t ) {
        ++c.lines;
        return true;
      } );
    } );
    size_t first = 0;
    for ( auto& c : chunks ) c.firstPoint = first, first += c.lines;
    inParallel( [&]( Chunk& c ) {  // pass 2: parse
      size_t i = c.firstPoint;
      forLines( data, c, np, [&]( const char** tb, const char** te, size_t total ) {
        if ( i >= h.count ) return false;  // the reference stops after pointCount points
        if ( total < np ) {
          c.shortLine = true;
          return false;
        }
        xyz[3 * i]     = coordinate( tb[h.ix], te[h.ix] );
        xyz[3 * i + 1] = coordinate( tb[h.iy], te[h.iy] );
        xyz[3 * i + 2] = coordinate( tb[h.iz], te[h.iz] );
        if ( rgb && h.colours ) {
          rgb[3 * i]     = colourByte( tb[h.ir], te[h.ir] );
          rgb[3 * i + 1] = colourByte( tb[h.ig], te[h.ig] );
          rgb[3 * i + 2] = colourByte( tb[h.ib], te[h.ib] );
        }
        ++i;
        return true;
      } );
    } );
    for ( auto& c : chunks )
      if ( c.shortLine ) return PCCB200_ERR_BAD_ARG;  // (the reference returns false at the first short line)
    return PCCB200_OK;
  }
  // binary (the reference reads little-endian host order): fixed record size
  size_t record = 0;
  for ( auto& p : h.props ) record += size_t( p.bytes );
  std::vector<size_t> offset( np );
  size_t              o = 0;
  for ( size_t a = 0; a < np; ++a ) offset[a] = o, o += size_t( h.props[a].bytes );
  const size_t avail = record ? ( size - h.bodyOffset ) / record : 0, pts = std::min( h.count, avail );
  auto         coord = [&]( const char* rec, int a ) -> int16_t {
    const char* p = rec + offset[a];
    if ( h.props[a].bytes == 2 ) {
      uint16_t v;
      memcpy( &v, p, 2 );
      return int16_t( v );
    }
    if ( h.props[a].bytes == 4 ) {
      float v;
      memcpy( &v, p, 4 );
      return int16_t( v );
    }
    double v;
    memcpy( &v, p, 8 );
    return int16_t( v );
  };
  for ( size_t i = 0; i < pts; ++i ) {
    const char* rec = data + h.bodyOffset + i * record;
    xyz[3 * i] = coord( rec, h.ix ), xyz[3 * i + 1] = coord( rec, h.iy ), xyz[3 * i + 2] = coord( rec, h.iz );
    if ( rgb && h.colours ) {
      rgb[3 * i]     = uint8_t( rec[offset[h.ir]] );
      rgb[3 * i + 1] = uint8_t( rec[offset[h.ig]] );
      rgb[3 * i + 2] = uint8_t( rec[offset[h.ib]] );
    }
  }
  return PCCB200_OK;
}

extern "C" int pccb200_ply_read( const char* path, int16_t* xyz, uint8_t* rgb, size_t capacity, size_t* n, int* hasColours ) {
  return readOne( path, xyz, rgb, capacity, n, hasColours, 8 );
}

// PCCGroupOfFrames::load (PccLibCommon/source/PCCGroupOfFrames.cpp:46-83): frame k of the group is the file sprintf( pattern,
// start + k ); the reference reads them one after the other (in parallel only in a TBB build) and cuts the group at the first
// file it cannot read. Here `threads` workers take the frames in order and share the remaining threads for the ascii bodies.
extern "C" int pccb200_ply_read_frames( const char* pathPattern, size_t startFrame, size_t endFrame, int16_t* const* xyz, uint8_t* const* rgb,
                                        const size_t* capacity, size_t* n, int* hasColours, int threads, size_t* framesRead ) {
  if ( framesRead ) *framesRead = 0;
  if ( !pathPattern || !n || endFrame < startFrame || ( xyz && !capacity ) ) return PCCB200_ERR_BAD_ARG;
  const size_t count = endFrame - startFrame;
  if ( count == 0 ) return PCCB200_OK;
  size_t hw = std::thread::hardware_concurrency();
  if ( hw == 0 ) hw = 1;
  const size_t total   = threads > 0 ? size_t( threads ) : hw;
  const size_t workers = std::max<size_t>( 1, std::min( total, count ) );
  const size_t inner   = std::max<size_t>( 1, std::min<size_t>( 8, total / workers ) );
  std::vector<int>    status( count, PCCB200_OK );
  std::atomic<size_t> next( 0 );
  auto work = [&]() {
    for ( ;; ) {
      const size_t k = next.fetch_add( 1 );
      if ( k >= count ) return;
      char name[4096];
      snprintf( name, sizeof( name ), pathPattern, startFrame + k );  // (a size_t behind the caller's %d pattern, as in the reference)
      int colours = 0;
      status[k]   = readOne( name, xyz ? xyz[k] : nullptr, ( xyz && rgb ) ? rgb[k] : nullptr, xyz ? capacity[k] : 0, &n[k], &colours, inner );
      if ( hasColours ) hasColours[k] = colours;
    }
  };
  std::vector<std::thread> pool;
  for ( size_t t = 1; t < workers; ++t ) pool.emplace_back( work );
  work();
  for ( auto& t : pool ) t.join();
  size_t good = 0;
  while ( good < count && status[good] == PCCB200_OK ) ++good;
  if ( framesRead ) *framesRead = good;
  return good == count ? PCCB200_OK : status[good];
}
