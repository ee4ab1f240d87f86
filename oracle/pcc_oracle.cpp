// pcc_oracle.cpp — CPU oracle (TEST INFRASTRUCTURE ONLY, see pcc_oracle.h).
//
// A from-scratch restatement of the TMC2 v24.0 hot-path algorithms, written for clarity, single-threaded,
// fp64 evaluated in the reference's operation order (build with -ffp-contract=off on x86-64).
// Citations are relative to the reference checkout: L/ = source/lib/, NF = dependencies/nanoflann/nanoflann.hpp.
#include "pcc_oracle.h"

#include <algorithm>
#include <map>
#include <cmath>
#include <cstring>
#include <limits>
#include <queue>
#include <unordered_map>
#include <vector>

namespace {

// =====================================================================================================
// a1. kd-tree with nanoflann's build + traversal semantics.
//     build : NF:858-866 (buildIndex), NF:1041-1089 (divideTree), NF:1103-1142 (middleSplit_),
//             NF:1154-1181 (planeSplit).   leaf size 10: L/PccLibCommon/source/PCCKdTree.cpp:58.
//     search: NF:1207-1254 (searchLevel), NF:1183-1200 (computeInitialDistances),
//             NF:110-131 (KNNResultSet::addPoint: ties keep the first-visited point).
//     metric: squared L2 accumulated in float (KDTreeVectorOfVectorsAdaptor.h:126-133), compared as double.
// =====================================================================================================
struct Node {
  int32_t child[2];  // -1,-1 for a leaf
  int32_t lo, hi;    // leaf: range in vind
  int32_t feat;      // internal: split dimension
  int32_t divlow, divhigh;
};

struct Tree {
  std::vector<int16_t>  pts;  // n x 3
  std::vector<uint32_t> vind;
  std::vector<Node>     nodes;
  size_t                n    = 0;
  int                   root = -1;
  int                   rootBox[3][2];
  static const int      kLeaf = 10;

  int16_t at( uint32_t i, int d ) const { return pts[3 * size_t( i ) + d]; }

  void minmax( size_t first, size_t count, int d, int& mn, int& mx ) const {
    mn = mx = at( vind[first], d );
    for ( size_t i = 1; i < count; ++i ) {
      int v = at( vind[first + i], d );
      mn    = std::min( mn, v );
      mx    = std::max( mx, v );
    }
  }

  // Two Hoare-style sweeps: [<cut | ==cut | >cut]; the resulting order of vind is part of the contract.
  void planeSplit( size_t first, size_t count, int d, int cut, size_t& lim1, size_t& lim2 ) {
    uint32_t* ind = &vind[first];
    size_t    l = 0, r = count - 1;
    for ( ;; ) {
      while ( l <= r && at( ind[l], d ) < cut ) ++l;
      while ( r != 0 && l <= r && at( ind[r], d ) >= cut ) --r;
      if ( l > r || r == 0 ) break;
      std::swap( ind[l], ind[r] );
      ++l, --r;
    }
    lim1 = l;
    r    = count - 1;
    for ( ;; ) {
      while ( l <= r && at( ind[l], d ) <= cut ) ++l;
      while ( r != 0 && l <= r && at( ind[r], d ) > cut ) --r;
      if ( l > r || r == 0 ) break;
      std::swap( ind[l], ind[r] );
      ++l, --r;
    }
    lim2 = l;
  }

  // box is the (loose) cell on entry and the tight bounds of the subtree on return.
  int divide( size_t left, size_t right, int box[3][2] ) {
    int id = int( nodes.size() );
    nodes.push_back( Node() );
    if ( right - left <= size_t( kLeaf ) ) {
      Node nd;
      nd.child[0] = nd.child[1] = -1;
      nd.lo = int32_t( left ), nd.hi = int32_t( right );
      nd.feat = nd.divlow = nd.divhigh = 0;
      for ( int d = 0; d < 3; ++d ) {
        int mn, mx;
        minmax( left, right - left, d, mn, mx );
        box[d][0] = mn, box[d][1] = mx;
      }
      nodes[id] = nd;
      return id;
    }
    const size_t count = right - left;
    // dimension of largest cell span (within 1e-5), tie-broken by the largest actual spread (first wins)
    int maxSpan = box[0][1] - box[0][0];
    for ( int d = 1; d < 3; ++d ) maxSpan = std::max( maxSpan, box[d][1] - box[d][0] );
    int feat = 0, bestSpread = -1;
    for ( int d = 0; d < 3; ++d ) {
      // ElementType is int16_t: the span is truncated to int16 before the comparison in double
      int16_t span = int16_t( box[d][1] - box[d][0] );
      if ( double( span ) > ( 1.0 - double( 0.00001 ) ) * double( int16_t( maxSpan ) ) ) {
        int mn, mx;
        minmax( left, count, d, mn, mx );
        int16_t spread = int16_t( mx - mn );
        if ( spread > bestSpread ) feat = d, bestSpread = spread;
      }
    }
    int split = ( box[feat][0] + box[feat][1] ) / 2;  // integer division (operands promoted from int16)
    int mn, mx;
    minmax( left, count, feat, mn, mx );
    int    cut = split < mn ? mn : ( split > mx ? mx : split );
    size_t lim1, lim2, idx;
    planeSplit( left, count, feat, cut, lim1, lim2 );
    if ( lim1 > count / 2 )
      idx = lim1;
    else if ( lim2 < count / 2 )
      idx = lim2;
    else
      idx = count / 2;
    int lbox[3][2], rbox[3][2];
    std::memcpy( lbox, box, sizeof( lbox ) );
    std::memcpy( rbox, box, sizeof( rbox ) );
    lbox[feat][1] = cut;
    rbox[feat][0] = cut;
    int c0        = divide( left, left + idx, lbox );
    int c1        = divide( left + idx, right, rbox );
    Node nd;
    nd.child[0] = c0, nd.child[1] = c1;
    nd.lo = nd.hi = 0;
    nd.feat       = feat;
    nd.divlow     = lbox[feat][1];
    nd.divhigh    = rbox[feat][0];
    nodes[id]     = nd;
    for ( int d = 0; d < 3; ++d ) {
      box[d][0] = std::min( lbox[d][0], rbox[d][0] );
      box[d][1] = std::max( lbox[d][1], rbox[d][1] );
    }
    return id;
  }

  void build( const int16_t* xyz, size_t count ) {
    n = count;
    pts.assign( xyz, xyz + 3 * n );
    vind.resize( n );
    for ( size_t i = 0; i < n; ++i ) vind[i] = uint32_t( i );
    nodes.clear();
    root = -1;
    if ( n == 0 ) return;
    for ( int d = 0; d < 3; ++d ) {
      int mn, mx;
      minmax( 0, n, d, mn, mx );
      rootBox[d][0] = mn, rootBox[d][1] = mx;
    }
    nodes.reserve( n / 4 + 16 );
    root = divide( 0, n, rootBox );
  }

  double dist2( const int16_t* q, uint32_t i ) const {
    float s = 0;
    for ( int d = 0; d < 3; ++d ) {
      const float e = float( q[d] ) - float( at( i, d ) );
      s += e * e;
    }
    return double( s );
  }

  // RS must provide worst() and add(dist, index)
  template <class RS>
  void descend( RS& rs, const int16_t* q, int node, double mindist, double side[3] ) const {
    const Node& nd = nodes[node];
    if ( nd.child[0] < 0 ) {
      const double worstAtEntry = rs.worst();
      for ( int32_t i = nd.lo; i < nd.hi; ++i ) {
        const double d = dist2( q, vind[i] );
        if ( d < worstAtEntry ) rs.add( d, vind[i] );
      }
      return;
    }
    const int    f     = nd.feat;
    const double v     = double( q[f] );
    const double diff1 = v - double( nd.divlow ), diff2 = v - double( nd.divhigh );
    int          nearC, farC;
    double       cut;
    if ( diff1 + diff2 < 0 ) {
      nearC = nd.child[0], farC = nd.child[1];
      cut   = diff2 * diff2;
    } else {
      nearC = nd.child[1], farC = nd.child[0];
      cut   = diff1 * diff1;
    }
    descend( rs, q, nearC, mindist, side );
    const double saved = side[f];
    mindist            = mindist + cut - saved;
    side[f]            = cut;
    if ( mindist <= rs.worst() ) descend( rs, q, farC, mindist, side );
    side[f] = saved;
  }

  template <class RS>
  void search( RS& rs, const int16_t* q ) const {
    if ( n == 0 ) return;
    double side[3] = {0, 0, 0}, mind = 0;
    for ( int d = 0; d < 3; ++d ) {
      if ( q[d] < rootBox[d][0] ) {
        double e = double( q[d] ) - double( rootBox[d][0] );
        side[d]  = e * e;
        mind += side[d];
      }
      if ( q[d] > rootBox[d][1] ) {
        double e = double( q[d] ) - double( rootBox[d][1] );
        side[d]  = e * e;
        mind += side[d];
      }
    }
    descend( rs, q, root, mind, side );
  }
};

struct KnnSet {
  int       cap, cnt = 0;
  uint32_t* idx;
  double*   dist;
  KnnSet( int k, uint32_t* i, double* d ) : cap( k ), idx( i ), dist( d ) {
    if ( cap ) dist[cap - 1] = std::numeric_limits<double>::max();
  }
  double worst() const { return dist[cap - 1]; }
  void   add( double d, uint32_t i ) {
    int p = cnt;
    while ( p > 0 && dist[p - 1] > d ) {  // strictly greater: equal distances keep earlier arrivals in front
      if ( p < cap ) dist[p] = dist[p - 1], idx[p] = idx[p - 1];
      --p;
    }
    if ( p < cap ) dist[p] = d, idx[p] = i;
    if ( cnt < cap ) ++cnt;
  }
};

struct RadiusSet {
  double                                    r2;
  std::vector<std::pair<double, uint32_t>>& out;
  RadiusSet( double r, std::vector<std::pair<double, uint32_t>>& o ) : r2( r ), out( o ) {}
  double worst() const { return r2; }
  void   add( double d, uint32_t i ) {
    if ( d < r2 ) out.emplace_back( d, i );
  }
};

inline double dot3( const double* a, const double* b ) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// =====================================================================================================
// a2. Symmetric 3x3 eigen-decomposition by quaternion Jacobi sweeps (L/PccLibCommon/include/PCCMath.h:505-598).
//     The floating-point evaluation order is part of the contract (normals feed an argmax).
// =====================================================================================================
void diagonalize( const double A[3][3], double Q[3][3], double D[3][3] ) {
  double q[4] = {0, 0, 0, 1};
  for ( int step = 0; step < 24; ++step ) {
    const double xx = q[0] * q[0], yy = q[1] * q[1], zz = q[2] * q[2], ww = q[3] * q[3];
    Q[0][0] = ( xx - yy - zz + ww );
    Q[1][1] = ( -xx + yy - zz + ww );
    Q[2][2] = ( -xx - yy + zz + ww );
    double a = q[0] * q[1], b = q[2] * q[3];
    Q[1][0] = 2.0 * ( a + b ), Q[0][1] = 2.0 * ( a - b );
    a = q[0] * q[2], b = q[1] * q[3];
    Q[2][0] = 2.0 * ( a - b ), Q[0][2] = 2.0 * ( a + b );
    a = q[1] * q[2], b = q[0] * q[3];
    Q[2][1] = 2.0 * ( a + b ), Q[1][2] = 2.0 * ( a - b );
    double AQ[3][3];  // A*Q with A symmetric: row r uses A[min][max]
    for ( int r = 0; r < 3; ++r )
      for ( int c = 0; c < 3; ++c ) {
        const double a0 = A[std::min( r, 0 )][std::max( r, 0 )], a1 = A[std::min( r, 1 )][std::max( r, 1 )],
                     a2 = A[std::min( r, 2 )][std::max( r, 2 )];
        AQ[r][c]        = Q[0][c] * a0 + Q[1][c] * a1 + Q[2][c] * a2;
      }
    for ( int r = 0; r < 3; ++r )
      for ( int c = 0; c < 3; ++c ) D[r][c] = AQ[0][r] * Q[0][c] + AQ[1][r] * Q[1][c] + AQ[2][r] * Q[2][c];
    const double o[3] = {D[1][2], D[0][2], D[0][1]};
    const double m[3] = {std::fabs( o[0] ), std::fabs( o[1] ), std::fabs( o[2] )};
    const int    k0   = ( m[0] > m[1] && m[0] > m[2] ) ? 0 : ( m[1] > m[2] ) ? 1 : 2;
    const int    k1 = ( k0 + 1 ) % 3, k2 = ( k0 + 2 ) % 3;
    if ( o[k0] == 0.0 ) break;
    double       thet = ( D[k2][k2] - D[k1][k1] ) / ( 2.0 * o[k0] );
    const double sgn  = ( thet > 0.0 ) ? 1.0 : -1.0;
    thet *= sgn;
    const double t = sgn / ( thet + ( ( thet < 1.E6 ) ? std::sqrt( thet * thet + 1.0 ) : thet ) );
    const double c = 1.0 / std::sqrt( t * t + 1.0 );
    if ( c == 1.0 ) break;
    double jr[4] = {0, 0, 0, 0};
    jr[k0]       = sgn * std::sqrt( ( 1.0 - c ) / 2.0 );
    jr[k0] *= -1.0;
    jr[3] = std::sqrt( 1.0 - jr[k0] * jr[k0] );
    if ( jr[3] == 1.0 ) break;
    // in-place quaternion product: later components see the already-updated earlier ones
    q[0] = ( q[3] * jr[0] + q[0] * jr[3] + q[1] * jr[2] - q[2] * jr[1] );
    q[1] = ( q[3] * jr[1] - q[0] * jr[2] + q[1] * jr[3] + q[2] * jr[0] );
    q[2] = ( q[3] * jr[2] + q[0] * jr[1] - q[1] * jr[0] + q[2] * jr[3] );
    q[3] = ( q[3] * jr[3] - q[0] * jr[0] - q[1] * jr[1] - q[2] * jr[2] );
    const double mq = std::sqrt( q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3] );
    for ( int i = 0; i < 4; ++i ) q[i] /= mq;
  }
}

// L/PccLibEncoder/source/PCCNormalsGenerator.cpp:71-162 (computeNormal), view point = origin.
void normalOfPoint( const int16_t* xyz, size_t i, const uint32_t* nb, int cnt, double* out ) {
  double nrm[3] = {0, 0, 0};
  if ( cnt > 1 ) {
    double bary[3] = {0, 0, 0};
    for ( int j = 0; j < cnt; ++j )
      for ( int d = 0; d < 3; ++d ) bary[d] = bary[d] + double( xyz[3 * size_t( nb[j] ) + d] );
    for ( int d = 0; d < 3; ++d ) bary[d] /= double( cnt );
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for ( int j = 0; j < cnt; ++j ) {
      double p[3];
      for ( int d = 0; d < 3; ++d ) p[d] = double( xyz[3 * size_t( nb[j] ) + d] ) - bary[d];
      C[0][0] += p[0] * p[0], C[1][1] += p[1] * p[1], C[2][2] += p[2] * p[2];
      C[0][1] += p[0] * p[1], C[0][2] += p[0] * p[2], C[1][2] += p[1] * p[2];
    }
    C[1][0] = C[0][1], C[2][0] = C[0][2], C[2][1] = C[1][2];
    const double den = double( cnt ) - 1.0;
    for ( auto& row : C )
      for ( double& v : row ) v /= den;
    double Q[3][3], D[3][3];
    diagonalize( C, Q, D );
    const double e0 = std::fabs( D[0][0] ), e1 = std::fabs( D[1][1] ), e2 = std::fabs( D[2][2] );
    const int    col = ( e0 < e1 && e0 < e2 ) ? 0 : ( e1 < e2 ) ? 1 : 2;
    for ( int d = 0; d < 3; ++d ) nrm[d] = Q[d][col];
  }
  // flip towards the view point (0,0,0): normal * (viewPoint - point) < 0
  const double vp[3] = {0.0 - double( xyz[3 * i] ), 0.0 - double( xyz[3 * i + 1] ), 0.0 - double( xyz[3 * i + 2] )};
  const bool   flip  = dot3( nrm, vp ) < 0.0;
  for ( int d = 0; d < 3; ++d ) out[d] = flip ? -nrm[d] : nrm[d];
}

int rowCount( const uint32_t* row, int k ) {
  int c = 0;
  while ( c < k && row[c] != 0xFFFFFFFFu ) ++c;
  return c;
}

}  // namespace

// ======================================================================================================
extern "C" {

void* pcco_kdtree_build( const int16_t* xyz, size_t n ) {
  Tree* t = new Tree();
  t->build( xyz, n );
  return t;
}
void pcco_kdtree_free( void* tree ) { delete static_cast<Tree*>( tree ); }
void pcco_kdtree_vind( void* tree, uint32_t* vind ) {
  Tree* t = static_cast<Tree*>( tree );
  std::copy( t->vind.begin(), t->vind.end(), vind );
}

void pcco_knn( void* tree, const int16_t* q, size_t nq, int k, uint32_t* idx, float* dist2 ) {
  const Tree*           t = static_cast<Tree*>( tree );
  std::vector<double>   d( k );
  std::vector<uint32_t> id( k );
  for ( size_t i = 0; i < nq; ++i ) {
    KnnSet rs( k, id.data(), d.data() );
    t->search( rs, q + 3 * i );
    for ( int j = 0; j < k; ++j ) {
      idx[i * k + j]   = j < rs.cnt ? id[j] : 0xFFFFFFFFu;
      dist2[i * k + j] = j < rs.cnt ? float( d[j] ) : -1.0f;
    }
  }
}

size_t pcco_radius( void* tree, const int16_t* q, size_t nq, double radius2, size_t max_results, uint64_t* offsets,
                    uint32_t* idx, float* dist2 ) {
  const Tree*                              t = static_cast<Tree*>( tree );
  std::vector<std::pair<double, uint32_t>> found;
  size_t                                   total = 0;
  for ( size_t i = 0; i < nq; ++i ) {
    found.clear();
    RadiusSet rs( radius2, found );
    t->search( rs, q + 3 * i );
    std::sort( found.begin(), found.end() );  // (dist, index): NF:160-166 IndexDist_Sorter is a total order
    if ( found.size() > max_results ) found.resize( max_results );
    offsets[i] = total;
    if ( idx )
      for ( size_t j = 0; j < found.size(); ++j ) {
        idx[total + j] = found[j].second;
        if ( dist2 ) dist2[total + j] = float( found[j].first );
      }
    total += found.size();
  }
  offsets[nq] = total;
  return total;
}

void pcco_normals( const int16_t* xyz, size_t n, const uint32_t* nbr, int k, double* normals ) {
  for ( size_t i = 0; i < n; ++i ) normalOfPoint( xyz, i, nbr + i * k, rowCount( nbr + i * k, k ), normals + 3 * i );
}

// L/PccLibEncoder/source/PCCNormalsGenerator.cpp:198-242 (orientNormals, SPANNING_TREE) + :521-548 (addNeighbors)
// Edge order: weight, then start index, then end index (PCCNormalsGenerator.h:64-72); max-heap.
void pcco_orient_normals( const int16_t* xyz, size_t n, const uint32_t* nbr, int k, double* normals ) {
  struct Edge {
    double   w;
    uint32_t s, e;
    bool     operator<( const Edge& o ) const {
      if ( w == o.w ) return s == o.s ? e < o.e : s < o.s;
      return w < o.w;
    }
  };
  std::priority_queue<Edge> heap;
  std::vector<uint8_t>      visited( n, 0 );
  double                    acc[3];
  size_t                    accCount = 0;
  auto                      expand   = [&]( uint32_t cur ) {
    acc[0] = acc[1] = acc[2] = 0.0;
    accCount                 = 0;
    const uint32_t* row      = nbr + size_t( cur ) * k;
    const int       cnt      = rowCount( row, k );
    for ( int j = 0; j < cnt; ++j ) {
      const uint32_t o = row[j];
      if ( !visited[o] ) {
        heap.push( Edge{std::fabs( dot3( normals + 3 * size_t( cur ), normals + 3 * size_t( o ) ) ), cur, o} );
      } else if ( o != cur ) {
        for ( int d = 0; d < 3; ++d ) acc[d] = acc[d] + normals[3 * size_t( o ) + d];
        ++accCount;
      }
    }
  };
  for ( size_t seed = 0; seed < n; ++seed ) {
    if ( visited[seed] ) continue;
    visited[seed] = 1;
    expand( uint32_t( seed ) );
    if ( accCount == 0 ) {
      if ( seed != 0 ) {
        for ( int d = 0; d < 3; ++d ) acc[d] = normals[3 * ( seed - 1 ) + d];
      } else {
        for ( int d = 0; d < 3; ++d ) acc[d] = 0.0 - double( xyz[3 * seed + d] );
      }
    }
    if ( dot3( normals + 3 * seed, acc ) < 0.0 )
      for ( int d = 0; d < 3; ++d ) normals[3 * seed + d] = -normals[3 * seed + d];
    while ( !heap.empty() ) {
      const Edge e = heap.top();
      heap.pop();
      if ( visited[e.e] ) continue;
      visited[e.e] = 1;
      if ( dot3( normals + 3 * size_t( e.s ), normals + 3 * size_t( e.e ) ) < 0.0 )
        for ( int d = 0; d < 3; ++d ) normals[3 * size_t( e.e ) + d] = -normals[3 * size_t( e.e ) + d];
      expand( e.e );
    }
  }
  size_t neg = 0;
  for ( size_t i = 0; i < n; ++i ) {
    const double vp[3] = {0.0 - double( xyz[3 * i] ), 0.0 - double( xyz[3 * i + 1] ), 0.0 - double( xyz[3 * i + 2] )};
    neg += dot3( normals + 3 * i, vp ) < 0.0;
  }
  if ( neg > ( n + 1 ) / 2 )
    for ( size_t i = 0; i < 3 * n; ++i ) normals[i] = -normals[i];
}

// L/PccLibEncoder/source/PCCEncoder.cpp:3569-3626 (calculateWeightNormal, enhancedPP on)
void pcco_weight_normal( const int16_t* xyz, size_t n, int bits, double minW, double w[3] ) {
  const size_t         side = size_t( 1 ) << bits;
  std::vector<uint8_t> face( 3 * side * side, 0 );
  for ( size_t i = 0; i < n; ++i ) {
    int p[3];
    for ( int d = 0; d < 3; ++d ) p[d] = std::max( 0, std::min( int( side - 1 ), int( xyz[3 * i + d] ) ) );
    face[size_t( p[2] ) * side + p[1]]                   = 1;  // YZ plane  -> axis 0
    face[size_t( p[0] ) * side + p[2] + side * side]     = 1;  // ZX plane  -> axis 1
    face[size_t( p[1] ) * side + p[0] + 2 * side * side] = 1;  // XY plane  -> axis 2
  }
  struct Cnt {
    int idx, value;
  } c[3];
  for ( int a = 0; a < 3; ++a ) {
    c[a].idx = a, c[a].value = 0;
    for ( size_t i = 0; i < side * side; ++i ) c[a].value += face[a * side * side + i];
  }
  // std::sort with comp1 (ascending by value); 3 elements -> insertion sort, stable for ties
  std::stable_sort( c, c + 3, []( const Cnt& a, const Cnt& b ) { return a.value < b.value; } );
  double ax[3];
  const double r0 = double( c[0].value ) / double( c[2].value ), r1 = double( c[1].value ) / double( c[2].value );
  if ( r0 >= minW ) {
    ax[c[0].idx] = r0, ax[c[1].idx] = r1, ax[c[2].idx] = 1.0;
  } else {
    ax[c[0].idx] = minW, ax[c[2].idx] = 1.0;
    ax[c[1].idx] = minW + ( r1 - r0 ) / ( 1.0 - r0 ) * ( 1 - minW );
  }
  w[0] = ax[0], w[1] = ax[1], w[2] = ax[2];
}

// L/PccLibEncoder/source/PCCPatchSegmenter.cpp:226-265: orientation 0 is scored WITHOUT its axis weight.
void pcco_initial_segmentation( const double* normals, size_t n, const double w[3], uint8_t* partition ) {
  static const double O[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
  const double        wt[6]   = {w[0], w[1], w[2], w[0], w[1], w[2]};
  for ( size_t i = 0; i < n; ++i ) {
    const double* nr   = normals + 3 * i;
    int           best = 0;
    double        bs   = dot3( nr, O[0] );
    for ( int j = 1; j < 6; ++j ) {
      const double s = dot3( nr, O[j] ) * wt[j];
      if ( s > bs ) bs = s, best = j;
    }
    partition[i] = uint8_t( best );
  }
}


// ======================================================================================================
// a6. Grid-based refinement (L/PccLibEncoder/source/PCCPatchSegmenter.cpp:1386-1561; voxel classes
//     L/PccLibEncoder/include/PCCPatchSegmenter.h:430-510).
// ======================================================================================================
namespace {
enum : uint8_t { NO_EDGE = 0x00, INDIRECT_EDGE = 0x01, M_DIRECT_EDGE = 0x10, S_DIRECT_EDGE = 0x11 };

struct Voxel {
  std::vector<uint32_t> pts;
  uint16_t              score[6] = {0, 0, 0, 0, 0, 0};
  uint8_t               edge = 0, ppi = 0, dirty = 1;

  void recount( const uint8_t* partition ) {
    for ( auto& s : score ) s = 0;
    for ( uint32_t j : pts ) ++score[partition[j]];
    if ( !dirty ) return;  // class and PPI are only re-derived for voxels whose points were re-labelled
    if ( edge != S_DIRECT_EDGE ) {
      int used = 0;
      for ( auto s : score ) used += s != 0;
      edge = used == 1 ? NO_EDGE : M_DIRECT_EDGE;
    }
    ppi   = uint8_t( std::max_element( score, score + 6 ) - score );
    dirty = 0;
  }
};
}  // namespace

void pcco_refine_segmentation( const int16_t* xyz, const double* normals, size_t n, const pccb200_seg_params* prm,
                               uint8_t* partition ) {
  if ( n == 0 ) return;
  static const double O[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
  const size_t        voxDim  = size_t( prm->voxel_dim_refine );
  int16_t             geoMax  = xyz[0];
  for ( size_t i = 0; i < 3 * n; ++i ) geoMax = std::max( geoMax, xyz[i] );
  size_t geoRange = 1;
  for ( size_t i = size_t( geoMax - 1 ); i != 0u; i >>= 1, geoRange <<= 1 ) {}
  size_t voxShift = 0, gridShift = 0;
  for ( size_t i = voxDim; i > 1; ++voxShift, i >>= 1 ) {}
  const size_t gridDim = geoRange >> voxShift;
  for ( size_t i = gridDim; i > 1; ++gridShift, i >>= 1 ) {}
  const size_t half = voxDim >> 1;
  auto         key  = [&]( size_t x, size_t y, size_t z ) { return x + ( y << gridShift ) + ( z << ( 2 * gridShift ) ); };

  // voxels in order of first appearance; a voxel is identified by its key (keys of border voxels may alias)
  std::vector<Voxel>                 vox;
  std::vector<int16_t>               centers;
  std::unordered_map<size_t, size_t> index;
  for ( size_t i = 0; i < n; ++i ) {
    const size_t x = ( size_t( xyz[3 * i] ) + half ) >> voxShift, y = ( size_t( xyz[3 * i + 1] ) + half ) >> voxShift,
                 z = ( size_t( xyz[3 * i + 2] ) + half ) >> voxShift;
    const size_t k  = key( x, y, z );
    auto         it = index.find( k );
    if ( it == index.end() ) {
      it = index.emplace( k, vox.size() ).first;
      vox.emplace_back();
      centers.push_back( int16_t( x ) ), centers.push_back( int16_t( y ) ), centers.push_back( int16_t( z ) );
    }
    vox[it->second].pts.push_back( uint32_t( i ) );
  }
  const size_t V = vox.size();
  // the reference re-looks each voxel up by the key of its centre: aliasing keys resolve to the same cell
  std::vector<size_t> cell( V );
  for ( size_t v = 0; v < V; ++v ) cell[v] = index[key( size_t( centers[3 * v] ), size_t( centers[3 * v + 1] ), size_t( centers[3 * v + 2] ) )];
  for ( size_t v = 0; v < V; ++v ) {
    Voxel& c = vox[cell[v]];
    c.edge   = uint8_t( c.pts.size() ) == 1 ? S_DIRECT_EDGE : M_DIRECT_EDGE;
    c.recount( partition );
  }
  // adjacency: voxels with centre distance^2 < (searchRadius >> shift), sorted by (distance, index)
  Tree tree;
  tree.build( centers.data(), V );
  const double                       r2 = double( size_t( prm->search_radius_refine ) >> voxShift );
  std::vector<std::vector<uint32_t>> adj( V ), near( V );
  std::vector<double>                weight( V );
  const size_t                       nearRange = voxDim >= 4 ? 1 : 2;
  std::vector<std::pair<double, uint32_t>> found;
  for ( size_t v = 0; v < V; ++v ) {
    found.clear();
    RadiusSet rs( r2, found );
    tree.search( rs, &centers[3 * v] );
    std::sort( found.begin(), found.end() );
    if ( found.size() > 32767 ) found.resize( 32767 );
    size_t nn = 0;
    for ( auto& f : found ) {
      const uint32_t o = f.second;
      adj[v].push_back( o );
      if ( size_t( std::abs( centers[3 * v] - centers[3 * o] ) ) <= nearRange &&
           size_t( std::abs( centers[3 * v + 1] - centers[3 * o + 1] ) ) <= nearRange &&
           size_t( std::abs( centers[3 * v + 2] - centers[3 * o + 2] ) ) <= nearRange )
        near[v].push_back( o );
      nn += uint8_t( vox[cell[o]].pts.size() );
      if ( nn >= size_t( prm->max_nn_count_refine ) ) break;
    }
    weight[v] = prm->lambda_refine / double( nn );
  }
  for ( int iter = 0; iter < std::max( 1, prm->iteration_count_refine ); ++iter ) {
    for ( size_t v = 0; v < V; ++v ) {
      Voxel&        c        = vox[cell[v]];
      const uint8_t edgeHere = c.edge;
      if ( edgeHere == NO_EDGE ) continue;
      uint16_t smooth[6] = {0, 0, 0, 0, 0, 0};
      for ( uint32_t o : adj[v] )
        for ( int k = 0; k < 6; ++k ) smooth[k] = uint16_t( smooth[k] + vox[cell[o]].score[k] );
      const size_t top = size_t( std::max_element( smooth, smooth + 6 ) - smooth );
      for ( uint32_t o : near[v] ) {
        Voxel& d = vox[cell[o]];
        if ( d.edge == NO_EDGE && d.ppi != top ) d.edge = INDIRECT_EDGE;
      }
      if ( edgeHere != M_DIRECT_EDGE ) {
        int used = 0;
        for ( auto s : smooth ) used += s != 0;
        if ( used == 1 && smooth[c.ppi] > 0 ) continue;
      }
      for ( uint32_t j : c.pts ) {
        double sc[6];
        for ( int k = 0; k < 6; ++k ) sc[k] = dot3( normals + 3 * size_t( j ), O[k] ) + weight[v] * double( smooth[k] );
        partition[j] = uint8_t( std::max_element( sc, sc + 6 ) - sc );
      }
      c.dirty = 1;
    }
    for ( size_t v = 0; v < V; ++v ) vox[cell[v]].recount( partition );
  }
}

// ======================================================================================================
// a7–a11. Patch segmentation (L/PccLibEncoder/source/PCCPatchSegmenter.cpp:537-1320, CTC path: no EOM,
//         6 projection planes, no partitioning/expansion/gradient separation), resampling (:362-470).
// ======================================================================================================
namespace {
// GPAPatchData (L/PccLibCommon/include/PCCPatch.h:42-71): a trial placement of the global patch allocation
struct OGpa {
  bool                 matched = false, global = false;
  int                  track = -1, sizeU0 = 0, sizeV0 = 0, u0 = -1, v0 = -1, orient = -1;
  std::vector<uint8_t> occ;
  void                 reset() { *this = OGpa(); }
};
struct OPatch {
  pccb200_patch        m;
  std::vector<int16_t> depth[2];
  std::vector<uint8_t> occ;
  OGpa                 cur, pre;  // random-access packing only
};
struct OPatchList {
  std::vector<OPatch> patches;
};
const int kViewAxes[6][4] = {{0, 2, 1, 0}, {1, 2, 0, 0}, {2, 0, 1, 0}, {0, 2, 1, 1}, {1, 2, 0, 1}, {2, 0, 1, 1}};  // normal,tangent,bitangent,mode
const int16_t kInfDepth   = 32767;
}  // namespace

void* pcco_segment_patches( const int16_t* xyz, const uint8_t* rgb, size_t n, const uint32_t* nbr, int k,
                            const uint8_t* partition, const pccb200_seg_params* prm ) {
  OPatchList*          out = new OPatchList();
  std::vector<double>  rawDist( n, std::numeric_limits<double>::max() );
  std::vector<uint32_t> raw( n );
  for ( size_t i = 0; i < n; ++i ) raw[i] = uint32_t( i );
  std::vector<int16_t> resampled;  // x,y,z triples
  const int            occRes = prm->occupancy_resolution;
  const int            minLevel = prm->min_level, thickness = prm->surface_thickness;
  while ( !raw.empty() ) {
    // ---- connected components over the directed k-NN graph, seeds in ascending index among far raw points
    std::vector<uint8_t> flag( n, 0 );
    for ( uint32_t i : raw ) flag[i] = 1;
    std::vector<std::vector<uint32_t>> comps;
    std::vector<uint32_t>              stack;
    for ( uint32_t i : raw ) {
      if ( !flag[i] || !( rawDist[i] > prm->max_allowed_dist2_raw_detection ) ) continue;
      flag[i] = 0;
      comps.emplace_back();
      auto&         cc = comps.back();
      const uint8_t cl = partition[i];
      stack.push_back( i ), cc.push_back( i );
      while ( !stack.empty() ) {
        const uint32_t cur = stack.back();
        stack.pop_back();
        const uint32_t* row = nbr + size_t( cur ) * k;
        for ( int j = 0; j < k && row[j] != 0xFFFFFFFFu; ++j ) {
          const uint32_t o = row[j];
          if ( partition[o] == cl && flag[o] ) flag[o] = 0, stack.push_back( o ), cc.push_back( o );
        }
      }
      if ( cc.size() < size_t( prm->min_point_count_per_cc ) ) comps.pop_back();
    }
    if ( comps.empty() ) break;
    for ( auto& cc : comps ) {
      out->patches.emplace_back();
      OPatch&        P = out->patches.back();
      pccb200_patch& m = P.m;
      std::memset( &m, 0, sizeof( m ) );
      m.index   = int32_t( out->patches.size() - 1 );
      m.view_id = partition[cc[0]];
      m.normal_axis = kViewAxes[m.view_id][0], m.tangent_axis = kViewAxes[m.view_id][1];
      m.bitangent_axis = kViewAxes[m.view_id][2], m.projection_mode = kViewAxes[m.view_id][3];
      m.u0 = m.v0 = m.orientation = 0;  // PCCPatch defaults before packing
      m.best_match_idx = -1;
      const int na = m.normal_axis, ta = m.tangent_axis, ba = m.bitangent_axis, mode = m.projection_mode;
      const int dir = 1 - 2 * mode;
      auto      C   = [&]( uint32_t i, int axis ) { return int( xyz[3 * size_t( i ) + axis] ); };
      if ( prm->enable_patch_splitting ) {  // keep the part within maxPatchSize of the minimum corner
        int minU = 32767, minV = 32767;
        for ( uint32_t i : cc ) minU = std::min( minU, C( i, ta ) ), minV = std::min( minV, C( i, ba ) );
        std::vector<uint32_t> kept;
        for ( uint32_t i : cc )
          if ( C( i, ta ) - minU < prm->max_patch_size && C( i, ba ) - minV < prm->max_patch_size ) kept.push_back( i );
        cc.swap( kept );
        if ( cc.empty() ) continue;  // (the reference leaves an empty patch in the list here as well)
      }
      int mn[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, mx[3] = {0, 0, 0};
      for ( uint32_t i : cc )
        for ( int a = 0; a < 3; ++a ) mn[a] = std::min( mn[a], C( i, a ) ), mx[a] = std::max( mx[a], C( i, a ) );
      m.size_u = 1 + mx[ta] - mn[ta], m.size_v = 1 + mx[ba] - mn[ba];
      m.u1 = mn[ta], m.v1 = mn[ba];
      const size_t px = size_t( m.size_u ) * m.size_v;
      P.depth[0].assign( px, kInfDepth );
      std::vector<uint32_t> owner( px, 0xFFFFFFFFu );
      int                   maxU = 0, maxV = 0, extreme = mode == 0 ? kInfDepth : 0;
      for ( uint32_t i : cc ) {  // depth0 = nearest (mode 0) / farthest (mode 1) point per pixel
        const int    d = C( i, na ), u = C( i, ta ) - m.u1, v = C( i, ba ) - m.v1;
        const size_t p = size_t( v ) * m.size_u + u;
        const bool   better = mode == 0 ? P.depth[0][p] > d : ( P.depth[0][p] == kInfDepth || P.depth[0][p] < d );
        if ( !better ) continue;
        P.depth[0][p] = int16_t( d ), owner[p] = i;
        maxU = std::max( maxU, u ), maxV = std::max( maxV, v );
        extreme = mode == 0 ? std::min( extreme, d ) : std::max( extreme, d );
      }
      m.d1 = mode == 0 ? ( extreme / minLevel ) * minLevel : int( std::ceil( double( extreme ) / double( minLevel ) ) ) * minLevel;
      m.size_u0 = maxU / occRes + 1, m.size_v0 = maxV / occRes + 1;
      m.size_2d_x = int( std::ceil( double( maxU + 1 ) / double( prm->quantizer_size_x ) ) * prm->quantizer_size_x );
      m.size_2d_y = int( std::ceil( double( maxV + 1 ) / double( prm->quantizer_size_y ) ) * prm->quantizer_size_y );
      P.occ.assign( size_t( m.size_u0 ) * m.size_v0, 0 );
      // per-block peak filter
      std::vector<int16_t> peak( P.occ.size(), mode == 0 ? kInfDepth : int16_t( 0 ) );
      for ( int v = 0; v < m.size_v; ++v )
        for ( int u = 0; u < m.size_u; ++u ) {
          const int16_t d = P.depth[0][size_t( v ) * m.size_u + u];
          if ( d == kInfDepth ) continue;
          int16_t& pk = peak[size_t( v / occRes ) * m.size_u0 + u / occRes];
          pk          = mode == 0 ? std::min( pk, d ) : std::max( pk, d );
        }
      for ( int v = 0; v < m.size_v; ++v )
        for ( int u = 0; u < m.size_u; ++u ) {
          const size_t  p = size_t( v ) * m.size_u + u;
          const int16_t d = P.depth[0][p];
          if ( d == kInfDepth ) continue;
          const int16_t pk = peak[size_t( v / occRes ) * m.size_u0 + u / occRes];
          const int16_t a  = int16_t( std::abs( d - pk ) );
          const int16_t b  = int16_t( thickness + dir * d );
          const int16_t c  = int16_t( dir * m.d1 + prm->max_allowed_depth );
          if ( a > 32 || b > c ) P.depth[0][p] = kInfDepth, owner[p] = 0xFFFFFFFFu;
        }
      // depth1: farthest point within surfaceThickness of depth0 whose colour is close to the depth0 point's
      P.depth[1] = P.depth[0];
      if ( thickness > 0 )
        for ( uint32_t i : cc ) {
          const int     d = C( i, na ), u = C( i, ta ) - m.u1, v = C( i, ba ) - m.v1;
          const size_t  p  = size_t( v ) * m.size_u + u;
          const int16_t d0 = P.depth[0][p];
          if ( !( d0 < kInfDepth ) ) continue;
          const int16_t  delta = int16_t( dir * ( d - d0 ) );
          const uint8_t *ci = rgb + 3 * size_t( i ), *c0 = rgb + 3 * size_t( owner[p] );
          const bool similar = std::abs( int( c0[0] ) - ci[0] ) < 128 && std::abs( int( c0[1] ) - ci[1] ) < 128 &&
                               std::abs( int( c0[2] ) - ci[2] ) < 128;
          if ( delta <= thickness && delta >= 0 && similar && dir * ( d - P.depth[1][p] ) > 0 ) P.depth[1][p] = int16_t( d );
        }
      // resample: D0 then D1 point per occupied pixel (raster order), depths re-based to d1
      int sizeD = 0, d0Count = 0;
      for ( int v = 0; v < m.size_v; ++v )
        for ( int u = 0; u < m.size_u; ++u ) {
          const size_t p = size_t( v ) * m.size_u + u;
          if ( !( P.depth[0][p] < kInfDepth ) ) continue;
          P.occ[size_t( v / occRes ) * m.size_u0 + u / occRes] = 1;
          for ( int map = 0; map < 2; ++map ) {
            int16_t q[3];
            q[na] = P.depth[map][p], q[ta] = int16_t( u + m.u1 ), q[ba] = int16_t( v + m.v1 );
            resampled.insert( resampled.end(), q, q + 3 );
          }
          ++d0Count;
          for ( int map = 0; map < 2; ++map ) {
            P.depth[map][p] = int16_t( dir * ( P.depth[map][p] - int16_t( m.d1 ) ) );
            sizeD           = std::max( sizeD, int( P.depth[map][p] ) );
          }
        }
      m.d0_count     = d0Count;
      m.size_d_pixel = sizeD;
      const int bd   = std::min( prm->geometry_bitdepth_3d, prm->geometry_bitdepth_2d );
      sizeD          = std::min( ( 1 << bd ) - 1, sizeD );
      const int lv   = int( std::log2( double( minLevel ) ) );
      int       qd   = sizeD == 0 ? 0 : ( ( sizeD - 1 ) / minLevel + 1 );
      qd             = std::min( qd, ( 1 << ( bd - lv ) ) - 1 );
      m.size_d       = qd == 0 ? 0 : qd * minLevel - 1;
    }
    // ---- residual: points farther than sqrt(selection) from everything resampled so far stay raw
    Tree rt;
    rt.build( resampled.data(), resampled.size() / 3 );
    raw.clear();
    for ( size_t i = 0; i < n; ++i ) {
      uint32_t id;
      double   d;
      KnnSet   rs( 1, &id, &d );
      rt.search( rs, xyz + 3 * i );
      rawDist[i] = d;
      if ( d > prm->max_allowed_dist2_raw_selection ) raw.push_back( uint32_t( i ) );
    }
  }
  return out;
}

int    pcco_patches_count( void* h ) { return int( static_cast<OPatchList*>( h )->patches.size() ); }
size_t pcco_patches_depth_elems( void* h ) {
  size_t s = 0;
  for ( auto& p : static_cast<OPatchList*>( h )->patches ) s += 2 * size_t( p.m.size_u ) * p.m.size_v;
  return s;
}
size_t pcco_patches_occ_elems( void* h ) {
  size_t s = 0;
  for ( auto& p : static_cast<OPatchList*>( h )->patches ) s += size_t( p.m.size_u0 ) * p.m.size_v0;
  return s;
}
void pcco_patches_get( void* h, pccb200_patch* out, int16_t* depth, uint8_t* occ ) {
  int64_t dOff = 0, oOff = 0;
  size_t  i    = 0;
  for ( auto& p : static_cast<OPatchList*>( h )->patches ) {
    out[i]              = p.m;
    out[i].depth_offset = dOff, out[i].occ_offset = oOff;
    ++i;
    const size_t px = size_t( p.m.size_u ) * p.m.size_v;
    for ( int m = 0; m < 2; ++m )
      for ( size_t j = 0; j < px; ++j ) depth[dOff + m * px + j] = j < p.depth[m].size() ? p.depth[m][j] : int16_t( 0 );
    dOff += 2 * px;
    std::copy( p.occ.begin(), p.occ.end(), occ + oOff );
    oOff += p.occ.size();
  }
}
void pcco_patches_free( void* h ) { delete static_cast<OPatchList*>( h ); }

void* pcco_segment_frame( const int16_t* xyz, const uint8_t* rgb, size_t n, const pccb200_seg_params* p ) {
  if ( n == 0 ) return new OPatchList();
  const int             k = p->nn_normal_estimation;
  Tree                  t;
  t.build( xyz, n );
  std::vector<uint32_t> nbr( n * k );
  std::vector<float>    d( n * k );
  pcco_knn( &t, xyz, n, k, nbr.data(), d.data() );
  std::vector<double> normals( 3 * n );
  pcco_normals( xyz, n, nbr.data(), k, normals.data() );
  if ( p->normal_orientation == 1 ) pcco_orient_normals( xyz, n, nbr.data(), k, normals.data() );
  std::vector<uint8_t> part( n );
  pcco_initial_segmentation( normals.data(), n, p->weight_normal, part.data() );
  pcco_refine_segmentation( xyz, normals.data(), n, p, part.data() );
  return pcco_segment_patches( xyz, rgb, n, nbr.data(), p->max_nn_count_patch_seg, part.data(), p );
}


// ======================================================================================================
// a13–a26. Packing, occupancy / geometry / attribute image formation, reconstruction, colour transfer, padding
// for one GOF (CTC all-intra path), in the order PCCEncoder::encode runs them (L/PccLibEncoder/source/PCCEncoder.cpp
// :103-424) with the video codec treated as lossless (decoded == source).
// ======================================================================================================
namespace {

struct OFrame {
  OPatchList            pl;  // packed order
  size_t                width = 0, height = 0;
  std::vector<uint8_t>  occupancy, omVideo, recRgb;
  std::vector<uint32_t> blockToPatch, pointToPixel, recPartition;
  std::vector<uint16_t> geo[2], recBoundary, attrRaw[2], attr[2];
  std::vector<uint8_t>  attrYuv[2];  // 8-bit YUV 4:2:0 of the padded attribute frames (Y, U, V planes)
  std::vector<int16_t>  recXyz;
};
struct OGof {
  std::vector<OFrame> frames;
};

enum { OR_DEFAULT = 0, OR_SWAP = 1 };  // PATCH_ORIENTATION_DEFAULT / _SWAP (PCCBitstreamCommon.h:112-122)  // PCCPatchOrientation values used by the 2-orientation flexible packer

// PCCPatch::patchBlock2CanvasBlock / patch2Canvas for the two orientations in use (L/PccLibCommon/source/PCCPatch.cpp:192-308)
inline void blockToCanvas( const pccb200_patch& m, int ub, int vb, int& x, int& y ) {
  if ( m.orientation == OR_DEFAULT ) x = ub + m.u0, y = vb + m.v0;
  else x = vb + m.u0, y = ub + m.v0;
}
inline void pixelToCanvas( const pccb200_patch& m, int occRes, int u, int v, int& x, int& y ) {
  if ( m.orientation == OR_DEFAULT ) x = u + m.u0 * occRes, y = v + m.v0 * occRes;
  else x = v + m.u0 * occRes, y = u + m.v0 * occRes;
}

// PCCPatch::gt (PCCPatch.cpp:349-371): larger max dimension first, then larger min dimension, then lower index
bool patchBefore( const OPatch& a, const OPatch& b ) {
  const int amax = std::max( a.m.size_u0, a.m.size_v0 ), amin = std::min( a.m.size_u0, a.m.size_v0 );
  const int bmax = std::max( b.m.size_u0, b.m.size_v0 ), bmin = std::min( b.m.size_u0, b.m.size_v0 );
  return amax != bmax ? amax > bmax : ( amin != bmin ? amin > bmin : a.m.index < b.m.index );
}

// PCCEncoder::packFlexible (PCCEncoder.cpp:2306-2449), packingStrategy 1, two orientations, safeguard 0, lowDelay off
void packFrame( OFrame& F, int occRes, int presetWidth, int numTilesHor, double tileRatio ) {
  auto& P  = F.pl.patches;
  F.width  = presetWidth;
  if ( P.empty() ) return;  // height stays at its preset
  std::sort( P.begin(), P.end(), patchBefore );  // gt() is a strict total order: any sort gives this result
  int sizeU = presetWidth / occRes, sizeV = std::max( P[0].m.size_v0, P[0].m.size_u0 );
  for ( auto& p : P ) sizeU = std::max( sizeU, p.m.size_u0 + 1 );
  const int tileW = sizeU / numTilesHor, tileH = int( tileW * tileRatio );
  sizeV           = sizeV >= tileH ? sizeV : tileH;
  size_t height   = size_t( sizeV ) * occRes;
  std::vector<uint8_t> canvas( size_t( sizeU ) * sizeV, 0 );
  for ( auto& p : P ) {
    pccb200_patch& m = p.m;
    bool           found = false;
    while ( !found ) {
      for ( int v = 0; v < sizeV && !found; ++v )
        for ( int u = 0; u < sizeU && !found; ++u ) {
          m.u0 = u, m.v0 = v;
          for ( int o = 0; o < 2 && !found; ++o ) {
            m.orientation = ( m.size_u0 > m.size_v0 ) ? ( o == 0 ? OR_SWAP : OR_DEFAULT ) : ( o == 0 ? OR_DEFAULT : OR_SWAP );
            bool fits     = true;  // every block of the bounding box must lie on the canvas and be free
            for ( int vb = 0; vb < m.size_v0 && fits; ++vb )
              for ( int ub = 0; ub < m.size_u0 && fits; ++ub ) {
                int x, y;
                blockToCanvas( m, ub, vb, x, y );
                if ( x >= sizeU || y >= sizeV || canvas[size_t( y ) * sizeU + x] ) fits = false;
              }
            found = fits;
          }
        }
      if ( !found ) {
        sizeV *= 2;
        canvas.resize( size_t( sizeU ) * sizeV, 0 );
      }
    }
    for ( int vb = 0; vb < m.size_v0; ++vb )  // only occupied blocks are taken
      for ( int ub = 0; ub < m.size_u0; ++ub ) {
        int x, y;
        blockToCanvas( m, ub, vb, x, y );
        canvas[size_t( y ) * sizeU + x] |= p.occ[size_t( vb ) * m.size_u0 + ub];
      }
    const int rows = m.orientation == OR_DEFAULT ? m.size_v0 : m.size_u0;
    height         = std::max( height, size_t( m.v0 + rows ) * occRes );
  }
  F.height = height;
}

// =====================================================================================================================
// a15: random-access packing (constrainedPack 1, globalPatchAllocation 1; packingStrategy 1, two orientations, safeguard 0,
// lowDelayEncoding off, one tile, no raw / EOM patches)
// =====================================================================================================================
// canvas of occupancy blocks with the reference's fit / mark rules (PCCPatch::checkFitPatchCanvas, PCCPatch.cpp:310-335; ...ForGPA :667-692)
struct BlockCanvas {
  int                  sizeU = 0, sizeV = 0;
  std::vector<uint8_t> c;
  BlockCanvas( int u, int v ) : sizeU( u ), sizeV( v ), c( size_t( u ) * v, 0 ) {}
  static void at( int u0, int v0, int orient, int ub, int vb, int& x, int& y ) {
    if ( orient == OR_DEFAULT ) x = ub + u0, y = vb + v0;
    else x = vb + u0, y = ub + v0;
  }
  bool fits( int u0, int v0, int orient, int sU0, int sV0 ) const {  // the whole bounding box must be on the canvas and free
    for ( int vb = 0; vb < sV0; ++vb )
      for ( int ub = 0; ub < sU0; ++ub ) {
        int x, y;
        at( u0, v0, orient, ub, vb, x, y );
        if ( x >= sizeU || y >= sizeV || c[size_t( y ) * sizeU + x] ) return false;
      }
    return true;
  }
  void mark( int u0, int v0, int orient, int sU0, int sV0, const std::vector<uint8_t>& occ, int occStride ) {  // occupied blocks only
    for ( int vb = 0; vb < sV0; ++vb )
      for ( int ub = 0; ub < sU0; ++ub ) {
        int x, y;
        at( u0, v0, orient, ub, vb, x, y );
        c[size_t( y ) * sizeU + x] |= occ[size_t( vb ) * occStride + ub];
      }
  }
  void grow() {
    sizeV *= 2;
    c.resize( size_t( sizeU ) * sizeV, 0 );
  }
};
enum PlaceMode { PLACE_BEST_EFFORT, PLACE_MATCHED, PLACE_KNOWN_ORIENTATION, PLACE_STICKY };
// One placement. BEST_EFFORT: raster scan, two orientations per position, the order chosen by the aspect of (aspU0, aspV0)
// (g_orientationHorizontal / g_orientationVertical, PCCCommon.h:131-150). MATCHED: the given position first, then a raster scan
// with the given orientation (its inclusive loop bounds only add positions that can never fit). KNOWN_ORIENTATION: raster scan
// with the given orientation. STICKY (a union patch without a reference orientation while a reference frame is in use,
// PCCEncoder.cpp:7171-7193): the first position tries both orientations; if neither fits, the orientation tried last has
// overwritten the "unknown" marker, so every later position is tried with that one only. The canvas doubles its height until
// the patch fits.
void placeOne( BlockCanvas& cv, int sU0, int sV0, int aspU0, int aspV0, PlaceMode mode, int& u0, int& v0, int& orient ) {
  for ( ;; ) {
    if ( mode == PLACE_MATCHED && cv.fits( u0, v0, orient, sU0, sV0 ) ) return;
    for ( int v = 0; v < cv.sizeV; ++v )
      for ( int u = 0; u < cv.sizeU; ++u ) {
        if ( mode == PLACE_BEST_EFFORT || ( mode == PLACE_STICKY && orient == -1 ) ) {
          for ( int o = 0; o < 2; ++o ) {
            const int tryO = aspU0 > aspV0 ? ( o == 0 ? OR_SWAP : OR_DEFAULT ) : ( o == 0 ? OR_DEFAULT : OR_SWAP );
            if ( cv.fits( u, v, tryO, sU0, sV0 ) ) {
              u0 = u, v0 = v, orient = tryO;
              return;
            }
            if ( mode == PLACE_STICKY ) orient = tryO;
          }
        } else if ( cv.fits( u, v, orient, sU0, sV0 ) ) {
          u0 = u, v0 = v;
          return;
        }
      }
    cv.grow();
  }
}
inline bool switched( int orient ) { return orient != OR_DEFAULT; }  // isPatchDimensionSwitched for DEFAULT / SWAP
inline void extend( int u0, int v0, int orient, int sU0, int sV0, int occRes, size_t& width, size_t& height ) {
  height = std::max( height, size_t( v0 + ( switched( orient ) ? sU0 : sV0 ) ) * occRes );
  width  = std::max( width, size_t( u0 + ( switched( orient ) ? sV0 : sU0 ) ) * occRes );
}
// pcc::computeIOU on the 3-D bounding rectangles (u1, v1, sizeU, sizeV) (PCCPatchSegmenter.cpp:1563-1570, Rect :393-425)
float rectIou( const pccb200_patch& a, const pccb200_patch& b ) {
  const int x1 = std::max( a.u1, b.u1 ), y1 = std::max( a.v1, b.v1 );
  int       w = std::min( a.u1 + a.size_u, b.u1 + b.size_u ) - x1, h = std::min( a.v1 + a.size_v, b.v1 + b.size_v ) - y1;
  if ( w <= 0 || h <= 0 ) w = h = 0;
  const int inter = w * h, uni = a.size_u * a.size_v + b.size_u * b.size_v - inter;
  return static_cast<float>( inter ) / uni;
}

// PCCEncoder::spatialConsistencyPackFlexible (PCCEncoder.cpp:1183-1412)
void packFrameAfter( OFrame& F, const OFrame& prev, int occRes, int presetWidth, int numTilesHor, double tileRatio ) {
  auto& P = F.pl.patches;
  F.width = presetWidth;
  if ( P.empty() ) return;
  std::sort( P.begin(), P.end(), patchBefore );
  int sizeU = presetWidth / occRes, sizeV = std::max( P[0].m.size_u0, P[0].m.size_v0 );
  // every patch of the previous frame, in its order, takes the unmatched patch of the same view with the largest IoU (> 0.2)
  std::vector<OPatch> order;
  for ( size_t pi = 0; pi < prev.pl.patches.size(); ++pi ) {
    const pccb200_patch& q = prev.pl.patches[pi].m;
    float                best = 0.0F;
    int                  bestIdx = -1;
    for ( size_t ci = 0; ci < P.size(); ++ci )
      if ( P[ci].m.view_id == q.view_id && P[ci].m.best_match_idx == -1 ) {
        const float iou = rectIou( q, P[ci].m );
        if ( iou > best ) best = iou, bestIdx = int( ci );
      }
    if ( best > 0.2F ) {
      P[bestIdx].m.best_match_idx = int( pi );
      order.push_back( P[bestIdx] );
    }
  }
  for ( auto& p : P )
    if ( p.m.best_match_idx == -1 ) order.push_back( p );
  P = std::move( order );
  for ( auto& p : P ) sizeU = std::max( sizeU, p.m.size_u0 + 1 );
  const int tileW = sizeU / numTilesHor, tileH = int( tileW * tileRatio );
  sizeV           = sizeV >= tileH ? sizeV : tileH;
  size_t      width = size_t( sizeU ) * occRes, height = size_t( sizeV ) * occRes;
  BlockCanvas cv( sizeU, sizeV );
  for ( auto& p : P ) {
    pccb200_patch& m = p.m;
    if ( m.best_match_idx != -1 ) {
      const pccb200_patch& q = prev.pl.patches[m.best_match_idx].m;
      m.u0 = q.u0, m.v0 = q.v0, m.orientation = q.orientation;
      placeOne( cv, m.size_u0, m.size_v0, m.size_u0, m.size_v0, PLACE_MATCHED, m.u0, m.v0, m.orientation );
    } else {
      placeOne( cv, m.size_u0, m.size_v0, m.size_u0, m.size_v0, PLACE_BEST_EFFORT, m.u0, m.v0, m.orientation );
    }
    cv.mark( m.u0, m.v0, m.orientation, m.size_u0, m.size_v0, p.occ, m.size_u0 );
    extend( m.u0, m.v0, m.orientation, m.size_u0, m.size_v0, occRes, width, height );
  }
  F.width = width, F.height = height;
}

// ---- PCCEncoder::performDataAdaptiveGPAMethod and helpers (PCCEncoder.cpp:6821-7860)
struct GpaState {
  std::vector<OFrame>* frames;
  int                  occRes;
  size_t               frameWidthIn, frameHeightIn;  // minimumImageWidth / Height
  std::vector<size_t>  curW, curH, preW, preH;       // Cur / PrePCCGPAFrameSize per frame
};
typedef std::map<size_t, std::vector<std::pair<size_t, size_t>>> Tracks;  // GlobalPatches: track -> (frame, patch)
struct UnionPatch {
  int                  sizeU0 = 0, sizeV0 = 0, u0 = 0, v0 = 0, orient = 0;
  std::vector<uint8_t> occ;
};
typedef std::map<size_t, UnionPatch> Unions;

// packingFirstFrame (:7226-7356): the first frame of a sub-context is packed on its own (against frame-1 when there is one)
void gpaPackFirstFrame( GpaState& S, size_t f, bool hasRef ) {
  OFrame& F = ( *S.frames )[f];
  auto&   P = F.pl.patches;
  int     sizeU = int( F.width ) / S.occRes, sizeV = 0;
  for ( auto& p : P ) sizeV = std::max( sizeV, std::max( p.m.size_u0, p.m.size_v0 ) );
  for ( auto& p : P ) sizeU = std::max( sizeU, p.m.size_u0 + 1 );
  size_t      width = size_t( sizeU ) * S.occRes, height = size_t( sizeV ) * S.occRes;
  BlockCanvas cv( sizeU, sizeV );
  for ( auto& p : P ) {
    OGpa& g  = p.cur;
    g.occ    = p.occ;
    g.sizeU0 = p.m.size_u0, g.sizeV0 = p.m.size_v0;
    if ( p.m.best_match_idx != -1 && hasRef ) {
      const pccb200_patch& q = ( *S.frames )[f - 1].pl.patches[p.m.best_match_idx].m;
      g.u0 = q.u0, g.v0 = q.v0, g.orient = q.orientation;
      placeOne( cv, g.sizeU0, g.sizeV0, g.sizeU0, g.sizeV0, PLACE_MATCHED, g.u0, g.v0, g.orient );
    } else {
      placeOne( cv, g.sizeU0, g.sizeV0, g.sizeU0, g.sizeV0, PLACE_BEST_EFFORT, g.u0, g.v0, g.orient );
    }
    cv.mark( g.u0, g.v0, g.orient, g.sizeU0, g.sizeV0, p.occ, p.m.size_u0 );
    extend( g.u0, g.v0, g.orient, g.sizeU0, g.sizeV0, S.occRes, width, height );
  }
  S.curW[f] = width, S.curH[f] = height;
}

// generateGlobalPatches (:7003-7060): extend every live track by its best match in frame f, or end it
void gpaExtendTracks( GpaState& S, size_t f, Tracks& tracks, size_t preIndex ) {
  auto& C = ( *S.frames )[f].pl.patches;
  for ( auto& t : tracks ) {
    auto& tp = t.second;
    if ( tp.empty() ) continue;
    const pccb200_patch& q = ( *S.frames )[tp[preIndex].first].pl.patches[tp[preIndex].second].m;
    float                best = 0.0F;
    int                  bestIdx = -1;
    for ( size_t ci = 0; ci < C.size(); ++ci )
      if ( q.view_id == C[ci].m.view_id && !C[ci].cur.matched ) {
        const float iou = rectIou( q, C[ci].m );
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
      OGpa& g  = ( *S.frames )[fp.first].pl.patches[fp.second].cur;
      g.global = true, g.track = int( t.first );
    }
}

// unionPatchGenerationAndPacking (:7062-7224): one union patch per live track, packed best effort (or with the orientation
// of the reference frame's matched patch); returns the height of the union packing in pixels
size_t gpaPackUnions( GpaState& S, const Tracks& tracks, size_t frameWidth, Unions& unions, int refFrame, bool useRef ) {
  unions.clear();
  for ( auto& t : tracks ) {
    if ( t.second.empty() ) continue;
    UnionPatch U;
    for ( auto& fp : t.second ) {
      const pccb200_patch& m = ( *S.frames )[fp.first].pl.patches[fp.second].m;
      U.sizeU0 = std::max( U.sizeU0, m.size_u0 ), U.sizeV0 = std::max( U.sizeV0, m.size_v0 );
    }
    U.occ.assign( size_t( U.sizeU0 ) * U.sizeV0, 0 );
    U.orient = 0;  // PCCPatch() default
    if ( useRef ) {
      const int matched = ( *S.frames )[t.second[0].first].pl.patches[t.second[0].second].m.best_match_idx;
      U.orient          = matched == -1 ? -1 : ( *S.frames )[refFrame].pl.patches[matched].m.orientation;
    }
    for ( auto& fp : t.second ) {
      const OPatch& p = ( *S.frames )[fp.first].pl.patches[fp.second];
      for ( int v = 0; v < p.m.size_v0; ++v )
        for ( int u = 0; u < p.m.size_u0; ++u )
          if ( p.occ[size_t( v ) * p.m.size_u0 + u] ) U.occ[size_t( v ) * U.sizeU0 + u] = 1;
    }
    unions[t.first] = std::move( U );
  }
  int sizeU = int( frameWidth ) / S.occRes, sizeV = 0;
  for ( auto& u : unions ) sizeU = std::max( sizeU, u.second.sizeU0 + 1 ), sizeV = std::max( sizeV, u.second.sizeV0 + 1 );
  size_t      width = size_t( sizeU ) * S.occRes, height = size_t( sizeV ) * S.occRes;
  BlockCanvas cv( sizeU, sizeV );
  for ( auto& it : unions ) {
    UnionPatch& U = it.second;
    placeOne( cv, U.sizeU0, U.sizeV0, U.sizeU0, U.sizeV0, !useRef ? PLACE_BEST_EFFORT : ( U.orient != -1 ? PLACE_KNOWN_ORIENTATION : PLACE_STICKY ), U.u0, U.v0,
              U.orient );
    cv.mark( U.u0, U.v0, U.orient, U.sizeU0, U.sizeV0, U.occ, U.sizeU0 );
    extend( U.u0, U.v0, U.orient, U.sizeU0, U.sizeV0, S.occRes, width, height );
  }
  return height;
}

// updateGPAPatchInformation (:7493-7529): global patches take the size of their union (their own occupancy, re-strided)
void gpaAdoptUnionSizes( GpaState& S, size_t first, size_t second, Unions& unions ) {
  for ( size_t f = first; f < second; ++f )
    for ( auto& p : ( *S.frames )[f].pl.patches ) {
      OGpa& g = p.cur;
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

// performGPAPacking (:7531-7660): global patches at the position of their union, the others packed around them; returns
// whether this trial packing is rejected
bool gpaPackSubContext( GpaState& S, size_t first, size_t second, Unions& unions, size_t unionsHeight, bool useRef ) {
  bool   tooHigh = false;
  size_t bad     = 0;
  for ( size_t f = first; f < second; ++f ) {
    OFrame& F = ( *S.frames )[f];
    auto&   P = F.pl.patches;
    if ( P.empty() ) return false;
    const auto& prevP = ( *S.frames )[f > 0 ? f - 1 : 0].pl.patches;
    int         sizeU = int( S.frameWidthIn ) / S.occRes;
    const int   sizeV = int( unionsHeight ) / S.occRes;
    for ( auto& p : P ) sizeU = std::max( sizeU, p.cur.sizeU0 + 1 );
    size_t      width = size_t( sizeU ) * S.occRes, height = size_t( sizeV ) * S.occRes;
    BlockCanvas cv( sizeU, sizeV );
    for ( auto& p : P ) {
      OGpa& g = p.cur;
      if ( !g.global ) continue;
      const UnionPatch& U = unions[size_t( g.track )];
      g.u0 = U.u0, g.v0 = U.v0, g.orient = U.orient;
      cv.mark( g.u0, g.v0, g.orient, g.sizeU0, g.sizeV0, g.occ, g.sizeU0 );
      extend( g.u0, g.v0, g.orient, g.sizeU0, g.sizeV0, S.occRes, width, height );
    }
    for ( auto& p : P ) {
      OGpa& g = p.cur;
      if ( g.global ) continue;
      if ( f == 0 || ( f == first && !useRef ) || p.m.best_match_idx == -1 ) {
        // packingWithoutRefForFirstFrameNoglobalPatch (:7662-7745) / the unmatched branch of ...WithRef... (:7824-7858)
        placeOne( cv, g.sizeU0, g.sizeV0, p.m.size_u0, p.m.size_v0, PLACE_BEST_EFFORT, g.u0, g.v0, g.orient );
      } else {
        // packingWithRefForFirstFrameNoglobalPatch (:7746-7822): the matched patch's final placement when it lies before the
        // sub-context, its trial placement otherwise
        const OPatch& q = prevP[p.m.best_match_idx];
        if ( f == first ) g.u0 = q.m.u0, g.v0 = q.m.v0, g.orient = q.m.orientation;
        else g.u0 = q.cur.u0, g.v0 = q.cur.v0, g.orient = q.cur.orient;
        placeOne( cv, g.sizeU0, g.sizeV0, g.sizeU0, g.sizeV0, PLACE_MATCHED, g.u0, g.v0, g.orient );
      }
      cv.mark( g.u0, g.v0, g.orient, g.sizeU0, g.sizeV0, p.occ, p.m.size_u0 );
      extend( g.u0, g.v0, g.orient, g.sizeU0, g.sizeV0, S.occRes, width, height );
    }
    S.curW[f] = width, S.curH[f] = height;
    if ( height > S.frameHeightIn ) {
      tooHigh = true;
      break;
    }
    if ( double( height ) / double( F.height ) >= 1.10 ) ++bad;  // BAD_HEIGHT_THRESHOLD
  }
  return tooHigh || bad > 2;                                     // BAD_CONDITION_THRESHOLD
}

// updatePatchInformation (:7358-7491): the accepted (previous) trial becomes the packing of the sub-context [first, second)
void gpaCommit( GpaState& S, size_t first, size_t second ) {
  auto& FR = *S.frames;
  for ( size_t f = first; f < second; ++f ) {
    FR[f].width = S.preW[f], FR[f].height = S.preH[f];
    for ( auto& p : FR[f].pl.patches ) {
      p.m.size_u0 = p.pre.sizeU0, p.m.size_v0 = p.pre.sizeV0, p.occ = p.pre.occ;
      p.m.u0 = p.pre.u0, p.m.v0 = p.pre.v0, p.m.orientation = p.pre.orient;
      p.m.is_global = p.pre.global ? 1 : 0;
    }
  }
  if ( second - first == 1 ) {
    for ( auto& p : FR[first].pl.patches ) p.m.best_match_idx = -1;
    return;
  }
  int globalCount = 0;
  for ( size_t f = first; f < second; ++f ) {
    auto& P = FR[f].pl.patches;
    for ( size_t i = 0; i < P.size(); ++i ) P[i].m.index = int( i );
    std::vector<OPatch> old = P;
    globalCount             = 0;
    for ( auto& p : old ) globalCount += p.m.is_global;
    P.clear();
    if ( f == first ) {
      for ( auto& p : old )
        if ( p.m.is_global ) P.push_back( p );
    } else {  // global patches follow the order of the patches they match in the previous frame
      for ( int i = 0; i < int( FR[f - 1].pl.patches.size() ); ++i )
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
    auto& P = FR[f].pl.patches;
    for ( int i = 0; i < globalCount; ++i ) {
      if ( f > first ) P[i].m.best_match_idx = i;
      P[i].m.index = i;
    }
    if ( f == second - 1 ) {
      for ( int i = globalCount; i < int( P.size() ); ++i ) P[i].m.index = i;
      continue;
    }
    auto&             N = FR[f + 1].pl.patches;
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
  for ( auto& p : FR[first].pl.patches ) p.m.best_match_idx = -1;
}

void gpaRun( std::vector<OFrame>& frames, int occRes, size_t frameWidthIn, size_t frameHeightIn ) {
  const size_t n = frames.size();
  GpaState     S{ &frames, occRes, frameWidthIn, frameHeightIn, std::vector<size_t>( n, 0 ), std::vector<size_t>( n, 0 ), std::vector<size_t>( n, 0 ),
              std::vector<size_t>( n, 0 ) };
  size_t       preFirst = 0, preSecond = 0;
  Tracks       tracks;
  Unions       unionsCur;
  bool         start = true;
  auto clearCur = [&]( size_t a, size_t b ) {
    for ( size_t j = a; j < b; ++j )
      for ( auto& p : frames[j].pl.patches ) p.cur.reset();
  };
  for ( size_t f = 0; f < n; ++f ) {
    bool useRef = true;
    if ( start ) {
      // initializeSubContext (:6971-6990)
      preFirst = f, preSecond = f + 1;
      tracks.clear();
      auto& P = frames[f].pl.patches;
      for ( size_t i = 0; i < P.size(); ++i ) {
        tracks[i].emplace_back( f, i );
        P[i].cur.global = true, P[i].cur.track = int( i );
      }
      if ( preFirst == 0 ) useRef = false;
      gpaPackFirstFrame( S, f, useRef );
      S.preW[f] = S.curW[f], S.preH[f] = S.curH[f];
      S.curW[f] = S.curH[f] = 0;
      for ( auto& p : P ) {
        p.pre = p.cur;
        p.cur.reset();
      }
      if ( f == n - 1 ) {
        gpaCommit( S, preFirst, preSecond );
        break;
      }
      start = false;
      continue;
    }
    const size_t curFirst = preFirst, curSecond = f + 1;
    int          refFrame = int( curFirst ) - 1;
    if ( curFirst == 0 ) useRef = false, refFrame = -1;
    clearCur( curFirst, curSecond );
    gpaExtendTracks( S, f, tracks, f - curFirst - 1 );
    const size_t unionsHeight = gpaPackUnions( S, tracks, frames[f].width, unionsCur, refFrame, useRef );
    bool         badCount = double( unionsCur.size() ) / tracks.size() < 0.15;
    const bool   badHeight = unionsHeight > frameHeightIn;
    if ( unionsHeight == 0 ) badCount = true;
    bool badPacking = false;
    if ( !badCount && !badHeight ) {
      gpaAdoptUnionSizes( S, curFirst, curSecond, unionsCur );
      badPacking = gpaPackSubContext( S, curFirst, curSecond, unionsCur, unionsHeight, useRef );
    }
    if ( badCount || badHeight || badPacking ) {
      clearCur( curFirst, curSecond );
      unionsCur.clear();
      tracks.clear();
      start = true;
      --f;  // the frame that broke the sub-context opens the next one
      gpaCommit( S, preFirst, preSecond );
    } else {
      for ( size_t j = curFirst; j < curSecond; ++j ) {
        S.preW[j] = S.curW[j], S.preH[j] = S.curH[j];
        for ( auto& p : frames[j].pl.patches ) p.pre = p.cur;
      }
      preFirst = curFirst, preSecond = curSecond;
      clearCur( curFirst, curSecond );
      unionsCur.clear();
      if ( f == n - 1 ) {
        gpaCommit( S, preFirst, preSecond );
        break;
      }
    }
  }
}

// ---- §8f-2: RGB444 -> YUV420 as PCCVideoEncoder::compress does it before the attribute video goes to the codec
// (L/PccLibEncoder/source/PCCVideoEncoder.cpp:331-353: PCCInternalColorConverter "RGB444ToYUV420_8_4"; source
//  L/PccLibColorConverter/source/PCCInternalColorConverter.cpp:407-424, 553-593, 642-666 and the inline filters of
//  L/PccLibColorConverter/include/PCCInternalColorConverter.h:153-183). Filter 4 (DF_GS, the default of compress()):
//  the taps are the reference's own constants, float( c * 512 ).
const float kGsHor[15] = {float( -0.01716352771649 * 512 ), float( 0.0 ), float( +0.04066666714886 * 512 ), float( 0.0 ),
                          float( -0.09154810319329 * 512 ), float( 0.0 ), float( 0.31577823859943 * 512 ), float( 0.50453345032298 * 512 ),
                          float( 0.31577823859943 * 512 ), float( 0.0 ), float( -0.09154810319329 * 512 ), float( 0.0 ),
                          float( 0.04066666714886 * 512 ), float( 0.0 ), float( -0.01716352771649 * 512 )};
const float kGsVer[16] = {float( -0.00945406160902 * 512 ), float( -0.01539537217249 * 512 ), float( 0.02360533018213 * 512 ),
                          float( 0.03519540819902 * 512 ), float( -0.05254456550808 * 512 ), float( -0.08189331229717 * 512 ),
                          float( 0.14630826357715 * 512 ), float( 0.45417830962846 * 512 ), float( 0.45417830962846 * 512 ),
                          float( 0.14630826357715 * 512 ), float( -0.08189331229717 * 512 ), float( -0.05254456550808 * 512 ),
                          float( 0.03519540819902 * 512 ), float( 0.02360533018213 * 512 ), float( -0.01539537217249 * 512 ),
                          float( -0.00945406160902 * 512 )};
inline uint8_t quantise8( float v, bool chroma ) {  // floatYUVToYUV, nbyte 1
  float r = std::round( float( 255. * double( v ) + ( chroma ? 128. : 0. ) ) );
  r       = r < 0.f ? 0.f : r;
  r       = r > 255.f ? 255.f : r;
  return uint8_t( uint16_t( r ) );
}
void rgbToYuv420( const std::vector<uint16_t>& rgbPlanes, size_t W, size_t H, std::vector<uint8_t>& out ) {
  const size_t       Q = W * H, w2 = W / 2, h2 = H / 2;
  std::vector<float> U( Q ), V( Q );
  out.assign( Q + 2 * w2 * h2, 0 );
  for ( size_t i = 0; i < Q; ++i ) {
    const float  r = float( rgbPlanes[i] ) / 255.f, g = float( rgbPlanes[Q + i] ) / 255.f, b = float( rgbPlanes[2 * Q + i] ) / 255.f;
    const double y = 0.212600 * r + 0.715200 * g + 0.072200 * b, u = -0.114572 * r - 0.385428 * g + 0.500000 * b,
                 v = 0.500000 * r - 0.454153 * g - 0.045847 * b;
    out[i] = quantise8( float( y < 0.0 ? 0.0 : ( y > 1.0 ? 1.0 : y ) ), false );
    U[i]   = float( u < -0.5 ? -0.5 : ( u > 0.5 ? 0.5 : u ) );
    V[i]   = float( v < -0.5 ? -0.5 : ( v > 0.5 ? 0.5 : v ) );
  }
  const float scale = 1.0f / float( 1 << 9 );
  auto        down  = [&]( const std::vector<float>& in, uint8_t* dst ) {
    std::vector<float> tmp( w2 * H );
    for ( size_t i = 0; i < H; ++i )
      for ( size_t j = 0; j < w2; ++j ) {
        double acc = 0;
        for ( int t = 0; t < 15; ++t ) {
          const long x = std::min<long>( std::max<long>( long( 2 * j ) + t - 7, 0 ), long( W ) - 1 );
          acc += double( kGsHor[t] ) * double( in[i * W + size_t( x )] );
        }
        tmp[i * w2 + j] = float( ( acc + double( 0.f ) ) * double( scale ) );
      }
    for ( size_t i = 0; i < h2; ++i )
      for ( size_t j = 0; j < w2; ++j ) {
        double acc = 0;
        for ( int t = 0; t < 16; ++t ) {
          const long y = std::min<long>( std::max<long>( long( 2 * i ) + t - 7, 0 ), long( H ) - 1 );
          acc += double( kGsVer[t] ) * double( tmp[size_t( y ) * w2 + j] );
        }
        dst[i * w2 + j] = quantise8( float( ( acc + double( 0.f ) ) * double( scale ) ), true );
      }
  };
  down( U, out.data() + Q );
  down( V, out.data() + Q + w2 * h2 );
}

// weighted mean used by the push-pull filter (PCCEncoder.cpp:6357-6369)
inline int mean4w( int p1, int w1, int p2, int w2, int p3, int w3, int p4, int w4 ) {
  return ( p1 * w1 + p2 * w2 + p3 * w3 + p4 * w4 ) / ( w1 + w2 + w3 + w4 );
}

struct Img3 {
  size_t                w = 0, h = 0;
  std::vector<uint16_t> c[3];
  void                  alloc( size_t W, size_t H ) {
    w = W, h = H;
    for ( auto& ch : c ) ch.assign( W * H, 0 );
  }
};

// PCCEncoder::pushPullMip (PCCEncoder.cpp:6372-6441)
void pullLevel( const Img3& img, const std::vector<uint8_t>& occ, Img3& mip, std::vector<uint8_t>& mipOcc ) {
  const size_t W = img.w, H = img.h, nw = ( W + 1 ) >> 1, nh = ( H + 1 ) >> 1;
  mip.alloc( nw, nh );
  mipOcc.assign( nw * nh, 0 );
  for ( size_t y = 0; y < nh; ++y )
    for ( size_t x = 0; x < nw; ++x ) {
      const size_t X = x << 1, Y = y << 1;
      const bool   in2 = X + 1 < W, in3 = Y + 1 < H;
      const int    w1 = occ[X + W * Y] ? 255 : 0, w2 = ( in2 && occ[X + 1 + W * Y] ) ? 255 : 0,
                w3 = ( in3 && occ[X + W * ( Y + 1 )] ) ? 255 : 0, w4 = ( in2 && in3 && occ[X + 1 + W * ( Y + 1 )] ) ? 255 : 0;
      if ( w1 + w2 + w3 + w4 == 0 ) continue;
      for ( int ch = 0; ch < 3; ++ch ) {
        const auto& s  = img.c[ch];
        const int   v1 = uint8_t( s[X + W * Y] ), v2 = in2 ? uint8_t( s[X + 1 + W * Y] ) : 0, v3 = in3 ? uint8_t( s[X + W * ( Y + 1 )] ) : 0,
                  v4 = ( in2 && in3 ) ? uint8_t( s[X + 1 + W * ( Y + 1 )] ) : 0;
        mip.c[ch][x + nw * y] = uint16_t( mean4w( v1, w1, v2, w2, v3, w3, v4, w4 ) );
      }
      mipOcc[x + nw * y] = 1;
    }
}

// PCCEncoder::pushPullFill (PCCEncoder.cpp:6444-6540)
void pushLevel( Img3& img, const Img3& mip, const std::vector<uint8_t>& occ, int iters ) {
  const int W = int( img.w ), H = int( img.h ), w = int( mip.w ), h = int( mip.h );
  for ( int Y = 0; Y < H; ++Y )
    for ( int X = 0; X < W; ++X ) {
      if ( occ[X + size_t( W ) * Y] ) continue;
      const int  x = X >> 1, y = Y >> 1;
      const int  dx = ( X & 1 ) ? 1 : -1, dy = ( Y & 1 ) ? 1 : -1;  // the diagonal neighbour the sample leans towards
      const bool okx = dx < 0 ? x > 0 : x < w - 1, oky = dy < 0 ? y > 0 : y < h - 1;
      for ( int ch = 0; ch < 3; ++ch ) {
        const auto& m  = mip.c[ch];
        const int   v  = m[x + size_t( w ) * y];
        const int   vx = okx ? m[x + dx + size_t( w ) * y] : 0, vy = oky ? m[x + size_t( w ) * ( y + dy )] : 0,
                  vd = ( okx && oky ) ? m[x + dx + size_t( w ) * ( y + dy )] : 0;
        img.c[ch][X + size_t( W ) * Y] = uint16_t( mean4w( v, 144, vx, okx ? 48 : 0, vy, oky ? 48 : 0, vd, ( okx && oky ) ? 16 : 0 ) );
      }
    }
  Img3 tmp = img;
  for ( int n = 0; n < iters; ++n ) {
    for ( int y = 0; y < H; ++y )
      for ( int x = 0; x < W; ++x ) {
        if ( occ[x + size_t( W ) * y] ) continue;
        const int x1 = x > 0 ? x - 1 : x, y1 = y > 0 ? y - 1 : y, x2 = x < W - 1 ? x + 1 : x, y2 = y < H - 1 ? y + 1 : y;
        for ( int ch = 0; ch < 3; ++ch ) {
          const auto& s   = img.c[ch];
          auto        at  = [&]( int xx, int yy ) { return int( s[xx + size_t( W ) * yy] ); };
          const int   val = at( x1, y1 ) + at( x2, y1 ) + at( x1, y2 ) + at( x2, y2 ) + at( x1, y ) + at( x2, y ) + at( x, y1 ) + at( x, y2 );
          tmp.c[ch][x + size_t( W ) * y] = uint16_t( ( val + 4 ) >> 3 );
        }
      }
    std::swap( img, tmp );
  }
}

// PCCEncoder::dilateSmoothedPushPull (PCCEncoder.cpp:6542-6591)
void pushPull( Img3& image, const std::vector<uint8_t>& occ ) {
  std::vector<Img3>                 mips;
  std::vector<std::vector<uint8_t>> mocc;
  for ( ;; ) {
    mips.emplace_back(), mocc.emplace_back();
    const size_t l = mips.size() - 1;
    if ( l > 0 ) pullLevel( mips[l - 1], mocc[l - 1], mips[l], mocc[l] );
    else pullLevel( image, occ, mips[0], mocc[0] );
    if ( mips[l].w <= 4 || mips[l].h <= 4 ) break;
  }
  int iters = 4;
  for ( int i = int( mips.size() ) - 1; i >= 0; --i ) {
    if ( i > 0 ) pushLevel( mips[i - 1], mips[i], mocc[i - 1], iters );
    else pushLevel( image, mips[0], occ, iters );
    iters = std::min( iters + 1, 16 );
  }
}

}  // namespace

void* pcco_encode_gof_canvas( int nframes, const int16_t* const* xyz, const uint8_t* const* rgb, const size_t* n,
                              const pccb200_seg_params* prm, int occPrec, int stopAfter, size_t forceW, size_t forceH );

void* pcco_encode_gof( int nframes, const int16_t* const* xyz, const uint8_t* const* rgb, const size_t* n,
                       const pccb200_seg_params* prm, int occPrec, int stopAfter ) {
  return pcco_encode_gof_canvas( nframes, xyz, rgb, n, prm, occPrec, stopAfter, 0, 0 );
}

// forceW/forceH (0 = none): lower bound for the canvas, i.e. the all-reduced size when the GOF's frames are sharded over ranks
void* pcco_encode_gof_canvas( int nframes, const int16_t* const* xyz, const uint8_t* const* rgb, const size_t* n,
                              const pccb200_seg_params* prm, int occPrec, int stopAfter, size_t forceW, size_t forceH ) {
  OGof* G = new OGof();
  G->frames.resize( nframes );
  const int occRes = prm->occupancy_resolution;
  const int minW = prm->geometry_bitdepth_3d > 11 ? 2560 : 1280, minH = 1280;
  // ---- a1..a11 per frame, then a13 packing
  for ( int f = 0; f < nframes; ++f ) {
    OPatchList* pl = static_cast<OPatchList*>( pcco_segment_frame( xyz[f], rgb[f], n[f], prm ) );
    G->frames[f].pl = std::move( *pl );
    delete pl;
    G->frames[f].height = minH;
    if ( f == 0 || prm->global_patch_allocation == 0 ) packFrame( G->frames[f], occRes, minW, 2, 1.0 );
    else packFrameAfter( G->frames[f], G->frames[f - 1], occRes, minW, 2, 1.0 );
  }
  if ( prm->global_patch_allocation != 0 && nframes > 0 && !G->frames[0].pl.patches.empty() && !getenv( "PCCO_DEBUG_NO_GPA" ) ) {
    // PCCEncoder::placeSegments (:4807-4827): common tile size, global patch allocation over the GOF, common tile size again
    size_t tw = minW, th = minH;
    for ( auto& F : G->frames ) tw = std::max( tw, F.width ), th = std::max( th, F.height );
    for ( auto& F : G->frames ) F.width = tw, F.height = th;
    gpaRun( G->frames, occRes, minW, minH );
  }
  // ---- a14: one canvas size for the GOF (resizeTileGeometryVideo + resizeGeometryVideo, PCCEncoder.cpp:5546-5634)
  size_t W = minW, H = minH;
  for ( auto& F : G->frames ) W = std::max( W, F.width ), H = std::max( H, F.height );
  W = size_t( std::ceil( double( W ) / 64.0 ) * 64 ), H = size_t( std::ceil( double( H ) / 64.0 ) * 64 );
  W = std::max( W, forceW ), H = std::max( H, forceH );
  for ( auto& F : G->frames ) F.width = W, F.height = H;
  if ( stopAfter == 1 ) return G;
  const size_t bw = W / occRes, bh = H / occRes, ow = W / occPrec, oh = H / occPrec;
  for ( int f = 0; f < nframes; ++f ) {
    OFrame& F = G->frames[f];
    auto&   P = F.pl.patches;
    // ---- a16 generateOccupancyMap (PCCEncoder.cpp:3768-3784)
    F.occupancy.assign( W * H, 0 );
    for ( auto& p : P )
      for ( int v = 0; v < p.m.size_v; ++v )
        for ( int u = 0; u < p.m.size_u; ++u )
          if ( p.depth[0][size_t( v ) * p.m.size_u + u] < kInfDepth ) {
            int x, y;
            pixelToCanvas( p.m, occRes, u, v, x, y );
            F.occupancy[size_t( y ) * W + x] = 1;
          }
    // ---- a17 generateOccupancyMapVideo (PCCEncoder.cpp:806-861): OR over each precision x precision cell
    F.omVideo.assign( ow * oh, 0 );
    for ( size_t y = 0; y < H; ++y )
      for ( size_t x = 0; x < W; ++x )
        if ( F.occupancy[y * W + x] ) F.omVideo[( y / occPrec ) * ow + x / occPrec] = 1;
    // ---- a18 generateBlockToPatchFromOccupancyMapVideo (L/PccLibCommon/source/PCCCodec.cpp:1736-1774): last patch wins
    F.blockToPatch.assign( bw * bh, 0 );
    for ( size_t pi = 0; pi < P.size(); ++pi ) {
      const pccb200_patch& m = P[pi].m;
      for ( int vb = 0; vb < m.size_v0; ++vb )
        for ( int ub = 0; ub < m.size_u0; ++ub ) {
          int bx, by;
          blockToCanvas( m, ub, vb, bx, by );
          bool any = false;
          for ( int v1 = 0; v1 < occRes && !any; ++v1 )
            for ( int u1 = 0; u1 < occRes && !any; ++u1 ) {
              int x, y;
              pixelToCanvas( m, occRes, ub * occRes + u1, vb * occRes + v1, x, y );
              any = F.omVideo[size_t( y / occPrec ) * ow + x / occPrec] != 0;
            }
          if ( any ) F.blockToPatch[size_t( by ) * bw + bx] = uint32_t( pi + 1 );
        }
    }
    // ---- a19 generateIntraImage (PCCEncoder.cpp:3929-3960) + a20 dilate3DPadding (:5951-6130, geometryPadding 0)
    for ( int map = 0; map < 2; ++map ) {
      auto& img = F.geo[map];
      img.assign( W * H, 0 );
      for ( auto& p : P )
        for ( int v = 0; v < p.m.size_v; ++v )
          for ( int u = 0; u < p.m.size_u; ++u ) {
            const int16_t d = p.depth[map][size_t( v ) * p.m.size_u + u];
            if ( d < kInfDepth ) {
              int x, y;
              pixelToCanvas( p.m, occRes, u, v, x, y );
              img[size_t( y ) * W + x] = uint16_t( d );
            }
          }
      // fill mask: original occupancy restricted to cells active in the (lossless) occupancy video == the occupancy
      std::vector<uint32_t> state( W * H, 0 );
      for ( size_t i = 0; i < W * H; ++i ) state[i] = F.occupancy[i] ? 1 : 0;
      for ( size_t vb = 0; vb < bh; ++vb )
        for ( size_t ub = 0; ub < bw; ++ub ) {
          const size_t x0 = ub * occRes, y0 = vb * occRes;
          size_t       filled = 0;
          for ( int v = 0; v < occRes; ++v )
            for ( int u = 0; u < occRes; ++u ) filled += state[( y0 + v ) * W + x0 + u] == 1;
          if ( filled == 0 ) {  // empty block: smear the left (else the upper) neighbour's border pixels, in raster order
            for ( int v = 0; v < occRes; ++v )
              for ( int u = 0; u < occRes; ++u ) {
                const size_t x = x0 + u, y = y0 + v;
                if ( ub > 0 ) img[y * W + x] = img[y * W + x - 1];
                else if ( vb > 0 ) img[y * W + x] = img[( y - 1 ) * W + x];
              }
            continue;
          }
          uint32_t iteration = 1;
          std::vector<int>      sum( size_t( occRes ) * occRes );
          std::vector<uint32_t> cnt( size_t( occRes ) * occRes );
          while ( filled < size_t( occRes ) * occRes ) {
            std::fill( sum.begin(), sum.end(), 0 ), std::fill( cnt.begin(), cnt.end(), 0u );
            static const int nb[4][2] = {{0, -1}, {-1, 0}, {1, 0}, {0, 1}};
            for ( int v = 0; v < occRes; ++v )
              for ( int u = 0; u < occRes; ++u ) {
                if ( state[( y0 + v ) * W + x0 + u] != iteration ) continue;
                for ( auto& d : nb ) {
                  const int u3 = u + d[0], v3 = v + d[1];
                  if ( u3 < 0 || v3 < 0 || u3 >= occRes || v3 >= occRes || state[( y0 + v3 ) * W + x0 + u3] != 0 ) continue;
                  sum[size_t( v3 ) * occRes + u3] += img[( y0 + v ) * W + x0 + u];
                  ++cnt[size_t( v3 ) * occRes + u3];
                }
              }
            for ( int v = 0; v < occRes; ++v )
              for ( int u = 0; u < occRes; ++u ) {
                const uint32_t c = cnt[size_t( v ) * occRes + u];
                if ( !c ) continue;
                ++filled;
                state[( y0 + v ) * W + x0 + u] = iteration + 1;
                img[( y0 + v ) * W + x0 + u]   = uint16_t( ( sum[size_t( v ) * occRes + u] + int( c / 2 ) ) / int( c ) );
              }
            ++iteration;
          }
        }
    }
    // ---- a21 dilateGroupGeometryVideo (PCCEncoder.cpp:3717-3739)
    for ( size_t y = 0; y < H; ++y )
      for ( size_t x = 0; x < W; ++x )
        if ( F.omVideo[( y / occPrec ) * ow + x / occPrec] == 0 ) {
          const uint32_t avg = ( uint32_t( F.geo[0][y * W + x] ) + uint32_t( F.geo[1][y * W + x] ) + 1 ) >> 1;
          F.geo[0][y * W + x] = F.geo[1][y * W + x] = uint16_t( avg );
        }
  }
  if ( stopAfter == 2 ) return G;
  for ( int f = 0; f < nframes; ++f ) {
    OFrame& F = G->frames[f];
    auto&   P = F.pl.patches;
    // ---- a22 generatePointCloud (PCCCodec.cpp:519-980): the tile occupancy becomes the up-sampled occupancy video
    std::vector<uint8_t> occ( W * H );
    for ( size_t y = 0; y < H; ++y )
      for ( size_t x = 0; x < W; ++x ) occ[y * W + x] = F.omVideo[( y / occPrec ) * ow + x / occPrec];
    for ( size_t pi = 0; pi < P.size(); ++pi ) {
      const pccb200_patch& m = P[pi].m;
      for ( int vb = 0; vb < m.size_v0; ++vb )
        for ( int ub = 0; ub < m.size_u0; ++ub ) {
          int bx, by;
          blockToCanvas( m, ub, vb, bx, by );
          if ( F.blockToPatch[size_t( by ) * bw + bx] != pi + 1 ) continue;
          for ( int v1 = 0; v1 < occRes; ++v1 )
            for ( int u1 = 0; u1 < occRes; ++u1 ) {
              const int u = ub * occRes + u1, v = vb * occRes + v1;
              int       x, y;
              pixelToCanvas( m, occRes, u, v, x, y );
              if ( !occ[size_t( y ) * W + x] ) continue;
              int16_t pt[2][3];
              for ( int map = 0; map < 2; ++map ) {  // PCCPatch::generatePoint (PCCPatch.h:177-207)
                const uint16_t depth = F.geo[map][size_t( y ) * W + x];
                double         nc    = 0;
                if ( m.projection_mode == 0 ) nc = double( depth ) + double( m.d1 );
                else {
                  const double t = double( m.d1 ) - double( depth );
                  if ( t > 0 ) nc = t;
                }
                pt[map][m.normal_axis]    = int16_t( nc );
                pt[map][m.tangent_axis]   = int16_t( double( u ) + m.u1 );
                pt[map][m.bitangent_axis] = int16_t( double( v ) + m.v1 );
              }
              for ( int map = 0; map < 2; ++map ) {
                if ( map == 1 && pt[1][0] == pt[0][0] && pt[1][1] == pt[0][1] && pt[1][2] == pt[0][2] ) continue;  // removeDuplicatePoints
                F.recXyz.insert( F.recXyz.end(), pt[map], pt[map] + 3 );
                F.pointToPixel.push_back( uint32_t( x ) ), F.pointToPixel.push_back( uint32_t( y ) ), F.pointToPixel.push_back( uint32_t( map ) );
                F.recPartition.push_back( uint32_t( pi ) );
              }
            }
        }
    }
    // identifyBoundaryPoints (PCCCodec.cpp:268-327): 3x3 then 5x5 ring of the occupancy, plus the image border
    const size_t R = F.recPartition.size();
    F.recBoundary.assign( R, 0 );
    auto at = [&]( long x, long y ) { return occ[size_t( y ) * W + size_t( x )]; };
    for ( size_t i = 0; i < R; ++i ) {
      const long x = F.pointToPixel[3 * i], y = F.pointToPixel[3 * i + 1];
      if ( !at( x, y ) ) continue;
      bool b = false;
      const bool yIn = y > 0 && y < long( H ) - 1, xIn = x > 0 && x < long( W ) - 1;
      if ( yIn && ( !at( x, y - 1 ) || !at( x, y + 1 ) ) ) b = true;
      if ( !b && xIn && ( !at( x + 1, y ) || !at( x - 1, y ) ) ) b = true;
      if ( !b && yIn && x > 0 && ( !at( x - 1, y - 1 ) || !at( x - 1, y + 1 ) ) ) b = true;
      if ( !b && yIn && x < long( W ) - 1 && ( !at( x + 1, y - 1 ) || !at( x + 1, y + 1 ) ) ) b = true;
      if ( y == 0 || y == long( H ) - 1 || x == 0 || x == long( W ) - 1 ) b = true;
      if ( !b ) {
        for ( int ix = -2; ix <= 2 && !b; ++ix )
          for ( int iy = -2; iy <= 2 && !b; ++iy )
            if ( ( std::abs( ix ) > 1 || std::abs( iy ) > 1 ) && y + iy >= 0 && y + iy < long( H ) && x + ix >= 0 && x + ix < long( W ) &&
                 !at( x + ix, y + iy ) )
              b = true;
        if ( y == 1 || y == long( H ) - 2 || x == 1 || x == long( W ) - 2 ) b = true;
      }
      F.recBoundary[i] = b ? 1 : 0;
    }
    if ( stopAfter == 3 ) continue;
    // ---- a23 transferColors (L/PccLibCommon/source/PCCPointSet.cpp:807-1124), CTC values: 8 fwd / 1 bwd neighbours,
    //      distance-weighted, offsets 4, all distance/colour gates disabled (1000 >= 512), search range 0, fixWeight.
    F.recRgb.assign( 3 * R, 0 );
    if ( R && n[f] ) {
      Tree src, tgt;
      src.build( xyz[f], n[f] );
      tgt.build( F.recXyz.data(), R );
      std::vector<uint8_t> fwd( 3 * R );
      for ( size_t i = 0; i < R; ++i ) {
        uint32_t id[8];
        double   d[8];
        KnnSet   rs( 8, id, d );
        src.search( rs, &F.recXyz[3 * i] );
        const uint8_t* c0 = rgb[f] + 3 * size_t( id[0] );
        if ( d[0] < 0.0001 || rs.cnt == 1 ) {
          for ( int k = 0; k < 3; ++k ) fwd[3 * i + k] = c0[k];
        } else {
          double acc[3] = {0, 0, 0}, sw = 0;
          for ( int j = 0; j < rs.cnt; ++j ) {
            const double wgt = 1 / ( d[j] + 4.0 );
            for ( int k = 0; k < 3; ++k ) acc[k] += rgb[f][3 * size_t( id[j] ) + k] * wgt;
            sw += wgt;
          }
          for ( int k = 0; k < 3; ++k ) {
            const double v = std::round( acc[k] / sw );
            fwd[3 * i + k] = uint8_t( v < 0.0 ? 0.0 : ( v > 255.0 ? 255.0 : v ) );
          }
        }
      }
      struct Vote {
        double  dist;
        uint8_t c[3];
      };
      std::vector<std::vector<Vote>> votes( R );
      for ( size_t s = 0; s < n[f]; ++s ) {
        uint32_t id;
        double   d;
        KnnSet   rs( 1, &id, &d );
        tgt.search( rs, xyz[f] + 3 * s );
        votes[id].push_back( Vote{d, {rgb[f][3 * s], rgb[f][3 * s + 1], rgb[f][3 * s + 2]}} );
      }
      for ( size_t i = 0; i < R; ++i ) {
        auto& v = votes[i];
        if ( v.empty() ) {
          for ( int k = 0; k < 3; ++k ) F.recRgb[3 * i + k] = fwd[3 * i + k];
          continue;
        }
        std::sort( v.begin(), v.end(), []( const Vote& a, const Vote& b ) { return a.dist < b.dist; } );
        double c2[3] = {0, 0, 0};
        if ( v[0].dist < 0.0001 || v.size() == 1 ) {
          for ( int k = 0; k < 3; ++k ) c2[k] = v[0].c[k];
        } else {
          double sw = 0;
          for ( auto& e : v ) {
            const double wgt = 1 / ( std::sqrt( e.dist ) + 4.0 );
            for ( int k = 0; k < 3; ++k ) c2[k] += ( e.c[k] * wgt );
            sw += wgt;
          }
          for ( int k = 0; k < 3; ++k ) c2[k] /= sw;
        }
        for ( int k = 0; k < 3; ++k ) {
          const double r = std::round( 0.0 * double( fwd[3 * i + k] ) + 1.0 * c2[k] );
          F.recRgb[3 * i + k] = uint8_t( r < 0.0 ? 0.0 : ( r > 255.0 ? 255.0 : r ) );
        }
      }
    }
    // (a24 presmoothPointCloudColor only touches points of boundary type 2, which no point has at this stage.)
    // ---- a25 generateAttributeVideo(tile) (PCCEncoder.cpp:6736-6819): T1 shows the D1 colour, else the D0 colour
    Img3 T[2];
    T[0].alloc( W, H ), T[1].alloc( W, H );
    std::vector<uint8_t> hasD1( W * H, 0 );
    for ( size_t i = 0; i < R; ++i ) {
      const size_t x = F.pointToPixel[3 * i], y = F.pointToPixel[3 * i + 1], map = F.pointToPixel[3 * i + 2];
      for ( int k = 0; k < 3; ++k ) T[map].c[k][y * W + x] = F.recRgb[3 * i + k];
      if ( map == 0 ) {
        if ( !hasD1[y * W + x] )
          for ( int k = 0; k < 3; ++k ) T[1].c[k][y * W + x] = F.recRgb[3 * i + k];
      } else {
        hasD1[y * W + x] = 1;
      }
    }
    for ( int map = 0; map < 2; ++map ) {
      F.attrRaw[map].resize( 3 * W * H );
      for ( int k = 0; k < 3; ++k ) std::copy( T[map].c[k].begin(), T[map].c[k].end(), F.attrRaw[map].begin() + k * W * H );
    }
    // ---- a26 push-pull background fill with the block-precision occupancy
    for ( int map = 0; map < 2; ++map ) pushPull( T[map], occ );
    // group dilation of the attribute maps (PCCEncoder.cpp:391-413, CTC: two maps in one stream, groupDilation on): where the
    // block-precision occupancy is empty, T0 = T1 = rounded mean of their 8-bit values
    for ( size_t q = 0; q < size_t( W ) * H; ++q )
      if ( !occ[q] )
        for ( int k = 0; k < 3; ++k ) {
          const uint32_t mean = ( uint32_t( uint8_t( T[0].c[k][q] ) ) + uint32_t( uint8_t( T[1].c[k][q] ) ) + 1 ) >> 1;
          T[0].c[k][q] = T[1].c[k][q] = uint8_t( mean );
        }
    for ( int map = 0; map < 2; ++map ) {
      F.attr[map].resize( 3 * W * H );
      for ( int k = 0; k < 3; ++k ) std::copy( T[map].c[k].begin(), T[map].c[k].end(), F.attr[map].begin() + k * W * H );
      rgbToYuv420( F.attr[map], W, H, F.attrYuv[map] );
    }
  }
  return G;
}

// a13-a15 alone on caller-given patch lists (metadata + block occupancy): the packing of pcco_encode_gof without a segmentation
void* pcco_pack_gof( int nframes, const int* counts, const pccb200_patch* patches, const uint8_t* occ, const int64_t* occBase, int ra, int bits ) {
  OGof* G = new OGof();
  G->frames.resize( nframes );
  const int occRes = 16, minW = bits + 1 > 11 ? 2560 : 1280, minH = 1280;
  size_t    at     = 0;
  for ( int f = 0; f < nframes; ++f ) {
    auto& P = G->frames[f].pl.patches;
    P.resize( counts[f] );
    for ( int i = 0; i < counts[f]; ++i, ++at ) {
      P[i].m                = patches[at];
      P[i].m.best_match_idx = -1, P[i].m.is_global = 0, P[i].m.u0 = P[i].m.v0 = P[i].m.orientation = 0;
      const uint8_t* o = occ + occBase[f] + patches[at].occ_offset;
      P[i].occ.assign( o, o + size_t( P[i].m.size_u0 ) * P[i].m.size_v0 );
    }
    G->frames[f].height = minH;
    if ( f == 0 || !ra ) packFrame( G->frames[f], occRes, minW, 2, 1.0 );
    else packFrameAfter( G->frames[f], G->frames[f - 1], occRes, minW, 2, 1.0 );
  }
  if ( ra && nframes > 0 && !G->frames[0].pl.patches.empty() ) {
    size_t tw = minW, th = minH;
    for ( auto& F : G->frames ) tw = std::max( tw, F.width ), th = std::max( th, F.height );
    for ( auto& F : G->frames ) F.width = tw, F.height = th;
    gpaRun( G->frames, occRes, minW, minH );
  }
  size_t W = minW, H = minH;
  for ( auto& F : G->frames ) W = std::max( W, F.width ), H = std::max( H, F.height );
  W = size_t( std::ceil( double( W ) / 64.0 ) * 64 ), H = size_t( std::ceil( double( H ) / 64.0 ) * 64 );
  for ( auto& F : G->frames ) F.width = W, F.height = H;
  return G;
}

void pcco_gof_free( void* h ) { delete static_cast<OGof*>( h ); }
void pcco_gof_dims( void* h, int f, size_t* w, size_t* hgt, size_t* recPoints ) {
  auto& F = static_cast<OGof*>( h )->frames[f];
  *w = F.width, *hgt = F.height, *recPoints = F.recPartition.size();
}
void*  pcco_gof_patches( void* h, int f ) { return &static_cast<OGof*>( h )->frames[f].pl; }
size_t pcco_gof_get( void* h, int f, int what, void* dst ) {
  auto& R   = static_cast<OGof*>( h )->frames[f];
  auto  put = [&]( const void* src, size_t bytes ) {
    if ( dst && bytes ) std::memcpy( dst, src, bytes );
  };
  switch ( what ) {
    case 1: put( R.occupancy.data(), R.occupancy.size() ); return R.occupancy.size();
    case 2: put( R.omVideo.data(), R.omVideo.size() ); return R.omVideo.size();
    case 3: put( R.blockToPatch.data(), R.blockToPatch.size() * 4 ); return R.blockToPatch.size();
    case 4: put( R.geo[0].data(), R.geo[0].size() * 2 ); return R.geo[0].size();
    case 5: put( R.geo[1].data(), R.geo[1].size() * 2 ); return R.geo[1].size();
    case 6: put( R.recXyz.data(), R.recXyz.size() * 2 ); return R.recXyz.size();
    case 7: put( R.pointToPixel.data(), R.pointToPixel.size() * 4 ); return R.pointToPixel.size();
    case 8: put( R.recPartition.data(), R.recPartition.size() * 4 ); return R.recPartition.size();
    case 9: put( R.recBoundary.data(), R.recBoundary.size() * 2 ); return R.recBoundary.size();
    case 10: put( R.recRgb.data(), R.recRgb.size() ); return R.recRgb.size();
    case 11: put( R.attrRaw[0].data(), R.attrRaw[0].size() * 2 ); return R.attrRaw[0].size();
    case 12: put( R.attrRaw[1].data(), R.attrRaw[1].size() * 2 ); return R.attrRaw[1].size();
    case 13: put( R.attr[0].data(), R.attr[0].size() * 2 ); return R.attr[0].size();
    case 14: put( R.attr[1].data(), R.attr[1].size() * 2 ); return R.attr[1].size();
    case 15: put( R.attrYuv[0].data(), R.attrYuv[0].size() ); return R.attrYuv[0].size();
    case 16: put( R.attrYuv[1].data(), R.attrYuv[1].size() ); return R.attrYuv[1].size();
    default: return 0;
  }
}

// ---- §8f-1 (first stage of the post-reconstruction chain): grid-based geometry smoothing, PCCCodec::smoothPointCloudPostprocess
// with gridSmoothing (L/PccLibCommon/source/PCCCodec.cpp:54-150), addGridCentroid (:982-1000), gridFiltering (:1002-1065),
// smoothPointCloudGrid (:1067-1106). In: the reconstructed cloud as generatePointCloud leaves it (positions, boundary point
// types, patch index per point). Out: boundary points whose tri-linearly weighted neighbourhood centroid lies far enough away
// move onto it and get boundary type 3.
void pcco_smooth_geometry( int16_t* xyz, uint16_t* boundary, const uint32_t* partition, size_t n, int gridSize, double threshold ) {
  if ( n == 0 ) return;
  int maxSize = std::max( std::max( xyz[0], xyz[1] ), xyz[2] );
  for ( size_t i = 0; i < 3 * n; ++i ) maxSize = std::max( maxSize, int( xyz[i] ) );
  const int w = ( maxSize + gridSize - 1 ) / gridSize, disth = std::max( gridSize / 2, 1 ), th = gridSize * w, half = gridSize / 2;
  auto      nearBorder = [&]( const int16_t* p ) {
    return p[0] < disth || p[1] < disth || p[2] < disth || th <= p[0] + disth || th <= p[1] + disth || th <= p[2] + disth;
  };
  // cells around boundary points, numbered by first appearance
  std::vector<int> cellIndex( size_t( w ) * w * w, -1 );
  int              cells = 0;
  for ( size_t i = 0; i < n; ++i ) {
    const int16_t* p = xyz + 3 * i;
    if ( boundary[i] != 1 || nearBorder( p ) ) continue;
    int q[3];
    for ( int k = 0; k < 3; ++k ) q[k] = p[k] / gridSize + ( p[k] % gridSize < half ? -1 : 0 );
    for ( int ix = 0; ix < 2; ++ix )
      for ( int iy = 0; iy < 2; ++iy )
        for ( int iz = 0; iz < 2; ++iz ) {
          int& c = cellIndex[size_t( q[0] + ix ) + size_t( q[1] + iy ) * w + size_t( q[2] + iz ) * w * w];
          if ( c == -1 ) c = cells++;
        }
  }
  // per cell: float centroid of all points in it (uint16 count, as the reference), first patch, "holds several patches" flag
  std::vector<uint16_t> count( cells, 0 );
  std::vector<float>    center( 3 * size_t( cells ), 0.f );
  std::vector<uint32_t> firstPatch( cells, 0 );
  std::vector<uint8_t>  mixed( cells, 0 );
  for ( size_t i = 0; i < n; ++i ) {
    const int16_t* p = xyz + 3 * i;
    if ( nearBorder( p ) ) continue;
    const int c = cellIndex[size_t( p[0] / gridSize ) + size_t( p[1] / gridSize ) * w + size_t( p[2] / gridSize ) * w * w];
    if ( c == -1 ) continue;
    const uint32_t patch = partition[i] + 1;
    if ( count[c] == 0 ) {
      firstPatch[c] = patch, mixed[c] = 0;
      center[3 * c] = center[3 * c + 1] = center[3 * c + 2] = 0.f;
    } else if ( !mixed[c] && firstPatch[c] != patch ) {
      mixed[c] = 1;
    }
    for ( int k = 0; k < 3; ++k ) center[3 * c + k] += float( p[k] );
    ++count[c];
  }
  for ( int c = 0; c < cells; ++c )
    if ( count[c] )
      for ( int k = 0; k < 3; ++k ) center[3 * c + k] /= float( count[c] );
  // per boundary point (independent of one another: the grid is frozen)
  const int g2 = gridSize * 2, w3 = w * w * w;
  for ( size_t i = 0; i < n; ++i ) {
    int16_t* p = xyz + 3 * i;
    if ( nearBorder( p ) || boundary[i] != 1 ) continue;
    int S[3], idx[2][2][2];
    for ( int k = 0; k < 3; ++k ) S[k] = p[k] / gridSize + ( ( p[k] - ( p[k] / gridSize ) * gridSize ) < half ? -1 : 0 );
    bool other = false;
    for ( int dz = 0; dz < 2; ++dz )
      for ( int dy = 0; dy < 2; ++dy )
        for ( int dx = 0; dx < 2; ++dx ) {
          const int t     = ( S[0] + dx ) + ( S[1] + dy ) * w + ( S[2] + dz ) * w * w;
          idx[dz][dy][dx] = t;
          const int c     = cellIndex[t];
          if ( mixed[c] && count[c] != 0 ) other = true;
        }
    if ( !other ) continue;
    const double cur[3] = {double( p[0] ), double( p[1] ), double( p[2] )};
    int          W[3], Q[3];
    for ( int k = 0; k < 3; ++k ) W[k] = ( p[k] - S[k] * gridSize - half ) * 2 + 1, Q[k] = g2 - W[k];
    double sum[3] = {0.0, 0.0, 0.0};
    int    cnt    = 0;
    for ( int dz = 0; dz < 2; ++dz )
      for ( int dy = 0; dy < 2; ++dy )
        for ( int dx = 0; dx < 2; ++dx ) {
          const int c = cellIndex[idx[dz][dy][dx]];
          double    v[3] = {cur[0], cur[1], cur[2]};
          if ( ( ( dx == 0 && dy == 0 && dz == 0 ) || idx[dz][dy][dx] < w3 ) && count[c] > 0 )
            for ( int k = 0; k < 3; ++k ) v[k] = double( center[3 * c + k] );
          const int wgt = ( dx ? W[0] : Q[0] ) * ( dy ? W[1] : Q[1] ) * ( dz ? W[2] : Q[2] );
          for ( int k = 0; k < 3; ++k ) {
            v[k] *= double( wgt );
            sum[k] += v[k];
          }
          cnt += wgt * count[c];
        }
    for ( int k = 0; k < 3; ++k ) sum[k] /= double( g2 * g2 * g2 );
    cnt /= g2 * g2 * g2;
    double centroid[3], d[3];
    for ( int k = 0; k < 3; ++k ) centroid[k] = sum[k] * double( cnt ), d[k] = cur[k] * double( cnt ) - centroid[k];
    const double dist2 = ( d[0] * d[0] + d[1] * d[1] + d[2] * d[2] ) / double( cnt ) + 0.5;
    if ( dist2 >= double( std::max( int( threshold ), cnt ) * 2 ) ) {
      for ( int k = 0; k < 3; ++k ) p[k] = int16_t( double( int64_t( centroid[k] / double( cnt ) + 0.5 ) ) );
      boundary[i] = 3;
    }
  }
}

// ---- §8f-1, second stage: colour transfer onto the smoothed cloud, PCCPointSet3::transferColors16bitBP
// (L/PccLibCommon/source/PCCPointSet.cpp:1126-1485) with the arguments PCCEncoder::encode / PCCDecoder::decode pass
// (L/PccLibEncoder/source/PCCEncoder.cpp:656-672): filter type 1, search range 0, lossy attributes, 8 forward / 1 backward
// neighbours, distance-weighted averages, skip-if-identical forward only, distance offsets 4, geometry / colour distance limits
// beyond their caps (= unlimited), no colour-outlier exclusion. Only the target points the geometry smoothing moved (boundary
// type 3) get a new colour; `source` is the cloud before smoothing with its 16-bit colours, `target` the smoothed one.
void pcco_transfer_colors16_smoothed( const int16_t* srcXyz, const uint16_t* srcCol, size_t S, const int16_t* tgtXyz, uint16_t* tgtCol,
                                      const uint16_t* tgtBoundary, size_t T ) {
  if ( S == 0 || T == 0 ) return;
  Tree* treeS = static_cast<Tree*>( pcco_kdtree_build( srcXyz, S ) );
  Tree* treeT = static_cast<Tree*>( pcco_kdtree_build( tgtXyz, T ) );
  std::vector<uint16_t> refined( tgtCol, tgtCol + 3 * T );  // forward result per target point
  std::vector<uint32_t> partSrc;                             // the sampled source points, in sampling order (with repetitions)
  double                d[8];
  uint32_t              id[8];
  auto clip16 = []( double v ) { return uint16_t( v < 0.0 ? 0.0 : ( v > 65535.0 ? 65535.0 : v ) ); };
  for ( size_t t = 0; t < T; ++t ) {
    if ( tgtBoundary[t] != 3 ) continue;
    KnnSet rs( 8, id, d );
    treeS->search( rs, tgtXyz + 3 * t );
    const int cnt = rs.cnt;
    for ( int i = 0; i < cnt; ++i ) partSrc.push_back( id[i] );
    if ( d[0] < 0.0001 || cnt == 1 ) {
      for ( int k = 0; k < 3; ++k ) refined[3 * t + k] = srcCol[3 * size_t( id[0] ) + k];
    } else {
      double acc[3] = {0.0, 0.0, 0.0}, sum = 0.0;
      for ( int i = 0; i < cnt; ++i ) {
        const double wgt = 1 / ( d[i] + 4.0 );
        for ( int k = 0; k < 3; ++k ) acc[k] += srcCol[3 * size_t( id[i] ) + k] * wgt;
        sum += wgt;
      }
      for ( int k = 0; k < 3; ++k ) refined[3 * t + k] = clip16( std::round( acc[k] / sum ) );
    }
  }
  // backward: every sampled source point votes for its nearest target point if their colours are close
  struct Vote {
    double   dist;
    uint16_t c[3];
  };
  std::vector<std::vector<Vote>> votes( T );
  for ( uint32_t si : partSrc ) {
    KnnSet rs( 1, id, d );
    treeT->search( rs, srcXyz + 3 * size_t( si ) );
    if ( rs.cnt < 1 ) continue;
    const uint16_t* c  = srcCol + 3 * size_t( si );
    const uint16_t* tc = tgtCol + 3 * size_t( id[0] );
    if ( std::abs( int( c[0] ) - int( tc[0] ) ) < 40 && std::abs( int( c[1] ) - int( tc[1] ) ) < 40 && std::abs( int( c[2] ) - int( tc[2] ) ) < 40 )
      votes[id[0]].push_back( Vote{d[0], {c[0], c[1], c[2]}} );
  }
  for ( auto& v : votes ) std::sort( v.begin(), v.end(), []( Vote& a, Vote& b ) { return a.dist < b.dist; } );
  std::vector<uint16_t> out( tgtCol, tgtCol + 3 * T );
  for ( size_t t = 0; t < T; ++t ) {
    if ( tgtBoundary[t] != 3 ) continue;
    const auto& v = votes[t];
    if ( v.empty() ) {
      for ( int k = 0; k < 3; ++k ) out[3 * t + k] = refined[3 * t + k];
      continue;
    }
    double c2[3] = {0.0, 0.0, 0.0};
    if ( v.size() == 1 ) {
      for ( int k = 0; k < 3; ++k ) c2[k] = v[0].c[k];
    } else {
      double sum = 0.0;
      for ( auto& e : v ) {
        const double wgt = 1 / ( std::sqrt( e.dist ) + 4.0 );
        for ( int k = 0; k < 3; ++k ) c2[k] += ( e.c[k] * wgt );
        sum += wgt;
      }
      for ( int k = 0; k < 3; ++k ) c2[k] /= sum;
    }
    for ( int k = 0; k < 3; ++k ) {
      const double c1 = double( refined[3 * t + k] );
      double       v0 = std::round( 0.0 * c1 + 1.0 * c2[k] );  // fixWeight: w = 0 (m42538)
      v0              = v0 < 0.0 ? 0.0 : ( v0 > 65535.0 ? 65535.0 : v0 );
      out[3 * t + k]  = uint16_t( v0 );
    }
  }
  std::copy( out.begin(), out.end(), tgtCol );
  pcco_kdtree_free( treeS );
  pcco_kdtree_free( treeT );
}

// ---- §8f-1, the two ends of the chain around colorPointCloud (a gather of these values by pointToPixel):
// (1) YUV 4:2:0 (8 bit, as decoded) -> YUV 4:4:4 (16 bit): PCCInternalColorConverter "YUV420ToYUV444_8_0", the inverse conversion
//     of PCCVideoEncoder::compress / PCCVideoDecoder::decompress (L/PccLibColorConverter/source/PCCInternalColorConverter.cpp:
//     467-485, 595-610, 669-695; up-sampling filter 0, the default; inline filters L/PccLibColorConverter/include/
//     PCCInternalColorConverter.h:185-247: float accumulation in tap order). in: Y (W*H), U, V ((W/2)*(H/2)); out: 3 planes W*H.
void pcco_yuv420_to_yuv444_16( const uint8_t* yuv420, size_t W, size_t H, uint16_t* yuv444 ) {
  const size_t Q = W * H, w2 = W / 2, h2 = H / 2;
  auto toFloat = [&]( uint8_t v, bool chroma ) {  // YUVtoFloatYUV, one byte per sample
    const float f = float( ( 1.0 / 255. ) * double( int( v ) - ( chroma ? 128 : 0 ) ) );
    const float lo = chroma ? -0.5f : 0.f, hi = chroma ? 0.5f : 1.f;
    return f < lo ? lo : ( f > hi ? hi : f );
  };
  auto quantise16 = [&]( float v, bool chroma ) {  // floatYUVToYUV, two bytes per sample
    float r = std::round( float( 65535. * double( v ) + ( chroma ? 32768. : 0. ) ) );
    r       = r < 0.f ? 0.f : ( r > 65535.f ? 65535.f : r );
    return uint16_t( r );
  };
  for ( size_t i = 0; i < Q; ++i ) yuv444[i] = quantise16( toFloat( yuv420[i], false ), false );
  const float ver0[4] = {-8.0f, +64.0f, +216.0f, -16.0f}, hor1[4] = {-16.0f, +144.0f, +144.0f, -16.0f}, ver1[4] = {-16.0f, +216.0f, +64.0f, -8.0f};
  const float scale = 1.0f / float( 1 << 8 );
  auto clampI = []( long v, long lo, long hi ) { return v < lo ? lo : ( v > hi ? hi : v ); };
  for ( int c = 0; c < 2; ++c ) {
    const uint8_t*     src = yuv420 + Q + size_t( c ) * w2 * h2;
    std::vector<float> in( w2 * h2 ), tmp( w2 * H );
    for ( size_t i = 0; i < w2 * h2; ++i ) in[i] = toFloat( src[i], true );
    for ( size_t i = 0; i < h2; ++i )
      for ( size_t j = 0; j < w2; ++j ) {
        float a = 0, b = 0;
        for ( int t = 0; t < 4; ++t ) a += ver0[t] * in[size_t( clampI( long( i ) + t - 2, 0, long( h2 ) - 1 ) ) * w2 + j];
        for ( int t = 0; t < 4; ++t ) b += ver1[t] * in[size_t( clampI( long( i ) + 1 + t - 2, 0, long( h2 ) - 1 ) ) * w2 + j];
        tmp[( 2 * i ) * w2 + j]     = ( a + 0.f ) * scale;
        tmp[( 2 * i + 1 ) * w2 + j] = ( b + 0.f ) * scale;
      }
    uint16_t* dst = yuv444 + Q * ( 1 + c );
    for ( size_t i = 0; i < H; ++i )
      for ( size_t j = 0; j < w2; ++j ) {
        float a = 0, b = 0;
        a += 0.0f * tmp[i * w2 + size_t( clampI( long( j ) - 1, 0, long( w2 ) - 1 ) )];
        a += 256.0f * tmp[i * w2 + j];
        for ( int t = 0; t < 4; ++t ) b += hor1[t] * tmp[i * w2 + size_t( clampI( long( j ) + 1 + t - 2, 0, long( w2 ) - 1 ) )];
        dst[i * W + 2 * j]     = quantise16( ( a + 0.f ) * scale, true );
        dst[i * W + 2 * j + 1] = quantise16( ( b + 0.f ) * scale, true );
      }
  }
}
// (2) PCCPointSet3::convertYUV16ToRGB8 (L/PccLibCommon/include/PCCPointSet.h:133-166): the final 8-bit RGB of every point
void pcco_yuv16_to_rgb8( const uint16_t* yuv, size_t n, uint8_t* rgb ) {
  for ( size_t i = 0; i < n; ++i ) {
    const double wgt = 1.0 / 65535.0;
    double       y = wgt * double( yuv[3 * i] ), u = wgt * ( double( yuv[3 * i + 1] ) - 32768.0 ), v = wgt * ( double( yuv[3 * i + 2] ) - 32768.0 );
    y = std::min( std::max( y, 0.0 ), 1.0 ), u = std::min( std::max( u, -0.5 ), 0.5 ), v = std::min( std::max( v, -0.5 ), 0.5 );
    const double c[3] = {y + 1.57480 * v, y - 0.18733 * u - 0.46813 * v, y + 1.85563 * u};
    for ( int k = 0; k < 3; ++k ) {
      const double r = std::round( c[k] * 255 );
      rgb[3 * i + k] = uint8_t( r < 0.0 ? 0.0 : ( r > 255.0 ? 255.0 : r ) );
    }
  }
}

}  // extern "C"
