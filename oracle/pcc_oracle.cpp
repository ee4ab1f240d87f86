// pcc_oracle.cpp — CPU oracle (TEST INFRASTRUCTURE ONLY, see pcc_oracle.h).
//
// A from-scratch restatement of the TMC2 v24.0 hot-path algorithms, written for clarity, single-threaded,
// fp64 evaluated in the reference's operation order (build with -ffp-contract=off on x86-64).
// Citations are relative to the reference checkout: L/ = source/lib/, NF = dependencies/nanoflann/nanoflann.hpp.
#include "pcc_oracle.h"

#include <algorithm>
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
struct OPatch {
  pccb200_patch        m;
  std::vector<int16_t> depth[2];
  std::vector<uint8_t> occ;
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

}  // extern "C"
