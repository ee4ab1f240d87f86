// normals.cu — per-point PCA normals in fp64, evaluated in the reference's operation order.
// Replaces PCCNormalsGenerator3::computeNormal(s) (PccLibEncoder/source/PCCNormalsGenerator.cpp:71-185) and
// PCCDiagonalize (PccLibCommon/include/PCCMath.h:505-598).  Compiled with -fmad=false: every product and sum
// rounds separately, exactly as the reference's x86-64 SSE2 build does, because the normals feed an argmax.
#include "stages.cuh"

namespace pccb200 {

namespace {

// Symmetric 3x3 eigen-decomposition by at most 24 quaternion Jacobi rotations.
// A is given by its upper triangle a00 a01 a02 a11 a12 a22. Returns Q (columns = eigenvectors) and diag(D).
__device__ void jacobiEigen( const double A[6], double Q[3][3], double Dd[3] ) {
  const double a00 = A[0], a01 = A[1], a02 = A[2], a11 = A[3], a12 = A[4], a22 = A[5];
  double       q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 1.0;
  double       D[3][3];
  for ( int step = 0; step < 24; ++step ) {
    const double xx = q0 * q0, yy = q1 * q1, zz = q2 * q2, ww = q3 * q3;
    Q[0][0] = ( xx - yy - zz + ww );
    Q[1][1] = ( -xx + yy - zz + ww );
    Q[2][2] = ( -xx - yy + zz + ww );
    double a = q0 * q1, b = q2 * q3;
    Q[1][0] = 2.0 * ( a + b ), Q[0][1] = 2.0 * ( a - b );
    a = q0 * q2, b = q1 * q3;
    Q[2][0] = 2.0 * ( a - b ), Q[0][2] = 2.0 * ( a + b );
    a = q1 * q2, b = q0 * q3;
    Q[2][1] = 2.0 * ( a + b ), Q[1][2] = 2.0 * ( a - b );
    double AQ[3][3];
#pragma unroll
    for ( int c = 0; c < 3; ++c ) {
      AQ[0][c] = Q[0][c] * a00 + Q[1][c] * a01 + Q[2][c] * a02;
      AQ[1][c] = Q[0][c] * a01 + Q[1][c] * a11 + Q[2][c] * a12;
      AQ[2][c] = Q[0][c] * a02 + Q[1][c] * a12 + Q[2][c] * a22;
    }
#pragma unroll
    for ( int r = 0; r < 3; ++r )
#pragma unroll
      for ( int c = 0; c < 3; ++c ) D[r][c] = AQ[0][r] * Q[0][c] + AQ[1][r] * Q[1][c] + AQ[2][r] * Q[2][c];
    Dd[0] = D[0][0], Dd[1] = D[1][1], Dd[2] = D[2][2];
    const double o0 = D[1][2], o1 = D[0][2], o2 = D[0][1];
    const double m0 = fabs( o0 ), m1 = fabs( o1 ), m2 = fabs( o2 );
    const int    k0 = ( m0 > m1 && m0 > m2 ) ? 0 : ( m1 > m2 ) ? 1 : 2;
    const double ok = k0 == 0 ? o0 : ( k0 == 1 ? o1 : o2 );
    if ( ok == 0.0 ) break;
    // k1 = (k0+1)%3, k2 = (k0+2)%3
    const double dk1 = k0 == 0 ? D[1][1] : ( k0 == 1 ? D[2][2] : D[0][0] );
    const double dk2 = k0 == 0 ? D[2][2] : ( k0 == 1 ? D[0][0] : D[1][1] );
    double       thet = ( dk2 - dk1 ) / ( 2.0 * ok );
    const double sgn  = ( thet > 0.0 ) ? 1.0 : -1.0;
    thet *= sgn;
    const double t = sgn / ( thet + ( ( thet < 1.E6 ) ? sqrt( thet * thet + 1.0 ) : thet ) );
    const double c = 1.0 / sqrt( t * t + 1.0 );
    if ( c == 1.0 ) break;
    double jk = sgn * sqrt( ( 1.0 - c ) / 2.0 );
    jk *= -1.0;
    const double j3 = sqrt( 1.0 - jk * jk );
    if ( j3 == 1.0 ) break;
    const double j0 = k0 == 0 ? jk : 0.0, j1 = k0 == 1 ? jk : 0.0, j2 = k0 == 2 ? jk : 0.0;
    // in-place product: each component sees the already-updated earlier ones (as the reference does)
    q0 = ( q3 * j0 + q0 * j3 + q1 * j2 - q2 * j1 );
    q1 = ( q3 * j1 - q0 * j2 + q1 * j3 + q2 * j0 );
    q2 = ( q3 * j2 + q0 * j1 - q1 * j0 + q2 * j3 );
    q3 = ( q3 * j3 - q0 * j0 - q1 * j1 - q2 * j2 );
    const double mq = sqrt( q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3 );
    q0 /= mq, q1 /= mq, q2 /= mq, q3 /= mq;
  }
}

__global__ void __launch_bounds__( 128 )
    kNormals( const short4* __restrict__ pts, const uint32_t* __restrict__ nbr, int k, int n, double* __restrict__ normals ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  const uint32_t* row = nbr + size_t( i ) * k;
  int             cnt = 0;
  double          bx = 0.0, by = 0.0, bz = 0.0;
  for ( int j = 0; j < k; ++j ) {
    const uint32_t o = row[j];
    if ( o == 0xFFFFFFFFu ) break;
    const short4 p = pts[o];
    bx = bx + double( p.x ), by = by + double( p.y ), bz = bz + double( p.z );
    ++cnt;
  }
  double nx = 0.0, ny = 0.0, nz = 0.0;
  if ( cnt > 1 ) {
    bx /= double( cnt ), by /= double( cnt ), bz /= double( cnt );
    double C[6] = {0, 0, 0, 0, 0, 0};  // 00 01 02 11 12 22
    for ( int j = 0; j < cnt; ++j ) {
      const short4 p = pts[row[j]];
      const double x = double( p.x ) - bx, y = double( p.y ) - by, z = double( p.z ) - bz;
      C[0] += x * x, C[3] += y * y, C[5] += z * z;
      C[1] += x * y, C[2] += x * z, C[4] += y * z;
    }
    const double den = double( cnt ) - 1.0;
#pragma unroll
    for ( int e = 0; e < 6; ++e ) C[e] /= den;
    double Q[3][3], D[3];
    jacobiEigen( C, Q, D );
    const double e0 = fabs( D[0] ), e1 = fabs( D[1] ), e2 = fabs( D[2] );
    const int    col = ( e0 < e1 && e0 < e2 ) ? 0 : ( e1 < e2 ) ? 1 : 2;
    nx = Q[0][col], ny = Q[1][col], nz = Q[2][col];
  }
  // flip toward the view point (origin): normal * (0 - p) < 0
  const short4 me = pts[i];
  const double vx = 0.0 - double( me.x ), vy = 0.0 - double( me.y ), vz = 0.0 - double( me.z );
  const double dt = nx * vx + ny * vy + nz * vz;
  if ( dt < 0.0 ) nx = -nx, ny = -ny, nz = -nz;
  normals[3 * size_t( i ) + 0] = nx;
  normals[3 * size_t( i ) + 1] = ny;
  normals[3 * size_t( i ) + 2] = nz;
}

// PCCPatchSegmenter3::initialSegmentation (PccLibEncoder/source/PCCPatchSegmenter.cpp:226-265):
// orientation 0 is scored without its axis weight; the first maximum wins.
__global__ void kInitialSegmentation( const double* __restrict__ normals, int n, double w0, double w1, double w2,
                                      uint8_t* __restrict__ partition ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= n ) return;
  const double x = normals[3 * size_t( i )], y = normals[3 * size_t( i ) + 1], z = normals[3 * size_t( i ) + 2];
  // dot products with the six axis orientations keep the reference's three-term form (x*ox + y*oy + z*oz)
  const double d[6] = {x * 1.0 + y * 0.0 + z * 0.0,  x * 0.0 + y * 1.0 + z * 0.0,  x * 0.0 + y * 0.0 + z * 1.0,
                       x * -1.0 + y * 0.0 + z * 0.0, x * 0.0 + y * -1.0 + z * 0.0, x * 0.0 + y * 0.0 + z * -1.0};
  const double w[6] = {w0, w1, w2, w0, w1, w2};
  int          best = 0;
  double       bs   = d[0];
#pragma unroll
  for ( int j = 1; j < 6; ++j ) {
    const double sc = d[j] * w[j];
    if ( sc > bs ) bs = sc, best = j;
  }
  partition[i] = uint8_t( best );
}

}  // namespace

void computeNormals( const short4* pts, const uint32_t* nbr, int k, size_t n, double* normals, cudaStream_t s ) {
  if ( n == 0 ) return;
  kNormals<<<divUp( n, 128 ), 128, 0, s>>>( pts, nbr, k, int( n ), normals );
  PCC_LAUNCH_CHECK();
}

void initialSegmentation( const double* normals, size_t n, const double w[3], uint8_t* partition, cudaStream_t s ) {
  if ( n == 0 ) return;
  kInitialSegmentation<<<divUp( n, 256 ), 256, 0, s>>>( normals, int( n ), w[0], w[1], w[2], partition );
  PCC_LAUNCH_CHECK();
}

}  // namespace pccb200
