// Micro-benchmark: aggregate kernel-launch throughput of T host threads, each on its own stream (one context).
// Usage: launch_rate [launches_per_thread]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
__global__ void kEmpty( int* p ) { if ( p && threadIdx.x == 12345 ) *p = 1; }
int main( int argc, char** argv ) {
  const int per = argc > 1 ? atoi( argv[1] ) : 4000;
  cudaFree( 0 );
  for ( int mode = 0; mode < 2; ++mode )
    for ( int T : {1, 2, 4, 8, 16, 32} ) {
      std::vector<cudaStream_t> st( T );
      for ( auto& s : st ) cudaStreamCreateWithFlags( &s, cudaStreamNonBlocking );
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for ( int t = 0; t < T; ++t )
        th.emplace_back( [&, t]() {
          for ( int i = 0; i < per; ++i ) {
            kEmpty<<<1, 32, 0, st[t]>>>( nullptr );
            if ( mode == 1 && i % 10 == 9 ) cudaStreamSynchronize( st[t] );
          }
          cudaStreamSynchronize( st[t] );
        } );
      for ( auto& x : th ) x.join();
      const double sec = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
      printf( "mode %d (%s) threads %2d: %8.0f launches/s aggregate (%.2f us per launch per thread)\n", mode, mode ? "sync every 10" : "no syncs", T,
              double( T ) * per / sec, sec / per * 1e6 );
      for ( auto& s : st ) cudaStreamDestroy( s );
    }
  return 0;
}
