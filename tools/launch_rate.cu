// Micro-benchmark: aggregate kernel-launch throughput of T host threads, each on its own stream (one context).
// Usage: launch_rate [launches_per_thread]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
__global__ void kEmpty( int* p ) { if ( p && threadIdx.x == 12345 ) *p = 1; }
int main( int argc, char** argv ) {
  const int per = argc > 1 ? atoi( argv[1] ) : 4000;
  cudaFree( 0 );
  std::mutex gate;
  const char* names[] = {"no syncs", "sync every 10", "sync every 10, bursts of 10 under one process-wide mutex", "sync every 10, mutex per launch"};
  for ( int mode = 0; mode < 4; ++mode )
    for ( int T : {1, 4, 16, 32, 64, 128} ) {
      std::vector<cudaStream_t> st( T );
      for ( auto& s : st ) cudaStreamCreateWithFlags( &s, cudaStreamNonBlocking );
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for ( int t = 0; t < T; ++t )
        th.emplace_back( [&, t]() {
          if ( mode == 2 ) {
            for ( int i = 0; i < per; i += 10 ) {
              {
                std::lock_guard<std::mutex> lk( gate );
                for ( int j = 0; j < 10; ++j ) kEmpty<<<1, 32, 0, st[t]>>>( nullptr );
              }
              cudaStreamSynchronize( st[t] );
            }
          } else
          for ( int i = 0; i < per; ++i ) {
            if ( mode == 3 ) {
              std::lock_guard<std::mutex> lk( gate );
              kEmpty<<<1, 32, 0, st[t]>>>( nullptr );
            } else
              kEmpty<<<1, 32, 0, st[t]>>>( nullptr );
            if ( mode != 0 && i % 10 == 9 ) cudaStreamSynchronize( st[t] );
          }
          cudaStreamSynchronize( st[t] );
        } );
      for ( auto& x : th ) x.join();
      const double sec = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
      printf( "mode %d (%s) threads %2d: %8.0f launches/s aggregate (%.2f us per launch per thread)\n", mode, names[mode], T,
              double( T ) * per / sec, sec / per * 1e6 );
      for ( auto& s : st ) cudaStreamDestroy( s );
    }
  return 0;
}
