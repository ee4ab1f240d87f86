// common.cuh — shared device/host helpers for libpccb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>

#include "../../include/pccb200.h"

namespace pccb200 {

struct CudaError {
  cudaError_t code;
  const char* file;
  int         line;
};

#define PCC_CUDA( expr )                                                          \
  do {                                                                            \
    cudaError_t _e = ( expr );                                                    \
    if ( _e != cudaSuccess ) throw ::pccb200::CudaError{ _e, __FILE__, __LINE__ }; \
  } while ( 0 )

#define PCC_LAUNCH_CHECK() PCC_CUDA( cudaGetLastError() )

static inline int divUp( size_t a, size_t b ) { return int( ( a + b - 1 ) / b ); }

// Host wait for everything queued on `s` so far. BLOCKING (the thread sleeps until the device raises the event) instead of
// cudaStreamSynchronize's default spin: a GOF keeps one host thread per frame and several GOFs are in flight, far more threads
// than the host has cores - a spinning waiter would burn the time slice a thread with kernels to launch is waiting for.
struct BlockingEvent {
  cudaEvent_t ev  = nullptr;
  int         dev = -1;
  ~BlockingEvent() {
    if ( ev ) cudaEventDestroy( ev );
  }
};
static inline bool spinWaits() {  // PCCB200_SPIN_WAIT=1: the driver's default (lowest latency when threads <= cores)
  static const bool spin = [] {
    const char* e = getenv( "PCCB200_SPIN_WAIT" );
    return e && e[0] == '1';
  }();
  return spin;
}
static inline void streamWait( cudaStream_t s ) {
  if ( spinWaits() ) {
    PCC_CUDA( cudaStreamSynchronize( s ) );
    return;
  }
  thread_local BlockingEvent be;
  int                        dev = 0;
  PCC_CUDA( cudaGetDevice( &dev ) );
  if ( !be.ev || be.dev != dev ) {
    if ( be.ev ) cudaEventDestroy( be.ev );
    be.ev = nullptr;
    PCC_CUDA( cudaEventCreateWithFlags( &be.ev, cudaEventBlockingSync | cudaEventDisableTiming ) );
    be.dev = dev;
  }
  PCC_CUDA( cudaEventRecord( be.ev, s ) );
  PCC_CUDA( cudaEventSynchronize( be.ev ) );
}

// Grow-only device buffer; reused across frames so the steady state does no cudaMalloc.
template <typename T>
struct DevBuf {
  T*     p   = nullptr;
  size_t cap = 0;
  DevBuf()   = default;
  DevBuf( const DevBuf& )            = delete;
  DevBuf& operator=( const DevBuf& ) = delete;
  DevBuf( DevBuf&& o ) noexcept : p( o.p ), cap( o.cap ) { o.p = nullptr, o.cap = 0; }
  DevBuf& operator=( DevBuf&& o ) noexcept {
    if ( this != &o ) {
      release();
      p = o.p, cap = o.cap;
      o.p = nullptr, o.cap = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if ( p ) cudaFree( p );
    p   = nullptr;
    cap = 0;
  }
  // contents are NOT preserved
  void reserve( size_t n ) {
    if ( n <= cap ) return;
    release();
    size_t want = n + n / 8 + 64;
    PCC_CUDA( cudaMalloc( &p, want * sizeof( T ) ) );
    cap = want;
  }
  // contents preserved
  void grow( size_t n, cudaStream_t s ) {
    if ( n <= cap ) return;
    size_t want = n + n / 2 + 64;
    T*     q    = nullptr;
    PCC_CUDA( cudaMalloc( &q, want * sizeof( T ) ) );
    if ( p ) {
      PCC_CUDA( cudaMemcpyAsync( q, p, cap * sizeof( T ), cudaMemcpyDeviceToDevice, s ) );
      streamWait( s );
      cudaFree( p );
    }
    p   = q;
    cap = want;
  }
  operator T*() const { return p; }
};

// Pinned host staging buffer (grow-only).
template <typename T>
struct PinBuf {
  T*     p   = nullptr;
  size_t cap = 0;
  ~PinBuf() {
    if ( p ) cudaFreeHost( p );
  }
  void reserve( size_t n ) {
    if ( n <= cap ) return;
    if ( p ) cudaFreeHost( p );
    PCC_CUDA( cudaMallocHost( &p, ( n + 64 ) * sizeof( T ) ) );
    cap = n + 64;
  }
  operator T*() const { return p; }
};

// ---- device-wide exclusive scan of uint32 (scan.cu) ---------------------------------------------------
// out[i] = sum_{j<i} in[j], out has n+1 entries (out[n] = total). tmp must hold scanTmpElems(n) uint32.
size_t scanTmpElems( size_t n );
void   exclusiveScanU32( const uint32_t* in, uint32_t* out, size_t n, uint32_t* tmp, cudaStream_t s );

// Stage / kernel timing with CUDA events on the launching stream (enabled per context; see pccb200_profile_*).
struct Profiler {
  struct Span {
    const char* name;
    cudaEvent_t a, b;
  };
  bool              enabled = false;
  std::vector<Span> spans;
  struct Result {
    const char* first;   // stage name
    float       second;  // duration (ms)
    float       start;   // start offset (ms) from the origin event, or -1
  };
  std::vector<Result> results;  // filled by collect()
  cudaEvent_t         origin = nullptr;  // optional common time origin (recorded by the owner)
  void begin( const char* name, cudaStream_t s ) {
    if ( !enabled ) return;
    Span sp{ name, nullptr, nullptr };
    cudaEventCreate( &sp.a ), cudaEventCreate( &sp.b );
    cudaEventRecord( sp.a, s );
    spans.push_back( sp );
  }
  void end( cudaStream_t s ) {  // closes the most recent open span
    if ( !enabled ) return;
    // spans are strictly nested/sequential in this code base: record on the last span not yet closed
    for ( size_t i = spans.size(); i-- > 0; )
      if ( !closed[i] ) {
        cudaEventRecord( spans[i].b, s );
        closed[i] = true;
        return;
      }
  }
  std::vector<bool> closed;
  void              collect( cudaStream_t s ) {
    if ( !enabled ) return;
    streamWait( s );
    for ( size_t i = 0; i < spans.size(); ++i ) {
      float ms = 0, st = -1.f;
      if ( closed[i] ) cudaEventElapsedTime( &ms, spans[i].a, spans[i].b );
      if ( origin && cudaEventElapsedTime( &st, origin, spans[i].a ) != cudaSuccess ) st = -1.f, cudaGetLastError();
      results.push_back( Result{ spans[i].name, ms, st } );
      cudaEventDestroy( spans[i].a ), cudaEventDestroy( spans[i].b );
    }
    spans.clear(), closed.clear();
  }
};
struct ProfScope {
  Profiler*    p;
  cudaStream_t s;
  ProfScope( Profiler* prof, const char* name, cudaStream_t st ) : p( prof ), s( st ) {
    if ( p && p->enabled ) {
      p->begin( name, s );
      p->closed.push_back( false );
    }
  }
  ~ProfScope() {
    if ( p && p->enabled ) p->end( s );
  }
};

}  // namespace pccb200
