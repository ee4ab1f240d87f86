// scratch.cu — per-device pool of FrameScratch sets (stages.cuh). A frame keeps ~0.3 GB while its orientation walk runs but
// needs ~1.2 GB more inside the data-parallel stage groups before and after it; leasing that part lets a B200 hold several GOFs
// of walks at once (the walk is a single-warp latency chain per frame: throughput comes from frames in flight).
#include <stdlib.h>

#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>

#include "stages.cuh"

namespace pccb200 {

namespace {

struct Pool {
  std::mutex                                 m;
  std::condition_variable                    cv;
  std::vector<std::unique_ptr<FrameScratch>> all;
  std::vector<FrameScratch*>                 idle;
  int                                        limit = 0;
};

std::mutex           gPoolsMutex;
std::map<int, Pool*> gPools;  // never destroyed: the CUDA context may be gone before static destructors run

Pool& poolOf( int device ) {
  std::lock_guard<std::mutex> lk( gPoolsMutex );
  Pool*&                      p = gPools[device];
  if ( !p ) {
    p             = new Pool();
    const char* e = getenv( "PCCB200_SCRATCH_SETS" );
    p->limit      = e && atoi( e ) > 0 ? atoi( e ) : 24;
  }
  return *p;
}

}  // namespace

FrameScratch* acquireFrameScratch( int device ) {
  Pool&                        p = poolOf( device );
  std::unique_lock<std::mutex> lk( p.m );
  for ( ;; ) {
    if ( !p.idle.empty() ) {
      FrameScratch* s = p.idle.back();
      p.idle.pop_back();
      return s;
    }
    if ( int( p.all.size() ) < p.limit ) {
      p.all.emplace_back( new FrameScratch() );
      return p.all.back().get();
    }
    p.cv.wait( lk );
  }
}

void releaseFrameScratch( int device, FrameScratch* s ) {
  if ( !s ) return;
  Pool& p = poolOf( device );
  {
    std::lock_guard<std::mutex> lk( p.m );
    p.idle.push_back( s );
  }
  p.cv.notify_one();
}

int setFrameScratchSets( int device, int count ) {
  Pool&                       p = poolOf( device );
  std::lock_guard<std::mutex> lk( p.m );
  const int                   before = p.limit;
  if ( count >= 1 ) p.limit = count;
  p.cv.notify_all();
  return before;
}

}  // namespace pccb200
