// TEST-ONLY: mpeg-pcc-tmc2_b200/csrc/stdsort.cuh against std::sort itself (the libstdc++ of this toolchain) on tie-heavy inputs.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

#include "../mpeg-pcc-tmc2_b200/csrc/stdsort.cuh"

struct Item {
  double   key;
  uint32_t id;
};

extern "C" int stdsort_check( int maxLen, int repeats, int keyRange, unsigned seed, long* comparisons ) {
  std::mt19937 rng( seed );
  long         cmpA = 0, cmpB = 0;
  for ( int len = 0; len <= maxLen; ++len )
    for ( int r = 0; r < repeats; ++r ) {
      std::vector<Item> a( len );
      for ( int i = 0; i < len; ++i ) a[i] = Item{double( rng() % unsigned( keyRange ) ), uint32_t( i )};
      if ( r % 4 == 1 ) std::sort( a.begin(), a.end(), []( const Item& x, const Item& y ) { return x.key < y.key || ( x.key == y.key && x.id < y.id ); } );  // pre-sorted
      if ( r % 4 == 2 ) std::reverse( a.begin(), a.end() );
      std::vector<Item> b = a;
      std::sort( a.begin(), a.end(), [&]( Item& x, Item& y ) { ++cmpA; return x.key < y.key; } );
      pccb200::stdsort::sort( b.data(), b.data() + len, [&]( const Item& x, const Item& y ) { ++cmpB; return x.key < y.key; } );
      for ( int i = 0; i < len; ++i )
        if ( a[i].id != b[i].id ) return len * 1000 + r + 1;  // first mismatch
      // the depth-limit fallback on its own: std::partial_sort( first, last, last ) is make_heap + sort_heap, as in introsort
      std::vector<Item> c( len ), d;
      for ( int i = 0; i < len; ++i ) c[i] = Item{double( rng() % unsigned( keyRange ) ), uint32_t( i )};
      d = c;
      std::partial_sort( c.begin(), c.end(), c.end(), []( const Item& x, const Item& y ) { return x.key < y.key; } );
      pccb200::stdsort::heapSort( d.data(), d.data() + len, []( const Item& x, const Item& y ) { return x.key < y.key; } );
      for ( int i = 0; i < len; ++i )
        if ( c[i].id != d[i].id ) return -( len * 1000 + r + 1 );
    }
  if ( comparisons ) comparisons[0] = cmpA, comparisons[1] = cmpB;
  return 0;
}
