// stdsort.cuh — the permutation libstdc++'s std::sort produces, reproduced step by step, usable on the host and in device code.
//
// Why: the reference sorts small per-point candidate lists with std::sort on a key that has ties
// (PccLibCommon/source/PCCPointSet.cpp:955 and :1291-1295: colour votes by distance) and then sums floating-point terms in the
// sorted order. std::sort is not stable; which of two equal keys comes first is an artefact of the algorithm. Up to 16 elements
// libstdc++ runs a plain insertion sort (stable - what color.cu relies on); above 16 it is introsort: median-of-three quicksort
// down to 16-element runs, heap sort when the depth limit 2*floor(log2 n) is hit, then one final insertion pass. Bit-exact
// parity of the sums therefore needs exactly this sequence of comparisons and moves (GCC 13 libstdc++, bits/stl_algo.h and
// bits/stl_heap.h; the algorithm is unchanged since GCC 4.x). tests/test_stdsort.py checks it against std::sort itself on
// tie-heavy inputs of every length up to a few hundred.
#pragma once
#if defined( __CUDACC__ )
#define PCC_HD __host__ __device__
#else
#define PCC_HD
#endif

namespace pccb200 {
namespace stdsort {

template <class T>
PCC_HD inline void swapT( T& a, T& b ) {
  T t = a;
  a   = b;
  b   = t;
}

template <class T, class Less>
PCC_HD inline void unguardedLinearInsert( T* last, Less less ) {
  T  val  = *last;
  T* next = last - 1;
  while ( less( val, *next ) ) {
    *last = *next;
    last  = next;
    --next;
  }
  *last = val;
}
template <class T, class Less>
PCC_HD inline void insertionSort( T* first, T* last, Less less ) {
  if ( first == last ) return;
  for ( T* i = first + 1; i != last; ++i ) {
    if ( less( *i, *first ) ) {
      T val = *i;
      for ( T* p = i; p != first; --p ) *p = *( p - 1 );
      *first = val;
    } else {
      unguardedLinearInsert( i, less );
    }
  }
}
template <class T, class Less>
PCC_HD inline void adjustHeap( T* first, long hole, long len, T value, Less less ) {
  const long top    = hole;
  long       second = hole;
  while ( second < ( len - 1 ) / 2 ) {
    second = 2 * ( second + 1 );
    if ( less( first[second], first[second - 1] ) ) --second;
    first[hole] = first[second];
    hole        = second;
  }
  if ( ( len & 1 ) == 0 && second == ( len - 2 ) / 2 ) {
    second      = 2 * ( second + 1 );
    first[hole] = first[second - 1];
    hole        = second - 1;
  }
  long parent = ( hole - 1 ) / 2;  // __push_heap
  while ( hole > top && less( first[parent], value ) ) {
    first[hole] = first[parent];
    hole        = parent;
    parent      = ( hole - 1 ) / 2;
  }
  first[hole] = value;
}
template <class T, class Less>
PCC_HD inline void heapSort( T* first, T* last, Less less ) {  // __partial_sort( first, last, last ): make_heap + sort_heap
  const long len = long( last - first );
  if ( len >= 2 ) {
    for ( long parent = ( len - 2 ) / 2;; --parent ) {
      T value = first[parent];
      adjustHeap( first, parent, len, value, less );
      if ( parent == 0 ) break;
    }
  }
  while ( last - first > 1 ) {
    --last;
    T value = *last;
    *last   = *first;
    adjustHeap( first, 0L, long( last - first ), value, less );
  }
}
template <class T, class Less>
PCC_HD inline void moveMedianToFirst( T* result, T* a, T* b, T* c, Less less ) {
  if ( less( *a, *b ) ) {
    if ( less( *b, *c ) ) swapT( *result, *b );
    else if ( less( *a, *c ) ) swapT( *result, *c );
    else swapT( *result, *a );
  } else if ( less( *a, *c ) ) {
    swapT( *result, *a );
  } else if ( less( *b, *c ) ) {
    swapT( *result, *c );
  } else {
    swapT( *result, *b );
  }
}
template <class T, class Less>
PCC_HD inline T* unguardedPartition( T* first, T* last, T* pivot, Less less ) {
  for ( ;; ) {
    while ( less( *first, *pivot ) ) ++first;
    --last;
    while ( less( *pivot, *last ) ) --last;
    if ( !( first < last ) ) return first;
    swapT( *first, *last );
    ++first;
  }
}

// std::sort( first, last, less ) for random-access ranges of trivially copyable T
template <class T, class Less>
PCC_HD inline void sort( T* first, T* last, Less less ) {
  const long n = long( last - first );
  if ( n <= 0 ) return;
  // __introsort_loop with depth limit 2 * floor( log2( n ) ); the recursion on the right part becomes an explicit stack
  int lg = 0;
  while ( ( 1L << ( lg + 1 ) ) <= n ) ++lg;
  struct Range {
    T*  first;
    T*  last;
    int depth;
  };
  Range stack[64];
  int   sp    = 0;
  stack[sp++] = Range{first, last, 2 * lg};
  while ( sp > 0 ) {
    Range r = stack[--sp];
    while ( r.last - r.first > 16 ) {
      if ( r.depth == 0 ) {
        heapSort( r.first, r.last, less );
        break;
      }
      --r.depth;
      T* mid = r.first + ( r.last - r.first ) / 2;
      moveMedianToFirst( r.first, r.first + 1, mid, r.last - 1, less );
      T* cut = unguardedPartition( r.first + 1, r.last, r.first, less );
      // the library recurses into [cut, last) first and then loops on [first, cut): the two parts are disjoint, so the order in
      // which they are finished does not change the result; push the right part, continue with the left
      if ( sp < 64 ) stack[sp++] = Range{cut, r.last, r.depth};
      r.last = cut;
    }
  }
  // __final_insertion_sort
  if ( n > 16 ) {
    insertionSort( first, first + 16, less );
    for ( T* i = first + 16; i != last; ++i ) unguardedLinearInsert( i, less );
  } else {
    insertionSort( first, last, less );
  }
}

}  // namespace stdsort
}  // namespace pccb200
