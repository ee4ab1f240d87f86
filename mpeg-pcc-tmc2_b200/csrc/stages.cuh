// stages.cuh — internal stage interfaces of libpccb200 (device pointers, one stream per frame).
#pragma once
#include "common.cuh"
#include "kdtree.cuh"

namespace pccb200 {

// normals.cu
void computeNormals( const short4* pts, const uint32_t* nbr, int k, size_t n, double* normals, cudaStream_t s );
void initialSegmentation( const double* normals, size_t n, const double w[3], uint8_t* partition, cudaStream_t s );

// util.cu
void packXyz( const int16_t* xyz3, size_t n, short4* out, cudaStream_t s );  // device int16 AoS (n x 3) -> short4

}  // namespace pccb200
