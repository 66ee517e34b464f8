// Single translation unit of libist_b200.so (C ABI in include/ist_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC ist_b200.cu
#include "plan_impl.cuh"
#include "ops_impl.cuh"
#include "lbfgs_impl.cuh"
#include "image_impl.cuh"
