// What sm_100a issues for the 8-bit x4 SIMD intrinsics the north star's "8-bit lanes" would need
// (per-byte saturating add and per-byte max), next to the 16-bit x2 DPX forms the scan kernel uses and
// dp4a.  Compile only:  nvcc -gencode arch=compute_100a,code=sm_100a -cubin -o simd8.cubin simd8_sass.cu
//                       cuobjdump -sass simd8.cubin        (summarised in profiles/r1_simd8_sass.txt)
#include <cuda_runtime.h>
extern "C" __global__ void k_vaddss4(unsigned *o, const unsigned *a, const unsigned *b) { o[threadIdx.x] = __vaddss4(a[threadIdx.x], b[threadIdx.x]); }
extern "C" __global__ void k_vmaxs4(unsigned *o, const unsigned *a, const unsigned *b) { o[threadIdx.x] = __vmaxs4(a[threadIdx.x], b[threadIdx.x]); }
extern "C" __global__ void k_vmaxu4(unsigned *o, const unsigned *a, const unsigned *b) { o[threadIdx.x] = __vmaxu4(a[threadIdx.x], b[threadIdx.x]); }
extern "C" __global__ void k_vsubus4(unsigned *o, const unsigned *a, const unsigned *b) { o[threadIdx.x] = __vsubus4(a[threadIdx.x], b[threadIdx.x]); }
extern "C" __global__ void k_dp4a(int *o, const int *a, const int *b) { o[threadIdx.x] = __dp4a(a[threadIdx.x], b[threadIdx.x], o[threadIdx.x]); }
extern "C" __global__ void k_vmaxs2(unsigned *o, const unsigned *a, const unsigned *b) { o[threadIdx.x] = __vmaxs2(a[threadIdx.x], b[threadIdx.x]); }
extern "C" __global__ void k_viaddmax_s16x2_relu(unsigned *o, const unsigned *a, const unsigned *b) { o[threadIdx.x] = __viaddmax_s16x2_relu(a[threadIdx.x], b[threadIdx.x], o[threadIdx.x]); }
extern "C" __global__ void k_vimax3_s16x2_relu(unsigned *o, const unsigned *a, const unsigned *b) { o[threadIdx.x] = __vimax3_s16x2_relu(a[threadIdx.x], b[threadIdx.x], o[threadIdx.x]); }
