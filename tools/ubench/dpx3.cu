// Issue-rate probe: DPX s16x2 ops with 1, 2 or 3 *varying* register sources, and the SW cell
// sequence as the scan kernel issues it.  Output: warp instructions / clk / SMSP.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned int u32;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)
__device__ __forceinline__ u32 hadd2_u(u32 a, u32 b) { u32 r; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int OP, int CH>
__global__ void __launch_bounds__(512, 1) k(u32 *out, const u32 *in, int iters, long long *cyc)
{
  u32 x[CH], y[CH], z[CH], w[CH];
  const u32 c0 = in[0], c1 = in[1], c2 = in[2];
#pragma unroll
  for (int i = 0; i < CH; i++) { x[i] = in[4 + i] + threadIdx.x; y[i] = in[20 + i] ^ threadIdx.x; z[i] = in[36 + i] + 3 * threadIdx.x; w[i] = in[52 + i]; }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < CH; i++)
      {
        if (OP == 0) x[i] = __vimax3_s16x2_relu(x[i], c0, c1);                 // 1 varying
        if (OP == 1) x[i] = __vimax3_s16x2_relu(x[i], y[i], c1);               // 2 varying
        if (OP == 2) { x[i] = __vimax3_s16x2_relu(x[i], y[i], z[i]); y[i] = __vimax3_s16x2_relu(y[i], z[i], x[i]); z[i] = __vimax3_s16x2_relu(z[i], x[i], y[i]); }  // 3 varying
        if (OP == 3) x[i] = __viaddmax_s16x2_relu(x[i], c0, y[i]);             // add const, max var
        if (OP == 4) { x[i] = __viaddmax_s16x2_relu(x[i], y[i], z[i]); y[i] = __viaddmax_s16x2_relu(y[i], z[i], x[i]); z[i] = __viaddmax_s16x2_relu(z[i], x[i], y[i]); }
        if (OP == 5)
        { // the hybrid cell: x = hd chain (h), y = e, z = f, w = smax
          u32 a = hadd2_u(x[i], c0);
          u32 h = __vimax3_s16x2_relu(a, y[i], z[i]);
          w[i] = __vmaxs2(w[i], h);
          u32 hq = hadd2_u(h, c1);
          y[i] = __viaddmax_s16x2_relu(y[i], c2, hq);
          z[i] = __viaddmax_s16x2_relu(z[i], c2, hq);
          x[i] = h;
        }
        if (OP == 6)
        { // same without the running maximum
          u32 a = hadd2_u(x[i], c0);
          u32 h = __vimax3_s16x2_relu(a, y[i], z[i]);
          u32 hq = hadd2_u(h, c1);
          y[i] = __viaddmax_s16x2_relu(y[i], c2, hq);
          z[i] = __viaddmax_s16x2_relu(z[i], c2, hq);
          x[i] = h;
        }
        if (OP == 7)
        { // DPX part only (adds replaced by nothing): 1 vimax3 + 2 viaddmax
          u32 h = __vimax3_s16x2_relu(x[i], y[i], z[i]);
          y[i] = __viaddmax_s16x2_relu(y[i], c2, h);
          z[i] = __viaddmax_s16x2_relu(z[i], c2, h);
          x[i] = h ^ c0;
        }
        if (OP == 8)
        { // int16 cell
          u32 t = __viaddmax_s16x2(x[i], c0, y[i]);
          u32 h = __vimax_s16x2_relu(t, z[i]);
          w[i] = __vmaxs2(w[i], h);
          u32 hq = __vadd2(h, c1);
          y[i] = __viaddmax_s16x2(y[i], c2, hq);
          z[i] = __viaddmax_s16x2(z[i], c2, hq);
          x[i] = h;
        }
      }
  }
  long long t1 = clock64();
  u32 acc = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) acc ^= x[i] ^ y[i] ^ z[i] ^ w[i];
  if (acc == 0x12345677u) out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static const int ops_per[9] = {1, 1, 3, 1, 3, 6, 5, 4, 6};
static const char *names[9] = {"vimax3_relu 1 varying src", "vimax3_relu 2 varying", "vimax3_relu 3 varying", "viaddmax const add", "viaddmax 3 varying",
  "hybrid cell (6 instr)", "hybrid cell no smax (5 instr)", "dpx part of the cell (3 dpx + 1 lop)", "int16 cell (6 instr)"};

template <int OP, int CH> void run(int threads, u32 *dout, u32 *din, long long *dcyc, int nsm)
{
  const int iters = 2000;
  k<OP, CH><<<nsm, threads>>>(dout, din, 10, dcyc);
  CK(cudaDeviceSynchronize());
  k<OP, CH><<<nsm, threads>>>(dout, din, iters, dcyc);
  CK(cudaDeviceSynchronize());
  long long h[256]; CK(cudaMemcpy(h, dcyc, 8 * nsm, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < nsm; i++) avg += h[i]; avg /= nsm;
  double winstr = (double)iters * 4 * CH * ops_per[OP] * (threads / 32);
  printf("%-40s thr=%4d ch=%d  ipc/SMSP=%.3f  cycles per cell-step per SMSP=%.2f\n", names[OP], threads, CH, winstr / avg / 4,
         avg * 4 / ((double)iters * 4 * CH * (threads / 32)));
}

int main()
{
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  u32 hin[80]; for (int i = 0; i < 80; i++) hin[i] = 0x00030005u + i * 0x00010001u;
  hin[0] = 0x00040003u; hin[1] = 0x800c800cu; hin[2] = 0xfffffffeu - 0x10000u;
  u32 *din, *dout; long long *dcyc;
  CK(cudaMalloc(&din, sizeof hin)); CK(cudaMalloc(&dout, 4096 * 4)); CK(cudaMalloc(&dcyc, 8 * 256));
  CK(cudaMemcpy(din, hin, sizeof hin, cudaMemcpyHostToDevice));
#define RUN(OP) run<OP, 4>(512, dout, din, dcyc, nsm); run<OP, 8>(512, dout, din, dcyc, nsm); run<OP, 4>(256, dout, din, dcyc, nsm);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
  return 0;
}
