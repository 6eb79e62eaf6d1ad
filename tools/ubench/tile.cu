// Tile-level probe: the scan kernel's R x 4 DP tile in isolation, in several formulations, to find
// what bounds it.  Reports clk per warp cell-pair per SMSP (lower is better) at 4 warps / SMSP.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../swipe_b200/csrc/sw_kernels.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

template <int V>
__device__ __forceinline__ void cell(u32 hd, u32 s, u32 &e, u32 &f, u32 &h, u32 &smax, const u32 negq, const u32 negr)
{
  if (V == 0 || V == 3 || V == 6) swb_cell<SWB_MODE_HYBRID>(hd, s, e, f, h, smax, negq, negr);
  if (V == 1)
  { // no running maximum
    u32 a = swb_hadd2(hd, s);
    h = __vimax3_s16x2_relu(a, e, f);
    u32 hq = swb_hadd2(h, negq);
    e = __viaddmax_s16x2_relu(e, negr, hq);
    f = __viaddmax_s16x2_relu(f, negr, hq);
  }
  if (V == 2)
  { // immediates for the penalties (q = 12, r = 1 as fp16 pattern / two's complement)
    u32 a = swb_hadd2(hd, s);
    h = __vimax3_s16x2_relu(a, e, f);
    smax = __vmaxs2(smax, h);
    u32 hq = swb_hadd2(h, 0x800c800cu);
    e = __viaddmax_s16x2_relu(e, 0xffffffffu, hq);
    f = __viaddmax_s16x2_relu(f, 0xffffffffu, hq);
  }
  if (V == 7)
  { // F through the FMA pipe + a 2-input fp16 max; H - q with relu so that E, F stay >= 0
    u32 a = swb_hadd2(hd, s);
    h = __vimax3_s16x2_relu(a, e, f);
    smax = __vmaxs2(smax, h);
    u32 hq = swb_hadd2_relu(h, negq);
    e = __viaddmax_s16x2_relu(e, negr, hq);
    f = swb_hmax2(swb_hadd2(f, 0x80018001u), hq);
  }
  if (V == 8)
  { // E and F both through FMA pipe + fp16 max
    u32 a = swb_hadd2(hd, s);
    h = __vimax3_s16x2_relu(a, e, f);
    smax = __vmaxs2(smax, h);
    u32 hq = swb_hadd2_relu(h, negq);
    e = swb_hmax2(swb_hadd2(e, 0x80018001u), hq);
    f = swb_hmax2(swb_hadd2(f, 0x80018001u), hq);
  }
  if (V == 9)
  { // as 7, running maximum with the 2-input fp16 max
    u32 a = swb_hadd2(hd, s);
    h = __vimax3_s16x2_relu(a, e, f);
    smax = swb_hmax2(smax, h);
    u32 hq = swb_hadd2_relu(h, negq);
    e = __viaddmax_s16x2_relu(e, negr, hq);
    f = swb_hmax2(swb_hadd2(f, 0x80018001u), hq);
  }
  if (V == 10)
  { // everything 2-input: h by two fp16 max (E, F >= 0 make the relu implicit)
    u32 a = swb_hadd2(hd, s);
    h = swb_hmax2(swb_hmax2(a, e), f);
    smax = swb_hmax2(smax, h);
    u32 hq = swb_hadd2_relu(h, negq);
    e = swb_hmax2(swb_hadd2(e, 0x80018001u), hq);
    f = swb_hmax2(swb_hadd2(f, 0x80018001u), hq);
  }
  if (V == 4) swb_cell<SWB_MODE_INT16>(hd, s, e, f, h, smax, negq, negr);
  if (V == 5)
  { // E' = E + q formulation: no H - q add; a on the FMA pipe
    u32 a = swb_hadd2(hd, s);
    u32 m = __vmaxs2(e, f);
    h = __viaddmax_s16x2_relu(m, negq, a);
    smax = __vmaxs2(smax, h);
    e = __viaddmax_s16x2(e, negr, h);
    f = __viaddmax_s16x2(f, negr, h);
  }
}

template <int V, int R>
__global__ void __launch_bounds__(128, 4) tile(u32 *out, const u32 *in, int steps, long long *cyc)
{
  extern __shared__ uint4 sm4[];
  const u32 sbase = (u32)__cvta_generic_to_shared(sm4);
  for (int i = threadIdx.x; i < 53248 / 4; i += blockDim.x) ((u32 *)sm4)[i] = in[i & 63] & 0x00070007u;
  __syncthreads();
  u32 H[R], E[R], rq[R];
#pragma unroll
  for (int i = 0; i < R; i++) { H[i] = 0; E[i] = 0; rq[i] = sbase + (((in[64 + i] + (threadIdx.x >> 3) * 7) % 22u) * 128u + (threadIdx.x & 7) * 16u); }
  u32 smax = 0, dtop = 0, ih0 = 0, ih1 = 0, ih2 = 0, ih3 = 0, if0 = 0, if1 = 0, if2 = 0, if3 = 0;
  const u32 negq = in[100], negr = in[101];
  u32 roff = (threadIdx.x >> 3) * 2816u % 45056u;   // per-stage ring offset, like the kernel
  long long t0 = clock64();
#pragma unroll 1
  for (int t = 0; t < steps; t++)
  {
    u32 hup0 = ih0, hup1 = ih1, hup2 = ih2, hup3 = ih3, f0 = if0, f1 = if1, f2 = if2, f3 = if3;
    u32 dg = dtop;
    dtop = ih3;
#pragma unroll
    for (int i = 0; i < R; i++)
    {
      uint4 sc;
      if (V == 3) sc = make_uint4(rq[i], rq[i] ^ roff, rq[i] + roff, roff);   // no shared-memory load
      else sc = swb_lds128(rq[i] + roff);
      u32 hd = dg, e = E[i], h;
      dg = H[i];
      cell<V>(hd, sc.x, e, f0, h, smax, negq, negr); hd = hup0; hup0 = h;
      cell<V>(hd, sc.y, e, f1, h, smax, negq, negr); hd = hup1; hup1 = h;
      cell<V>(hd, sc.z, e, f2, h, smax, negq, negr); hd = hup2; hup2 = h;
      cell<V>(hd, sc.w, e, f3, h, smax, negq, negr); hup3 = h;
      H[i] = h; E[i] = e;
    }
    ih0 = hup0; ih1 = hup1; ih2 = hup2; ih3 = hup3; if0 = f0; if1 = f1; if2 = f2; if3 = f3;
    roff = roff + 2816u == 47872u ? 0u : roff + 2816u;
    if (V == 6) __syncthreads();
  }
  long long t1 = clock64();
  u32 acc = smax ^ dtop;
#pragma unroll
  for (int i = 0; i < R; i++) acc ^= H[i] ^ E[i];
  if (acc == 0x12345677u) out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static const char *names[] = {"hybrid cell as in the kernel", "hybrid, no running max", "hybrid, immediate penalties",
                              "hybrid, scores from registers (no LDS)", "int16 cell", "E+q formulation", "hybrid + __syncthreads per step",
                              "F via hadd2+hmax2 (ALU 6 / FMA 6)", "E and F via hadd2+hmax2 (ALU 5 / FMA 8)", "as 7, smax by hmax2", "all 2-input fp16 max"};

template <int V, int R> void run(u32 *dout, u32 *din, long long *dcyc, int nsm, int oversub = 1)
{
  const int steps = 3000 / oversub, ctas = nsm * 4 * oversub;
  CK(cudaFuncSetAttribute((const void *)tile<V, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 53248));
  tile<V, R><<<ctas, 128, 53248>>>(dout, din, 10, dcyc);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  tile<V, R><<<ctas, 128, 53248>>>(dout, din, steps, dcyc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  static long long h[65536]; CK(cudaMemcpy(h, dcyc, 8 * ctas, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < ctas; i++) avg += h[i]; avg /= ctas;
  // per SMSP: 4 warps (one per resident CTA) each doing steps * R * 4 cell pairs
  long long mn = h[0], mx = h[0]; for (int i = 0; i < ctas; i++) { if (h[i] < mn) mn = h[i]; if (h[i] > mx) mx = h[i]; }
  const double cellpairs = (double)steps * R * 4 * 32 * 4 * ctas;   // cells = 2 x this
  printf("%-44s R=%2d x%d clock64: %.2f clk/cell-pair/SMSP (cta min %.2f max %.2f)   events: %.3f ms = %.0f GCUPS\n", names[V], R, oversub,
         avg / ((double)steps * R * 4 * 4), mn / ((double)steps * R * 4 * 4), mx / ((double)steps * R * 4 * 4), ms,
         2.0 * cellpairs / (ms * 1e-3) * 1e-9);
}

int main()
{
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  u32 hin[128]; for (int i = 0; i < 128; i++) hin[i] = 0x00030005u + i * 0x00110013u;
  hin[100] = 0x800c800cu; hin[101] = 0xffffffffu;
  u32 *din, *dout; long long *dcyc;
  CK(cudaMalloc(&din, sizeof hin)); CK(cudaMalloc(&dout, 4096 * 4)); CK(cudaMalloc(&dcyc, 8 * 65536));
  CK(cudaMemcpy(din, hin, sizeof hin, cudaMemcpyHostToDevice));
  run<0, 24>(dout, din, dcyc, nsm); run<1, 24>(dout, din, dcyc, nsm); run<2, 24>(dout, din, dcyc, nsm);
  run<3, 24>(dout, din, dcyc, nsm); run<4, 24>(dout, din, dcyc, nsm); run<5, 24>(dout, din, dcyc, nsm);
  run<7, 24>(dout, din, dcyc, nsm); run<8, 24>(dout, din, dcyc, nsm); run<9, 24>(dout, din, dcyc, nsm); run<10, 24>(dout, din, dcyc, nsm);
  run<6, 24>(dout, din, dcyc, nsm); run<0, 12>(dout, din, dcyc, nsm); run<0, 8>(dout, din, dcyc, nsm); run<0, 16>(dout, din, dcyc, nsm); run<0, 20>(dout, din, dcyc, nsm); run<3, 12>(dout, din, dcyc, nsm);
  // steady state: more CTAs than resident slots, the SM refills as CTAs finish
  run<0, 24>(dout, din, dcyc, nsm, 4); run<0, 24>(dout, din, dcyc, nsm, 16); run<1, 24>(dout, din, dcyc, nsm, 16); run<6, 24>(dout, din, dcyc, nsm, 16);
  return 0;
}
