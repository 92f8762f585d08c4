// Issue-rate probe for sm_100a: do the FP64 pipe and the ALU pipe (FSEL, LOP3, ISETP ...) run
// concurrently, and which selects can be moved to the FMA pipe (predicated IMAD)?
// Every test runs `W` warps per SM sub-partition for a fixed number of loop iterations and
// reports warp-instructions per cycle per sub-partition (clock64 on one SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

// ---- instruction streams: K independent chains per thread, asm volatile so nothing is folded
#define D_OP(x, y) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(y))
#define DS_OP(p, x, y) asm volatile("{.reg .pred q; setp.gt.f64 q, %1, %2; selp.u32 %0, 1, %0, q;}" : "+r"(p) : "d"(x), "d"(y))
#define A_OP(x, y) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(x) : "r"(y))
#define SEL_OP(x, y, c) asm volatile("{.reg .pred q; setp.ne.u32 q, %2, 0; selp.b32 %0, %1, %0, q;}" : "+r"(x) : "r"(y), "r"(c))
#define I_OP(x, y, z) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
// x = (y > x) ? y : x on the 64-bit pair (lo,hi): compare on the FP64 pipe, then either two FSEL (ALU pipe)
// or two predicated IMAD with an opaque multiplier 1 (FMA pipe)
#define MAXSEL_OP(x, y) asm volatile("{.reg .pred q; setp.gt.f64 q, %1, %0; selp.f64 %0, %1, %0, q;}" : "+d"(x) : "d"(y))
#define MAXMAD_OP(x, y, one) asm volatile("{.reg .pred q; .reg .b32 xl, xh, yl, yh; setp.gt.f64 q, %1, %0; mov.b64 {xl,xh}, %0; mov.b64 {yl,yh}, %1; @q mad.lo.u32 xl, yl, %2, 0; @q mad.lo.u32 xh, yh, %2, 0; mov.b64 %0, {xl,xh};}" : "+d"(x) : "d"(y), "r"(one))
#define F_OP(x, y, z) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(y), "f"(z))

template <int MODE>
__global__ void k_probe(long long* out, int n, double dseed, unsigned useed) {
  dseed += threadIdx.x * 0.5; useed += threadIdx.x * 7u;
  double d0 = dseed, d1 = dseed + 1, d2 = dseed + 2, d3 = dseed + 3, d4 = dseed + 4, d5 = dseed + 5,
         d6 = dseed + 6, d7 = dseed + 7;
  unsigned a0 = useed, a1 = useed + 1, a2 = useed + 2, a3 = useed + 3, a4 = useed + 4, a5 = useed + 5,
           a6 = useed + 6, a7 = useed + 7;
  unsigned i0 = useed * 3, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6,
           i7 = i0 + 7;
  float f0 = (float)dseed, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6,
        f7 = f0 + 7;
  const double dy = dseed * 1e-9;
  const unsigned uy = useed | 1u, uz = (useed >> 3) + threadIdx.x;
  const unsigned one = (useed & 0u) + (n > 0 ? 1u : 0u) + (threadIdx.x >> 12);
  const float fy = 1.0f + (float)dseed * 1e-7f, fz = 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (MODE == 0) {  // FP64 only: 8 DADD
        D_OP(d0, dy); D_OP(d1, dy); D_OP(d2, dy); D_OP(d3, dy); D_OP(d4, dy); D_OP(d5, dy); D_OP(d6, dy); D_OP(d7, dy);
      } else if (MODE == 1) {  // ALU only: 8 LOP3
        A_OP(a0, uy); A_OP(a1, uy); A_OP(a2, uy); A_OP(a3, uy); A_OP(a4, uy); A_OP(a5, uy); A_OP(a6, uy); A_OP(a7, uy);
      } else if (MODE == 2) {  // 4 DADD + 4 LOP3 interleaved
        D_OP(d0, dy); A_OP(a0, uy); D_OP(d1, dy); A_OP(a1, uy); D_OP(d2, dy); A_OP(a2, uy); D_OP(d3, dy); A_OP(a3, uy);
      } else if (MODE == 3) {  // 4 DADD + 4 IMAD interleaved
        D_OP(d0, dy); I_OP(i0, uy, uz); D_OP(d1, dy); I_OP(i1, uy, uz); D_OP(d2, dy); I_OP(i2, uy, uz); D_OP(d3, dy); I_OP(i3, uy, uz);
      } else if (MODE == 4) {  // 4 LOP3 + 4 IMAD interleaved
        A_OP(a0, uy); I_OP(i0, uy, uz); A_OP(a1, uy); I_OP(i1, uy, uz); A_OP(a2, uy); I_OP(i2, uy, uz); A_OP(a3, uy); I_OP(i3, uy, uz);
      } else if (MODE == 5) {  // IMAD only
        I_OP(i0, uy, uz); I_OP(i1, uy, uz); I_OP(i2, uy, uz); I_OP(i3, uy, uz); I_OP(i4, uy, uz); I_OP(i5, uy, uz); I_OP(i6, uy, uz); I_OP(i7, uy, uz);
      } else if (MODE == 6) {  // FFMA only
        F_OP(f0, fy, fz); F_OP(f1, fy, fz); F_OP(f2, fy, fz); F_OP(f3, fy, fz); F_OP(f4, fy, fz); F_OP(f5, fy, fz); F_OP(f6, fy, fz); F_OP(f7, fy, fz);
      } else if (MODE == 7) {  // 4 DADD + 4 FFMA
        D_OP(d0, dy); F_OP(f0, fy, fz); D_OP(d1, dy); F_OP(f1, fy, fz); D_OP(d2, dy); F_OP(f2, fy, fz); D_OP(d3, dy); F_OP(f3, fy, fz);
      } else if (MODE == 8) {  // 2 DADD + 2 LOP3 + 4 IMAD: does a third pipe fill the gaps?
        D_OP(d0, dy); I_OP(i0, uy, uz); A_OP(a0, uy); I_OP(i1, uy, uz); D_OP(d1, dy); I_OP(i2, uy, uz); A_OP(a1, uy); I_OP(i3, uy, uz);
      } else if (MODE == 9) {  // select pairs as the kernel has them: DSETP + 2 SEL, 4 chains
        SEL_OP(a0, uy, a4); SEL_OP(a1, uy, a5); SEL_OP(a2, uy, a6); SEL_OP(a3, uy, a7);
        SEL_OP(a4, uz, a0); SEL_OP(a5, uz, a1); SEL_OP(a6, uz, a2); SEL_OP(a7, uz, a3);
      } else if (MODE == 10) {  // 8 DADD : 4 LOP3 ... FP64-heavy mix 2:1
        D_OP(d0, dy); D_OP(d1, dy); A_OP(a0, uy); D_OP(d2, dy); D_OP(d3, dy); A_OP(a1, uy); D_OP(d4, dy); D_OP(d5, dy);
      } else if (MODE == 11) {  // DSETP->predicate->SEL (FP64 compare feeding ALU)
        DS_OP(a0, d0, dy); DS_OP(a1, d1, dy); DS_OP(a2, d2, dy); DS_OP(a3, d3, dy);
        DS_OP(a4, d4, dy); DS_OP(a5, d5, dy); DS_OP(a6, d6, dy); DS_OP(a7, d7, dy);
      } else if (MODE == 12) {  // running max, select by 2 FSEL
        MAXSEL_OP(d0, d4); MAXSEL_OP(d1, d5); MAXSEL_OP(d2, d6); MAXSEL_OP(d3, d7);
        MAXSEL_OP(d4, d0); MAXSEL_OP(d5, d1); MAXSEL_OP(d6, d2); MAXSEL_OP(d7, d3);
      } else if (MODE == 13) {  // running max, select by 2 predicated IMAD
        MAXMAD_OP(d0, d4, one); MAXMAD_OP(d1, d5, one); MAXMAD_OP(d2, d6, one); MAXMAD_OP(d3, d7, one);
        MAXMAD_OP(d4, d0, one); MAXMAD_OP(d5, d1, one); MAXMAD_OP(d6, d2, one); MAXMAD_OP(d7, d3, one);
      } else if (MODE == 14) {  // max by FSEL + independent DADD work: 8x(DSETP+2FSEL) interleaved with 8 DADD
        MAXSEL_OP(d0, d4); D_OP(d4, dy); MAXSEL_OP(d1, d5); D_OP(d5, dy); MAXSEL_OP(d2, d6); D_OP(d6, dy); MAXSEL_OP(d3, d7); D_OP(d7, dy);
      } else if (MODE == 15) {  // same with predicated IMAD
        MAXMAD_OP(d0, d4, one); D_OP(d4, dy); MAXMAD_OP(d1, d5, one); D_OP(d5, dy); MAXMAD_OP(d2, d6, one); D_OP(d6, dy); MAXMAD_OP(d3, d7, one); D_OP(d7, dy);
      }
    }
  }
  const long long t1 = clock64();
  // keep everything alive
  double ds = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
  unsigned as = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ i0 ^ i1 ^ i2 ^ i3 ^ i4 ^ i5 ^ i6 ^ i7;
  float fs = f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
  if (ds == 1.2345 && as == 77u && fs == 3.3f) out[1] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int per_iter, long long* dout) {
  for (int wps = 1; wps <= 8; wps *= 2) {   // warps per sub-partition
    const int threads = 128 * wps;           // 4 sub-partitions
    if (threads > 1024) break;
    k_probe<MODE><<<148, threads>>>(dout, ITER, 1.0, 12345u);
    cudaDeviceSynchronize();
    k_probe<MODE><<<148, threads>>>(dout, ITER, 1.0, 12345u);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, dout, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double inst = (double)ITER * 4 * per_iter * wps;   // warp-instructions per sub-partition
    printf("%-34s warps/SMSP %d: %9lld cycles, %.3f warp-inst/clk/SMSP%s\n", name, wps, cyc, inst / (double)cyc,
           e == cudaSuccess ? "" : " (ERROR)");
  }
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 64);
  cudaMemset(dout, 0, 64);
  run<0>("DADD x8", 8, dout);
  run<1>("LOP3 x8", 8, dout);
  run<2>("DADD x4 + LOP3 x4", 8, dout);
  run<3>("DADD x4 + IMAD x4", 8, dout);
  run<4>("LOP3 x4 + IMAD x4", 8, dout);
  run<5>("IMAD x8", 8, dout);
  run<6>("FFMA x8", 8, dout);
  run<7>("DADD x4 + FFMA x4", 8, dout);
  run<8>("DADD x2 + LOP3 x2 + IMAD x4", 8, dout);
  run<9>("ISETP+SEL x8 (16 instr)", 16, dout);
  run<10>("DADD x6 + LOP3 x2", 8, dout);
  run<11>("DSETP+SEL x8 (16 instr)", 16, dout);
  run<12>("max: DSETP+2FSEL x8 (24 instr)", 24, dout);
  run<13>("max: DSETP+2 @p IMAD x8 (24)", 24, dout);
  run<14>("(DSETP+2FSEL+DADD) x4 (16)", 16, dout);
  run<15>("(DSETP+2@pIMAD+DADD) x4 (16)", 16, dout);
  return 0;
}
