// peaks.cu -- measured pipe peaks of the GPU the runtime is running on ("butterflies/s
// resp. MACs/s against the measured IMAD peak, to be measured once with a synthetic kernel").
//
// The NTT and base-conversion kernels are bound by the integer multiply pipe, not by HBM, so their
// roofline denominator is an instruction rate.  Each kernel below issues long unrolled sequences
// of ONE instruction kind over 8 independent accumulator chains per thread (enough ILP and warps
// to saturate the pipe) and reports thread-instructions per second:
//   [0] IMAD.WIDE.U32  (32x32+64 -> 64: the building block of every 64-bit modular product)
//   [1] IMAD           (32x32+32 -> 32 low word)
//   [2] IADD3          (ALU pipe)
//   [3] DFMA           (FP64 pipe, encode/decode FFT)
//   [4] a 1:1 interleave of IMAD.WIDE.U32 and IADD3 (do the two pipes issue side by side?)
//   [5] IMAD.HI.U32
//   [6] a 1:1 interleave of IMAD.WIDE.U32 and DFMA (rate of each)
// bench.py calls ace_measure_pipe_peaks() live and divides the NTT's butterflies x IMADs by [0].
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/ace_b200.h"

namespace {

constexpr int kChains = 8, kIters = 2048, kThreads = 256;

template <int KIND>
__global__ void __launch_bounds__(kThreads) pipe_peak_kernel(uint64_t* out, uint32_t a, uint32_t b) {
  uint64_t acc[kChains];
  uint32_t x[kChains];
  double   d[kChains];
#pragma unroll
  for (int c = 0; c < kChains; c++) {
    acc[c] = threadIdx.x * 7 + c;
    x[c]   = threadIdx.x * 3 + c + a;
    d[c]   = (double)(threadIdx.x + c);
  }
  const double da = (double)a * 1e-9, db = (double)b * 1e-9;
#pragma unroll 1
  for (int it = 0; it < kIters / 8; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int c = 0; c < kChains; c++) {
        if (KIND == 0 || KIND == 4)
          acc[c] = (uint64_t)(uint32_t)acc[c] * b + acc[c];  // IMAD.WIDE.U32 R, R.lo, b, R
        if (KIND == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
        if (KIND == 2 || KIND == 4)
          asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(x[c]) : "r"(a), "r"(b));  // ptxas: one IADD3
        if (KIND == 3 || KIND == 6) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(da), "d"(db));
        if (KIND == 5) x[c] = __umulhi(x[c], b) + a;  // IMAD.HI.U32
        if (KIND == 6) acc[c] = (uint64_t)(uint32_t)acc[c] * b + acc[c];
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int c = 0; c < kChains; c++) {
    if (KIND == 0 || KIND == 4 || KIND == 6) s += acc[c];
    if (KIND == 1 || KIND == 2 || KIND == 4 || KIND == 5) s ^= x[c];
    if (KIND == 3 || KIND == 6) s += (uint64_t)d[c];
  }
  if (s == 0x1234567) out[0] = s;  // never true in practice; keeps the chains alive
}

template <int KIND>
double measure(cudaStream_t st, uint64_t* scratch, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, st);
    pipe_peak_kernel<KIND><<<blocks, kThreads, 0, st>>>(scratch, 3 + rep, 5);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // thread-instructions of the measured kind (KIND 4 counts the IMAD.WIDE half only)
  const double ops = (double)blocks * kThreads * kIters * kChains;
  return ops / (best * 1e-3) * 1e-9;  // G thread-instr / s
}

}  // namespace

extern "C" int ace_measure_pipe_peaks(int device, double* gops, int n) {
  if (cudaSetDevice(device) != cudaSuccess) return -2;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
  uint64_t* scratch = nullptr;
  if (cudaMalloc(&scratch, 64) != cudaSuccess) return -1;
  cudaStream_t st;
  cudaStreamCreate(&st);
  const int blocks = prop.multiProcessorCount * 8;  // 8 resident CTAs of 256 threads per SM
  double    r[7];
  r[0] = measure<0>(st, scratch, blocks);
  r[1] = measure<1>(st, scratch, blocks);
  r[2] = measure<2>(st, scratch, blocks);
  r[3] = measure<3>(st, scratch, blocks);
  r[4] = measure<4>(st, scratch, blocks);
  r[5] = measure<5>(st, scratch, blocks);
  r[6] = measure<6>(st, scratch, blocks);
  for (int i = 0; i < n && i < 7; i++) gops[i] = r[i];
  cudaStreamDestroy(st);
  cudaFree(scratch);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
