// Micro-benchmark: fp64 throughput of DFMA vs DMMA (mma.sync f64) on sm_100a.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a[16], b = 1.0000001, c = 0.9999999;
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma884_kernel(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma16816_kernel(double* out, int iters) {
  double c[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3 + i;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 4; ++i) b[i] = 1e-9 * threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// both pipes at once: do DFMA and DMMA overlap or share one datapath?
__global__ void mixed_kernel(double* out, int iters) {
  double c[4][2], a[8], x = 1.0 + threadIdx.x * 1e-9, y = 1e-9 * threadIdx.x;
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(x), "d"(y));
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int thr = warps * 32 > 1024 ? 1024 : warps * 32, blocks = sms * (warps * 32 / thr);
    float ms = time_it([&] { dfma_kernel<<<blocks, thr>>>(out, iters); });
    printf("DFMA      warps/SM=%2d  %.2f TFLOP/s\n", warps, 2.0 * 16 * iters * (double)blocks * thr / ms * 1e-9);
    ms = time_it([&] { dmma884_kernel<<<blocks, thr>>>(out, iters); });
    printf("DMMA884   warps/SM=%2d  %.2f TFLOP/s\n", warps, 2.0 * 256 * 8 * iters * (double)blocks * (thr / 32) / ms * 1e-9);
    ms = time_it([&] { dmma16816_kernel<<<blocks, thr>>>(out, iters); });
    printf("DMMA16816 warps/SM=%2d  %.2f TFLOP/s\n", warps, 2.0 * 2048 * 4 * iters * (double)blocks * (thr / 32) / ms * 1e-9);
    ms = time_it([&] { mixed_kernel<<<blocks, thr>>>(out, iters); });
    printf("MIXED     warps/SM=%2d  %.2f TFLOP/s (dmma %.2f + dfma %.2f)\n", warps,
           (2.0 * 256 * 4 + 2.0 * 32 * 8) * iters * (double)blocks * (thr / 32) / ms * 1e-9,
           2.0 * 256 * 4 * iters * (double)blocks * (thr / 32) / ms * 1e-9, 2.0 * 32 * 8 * iters * (double)blocks * (thr / 32) / ms * 1e-9);
  }
  return 0;
}
