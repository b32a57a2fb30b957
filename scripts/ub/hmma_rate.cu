// Micro-benchmark: issue rate of mma.sync.m16n8k16 bf16 (HMMA.16816.F32.BF16) per SM on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(float* out, int iters, long long* cyc) {
  float d[ILP][4];
  for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) d[i][e] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) s += d[i][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
__global__ void k8(float* out, int iters, long long* cyc) {
  float d[ILP][4];
  for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) d[i][e] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, b0 = a0 * 11;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a0), "r"(a1), "r"(b0));
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) s += d[i][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run8(int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k8<ILP><<<148, warps * 32>>>(out, iters, cyc);
  k8<ILP><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = double(h) / (double(iters) * ILP);
  printf("k8: warps/SM %2d ILP %d: %.2f clk per HMMA.1688 per warp, %.2f per SMSP\n", warps, ILP, per,
         per / (warps / 4.0 > 1 ? warps / 4.0 : 1));
  cudaFree(out); cudaFree(cyc);
}
template <int ILP> void run(int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<ILP><<<148, warps * 32>>>(out, iters, cyc);
  k<ILP><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = double(h) / (double(iters) * ILP);            // cycles per HMMA per warp
  printf("warps/SM %2d ILP %d: %.2f clk per HMMA per warp, %.2f clk per HMMA per SMSP\n", warps, ILP, per,
         per / (warps / 4.0 > 1 ? warps / 4.0 : 1));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1>(4); run<2>(4); run<4>(4); run<8>(4);
  run<1>(8); run<4>(8); run<4>(16); run<8>(16);
  run8<1>(4); run8<4>(4); run8<8>(4); run8<8>(16);
  return 0;
}
