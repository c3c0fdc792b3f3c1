// Micro-benchmark: legacy mma.sync throughput on sm_100a (tf32 m16n8k8, bf16 m16n8k16), plus an
// fp32 FFMA ceiling, to size the NMF contraction design.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_tf32(float* out, int iters) {
  float c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  unsigned a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, b0 = 4, b1 = 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_bf16(float* out, int iters) {
  float c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  unsigned a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, b0 = 4, b1 = 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_ffma(float* out, int iters, float x) {
  float c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fmaf(c[i], x, 1.0f);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed fp32x2 FMA (FFMA2, sm_100+): does one issue slot buy two FMAs per lane?
template <int ILP>
__global__ void k_ffma2(float* out, int iters, float x) {
  float2 c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  const float2 xx = make_float2(x, x), one = make_float2(1.0f, 1.0f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = __ffma2_rn(c[i], xx, one);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i].x + c[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 interleaved 1:1 with integer ALU work: the issue-bound regime of the covariance kernels
template <int ILP, bool PACKED>
__global__ void k_mix(float* out, int iters, float x) {
  float2 c[ILP];
  unsigned u[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    c[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    u[i] = threadIdx.x * 2654435761u + i;
  }
  const float2 xx = make_float2(x, x), one = make_float2(1.0f, 1.0f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (PACKED) {
        c[i] = __ffma2_rn(c[i], xx, one);
      } else {
        c[i].x = fmaf(c[i].x, x, 1.0f);
        c[i].y = fmaf(c[i].y, x, 1.0f);
      }
      u[i] = (u[i] ^ (u[i] >> 7)) + 0x9e3779b9u;
      u[i] = (u[i] ^ (u[i] << 3)) + it;
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i].x + c[i].y + (float)u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    const int blocks = sms * (warps >= 16 ? 2 : 1), threads = (warps >= 16 ? warps / 2 : warps) * 32;
    const double nw = (double)blocks * threads / 32;
    float ms = time_ms([&] { k_tf32<8><<<blocks, threads>>>(out, iters); });
    printf("tf32 m16n8k8  warps/SM=%2d : %8.1f TFLOP/s\n", warps, nw * iters * 8 * (2.0 * 16 * 8 * 8) / ms / 1e9);
    ms = time_ms([&] { k_bf16<8><<<blocks, threads>>>(out, iters); });
    printf("bf16 m16n8k16 warps/SM=%2d : %8.1f TFLOP/s\n", warps, nw * iters * 8 * (2.0 * 16 * 8 * 16) / ms / 1e9);
    ms = time_ms([&] { k_ffma<8><<<blocks, threads>>>(out, iters * 8, 1.0001f); });
    printf("fp32 FFMA     warps/SM=%2d : %8.1f TFLOP/s\n", warps, nw * 32 * iters * 8.0 * 8 * 2 / ms / 1e9);
    ms = time_ms([&] { k_ffma2<8><<<blocks, threads>>>(out, iters * 8, 1.0001f); });
    printf("fp32 FFMA2    warps/SM=%2d : %8.1f TFLOP/s\n", warps, nw * 32 * iters * 8.0 * 8 * 4 / ms / 1e9);
    ms = time_ms([&] { k_mix<8, false><<<blocks, threads>>>(out, iters * 4, 1.0001f); });
    const float ms2 = time_ms([&] { k_mix<8, true><<<blocks, threads>>>(out, iters * 4, 1.0001f); });
    printf("2 FFMA + 4 int ALU vs 1 FFMA2 + 4 int ALU, warps/SM=%2d : %.3f ms vs %.3f ms\n", warps, ms, ms2);
  }
  printf("SMs=%d clock=%d MHz\n", sms, p.clockRate / 1000);
  return 0;
}
