// Store-pattern microbenchmark for the GEMM epilogue (diagnostics; not part of the library).
// A persistent grid of 148 CTAs x `warps` warps writes an [M, N] bf16 matrix tile by tile (128 x BN tiles), the way the
// epilogue warps do, with three lane->address mappings:
//   0  row-per-lane  : lane = row, 64 contiguous bytes per lane per chunk as two 256-bit stores (the register epilogue)
//   1  coalesced     : lanes cover 1024 contiguous bytes of ONE row per instruction (what a TMA store / transposed epilogue emits)
//   2  row-per-lane, 128-bit stores x4
//   3  half-warp-per-row: 16 lanes x 32 B cover 512 contiguous bytes of a row, 2 rows per instruction
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/ubench_store tools/ubench_store.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void stg256(void* ptr, uint32_t v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ void stg128(void* ptr, uint32_t v) {
  asm volatile("st.global.v4.b32 [%0], {%1, %1, %1, %1};" ::"l"(ptr), "r"(v) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) store_kernel(uint16_t* out, int M, int N, int BN, int warps, int delay) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps) return;
  const int tiles_m = M / 128, tiles_n = N / BN;
  const int q = warp & 3, wg = warp >> 2, wgs = warps >> 2;
  for (int t = blockIdx.x; t < tiles_m * tiles_n; t += gridDim.x) {
    const int tm = t / tiles_n, tn = t % tiles_n;
    for (int c0 = wg * 32; c0 < BN; c0 += 32 * wgs) {
      if (delay) __nanosleep(delay);
      const uint32_t v = t + c0;
      if (MODE == 0) {
        uint16_t* o = out + (size_t)(tm * 128 + q * 32 + lane) * N + tn * BN + c0;
        stg256(o, v);
        stg256(o + 16, v);
      } else if (MODE == 2) {
        uint16_t* o = out + (size_t)(tm * 128 + q * 32 + lane) * N + tn * BN + c0;
        stg128(o, v); stg128(o + 8, v); stg128(o + 16, v); stg128(o + 24, v);
      } else if (MODE == 1) {
        // same 32 rows x 32 cols region (2 KB), but each instruction covers 16 rows x 64 B... a 32-column chunk only has
        // 64 B per row, so "coalesced" here means 2 lanes per row: 16 rows per instruction
        uint16_t* o = out + (size_t)(tm * 128 + q * 32 + (lane >> 1)) * N + tn * BN + c0 + (lane & 1) * 16;
        stg256(o, v);
        stg256(o + (size_t)16 * N, v);
      } else {
        // whole-tile view: the warp owns 32 rows x BN columns over the chunk loop; here each instruction writes 2 rows x 512 B
        // (only valid when BN % 256 == 0 and wgs == 1): c0 indexes 32-column chunks -> remap to (row pair, 256-col half)
        const int idx = c0 / 32;                      // 0 .. BN/32-1
        const int per_row = BN / 256;                 // instructions per row pair... BN=256 -> 1
        for (int rp = idx; rp < 16 * per_row; rp += BN / 32) {
          const int r = (rp / per_row) * 2 + (lane >> 4), cc = (rp % per_row) * 256 + (lane & 15) * 16;
          uint16_t* o = out + (size_t)(tm * 128 + q * 32 + r) * N + tn * BN + cc;
          stg256(o, v);
        }
      }
    }
  }
}

template <int MODE>
float run(uint16_t* out, int M, int N, int BN, int warps, int delay, int grid) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) store_kernel<MODE><<<grid, 512>>>(out, M, N, BN, warps, delay);
  cudaEventRecord(a);
  const int reps = 10;
  for (int i = 0; i < reps; ++i) store_kernel<MODE><<<grid, 512>>>(out, M, N, BN, warps, delay);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  if (cudaGetLastError() != cudaSuccess) { printf("error\n"); exit(1); }
  return ms / reps * 1e3f;
}

int main() {
  const int M = 65536;
  uint16_t* out;
  cudaMalloc(&out, (size_t)M * 2560 * 2);
  const int shapes[][2] = {{2560, 256}, {320, 160}, {1280, 256}, {640, 128}};
  for (auto& s : shapes) {
    const int N = s[0], BN = s[1];
    const double mb = (double)M * N * 2 / 1e6;
    for (int warps : {8, 16}) {
      for (int grid : {148, 296}) {
        if (grid == 296 && warps == 16) continue;
        float t0 = run<0>(out, M, N, BN, warps, 0, grid), t1 = run<1>(out, M, N, BN, warps, 0, grid),
              t2 = run<2>(out, M, N, BN, warps, 0, grid);
        float t3 = (BN % 256 == 0 && warps == 4) ? run<3>(out, M, N, BN, warps, 0, grid) : 0.f;
        printf("N=%4d BN=%3d warps=%2d grid=%3d (%.0f MB): row/lane 2x256b %6.1f us %5.0f GB/s | 2 lanes/row %6.1f us %5.0f GB/s | "
               "row/lane 4x128b %6.1f us %5.0f GB/s | t3 %.1f\n",
               N, BN, warps, grid, mb, t0, mb / t0 * 1e3, t1, mb / t1 * 1e3, t2, mb / t2 * 1e3, t3);
      }
    }
  }
  return 0;
}
