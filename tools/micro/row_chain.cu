// Microbenchmark of the PGS row-update chain for ONE warp (the tail of a launch is one such warp alone on its SM):
//   u += -c * d;  nl = clamp(u * invd + base);  dl = nl - lam;  d = broadcast(dl from lane i)
// Variants: A = shuffle in the kernel body, B = shuffle inside a __noinline__ function (ptxas cannot prove
// convergence there and guards every shuffle with BRA.DIV), C = broadcast through shared memory (STS + LDS),
// D = like B but 4 LDS + 4 FFMA + selects around the chain (the compiled loop of the solver).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o row_chain row_chain.cu ; run: ./row_chain
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
#define N_ROWS 4096

__device__ __forceinline__ float row_update(float& u, float invd, float base, float lo, float hi, float lam, float c, float d) {
  u = fmaf(-c, d, u);
  float nl = fmaf(u, invd, base);
  nl = fminf(fmaxf(nl, lo), hi);
  return nl - lam;
}

__global__ void kA(float* out, long long* cyc, float invd, float base, float lo, float hi, float c) {
  float u = threadIdx.x * 0.001f, lam = 0.f, d = 0.f;
  const long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_ROWS; i++) {
    const float dl = row_update(u, invd, base, lo, hi, lam, c, d);
    d = __shfl_sync(FULL, dl, i & 15, 16);
    if ((threadIdx.x & 15) == (i & 15)) lam += dl;
  }
  const long long t1 = clock64();
  out[threadIdx.x] = u + lam;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__device__ __noinline__ float loopB(float u, float invd, float base, float lo, float hi, float c, int n) {
  float lam = 0.f, d = 0.f;
#pragma unroll 4
  for (int i = 0; i < n; i++) {
    const float dl = row_update(u, invd, base, lo, hi, lam, c, d);
    d = __shfl_sync(FULL, dl, i & 15, 16);
    if ((threadIdx.x & 15) == (i & 15)) lam += dl;
  }
  return u + lam;
}
__global__ void kB(float* out, long long* cyc, float invd, float base, float lo, float hi, float c, int n) {
  const long long t0 = clock64();
  const float r = loopB(threadIdx.x * 0.001f, invd, base, lo, hi, c, n);
  const long long t1 = clock64();
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
}

__global__ void kC(float* out, long long* cyc, float invd, float base, float lo, float hi, float c) {
  __shared__ float bc[2][32];
  float u = threadIdx.x * 0.001f, lam = 0.f, d = 0.f;
  const int half = threadIdx.x >> 4;
  const long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_ROWS; i++) {
    const float dl = row_update(u, invd, base, lo, hi, lam, c, d);
    if ((threadIdx.x & 15) == (i & 15)) { bc[i & 1][half] = dl; lam += dl; }
    __syncwarp();
    d = bc[i & 1][half];
  }
  const long long t1 = clock64();
  out[threadIdx.x] = u + lam;
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
}

__device__ __noinline__ float loopD(const float* __restrict__ tab, float u0, float invd, float base, float lo, float hi, int n, unsigned mk) {
  float u[3] = {u0, u0 * 0.5f, u0 * 0.25f}, mu = 0.f, lam = 0.f;
  const int lane = threadIdx.x & 15;
#pragma unroll 4
  for (int i = 0; i < n; i++) {
    const float* row = tab + (i & 31) * 64;
    const float c0 = row[lane], c1 = row[16 + lane], c2 = row[32 + lane], cw = row[48 + lane];
    const bool active = (mk >> (i & 31)) & 1u;
    float nl = fmaf(u[0], invd, base);
    nl = fminf(fmaxf(nl, lo), hi);
    const float dl = active ? nl - lam : 0.f;
    const float d = __shfl_sync(FULL, dl, i & 15, 16);
    lam = (lane == (i & 15) && active) ? nl : lam;
    mu = fmaf(-cw, d, mu);
    u[0] = fmaf(-c0, d, u[0]); u[1] = fmaf(-c1, d, u[1]); u[2] = fmaf(-c2, d, u[2]);
  }
  return u[0] + u[1] + u[2] + mu + lam;
}
__global__ void kD(float* out, long long* cyc, float invd, float base, float lo, float hi, int n, unsigned mk) {
  __shared__ float tab[32 * 64];
  for (int k = threadIdx.x; k < 32 * 64; k += 32) tab[k] = 1e-3f * (k % 7);
  __syncwarp();
  const long long t0 = clock64();
  const float r = loopD(tab, threadIdx.x * 0.001f, invd, base, lo, hi, n, mk);
  const long long t1 = clock64();
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
}

// E: like D, but the trip count is loaded per thread (ptxas cannot prove the loop uniform -> BRA.DIV before each shuffle)
__device__ __noinline__ float loopE(const float* __restrict__ tab, float u0, float invd, float base, float lo, float hi, const int* np, unsigned mk) {
  float u[3] = {u0, u0 * 0.5f, u0 * 0.25f}, mu = 0.f, lam = 0.f;
  const int lane = threadIdx.x & 15;
  const int n = np[threadIdx.x];
#pragma unroll 4
  for (int i = 0; i < n; i++) {
    const float* row = tab + (i & 31) * 64;
    const float c0 = row[lane], c1 = row[16 + lane], c2 = row[32 + lane], cw = row[48 + lane];
    const bool active = (mk >> (i & 31)) & 1u;
    float nl = fmaf(u[0], invd, base);
    nl = fminf(fmaxf(nl, lo), hi);
    const float dl = active ? nl - lam : 0.f;
    const float d = __shfl_sync(FULL, dl, i & 15, 16);
    lam = (lane == (i & 15) && active) ? nl : lam;
    mu = fmaf(-cw, d, mu);
    u[0] = fmaf(-c0, d, u[0]); u[1] = fmaf(-c1, d, u[1]); u[2] = fmaf(-c2, d, u[2]);
  }
  return u[0] + u[1] + u[2] + mu + lam;
}
__global__ void kE(float* out, long long* cyc, float invd, float base, float lo, float hi, const int* np, unsigned mk) {
  __shared__ float tab[32 * 64];
  for (int k = threadIdx.x; k < 32 * 64; k += 32) tab[k] = 1e-3f * (k % 7);
  __syncwarp();
  const long long t0 = clock64();
  const float r = loopE(tab, threadIdx.x * 0.001f, invd, base, lo, hi, np, mk);
  const long long t1 = clock64();
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
}

__device__ __noinline__ float loopG(const float* __restrict__ tab, float u0, float invd, float base, float lo, float hi, int n, unsigned mk) {
  float u[3] = {u0, u0 * 0.5f, u0 * 0.25f}, mu = 0.f, lam = 0.f;
  const int lane = threadIdx.x & 15;
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    const float* row = tab + (i & 31) * 64;
    const float c0 = row[lane], c1 = row[16 + lane], c2 = row[32 + lane], cw = row[48 + lane];
    const bool active = (mk >> (i & 31)) & 1u;
    float nl = fmaf(u[0], invd, base);
    nl = fminf(fmaxf(nl, lo), hi);
    const float dl = active ? nl - lam : 0.f;
    const float d = __shfl_sync(FULL, dl, i & 15, 16);
    lam = (lane == (i & 15) && active) ? nl : lam;
    mu = fmaf(-cw, d, mu);
    u[0] = fmaf(-c0, d, u[0]); u[1] = fmaf(-c1, d, u[1]); u[2] = fmaf(-c2, d, u[2]);
  }
  return u[0] + u[1] + u[2] + mu + lam;
}
__global__ void kG(float* out, long long* cyc, float invd, float base, float lo, float hi, int n, unsigned mk) {
  __shared__ float tab[32 * 64];
  for (int k = threadIdx.x; k < 32 * 64; k += 32) tab[k] = 1e-3f * (k % 7);
  __syncwarp();
  const long long t0 = clock64();
  const float r = loopG(tab, threadIdx.x * 0.001f, invd, base, lo, hi, n, mk);
  const long long t1 = clock64();
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
}

// F: like E, not unrolled, but the trip count is loaded per thread (ptxas cannot prove the loop uniform -> BRA.DIV before each shuffle)
__device__ __noinline__ float loopF(const float* __restrict__ tab, float u0, float invd, float base, float lo, float hi, const int* np, unsigned mk) {
  float u[3] = {u0, u0 * 0.5f, u0 * 0.25f}, mu = 0.f, lam = 0.f;
  const int lane = threadIdx.x & 15;
  const int n = np[threadIdx.x];
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    const float* row = tab + (i & 31) * 64;
    const float c0 = row[lane], c1 = row[16 + lane], c2 = row[32 + lane], cw = row[48 + lane];
    const bool active = (mk >> (i & 31)) & 1u;
    float nl = fmaf(u[0], invd, base);
    nl = fminf(fmaxf(nl, lo), hi);
    const float dl = active ? nl - lam : 0.f;
    const float d = __shfl_sync(FULL, dl, i & 15, 16);
    lam = (lane == (i & 15) && active) ? nl : lam;
    mu = fmaf(-cw, d, mu);
    u[0] = fmaf(-c0, d, u[0]); u[1] = fmaf(-c1, d, u[1]); u[2] = fmaf(-c2, d, u[2]);
  }
  return u[0] + u[1] + u[2] + mu + lam;
}
__global__ void kF(float* out, long long* cyc, float invd, float base, float lo, float hi, const int* np, unsigned mk) {
  __shared__ float tab[32 * 64];
  for (int k = threadIdx.x; k < 32 * 64; k += 32) tab[k] = 1e-3f * (k % 7);
  __syncwarp();
  const long long t0 = clock64();
  const float r = loopF(tab, threadIdx.x * 0.001f, invd, base, lo, hi, np, mk);
  const long long t1 = clock64();
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 128); cudaMallocManaged(&cyc, 64);
  int* np; cudaMallocManaged(&np, 128); for (int i = 0; i < 32; i++) np[i] = N_ROWS;
  for (int rep = 0; rep < 2; rep++) {
    kA<<<1, 32>>>(out, cyc, 0.5f, 0.1f, -1.f, 1.f, 0.3f);
    kB<<<1, 32>>>(out, cyc, 0.5f, 0.1f, -1.f, 1.f, 0.3f, N_ROWS);
    kC<<<1, 32>>>(out, cyc, 0.5f, 0.1f, -1.f, 1.f, 0.3f);
    kD<<<1, 32>>>(out, cyc, 0.5f, 0.1f, -1.f, 1.f, N_ROWS, 0xffffffffu);
    kE<<<1, 32>>>(out, cyc, 0.5f, 0.1f, -1.f, 1.f, np, 0xffffffffu);
    kG<<<1, 32>>>(out, cyc, 0.5f, 0.1f, -1.f, 1.f, N_ROWS, 0xffffffffu);
    kF<<<1, 32>>>(out, cyc, 0.5f, 0.1f, -1.f, 1.f, np, 0xffffffffu);
    cudaDeviceSynchronize();
  }
  printf("cycles per row update: A (shuffle, kernel body) %.1f | B (shuffle, noinline: BRA.DIV) %.1f | C (smem broadcast) %.1f | D (solver-like loop) %.1f | E (D with per-thread trip count) %.1f | G (D, unroll 1) %.1f | F (E, unroll 1: BRA.DIV per row) %.1f\n",
         cyc[0] / (double)N_ROWS, cyc[1] / (double)N_ROWS, cyc[2] / (double)N_ROWS, cyc[3] / (double)N_ROWS, cyc[4] / (double)N_ROWS, cyc[5] / (double)N_ROWS, cyc[6] / (double)N_ROWS);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
