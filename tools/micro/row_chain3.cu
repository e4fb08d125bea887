// Round-2 microbenchmark, part 3: the LOOK-AHEAD row update for the tail of a launch.
//
// In the delta form (row_chain2.cu, variant H) lane r owns the running error u_r of row r; a row update is
//     dl_r = clamp(u_r invd_r + c_r, lo'_r, hi'_r);   broadcast dl_r (SHFL);   u_j -= A[j][r] dl_r  for all j
// and the dependent chain of consecutive rows is FFMA FMNMX FMNMX SHFL FFMA (~50 cycles, the SHFL is half of it, and more
// when the SM's MIO pipe is busy with other warps).
//
// Look-ahead (variant L<K>): every lane of the group computes every dl_r REDUNDANTLY, so nothing is broadcast on the chain.
// What a non-owner lacks is u_r; the owner ships a SNAPSHOT of it K rows early (one SHFL, off the chain), holding every impulse
// change up to row r-K-1, and every lane applies the last K changes itself, in the same order the owner does:
//     u = snapshot_r;  u = fma(-A[r][r-K], dl_{r-K}, u); ...; u = fma(-A[r][r-1], dl_{r-1}, u);  dl_r = clamp(...)
// -> same floating-point operations in the same order as the serial form (bit-identical iterates), chain = FFMA FFMA FMNMX
// FMNMX (no MIO instruction).  Per row: 1 SHFL + 2 group-uniform LDS.128 (band coefficients | invd, per-phase constants) +
// NSG + 1 column loads, K + 3 + NSG + 2 FP instructions.
//
// Measured for ONE warp alone on an SM, and for one warp sharing the SM with 16 warps that run the H loop (the situation of a
// tail block next to the main launch's blocks).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o row_chain3 row_chain3.cu ; run: ./row_chain3
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
#define N_ROWS 3072

__device__ __noinline__ float loopH(const float* __restrict__ tab, float u0, float invd, float cc, float lo, float hi, int n) {
  float u[3] = {u0, u0 * 0.5f, u0 * 0.25f}, mu = 0.f, lam = 0.f;
  const int lane = threadIdx.x & 15;
#pragma unroll 4
  for (int i = 0; i < n; i++) {
    const float* row = tab + (i & 31) * 64;
    const float c0 = row[lane], c1 = row[16 + lane], c2 = row[32 + lane], cw = row[48 + lane];
    const float dl = fminf(fmaxf(fmaf(u[0], invd, cc), lo), hi);
    const float d = __shfl_sync(FULL, dl, i & 15, 16);
    lam += (lane == (i & 15)) ? dl : 0.f;
    mu = fmaf(-cw, d, mu);
    u[0] = fmaf(-c0, d, u[0]); u[1] = fmaf(-c1, d, u[1]); u[2] = fmaf(-c2, d, u[2]);
  }
  return u[0] + u[1] + u[2] + mu + lam;
}

// look-ahead, K near terms (K <= 3: band = {A[r][r-3], A[r][r-2], A[r][r-1], invd_r})
template <int K, int UNR>
__device__ __noinline__ float loopL(const float* __restrict__ tab, const float4* __restrict__ band, const float4* __restrict__ cst,
                                    float u0, int n) {
  float u[3] = {u0, u0 * 0.5f, u0 * 0.25f}, mu = 0.f, lam = 0.f;
  const int lane = threadIdx.x & 15;
  float d1 = 0.f, d2 = 0.f, d3 = 0.f;
  float P[4];
  // prologue: snapshots of the first K rows
#pragma unroll
  for (int j = 0; j < K; j++) {
    const int gi = j % 48, s = gi >> 4;
    P[j] = __shfl_sync(FULL, s == 0 ? u[0] : (s == 1 ? u[1] : u[2]), gi & 15, 16);
  }
  int r = 0;   // row index (three sets of 16)
#pragma unroll UNR
  for (int i = 0; i < n; i++) {
    {
      const int gi = r + K >= 48 ? r + K - 48 : r + K, s = gi >> 4;
      P[K] = __shfl_sync(FULL, s == 0 ? u[0] : (s == 1 ? u[1] : u[2]), gi & 15, 16);
    }
    const float* row = tab + (i & 31) * 64;
    const float c0 = row[lane], c1 = row[16 + lane], c2 = row[32 + lane], cw = row[48 + lane];
    const float4 b = band[r], c = cst[r];
    float uu = P[0];
    if (K >= 3) uu = fmaf(-b.x, d3, uu);
    if (K >= 2) uu = fmaf(-b.y, d2, uu);
    uu = fmaf(-b.z, d1, uu);
    const float dl = fminf(fmaxf(fmaf(uu, b.w, c.x), c.y), c.z);
    lam += (lane == (r & 15)) ? dl : 0.f;    // (the set test is folded into the constants in the kernel)
    mu = fmaf(-cw, dl, mu);
    u[0] = fmaf(-c0, dl, u[0]); u[1] = fmaf(-c1, dl, u[1]); u[2] = fmaf(-c2, dl, u[2]);
    d3 = d2; d2 = d1; d1 = dl;
#pragma unroll
    for (int j = 0; j < K; j++) P[j] = P[j + 1];
    r = r + 1 == 48 ? 0 : r + 1;
  }
  return u[0] + u[1] + u[2] + mu + lam;
}

// look-ahead with the row's table entries loaded PF rows ahead (manual software pipelining: ptxas does not hoist the loads of a
// rolled loop across iterations, and an in-order warp stalls on the first consumer of a late load)
struct RowData { float4 b, c; float c0, c1, c2, cw; };
__device__ __forceinline__ RowData load_row(const float* tab, const float4* band, const float4* cst, int i, int r, int lane) {
  RowData d;
  const float* row = tab + (i & 31) * 64;
  d.c0 = row[lane]; d.c1 = row[16 + lane]; d.c2 = row[32 + lane]; d.cw = row[48 + lane];
  d.b = band[r]; d.c = cst[r];
  return d;
}
template <int PF, int UNR>
__device__ __noinline__ float loopM(const float* __restrict__ tab, const float4* __restrict__ band, const float4* __restrict__ cst,
                                    float u0, int n) {
  constexpr int K = 3;
  float u[3] = {u0, u0 * 0.5f, u0 * 0.25f}, mu = 0.f, lam = 0.f;
  const int lane = threadIdx.x & 15;
  float d1 = 0.f, d2 = 0.f, d3 = 0.f;
  float P[4];
#pragma unroll
  for (int j = 0; j < K; j++) P[j] = __shfl_sync(FULL, u[0], j, 16);
  RowData q[PF + 1];
#pragma unroll
  for (int j = 0; j < PF; j++) q[j] = load_row(tab, band, cst, j, j, lane);
  int r = 0, rp = PF;   // row index, prefetched row index
#pragma unroll UNR
  for (int i = 0; i < n; i++) {
    {
      const int gi = r + K >= 48 ? r + K - 48 : r + K, s = gi >> 4;
      P[K] = __shfl_sync(FULL, s == 0 ? u[0] : (s == 1 ? u[1] : u[2]), gi & 15, 16);
    }
    q[PF] = load_row(tab, band, cst, i + PF, rp, lane);
    const RowData d = q[0];
    float uu = fmaf(-d.b.x, d3, P[0]);
    uu = fmaf(-d.b.y, d2, uu);
    uu = fmaf(-d.b.z, d1, uu);
    const float dl = fminf(fmaxf(fmaf(uu, d.b.w, d.c.x), d.c.y), d.c.z);
    lam += (lane == (r & 15)) ? dl : 0.f;
    mu = fmaf(-d.cw, dl, mu);
    u[0] = fmaf(-d.c0, dl, u[0]); u[1] = fmaf(-d.c1, dl, u[1]); u[2] = fmaf(-d.c2, dl, u[2]);
    d3 = d2; d2 = d1; d1 = dl;
#pragma unroll
    for (int j = 0; j < K; j++) P[j] = P[j + 1];
#pragma unroll
    for (int j = 0; j < PF; j++) q[j] = q[j + 1];
    r = r + 1 == 48 ? 0 : r + 1;
    rp = rp + 1 == 48 ? 0 : rp + 1;
  }
  return u[0] + u[1] + u[2] + mu + lam;
}

template <int V>
__global__ void kern(float* out, long long* cyc, int slot, int n) {
  __shared__ float tab[36 * 64 + 512];
  __shared__ float4 band[2][48], cst[2][48];
  for (int i = threadIdx.x; i < 36 * 64 + 512; i += blockDim.x) tab[i] = 1e-3f * (i % 7);
  for (int i = threadIdx.x; i < 96; i += blockDim.x) {
    band[0][i] = make_float4(1e-3f, 2e-3f, 3e-3f, 0.5f);
    cst[0][i] = make_float4(0.1f, -1.f, 1.f, 0.f);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, g = (threadIdx.x >> 4) & 1;
  const long long t0 = clock64();
  float r;
  if (warp == 0) {
    if (V == 0) r = loopH(tab, threadIdx.x * 0.001f, 0.5f, 0.1f, -1.f, 1.f, n);
    if (V == 1) r = loopL<1, 4>(tab, band[g], cst[g], threadIdx.x * 0.001f, n);
    if (V == 2) r = loopL<2, 4>(tab, band[g], cst[g], threadIdx.x * 0.001f, n);
    if (V == 3) r = loopL<3, 4>(tab, band[g], cst[g], threadIdx.x * 0.001f, n);
    if (V == 4) r = loopL<3, 1>(tab, band[g], cst[g], threadIdx.x * 0.001f, n);
    if (V == 5) r = loopL<2, 1>(tab, band[g], cst[g], threadIdx.x * 0.001f, n);
    if (V == 6) r = loopM<1, 1>(tab, band[g], cst[g], threadIdx.x * 0.001f, n);
    if (V == 7) r = loopM<2, 1>(tab, band[g], cst[g], threadIdx.x * 0.001f, n);
    if (V == 8) r = loopM<1, 4>(tab, band[g], cst[g], threadIdx.x * 0.001f, n);
    if (V == 9) r = loopM<3, 4>(tab, band[g], cst[g], threadIdx.x * 0.001f, n);
    if (V == 10) r = loopM<2, 3>(tab, band[g], cst[g], threadIdx.x * 0.001f, n);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[slot] = t1 - t0;
  } else {
    r = loopH(tab, threadIdx.x * 0.001f, 0.5f, 0.1f, -1.f, 1.f, 3 * n);   // background load: outlasts warp 0
  }
  out[threadIdx.x] = r;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 8192); cudaMallocManaged(&cyc, 64 * sizeof(long long));
  const char* names[11] = {"H delta form", "L K=1 unroll 4", "L K=2 unroll 4", "L K=3 unroll 4", "L K=3 rolled", "L K=2 rolled",
                          "M K=3 prefetch 1", "M K=3 prefetch 2", "M pf 1 unroll 4", "M pf 3 unroll 4", "M pf 2 unroll 3"};
  for (int rep = 0; rep < 2; rep++) {
#define RUN(V) kern<V><<<1, 32>>>(out, cyc, V, N_ROWS); kern<V><<<1, 32 * 17>>>(out, cyc, 16 + V, N_ROWS); \
               kern<V><<<1, 32 * 9>>>(out, cyc, 32 + V, N_ROWS);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10)
    cudaDeviceSynchronize();
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  for (int v = 0; v < 11; v++)
    printf("%-18s lone warp %6.1f cycles/row | next to 8 busy warps %6.1f | next to 16 busy warps %6.1f\n", names[v],
           cyc[v] / (double)N_ROWS, cyc[32 + v] / (double)N_ROWS, cyc[16 + v] / (double)N_ROWS);
  return 0;
}
