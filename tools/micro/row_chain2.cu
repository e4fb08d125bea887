// Round-2 microbenchmark of PGS row-update formulations (see row_chain.cu for round 1).  Every kernel runs the same
// serial Gauss-Seidel recurrence over N_ROWS row updates with 16-lane groups (two environments per warp) and is timed
// for ONE warp alone on an SM (the tail of a launch) and for 16 warps on one SM (4 per scheduler: the bulk of a launch).
//   D : round-1 solver loop (select on `active`, lam select, nl - lam on the chain)        FFMA FMNMX FMNMX FADD SEL SHFL FFMA
//   H : delta form: dl = clamp(u * invd + c, lo', hi') with c = base - lam, lo' = lo - lam, hi' = hi - lam prepared per
//       sweep off the chain (inactive rows: lo' = hi' = 0)                                  FFMA FMNMX FMNMX SHFL FFMA
//   J : H with the broadcast done by redux.sync.max.f32 / min (one per group) instead of SHFL
//   K : replicated: every lane carries all R (= 12) running errors u_j, no broadcast at all: R FFMA per row
//   K2: K with packed fma.rn.f32x2
//   P : H, two rows per step: lane i+1 also carries row i's state and recomputes dl_i itself (one SHFL latency per 2 rows)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o row_chain2 row_chain2.cu ; run: ./row_chain2
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
#define N_ROWS 3072
#define NSG 3

struct Res { long long lone[8], full[8]; };

__device__ __forceinline__ void tick(long long* dst, int k, long long t0) {
  const long long t1 = clock64();
  if (threadIdx.x == 0) dst[k] = t1 - t0;
}

// ---- D: the round-1 loop -------------------------------------------------------------------------------------------
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
// ---- H: delta form -------------------------------------------------------------------------------------------------
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
// ---- J: redux broadcast ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float redux_bcast(float v, bool owner, unsigned mask) {
  // owner lane contributes v, the others -inf: max = v
  const float x = owner ? v : __int_as_float(0xff800000);
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "r"(mask));
  return r;
}
__device__ __noinline__ float loopJ(const float* __restrict__ tab, float u0, float invd, float cc, float lo, float hi, int n) {
  float u[3] = {u0, u0 * 0.5f, u0 * 0.25f}, mu = 0.f, lam = 0.f;
  const int lane = threadIdx.x & 15;
  const unsigned gm = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;
#pragma unroll 4
  for (int i = 0; i < n; i++) {
    const float* row = tab + (i & 31) * 64;
    const float c0 = row[lane], c1 = row[16 + lane], c2 = row[32 + lane], cw = row[48 + lane];
    const float dl = fminf(fmaxf(fmaf(u[0], invd, cc), lo), hi);
    const bool own = lane == (i & 15);
    const float d = redux_bcast(dl, own, gm);
    lam += own ? dl : 0.f;
    mu = fmaf(-cw, d, mu);
    u[0] = fmaf(-c0, d, u[0]); u[1] = fmaf(-c1, d, u[1]); u[2] = fmaf(-c2, d, u[2]);
  }
  return u[0] + u[1] + u[2] + mu + lam;
}
// ---- K: replicated, R = 12 rows ------------------------------------------------------------------------------------
#define RK 12
__device__ __noinline__ float loopK(const float* __restrict__ tab, float u0, float invd, float lo, float hi, int nsweeps) {
  float u[RK], lam[RK];
#pragma unroll
  for (int j = 0; j < RK; j++) { u[j] = u0 * (1.f + 0.1f * j); lam[j] = 0.f; }
  const int g = (threadIdx.x >> 4) & 1;
  const float4* A4 = reinterpret_cast<const float4*>(tab + g * 256);   // 12 x 12 (stride 16), per group
  for (int s = 0; s < nsweeps; s++) {
#pragma unroll
    for (int i = 0; i < RK; i++) {
      const float4 a0 = A4[i * 4], a1 = A4[i * 4 + 1], a2 = A4[i * 4 + 2];
      float nl = fmaf(u[i], invd, lam[i]);
      nl = fminf(fmaxf(nl, lo), hi);
      const float dl = nl - lam[i];
      lam[i] = nl;
      u[0] = fmaf(-a0.x, dl, u[0]); u[1] = fmaf(-a0.y, dl, u[1]); u[2] = fmaf(-a0.z, dl, u[2]); u[3] = fmaf(-a0.w, dl, u[3]);
      u[4] = fmaf(-a1.x, dl, u[4]); u[5] = fmaf(-a1.y, dl, u[5]); u[6] = fmaf(-a1.z, dl, u[6]); u[7] = fmaf(-a1.w, dl, u[7]);
      u[8] = fmaf(-a2.x, dl, u[8]); u[9] = fmaf(-a2.y, dl, u[9]); u[10] = fmaf(-a2.z, dl, u[10]); u[11] = fmaf(-a2.w, dl, u[11]);
    }
  }
  float r = 0.f;
#pragma unroll
  for (int j = 0; j < RK; j++) r += u[j] + lam[j];
  return r;
}
// ---- K2: replicated with packed FMA ----------------------------------------------------------------------------------
__device__ __forceinline__ void ffma2(float& x0, float& x1, float a0, float a1, float b) {
  // (x0, x1) += (a0, a1) * (b, b)
  unsigned long long x, a, bb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x) : "l"(a), "l"(bb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
}
__device__ __noinline__ float loopK2(const float* __restrict__ tab, float u0, float invd, float lo, float hi, int nsweeps) {
  float u[RK], lam[RK];
#pragma unroll
  for (int j = 0; j < RK; j++) { u[j] = u0 * (1.f + 0.1f * j); lam[j] = 0.f; }
  const int g = (threadIdx.x >> 4) & 1;
  const float4* A4 = reinterpret_cast<const float4*>(tab + g * 256);
  for (int s = 0; s < nsweeps; s++) {
#pragma unroll
    for (int i = 0; i < RK; i++) {
      const float4 a0 = A4[i * 4], a1 = A4[i * 4 + 1], a2 = A4[i * 4 + 2];
      float nl = fmaf(u[i], invd, lam[i]);
      nl = fminf(fmaxf(nl, lo), hi);
      const float dl = lam[i] - nl;    // -dl
      lam[i] = nl;
      ffma2(u[0], u[1], a0.x, a0.y, dl); ffma2(u[2], u[3], a0.z, a0.w, dl);
      ffma2(u[4], u[5], a1.x, a1.y, dl); ffma2(u[6], u[7], a1.z, a1.w, dl);
      ffma2(u[8], u[9], a2.x, a2.y, dl); ffma2(u[10], u[11], a2.z, a2.w, dl);
    }
  }
  float r = 0.f;
#pragma unroll
  for (int j = 0; j < RK; j++) r += u[j] + lam[j];
  return r;
}
// ---- P: two rows per broadcast --------------------------------------------------------------------------------------
__device__ __noinline__ float loopP(const float* __restrict__ tab, float u0, float invd, float cc, float lo, float hi, int n) {
  // lane L owns row L and mirrors row L-1 (up = its running error); rows are visited in pairs (i, i+1), i even
  float u[3] = {u0, u0 * 0.5f, u0 * 0.25f}, up = u0 * 0.9f, mu = 0.f, lam = 0.f;
  const int lane = threadIdx.x & 15;
  const float a_prev = tab[lane];   // A[L][L-1]
#pragma unroll 2
  for (int i = 0; i < n; i += 2) {
    const float* row = tab + (i & 31) * 64;
    const float c0 = row[lane], c1 = row[16 + lane], c2 = row[32 + lane], cw = row[48 + lane];
    const float e0 = row[64 + lane], e1 = row[80 + lane], e2 = row[96 + lane], ew = row[112 + lane];
    const float p0 = row[(lane + 15) & 15], p1 = row[64 + ((lane + 15) & 15)];       // coefficients of the mirrored row
    const bool second = lane == ((i + 1) & 15);
    const float dlp = fminf(fmaxf(fmaf(up, invd, cc), lo), hi);      // row i as seen by its mirror lane
    const float ue = fmaf(second ? -a_prev : 0.f, dlp, u[0]);
    const float dl = fminf(fmaxf(fmaf(ue, invd, cc), lo), hi);
    const float d0 = __shfl_sync(FULL, dl, i & 15, 16), d1 = __shfl_sync(FULL, dl, (i + 1) & 15, 16);
    lam += (lane == (i & 15) || second) ? dl : 0.f;
    mu = fmaf(-ew, d1, fmaf(-cw, d0, mu));
    u[0] = fmaf(-e0, d1, fmaf(-c0, d0, u[0])); u[1] = fmaf(-e1, d1, fmaf(-c1, d0, u[1])); u[2] = fmaf(-e2, d1, fmaf(-c2, d0, u[2]));
    up = fmaf(-p1, d1, fmaf(-p0, d0, up));
  }
  return u[0] + u[1] + u[2] + mu + lam + up;
}

template <int V>
__global__ void kern(float* out, long long* cyc, int k, float invd, float base, float lo, float hi, int n, unsigned mk) {
  __shared__ float tab[36 * 64 + 512];
  for (int i = threadIdx.x; i < 36 * 64 + 512; i += blockDim.x) tab[i] = 1e-3f * (i % 7);
  __syncthreads();
  const long long t0 = clock64();
  float r;
  if (V == 0) r = loopD(tab, threadIdx.x * 0.001f, invd, base, lo, hi, n, mk);
  if (V == 1) r = loopH(tab, threadIdx.x * 0.001f, invd, base, lo, hi, n);
  if (V == 2) r = loopJ(tab, threadIdx.x * 0.001f, invd, base, lo, hi, n);
  if (V == 3) r = loopK(tab, threadIdx.x * 0.001f, invd, lo, hi, n / RK);
  if (V == 4) r = loopK2(tab, threadIdx.x * 0.001f, invd, lo, hi, n / RK);
  if (V == 5) r = loopP(tab, threadIdx.x * 0.001f, invd, base, lo, hi, n);
  __syncthreads();
  tick(cyc, k, t0);
  out[threadIdx.x] = r;
}

int main() {
  float* out; Res* res;
  cudaMalloc(&out, 4096); cudaMallocManaged(&res, sizeof(Res));
  const char* names[6] = {"D round-1 loop", "H delta form", "J redux broadcast", "K replicated R=12", "K2 replicated f32x2", "P two rows per shuffle"};
  for (int rep = 0; rep < 2; rep++) {
#define RUN(V) kern<V><<<1, 32>>>(out, res->lone, V, 0.5f, 0.1f, -1.f, 1.f, N_ROWS, 0xffffffffu); \
               kern<V><<<1, 512>>>(out, res->full, V, 0.5f, 0.1f, -1.f, 1.f, N_ROWS, 0xffffffffu);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    cudaDeviceSynchronize();
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  for (int v = 0; v < 6; v++)
    printf("%-26s lone warp %6.1f cycles/row | 16 warps on the SM %6.1f cycles/row per warp (%.1f per scheduler slot)\n", names[v],
           res->lone[v] / (double)N_ROWS, res->full[v] / (double)N_ROWS, res->full[v] / (double)N_ROWS / 4);
  return 0;
}
