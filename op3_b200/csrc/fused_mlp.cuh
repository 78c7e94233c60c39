// CTA-resident chains of small dense layers (Mlp.forward, op3/torch/networks.py:63-74, and its backward).
//
// The dynamics network (physics_network.py:94-141: eleven one-hidden-layer MLPs over M = B*K slot rows and B*K*(K-1) pair
// rows) and the head of the refinement network (refinement_network.py:203-219: FC, LSTM cell, two linear heads) are row-local
// chains of layers no wider than 512: launched layer by layer they are pure launch latency (a training step of the pickplace
// variant spent 370 launches and 5 ms on them).  Here ONE CTA owns a tile of up to 16 rows (for the dynamics: every slot and
// every ordered slot pair of the same samples) and walks the whole chain with the activations in shared memory; only the
// weights stream in (L2 -> shared memory, cp.async, 4-stage ring of 128 x 32 tiles) and only what the backward pass or the
// caller needs is written out.  Two arithmetic paths behind one interface: fp32 FMA (exact mode; its inner loop is bound by the
// shared-memory reads of the activation panel, 1 LDS.128 per 4 FMAs) and warp-level tensor-core MMAs with 3xTF32 operand
// splitting (every other mode; fp32-grade like the convolutions of those modes, 6 scalar LDS per 3 MMAs).
// Panels are [16][pitch] with pitch = 4 mod 32 floats (PITCH()) so that the MMA fragment loads are conflict-free.
#pragma once
#include "common.cuh"

namespace op3 {
namespace fused {

#ifndef OP3_FUSED_WARPS
#define OP3_FUSED_WARPS 16              /* warps per CTA (8 or 16): the chains are latency-bound, a warp owns 8 output columns */
#endif
constexpr int FT = 32 * OP3_FUSED_WARPS; // threads per CTA: TN output columns x 4 contraction groups
constexpr int RTM = 16;                 // rows of one activation panel
constexpr int TN = 8 * OP3_FUSED_WARPS, TK = 32;   // weight tile: TN output columns x 32 contraction steps
constexpr int PNT = TK + 4;             // pitch of a [column][k] tile (float4 reads, conflict-free over 8 lanes)
constexpr int PNN = TN + 8;             // pitch of a [k][column] tile (8 mod 32: conflict-free mma B fragments)
constexpr int WTILE = TN * PNT > TK * PNN ? TN * PNT : TK * PNN;
static_assert(FT == 4 * TN && PNN % 32 == 8, "thread maps below: 4 contraction groups of TN threads, two 16-byte chunks per thread and tile");
constexpr int NST = 4;                  // tiles in flight
constexpr int GEMM_SMEM_FLOATS = NST * WTILE + 4 * RTM * TN;     // weight ring + cross-group reduction buffer

// What happens to one output value v = sum_k X[r][k] * Wop[k][n], in this order (every step optional):
//   v += bias[n]; v = act(v); v *= ELU'(gy[r][n]) for n < gy_cols; v += so[r][n] (so_acc); so[r][n] = v; go[r][n] (+)= v.
// A plain struct instead of a functor: ONE out-of-line copy of the GEMM loop serves all ~25 layers of a chain.  (Inlined per
// layer the dynamics kernel was 300 KB of straight-line code walked once -- it ran at instruction-cache-miss speed, 1 us per
// weight tile.)
struct Epi {
  const float* bias = nullptr;
  int act = ACT_NONE;
  const float* gy = nullptr;   // activation OUTPUT whose ELU' multiplies v (backward), global, pre-offset to the row tile
  int ldgy = 0, gy_cols = 1 << 30;
  float* so = nullptr;         // shared-memory panel
  int ldso = 0, so_acc = 0;
  float* go = nullptr;         // global rows, pre-offset to the row tile
  int ldgo = 0, go_acc = 0;
};
__device__ __forceinline__ Epi epi_fwd(const float* bias, int act, float* so, int ldso, float* go, int ldgo) {
  Epi e; e.bias = bias; e.act = act; e.so = so; e.ldso = ldso; e.go = go; e.ldgo = ldgo; return e;
}
__device__ __forceinline__ Epi epi_bwd(const float* gy, int ldgy, float* so, int ldso, int so_acc, float* go, int ldgo, int go_acc) {
  Epi e; e.gy = gy; e.ldgy = ldgy; e.so = so; e.ldso = ldso; e.so_acc = so_acc; e.go = go; e.ldgo = ldgo; e.go_acc = go_acc; return e;
}

#ifdef OP3_FUSED_TRACE
__device__ unsigned long long g_fused_trace[8];   // [0] kernel cycles, [1] inside cta_gemm, [2] tile waits, [3] epilogues, [4] tiles, [5] calls
#define FTRACE(i, v) do { if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&g_fused_trace[i], (unsigned long long)(v)); } while (0)
#else
#define FTRACE(i, v) do { } while (0)
#endif

// Y[r][n] = sum_k X[r][k] * Wop[k][n] for r < rows (<= RT), n < ncols, each handed to the epilogue exactly once.
//   NN = false: Wop[k][n] = W[n * ldw + k]   (y = x W^T with an nn.Linear weight [out][in]: forward)
//   NN = true : Wop[k][n] = W[k * ldw + n]   (dx = dy W: backward)
// Xs: shared-memory panel of RT rows, row pitch ldx (multiple of 4), readable and FINITE up to column roundup32(kdim) in all RT
// rows (rows >= `rows` are computed and dropped: the row loop has no branches, so the loads of all rows are in flight together).
// kdim % 4 == 0 (NN = false) / ncols % 4 == 0 (NN = true); W rows 16-byte aligned.  `wt`: GEMM_SMEM_FLOATS floats.
// The summation order is fixed (k ascending inside each of four interleaved groups, groups added 0..3): deterministic and
// independent of RT.  Ends with a CTA barrier: what the epilogue wrote is visible to every thread afterwards.
template <bool NN, int RT>
__device__ __noinline__ void cta_gemm_rt(const float* Xs, int ldx, int rows, const float* __restrict__ W, int ldw, int kdim,
                                         int ncols, float* wt, const Epi ep) {
  static_assert((NST & (NST - 1)) == 0, "NST must be a power of two");
  const int tid = threadIdx.x, n = tid % TN, kq = tid / TN;
  float* red = wt + NST * WTILE;
  const int nkt = (kdim + TK - 1) / TK, nnt = (ncols + TN - 1) / TN, total = nkt * nnt;
  // every thread copies two 16-byte chunks per tile: (row crow, crow + rstep) x 4 consecutive elements at ccol
  const int crow = NN ? tid / (TN / 4) : (tid >> 3);
  const int ccol = NN ? (tid % (TN / 4)) * 4 : (tid & 7) * 4;
  constexpr int rstep = NN ? 4 * FT / TN : FT / 8;
  const int soff = NN ? crow * PNN + ccol : crow * PNT + ccol;
  constexpr int sstep = NN ? rstep * PNN : rstep * PNT;
  int pt = 0, pn0 = 0, pk0 = 0;                              // prefetch cursor: tile index, its first column and first k
  auto issue = [&]() {
    if (pt < total) {
      float* dst = wt + (pt & (NST - 1)) * WTILE + soff;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int off;
        bool ok;
        if (!NN) {
          const int nn = pn0 + crow + i * rstep, kk = pk0 + ccol;
          ok = nn < ncols && kk < kdim;
          off = nn * ldw + kk;
        } else {
          const int kk = pk0 + crow + i * rstep, nn = pn0 + ccol;
          ok = kk < kdim && nn < ncols;
          off = kk * ldw + nn;
        }
        cp_async16(dst + i * sstep, W + (ok ? off : 0), ok);
      }
      ++pt;
      pk0 += TK;
      if (pk0 >= kdim) { pk0 = 0; pn0 += TN; }
    }
    cp_async_commit();                  // always: keeps the group count in step with the tile index
  };
#ifdef OP3_FUSED_TRACE
  const long long tr0 = clock64();
  FTRACE(4, total); FTRACE(5, 1);
#endif
#pragma unroll
  for (int s = 0; s < NST - 1; ++s) issue();
  int t = 0;
  for (int nt = 0; nt < nnt; ++nt) {
    float acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = 0.f;
    for (int kt = 0; kt < nkt; ++kt, ++t) {
#ifdef OP3_FUSED_TRACE
      const long long tw0 = clock64();
#endif
      cp_async_wait<NST - 2>();
      __syncthreads();                  // tile t has landed; everyone is done with the buffer tile t + NST - 1 overwrites
#ifdef OP3_FUSED_TRACE
      FTRACE(2, clock64() - tw0);
#endif
#ifdef OP3_FUSED_TRACE
      const long long ti0 = clock64();
#endif
      issue();
#ifdef OP3_FUSED_TRACE
      const long long ti1 = clock64();
      FTRACE(6, ti1 - ti0);
#endif
      const float* ws = wt + (t & (NST - 1)) * WTILE;
      float w[8];
      if (!NN) {
        const float4 a = *reinterpret_cast<const float4*>(ws + n * PNT + kq * 8);
        const float4 b = *reinterpret_cast<const float4*>(ws + n * PNT + kq * 8 + 4);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = ws[(kq * 8 + j) * PNN + n];
      }
      const float* xp = Xs + kt * TK + kq * 8;
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const float4 x0 = *reinterpret_cast<const float4*>(xp + r * ldx);
        const float4 x1 = *reinterpret_cast<const float4*>(xp + r * ldx + 4);
        float a = acc[r];
        a = fmaf(w[0], x0.x, a); a = fmaf(w[1], x0.y, a); a = fmaf(w[2], x0.z, a); a = fmaf(w[3], x0.w, a);
        a = fmaf(w[4], x1.x, a); a = fmaf(w[5], x1.y, a); a = fmaf(w[6], x1.z, a); a = fmaf(w[7], x1.w, a);
        acc[r] = a;
      }
#ifdef OP3_FUSED_TRACE
      FTRACE(7, clock64() - ti1);
#endif
    }
#ifdef OP3_FUSED_TRACE
    const long long te0 = clock64();
#endif
#pragma unroll
    for (int r = 0; r < RT; ++r) red[(kq * RT + r) * TN + n] = acc[r];
    __syncthreads();
    for (int e = tid; e < rows * TN; e += FT) {
      const int r = e / TN, i = e % TN, c = nt * TN + i;
      if (c < ncols) {
        float v = ((red[(0 * RT + r) * TN + i] + red[(1 * RT + r) * TN + i]) + red[(2 * RT + r) * TN + i]) +
                  red[(3 * RT + r) * TN + i];
        if (ep.bias) v += __ldg(ep.bias + c);
        if (ep.act == ACT_ELU) v = elu_f(v);
        else if (ep.act == ACT_SIGMOID) v = sigmoid_f(v);
        if (ep.gy && c < ep.gy_cols) v *= elu_grad_from_out(ep.gy[(size_t)r * ep.ldgy + c]);
        if (ep.so) {
          if (ep.so_acc) v += ep.so[r * ep.ldso + c];
          ep.so[r * ep.ldso + c] = v;
        }
        if (ep.go) {
          float* g = ep.go + (size_t)r * ep.ldgo + c;
          *g = ep.go_acc ? *g + v : v;
        }
      }
    }
#ifdef OP3_FUSED_TRACE
    FTRACE(3, clock64() - te0);
#endif
  }
  cp_async_wait<0>();
  __syncthreads();
#ifdef OP3_FUSED_TRACE
  FTRACE(1, clock64() - tr0);
#endif
}

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// D (16x8, fp32) += A (16x8, tf32, row) * B (8x8, tf32, col)
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Same contract as cta_gemm_rt on the tensor cores: warp w owns columns 8w .. 8w+7 of the TN-column tile and all 16 panel
// rows (one m16n8k8 accumulator, no cross-warp reduction).  Operands are split x = hi + lo with hi = rn_tf32(x), lo =
// rn_tf32(x - hi) and the three products lo*hi + hi*lo + hi*hi accumulate in fp32 (3xTF32: ~1e-6 relative, like the tf32x3
// convolutions).  ldx = 4 mod 32 keeps the A-fragment loads conflict-free.
template <bool NN>
__device__ __noinline__ void cta_gemm_mma(const float* Xs, int ldx, int rows, const float* __restrict__ W, int ldw, int kdim,
                                          int ncols, float* wt, const Epi ep) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  const int nkt = (kdim + TK - 1) / TK, nnt = (ncols + TN - 1) / TN, total = nkt * nnt;
  const int crow = NN ? tid / (TN / 4) : (tid >> 3);
  const int ccol = NN ? (tid % (TN / 4)) * 4 : (tid & 7) * 4;
  constexpr int rstep = NN ? 4 * FT / TN : FT / 8;
  const int soff = NN ? crow * PNN + ccol : crow * PNT + ccol;
  constexpr int sstep = NN ? rstep * PNN : rstep * PNT;
  int pt = 0, pn0 = 0, pk0 = 0;
  auto issue = [&]() {
    if (pt < total) {
      float* dst = wt + (pt & (NST - 1)) * WTILE + soff;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int off;
        bool ok;
        if (!NN) {
          const int nn = pn0 + crow + i * rstep, kk = pk0 + ccol;
          ok = nn < ncols && kk < kdim;
          off = nn * ldw + kk;
        } else {
          const int kk = pk0 + crow + i * rstep, nn = pn0 + ccol;
          ok = kk < kdim && nn < ncols;
          off = kk * ldw + nn;
        }
        cp_async16(dst + i * sstep, W + (ok ? off : 0), ok);
      }
      ++pt;
      pk0 += TK;
      if (pk0 >= kdim) { pk0 = 0; pn0 += TN; }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < NST - 1; ++s) issue();
  const float* xa = Xs + gid * ldx + tig;
  const float* xb = xa + 8 * ldx;
  int t = 0;
  for (int nt = 0; nt < nnt; ++nt) {
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    for (int kt = 0; kt < nkt; ++kt, ++t) {
      cp_async_wait<NST - 2>();
      __syncthreads();
      issue();
      const float* ws = wt + (t & (NST - 1)) * WTILE;
      const float* wp = NN ? ws + tig * PNN + 8 * warp + gid : ws + (8 * warp + gid) * PNT + tig;
#pragma unroll
      for (int ks = 0; ks < TK / 8; ++ks) {
        const int kk = kt * TK + ks * 8;
        float xr[4] = {xa[kk], xb[kk], xa[kk + 4], xb[kk + 4]};
        float wr[2];
        if (NN) { wr[0] = wp[ks * 8 * PNN]; wr[1] = wp[(ks * 8 + 4) * PNN]; }
        else { wr[0] = wp[ks * 8]; wr[1] = wp[ks * 8 + 4]; }
        uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ah[i] = f2tf32(xr[i]); al[i] = f2tf32(xr[i] - __uint_as_float(ah[i])); }
#pragma unroll
        for (int i = 0; i < 2; ++i) { bh[i] = f2tf32(wr[i]); bl[i] = f2tf32(wr[i] - __uint_as_float(bh[i])); }
        mma_tf32(c, al, bh);
        mma_tf32(c, ah, bl);
        mma_tf32(c, ah, bh);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = gid + (q >> 1) * 8, col = nt * TN + 8 * warp + 2 * tig + (q & 1);
      if (r < rows && col < ncols) {
        float v = c[q];
        if (ep.bias) v += __ldg(ep.bias + col);
        if (ep.act == ACT_ELU) v = elu_f(v);
        else if (ep.act == ACT_SIGMOID) v = sigmoid_f(v);
        if (ep.gy && col < ep.gy_cols) v *= elu_grad_from_out(ep.gy[(size_t)r * ep.ldgy + col]);
        if (ep.so) {
          if (ep.so_acc) v += ep.so[r * ep.ldso + col];
          ep.so[r * ep.ldso + col] = v;
        }
        if (ep.go) {
          float* g = ep.go + (size_t)r * ep.ldgo + col;
          *g = ep.go_acc ? *g + v : v;
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

// `tc`: tensor-core path (every precision mode but the exact fp32 one).  fp32 path: the instance with the next row tile
// (4, 8, 16) -- a chain over 4 slot rows should not pay for 16.
template <bool NN>
__device__ __forceinline__ void cta_gemm(bool tc, const float* Xs, int ldx, int rows, const float* __restrict__ W, int ldw,
                                         int kdim, int ncols, float* wt, const Epi& ep) {
  if (tc) cta_gemm_mma<NN>(Xs, ldx, rows, W, ldw, kdim, ncols, wt, ep);
  else if (rows <= 4) cta_gemm_rt<NN, 4>(Xs, ldx, rows, W, ldw, kdim, ncols, wt, ep);
  else if (rows <= 8) cta_gemm_rt<NN, 8>(Xs, ldx, rows, W, ldw, kdim, ncols, wt, ep);
  else cta_gemm_rt<NN, 16>(Xs, ldx, rows, W, ldw, kdim, ncols, wt, ep);
}

// panel row pitch for `w` columns: w + 4 floats (= 4 mod 32 for the widths used here, all multiples of 32)
__host__ __device__ constexpr int PITCH(int w) { return w + 4; }

// rows x cols floats from global (row pitch ldg) into a shared panel (row pitch lds); cols % 4 == 0, 16-byte aligned rows
__device__ __forceinline__ void load_panel(float* Ps, int lds, const float* g, int ldg, int rows, int cols) {
  const int c4 = cols >> 2;
  for (int e = threadIdx.x; e < rows * c4; e += FT) {
    const int r = e / c4, c = (e - r * c4) * 4;
    *reinterpret_cast<float4*>(Ps + r * lds + c) = *reinterpret_cast<const float4*>(g + (size_t)r * ldg + c);
  }
}
__device__ __forceinline__ void zero_panel(float* Ps, int n) {
  for (int e = threadIdx.x; e < n; e += FT) Ps[e] = 0.f;
}

__device__ __forceinline__ float act_apply(float v, int act) {
  return act == ACT_ELU ? elu_f(v) : (act == ACT_SIGMOID ? sigmoid_f(v) : v);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace fused
}  // namespace op3
