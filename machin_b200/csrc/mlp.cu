// Small-MLP forward / backward for the learner's Q / critic networks, float32 SIMT (sm_100a).
//
// The reference leaves the value networks to the user's nn.Module; for the plain Linear+ReLU stacks its
// tests and examples use (test_apex.py:25-40, examples/framework_examples/dqn_apex.py) the eager path is
// ~75 library launches per update (cuBLAS sgemm + epilogue + clamp + bias-grad reductions), which is what
// the learner step spends its time on once replay sampling is on the device.  These kernels cut that to one
// launch per layer per direction.  float32 FFMA with a fixed summation order (no tensor cores: TF32/bf16
// would break the 1e-5 parity bar; no atomics on results: split-K partials are summed in a fixed order by
// the last CTA to arrive).
//
// One generic tiled GEMM   C[i, j] = sum_k A(i, k) * B(j, k)   with arbitrary strides:
//   forward        y[m, n]  = sum_k x[m, k]  * w[n, k]   (+ bias[n], ReLU)
//   input grad     dx[m, k] = sum_n dy[m, n] * w[n, k]   (* (act[m, k] > 0))
//   weight grad    dw[n, k] = sum_m dy[m, n] * x[m, k]   ;  db[n] = sum_m dy[m, n]  (virtual column of ones)
// Tile 64x64, 256 threads, 4x4 micro-tile per thread, K-step 16; split-K over gridDim.z so that the grid
// reaches ~1 CTA per SM even though the matrices are small.
#include <cooperative_groups.h>

#include <stdlib.h>

#include <new>
#include <vector>

#include "mb_common.cuh"
#include "td_device.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int TM = 64, TN = 64, TK = 32, NT = 256;
constexpr int MT = TM / 16;   // register tile per thread (16 x 16 threads): 4 x 4 for 64-wide tiles, 8 x 8 for 128

struct GemmArgs {
    const float *A;
    const float *B;
    float *C;
    int64_t a_i, a_k;      // A(i, k) = A[i * a_i + k * a_k]
    int64_t b_j, b_k;      // B(j, k) = B[j * b_j + k * b_k]
    int64_t c_i;           // C[i * c_i + j]
    int M, N, K;           // output M x N, reduction K
    int ones_col;          // >= 0: B(ones_col, k) == 1 for all k, and that output column goes to `extra`
    const float *bias;     // per-column bias (forward) or NULL
    const float *mask;     // same shape as C: keep acc where mask > 0 (input grad) or NULL
    float *extra;          // db target when ones_col >= 0
    int relu;
    float *partial;        // [splits, M, N] workspace when splits > 1
    int *tile_counter;     // one int per output tile, zero between launches
    int k_per_split;
};

__device__ __forceinline__ void epilogue_store(const GemmArgs &g, int i, int j, float v) {
    if (i >= g.M || j >= g.N) return;
    if (g.ones_col >= 0 && j == g.ones_col) {
        g.extra[i] = v;
        return;
    }
    if (g.bias) v = __fadd_rn(v, g.bias[j]);
    if (g.relu) v = fmaxf(v, 0.0f);
    if (g.mask) v = g.mask[(int64_t)i * g.c_i + j] > 0.0f ? v : 0.0f;
    g.C[(int64_t)i * g.c_i + j] = v;
}

struct GemmSmem {
    float As[TK][TM + 4];
    float Bs[TK][TN + 4];
    int last;
};

// One work item = one 64x64 output tile x one K-split.  Called by the stand-alone GEMM kernels and by
// the persistent program executor below.
template <bool kAKfast, bool kBKfast>
__device__ __forceinline__ void gemm_item(const GemmArgs &g, GemmSmem &sm, int tile_x, int tile_y, int tiles_x,
                                          int split, int splits) {
    const int tid = threadIdx.x;
    const int i0 = tile_y * TM, j0 = tile_x * TN;
    const int k_begin = split * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);
    // 16 x 16 threads, 8 x 8 outputs each, as two 4-wide halves 64 apart: with an 8 x 8 register tile the
    // inner loop needs 4 LDS.128 per 64 FFMA and is FFMA-bound (a 4 x 4 tile is shared-memory-bandwidth bound:
    // measured 86 cycles per k for 8 warps); the 64-apart halves keep every LDS.128 contiguous across lanes.
    const int ty = tid >> 4, tx = tid & 15;
    float acc[MT][MT];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b) acc[a][b] = 0.0f;
    // Global -> register -> shared staging, software-pipelined: all loads of a K-step are issued back to
    // back (one memory latency per step), and the next step's loads are in flight while this one is multiplied.
    constexpr int UA = TM * TK / NT, UB = TN * TK / NT;
    float ra[UA], rb[UB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < UA; ++u) {
            const int e = tid + u * NT;
            int i, k;
            if (kAKfast) { i = e / TK; k = e % TK; } else { k = e / TM; i = e % TM; }
            const int gi = i0 + i, gk = k0 + k;
            ra[u] = (gi < g.M && gk < k_end) ? __ldcg(g.A + (int64_t)gi * g.a_i + (int64_t)gk * g.a_k) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int e = tid + u * NT;
            int j, k;
            if (kBKfast) { j = e / TK; k = e % TK; } else { k = e / TN; j = e % TN; }
            const int gj = j0 + j, gk = k0 + k;
            float v = 0.0f;
            if (gj < g.N && gk < k_end)
                v = (gj == g.ones_col) ? 1.0f : __ldcg(g.B + (int64_t)gj * g.b_j + (int64_t)gk * g.b_k);
            rb[u] = v;
        }
    };
    auto park = [&]() {
#pragma unroll
        for (int u = 0; u < UA; ++u) {
            const int e = tid + u * NT;
            if (kAKfast) sm.As[e % TK][e / TK] = ra[u]; else sm.As[e / TM][e % TM] = ra[u];
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int e = tid + u * NT;
            if (kBKfast) sm.Bs[e % TK][e / TK] = rb[u]; else sm.Bs[e / TN][e % TN] = rb[u];
        }
    };
    if (k_begin < k_end) fetch(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += TK) {
        park();
        __syncthreads();
        if (k0 + TK < k_end) fetch(k0 + TK);
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float av[MT], bv[MT];
#pragma unroll
            for (int h = 0; h < MT / 4; ++h) {
                const float4 a = *reinterpret_cast<const float4 *>(&sm.As[k][h * 64 + ty * 4]);
                const float4 b = *reinterpret_cast<const float4 *>(&sm.Bs[k][h * 64 + tx * 4]);
                av[4 * h] = a.x; av[4 * h + 1] = a.y; av[4 * h + 2] = a.z; av[4 * h + 3] = a.w;
                bv[4 * h] = b.x; bv[4 * h + 1] = b.y; bv[4 * h + 2] = b.z; bv[4 * h + 3] = b.w;
            }
#pragma unroll
            for (int x = 0; x < MT; ++x)
#pragma unroll
                for (int y = 0; y < MT; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
        }
        __syncthreads();
    }
    auto row_of = [&](int x) { return i0 + (x / 4) * 64 + ty * 4 + (x & 3); };
    auto col_of = [&](int y) { return j0 + (y / 4) * 64 + tx * 4 + (y & 3); };
    if (splits == 1) {
#pragma unroll
        for (int x = 0; x < MT; ++x)
#pragma unroll
            for (int y = 0; y < MT; ++y) epilogue_store(g, row_of(x), col_of(y), acc[x][y]);
        return;
    }
    // split-K: park the partial tile, the last CTA of this tile sums the partials in split order
    const int64_t MN = (int64_t)g.M * g.N;
    float *mine = g.partial + (int64_t)split * MN;
#pragma unroll
    for (int x = 0; x < MT; ++x)
#pragma unroll
        for (int y = 0; y < MT; ++y) {
            const int i = row_of(x), j = col_of(y);
            if (i < g.M && j < g.N) mine[(int64_t)i * g.N + j] = acc[x][y];
        }
    __threadfence();
    __syncthreads();
    const int tile = tile_y * tiles_x + tile_x;
    if (tid == 0) sm.last = (atomicAdd(g.tile_counter + tile, 1) == splits - 1);
    __syncthreads();
    if (!sm.last) return;
    __threadfence();
    // splits outermost: the MT*MT loads of one split are independent and in flight together (with the
    // outputs outermost this was splits * MT * MT dependent L2 round trips: ~20 us for a 4-way split)
#pragma unroll
    for (int x = 0; x < MT; ++x)
#pragma unroll
        for (int y = 0; y < MT; ++y) acc[x][y] = 0.0f;
    for (int sp = 0; sp < splits; ++sp) {
        const float *part = g.partial + (int64_t)sp * MN;
#pragma unroll
        for (int x = 0; x < MT; ++x)
#pragma unroll
            for (int y = 0; y < MT; ++y) {
                const int i = row_of(x), j = col_of(y);
                if (i < g.M && j < g.N) acc[x][y] = __fadd_rn(acc[x][y], __ldcg(part + (int64_t)i * g.N + j));
            }
    }
#pragma unroll
    for (int x = 0; x < MT; ++x)
#pragma unroll
        for (int y = 0; y < MT; ++y) epilogue_store(g, row_of(x), col_of(y), acc[x][y]);
    if (tid == 0) g.tile_counter[tile] = 0;
    __syncthreads();
}

template <bool kAKfast, bool kBKfast>
__global__ void __launch_bounds__(NT) gemm_kernel(GemmArgs g) {
    __shared__ GemmSmem sm;
    gemm_item<kAKfast, kBKfast>(g, sm, blockIdx.x, blockIdx.y, gridDim.x, blockIdx.z, gridDim.z);
}

// ---- small-problem GEMM: the learner's layers are 512 x 256 x 256 and smaller -----------------------------------
// With 64 x 64 tiles such a layer is 32 tiles (64 with a grid-level K-split and its reduction pass): most of the 148
// SMs idle and each CTA walks the whole K serially -- measured 11 us per layer.  Here the tile is 32 x 32 (128 CTAs for
// 512 x 256), and the K range of a tile is cut over KG groups of 64 threads INSIDE the CTA: each group multiplies its
// K-slice out of its own shared-memory stage (same register-staged software pipeline as above, 4 x 4 outputs per
// thread), the groups' partial tiles are added in group order through shared memory (a fixed summation order: the
// result does not depend on scheduling), and group 0 runs the fused epilogue.  No workspace, no second pass.
constexpr int ST = 32, SG = 64, STK = 32;     // tile edge, threads per K-group, K-step

template <int KG>
struct SmallSmem {
    float As[KG][STK][ST + 4];
    float Bs[KG][STK][ST + 4];
};

template <int KG, bool kAKfast, bool kBKfast>
__global__ void __launch_bounds__(SG * KG) gemm_small_kernel(GemmArgs g) {
    __shared__ SmallSmem<KG> sm;
    const int tid = threadIdx.x, grp = tid / SG, lt = tid % SG;
    const int i0 = blockIdx.y * ST, j0 = blockIdx.x * ST;
    const int k_begin = grp * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);
    const int ty = lt >> 3, tx = lt & 7;      // 8 x 8 threads, 4 x 4 outputs each
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
    constexpr int U = ST * STK / SG;          // 16 elements of A and of B per thread and K-step
    float ra[U], rb[U];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = lt + u * SG;
            int i, k;
            if (kAKfast) { i = e / STK; k = e % STK; } else { k = e / ST; i = e % ST; }
            const int gi = i0 + i, gk = k0 + k;
            ra[u] = (gi < g.M && gk < k_end) ? __ldcg(g.A + (int64_t)gi * g.a_i + (int64_t)gk * g.a_k) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = lt + u * SG;
            int j, k;
            if (kBKfast) { j = e / STK; k = e % STK; } else { k = e / ST; j = e % ST; }
            const int gj = j0 + j, gk = k0 + k;
            float v = 0.0f;
            if (gj < g.N && gk < k_end)
                v = (gj == g.ones_col) ? 1.0f : __ldcg(g.B + (int64_t)gj * g.b_j + (int64_t)gk * g.b_k);
            rb[u] = v;
        }
    };
    auto park = [&]() {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = lt + u * SG;
            if (kAKfast) sm.As[grp][e % STK][e / STK] = ra[u]; else sm.As[grp][e / ST][e % ST] = ra[u];
            if (kBKfast) sm.Bs[grp][e % STK][e / STK] = rb[u]; else sm.Bs[grp][e / ST][e % ST] = rb[u];
        }
    };
    // every group runs the same number of steps (g.k_per_split is a multiple of STK; an empty slice multiplies zeros),
    // so the CTA-wide barriers line up
    const int steps = g.k_per_split / STK;
    if (k_begin < k_end) fetch(k_begin);
    else {
#pragma unroll
        for (int u = 0; u < U; ++u) ra[u] = rb[u] = 0.0f;
    }
    for (int s = 0; s < steps; ++s) {
        const int k0 = k_begin + s * STK;
        park();
        __syncthreads();
        if (s + 1 < steps) {
            if (k0 + STK < k_end) fetch(k0 + STK);
            else {
#pragma unroll
                for (int u = 0; u < U; ++u) ra[u] = rb[u] = 0.0f;
            }
        }
#pragma unroll
        for (int k = 0; k < STK; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&sm.As[grp][k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&sm.Bs[grp][k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
        }
        __syncthreads();
    }
    // partial tiles of groups 1.. -> their own (now idle) A stage, as [32][32+4]; group 0 adds them in group order
    if (KG > 1) {
        if (grp > 0) {
#pragma unroll
            for (int x = 0; x < 4; ++x)
                *reinterpret_cast<float4 *>(&sm.As[grp][ty * 4 + x][tx * 4]) = make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]);
        }
        __syncthreads();
        if (grp == 0) {
#pragma unroll
            for (int q = 1; q < KG; ++q)
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const float4 p = *reinterpret_cast<const float4 *>(&sm.As[q][ty * 4 + x][tx * 4]);
                    acc[x][0] = __fadd_rn(acc[x][0], p.x);
                    acc[x][1] = __fadd_rn(acc[x][1], p.y);
                    acc[x][2] = __fadd_rn(acc[x][2], p.z);
                    acc[x][3] = __fadd_rn(acc[x][3], p.w);
                }
        }
    }
    if (grp == 0) {
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) epilogue_store(g, i0 + ty * 4 + x, j0 + tx * 4 + y, acc[x][y]);
    }
}

// ---- whole-K GEMM: the learner's layers have K <= 512, so a 32 x 32 output tile's ENTIRE operand panels fit in shared
// memory (2 x 32 x K floats: 64 KB at K = 256).  The kernel above still walks K in steps of 32 (fetch -> park -> barrier ->
// multiply -> barrier per step): measured 8-9 us per layer, every warp spending most of its issue slots on address
// arithmetic, staging stores and barrier waits.  Here the panels are copied ONCE with cp.async (16-byte copies straight
// from L2 into shared memory, no registers, all in flight together), one barrier, then the four K-groups of 64 threads
// multiply their quarter of K out of shared memory with no further synchronisation (per 4 k: 8 LDS.128 + 64 FFMA per
// thread), and the partial tiles are added in group order like above.  One global round trip instead of K / 32.
// Operand layouts (C[i, j] = sum_k A(i, k) B(j, k)):  k-fast -> panel [32][Kp] (Kp = K padded to 4, + 4 floats of row
// padding against bank conflicts);  k-slow -> panel [Kp][36].  Anything a 16-byte copy cannot express (row starts that
// are not 16-byte aligned, ragged edges, the virtual column of ones) is filled by plain loads.
constexpr int WK_MAX_K = 512;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Fill one operand panel.  kKfast: panel[r][k] (row stride ld = Kp + 4), else panel[k][r] (row stride 36).
// rows = tile rows that exist (< 32 at the edge), r0 = first row, stride_r / stride_k = element strides in global memory.
template <bool kKfast>
__device__ __forceinline__ void wk_fill(float *panel, const float *__restrict__ src, int64_t stride_r, int64_t stride_k,
                                        int r0, int n_rows, int K, int Kp, int ones_row, int tid, int nthreads) {
    const int ld = kKfast ? Kp + 4 : 36;
    if (kKfast) {
        // chunks of 4 consecutive k of one row
        const int chunks_per_row = Kp >> 2;
        const bool fast = stride_k == 1 && (stride_r & 3) == 0 && ((uintptr_t)src & 15) == 0;
        for (int c = tid; c < 32 * chunks_per_row; c += nthreads) {
            const int r = c / chunks_per_row, k = (c - r * chunks_per_row) << 2;
            float *dst = panel + r * ld + k;
            const int gr = r0 + r;
            if (gr == ones_row) {
                dst[0] = dst[1] = dst[2] = dst[3] = 1.0f;      // (k beyond K multiplies zeros of the other panel)
            } else if (gr < n_rows && fast && k + 4 <= K) {
                cp_async16(dst, src + (int64_t)gr * stride_r + k);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    dst[q] = (gr < n_rows && k + q < K) ? __ldcg(src + (int64_t)gr * stride_r + (int64_t)(k + q) * stride_k) : 0.0f;
            }
        }
    } else {
        // chunks of 4 consecutive rows at one k
        const bool fast = stride_r == 1 && (stride_k & 3) == 0 && ((uintptr_t)src & 15) == 0 && (r0 & 3) == 0;
        for (int c = tid; c < Kp * 8; c += nthreads) {
            const int k = c >> 3, r = (c & 7) << 2;
            float *dst = panel + k * ld + r;
            const int gr = r0 + r;
            if (k < K && fast && gr + 4 <= n_rows && !(ones_row >= gr && ones_row < gr + 4)) {
                cp_async16(dst, src + (int64_t)k * stride_k + gr);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float v = 0.0f;
                    if (gr + q == ones_row) v = 1.0f;
                    else if (k < K && gr + q < n_rows) v = __ldcg(src + (int64_t)(gr + q) * stride_r + (int64_t)k * stride_k);
                    dst[q] = v;
                }
            }
        }
    }
}

template <int KG, bool kAKfast, bool kBKfast>     // KG K-groups of 64 threads: 4 (K < 128) or 8
__global__ void __launch_bounds__(64 * KG) gemm_wk_kernel(GemmArgs g, int Kp) {
    MB_DYNAMIC_SMEM(dyn);
    const int lda = kAKfast ? Kp + 4 : 36, ldb = kBKfast ? Kp + 4 : 36;
    float *As = reinterpret_cast<float *>(dyn);
    const int a_floats = kAKfast ? 32 * lda : Kp * lda;
    constexpr int kRed = (KG - 1) * 64 * 16;                               // reduction scratch: (KG - 1) groups x 64 threads x 16
    float *Bs = As + (a_floats < kRed ? kRed : a_floats);                  // (the A region also hosts it)
    const int tid = threadIdx.x, grp = tid >> 6, lt = tid & 63;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    wk_fill<kAKfast>(As, g.A, g.a_i, g.a_k, i0, g.M, g.K, Kp, -1, tid, 64 * KG);
    // B's row index is the output column j; the virtual column of ones (db) is row `ones_col` of B
    wk_fill<kBKfast>(Bs, g.B, g.b_j, g.b_k, j0, g.ones_col >= 0 ? g.N - 1 : g.N, g.K, Kp, g.ones_col, tid, 64 * KG);
    cp_async_wait_all();
    __syncthreads();
    const int ty = lt >> 3, tx = lt & 7;          // 8 x 8 threads per K-group, 4 x 4 outputs each
    const int kq = ((Kp >> 2) + KG - 1) / KG * 4;  // k's per group (multiple of 4)
    const int k_begin = grp * kq, k_end = min(Kp, k_begin + kq);
    float acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = 0.0f;
    // software-pipelined: the 8 LDS.128 of the next 4 k are in flight while the 64 FFMA of the current ones issue
    // (two warps per scheduler cannot hide the shared-memory latency by themselves)
    float a[2][4][4], b[2][4][4];                   // [buffer][row / col][kk]
    auto load = [&](int buf, int k) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (kAKfast) {   // rows ty + 8q: consecutive rows across lanes, row stride Kp + 4 floats -> conflict-free LDS.128
                const float4 v = *reinterpret_cast<const float4 *>(As + (ty + 8 * q) * lda + k);
                a[buf][q][0] = v.x; a[buf][q][1] = v.y; a[buf][q][2] = v.z; a[buf][q][3] = v.w;
            } else {
                const float4 v = *reinterpret_cast<const float4 *>(As + (k + q) * lda + ty * 4);
                a[buf][0][q] = v.x; a[buf][1][q] = v.y; a[buf][2][q] = v.z; a[buf][3][q] = v.w;
            }
            if (kBKfast) {   // columns tx + 8q (same reason)
                const float4 v = *reinterpret_cast<const float4 *>(Bs + (tx + 8 * q) * ldb + k);
                b[buf][q][0] = v.x; b[buf][q][1] = v.y; b[buf][q][2] = v.z; b[buf][q][3] = v.w;
            } else {
                const float4 v = *reinterpret_cast<const float4 *>(Bs + (k + q) * ldb + tx * 4);
                b[buf][0][q] = v.x; b[buf][1][q] = v.y; b[buf][2][q] = v.z; b[buf][3][q] = v.w;
            }
        }
    };
    auto mult = [&](int buf) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[buf][x][kk], b[buf][y][kk], acc[x][y]);
    };
    if (k_begin < k_end) load(0, k_begin);
    for (int k = k_begin; k < k_end; k += 8) {
        if (k + 4 < k_end) load(1, k + 4);
        mult(0);
        if (k + 4 < k_end) {
            if (k + 8 < k_end) load(0, k + 8);
            mult(1);
        }
    }
    // partial tiles of groups 1..3 -> the (now idle) A panel as [32][36]; group 0 adds them in group order
    __syncthreads();
    float *red = As;                                // (KG - 1) groups x 64 threads x 16 floats
    // (the scratch is indexed by thread, not by output position: slot [lt][x] holds that thread's 4 values of row x)
    if (grp > 0) {
#pragma unroll
        for (int x = 0; x < 4; ++x)
            *reinterpret_cast<float4 *>(red + ((grp - 1) * 64 + lt) * 16 + x * 4) =
                make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]);
    }
    __syncthreads();
    if (grp == 0) {
#pragma unroll
        for (int q = 0; q < KG - 1; ++q)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const float4 p = *reinterpret_cast<const float4 *>(red + (q * 64 + lt) * 16 + x * 4);
                acc[x][0] = __fadd_rn(acc[x][0], p.x);
                acc[x][1] = __fadd_rn(acc[x][1], p.y);
                acc[x][2] = __fadd_rn(acc[x][2], p.z);
                acc[x][3] = __fadd_rn(acc[x][3], p.w);
            }
        // Epilogue.  epilogue_store() per element would issue up to 32 dependent-looking global loads (bias / mask)
        // interleaved with the stores -- the compiler cannot hoist them across stores it must assume alias -- i.e. a
        // chain of L2 round trips longer than the multiply itself.  All loads first, then the arithmetic, then the stores.
        int rows[4], cols[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) rows[x] = i0 + (kAKfast ? ty + 8 * x : ty * 4 + x);
#pragma unroll
        for (int y = 0; y < 4; ++y) cols[y] = j0 + (kBKfast ? tx + 8 * y : tx * 4 + y);
        float bias_v[4], mask_v[4][4];
#pragma unroll
        for (int y = 0; y < 4; ++y) bias_v[y] = (g.bias && cols[y] < g.N) ? __ldg(g.bias + cols[y]) : 0.0f;
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y)
                mask_v[x][y] = (g.mask && rows[x] < g.M && cols[y] < g.N) ? __ldg(g.mask + (int64_t)rows[x] * g.c_i + cols[y]) : 1.0f;
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                float v = acc[x][y];
                if (g.bias) v = __fadd_rn(v, bias_v[y]);
                if (g.relu) v = fmaxf(v, 0.0f);
                if (g.mask) v = mask_v[x][y] > 0.0f ? v : 0.0f;
                acc[x][y] = v;
            }
        const bool vec = !kBKfast && g.ones_col < 0 && (g.c_i & 3) == 0 && (((uintptr_t)g.C) & 15) == 0 && cols[3] < g.N;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            if (rows[x] >= g.M) continue;
            if (vec) {
                *reinterpret_cast<float4 *>(g.C + (int64_t)rows[x] * g.c_i + cols[0]) = make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]);
            } else {
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    if (cols[y] >= g.N) continue;
                    if (g.ones_col >= 0 && cols[y] == g.ones_col) g.extra[rows[x]] = acc[x][y];
                    else g.C[(int64_t)rows[x] * g.c_i + cols[y]] = acc[x][y];
                }
            }
        }
    }
}

// ---- narrow outputs (N <= 8: the last layer of a Q-net or critic): one WARP per row -------------------------------------
// A 32 x 32 tile wastes 26 of its 32 columns on a 6-wide output and the launch is 16 CTAs walking K = 256 each (measured
// 7.7 us, as long as the 256-wide layer).  Here the N weight rows live in shared memory, every warp takes one input row:
// lane l multiplies elements 4l.. of the row (float4 loads, stride 128) against all N weight rows, the N partial sums are
// reduced with shuffles in a fixed order, lanes 0..N-1 add the bias and store.  64 CTAs for 512 rows, one L2 round trip.
constexpr int NARROW_MAX_N = 8;

__global__ void __launch_bounds__(256) gemm_narrow_kernel(GemmArgs g) {
    MB_DYNAMIC_SMEM(dyn);
    float *ws = reinterpret_cast<float *>(dyn);                 // [N][K]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int K = g.K, N = g.N;
    for (int c = tid; c < (N * K) >> 2; c += 256) {
        const int j = (c << 2) / K, k = (c << 2) - j * K;
        *reinterpret_cast<float4 *>(ws + j * K + k) = __ldg(reinterpret_cast<const float4 *>(g.B + (int64_t)j * g.b_j + k));
    }
    __syncthreads();
    for (int row = blockIdx.x * 8 + wid; row < g.M; row += gridDim.x * 8) {
        float acc[NARROW_MAX_N];
#pragma unroll
        for (int j = 0; j < NARROW_MAX_N; ++j) acc[j] = 0.0f;
        const float *x = g.A + (int64_t)row * g.a_i;
        for (int k = lane * 4; k < K; k += 128) {
            const float4 xv = __ldcg(reinterpret_cast<const float4 *>(x + k));
#pragma unroll
            for (int j = 0; j < NARROW_MAX_N; ++j)
                if (j < N) {
                    const float4 wv = *reinterpret_cast<const float4 *>(ws + j * K + k);
                    acc[j] = fmaf(xv.w, wv.w, fmaf(xv.z, wv.z, fmaf(xv.y, wv.y, fmaf(xv.x, wv.x, acc[j]))));
                }
        }
#pragma unroll
        for (int j = 0; j < NARROW_MAX_N; ++j)
            if (j < N) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[j] = __fadd_rn(acc[j], __shfl_xor_sync(0xffffffffu, acc[j], o));
            }
        float out = 0.0f;                                        // lane j keeps column j
#pragma unroll
        for (int j = 0; j < NARROW_MAX_N; ++j)
            if (lane == j) out = acc[j];
        if (lane < N) {
            if (g.bias) out = __fadd_rn(out, __ldg(g.bias + lane));
            if (g.relu) out = fmaxf(out, 0.0f);
            g.C[(int64_t)row * g.c_i + lane] = out;
        }
    }
}

bool launch_narrow(const GemmArgs &g, cudaStream_t s) {
    if (g.N > NARROW_MAX_N || g.N < 1 || g.K < 4 || (g.K & 3) || g.K > 2048 || g.a_k != 1 || g.b_k != 1 || (g.a_i & 3) || (g.b_j & 3) ||
        g.mask || g.ones_col >= 0 || (((uintptr_t)g.A | (uintptr_t)g.B) & 15) || (size_t)g.N * g.K * sizeof(float) > 48 * 1024)
        return false;
    int blocks = (g.M + 7) / 8;
    if (blocks > 4 * mb::kNumSMs) blocks = 4 * mb::kNumSMs;
    MB_PREFER_MAX_SMEM(gemm_narrow_kernel);
    gemm_narrow_kernel<<<blocks, 256, sizeof(float) * (size_t)g.N * g.K, s>>>(g);
    return true;
}

// MB_GEMM=wk (default) | small (32 x 32 tiles, K in steps of 32) | legacy (64 x 64 tiles, grid split-K): A/B timing
int gemm_mode() {
    const char *e = getenv("MB_GEMM");
    if (e && e[0] == 's') return 1;
    if (e && e[0] == 'l') return 2;
    const char *l = getenv("MB_GEMM_LEGACY");
    return (l && l[0] == '1') ? 2 : 0;
}

template <int KG, bool kA, bool kB>
bool launch_wk_groups(GemmArgs g, cudaStream_t s, int Kp) {
    const size_t a_floats = kA ? (size_t)32 * (Kp + 4) : (size_t)Kp * 36;
    const size_t b_floats = kB ? (size_t)32 * (Kp + 4) : (size_t)Kp * 36;
    const size_t red = (size_t)(KG - 1) * 64 * 16;           // the reduction scratch lives in the A panel
    const size_t dyn = sizeof(float) * ((a_floats < red ? red : a_floats) + b_floats);
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(gemm_wk_kernel<KG, kA, kB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr_done = true;
    }
    if (dyn > 200 * 1024) return false;
    const dim3 grid((g.N + 31) / 32, (g.M + 31) / 32, 1);
    MB_PREFER_MAX_SMEM(gemm_wk_kernel<KG, kA, kB>);
    gemm_wk_kernel<KG, kA, kB><<<grid, 64 * KG, dyn, s>>>(g, Kp);
    return true;
}

template <bool kA, bool kB>
bool launch_wk(GemmArgs g, cudaStream_t s) {
    const int tiles64 = ((g.N + TN - 1) / TN) * ((g.M + TM - 1) / TM);
    if (gemm_mode() != 0 || tiles64 >= 2 * mb::kNumSMs || g.K > WK_MAX_K || g.K < 1) return false;
    const int Kp = (g.K + 3) / 4 * 4;
    // more K-groups = more warps per scheduler to hide the shared-memory latency of the multiply (each group walks K / KG)
    return Kp >= 128 ? launch_wk_groups<8, kA, kB>(g, s, Kp) : launch_wk_groups<4, kA, kB>(g, s, Kp);
}

bool small_gemm_enabled() { return gemm_mode() != 2; }

template <bool kA, bool kB>
bool launch_small(GemmArgs g, cudaStream_t s) {
    const int tiles64 = ((g.N + TN - 1) / TN) * ((g.M + TM - 1) / TM);
    // (reductions longer than 1024 keep the grid-level split-K of the 64 x 64 kernel: more, shorter partial sums)
    if (!small_gemm_enabled() || tiles64 >= 2 * mb::kNumSMs || g.K > 1024) return false;
    const dim3 grid((g.N + ST - 1) / ST, (g.M + ST - 1) / ST, 1);
    const int ksteps = (g.K + STK - 1) / STK;
    if (ksteps >= 4) {
        g.k_per_split = (ksteps + 3) / 4 * STK;
        gemm_small_kernel<4, kA, kB><<<grid, SG * 4, 0, s>>>(g);
    } else if (ksteps >= 2) {
        g.k_per_split = (ksteps + 1) / 2 * STK;
        gemm_small_kernel<2, kA, kB><<<grid, SG * 2, 0, s>>>(g);
    } else {
        g.k_per_split = STK;
        gemm_small_kernel<1, kA, kB><<<grid, SG, 0, s>>>(g);
    }
    return true;
}

int pick_splits(int tiles, int K) {
    int s = 1;
    while (tiles * s * 2 <= mb::kNumSMs + mb::kNumSMs / 4 && (K / (s * 2)) >= 2 * TK && s < 32) s *= 2;
    return s;
}

template <bool kA, bool kB>
int launch(GemmArgs g, float *workspace, int64_t workspace_floats, int *counters, int n_counters, cudaStream_t s) {
    if ((kA && kB && gemm_mode() == 0 && launch_narrow(g, s)) || launch_wk<kA, kB>(g, s) || launch_small<kA, kB>(g, s)) {
        MB_LAUNCHED(1);
        return MB_OK;
    }
    const int tx = (g.N + TN - 1) / TN, ty = (g.M + TM - 1) / TM;
    int splits = pick_splits(tx * ty, g.K);
    if (splits > 1 && (!workspace || !counters || (int64_t)splits * g.M * g.N > workspace_floats || tx * ty > n_counters))
        splits = 1;
    g.k_per_split = ((g.K + splits - 1) / splits + TK - 1) / TK * TK;
    splits = (g.K + g.k_per_split - 1) / g.k_per_split;
    g.partial = workspace;
    g.tile_counter = counters;
    gemm_kernel<kA, kB><<<dim3(tx, ty, splits), NT, 0, s>>>(g);
    MB_LAUNCHED(1);
    return MB_OK;
}

// ---------------------------------------------------------------------------------------------
// Persistent "program" executor: the whole double-DQN value math of one learner update -- three forward
// passes (Q(s), Q(s'), Q_target(s')), the TD target / |delta| / loss / dloss/dq row math and the full
// backward pass -- in ONE cooperative launch.  A program is a list of ops grouped in phases; within a
// phase the CTAs of the grid pull work items (64x64 tile x K-split, or 256 TD rows) round-robin, phases
// are separated by grid.sync().  Replaces ~75 dependent library launches (or 16 of the stand-alone GEMM
// launches above) whose cost is launch gaps and single-wave latency, not arithmetic.
// ---------------------------------------------------------------------------------------------
enum OpKind { OP_GEMM_KK = 0, OP_GEMM_KJ = 1, OP_GEMM_IJ = 2, OP_TD_DQN = 3 };

struct ProgOp {
    GemmArgs g;
    TdDqnArgs td;
    int kind, phase, tiles_x, tiles_y, splits, items;
};

__global__ void __launch_bounds__(NT, 1) mlp_program_kernel(const ProgOp *__restrict__ ops, int n_ops, int n_phases) {
    cg::grid_group grid = cg::this_grid();
    __shared__ GemmSmem sm;
    __shared__ ProgOp s_op;
    int op0 = 0;
    for (int ph = 0; ph < n_phases; ++ph) {
        int op1 = op0, total = 0;
        while (op1 < n_ops && ops[op1].phase == ph) total += ops[op1++].items;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            int o = op0, local = item;
            while (local >= ops[o].items) local -= ops[o++].items;
            __syncthreads();  // previous item's readers of s_op are done
            for (int e = threadIdx.x; e < (int)(sizeof(ProgOp) / 4); e += NT)
                reinterpret_cast<int *>(&s_op)[e] = reinterpret_cast<const int *>(ops + o)[e];
            __syncthreads();
            if (s_op.kind == OP_TD_DQN) {
                const int64_t b = (int64_t)local * NT + threadIdx.x;
                if (b < s_op.td.batch) td_dqn_row(s_op.td, b);
            } else {
                const int tiles = s_op.tiles_x * s_op.tiles_y;
                const int split = local / tiles, tl = local % tiles;
                const int tyi = tl / s_op.tiles_x, txi = tl % s_op.tiles_x;
                if (s_op.kind == OP_GEMM_KK) gemm_item<true, true>(s_op.g, sm, txi, tyi, s_op.tiles_x, split, s_op.splits);
                else if (s_op.kind == OP_GEMM_KJ) gemm_item<true, false>(s_op.g, sm, txi, tyi, s_op.tiles_x, split, s_op.splits);
                else gemm_item<false, false>(s_op.g, sm, txi, tyi, s_op.tiles_x, split, s_op.splits);
            }
        }
        op0 = op1;
        if (ph + 1 < n_phases) grid.sync();
    }
}


// ---------------------------------------------------------------------------------------------
// Row-tile program: the same value math without any grid-wide barrier.
// Forward of all three passes, the TD row math and the backward-input chain of ONE batch row depend on that
// row only, so a CTA that owns RB consecutive batch rows can run all of it back to back with the activations
// in shared memory; the weights (0.35 MB for 84-256-256-6) stream from L2 through L1 straight into registers
// (forward from a k-major copy, refreshed by a small kernel in front, so that a thread's four output columns
// are one 16-byte load; backward-input from the parameter itself, which is already k-major for that product).
// The online net runs Q(s) and Q(s') as one 2*RB-row pass, so every weight load feeds 16 rows.  What does
// need all rows -- dW = dz^T x and db -- follows as ONE launch of independent tile x K-split items (the
// executor above with a single phase: no grid.sync).  Three launches per update instead of one cooperative
// launch with six grid barriers (or ~35 library launches).
// ---------------------------------------------------------------------------------------------
#include "rows_program.cuh"

struct TransposeList {
    const float *src[2 * RL];
    float *dst[2 * RL];
    int rows[2 * RL], cols[2 * RL];   // src is [rows, cols] row-major; dst[c * rows + r] = src[r * cols + c]
    int64_t start[2 * RL + 1];
    int n;
};

__global__ void __launch_bounds__(256) transpose_multi_kernel(TransposeList l) {
    const int64_t total = l.start[l.n];
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int t = 0;
        while (e >= l.start[t + 1]) ++t;
        const int64_t o = e - l.start[t];                 // index into dst: coalesced writes, strided (L2-resident) reads
        const int c = (int)(o / l.rows[t]), r = (int)(o % l.rows[t]);
        l.dst[t][o] = __ldg(l.src[t] + (int64_t)r * l.cols[t] + c);
    }
}

// fused multi-tensor Adam (torch.optim.Adam semantics, amsgrad = False, maximize = False)
struct AdamList {
    float *p[MB_MAX_SOFT_TENSORS];
    const float *g[MB_MAX_SOFT_TENSORS];
    float *m[MB_MAX_SOFT_TENSORS];
    float *v[MB_MAX_SOFT_TENSORS];
    int64_t start[MB_MAX_SOFT_TENSORS + 1];
    int n;
};

__global__ void __launch_bounds__(256) adam_kernel(AdamList l, float lr, float beta1, float beta2, float eps,
                                                   float weight_decay, float *step_count) {
    const int64_t total = l.start[l.n];
    const float t = *step_count + 1.0f;
    const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
    const float step_size = lr / bc1, bc2_sqrt = sqrtf(bc2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int k = 0;
        while (i >= l.start[k + 1]) ++k;
        const int64_t j = i - l.start[k];
        float grad = l.g[k][j];
        const float p = l.p[k][j];
        if (weight_decay != 0.0f) grad = fmaf(p, weight_decay, grad);
        const float m = l.m[k][j] + (grad - l.m[k][j]) * (1.0f - beta1);          // lerp, like torch
        const float v = l.v[k][j] * beta2 + (1.0f - beta2) * grad * grad;
        l.m[k][j] = m;
        l.v[k][j] = v;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        l.p[k][j] = p - step_size * (m / denom);
    }
}

__global__ void adam_step_inc(float *step_count) { *step_count += 1.0f; }

#include "optim_device.cuh"   // adam_soft_kernel: the Adam step and the soft target update in one launch

}  // namespace

extern "C" {

int mb_mlp_linear_forward(const float *d_x, const float *d_w, const float *d_b, float *d_y, int M, int N, int K,
                          int relu, float *d_workspace, int64_t workspace_floats, int *d_counters, int n_counters,
                          void *stream) {
    if (!d_x || !d_w || !d_y || M <= 0 || N <= 0 || K <= 0) return MB_ERR_INVALID;
    GemmArgs g{};
    g.A = d_x; g.a_i = K; g.a_k = 1;
    g.B = d_w; g.b_j = K; g.b_k = 1;
    g.C = d_y; g.c_i = N;
    g.M = M; g.N = N; g.K = K;
    g.ones_col = -1; g.bias = d_b; g.relu = relu;
    return launch<true, true>(g, d_workspace, workspace_floats, d_counters, n_counters, (cudaStream_t)stream);
}

int mb_mlp_linear_backward_input(const float *d_dy, const float *d_w, const float *d_act_mask, float *d_dx, int M,
                                 int N, int K, float *d_workspace, int64_t workspace_floats, int *d_counters,
                                 int n_counters, void *stream) {
    if (!d_dy || !d_w || !d_dx || M <= 0 || N <= 0 || K <= 0) return MB_ERR_INVALID;
    GemmArgs g{};
    g.A = d_dy; g.a_i = N; g.a_k = 1;         // A(m, n)
    g.B = d_w; g.b_j = 1; g.b_k = K;          // B(kk, n) = w[n, kk]
    g.C = d_dx; g.c_i = K;
    g.M = M; g.N = K; g.K = N;
    g.ones_col = -1; g.mask = d_act_mask;
    return launch<true, false>(g, d_workspace, workspace_floats, d_counters, n_counters, (cudaStream_t)stream);
}

int mb_mlp_linear_backward_weight(const float *d_dy, const float *d_x, float *d_dw, float *d_db, int M, int N, int K,
                                  float *d_workspace, int64_t workspace_floats, int *d_counters, int n_counters,
                                  void *stream) {
    if (!d_dy || !d_x || !d_dw || M <= 0 || N <= 0 || K <= 0) return MB_ERR_INVALID;
    GemmArgs g{};
    g.A = d_dy; g.a_i = 1; g.a_k = N;         // A(n, m) = dy[m, n]
    g.B = d_x; g.b_j = 1; g.b_k = K;          // B(kk, m) = x[m, kk]
    g.C = d_dw; g.c_i = K;
    g.M = N; g.N = d_db ? K + 1 : K; g.K = M;
    g.ones_col = d_db ? K : -1; g.extra = d_db;
    return launch<false, false>(g, d_workspace, workspace_floats, d_counters, n_counters, (cudaStream_t)stream);
}

struct mb_mlp_program {
    int device;
    ProgOp *d_ops;
    int n_ops, n_phases, grid;
    std::vector<void *> owned;   // device allocations (activations, partials, counters)
    int rows;                    // row-tile variant: transpose refresh + dqn_rows_kernel + single-phase dW program
    int rows_cpt;                // its thread layout: 4 or 2 output columns per thread (MB_ROWS_CPT, default 4)
    RowsArgs ra;
    TransposeList tl;
    size_t rows_smem;
    int rows_grid;
};

namespace {

struct Builder {
    std::vector<ProgOp> ops;
    std::vector<void *> *owned;
    int grid;
    bool failed = false;

    void *alloc(size_t bytes, bool zero) {
        void *p = nullptr;
        if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) { failed = true; cudaGetLastError(); return nullptr; }
        if (zero) cudaMemset(p, 0, bytes ? bytes : 16);
        owned->push_back(p);
        return p;
    }
    void add_gemm(int phase, int kind, GemmArgs g) {
        ProgOp op{};
        op.kind = kind; op.phase = phase;
        op.tiles_x = (g.N + TN - 1) / TN; op.tiles_y = (g.M + TM - 1) / TM;
        const int tiles = op.tiles_x * op.tiles_y;
        int splits = 1;   // enough K-splits that one op offers up to ~1 item per CTA, each with >= 2 K-steps
        while (tiles * splits * 2 <= grid + grid / 4 && g.K / (splits * 2) >= 2 * TK && splits < 32) splits *= 2;
        g.k_per_split = ((g.K + splits - 1) / splits + TK - 1) / TK * TK;
        splits = (g.K + g.k_per_split - 1) / g.k_per_split;
        if (splits > 1) {
            g.partial = (float *)alloc(sizeof(float) * (size_t)splits * g.M * g.N, false);
            g.tile_counter = (int *)alloc(sizeof(int) * (size_t)tiles, true);
        }
        op.g = g; op.splits = splits; op.items = tiles * splits;
        ops.push_back(op);
    }
};

}  // namespace

int mb_dqn_mlp_program_create(int device, int n_layers, const int *sizes, int batch, float *const *d_params,
                              float *const *d_target_params, float *const *d_grads, const float *d_state,
                              const float *d_next_state, const int64_t *d_action, const float *d_reward,
                              const uint8_t *d_terminal, const double *d_is_weight, float discount, int mode,
                              float *d_y, float *d_action_value, float *d_abs_error, double *d_loss_elem,
                              mb_mlp_program **out) {
    if (!out || n_layers < 1 || n_layers > 8 || batch < 1 || !sizes || !d_params || !d_target_params || !d_grads)
        return MB_ERR_INVALID;
    if (mb_device_count() <= device || device < 0) return MB_ERR_NO_DEVICE;
    mb::DeviceGuard guard(device);
    mb_mlp_program *P = new (std::nothrow) mb_mlp_program();
    if (!P) return MB_ERR_NOMEM;
    P->device = device;
    int per_sm = 0, sms = mb::kNumSMs;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mlp_program_kernel, NT, 0);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    P->grid = sms * (per_sm >= 2 ? 2 : 1);
    Builder B;
    B.owned = &P->owned;
    B.grid = P->grid;
    const int L = n_layers, A = sizes[L];
    const int64_t Bn = batch;
    // activations: pass 0 = Q(s) online, 1 = Q(s') online, 2 = Q_target(s')
    float *act[3][9];
    const float *in[3] = {d_state, d_next_state, d_next_state};
    for (int p = 0; p < 3; ++p)
        for (int l = 0; l < L; ++l) act[p][l] = (float *)B.alloc(sizeof(float) * Bn * sizes[l + 1], false);
    float *dz[9];
    for (int l = 0; l < L; ++l) dz[l] = (float *)B.alloc(sizeof(float) * Bn * sizes[l + 1], false);  // grad wrt layer-l output
    int phase = 0;
    for (int l = 0; l < L; ++l, ++phase)
        for (int p = 0; p < 3; ++p) {
            float *const *prm = (p == 2) ? d_target_params : d_params;
            GemmArgs g{};
            g.A = l == 0 ? in[p] : act[p][l - 1]; g.a_i = sizes[l]; g.a_k = 1;
            g.B = prm[2 * l]; g.b_j = sizes[l]; g.b_k = 1;
            g.C = act[p][l]; g.c_i = sizes[l + 1];
            g.M = batch; g.N = sizes[l + 1]; g.K = sizes[l];
            g.ones_col = -1; g.bias = prm[2 * l + 1]; g.relu = l < L - 1;
            B.add_gemm(phase, OP_GEMM_KK, g);
        }
    {   // TD row math: q = act[0][L-1], qn = act[1][L-1], qtn = act[2][L-1]; dloss/dq -> dz[L-1]
        ProgOp op{};
        op.kind = OP_TD_DQN; op.phase = phase++;
        op.td = TdDqnArgs{act[0][L - 1], act[1][L - 1], act[2][L - 1], Bn, A, d_action, d_reward, d_terminal, d_is_weight,
                          discount, mode, d_y, d_action_value, d_abs_error, d_loss_elem, dz[L - 1]};
        op.items = (batch + NT - 1) / NT;
        B.ops.push_back(op);
    }
    for (int l = L - 1; l >= 0; --l, ++phase) {
        const float *x = l == 0 ? d_state : act[0][l - 1];
        GemmArgs w{};   // dW_l[N, K] = dz_l^T x ; db_l = column sums of dz_l
        w.A = dz[l]; w.a_i = 1; w.a_k = sizes[l + 1];
        w.B = x; w.b_j = 1; w.b_k = sizes[l];
        w.C = d_grads[2 * l]; w.c_i = sizes[l];
        w.M = sizes[l + 1]; w.N = sizes[l] + 1; w.K = batch;
        w.ones_col = sizes[l]; w.extra = d_grads[2 * l + 1];
        B.add_gemm(phase, OP_GEMM_IJ, w);
        if (l > 0) {    // dz_{l-1}[M, K] = (dz_l W_l) * (act_{l-1} > 0)
            GemmArgs d{};
            d.A = dz[l]; d.a_i = sizes[l + 1]; d.a_k = 1;
            d.B = d_params[2 * l]; d.b_j = 1; d.b_k = sizes[l];
            d.C = dz[l - 1]; d.c_i = sizes[l];
            d.M = batch; d.N = sizes[l]; d.K = sizes[l + 1];
            d.ones_col = -1; d.mask = act[0][l - 1];
            B.add_gemm(phase, OP_GEMM_KJ, d);
        }
    }
    if (!B.failed) P->d_ops = (ProgOp *)B.alloc(sizeof(ProgOp) * B.ops.size(), false);
    if (B.failed || cudaMemcpy(P->d_ops, B.ops.data(), sizeof(ProgOp) * B.ops.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        mb_mlp_program_destroy(P);
        return MB_ERR_NOMEM;
    }
    P->n_ops = (int)B.ops.size();
    P->n_phases = phase;
    int max_items = 1;
    for (int ph = 0; ph < phase; ++ph) {
        int tot = 0;
        for (const ProgOp &op : B.ops) if (op.phase == ph) tot += op.items;
        if (tot > max_items) max_items = tot;
    }
    if (P->grid > max_items) P->grid = max_items;
    MB_CUDA(cudaDeviceSynchronize());
    *out = P;
    return MB_OK;
}

int mb_dqn_mlp_rows_create(int device, int n_layers, const int *sizes, int batch, float *const *d_params,
                           float *const *d_target_params, float *const *d_grads, const float *d_state,
                           const float *d_next_state, const int64_t *d_action, const float *d_reward,
                           const uint8_t *d_terminal, const double *d_is_weight, float discount, int mode,
                           float *d_y, float *d_action_value, float *d_abs_error, double *d_loss_elem,
                           mb_mlp_program **out) {
    if (!out || n_layers < 1 || n_layers > RL || batch < 1 || !sizes || !d_params || !d_target_params || !d_grads)
        return MB_ERR_INVALID;
    for (int l = 1; l <= n_layers; ++l)
        if (sizes[l] < 1 || sizes[l] > RW) return MB_ERR_INVALID;     // one CTA sweeps a whole layer output
    if (sizes[0] < 1) return MB_ERR_INVALID;
    if (mb_device_count() <= device || device < 0) return MB_ERR_NO_DEVICE;
    mb::DeviceGuard guard(device);
    mb_mlp_program *P = new (std::nothrow) mb_mlp_program();
    if (!P) return MB_ERR_NOMEM;
    P->device = device;
    P->rows = 1;
    int sms = mb::kNumSMs;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    Builder B;
    B.owned = &P->owned;
    B.grid = sms;
    const int L = n_layers, A = sizes[L];
    const int64_t Bn = batch;
    RowsArgs &ra = P->ra;
    ra = RowsArgs{};
    ra.L = L;
    ra.batch = batch;
    for (int l = 0; l <= L; ++l) ra.sizes[l] = sizes[l];
    ra.state = d_state;
    ra.next_state = d_next_state;
    TransposeList &tl = P->tl;
    tl = TransposeList{};
    tl.start[0] = 0;
    for (int l = 0; l < L; ++l) {
        const size_t wn = (size_t)sizes[l] * sizes[l + 1];
        float *wt = (float *)B.alloc(sizeof(float) * wn, true), *twt = (float *)B.alloc(sizeof(float) * wn, true);
        ra.W[l] = d_params[2 * l];
        ra.b[l] = d_params[2 * l + 1];
        ra.tb[l] = d_target_params[2 * l + 1];
        ra.Wt[l] = wt;
        ra.tWt[l] = twt;
        ra.h[l] = (float *)B.alloc(sizeof(float) * Bn * sizes[l + 1], true);
        ra.dz[l] = (float *)B.alloc(sizeof(float) * Bn * sizes[l + 1], true);
        for (int net = 0; net < 2; ++net) {
            const int t = tl.n++;
            tl.src[t] = net == 0 ? d_params[2 * l] : d_target_params[2 * l];
            tl.dst[t] = net == 0 ? wt : twt;
            tl.rows[t] = sizes[l + 1];      // parameter is [out, in]
            tl.cols[t] = sizes[l];
            tl.start[t + 1] = tl.start[t] + (int64_t)wn;
        }
    }
    ra.qn = (float *)B.alloc(sizeof(float) * Bn * A, true);
    ra.qtn = (float *)B.alloc(sizeof(float) * Bn * A, true);
    ra.td = TdDqnArgs{ra.h[L - 1], ra.qn, ra.qtn, Bn, A, d_action, d_reward, d_terminal, d_is_weight,
                      discount, mode, d_y, d_action_value, d_abs_error, d_loss_elem, ra.dz[L - 1]};
    // dW_l[N, K] = dz_l^T x_l ; db_l = column sums of dz_l -- independent items, one phase
    for (int l = 0; l < L; ++l) {
        GemmArgs w{};
        w.A = ra.dz[l]; w.a_i = 1; w.a_k = sizes[l + 1];
        w.B = l == 0 ? d_state : ra.h[l - 1]; w.b_j = 1; w.b_k = sizes[l];
        w.C = d_grads[2 * l]; w.c_i = sizes[l];
        w.M = sizes[l + 1]; w.N = sizes[l] + 1; w.K = batch;
        w.ones_col = sizes[l]; w.extra = d_grads[2 * l + 1];
        B.add_gemm(0, OP_GEMM_IJ, w);
    }
    if (!B.failed) P->d_ops = (ProgOp *)B.alloc(sizeof(ProgOp) * B.ops.size(), false);
    if (B.failed || cudaMemcpy(P->d_ops, B.ops.data(), sizeof(ProgOp) * B.ops.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        mb_mlp_program_destroy(P);
        return MB_ERR_NOMEM;
    }
    P->n_ops = (int)B.ops.size();
    P->n_phases = 1;
    int items = 0;
    for (const ProgOp &op : B.ops) items += op.items;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mlp_program_kernel, NT, 0);
    P->grid = sms * (per_sm >= 2 ? 2 : 1);
    if (P->grid > items) P->grid = items;
    if (P->grid < 1) P->grid = 1;
    const int xp = ((sizes[0] + 3) & ~3) + 4;
    P->rows_smem = sizeof(float) * (size_t)(2 * RB) * (size_t)(xp + 2 * RP);
    P->rows_grid = (batch + RB - 1) / RB;
    // one fixed opt-in limit for every program on this device (a later, smaller program must not lower it)
    {
        const char *e = getenv("MB_ROWS_CPT");
        P->rows_cpt = (e && e[0] == '2') ? 2 : 4;
    }
    if (P->rows_smem > 160 * 1024 ||
        cudaFuncSetAttribute(dqn_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(dqn_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess) {
        cudaGetLastError();
        mb_mlp_program_destroy(P);
        return MB_ERR_INVALID;
    }
    MB_CUDA(cudaDeviceSynchronize());
    *out = P;
    return MB_OK;
}

int mb_mlp_program_run_part(mb_mlp_program *P, int part, void *stream) {
    if (!P || part < 0 || part > 2) return MB_ERR_INVALID;
    mb::DeviceGuard guard(P->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (part == 0) {                                   // refresh the k-major weight copies (row-tile program only)
        if (!P->rows) return MB_OK;
        const int64_t total = P->tl.start[P->tl.n];
        int64_t blocks = (total + 255) / 256;
        if (blocks > (int64_t)mb::kNumSMs * 4) blocks = (int64_t)mb::kNumSMs * 4;
        transpose_multi_kernel<<<(unsigned)blocks, 256, 0, s>>>(P->tl);
        MB_LAUNCHED(1);
        return MB_OK;
    }
    if (P->rows && part == 1) {                        // forward x3 + TD + backward-input, one CTA per 8 batch rows
        if (P->rows_cpt == 2) dqn_rows_kernel<2><<<(unsigned)P->rows_grid, RT, P->rows_smem, s>>>(P->ra);
        else dqn_rows_kernel<4><<<(unsigned)P->rows_grid, RT, P->rows_smem, s>>>(P->ra);
        MB_LAUNCHED(1);
        return MB_OK;
    }
    if (!P->rows && part == 2) return MB_OK;           // the cooperative program does everything in part 1
    void *args[] = {&P->d_ops, &P->n_ops, &P->n_phases};
    MB_CUDA(cudaLaunchCooperativeKernel((const void *)mlp_program_kernel, dim3((unsigned)P->grid), dim3(NT), args, 0, s));
    mb::count_launch(1);
    return MB_OK;
}

int mb_mlp_program_run(mb_mlp_program *P, void *stream) {
    for (int part = 0; part < 3; ++part) {
        const int r = mb_mlp_program_run_part(P, part, stream);
        if (r != MB_OK) return r;
    }
    return MB_OK;
}

int mb_mlp_program_destroy(mb_mlp_program *P) {
    if (!P) return MB_OK;
    mb::DeviceGuard guard(P->device);
    for (void *p : P->owned) cudaFree(p);
    delete P;
    return MB_OK;
}

int mb_adam_soft_step_multi(float *const *d_params, const float *const *d_grads, float *const *d_exp_avg,
                            float *const *d_exp_avg_sq, float *const *d_target_params, const int64_t *sizes,
                            int n_tensors, float *d_step_count, unsigned *d_ticket, double lr, double beta1, double beta2,
                            double eps, double weight_decay, double update_rate, void *stream) {
    if (n_tensors < 0 || n_tensors > MB_MAX_SOFT_TENSORS || !d_step_count || !d_ticket || (n_tensors > 0 && (!d_params || !d_grads || !d_exp_avg || !d_exp_avg_sq ||
                                                              !d_target_params || !sizes)))
        return MB_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    for (int base = 0; base < n_tensors; base += MB_MAX_SOFT_TENSORS) {
        AdamSoftList l;
        l.n = n_tensors - base < MB_MAX_SOFT_TENSORS ? n_tensors - base : MB_MAX_SOFT_TENSORS;
        l.start[0] = 0;
        for (int k = 0; k < l.n; ++k) {
            l.p[k] = d_params[base + k]; l.g[k] = d_grads[base + k];
            l.m[k] = d_exp_avg[base + k]; l.v[k] = d_exp_avg_sq[base + k];
            l.t[k] = d_target_params[base + k];
            l.start[k + 1] = l.start[k] + sizes[base + k];
        }
        const int64_t total = l.start[l.n];
        if (total == 0) continue;
        int64_t blocks = (total + 255) / 256;
        if (blocks > (int64_t)mb::kNumSMs * 8) blocks = (int64_t)mb::kNumSMs * 8;
        // python: target * (1.0 - update_rate) + param * update_rate ; both scalars become float32 (csrc/td.cu)
        MB_PREFER_MAX_SMEM(adam_soft_kernel);
        adam_soft_kernel<<<(unsigned)blocks, 256, 0, s>>>(l, (float)lr, (float)beta1, (float)beta2, (float)eps,
                                                          (float)weight_decay, d_step_count, d_ticket,
                                                          (float)(1.0 - update_rate), (float)update_rate);
        MB_LAUNCHED(1);                                   // (one chunk: the kernel advances the step counter itself)
    }
    return MB_OK;
}

int mb_adam_step_multi(float *const *d_params, const float *const *d_grads, float *const *d_exp_avg,
                       float *const *d_exp_avg_sq, const int64_t *sizes, int n_tensors, float *d_step_count,
                       double lr, double beta1, double beta2, double eps, double weight_decay, void *stream) {
    if (n_tensors < 0 || !d_step_count) return MB_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    for (int base = 0; base < n_tensors; base += MB_MAX_SOFT_TENSORS) {
        AdamList l;
        l.n = n_tensors - base < MB_MAX_SOFT_TENSORS ? n_tensors - base : MB_MAX_SOFT_TENSORS;
        l.start[0] = 0;
        for (int k = 0; k < l.n; ++k) {
            l.p[k] = d_params[base + k]; l.g[k] = d_grads[base + k];
            l.m[k] = d_exp_avg[base + k]; l.v[k] = d_exp_avg_sq[base + k];
            l.start[k + 1] = l.start[k] + sizes[base + k];
        }
        const int64_t total = l.start[l.n];
        if (total == 0) continue;
        int64_t blocks = (total + 255) / 256;
        if (blocks > (int64_t)mb::kNumSMs * 8) blocks = (int64_t)mb::kNumSMs * 8;
        adam_kernel<<<(unsigned)blocks, 256, 0, s>>>(l, (float)lr, (float)beta1, (float)beta2, (float)eps,
                                                     (float)weight_decay, d_step_count);
        mb::count_launch(1);
    }
    adam_step_inc<<<1, 1, 0, s>>>(d_step_count);
    MB_LAUNCHED(1);
    return MB_OK;
}

}  // extern "C"
