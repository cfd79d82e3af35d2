// Row-tile DQN value math: device code only (no CUDA runtime calls), included by mlp.cu inside its anonymous
// namespace -- and by tests/cpu_emul/rows_harness.cpp, which compiles THIS SOURCE with g++ against a small shim
// (threads = OS threads, __syncthreads = a barrier) so that its indexing can be checked where no GPU exists.
// See the comment block in mlp.cu for the design.
#pragma once
#include <stdint.h>

#include "td_device.cuh"

constexpr int RB = 8;            // batch rows per CTA
constexpr int RT = 256;          // threads: 4 row groups x 64 column groups of 4
constexpr int RW = 256;          // widest layer output one CTA sweeps
constexpr int RP = RW + 4;       // activation row pitch in shared memory (floats, multiple of 4)
constexpr int RL = 8;            // layers

struct RowsArgs {
    int L, batch;
    int sizes[RL + 1];
    const float *W[RL];          // online parameters [out, in] (backward-input reads them as B(k = out, j = in))
    const float *Wt[RL];         // k-major copies [in, out] of the online / target weights (forward)
    const float *tWt[RL];
    const float *b[RL], *tb[RL];
    const float *state, *next_state;
    float *h[RL];                // Q(s) pass: output of layer l, [batch, sizes[l + 1]] (post-ReLU; last = q)
    float *qn, *qtn;             // [batch, A]
    float *dz[RL];               // gradient w.r.t. the output of layer l (pre-mask of the layer below)
    TdDqnArgs td;
};

// Thread layout, chosen at compile time: CPT output columns per thread.  RT threads form G = RT * CPT / RW row groups
// of RW / CPT threads; a thread owns MR = M / G rows x CPT columns of the M x N output.
//   CPT = 4: 4 row groups x 64 threads, one 16-byte weight load per k (4 LSU wavefronts per warp), MR = 4 (online);
//   CPT = 2: 2 row groups x 128 threads, one 8-byte weight load per k (2 wavefronts), MR = 8: every weight is
//            fetched by half as many row groups -- 4 LSU wavefronts per 16 FFMA instead of 5.
template <int CPT> struct RowsLayout {
    static constexpr int kThreadsPerGroup = RW / CPT;
    static constexpr int kGroups = RT / kThreadsPerGroup;
};

template <int CPT, bool kVec>
__device__ __forceinline__ void rows_load_b(const float *__restrict__ row, int j0, int N, float (&v)[CPT]) {
    if (kVec) {
        if (CPT == 4) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(row + j0));
            v[0] = x.x; v[1] = x.y; v[CPT - 2] = x.z; v[CPT - 1] = x.w;
        } else {
            const float2 x = __ldg(reinterpret_cast<const float2 *>(row + j0));
            v[0] = x.x; v[1] = x.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < CPT; ++c) v[c] = j0 + c < N ? __ldg(row + j0 + c) : 0.0f;
    }
}

// acc[r][c] = sum_k As[(ty * MR + r) * ap + k] * Bg[k * N + j0 + c]      (rows in shared memory, B k-major in global)
// The weights come straight from L2 and a CTA of 8 warps has an SM to itself, so nothing else hides the ~700-cycle
// latency: three register buffers of 8 k-steps rotate -- while one is multiplied (~320 cycles) the loads of the
// next two are in flight -- and the tail of K is loaded predicated in the same batches instead of one dependent
// load per step.  k ascends, one fmaf per term: the summation order does not depend on the buffering or the layout.
template <int MR, int CPT, bool kVec>
__device__ __forceinline__ void rows_matmul_impl(const float *As, int ap, int K, const float *__restrict__ Bg, int N,
                                                 float (&acc)[MR][CPT]) {
    constexpr int NG = RowsLayout<CPT>::kThreadsPerGroup;
    const int ty = threadIdx.x / NG, j0 = (threadIdx.x % NG) * CPT;
    const float *a0 = As + (ty * MR) * ap;
    auto load8 = [&](float (&b)[8][CPT], int k) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (k + u < K) {
                rows_load_b<CPT, kVec>(Bg + (int64_t)(k + u) * N, j0, N, b[u]);
            } else {
#pragma unroll
                for (int c = 0; c < CPT; ++c) b[u][c] = 0.0f;
            }
        }
    };
    auto mul8 = [&](const float (&b)[8][CPT], int k) {
        if (k >= K) return;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int kh = k + 4 * h;
            if (kh >= K) break;
            float av[MR][4];
#pragma unroll
            for (int r = 0; r < MR; ++r) {
                if (kh + 4 <= K) {                                // rows are 16-byte aligned at multiples of 4
                    const float4 a = *reinterpret_cast<const float4 *>(a0 + r * ap + kh);   // warp-wide broadcast
                    av[r][0] = a.x; av[r][1] = a.y; av[r][2] = a.z; av[r][3] = a.w;
                } else {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) av[r][kk] = kh + kk < K ? a0[r * ap + kh + kk] : 0.0f;
                }
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (kh + kk >= K) break;                          // exact term count: no +0 terms
#pragma unroll
                for (int r = 0; r < MR; ++r)
#pragma unroll
                    for (int c = 0; c < CPT; ++c) acc[r][c] = fmaf(av[r][kk], b[4 * h + kk][c], acc[r][c]);
            }
        }
    };
    float b0[8][CPT], b1[8][CPT], b2[8][CPT];
    load8(b0, 0);
    load8(b1, 8);
    for (int k = 0; k < K; k += 24) {
        load8(b2, k + 16);
        mul8(b0, k);
        load8(b0, k + 24);
        mul8(b1, k + 8);
        load8(b1, k + 32);
        mul8(b2, k + 16);
    }
}

template <int MR, int CPT>
__device__ __forceinline__ void rows_matmul(const float *As, int ap, int K, const float *__restrict__ Bg, int N,
                                            float (&acc)[MR][CPT]) {
#pragma unroll
    for (int r = 0; r < MR; ++r)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[r][c] = 0.0f;
    if ((int)(threadIdx.x % RowsLayout<CPT>::kThreadsPerGroup) * CPT >= N) return;
    // N % CPT == 0: a thread's columns all exist and every row of Bg is aligned for one vector load per k
    if (N % CPT == 0) rows_matmul_impl<MR, CPT, true>(As, ap, K, Bg, N, acc);
    else rows_matmul_impl<MR, CPT, false>(As, ap, K, Bg, N, acc);
}

// One forward layer for M = kGroups * MR rows held in shared memory.  save0: rows [0, RB) also go to global at batch
// row b0 + row (activations of the Q(s) pass / q); save1: rows [row1, row1 + RB) go to global (qn or qtn).
template <int MR, int CPT>
__device__ __forceinline__ void rows_forward_layer(const float *in, int ip, float *out, int K, int N,
                                                   const float *__restrict__ Wt, const float *__restrict__ bias,
                                                   bool relu, int b0, int batch, float *save0, float *save1, int row1) {
    constexpr int NG = RowsLayout<CPT>::kThreadsPerGroup;
    float acc[MR][CPT];
    rows_matmul<MR, CPT>(in, ip, K, Wt, N, acc);
    const int ty = threadIdx.x / NG, j0 = (threadIdx.x % NG) * CPT;
    if (j0 < N) {
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            const int row = ty * MR + r;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int j = j0 + c;
                if (j >= N) continue;
                float v = __fadd_rn(acc[r][c], __ldg(bias + j));
                if (relu) v = fmaxf(v, 0.0f);
                out[row * RP + j] = v;
                if (save0 && row < RB && b0 + row < batch) save0[(int64_t)(b0 + row) * N + j] = v;
                if (save1 && row >= row1 && row < row1 + RB && b0 + row - row1 < batch)
                    save1[(int64_t)(b0 + row - row1) * N + j] = v;
            }
        }
    }
    __syncthreads();
}

template <int CPT>
__global__ void __launch_bounds__(RT) dqn_rows_kernel(RowsArgs a) {
    constexpr int G = RowsLayout<CPT>::kGroups, NG = RowsLayout<CPT>::kThreadsPerGroup;
    static_assert((2 * RB) % G == 0 && RB % G == 0, "row groups must divide the row counts");
    extern __shared__ __align__(16) float rsm[];
    const int S = a.sizes[0], xp = ((S + 3) & ~3) + 4, L = a.L, A = a.sizes[L];
    float *X = rsm;                      // [2 * RB][xp]: rows 0..RB-1 = s, RB..2RB-1 = s'
    float *B0 = X + 2 * RB * xp;         // [2 * RB][RP]
    float *B1 = B0 + 2 * RB * RP;
    const int tid = threadIdx.x, b0 = blockIdx.x * RB;
    for (int e = tid; e < 2 * RB * xp; e += RT) {
        const int row = e / xp, k = e % xp, b = b0 + (row % RB);
        const float *src = row < RB ? a.state : a.next_state;
        X[e] = (k < S && b < a.batch) ? __ldg(src + (int64_t)b * S + k) : 0.0f;
    }
    __syncthreads();
    // ---- online net over [s; s'] (16 rows per weight load)
    {
        const float *in = X;
        int ip = xp;
        for (int l = 0; l < L; ++l) {
            float *out = (l & 1) ? B1 : B0;
            const bool last = l == L - 1;
            rows_forward_layer<2 * RB / G, CPT>(in, ip, out, a.sizes[l], a.sizes[l + 1], a.Wt[l], a.b[l], !last, b0, a.batch,
                                                a.h[l], last ? a.qn : nullptr, RB);
            in = out;
            ip = RP;
        }
    }
    // ---- target net over s'
    {
        const float *in = X + RB * xp;
        int ip = xp;
        for (int l = 0; l < L; ++l) {
            float *out = (l & 1) ? B1 : B0;
            const bool last = l == L - 1;
            rows_forward_layer<RB / G, CPT>(in, ip, out, a.sizes[l], a.sizes[l + 1], a.tWt[l], a.tb[l], !last, b0, a.batch,
                                            nullptr, last ? a.qtn : nullptr, 0);
            in = out;
            ip = RP;
        }
    }
    // ---- TD row math on this CTA's rows (q, qn, qtn were written by this CTA; the barrier above orders them)
    if (tid < RB && b0 + tid < a.batch) td_dqn_row(a.td, b0 + tid);
    __syncthreads();
    // ---- backward-input chain: dz[l-1] = (dz[l] W_l) * (h[l-1] > 0)
    float *cur = B0, *nxt = B1;
    for (int e = tid; e < RB * RP; e += RT) {
        const int row = e / RP, j = e % RP;
        cur[e] = (j < A && b0 + row < a.batch) ? a.td.grad_q[(int64_t)(b0 + row) * A + j] : 0.0f;
    }
    __syncthreads();
    for (int l = L - 1; l >= 1; --l) {
        const int K = a.sizes[l + 1], N = a.sizes[l];
        constexpr int MRB = RB / G;
        float acc[MRB][CPT];
        rows_matmul<MRB, CPT>(cur, RP, K, a.W[l], N, acc);
        const int ty = tid / NG, j0 = (tid % NG) * CPT;
        if (j0 < N) {
#pragma unroll
            for (int r = 0; r < MRB; ++r) {
                const int row = ty * MRB + r, b = b0 + row;
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int j = j0 + c;
                    if (j >= N) continue;
                    float v = 0.0f;
                    if (b < a.batch) {
                        v = a.h[l - 1][(int64_t)b * N + j] > 0.0f ? acc[r][c] : 0.0f;
                        a.dz[l - 1][(int64_t)b * N + j] = v;
                    }
                    nxt[row * RP + j] = v;
                }
            }
        }
        __syncthreads();
        float *t = cur; cur = nxt; nxt = t;
    }
}
