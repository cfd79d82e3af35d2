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
#include "mb_common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16, NT = 256;

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

template <bool kAKfast, bool kBKfast>
__global__ void __launch_bounds__(NT) gemm_kernel(GemmArgs g) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
    const int k_begin = blockIdx.z * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);
    const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
    for (int k0 = k_begin; k0 < k_end; k0 += TK) {
#pragma unroll
        for (int e = tid; e < TM * TK; e += NT) {
            int i, k;
            if (kAKfast) { i = e / TK; k = e % TK; } else { k = e / TM; i = e % TM; }
            const int gi = i0 + i, gk = k0 + k;
            As[k][i] = (gi < g.M && gk < k_end) ? g.A[(int64_t)gi * g.a_i + (int64_t)gk * g.a_k] : 0.0f;
        }
#pragma unroll
        for (int e = tid; e < TN * TK; e += NT) {
            int j, k;
            if (kBKfast) { j = e / TK; k = e % TK; } else { k = e / TN; j = e % TN; }
            const int gj = j0 + j, gk = k0 + k;
            float v = 0.0f;
            if (gj < g.N && gk < k_end) v = (gj == g.ones_col) ? 1.0f : g.B[(int64_t)gj * g.b_j + (int64_t)gk * g.b_k];
            Bs[k][j] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
        }
        __syncthreads();
    }
    const int splits = gridDim.z;
    if (splits == 1) {
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) epilogue_store(g, i0 + ty * 4 + x, j0 + tx * 4 + y, acc[x][y]);
        return;
    }
    // split-K: park the partial tile, the last CTA of this tile sums the partials in split order
    const int64_t MN = (int64_t)g.M * g.N;
    float *mine = g.partial + (int64_t)blockIdx.z * MN;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int i = i0 + ty * 4 + x, j = j0 + tx * 4 + y;
            if (i < g.M && j < g.N) mine[(int64_t)i * g.N + j] = acc[x][y];
        }
    __threadfence();
    __syncthreads();
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) s_last = (atomicAdd(g.tile_counter + tile, 1) == splits - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int i = i0 + ty * 4 + x, j = j0 + tx * 4 + y;
            if (i < g.M && j < g.N) {
                float v = 0.0f;
                for (int s = 0; s < splits; ++s) v = __fadd_rn(v, __ldcg(g.partial + (int64_t)s * MN + (int64_t)i * g.N + j));
                epilogue_store(g, i, j, v);
            }
        }
    if (tid == 0) g.tile_counter[tile] = 0;
}

int pick_splits(int tiles, int K) {
    int s = 1;
    while (tiles * s * 2 <= 160 && (K / (s * 2)) >= 32 && s < 16) s *= 2;
    return s;
}

template <bool kA, bool kB>
int launch(GemmArgs g, float *workspace, int64_t workspace_floats, int *counters, int n_counters, cudaStream_t s) {
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
