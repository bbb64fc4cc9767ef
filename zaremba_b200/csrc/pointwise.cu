// HBM-bound kernels of the path: embedding gather/scatter, dropout, LSTM cell pointwise
// math (validation engine), bias/column sums, softmax-NLL.  Coalesced row-major access;
// no tensor-core work here.
#include "kernels.h"

namespace zrb {

// ----------------------------------------------------------------------------------------
// Embedding gather fused with the first dropout site.   model.py:13-14, :105
// grid: one block per token row; threads stride the row in groups of 4 (one Philox call
// yields the keep flags of 4 consecutive elements).
// Algorithmic bytes per token: H*4 read (table row) + H*4 written (+ H*2 fp16 image).
// ----------------------------------------------------------------------------------------
__global__ void embed_dropout_fwd_kernel(const float* __restrict__ W, const int64_t* __restrict__ idx,
                                         float* __restrict__ out, __half* __restrict__ out_h, int64_t ld_h, int N,
                                         int H, int V, MaskSrc m) {
    int n = blockIdx.x;
    if (n >= N) return;
    int64_t row = idx[n];
    if (row < 0 || row >= V) row = 0;  // the reference would raise; never dereference OOB
    const float* src = W + row * (int64_t)H;
    int groups = (H + 3) >> 2;
    uint64_t n_total = (uint64_t)N * H;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        int j0 = g << 2;
        uint64_t e0 = (uint64_t)n * H + j0;
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (j0 + i < H) ? src[j0 + i] : 0.f;
        if (m.active) {
            // rows are H long; H % 4 != 0 makes groups straddle Philox quads -> per-element path
            if ((H & 3) == 0) {
                uint32_t bits = mask_keep4(m, e0 >> 2, n_total);
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = ((bits >> i) & 1u) ? v[i] * m.scale : 0.f;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] *= mask_mul1(m, e0 + i, n_total);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (j0 + i < H) {
                if (out) out[(int64_t)n * H + j0 + i] = v[i];
                if (out_h) out_h[(int64_t)n * ld_h + j0 + i] = __float2half_rn(v[i]);
            }
        }
    }
}

int embed_dropout_fwd(const float* W, const int64_t* idx, float* out, __half* out_h, int64_t ld_h, int N, int H,
                      int V, MaskSrc m, cudaStream_t s) {
    if (N == 0) return ZRB_OK;
    int threads = H >= 1024 ? 256 : 128;
    embed_dropout_fwd_kernel<<<N, threads, 0, s>>>(W, idx, out, out_h, ld_h, N, H, V, m);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

// Backward of the gather: scatter-add with fp32 atomics (duplicate tokens in a window hit
// the same row).  dW must be zeroed by the caller.
__global__ void embed_dropout_bwd_kernel(const float* __restrict__ dA, const int64_t* __restrict__ idx,
                                         float* __restrict__ dW, int N, int H, int V, MaskSrc m) {
    int n = blockIdx.x;
    if (n >= N) return;
    int64_t row = idx[n];
    if (row < 0 || row >= V) return;
    uint64_t n_total = (uint64_t)N * H;
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        float g = dA[(int64_t)n * H + j] * mask_mul1(m, (uint64_t)n * H + j, n_total);
        if (g != 0.f) atomicAdd(dW + row * (int64_t)H + j, g);
    }
}

int embed_dropout_bwd(const float* dA, const int64_t* idx, float* dW, int N, int H, int V, MaskSrc m,
                      cudaStream_t s) {
    if (N == 0) return ZRB_OK;
    embed_dropout_bwd_kernel<<<N, 256, 0, s>>>(dA, idx, dW, N, H, V, m);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

// ----------------------------------------------------------------------------------------
// LSTM cell pointwise (validation engine).  model.py:37-45 in nn.LSTM's (i,f,g,o) order.
// ----------------------------------------------------------------------------------------
__global__ void lstm_cell_fwd_kernel(float* __restrict__ pre, const float* __restrict__ c_prev,
                                     float* __restrict__ c_out, float* __restrict__ h_raw, float* __restrict__ y_out,
                                     int B, int H, int64_t elem_off, int64_t n_total, MaskSrc m) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)B * H) return;
    int b = (int)(tid / H), j = (int)(tid % H);
    float* row = pre + (int64_t)b * 4 * H;
    float i = sigmoidf_(row[j]);
    float f = sigmoidf_(row[H + j]);
    float g = tanhf(row[2 * H + j]);
    float o = sigmoidf_(row[3 * H + j]);
    float c = f * c_prev[tid] + i * g;
    float h = o * tanhf(c);
    row[j] = i; row[H + j] = f; row[2 * H + j] = g; row[3 * H + j] = o;
    c_out[tid] = c;
    h_raw[tid] = h;
    y_out[tid] = h * mask_mul1(m, (uint64_t)(elem_off + tid), (uint64_t)n_total);
}

int lstm_cell_fwd(float* pre, const float* c_prev, float* c_out, float* h_raw, float* y_out, int B, int H,
                  int64_t elem_off, int64_t n_total, MaskSrc m, cudaStream_t s) {
    int64_t n = (int64_t)B * H;
    lstm_cell_fwd_kernel<<<cdiv(n, 256), 256, 0, s>>>(pre, c_prev, c_out, h_raw, y_out, B, H, elem_off, n_total, m);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

// SURVEY 8a backward: dh = mask*dy + dh_rec; do = dh*tanh(c); dc += dh*o*(1-tanh^2 c); ...
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ dy_post, const float* __restrict__ dh_rec,
                                     float* __restrict__ dc, const float* __restrict__ gates,
                                     const float* __restrict__ c_t, const float* __restrict__ c_prev,
                                     float* __restrict__ dG, int B, int H, int64_t elem_off, int64_t n_total,
                                     MaskSrc m) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)B * H) return;
    int b = (int)(tid / H), j = (int)(tid % H);
    const float* row = gates + (int64_t)b * 4 * H;
    float i = row[j], f = row[H + j], g = row[2 * H + j], o = row[3 * H + j];
    float dh = dy_post[tid] * mask_mul1(m, (uint64_t)(elem_off + tid), (uint64_t)n_total);
    if (dh_rec) dh += dh_rec[tid];
    float tc = tanhf(c_t[tid]);
    float d_o = dh * tc;
    float dcc = dc[tid] + dh * o * (1.f - tc * tc);
    float d_i = dcc * g;
    float d_g = dcc * i;
    float d_f = dcc * c_prev[tid];
    dc[tid] = dcc * f;
    float* drow = dG + (int64_t)b * 4 * H;
    drow[j] = d_i * i * (1.f - i);
    drow[H + j] = d_f * f * (1.f - f);
    drow[2 * H + j] = d_g * (1.f - g * g);
    drow[3 * H + j] = d_o * o * (1.f - o);
}

int lstm_cell_bwd(const float* dy_post, const float* dh_rec, float* dc, const float* gates, const float* c_t,
                  const float* c_prev, float* dG, int B, int H, int64_t elem_off, int64_t n_total, MaskSrc m,
                  cudaStream_t s) {
    int64_t n = (int64_t)B * H;
    lstm_cell_bwd_kernel<<<cdiv(n, 256), 256, 0, s>>>(dy_post, dh_rec, dc, gates, c_t, c_prev, dG, B, H, elem_off,
                                                      n_total, m);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

// ----------------------------------------------------------------------------------------
__global__ void add_bias_kernel(float* __restrict__ C, const float* __restrict__ b1, const float* __restrict__ b2,
                                int64_t total, int M) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int j = (int)(i % M);
    float v = b1[j];
    if (b2) v += b2[j];
    C[i] += v;
}
int add_bias2(float* C, const float* b1, const float* b2, int N, int M, cudaStream_t s) {
    int64_t total = (int64_t)N * M;
    if (!total) return ZRB_OK;
    add_bias_kernel<<<cdiv(total, 256), 256, 0, s>>>(C, b1, b2, total, M);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}
int add_bias1(float* C, const float* b1, int N, int M, cudaStream_t s) { return add_bias2(C, b1, nullptr, N, M, s); }

// out[j] = sum_n A[n,j]: a block owns 32 columns; 8 row-lanes stride the rows, then a
// shared-memory tree over the 8 partials (deterministic).
__global__ void colsum_kernel(const float* __restrict__ A, float* __restrict__ out, float* __restrict__ out2, int N,
                              int M) {
    __shared__ float part[8][33];
    int col = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (col < M)
        for (int n = threadIdx.y; n < N; n += 8) acc += A[(int64_t)n * M + col];
    part[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && col < M) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += part[r][threadIdx.x];
        out[col] = t;
        if (out2) out2[col] = t;
    }
}
int colsum(const float* A, float* out, float* out2, int N, int M, cudaStream_t s) {
    dim3 blk(32, 8);
    colsum_kernel<<<cdiv(M, 32), blk, 0, s>>>(A, out, out2, N, M);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

// ----------------------------------------------------------------------------------------
// softmax-NLL forward (+ gradient) -- main.py:77-84.  One block per token row; the row is
// read twice from L2/HBM (max+sum pass fused as online softmax, then the write pass).
// The reference exponentiates without subtracting the max (overflows beyond ~88); the
// max-subtracted form is the same function where the reference is finite.
// Algorithmic bytes per token: V*4 read + V*4 written (dscores).
// ----------------------------------------------------------------------------------------
__global__ void softmax_nll_kernel(const float* __restrict__ scores, const int64_t* __restrict__ y, int N, int V,
                                   float gscale, float* __restrict__ row_loss, float* __restrict__ dscores,
                                   float* __restrict__ tgt_prob, __half* __restrict__ ds_h, int64_t ld_s,
                                   float h_scale) {
    __shared__ float s_m[32], s_s[32];
    int n = blockIdx.x;
    const float* row = scores + (int64_t)n * V;
    float mx = -INFINITY, sum = 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        float z = row[v];
        if (z > mx) {
            sum = sum * __expf(mx - z) + 1.f;
            mx = z;
        } else {
            sum += __expf(z - mx);
        }
    }
    // warp then block combine of (max, sum) pairs
    for (int o = 16; o > 0; o >>= 1) {
        float om = __shfl_xor_sync(0xffffffffu, mx, o);
        float os = __shfl_xor_sync(0xffffffffu, sum, o);
        float nm = fmaxf(mx, om);
        sum = (mx == -INFINITY ? 0.f : sum * __expf(mx - nm)) + (om == -INFINITY ? 0.f : os * __expf(om - nm));
        mx = nm;
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (l == 0) { s_m[w] = mx; s_s[w] = sum; }
    __syncthreads();
    if (w == 0) {
        mx = l < nw ? s_m[l] : -INFINITY;
        sum = l < nw ? s_s[l] : 0.f;
        for (int o = 16; o > 0; o >>= 1) {
            float om = __shfl_xor_sync(0xffffffffu, mx, o);
            float os = __shfl_xor_sync(0xffffffffu, sum, o);
            float nm = fmaxf(mx, om);
            sum = (mx == -INFINITY ? 0.f : sum * __expf(mx - nm)) + (om == -INFINITY ? 0.f : os * __expf(om - nm));
            mx = nm;
        }
        if (l == 0) { s_m[0] = mx; s_s[0] = sum; }
    }
    __syncthreads();
    mx = s_m[0];
    sum = s_s[0];
    int64_t tgt = y[n];
    float inv = 1.f / sum;
    if (threadIdx.x == 0) {
        float zt = (tgt >= 0 && tgt < V) ? row[tgt] : mx;
        row_loss[n] = -(zt - mx - logf(sum));
        if (tgt_prob) tgt_prob[n] = expf(zt - mx) * inv;
    }
    if (dscores || ds_h) {
        float* drow = dscores ? dscores + (int64_t)n * V : nullptr;
        __half* hrow = ds_h ? ds_h + (int64_t)n * ld_s : nullptr;
        for (int v = threadIdx.x; v < V; v += blockDim.x) {
            float p = __expf(row[v] - mx) * inv;
            if (v == tgt) p -= 1.f;
            p *= gscale;
            if (drow) drow[v] = p;
            if (hrow) hrow[v] = __float2half_rn(fminf(fmaxf(p * h_scale, -65504.f), 65504.f));
        }
    }
}

// Same function with the row held in registers: one block per token, every thread loads NV float4 chunks ONCE
// (the scalar kernel above reads the row twice and carries the online-softmax recurrence through every element),
// block max, one exp per element, block sum, then the gradient is written from the registers with 16-byte
// (fp32) / 8-byte (fp16 image) stores.  Needs V % 4 == 0, V <= 4 * 512 * NV and 16-byte aligned rows.
constexpr int kSmThreads = 512;
__device__ __forceinline__ float block_reduce_512(float v, float* sh, bool is_max) {
    for (int o = 16; o > 0; o >>= 1) {
        float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, t) : v + t;
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();                       // sh may still be read from the previous reduction
    if (l == 0) sh[w] = v;
    __syncthreads();
    v = l < kSmThreads / 32 ? sh[l] : (is_max ? -INFINITY : 0.f);
    for (int o = 16; o > 0; o >>= 1) {
        float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, t) : v + t;
    }
    return v;                              // every thread holds the block result (fixed tree: deterministic)
}
template <int NV>
__global__ void __launch_bounds__(kSmThreads) softmax_nll_reg_kernel(
    const float* __restrict__ scores, const int64_t* __restrict__ y, int N, int V, float gscale,
    float* __restrict__ row_loss, float* __restrict__ dscores, float* __restrict__ tgt_prob, __half* __restrict__ ds_h,
    int64_t ld_s, float h_scale) {
    __shared__ float sh[32];
    const int n = blockIdx.x;
    const float* row = scores + (int64_t)n * V;
    const float4* row4 = reinterpret_cast<const float4*>(row);
    const int nv4 = V >> 2;
    float4 z[NV];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int i = threadIdx.x + k * kSmThreads;
        if (i < nv4) {
            z[k] = __ldcs(row4 + i);
            mx = fmaxf(mx, fmaxf(fmaxf(z[k].x, z[k].y), fmaxf(z[k].z, z[k].w)));
        }
    }
    mx = block_reduce_512(mx, sh, true);
    if (mx == -INFINITY) mx = 0.f;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int i = threadIdx.x + k * kSmThreads;
        if (i < nv4) {
            z[k].x = __expf(z[k].x - mx); z[k].y = __expf(z[k].y - mx);
            z[k].z = __expf(z[k].z - mx); z[k].w = __expf(z[k].w - mx);
            sum += (z[k].x + z[k].y) + (z[k].z + z[k].w);
        }
    }
    sum = block_reduce_512(sum, sh, false);
    const float inv = 1.f / sum;
    const int64_t tgt = y[n];
    if (threadIdx.x == 0) {
        const float zt = (tgt >= 0 && tgt < V) ? row[tgt] : mx;
        row_loss[n] = -(zt - mx - logf(sum));
        if (tgt_prob) tgt_prob[n] = expf(zt - mx) * inv;
    }
    if (dscores || ds_h) {
        float4* drow = dscores ? reinterpret_cast<float4*>(dscores + (int64_t)n * V) : nullptr;
        uint2* hrow = ds_h ? reinterpret_cast<uint2*>(ds_h + (int64_t)n * ld_s) : nullptr;
        const float g = gscale * inv;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int i = threadIdx.x + k * kSmThreads;
            if (i < nv4) {
                float4 p = make_float4(z[k].x * g, z[k].y * g, z[k].z * g, z[k].w * g);
                const int64_t d = tgt - 4 * (int64_t)i;        // the target's position inside this chunk, if any
                if (d == 0) p.x -= gscale; else if (d == 1) p.y -= gscale; else if (d == 2) p.z -= gscale;
                else if (d == 3) p.w -= gscale;
                if (drow) drow[i] = p;
                if (hrow) {
                    const float lo = -65504.f, hi = 65504.f;
                    __half2 a = __floats2half2_rn(fminf(fmaxf(p.x * h_scale, lo), hi), fminf(fmaxf(p.y * h_scale, lo), hi));
                    __half2 b = __floats2half2_rn(fminf(fmaxf(p.z * h_scale, lo), hi), fminf(fmaxf(p.w * h_scale, lo), hi));
                    uint2 u;
                    u.x = *reinterpret_cast<uint32_t*>(&a);
                    u.y = *reinterpret_cast<uint32_t*>(&b);
                    hrow[i] = u;
                }
            }
        }
    }
}

// loss = (B / N) * sum_n row_loss[n]  (fixed-order tree: deterministic)
__global__ void loss_reduce_kernel(const float* __restrict__ row_loss, int N, float scale, float* __restrict__ loss) {
    __shared__ float sh[32];
    float acc = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) acc += row_loss[n];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) *loss = v * scale;
    }
}

int softmax_nll(const float* scores, const int64_t* y, int N, int V, int B, float* row_loss, float* loss,
                float* dscores, float* tgt_prob, cudaStream_t s, __half* ds_h, int64_t ld_s, float h_scale) {
    if (N == 0) return ZRB_OK;
    float gscale = (float)((double)B / (double)N);
    const bool vec = V % 4 == 0 && V <= 4 * kSmThreads * 8 && (((uintptr_t)scores) & 15) == 0 &&
                     (!dscores || (((uintptr_t)dscores) & 15) == 0) &&
                     (!ds_h || ((((uintptr_t)ds_h) & 7) == 0 && ld_s % 4 == 0));
    const int nv = vec ? (V / 4 + kSmThreads - 1) / kSmThreads : 0;
#define ZRB_SM_LAUNCH(NV) \
    softmax_nll_reg_kernel<NV><<<N, kSmThreads, 0, s>>>(scores, y, N, V, gscale, row_loss, dscores, tgt_prob, ds_h, ld_s, h_scale)
    if (vec && nv <= 2) ZRB_SM_LAUNCH(2);
    else if (vec && nv <= 4) ZRB_SM_LAUNCH(4);
    else if (vec && nv <= 6) ZRB_SM_LAUNCH(6);
    else if (vec) ZRB_SM_LAUNCH(8);
    else softmax_nll_kernel<<<N, 512, 0, s>>>(scores, y, N, V, gscale, row_loss, dscores, tgt_prob, ds_h, ld_s, h_scale);
#undef ZRB_SM_LAUNCH
    ZRB_KERNEL_CHECK();
    if (loss) {
        loss_reduce_kernel<<<1, 256, 0, s>>>(row_loss, N, gscale, loss);
        ZRB_KERNEL_CHECK();
    }
    return ZRB_OK;
}

// ---- sparse form of the embedding gradient (data parallel) ------------------------------------------------
// rows[n, :] = dropout-masked dA[n, :]: what embed_dropout_bwd would scatter, kept as N rows so that ranks can
// exchange 4 MB of rows instead of all-reducing the dense 60 MB table gradient.
__global__ void embed_rows_kernel(const float* __restrict__ dA, float* __restrict__ rows, int N, int H, MaskSrc m) {
    int n = blockIdx.x;
    uint64_t n_total = (uint64_t)N * H;
    for (int j = threadIdx.x; j < H; j += blockDim.x)
        rows[(int64_t)n * H + j] = dA[(int64_t)n * H + j] * mask_mul1(m, (uint64_t)n * H + j, n_total);
}
int embed_rows(const float* dA, float* rows, int N, int H, MaskSrc m, cudaStream_t s) {
    if (!N) return ZRB_OK;
    embed_rows_kernel<<<N, 256, 0, s>>>(dA, rows, N, H, m);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

// dW[id, :] = sum of rows[n, :] over all n with ids[n] == id, bit-identical on every rank that holds the same
// (ids, rows) arrays: the rows are accumulated as 64-bit fixed point (2^-40 resolution, |x| < 8e6) with integer
// atomics, whose result does not depend on the order of the additions, into the slot of the id's first
// occurrence (atomicMin), then converted back to fp32 once.
__global__ void embed_first_kernel(const int64_t* __restrict__ ids, int* __restrict__ first, int n_rows, int V) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_rows) return;
    int64_t id = ids[n];
    if (id >= 0 && id < V) atomicMin(first + id, n);
}
__global__ void embed_accum_kernel(const int64_t* __restrict__ ids, const float* __restrict__ rows,
                                   const int* __restrict__ first, long long* __restrict__ acc, int n_rows, int H, int V) {
    const int n = blockIdx.x;
    const int64_t id = ids[n];
    if (id < 0 || id >= V) return;
    long long* dst = acc + (int64_t)first[id] * H;
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        float v = rows[(int64_t)n * H + j];
        if (v != 0.f) atomicAdd((unsigned long long*)(dst + j), (unsigned long long)__float2ll_rn(v * 1099511627776.f));
    }
}
__global__ void embed_finish_kernel(const int64_t* __restrict__ ids, const int* __restrict__ first,
                                    const long long* __restrict__ acc, float* __restrict__ dW, int n_rows, int H, int V) {
    const int n = blockIdx.x;
    const int64_t id = ids[n];
    if (id < 0 || id >= V || first[id] != n) return;
    for (int j = threadIdx.x; j < H; j += blockDim.x)
        dW[id * (int64_t)H + j] = (float)((double)acc[(int64_t)n * H + j] * (1.0 / 1099511627776.0));
}
int embed_scatter_rows(const int64_t* ids, const float* rows, float* dW, int n_rows, int H, int V, int* first,
                       long long* acc, cudaStream_t s) {
    if (!n_rows) return ZRB_OK;
    ZRB_CUDA(cudaMemsetAsync(first, 0x7f, (size_t)V * sizeof(int), s));           // 0x7f7f7f7f > any row index
    ZRB_CUDA(cudaMemsetAsync(acc, 0, (size_t)n_rows * H * sizeof(long long), s));
    embed_first_kernel<<<cdiv(n_rows, 256), 256, 0, s>>>(ids, first, n_rows, V);
    ZRB_KERNEL_CHECK();
    embed_accum_kernel<<<n_rows, 256, 0, s>>>(ids, rows, first, acc, n_rows, H, V);
    ZRB_KERNEL_CHECK();
    embed_finish_kernel<<<n_rows, 256, 0, s>>>(ids, first, acc, dW, n_rows, H, V);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

// ---- touched-rows-only handling of the embedding gradient (single process) ---------------------------------
// The dense [V,H] gradient is non-zero only in the <= N rows of this window's tokens.  Instead of zero-filling,
// norm-reading and updating 60 MB per step, only those rows are touched: rows of the PREVIOUS step are cleared,
// the first occurrence of every id (atomicMin table) owns the row for the norm and the update.
__global__ void embed_zero_rows_kernel(float* __restrict__ dW, const int64_t* __restrict__ ids, int n, int H, int V) {
    const int64_t id = ids[blockIdx.x];
    if (id < 0 || id >= V) return;
    for (int j = threadIdx.x; j < H; j += blockDim.x) dW[id * (int64_t)H + j] = 0.f;
}
int embed_zero_rows(float* dW, const int64_t* ids, int n, int H, int V, cudaStream_t s) {
    if (!n) return ZRB_OK;
    embed_zero_rows_kernel<<<n, 256, 0, s>>>(dW, ids, n, H, V);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}
int embed_first_table(const int64_t* ids, int* first, int n, int V, cudaStream_t s) {
    ZRB_CUDA(cudaMemsetAsync(first, 0x7f, (size_t)V * sizeof(int), s));
    if (!n) return ZRB_OK;
    embed_first_kernel<<<cdiv(n, 256), 256, 0, s>>>(ids, first, n, V);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}
// partial[blockIdx.x] = sum of squares of the rows owned by this block's tokens (first occurrences only)
__global__ void embed_rows_sumsq_kernel(const float* __restrict__ dW, const int64_t* __restrict__ ids,
                                        const int* __restrict__ first, int n, int H, int V, float* __restrict__ partial) {
    __shared__ float sh[8];
    float acc = 0.f;
    for (int t = blockIdx.x; t < n; t += gridDim.x) {
        const int64_t id = ids[t];
        if (id < 0 || id >= V || first[id] != t) continue;
        for (int j = threadIdx.x; j < H; j += blockDim.x) {
            float v = dW[id * (int64_t)H + j];
            acc += v * v;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        partial[blockIdx.x] = t;
    }
}
int embed_rows_sumsq(const float* dW, const int64_t* ids, const int* first, int n, int H, int V, float* partial,
                     int nblocks, cudaStream_t s) {
    embed_rows_sumsq_kernel<<<nblocks, 256, 0, s>>>(dW, ids, first, n, H, V, partial);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}
__global__ void embed_rows_update_kernel(float* __restrict__ W, float* __restrict__ dW, const int64_t* __restrict__ ids,
                                         const int* __restrict__ first, int n, int H, int V, float lr,
                                         const float* __restrict__ scalars, bool write_g) {
    const int t = blockIdx.x;
    const int64_t id = ids[t];
    if (id < 0 || id >= V || first[id] != t) return;
    const float coef = scalars[1];
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        float g = dW[id * (int64_t)H + j] * coef;
        if (write_g) dW[id * (int64_t)H + j] = g;
        W[id * (int64_t)H + j] -= lr * g;
    }
}
int embed_rows_update(float* W, float* dW, const int64_t* ids, const int* first, int n, int H, int V, float lr,
                      const float* scalars, bool write_g, cudaStream_t s) {
    if (!n) return ZRB_OK;
    embed_rows_update_kernel<<<n, 256, 0, s>>>(W, dW, ids, first, n, H, V, lr, scalars, write_g);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

__global__ void dropout_mask_kernel(MaskSrc m, int64_t n, uint8_t* __restrict__ out) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * g >= n) return;
    uint32_t bits = m.active ? mask_keep4(m, (uint64_t)g, (uint64_t)n) : 0xFu;
    for (int i = 0; i < 4; ++i)
        if (4 * g + i < n) out[4 * g + i] = (bits >> i) & 1u;
}
int dropout_mask_bytes(MaskSrc m, int64_t n, uint8_t* out, cudaStream_t s) {
    if (!n) return ZRB_OK;
    dropout_mask_kernel<<<cdiv((n + 3) / 4, 256), 256, 0, s>>>(m, n, out);
    ZRB_KERNEL_CHECK();
    return ZRB_OK;
}

}  // namespace zrb
