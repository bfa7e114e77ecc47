// vg_opq.cu — OptimizedProductQuantizer.Train (internal/quantization/opq.go:89-193) with the
// Procrustes / one-sided Jacobi SVD of internal/quantization/svd.go, on the device.
//
// Alternating optimisation per outer iteration:
//   1. rotate every vector by the block-diagonal R (rotateVector, opq.go:196-214: simd.Dot per row)
//   2. ProductQuantizer.Train on the rotated vectors (vg_kmeans.cu: dev_pq_train)
//   3. y^ = Decode(Encode(rotated)); M_b = sum_i x_{i,b}^T y^_{i,b}, accumulated in SAMPLE ORDER with
//      separately rounded multiply and add (Go does not fuse on amd64) — one thread per matrix entry
//      walks the samples, 24 x 32 x 32 independent order-exact chains for 768-d / 96 subspaces
//   4. R_b = U V^T from the Jacobi SVD of M_b (tol 1e-5, <= 100 sweeps, reflection fixed on the
//      smallest singular value).  The sweep order (i,j) and every float32 rounding are the
//      reference's; one warp per block: lane 0 owns the sequential reductions, all lanes apply
//      the plane rotations (independent per row).
#include <vector>

#include "vg_kmeans.cuh"
#include "vg_quant.cuh"

namespace vg {

// rotateVector: out[r][b*bs + i] = simd.Dot(R_b[i], v_b) in AVX-512 order (half-warp per output).
__global__ void __launch_bounds__(256) opq_rotate_kernel(const float *v, int64_t n, int64_t dim, int bs, const float *rot, float *out) {
    const int64_t hwid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    const int64_t total = n * dim;
    const bool live = hwid < total;
    const int64_t o = live ? hwid : total - 1;
    const int64_t r = o / dim, d = o - r * dim;
    const int64_t b = d / bs, i = d - b * bs;
    const float *row = rot + (b * bs + i) * bs;
    const float *x = v + r * dim + b * bs;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const int epochs = bs >> 6;
    for (int e = 0; e < epochs; e++)
#pragma unroll
        for (int j = 0; j < 4; j++) a[j] = __fmaf_rn(row[e * 64 + j * 16 + lane], x[e * 64 + j * 16 + lane], a[j]);
    float tot = reduce16(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])));
    if (lane == 0 && live) {
        for (int t = epochs * 64; t < bs; t++) tot = __fmaf_rn(row[t], x[t], tot);
        out[o] = tot;
    }
}
vg_status dev_opq_rotate(const float *d_v, int64_t n, int64_t dim, int bs, const float *d_rot, float *d_out, cudaStream_t st) {
    const int64_t threads = n * dim * 16;
    if (threads <= 0) return VG_OK;
    opq_rotate_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_v, n, dim, bs, d_rot, d_out);
    VG_LAUNCHED();
    return VG_OK;
}

// Decode's inverse rotation (opq.go:244-262): dst[b*bs+i] = sum_j R_b[j][i] * src[b*bs+j], sequential, unfused.
__global__ void __launch_bounds__(256) opq_unrotate_kernel(const float *src, int64_t n, int64_t dim, int bs, const float *rot, float *dst) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n * dim) return;
    const int64_t r = o / dim, d = o - r * dim;
    const int64_t b = d / bs, i = d - b * bs;
    const float *x = src + r * dim + b * bs;
    float sum = 0.0f;
    for (int j = 0; j < bs; j++) sum = __fadd_rn(sum, __fmul_rn(rot[(b * bs + j) * bs + i], x[j]));
    dst[o] = sum;
}

// M[b][r][c] = sum over samples i (in order) of fl(x[i][b*bs+r] * y[i][b*bs+c])   (opq.go:160-178)
__global__ void __launch_bounds__(256) opq_accum_kernel(const float *x, const float *y, int64_t n, int64_t dim, int bs, int64_t total,
                                                        float *M) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t b = idx / ((int64_t)bs * bs);
    const int r = (int)((idx / bs) % bs), c = (int)(idx % bs);
    const float *xp = x + b * bs + r, *yp = y + b * bs + c;
    float acc = 0.0f;
    int64_t i = 0;
    for (; i + 4 <= n; i += 4) {
        float xv[4], yv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            xv[u] = __ldg(xp + (i + u) * dim);
            yv[u] = __ldg(yp + (i + u) * dim);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc = __fadd_rn(acc, __fmul_rn(xv[u], yv[u]));
    }
    for (; i < n; i++) acc = __fadd_rn(acc, __fmul_rn(__ldg(xp + i * dim), __ldg(yp + i * dim)));
    M[idx] = acc;
}

// computeProcrustesRotation (svd.go:129-182) for one block per warp.  work: [blocks][3][n*n] scratch (U, V, T).
__global__ void __launch_bounds__(32) opq_procrustes_kernel(const float *M_all, int n, float *work, float *sig_all, float *R_all) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int nn = n * n;
    float *u = work + (size_t)b * 3 * nn, *v = u + nn, *t = v + nn;
    float *sigma = sig_all + (size_t)b * n;
    float *R = R_all + (size_t)b * nn;
    for (int i = lane; i < nn; i += 32) {
        u[i] = M_all[(size_t)b * nn + i];
        v[i] = (i / n == i % n) ? 1.0f : 0.0f;
    }
    __syncwarp();
    // ---- svd: one-sided Jacobi (svd.go:13-127)
    for (int iter = 0; iter < 100; iter++) {
        int changed = 0;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                float c = 0.f, s = 0.f;
                int rot = 0;
                if (lane == 0) {
                    float alpha = 0.f, beta = 0.f, gamma = 0.f;
                    for (int k = 0; k < n; k++) {
                        const float ui = u[k * n + i], uj = u[k * n + j];
                        alpha = __fadd_rn(alpha, __fmul_rn(ui, ui));
                        beta = __fadd_rn(beta, __fmul_rn(uj, uj));
                        gamma = __fadd_rn(gamma, __fmul_rn(ui, uj));
                    }
                    if (!(alpha < 1e-12f || beta < 1e-12f)) {
                        const float ab = __fmul_rn(alpha, beta);
                        if (!(fabs((double)gamma) < 1e-5 * sqrt((double)ab))) {
                            rot = 1;
                            const float zeta = __fdiv_rn(__fsub_rn(beta, alpha), __fmul_rn(2.0f, gamma));
                            const float rt = (float)sqrt((double)__fadd_rn(1.0f, __fmul_rn(zeta, zeta)));
                            float tt;
                            if (zeta > 0.0f) tt = __fdiv_rn(1.0f, __fadd_rn(zeta, rt));
                            else tt = __fdiv_rn(-1.0f, __fadd_rn(-zeta, rt));
                            c = __fdiv_rn(1.0f, (float)sqrt((double)__fadd_rn(1.0f, __fmul_rn(tt, tt))));
                            s = __fmul_rn(c, tt);
                        }
                    }
                }
                rot = __shfl_sync(0xffffffffu, rot, 0);
                if (rot) {
                    c = __shfl_sync(0xffffffffu, c, 0);
                    s = __shfl_sync(0xffffffffu, s, 0);
                    changed = 1;
                    for (int k = lane; k < n; k += 32) {  // rows are independent
                        float t1 = u[k * n + i], t2 = u[k * n + j];
                        u[k * n + i] = __fsub_rn(__fmul_rn(c, t1), __fmul_rn(s, t2));
                        u[k * n + j] = __fadd_rn(__fmul_rn(s, t1), __fmul_rn(c, t2));
                        t1 = v[k * n + i];
                        t2 = v[k * n + j];
                        v[k * n + i] = __fsub_rn(__fmul_rn(c, t1), __fmul_rn(s, t2));
                        v[k * n + j] = __fadd_rn(__fmul_rn(s, t1), __fmul_rn(c, t2));
                    }
                }
                __syncwarp();
            }
        if (!changed) break;
    }
    // singular values = column norms of U; normalise the columns
    for (int j = lane; j < n; j += 32) {
        float sum = 0.f;
        for (int i = 0; i < n; i++) sum = __fadd_rn(sum, __fmul_rn(u[i * n + j], u[i * n + j]));
        const float sg = (float)sqrt((double)sum);
        sigma[j] = sg;
        if (sg > 1e-10f) {
            const float inv = __fdiv_rn(1.0f, sg);
            for (int i = 0; i < n; i++) u[i * n + j] = __fmul_rn(u[i * n + j], inv);
        }
    }
    __syncwarp();
    int mi = 0;
    {
        float ms = sigma[0];
        for (int i = 1; i < n; i++)
            if (sigma[i] < ms) {
                ms = sigma[i];
                mi = i;
            }
    }
    for (int pass = 0; pass < 2; pass++) {
        // R = U V^T
        for (int idx = lane; idx < nn; idx += 32) {
            const int i = idx / n, j = idx - i * n;
            float sum = 0.f;
            for (int k = 0; k < n; k++) sum = __fadd_rn(sum, __fmul_rn(u[i * n + k], v[j * n + k]));
            R[idx] = sum;
        }
        __syncwarp();
        if (pass == 1) break;
        // determinant (svd.go:184-216), Gaussian elimination with partial pivoting; row updates are independent
        for (int idx = lane; idx < nn; idx += 32) t[idx] = R[idx];
        __syncwarp();
        float det = 1.0f;
        bool zero = false;
        for (int i = 0; i < n && !zero; i++) {
            int pivot = i;
            for (int j = i + 1; j < n; j++)
                if (fabs((double)t[j * n + i]) > fabs((double)t[pivot * n + i])) pivot = j;
            if (pivot != i) {
                for (int k = lane; k < n; k += 32) {
                    const float tmp = t[i * n + k];
                    t[i * n + k] = t[pivot * n + k];
                    t[pivot * n + k] = tmp;
                }
                det = __fmul_rn(det, -1.0f);
                __syncwarp();
            }
            const float piv = t[i * n + i];
            if (piv == 0.0f) {
                det = 0.0f;
                zero = true;
                break;
            }
            det = __fmul_rn(det, piv);
            for (int j = i + 1 + lane; j < n; j += 32) {
                const float factor = __fdiv_rn(t[j * n + i], piv);
                for (int k = i + 1; k < n; k++) t[j * n + k] = __fsub_rn(t[j * n + k], __fmul_rn(factor, t[i * n + k]));
            }
            __syncwarp();
        }
        if (!(det < 0.0f)) break;
        for (int i = lane; i < n; i += 32) u[i * n + mi] = __fmul_rn(u[i * n + mi], -1.0f);
        __syncwarp();
    }
}

vg_status dev_opq_procrustes(const float *d_M, int blocks, int bs, float *d_R, float *d_sigma, cudaStream_t st) {
    DevBuf work, sig;
    VG_TRY(work.alloc((size_t)blocks * 3 * bs * bs * 4));
    float *sp = d_sigma;
    if (!sp) {
        VG_TRY(sig.alloc((size_t)blocks * bs * 4));
        sp = sig.as<float>();
    }
    opq_procrustes_kernel<<<blocks, 32, 0, st>>>(d_M, bs, work.as<float>(), sp, d_R);
    VG_LAUNCHED();
    return VG_OK;
}

static int64_t opq_block_size(int64_t dim, int64_t m) {  // NewOptimizedProductQuantizer, opq.go:41-58
    const int64_t sub = dim / m;
    int64_t bs = dim;
    if (dim > 64) {
        int64_t best = 1000;
        for (int64_t b = sub; b <= dim; b += sub)
            if (dim % b == 0) {
                const int64_t diff = b > 32 ? b - 32 : 32 - b;
                if (diff < best) {
                    best = diff;
                    bs = b;
                }
            }
    }
    return bs;
}

}  // namespace vg

using namespace vg;

extern "C" {

vg_status vg_opq_block_size(int64_t dim, int64_t m, int64_t *block_size) {
    if (dim <= 0 || m <= 0 || dim % m != 0) return fail(VG_ERR_INVALID, "dimension must be divisible by numSubvectors");
    *block_size = opq_block_size(dim, m);
    return VG_OK;
}

vg_status vg_opq_rotate(const float *h_vecs, int64_t n, int64_t dim, int64_t block, const float *h_rotations, int32_t inverse, float *h_out) {
    VG_ENTER();
    if (n <= 0) return VG_OK;
    if (block <= 0 || dim % block != 0) return fail(VG_ERR_INVALID, "OPQ needs block rotations with dim % block == 0");
    cudaStream_t st = stream();
    DevBuf v, rot, out;
    VG_TRY(v.alloc((size_t)n * dim * 4));
    VG_TRY(rot.alloc((size_t)dim * block * 4));
    VG_TRY(out.alloc((size_t)n * dim * 4));
    VG_TRY(staged_h2d(v.p, h_vecs, (size_t)n * dim * 4));
    VG_TRY(staged_h2d(rot.p, h_rotations, (size_t)dim * block * 4));
    if (inverse) {
        opq_unrotate_kernel<<<(unsigned)((n * dim + 255) / 256), 256, 0, st>>>(v.as<float>(), n, dim, (int)block, rot.as<float>(), out.as<float>());
        VG_LAUNCHED();
    } else {
        VG_TRY(dev_opq_rotate(v.as<float>(), n, dim, (int)block, rot.as<float>(), out.as<float>(), st));
    }
    VG_CUDA(cudaStreamSynchronize(st));
    return staged_d2h(h_out, out.p, (size_t)n * dim * 4);
}

vg_status vg_opq_procrustes(const float *h_M, int64_t blocks, int64_t n, float *h_R, float *h_sigma) {
    VG_ENTER();
    if (blocks <= 0 || n <= 0) return fail(VG_ERR_INVALID, "procrustes requires square matrix");
    cudaStream_t st = stream();
    DevBuf M, R, S;
    VG_TRY(M.alloc((size_t)blocks * n * n * 4));
    VG_TRY(R.alloc((size_t)blocks * n * n * 4));
    VG_TRY(S.alloc((size_t)blocks * n * 4));
    VG_TRY(staged_h2d(M.p, h_M, (size_t)blocks * n * n * 4));
    VG_TRY(dev_opq_procrustes(M.as<float>(), (int)blocks, (int)n, R.as<float>(), S.as<float>(), st));
    VG_CUDA(cudaStreamSynchronize(st));
    VG_TRY(staged_d2h(h_R, R.p, (size_t)blocks * n * n * 4));
    if (h_sigma) VG_TRY(staged_d2h(h_sigma, S.p, (size_t)blocks * n * 4));
    return VG_OK;
}

vg_status vg_opq_train(const float *h_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, int64_t opq_iters, int64_t pq_iters,
                       uint64_t seed, float *h_rotations, int8_t *h_codebooks, float *h_scales, float *h_offsets) {
    VG_ENTER();
    if (n <= 0) return fail(VG_ERR_INVALID, "no vectors provided for training");
    if (m <= 0 || dim <= 0 || dim % m != 0) return fail(VG_ERR_INVALID, "dimension must be divisible by numSubvectors");
    if (k <= 0 || k > 256) return fail(VG_ERR_INVALID, "numCentroids must be <= 256 for uint8 encoding");
    cudaStream_t st = stream();
    const int bs = (int)opq_block_size(dim, m);
    const int blocks = (int)(dim / bs);
    const int ds = (int)(dim / m);
    DevBuf x, xr, y, codes, rot, M, cent, cb, sc, of;
    VG_TRY(x.alloc((size_t)n * dim * 4));
    VG_TRY(staged_h2d(x.p, h_vecs, (size_t)n * dim * 4));
    VG_TRY(xr.alloc((size_t)n * dim * 4));
    VG_TRY(y.alloc((size_t)n * dim * 4));
    VG_TRY(codes.alloc((size_t)n * m));
    VG_TRY(rot.alloc((size_t)blocks * bs * bs * 4));
    VG_TRY(M.alloc((size_t)blocks * bs * bs * 4));
    {
        std::vector<float> id((size_t)blocks * bs * bs, 0.0f);
        for (int b = 0; b < blocks; b++)
            for (int i = 0; i < bs; i++) id[((size_t)b * bs + i) * bs + i] = 1.0f;
        VG_TRY(staged_h2d(rot.p, id.data(), id.size() * 4));
    }
    for (int64_t it = 0; it < opq_iters; it++) {
        VG_TRY(dev_opq_rotate(x.as<float>(), n, dim, bs, rot.as<float>(), xr.as<float>(), st));
        VG_TRY(dev_pq_train(xr.as<float>(), n, dim, m, k, pq_iters, seed + (uint64_t)it, cent, cb, sc, of, st));
        VG_TRY(dev_pq_encode(xr.as<float>(), n, dim, (int)m, (int)k, cb.as<int8_t>(), sc.as<float>(), of.as<float>(), codes.as<uint8_t>(), st));
        VG_TRY(dev_pq_decode(codes.as<uint8_t>(), n, dim, (int)m, (int)k, cb.as<int8_t>(), sc.as<float>(), of.as<float>(), y.as<float>(), st));
        const int64_t total = (int64_t)blocks * bs * bs;
        opq_accum_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x.as<float>(), y.as<float>(), n, dim, bs, total, M.as<float>());
        VG_LAUNCHED();
        VG_TRY(dev_opq_procrustes(M.as<float>(), blocks, bs, rot.as<float>(), nullptr, st));
    }
    VG_CUDA(cudaStreamSynchronize(st));
    VG_TRY(staged_d2h(h_rotations, rot.p, (size_t)blocks * bs * bs * 4));
    if (opq_iters > 0) {
        VG_TRY(staged_d2h(h_codebooks, cb.p, (size_t)m * k * ds));
        VG_TRY(staged_d2h(h_scales, sc.p, (size_t)m * 4));
        VG_TRY(staged_d2h(h_offsets, of.p, (size_t)m * 4));
    }
    return VG_OK;
}

}  // extern "C"
