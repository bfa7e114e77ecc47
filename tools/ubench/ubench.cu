// Micro-benchmarks that decide kernel design (run on the B200 box):
//   1. FP32 issue rate: FFMA (3-reg) vs packed fma.rn.f32x2 / add.rn.f32x2 (sm_100+)
//   2. shared-memory random lookup rate for 4/8/16-byte entries (PQ ADC tables)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) fp_kernel(float *out, int iters, float seed) {
    // 16 independent chains per thread
    float a[16];
    float2 b[8];
    for (int i = 0; i < 16; i++) a[i] = seed + i + threadIdx.x;
    for (int i = 0; i < 8; i++) b[i] = make_float2(seed + i, seed - i + threadIdx.x);
    float x = seed * 0.5f, y = seed * 0.25f;
    float2 x2 = make_float2(x, y), y2 = make_float2(y, x);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {  // scalar FFMA 3-reg: 16 per iter
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __fmaf_rn(a[i], x, y);
        } else if (MODE == 1) {  // packed FFMA2: 8 per iter = 16 flop-lanes
#pragma unroll
            for (int i = 0; i < 8; i++) b[i] = ffma2(b[i], x2, y2);
        } else if (MODE == 2) {  // scalar: FADD + FFMA (the scan inner loop): e = q + nrec; acc = fma(e,e,acc)
#pragma unroll
            for (int i = 0; i < 16; i++) { float e = __fadd_rn(x, a[i] ); a[i] = __fmaf_rn(e, e, a[i]); }
        } else if (MODE == 3) {  // packed: FADD2 + FFMA2
#pragma unroll
            for (int i = 0; i < 8; i++) { float2 e = fadd2(x2, b[i]); b[i] = ffma2(e, e, b[i]); }
        } else if (MODE == 4) {  // packed FADD2 only
#pragma unroll
            for (int i = 0; i < 8; i++) b[i] = fadd2(b[i], x2);
        } else if (MODE == 5) {  // scalar FADD only
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = __fadd_rn(a[i], x);
        }
    }
    float s = 0;
    for (int i = 0; i < 16; i++) s += a[i];
    for (int i = 0; i < 8; i++) s += b[i].x + b[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory lookups: table of ENT entries of W bytes, random indices from a per-thread LCG
template <int W>
__global__ void __launch_bounds__(256) lds_kernel(float *out, int iters, int ent) {
    extern __shared__ __align__(16) unsigned char sm[];
    float *t = reinterpret_cast<float *>(sm);
    for (int i = threadIdx.x; i < ent * (W / 4); i += blockDim.x) t[i] = (float)(i & 7);
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    const int groups = ent / 1024;  // 4 subspaces x 256 centroids per group
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            s = s * 1664525u + 1013904223u;
            const uint32_t g = (it * 2 + u) % groups;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t idx = g * 1024 + b * 256 + ((s >> (8 * b)) & 255u);
                if (W == 4) acc0 += t[idx];
                else if (W == 8) { float2 v = reinterpret_cast<const float2 *>(t)[idx]; acc0 += v.x; acc1 += v.y; }
                else { float4 v = reinterpret_cast<const float4 *>(t)[idx]; acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w; }
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, clk);
    float *out;
    CK(cudaMalloc(&out, 148 * 8 * 256 * 4 * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = p.multiProcessorCount * 8, iters = 20000;
    const char *names[] = {"FFMA scalar", "FFMA2 packed", "FADD+FFMA scalar", "FADD2+FFMA2 packed", "FADD2 packed", "FADD scalar"};
    const double lanes_per_iter[] = {16, 16, 32, 32, 16, 16};
    for (int mode = 0; mode < 6; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            switch (mode) {
                case 0: fp_kernel<0><<<blocks, 256>>>(out, iters, 1.0f); break;
                case 1: fp_kernel<1><<<blocks, 256>>>(out, iters, 1.0f); break;
                case 2: fp_kernel<2><<<blocks, 256>>>(out, iters, 1.0f); break;
                case 3: fp_kernel<3><<<blocks, 256>>>(out, iters, 1.0f); break;
                case 4: fp_kernel<4><<<blocks, 256>>>(out, iters, 1.0f); break;
                case 5: fp_kernel<5><<<blocks, 256>>>(out, iters, 1.0f); break;
            }
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 1) {
                double ops = (double)blocks * 256 * iters * lanes_per_iter[mode];
                printf("%-22s %8.3f ms  %7.2f T lane-ops/s  = %6.1f lane-ops/clk/SM (at 1.965 GHz)\n", names[mode], ms, ops / ms / 1e9,
                       ops / (ms * 1e-3) / p.multiProcessorCount / 1.965e9);
            }
        }
    }
    // LDS lookups
    for (int w = 4; w <= 16; w *= 2) {
        for (int ent : {256 * 16, 256 * 32}) {
            size_t bytes = (size_t)ent * w;
            if (bytes > 200 * 1024) continue;
            const int lb = p.multiProcessorCount * 1, it2 = 4000;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (w == 4) { cudaFuncSetAttribute(lds_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); lds_kernel<4><<<lb, 256, bytes>>>(out, it2, ent); }
                if (w == 8) { cudaFuncSetAttribute(lds_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); lds_kernel<8><<<lb, 256, bytes>>>(out, it2, ent); }
                if (w == 16) { cudaFuncSetAttribute(lds_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); lds_kernel<16><<<lb, 256, bytes>>>(out, it2, ent); }
                cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1));
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep == 1) {
                    double lk = (double)lb * 256 * it2 * 8;
                    printf("LDS.%-3d random, %6d entries (%3zu KB), 256 thr x1 CTA/SM: %7.3f ms  %6.2f lookups/clk/SM  %6.1f B/clk/SM\n", w * 8, ent,
                           bytes / 1024, ms, lk / (ms * 1e-3) / p.multiProcessorCount / 1.965e9, lk * w / (ms * 1e-3) / p.multiProcessorCount / 1.965e9);
                }
            }
        }
    }
    CK(cudaGetLastError());
    return 0;
}
