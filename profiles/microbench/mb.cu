// Round-2 design micro-benchmarks for the persistent training kernel (profiles/README.md).
// Measures on one B200: cluster barrier latency, all-gather of a 100x128 fp32 activation
// through L2 vs through DSMEM inside a 16-CTA cluster, mma.sync tf32 / FFMA / FFMA2 issue
// rates per SM, and single-CTA L2 read bandwidth.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

__global__ void k_cluster_barrier(int iters, long long* out) {
    cluster_sync_all();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) cluster_sync_all();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0);
}

// all-gather through global memory (L2): CTA c writes its [100 x 8] slice, barrier, reads all
__global__ void k_gather_l2(int iters, float* buf, long long* out, float* sink) {
    extern __shared__ float sm[];
    const int nct = gridDim.x, c = blockIdx.x, t = threadIdx.x, nt = blockDim.x;
    float acc = 0.f;
    cluster_sync_all();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float* b = buf + (it & 1) * 12800;
        // write own slice: 100 rows x 8 cols = 800 floats = 200 float4
        for (int i = t; i < 200; i += nt) {
            int r = i >> 1, h = i & 1;
            float4 v = make_float4(it + c, r, h, acc);
            *reinterpret_cast<float4*>(b + r * 128 + c * 8 + h * 4) = v;
        }
        cluster_sync_all();
        for (int i = t; i < 3200; i += nt) {
            float4 v = __ldcg(reinterpret_cast<const float4*>(b) + i);
            reinterpret_cast<float4*>(sm)[i] = v;
        }
        __syncthreads();
        acc += sm[(t * 7 + it) % 12800];
    }
    cluster_sync_all();
    long long t1 = clock64();
    if (t == 0 && c == 0) out[0] = t1 - t0;
    sink[c * nt + t] = acc;
    (void)nct;
}

// all-gather through DSMEM: CTA c pushes its slice into every CTA's shared memory
__global__ void k_gather_dsmem(int iters, long long* out, float* sink) {
    extern __shared__ float sm[];   // 2 x 12800 floats
    cg::cluster_group cl = cg::this_cluster();
    const int nct = cl.num_blocks(), c = cl.block_rank(), t = threadIdx.x, nt = blockDim.x;
    float acc = 0.f;
    cluster_sync_all();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float* b = sm + (it & 1) * 12800;
        for (int i = t; i < 200 * nct; i += nt) {
            int dst = i / 200, j = i - dst * 200;
            int r = j >> 1, h = j & 1;
            float* rb = cl.map_shared_rank(b, (dst + c) % nct);
            *reinterpret_cast<float4*>(rb + r * 128 + c * 8 + h * 4) = make_float4(it + c, r, h, acc);
        }
        cluster_sync_all();
        acc += b[(t * 7 + it) % 12800];
    }
    cluster_sync_all();
    long long t1 = clock64();
    if (t == 0 && c == 0) out[0] = t1 - t0;
    sink[c * nt + t] = acc;
}

// reduce-scatter through L2: every CTA writes a full [100x128] partial, barrier, CTA c sums its 8 columns over 16 partials
__global__ void k_rs_l2(int iters, float* buf, long long* out, float* sink) {
    const int nct = gridDim.x, c = blockIdx.x, t = threadIdx.x, nt = blockDim.x;
    float acc = 0.f;
    cluster_sync_all();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float* b = buf + (size_t)(it & 1) * 12800 * nct + (size_t)c * 12800;
        for (int i = t; i < 3200; i += nt)
            reinterpret_cast<float4*>(b)[i] = make_float4(i, it, c, acc);
        cluster_sync_all();
        const float* base = buf + (size_t)(it & 1) * 12800 * nct;
        for (int i = t; i < 200; i += nt) {
            int r = i >> 1, h = i & 1;
            float4 s = make_float4(0, 0, 0, 0);
            for (int q = 0; q < nct; ++q) {
                float4 v = __ldcg(reinterpret_cast<const float4*>(base + (size_t)q * 12800 + r * 128 + c * 8 + h * 4));
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            acc += s.x + s.y + s.z + s.w;
        }
    }
    cluster_sync_all();
    long long t1 = clock64();
    if (t == 0 && c == 0) out[0] = t1 - t0;
    sink[c * nt + t] = acc;
}

__global__ void k_mma_tf32(int iters, long long* out, float* sink) {
    float d[4][4] = {};
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    float s = 0; for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += d[j][q];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mma_bf16(int iters, long long* out, float* sink) {
    float d[4][4] = {};
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    float s = 0; for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += d[j][q];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma(int iters, long long* out, float* sink, float x, float y) {
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = threadIdx.x + j;
    float xs[4] = {x, x + 1, x + 2, x + 3};
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], xs[j & 3], y);
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    float s = 0; for (int j = 0; j < 16; ++j) s += a[j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(int iters, long long* out, float* sink, float x, float y) {
    unsigned long long a[8];
    float2 xv = make_float2(x, x + 1), yv = make_float2(y, y);
    unsigned long long xs, ys;
    xs = *reinterpret_cast<unsigned long long*>(&xv);
    ys = *reinterpret_cast<unsigned long long*>(&yv);
#pragma unroll
    for (int j = 0; j < 8; ++j) { float2 v = make_float2(threadIdx.x + j, j); a[j] = *reinterpret_cast<unsigned long long*>(&v); }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;\n" : "+l"(a[j]) : "l"(xs), "l"(ys));
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    float s = 0; for (int j = 0; j < 8; ++j) { float2 v = *reinterpret_cast<float2*>(&a[j]); s += v.x + v.y; }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// smem-operand FFMA: typical register-tiled GEMM inner loop, 8x4 micro-tile, operands from shared memory
__global__ void k_gemm_inner(int iters, long long* out, float* sink) {
    __shared__ float As[128 * 32];
    __shared__ float Bs[32 * 128];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) { As[i] = i * 0.001f; Bs[i] = i * 0.002f; }
    __syncthreads();
    float acc[8][4] = {};
    const int tr = (threadIdx.x >> 5) * 8 % 128, tc = (threadIdx.x & 31) * 4;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            float4 b = *reinterpret_cast<const float4*>(&Bs[k * 128 + tc]);
            float a[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) a[r] = As[(tr + r) * 32 + k];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                acc[r][0] = fmaf(a[r], b.x, acc[r][0]); acc[r][1] = fmaf(a[r], b.y, acc[r][1]);
                acc[r][2] = fmaf(a[r], b.z, acc[r][2]); acc[r][3] = fmaf(a[r], b.w, acc[r][3]);
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    float s = 0; for (int r = 0; r < 8; ++r) for (int q = 0; q < 4; ++q) s += acc[r][q];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_l2_read(int iters, const float4* src, int n4, long long* out, float* sink) {
    float4 s = make_float4(0, 0, 0, 0);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int i = threadIdx.x; i < n4; i += blockDim.x * 4) {
            float4 v0 = __ldcg(src + i);
            float4 v1 = (i + blockDim.x < n4) ? __ldcg(src + i + blockDim.x) : s;
            float4 v2 = (i + 2 * blockDim.x < n4) ? __ldcg(src + i + 2 * blockDim.x) : s;
            float4 v3 = (i + 3 * blockDim.x < n4) ? __ldcg(src + i + 3 * blockDim.x) : s;
            s.x += v0.x + v1.x + v2.x + v3.x; s.y += v0.y + v1.y + v2.y + v3.y;
            s.z += v0.z + v1.z + v2.z + v3.z; s.w += v0.w + v1.w + v2.w + v3.w;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s.x + s.y + s.z + s.w;
}

template <typename K, typename... A>
static int launch_cluster(K kern, int nct, int csz, int threads, size_t smem, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nct); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (csz > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ncl = -1;
    cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
    printf("  [cluster %d x %d thr, smem %zu] max active clusters %d\n", csz, threads, smem, ncl);
    CK(cudaLaunchKernelEx(&cfg, kern, args...));
    CK(cudaDeviceSynchronize());
    return 0;
}

int main() {
    long long* out; float* sink; float* buf;
    CK(cudaMallocManaged(&out, 64));
    CK(cudaMalloc(&sink, 148 * 1024 * 4 * 4));
    CK(cudaMalloc(&buf, 2 * 12800 * 16 * 4 + 1024 * 1024 * 4));
    CK(cudaMemset(buf, 0, 2 * 12800 * 16 * 4 + 1024 * 1024 * 4));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    const int IT = 2000;
    for (int csz : {2, 4, 8, 16}) for (int thr : {256, 512}) {
        if (launch_cluster(k_cluster_barrier, csz, csz, thr, 0, IT, out)) return 1;
        if (launch_cluster(k_cluster_barrier, csz, csz, thr, 0, IT, out)) return 1;
        printf("cluster_barrier csz=%d thr=%d: %.1f cycles/iter\n", csz, thr, (double)out[0] / IT);
    }
    // big-smem 16-cluster feasibility
    for (size_t smem : {(size_t)100 * 1024, (size_t)200 * 1024, (size_t)227 * 1024}) {
        if (launch_cluster(k_gather_dsmem, 16, 16, 512, smem, 10, out, sink)) printf("  launch failed at smem %zu\n", smem);
    }
    for (int csz : {8, 16}) for (int thr : {256, 512}) {
        launch_cluster(k_gather_l2, csz, csz, thr, 12800 * 4, IT, buf, out, sink);
        launch_cluster(k_gather_l2, csz, csz, thr, 12800 * 4, IT, buf, out, sink);
        printf("allgather_L2 [100x128 f32] csz=%d thr=%d: %.1f cycles/iter\n", csz, thr, (double)out[0] / IT);
        launch_cluster(k_gather_dsmem, csz, csz, thr, 2 * 12800 * 4, IT, out, sink);
        launch_cluster(k_gather_dsmem, csz, csz, thr, 2 * 12800 * 4, IT, out, sink);
        printf("allgather_DSMEM [100x128 f32] csz=%d thr=%d: %.1f cycles/iter\n", csz, thr, (double)out[0] / IT);
        launch_cluster(k_rs_l2, csz, csz, thr, 0, IT, buf, out, sink);
        launch_cluster(k_rs_l2, csz, csz, thr, 0, IT, buf, out, sink);
        printf("reducescatter_L2 [100x128 f32 partials] csz=%d thr=%d: %.1f cycles/iter\n", csz, thr, (double)out[0] / IT);
    }
    for (int thr : {128, 256, 512, 1024}) {
        int nw = thr / 32;
        k_mma_tf32<<<16, thr>>>(IT, out, sink); CK(cudaDeviceSynchronize());
        k_mma_tf32<<<16, thr>>>(IT, out, sink); CK(cudaDeviceSynchronize());
        printf("mma.sync m16n8k8 tf32 thr=%d: %.2f cycles per (4 MMA/warp); %.1f tf32 MAC/clk/SM\n", thr,
               (double)out[0] / IT, 4.0 * nw * 16 * 8 * 8 / ((double)out[0] / IT));
        k_mma_bf16<<<16, thr>>>(IT, out, sink); CK(cudaDeviceSynchronize());
        k_mma_bf16<<<16, thr>>>(IT, out, sink); CK(cudaDeviceSynchronize());
        printf("mma.sync m16n8k16 bf16 thr=%d: %.1f bf16 MAC/clk/SM\n", thr,
               4.0 * nw * 16 * 8 * 16 / ((double)out[0] / IT));
        k_ffma<<<16, thr>>>(IT, out, sink, 1.0001f, 0.5f); CK(cudaDeviceSynchronize());
        k_ffma<<<16, thr>>>(IT, out, sink, 1.0001f, 0.5f); CK(cudaDeviceSynchronize());
        printf("FFMA thr=%d: %.1f FMA/clk/SM\n", thr, 16.0 * thr / ((double)out[0] / IT));
        k_ffma2<<<16, thr>>>(IT, out, sink, 1.0001f, 0.5f); CK(cudaDeviceSynchronize());
        k_ffma2<<<16, thr>>>(IT, out, sink, 1.0001f, 0.5f); CK(cudaDeviceSynchronize());
        printf("FFMA2 thr=%d: %.1f FMA/clk/SM\n", thr, 16.0 * thr / ((double)out[0] / IT));
        k_gemm_inner<<<16, thr>>>(200, out, sink); CK(cudaDeviceSynchronize());
        k_gemm_inner<<<16, thr>>>(200, out, sink); CK(cudaDeviceSynchronize());
        printf("smem-operand 8x4 FFMA tile thr=%d: %.1f FMA/clk/SM\n", thr, 200.0 * 32 * 32 * thr / (double)out[0]);
    }
    for (int nct : {1, 16, 148}) for (int thr : {256, 512, 1024}) {
        int n4 = 360 * 1024 / 16;
        k_l2_read<<<nct, thr>>>(50, (const float4*)buf, n4, out, sink); CK(cudaDeviceSynchronize());
        k_l2_read<<<nct, thr>>>(50, (const float4*)buf, n4, out, sink); CK(cudaDeviceSynchronize());
        printf("L2 read 360KB ctas=%d thr=%d: %.1f B/clk/SM\n", nct, thr, 50.0 * 360 * 1024 / (double)out[0]);
    }
    return 0;
}
