// Microbenchmark behind DESIGN.md section 8: what the ACCESS PATTERN of the synthesis epilogue costs at equal occupancy.
// Every variant streams the same tensors -- C [R x W] read and rewritten in place, C2 [R x W] written -- from one persistent CTA of
// 512 threads per SM, with no arithmetic to speak of:
//   flat      grid-stride float2 over the flat array (what an elementwise kernel does)
//   quad      the current tc_rowgemm.cuh epilogue: 128-row x 256-column tiles, a warp instruction = 8 rows x 32 bytes
//             (tcgen05.ld 16x256b ownership), 16 columns per warp and round, loads of a round issued before its stores
//   rowsweep  128-row x NT-column tiles, a warp instruction = 256 contiguous bytes of ONE row (what a shared-memory transposed
//             epilogue would issue), rows of a tile dealt round-robin to the 16 warps
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/epi tools/microbench/epilogue_patterns.cu && /tmp/epi
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(512, 1) flat_kernel(float* __restrict__ C, float* __restrict__ C2, long n2) {
    float2* c = reinterpret_cast<float2*>(C);
    float2* d = reinterpret_cast<float2*>(C2);
    for (long i = (long)blockIdx.x * 512 + threadIdx.x; i < n2; i += (long)gridDim.x * 512) {
        float2 v = c[i];
        v.x += 1.f; v.y += 1.f;
        c[i] = v;
        d[i] = make_float2(v.x * 2.f, v.y * 2.f);
    }
}

template <int STREAMS, int DEPTH>
__global__ void __launch_bounds__(512, 1) quad_kernel(float* __restrict__ C, float* __restrict__ C2, long R, int W) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2;
    const int ntn = (W + 255) / 256;
    const long tiles = ((R + 127) / 128) * ntn;
    for (long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long rt = t / ntn;
        const int n_base = (int)(t - rt * ntn) * 256;
        for (int ci0 = half; ci0 < 16; ci0 += 4 * DEPTH) {
            float2 v[DEPTH][8];
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) {
                const int c0 = n_base + (ci0 + 4 * d) * 16 + 2 * (lane & 3);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const long row = rt * 128 + q * 32 + (e >> 1) * 8 + (lane >> 2);
                    v[d][e] = *reinterpret_cast<const float2*>(C + row * W + c0 + (e & 1) * 8);
                }
            }
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) {
                const int c0 = n_base + (ci0 + 4 * d) * 16 + 2 * (lane & 3);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const long row = rt * 128 + q * 32 + (e >> 1) * 8 + (lane >> 2);
                    const int c = c0 + (e & 1) * 8;
                    const float2 o = make_float2(v[d][e].x + 1.f, v[d][e].y + 1.f);
                    *reinterpret_cast<float2*>(C + row * W + c) = o;
                    if (STREAMS == 3) *reinterpret_cast<float2*>(C2 + row * W + c) = make_float2(o.x * 2.f, o.y * 2.f);
                }
            }
        }
    }
}

template <int NT, int STREAMS>
__global__ void __launch_bounds__(512, 1) rowsweep_kernel(float* __restrict__ C, float* __restrict__ C2, long R, int W) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntn = (W + NT - 1) / NT;
    const long tiles = ((R + 127) / 128) * ntn;
    constexpr int SEG = NT / 64;              // float2 per lane and row
    for (long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long rt = t / ntn;
        const int n_base = (int)(t - rt * ntn) * NT;
        for (int r0 = warp; r0 < 128; r0 += 32) {      // two rows in flight per warp
            float2 v[2][SEG];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long row = rt * 128 + r0 + 16 * u;
#pragma unroll
                for (int s = 0; s < SEG; ++s) {
                    const int c = n_base + 64 * s + 2 * lane;
                    v[u][s] = make_float2(0.f, 0.f);
                    v[u][s] = *reinterpret_cast<const float2*>(C + row * W + c);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long row = rt * 128 + r0 + 16 * u;
#pragma unroll
                for (int s = 0; s < SEG; ++s) {
                    const int c = n_base + 64 * s + 2 * lane;
                    {
                        const float2 o = make_float2(v[u][s].x + 1.f, v[u][s].y + 1.f);
                        *reinterpret_cast<float2*>(C + row * W + c) = o;
                        if (STREAMS == 3) *reinterpret_cast<float2*>(C2 + row * W + c) = make_float2(o.x * 2.f, o.y * 2.f);
                    }
                }
            }
        }
    }
}


// the proposed transposed epilogue's global phase: 128-row x 128-column tiles, warp w takes rows w, w+16, ..: one float4 per lane
// and row (512 contiguous bytes per warp instruction), the eight loads of a tile issued before the first use
template <int STREAMS>
__global__ void __launch_bounds__(512, 1) rowsweep4_kernel(float* __restrict__ C, float* __restrict__ C2, long R, int W) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntn = W / 128;
    const long tiles = (R / 128) * ntn;
    for (long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long rt = t / ntn;
        const int n_base = (int)(t - rt * ntn) * 128;
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(C + (rt * 128 + warp + 16 * u) * W + n_base + 4 * lane);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float* dst = C + (rt * 128 + warp + 16 * u) * W + n_base + 4 * lane;
            const float4 o = make_float4(v[u].x + 1.f, v[u].y + 1.f, v[u].z + 1.f, v[u].w + 1.f);
            *reinterpret_cast<float4*>(dst) = o;
            if (STREAMS == 3) *reinterpret_cast<float4*>(C2 + (dst - C)) = make_float4(o.x * 2.f, o.y * 2.f, o.z * 2.f, o.w * 2.f);
        }
    }
}

int main() {
    const long R = 32L * 32 * 480;             // multiple of 128
    const int W = 512;                         // multiple of every tile width: no bounds checks in the kernels
    float *C, *C2;
    CK(cudaMalloc(&C, R * W * 4));
    CK(cudaMalloc(&C2, R * W * 4));
    CK(cudaMemset(C, 0, R * W * 4));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, int streams, auto launch) {
        for (int i = 0; i < 2; ++i) launch();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) launch();
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-34s %8.3f ms  %7.0f GB/s\n", name, ms / 5, (double)streams * R * W * 4 / (ms / 5) / 1e6);
    };
    time("flat float2, 3 streams", 3, [&] { flat_kernel<<<sms, 512>>>(C, C2, R * W / 2); });
    time("flat float2 x4 CTAs, 3 streams", 3, [&] { flat_kernel<<<sms * 4, 512>>>(C, C2, R * W / 2); });
    time("quad, 3 streams, 8 loads", 3, [&] { quad_kernel<3, 1><<<sms, 512>>>(C, C2, R, W); });
    time("quad, 3 streams, 16 loads", 3, [&] { quad_kernel<3, 2><<<sms, 512>>>(C, C2, R, W); });
    time("quad, 2 streams, 8 loads", 2, [&] { quad_kernel<2, 1><<<sms, 512>>>(C, C2, R, W); });
    time("quad, 2 streams, 16 loads", 2, [&] { quad_kernel<2, 2><<<sms, 512>>>(C, C2, R, W); });
    time("quad, 2 streams, 32 loads", 2, [&] { quad_kernel<2, 4><<<sms, 512>>>(C, C2, R, W); });
    time("rowsweep4 NT=128, 3 streams", 3, [&] { rowsweep4_kernel<3><<<sms, 512>>>(C, C2, R, W); });
    time("rowsweep4 NT=128, 2 streams", 2, [&] { rowsweep4_kernel<2><<<sms, 512>>>(C, C2, R, W); });
    time("rowsweep NT=256, 3 streams", 3, [&] { rowsweep_kernel<256, 3><<<sms, 512>>>(C, C2, R, W); });
    time("rowsweep NT=256, 2 streams", 2, [&] { rowsweep_kernel<256, 2><<<sms, 512>>>(C, C2, R, W); });
    time("rowsweep NT=512, 2 streams", 2, [&] { rowsweep_kernel<512, 2><<<sms, 512>>>(C, C2, R, W); });
    return 0;
}
