// Hand-off latency between two CTAs on different SMs: (a) tagged 16-byte relaxed store/poll,
// (b) data + st.release flag / ld.relaxed poll + fence.  Reports ns per one-way hop.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/bin/pingpong scripts/pingpong.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void st_tagged(double2 *p, double v, double tag)
{
    asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v), "d"(tag) : "memory");
}
__device__ __forceinline__ double2 ld_tagged(const double2 *p)
{
    double2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ unsigned long long now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// mode 0: tagged; mode 1: flag + release; 128 values per hop (like one solve block); grid = 2 CTAs x 128 threads
__global__ void pingpong(double2 *buf, double *plain, int *flag, int rounds, int mode, unsigned long long *out)
{
    const int me = blockIdx.x, other = 1 - me, tid = threadIdx.x;
    __shared__ double sh[128];
    unsigned long long t0 = 0;
    if (tid == 0) t0 = now();
    double acc = 1.0;
    for (int r = 0; r < rounds; ++r)
    {
        const double tag = (double)(r + 1);
        const bool my_turn = (r & 1) == me;
        if (my_turn)
        {
            if (mode == 0)
                st_tagged(buf + me * 128 + tid, acc + tid, tag);
            else
            {
                plain[me * 128 + tid] = acc + tid;
                __syncthreads();
                if (tid == 0) st_release(flag + me, r + 1);
            }
        }
        else
        {
            if (mode == 0)
            {
                if (tid < 32)
                {
                    double2 v0, v1, v2, v3;
                    for (;;)
                    {
                        v0 = ld_tagged(buf + other * 128 + tid);
                        v1 = ld_tagged(buf + other * 128 + tid + 32);
                        v2 = ld_tagged(buf + other * 128 + tid + 64);
                        v3 = ld_tagged(buf + other * 128 + tid + 96);
                        if (v0.y == tag && v1.y == tag && v2.y == tag && v3.y == tag) break;
                    }
                    sh[tid] = v0.x; sh[tid + 32] = v1.x; sh[tid + 64] = v2.x; sh[tid + 96] = v3.x;
                }
                __syncthreads();
            }
            else
            {
                if (tid == 0)
                {
                    while (ld_relaxed(flag + other) != r + 1) {}
                    asm volatile("fence.acq_rel.gpu;" ::: "memory");
                }
                __syncthreads();
                sh[tid] = __ldcg(plain + other * 128 + tid);
                __syncthreads();
            }
            acc = sh[(tid + 1) & 127] * 0.5;
        }
        __syncthreads();
    }
    if (tid == 0) out[me] = now() - t0;
    if (acc == 12345.678) out[2] = 1;
}
int main()
{
    double2 *buf; double *plain; int *flag; unsigned long long *out, h[3];
    cudaMalloc(&buf, 256 * 16); cudaMalloc(&plain, 256 * 8); cudaMalloc(&flag, 8); cudaMalloc(&out, 24);
    const int rounds = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 2; ++rep)
        {
            cudaMemset(buf, 0, 256 * 16); cudaMemset(flag, 0, 8);
            pingpong<<<2, 128>>>(buf, plain, flag, rounds, mode, out);
            cudaDeviceSynchronize();
            cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
            printf("mode %d (%s): %.0f ns per hop (%s)\n", mode, mode ? "data + release flag + fence" : "tagged 16-byte values",
                   (double)h[0] / rounds, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
