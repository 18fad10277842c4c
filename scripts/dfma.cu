#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double x, double y){
  double a = x + threadIdx.x;
  long long t0 = clock64();
  #pragma unroll
  for (int i=0;i<256;i++) a = fma(a, y, x);
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x*blockDim.x] = a; if (threadIdx.x==0) cyc[blockIdx.x] = t1-t0;
}
template<int ILP> __global__ void thr(double* out, long long* cyc, double x, double y){
  double a[ILP];
  for (int j=0;j<ILP;j++) a[j] = x + threadIdx.x + j;
  long long t0 = clock64();
  #pragma unroll
  for (int i=0;i<64;i++)
    #pragma unroll
    for (int j=0;j<ILP;j++) a[j] = fma(a[j], y, x);
  long long t1 = clock64();
  double s=0; for (int j=0;j<ILP;j++) s+=a[j];
  out[threadIdx.x + blockIdx.x*blockDim.x] = s; if (threadIdx.x==0) cyc[blockIdx.x] = t1-t0;
}
__global__ void rs(double* out, long long* cyc, double x){
  double a = x + threadIdx.x;
  long long t0 = clock64();
  #pragma unroll
  for (int i=0;i<32;i++) a = rsqrt(a) + 2.0;
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x==0) cyc[0] = t1-t0;
}
__global__ void dm(double* out, long long* cyc, double x, double y){
  double c0=x,c1=y; double a = x + (threadIdx.x&3), b = y + (threadIdx.x>>2);
  long long t0 = clock64();
  #pragma unroll
  for (int i=0;i<256;i++) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x*blockDim.x] = c0+c1; if (threadIdx.x==0) cyc[blockIdx.x] = t1-t0;
}
template<int ILP> __global__ void dmt(double* out, long long* cyc, double x, double y){
  double c[ILP][2]; for(int j=0;j<ILP;j++){c[j][0]=x+j;c[j][1]=y;} double a = x + (threadIdx.x&3), b = y + (threadIdx.x>>2);
  long long t0 = clock64();
  #pragma unroll
  for (int i=0;i<64;i++)
   #pragma unroll
   for(int j=0;j<ILP;j++) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  long long t1 = clock64();
  double s=0; for(int j=0;j<ILP;j++) s+=c[j][0]+c[j][1];
  out[threadIdx.x + blockIdx.x*blockDim.x] = s; if (threadIdx.x==0) cyc[blockIdx.x] = t1-t0;
}
int main(){
  double* out; long long* cyc; cudaMalloc(&out, 8*1024*1024); cudaMalloc(&cyc, 8*4096); long long h[4096];
  lat<<<1,32>>>(out,cyc,1.0,0.999); cudaMemcpy(h,cyc,8,cudaMemcpyDeviceToHost); printf("DFMA dependent latency: %.1f cycles\n", h[0]/256.0);
  thr<8><<<1,32>>>(out,cyc,1.0,0.999); cudaMemcpy(h,cyc,8,cudaMemcpyDeviceToHost); printf("DFMA 1 warp ILP8: %.2f cycles/DFMA\n", h[0]/(64.0*8));
  thr<16><<<1,32>>>(out,cyc,1.0,0.999); cudaMemcpy(h,cyc,8,cudaMemcpyDeviceToHost); printf("DFMA 1 warp ILP16: %.2f cycles/DFMA\n", h[0]/(64.0*16));
  thr<8><<<1,128>>>(out,cyc,1.0,0.999); cudaMemcpy(h,cyc,8,cudaMemcpyDeviceToHost); printf("DFMA 4 warps (1/SMSP) ILP8: %.2f cycles/DFMA per warp\n", h[0]/(64.0*8));
  thr<8><<<1,256>>>(out,cyc,1.0,0.999); cudaMemcpy(h,cyc,8,cudaMemcpyDeviceToHost); printf("DFMA 8 warps ILP8: %.2f cycles/DFMA per warp\n", h[0]/(64.0*8));
  thr<8><<<1,1024>>>(out,cyc,1.0,0.999); cudaMemcpy(h,cyc,8,cudaMemcpyDeviceToHost); printf("DFMA 32 warps ILP8: %.2f cycles/DFMA per warp -> %.1f DFMA lanes/clk/SM\n", h[0]/(64.0*8), 32*32*64.0*8/h[0]);
  rs<<<1,32>>>(out,cyc,3.0); cudaMemcpy(h,cyc,8,cudaMemcpyDeviceToHost); printf("rsqrt(double)+add dependent: %.1f cycles\n", h[0]/32.0);
  dm<<<1,32>>>(out,cyc,1.0,0.5); cudaMemcpy(h,cyc,8,cudaMemcpyDeviceToHost); printf("DMMA dependent latency: %.1f cycles\n", h[0]/256.0);
  dmt<8><<<1,32>>>(out,cyc,1.0,0.5); cudaMemcpy(h,cyc,8,cudaMemcpyDeviceToHost); printf("DMMA 1 warp ILP8: %.2f cycles/DMMA\n", h[0]/(64.0*8));
  dmt<8><<<1,128>>>(out,cyc,1.0,0.5); cudaMemcpy(h,cyc,8,cudaMemcpyDeviceToHost); printf("DMMA 4 warps ILP8: %.2f cycles/DMMA per warp\n", h[0]/(64.0*8));
  dmt<8><<<1,512>>>(out,cyc,1.0,0.5); cudaMemcpy(h,cyc,8,cudaMemcpyDeviceToHost); printf("DMMA 16 warps ILP8: %.2f cycles/DMMA per warp -> %.1f flops/clk/SM\n", h[0]/(64.0*8), 16*64.0*8*512/h[0]);
  // full chip DMMA throughput
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  dmt<8><<<148*4,512>>>(out,cyc,1.0,0.5); cudaEventRecord(e0); dmt<8><<<148*4,512>>>(out,cyc,1.0,0.5); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1);
  printf("DMMA full chip: %.2f TFLOP/s\n", 148.0*4*16*64*8*512/ms/1e9);
  thr<8><<<148*4,1024>>>(out,cyc,1.0,0.999); cudaEventRecord(e0); thr<8><<<148*4,1024>>>(out,cyc,1.0,0.999); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
  printf("DFMA full chip: %.2f TFLOP/s  (%s)\n", 148.0*4*1024*64*8*2/ms/1e9, cudaGetErrorString(cudaGetLastError()));
}
