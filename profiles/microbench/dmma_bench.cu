// Microbenchmark: FP64 tensor (DMMA) shapes vs DFMA on sm_100a. Scratch, not product.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s @%d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int SHAPE> __global__ void __launch_bounds__(256) k_dmma(double* out, int iters) {
  double a[8], b[4]; double c[8][4];
  for (int i=0;i<8;i++) a[i]=threadIdx.x*1e-3+i; for(int i=0;i<4;i++) b[i]=threadIdx.x*2e-3+i;
  for (int j=0;j<8;j++) for(int i=0;i<4;i++) c[j][i]=0;
  for (int it=0; it<iters; ++it) {
    #pragma unroll
    for (int j=0;j<8;j++) {
      if (SHAPE==0) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
          : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a[j&7]), "d"(b[j&3]));
      } else if (SHAPE==1) {
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5},{%6},{%0,%1,%2,%3};\n"
          : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3]) : "d"(a[j&7]), "d"(a[(j+1)&7]), "d"(b[j&3]));
      } else if (SHAPE==2) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7},{%8,%9},{%0,%1,%2,%3};\n"
          : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[j&1]), "d"(b[2]));
      } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7,%8,%9,%10,%11},{%12,%13,%14,%15},{%0,%1,%2,%3};\n"
          : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
      }
    }
  }
  double s=0; for (int j=0;j<8;j++) for(int i=0;i<4;i++) s+=c[j][i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
  double c[16]; double a=threadIdx.x*1e-3, b=1.0000001;
  for (int i=0;i<16;i++) c[i]=i;
  for (int it=0; it<iters; ++it) {
    #pragma unroll
    for (int j=0;j<16;j++) c[j]=fma(c[j],b,a);
  }
  double s=0; for(int i=0;i<16;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  double* out; CK(cudaMalloc(&out, 148*8*256*sizeof(double)));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters=20000;
  const double flops_per_mma[4]={2.0*8*8*4, 2.0*16*8*4, 2.0*16*8*8, 2.0*16*8*16};
  const char* names[4]={"m8n8k4","m16n8k4","m16n8k8","m16n8k16"};
  for (int occ=1; occ<=4; occ*=2) {
   int grid=148*occ;
   for (int s=0;s<4;s++){
    float best=1e30f;
    for(int rep=0;rep<3;rep++){
      cudaEventRecord(e0);
      if(s==0) k_dmma<0><<<grid,256>>>(out,iters); else if(s==1) k_dmma<1><<<grid,256>>>(out,iters);
      else if(s==2) k_dmma<2><<<grid,256>>>(out,iters); else k_dmma<3><<<grid,256>>>(out,iters);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms;
    }
    double fl = (double)grid*8 /*warps*/ * iters * 8 * flops_per_mma[s];
    printf("DMMA %-9s grid=%d: %.3f ms  %.2f TFLOP/s\n", names[s], grid, best, fl/best*1e-9);
   }
   float best=1e30f;
   for(int rep=0;rep<3;rep++){ cudaEventRecord(e0); k_dfma<<<grid,256>>>(out,iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms;}
   double fl=(double)grid*256*iters*16*2;
   printf("DFMA grid=%d: %.3f ms  %.2f TFLOP/s\n", grid, best, fl/best*1e-9);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
