#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template<int MODE> __global__ void __launch_bounds__(256) k(uint32_t* out, int iters, uint32_t seed, uint32_t kk){
  uint32_t a[8], b[8];
  for(int i=0;i<8;i++){a[i]=seed+threadIdx.x*9u+i; b[i]=seed^(blockIdx.x*131u+threadIdx.x*7u+i*77u);}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int r=0;r<8;r++){
      #pragma unroll
      for(int i=0;i<8;i++){
        if(MODE==0){ a[i]=__viaddmin_u32(a[i], kk, b[i]); b[i]=__viaddmin_u32(b[i], 1u, a[i]); }        // 2 VIADDMNMX
        if(MODE==1){ a[i]=min(min(a[i], b[i]), kk+i); b[i]=min(min(b[i],a[i]^1), seed); }               // VIMNMX3?
        if(MODE==2){ a[i]=(a[i]&~3u)|b[i]; b[i]=(b[i]&0xfffffff0u)|a[i]; }                             // LOP3
        if(MODE==3){ asm("mad.lo.u32 %0,%1,%2,%3;":"=r"(a[i]):"r"(a[i]),"r"(kk),"r"(b[i])); asm("mad.lo.u32 %0,%1,%2,%3;":"=r"(b[i]):"r"(b[i]),"r"(kk),"r"(a[i])); } // IMAD
        if(MODE==4){ a[i]=__funnelshift_r(a[i], b[i], 2); b[i]=__funnelshift_r(b[i], a[i], 3); }       // SHF
        if(MODE==5){ a[i]=__viaddmin_u32(a[i], kk, b[i]); asm("mad.lo.u32 %0,%1,%2,%3;":"=r"(b[i]):"r"(b[i]),"r"(kk),"r"(a[i])); } // 1 VIADDMNMX + 1 IMAD
        if(MODE==6){ a[i]=(a[i]&~3u)|b[i]; asm("mad.lo.u32 %0,%1,%2,%3;":"=r"(b[i]):"r"(b[i]),"r"(kk),"r"(a[i])); } // LOP3 + IMAD
        if(MODE==7){ a[i]=min(a[i],b[i]); b[i]=min(b[i]+kk, a[i]+seed); } // VIMNMX + ...
      }
    }
  }
  uint32_t s=0; for(int i=0;i<8;i++) s+=a[i]^b[i]; if(s==0x12345678u) out[0]=s;
}
template<int M> void run(const char* name, uint32_t* d, int sms){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid=sms*8, iters=2048; float best=1e9;
  for(int rep=0;rep<4;rep++){ cudaEventRecord(e0); k<M><<<grid,256>>>(d,iters,12345u,3u+rep); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(rep&&ms<best)best=ms; }
  double inst=(double)grid*256*iters*64.0*2.0;
  printf("%-28s %.3f ms  %.2f T lane-instr/s (assuming 2 instr per pair)  = %.1f lane-instr/clk/SM @1.965GHz\n", name, best, inst/best/1e9, inst/(best*1e-3)/1.965e9/sms);
}
int main(){ uint32_t* d; cudaMalloc(&d,64); cudaDeviceProp p; cudaGetDeviceProperties(&p,0); int sms=p.multiProcessorCount;
 run<0>("VIADDMNMX x2",d,sms); run<1>("min3 x2",d,sms); run<2>("LOP3 x2",d,sms); run<3>("IMAD x2",d,sms); run<4>("SHF x2",d,sms); run<5>("VIADDMNMX+IMAD",d,sms); run<6>("LOP3+IMAD",d,sms); run<7>("mode7",d,sms); return 0; }
