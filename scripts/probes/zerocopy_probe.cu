// zerocopy_probe.cu - how fast can the GPU gather from / scatter into a pinned host AoS (P[] 536 B, SphP[] 264 B)
// directly over PCIe, compared with a staged cudaMemcpy?  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while(0)

// one thread per particle, fields read as 16-byte vectors where alignment allows (P stride 536 = 8 mod 16 -> use 8-byte loads)
__global__ void k_pull(long long n, const char *__restrict__ P, const char *__restrict__ S, double *__restrict__ out, int *__restrict__ iout)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  const double *p = (const double *) (P + i * 536);
  const double *s = (const double *) (S + i * 264);
  double a[11];
  a[0] = p[0]; a[1] = p[1]; a[2] = p[2]; a[3] = p[6];	// pos, mass
  a[4] = p[12]; a[5] = p[14];			// oldacc, hsml
  a[6] = s[0]; a[7] = s[2]; a[8] = s[3]; a[9] = s[4]; a[10] = s[11];	// entropy, velpred, dtentropy
  const int *q = (const int *) (P + i * 536 + 524);
  int tb = q[0], ty = q[2];
#pragma unroll
  for(int k = 0; k < 11; k++) out[k * n + i] = a[k];
  iout[i] = tb; iout[n + i] = ty;
}
// lane-cooperative variant: a warp handles 32 particles, each 64 B segment is read by 8 lanes x 8 B so the request is one 64 B burst
__global__ void k_pull_coop(long long n, const char *__restrict__ P, const char *__restrict__ S, double *__restrict__ out, int *__restrict__ iout)
{
  const int lane = threadIdx.x & 31;
  const long long w = (blockIdx.x * (long long) blockDim.x + threadIdx.x) >> 5;
  const long long base = w * 32;
  if(base >= n) return;
  // 4 particles per pass, 8 lanes each
  for(int pass = 0; pass < 8; pass++)
    {
      const long long i = base + pass * 4 + (lane >> 3);
      const int l = lane & 7;
      if(i >= n) continue;
      const double *p = (const double *) (P + i * 536);
      const double *s = (const double *) (S + i * 264);
      double v0 = p[l];		// bytes 0..63
      double v1 = p[8 + l];	// 64..127
      double v2 = s[l];		// 0..63
      double v3 = (l < 4) ? s[8 + l] : 0;	// 64..95
      int q = (l < 3) ? ((const int *) (P + i * 536 + 524))[l] : 0;
      out[(long long) l * n + i] = v0;
      out[(long long) (8 + l) * n + i] = v1;
      out[(long long) (16 + l) * n + i] = v2;
      if(l < 4) out[(long long) (24 + l) * n + i] = v3;
      if(l < 3) iout[(long long) l * n + i] = q;
    }
}
__global__ void k_push(long long n, char *__restrict__ P, char *__restrict__ S, const double *__restrict__ in)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  double *p = (double *) (P + i * 536);
  double *s = (double *) (S + i * 264);
  double a[17];
#pragma unroll
  for(int k = 0; k < 17; k++) a[k] = in[k * n + i];
  p[8] = a[0]; p[9] = a[1]; p[10] = a[2]; p[12] = a[3]; p[14] = a[4]; p[15] = a[5];	// GravAccel, OldAcc, Hsml, NumNgb
  *(float *) (P + i * 536 + 520) = (float) a[6];
  s[1] = a[7]; s[5] = a[8];	// Pressure, MaxSignalVel
  s[10] = a[9]; s[11] = a[10]; s[12] = a[11]; s[13] = a[12]; s[14] = a[13];	// Density, DtEntropy, HydroAccel
  s[19] = a[14]; s[20] = a[15]; s[21] = a[16];	// Dhsml, DivVel, CurlVel
}
int main(int argc, char **argv)
{
  long long n = argc > 1 ? atoll(argv[1]) : 1600000;
  char *P, *S;
  CK(cudaMallocHost(&P, n * 536));
  CK(cudaMallocHost(&S, n * 264));
  memset(P, 1, n * 536);
  memset(S, 2, n * 264);
  double *d; int *di; char *dP;
  CK(cudaMalloc(&d, n * 32 * 8));
  CK(cudaMalloc(&di, n * 4 * 4));
  CK(cudaMalloc(&dP, n * 536));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for(int rep = 0; rep < 2; rep++)
    {
      cudaEventRecord(e0);
      k_pull<<<(n + 127) / 128, 128>>>(n, P, S, d, di);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
      printf("k_pull (thread/particle)   %.2f ms  payload %.0f MB -> %.1f GB/s payload\n", ms, n * 100e-6, n * 100e-6 / ms);
      cudaEventRecord(e0);
      k_pull_coop<<<(n + 127) / 128, 128>>>(n, P, S, d, di);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
      printf("k_pull_coop (8 lanes/64B)  %.2f ms  bytes read %.0f MB -> %.1f GB/s\n", ms, n * 236e-6, n * 236e-6 / ms);
      cudaEventRecord(e0);
      k_push<<<(n + 127) / 128, 128>>>(n, P, S, d);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
      printf("k_push (thread/particle)   %.2f ms  payload %.0f MB -> %.1f GB/s payload\n", ms, n * 132e-6, n * 132e-6 / ms);
      // the copy engine's strided (2D) mode: rows of `w` bytes at the struct stride, no SM and no CPU involved
      for(int w : {24, 40, 68, 128})
	{
	  cudaEventRecord(e0);
	  CK(cudaMemcpy2DAsync(dP, 128, P, 536, w, n, cudaMemcpyHostToDevice));
	  cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
	  printf("memcpy2D H2D rows of %3d B      %.2f ms  %.0f MB -> %.1f GB/s\n", w, ms, n * (double) w * 1e-6, n * (double) w * 1e-6 / ms);
	  cudaEventRecord(e0);
	  CK(cudaMemcpy2DAsync(P + 64, 536, dP, 128, w, n, cudaMemcpyDeviceToHost));
	  cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
	  printf("memcpy2D D2H rows of %3d B      %.2f ms  %.0f MB -> %.1f GB/s\n", w, ms, n * (double) w * 1e-6, n * (double) w * 1e-6 / ms);
	}
      {
	// how many strided copies run at once?  k streams, each scattering 24-byte rows for n/k of the particles
	cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
	for(int k : {1, 2, 3, 4, 6})
	  {
	    cudaStream_t st[6];
	    for(int q = 0; q < k; q++) cudaStreamCreateWithFlags(&st[q], cudaStreamNonBlocking);
	    CK(cudaDeviceSynchronize());
	    cudaEventRecord(e0);
	    for(int q = 0; q < k; q++) CK(cudaStreamWaitEvent(st[q], e0, 0));
	    for(int q = 0; q < k; q++)
	      {
		long long a = n * q / k, b = n * (q + 1) / k;
		CK(cudaMemcpy2DAsync(P + 64 + a * 536, 536, dP + a * 128, 128, 24, b - a, cudaMemcpyDeviceToHost, st[q]));
	      }
	    for(int q = 0; q < k; q++) { cudaEvent_t d; cudaEventCreate(&d); cudaEventRecord(d, st[q]); cudaStreamWaitEvent(0, d, 0); }
	    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
	    printf("D2H 24-B rows, n split over %d streams: %.2f ms   (asyncEngineCount %d)\n", k, ms, prop.asyncEngineCount);
	    // same rows count per stream but k different fields for ALL particles (k x the work)
	    cudaEventRecord(e0);
	    for(int q = 0; q < k; q++) CK(cudaStreamWaitEvent(st[q], e0, 0));
	    for(int q = 0; q < k; q++) CK(cudaMemcpy2DAsync(P + 64 + 32 * q, 536, dP + 16 * q, 128, 16, n, cudaMemcpyDeviceToHost, st[q]));
	    for(int q = 0; q < k; q++) { cudaEvent_t d; cudaEventCreate(&d); cudaEventRecord(d, st[q]); cudaStreamWaitEvent(0, d, 0); }
	    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
	    printf("D2H %d different 16-B fields of all n, one stream each: %.2f ms\n", k, ms);
	    for(int q = 0; q < k; q++) cudaStreamDestroy(st[q]);
	  }
      }
      cudaEventRecord(e0);
      CK(cudaMemcpyAsync(dP, P, n * 536, cudaMemcpyHostToDevice));
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
      printf("memcpy H2D whole P[]       %.2f ms  %.0f MB -> %.1f GB/s\n", ms, n * 536e-6, n * 536e-6 / ms);
      cudaEventRecord(e0);
      CK(cudaMemcpyAsync(P, dP, n * 536, cudaMemcpyDeviceToHost));
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
      printf("memcpy D2H whole P[]       %.2f ms  %.0f MB -> %.1f GB/s\n", ms, n * 536e-6, n * 536e-6 / ms);
    }
  return 0;
}
