// fma_peaks.cu - measured (not derived) non-tensor FMA peaks of the GPU the roofline fractions are quoted against:
// FP32 FFMA and FP64 DFMA throughput of the CUDA cores, 8 independent dependency chains per thread, all SMs full.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peaks fma_peaks.cu && ./fma_peaks > profiles/r02_fma_peaks.json
#include <cstdio>
#include <cuda_runtime.h>

template <class T> __global__ void __launch_bounds__(256) k_fma(T *out, T a, T b, int iters)
{
  T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for(int i = 0; i < iters; i++)
    {
#pragma unroll 16
      for(int u = 0; u < 16; u++)
	{
	  x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
	  x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
	}
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <class T> double run(int sms, int iters)
{
  const int blocks = sms * 8, threads = 256;
  T *out;
  cudaMalloc(&out, sizeof(T) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_fma<T><<<blocks, threads>>>(out, (T) 1.0000001, (T) 1e-7, iters / 8);
  cudaDeviceSynchronize();
  double best = 0;
  for(int rep = 0; rep < 5; rep++)
    {
      cudaEventRecord(e0);
      k_fma<T><<<blocks, threads>>>(out, (T) 1.0000001, (T) 1e-7, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flop = 2.0 * 8 * 16 * (double) iters * blocks * threads;
      const double tf = flop / (ms * 1e-3) / 1e12;
      if(tf > best) best = tf;
    }
  cudaFree(out);
  return best;
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const double f32 = run<float>(p.multiProcessorCount, 4096), f64 = run<double>(p.multiProcessorCount, 2048);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_max\": %.0f, \"fp32_fma_tflops\": %.2f, \"fp64_fma_tflops\": %.2f, "
	 "\"how\": \"scripts/probes/fma_peaks.cu: 8 independent FMA chains per thread, 8 blocks x 256 threads per SM, best of 5, CUDA events\"}\n",
	 p.name, p.multiProcessorCount, p.clockRate / 1e3, f32, f64);
  return 0;
}
