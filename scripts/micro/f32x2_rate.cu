// Microbenchmark: issue rate of FADD vs FADD2 (add.f32x2) on sm_100a, alone and mixed with FMNMX
// (ALU pipe).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2_rate f32x2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1},%2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float add1(float a, float b) { float r; asm volatile("add.rn.f32 %0,%1,%2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float mnmx(float a, float b) { float r; asm volatile("min.f32 %0,%1,%2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const float* in, float* out, int iters)
{
	float x[16];
#pragma unroll
	for (int i = 0; i < 16; ++i) x[i] = in[(threadIdx.x + i) & 1023];
	const float c = in[5];
	if (MODE == 0) {  // 16 independent FADD per iteration
		for (int it = 0; it < iters; ++it) {
#pragma unroll
			for (int i = 0; i < 16; ++i) x[i] = add1(x[i], c);
		}
	} else if (MODE == 1) {  // 8 independent FADD2 per iteration (same lane-ops as MODE 0)
		u64 v[8]; const u64 cc = pk(c, c);
#pragma unroll
		for (int i = 0; i < 8; ++i) v[i] = pk(x[2 * i], x[2 * i + 1]);
		for (int it = 0; it < iters; ++it) {
#pragma unroll
			for (int i = 0; i < 8; ++i) v[i] = add2(v[i], cc);
		}
#pragma unroll
		for (int i = 0; i < 8; ++i) upk(v[i], x[2 * i], x[2 * i + 1]);
	} else if (MODE == 2) {  // 8 FADD + 8 FMNMX
		for (int it = 0; it < iters; ++it) {
#pragma unroll
			for (int i = 0; i < 8; ++i) x[i] = add1(x[i], c);
#pragma unroll
			for (int i = 8; i < 16; ++i) x[i] = mnmx(x[i], c);
		}
	} else if (MODE == 3) {  // 4 FADD2 + 8 FMNMX (same work as MODE 2)
		u64 v[4]; const u64 cc = pk(c, c);
#pragma unroll
		for (int i = 0; i < 4; ++i) v[i] = pk(x[2 * i], x[2 * i + 1]);
		for (int it = 0; it < iters; ++it) {
#pragma unroll
			for (int i = 0; i < 4; ++i) v[i] = add2(v[i], cc);
#pragma unroll
			for (int i = 8; i < 16; ++i) x[i] = mnmx(x[i], c);
		}
#pragma unroll
		for (int i = 0; i < 4; ++i) upk(v[i], x[2 * i], x[2 * i + 1]);
	} else if (MODE == 4) {  // 16 FMNMX
		for (int it = 0; it < iters; ++it) {
#pragma unroll
			for (int i = 0; i < 16; ++i) x[i] = mnmx(x[i], c);
		}
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, const float* in, float* out, double ops_per_iter_lane)
{
	const int iters = 20000;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k<MODE><<<148, 512>>>(in, out, 100);
	cudaEventRecord(e0);
	k<MODE><<<148, 512>>>(in, out, iters);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	// warp-instructions per SMSP per ns
	const double warps_per_smsp = 512 / 32 / 4.0;
	const double winst = warps_per_smsp * iters * ops_per_iter_lane;  // per SMSP
	printf("%-28s %8.3f ms  %6.3f warp-inst/ns/SMSP  (%.2f inst/clk at 1.965 GHz)\n", name, ms, winst / (ms * 1e6), winst / (ms * 1e6) / 1.965);
}

int main()
{
	float *in, *out; cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 512 * 4); cudaMemset(in, 0, 4096);
	run<0>("16 FADD", in, out, 16);
	run<1>("8 FADD2", in, out, 8);
	run<2>("8 FADD + 8 FMNMX", in, out, 16);
	run<3>("4 FADD2 + 8 FMNMX", in, out, 12);
	run<4>("16 FMNMX", in, out, 16);
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
