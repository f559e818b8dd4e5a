// How long do pinned-host / device allocations of a batch's size take on this box?
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
	cudaFree(0);
	const size_t sizes[] = {17u << 20, 23u << 20, 40u << 20, 160u << 20};
	for (size_t sz : sizes) {
		void *h = nullptr, *d = nullptr;
		double t0 = now();
		cudaHostAlloc(&h, sz, cudaHostAllocPortable);
		double t1 = now();
		cudaMalloc(&d, sz);
		double t2 = now();
		void* m = malloc(sz);
		memset(m, 1, sz);
		double t3 = now();
		cudaHostRegister(m, sz, cudaHostRegisterPortable);
		double t4 = now();
		cudaFreeHost(h);
		double t5 = now();
		cudaHostUnregister(m);
		double t6 = now();
		printf("%4zu MB: cudaHostAlloc %.3f s  cudaMalloc %.3f s  malloc+touch %.3f s  cudaHostRegister %.3f s  cudaFreeHost %.3f s  unregister %.3f s\n",
		       sz >> 20, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5);
		free(m); cudaFree(d);
	}
	return 0;
}
