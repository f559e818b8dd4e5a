// Allocation costs that decide the start-up of the streaming pipeline: pinned / device allocations of batch size, alone,
// after a 30 GB scratch arena exists, and from three host threads at once.  nvcc -O2 -o alloc_cost alloc_cost.cu
#include <sys/mman.h>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void one(const char* tag, size_t sz)
{
	void *h = nullptr, *d = nullptr;
	const double t0 = now();
	cudaHostAlloc(&h, sz, cudaHostAllocPortable);
	const double t1 = now();
	cudaMalloc(&d, sz);
	const double t2 = now();
	printf("%-28s %4zu MB: cudaHostAlloc %.4f s (%.2f GB/s)  cudaMalloc %.4f s\n", tag, sz >> 20, t1 - t0, sz / (t1 - t0) / 1e9, t2 - t1);
	cudaFreeHost(h); cudaFree(d);
}
int main()
{
	cudaFree(0);
	const size_t sizes[] = {20u << 20, 40u << 20, 160u << 20};
	for (size_t sz : sizes) one("fresh context", sz);
	void* big = nullptr;
	double t0 = now();
	cudaMalloc(&big, (size_t)30 << 30);
	printf("cudaMalloc 30 GB: %.4f s\n", now() - t0);
	for (size_t sz : sizes) one("with 30 GB arena", sz);
	for (size_t sz : sizes) {
		t0 = now();
		std::vector<std::thread> th;
		for (int k = 0; k < 3; k++) th.emplace_back([sz] { void* h; cudaHostAlloc(&h, sz, cudaHostAllocPortable); cudaFreeHost(h); });
		for (auto& t : th) t.join();
		printf("3 threads x (cudaHostAlloc + cudaFreeHost) %4zu MB: %.4f s total\n", sz >> 20, now() - t0);
	}
	// transparent huge pages: anonymous mapping + MADV_HUGEPAGE + first touch, then cudaHostRegister
	for (size_t sz : sizes) {
		const size_t al = (sz + (2u << 20) - 1) & ~((size_t)(2u << 20) - 1);
		double t0 = now();
		void* p = mmap(nullptr, al, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
		madvise(p, al, MADV_HUGEPAGE);
		memset(p, 0, al);
		double t1 = now();
		cudaError_t e = cudaHostRegister(p, al, cudaHostRegisterPortable);
		double t2 = now();
		printf("THP %4zu MB: mmap+madvise+touch %.4f s  cudaHostRegister %.4f s (%s)  total %.2f GB/s\n", sz >> 20, t1 - t0, t2 - t1,
		       cudaGetErrorString(e), sz / (t2 - t0) / 1e9);
		cudaHostUnregister(p); munmap(p, al);
	}
	for (size_t sz : sizes) {   // the same without the huge-page hint
		double t0 = now();
		void* p = mmap(nullptr, sz, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
		memset(p, 0, sz);
		double t1 = now();
		cudaHostRegister(p, sz, cudaHostRegisterPortable);
		double t2 = now();
		printf("4K  %4zu MB: mmap+touch %.4f s  cudaHostRegister %.4f s  total %.2f GB/s\n", sz >> 20, t1 - t0, t2 - t1, sz / (t2 - t0) / 1e9);
		cudaHostUnregister(p); munmap(p, sz);
	}
	{   // pinned allocation while a 30 GB cudaMalloc runs on another thread
		cudaFree(big);
		std::thread a([&] { double t = now(); cudaMalloc(&big, (size_t)30 << 30); printf("  concurrent cudaMalloc 30 GB: %.4f s\n", now() - t); });
		one("during the 30 GB cudaMalloc", 20u << 20);
		a.join();
	}
	return 0;
}
