// Microbenchmark: thread-per-read against warp-per-read for ONE segment kind, the 49-HMM x 6-column
// barcode segment of cfg2 in the backward direction, without the Mb/Ib hand-off stores (what k_backward<false> does).
// Both kernels run the same per-(HMM, position) work with the product's logsum (8 instructions, 64 KB table in shared
// memory, lane-random gathers): 3 logsums per column for 5 columns plus the ordered fold of the 49 contributions into
// the silent state, 16 logsums per (HMM, position) -- the live count of the real kernel is 17.  The numbers are not
// TagDust's (random transitions / emissions), the amount and the dependency structure of the work are.
//
//   A  thread per read: 148 CTAs x 512 threads; HMM-outer, position-inner; the silent row cs[i] is read-modify-written
//      through global memory once per HMM (like k_backward), 12 state registers.
//   C  (below, added after B's result) one warp per HMM with 32 reads per warp, one CTA per 32 reads.
//   B  warp per read: lane = HMM (49 HMMs = two rounds of 32 lanes, 15 lanes idle in the second), position loop serial;
//      the ordered 49-term fold of a position is NOT on the recurrence's critical path (the HMMs read the NEXT segment's
//      silent row), so it is done afterwards for 32 positions at once, lane = position, from a [HMM][position] tile in
//      shared memory -- the best case for this mapping: no serial fold on one lane, no silent-state traffic per HMM.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o warp_per_read warp_per_read.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int H = 49, NC = 6, L = 150, READS = 148 * 512;
constexpr float NEG = -1e30f;
typedef uint32_t TabAddr;

__device__ __forceinline__ float LS(float a, float b, TabAddr tab)
{
	const float mx = fmaxf(a, b);
	const float p = fminf(fabsf(a - b) * 1000.0f, 15999.0f);
	const uint32_t bits = __float_as_uint(__fadd_rz(p, 8388608.0f));
	float t;
	asm("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(tab + (bits << 2)));
	return mx + t;
}

struct Model {
	float tMM, tMI, tIM, tII, tMD, tSM, tSkip;
};

// one position of one HMM: returns the HMM's contribution to the silent state of this position
__device__ __forceinline__ float hmm_step(float (&Mb)[NC], float (&Ib)[NC], const float* em, int x, float pn, const Model& m, TabAddr tab)
{
	float nM[NC], nI[NC];
	nM[NC - 1] = pn + m.tSkip;
	nI[NC - 1] = NEG;
	const float eI = em[NC * 4 + x];
#pragma unroll
	for (int c = NC - 2; c >= 0; --c) {
		const float eMn = em[(c + 1) * 4 + x];
		const float a = Mb[c + 1] + m.tMM + eMn, b = Ib[c] + eI;
		float v = LS(a, b + m.tMI, tab);
		v = LS(v, nM[c + 1] + m.tMD, tab);
		nM[c] = v;
		nI[c] = LS(a + m.tIM, b + m.tII, tab);
	}
#pragma unroll
	for (int c = 0; c < NC; ++c) { Mb[c] = nM[c]; Ib[c] = nI[c]; }
	return nM[0] + m.tSM + em[x];
}

// ---- A: thread per read
__global__ void __launch_bounds__(512, 1) k_thread(const float* tabg, const float* emg, const uint8_t* seq, const float* ps, float* cs, Model m)
{
	extern __shared__ float smem[];
	float* tabs = smem;
	float* ems = smem + 16000;
	for (int i = threadIdx.x; i < 16000; i += blockDim.x) tabs[i] = tabg[i];
	for (int i = threadIdx.x; i < H * (NC + 1) * 4; i += blockDim.x) ems[i] = emg[i];
	__syncthreads();
	const TabAddr tab = (uint32_t)__cvta_generic_to_shared(tabs) - (0x4B000000u << 2);
	const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t stride = (size_t)gridDim.x * blockDim.x;   // arrays are [position][read]
	for (int h = 0; h < H; ++h) {
		float Mb[NC], Ib[NC];
#pragma unroll
		for (int c = 0; c < NC; ++c) { Mb[c] = NEG; Ib[c] = NEG; }
		const float* em = ems + h * (NC + 1) * 4;
		for (int i = L; i >= 1; --i) {
			const int x = seq[(size_t)i * stride + r];
			const float pn = ps[(size_t)(i + 1) * stride + r];
			const float contrib = hmm_step(Mb, Ib, em, x, pn, m, tab);
			float c0 = cs[(size_t)i * stride + r];
			cs[(size_t)i * stride + r] = LS(c0, contrib, tab);
		}
	}
}

// ---- B: warp per read
constexpr int WARPS = 16;   // per CTA: 16 tiles of 64 x 33 floats = 132 KB beside the 64 KB table
constexpr int TP = 33;      // padded tile row: the lanes of a round write different rows of one column
__global__ void __launch_bounds__(WARPS * 32, 1) k_warp(const float* tabg, const float* emg, const uint8_t* seq, const float* ps, float* cs, Model m, int reads)
{
	extern __shared__ float smem[];
	float* tabs = smem;
	float* ems = smem + 16000;
	float* tiles = ems + H * (NC + 1) * 4 + 12;   // [warp][64 HMM slots][32 positions, padded]
	for (int i = threadIdx.x; i < 16000; i += blockDim.x) tabs[i] = tabg[i];
	for (int i = threadIdx.x; i < H * (NC + 1) * 4; i += blockDim.x) ems[i] = emg[i];
	__syncthreads();
	const TabAddr tab = (uint32_t)__cvta_generic_to_shared(tabs) - (0x4B000000u << 2);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	float* tile = tiles + (size_t)warp * 64 * TP;
	const int h0 = lane, h1 = 32 + lane;
	const bool two = h1 < H;
	const float* em0 = ems + h0 * (NC + 1) * 4;
	const float* em1 = ems + (two ? h1 : 0) * (NC + 1) * 4;
	for (size_t r = (size_t)blockIdx.x * WARPS + warp; r < (size_t)reads; r += (size_t)gridDim.x * WARPS) {
		float M0[NC], I0[NC], M1[NC], I1[NC];
#pragma unroll
		for (int c = 0; c < NC; ++c) { M0[c] = NEG; I0[c] = NEG; M1[c] = NEG; I1[c] = NEG; }
		for (int ib = L; ib >= 1; ib -= 32) {
			const int nb = ib >= 32 ? 32 : ib;
			// lane p holds the base and the next-segment silent value of position ib - p
			const int xi = (lane < nb) ? seq[r * (L + 2) + (ib - lane)] : 0;   // this kernel's arrays are [read][position]
			const float pi = (lane < nb) ? ps[r * (L + 2) + (ib - lane + 1)] : NEG;
			for (int p = 0; p < nb; ++p) {
				const int x = __shfl_sync(0xffffffffu, xi, p);
				const float pn = __shfl_sync(0xffffffffu, pi, p);
				tile[h0 * TP + p] = hmm_step(M0, I0, em0, x, pn, m, tab);
				if (two) tile[h1 * TP + p] = hmm_step(M1, I1, em1, x, pn, m, tab);
			}
			__syncwarp();
			// the ordered fold of the 49 contributions, lane = position
			if (lane < nb) {
				float c0 = NEG;
#pragma unroll 7
				for (int h = 0; h < H; ++h) c0 = LS(c0, tile[h * TP + lane], tab);
				cs[r * (L + 2) + (ib - lane)] = c0;
			}
			__syncwarp();
		}
	}
}

// ---- C: one warp per HMM, 32 reads per warp, one CTA per group of 32 reads ("thread per (read, HMM)")
// No idle lanes (lane = read, as in A) and no silent-state traffic (as in B): the 49 HMMs of a segment run as 25 warps of
// one CTA (warp w: HMMs w and w + 25), every warp walks the positions of its HMMs in blocks of 25 and leaves the
// contributions in a [HMM][position][read] tile in shared memory; after a barrier warp p folds position p of the block
// (lane = read, the 49 terms in order), after a second barrier the next block starts.  12-24 state registers per thread.
constexpr int CW = 25, PB = 25;
__global__ void __launch_bounds__(CW * 32, 1) k_cta(const float* tabg, const float* emg, const uint8_t* seq, const float* ps, float* cs, Model m, int reads)
{
	extern __shared__ float smem[];
	float* tabs = smem;
	float* ems = smem + 16000;
	float* tile = ems + H * (NC + 1) * 4 + 12;   // [HMM][PB][32 reads]
	for (int i = threadIdx.x; i < 16000; i += blockDim.x) tabs[i] = tabg[i];
	for (int i = threadIdx.x; i < H * (NC + 1) * 4; i += blockDim.x) ems[i] = emg[i];
	__syncthreads();
	const TabAddr tab = (uint32_t)__cvta_generic_to_shared(tabs) - (0x4B000000u << 2);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int h0 = warp, h1 = warp + CW;
	const bool two = h1 < H;
	const float* em0 = ems + h0 * (NC + 1) * 4;
	const float* em1 = ems + (two ? h1 : 0) * (NC + 1) * 4;
	const size_t stride = (size_t)reads;   // arrays are [position][read], like A
	for (size_t g = blockIdx.x; g * 32 < (size_t)reads; g += gridDim.x) {
		const size_t r = g * 32 + lane;
		float M0[NC], I0[NC], M1[NC], I1[NC];
#pragma unroll
		for (int c = 0; c < NC; ++c) { M0[c] = NEG; I0[c] = NEG; M1[c] = NEG; I1[c] = NEG; }
		for (int ib = L; ib >= 1; ib -= PB) {
			const int nb = ib >= PB ? PB : ib;
			for (int p = 0; p < nb; ++p) {
				const int i = ib - p;
				const int x = seq[(size_t)i * stride + r];
				const float pn = ps[(size_t)(i + 1) * stride + r];
				tile[(h0 * PB + p) * 32 + lane] = hmm_step(M0, I0, em0, x, pn, m, tab);
				if (two) tile[(h1 * PB + p) * 32 + lane] = hmm_step(M1, I1, em1, x, pn, m, tab);
			}
			__syncthreads();
			if (warp < nb) {
				float c0 = NEG;
#pragma unroll 7
				for (int h = 0; h < H; ++h) c0 = LS(c0, tile[(h * PB + warp) * 32 + lane], tab);
				cs[(size_t)(ib - warp) * stride + r] = c0;
			}
			__syncthreads();
		}
	}
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main()
{
	std::vector<float> tab(16000);
	for (int i = 0; i < 16000; i++) tab[i] = i >= 15700 ? 0.0f : (float)log(1.0 + exp(-(double)i / 1000.0));
	std::vector<float> em((size_t)H * (NC + 1) * 4);
	unsigned s = 12345;
	auto rnd = [&] { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); };
	for (auto& e : em) e = logf(0.02f + 0.9f * rnd());
	const size_t cells = (size_t)(L + 2) * READS;
	std::vector<uint8_t> seq(cells);
	for (auto& c : seq) c = (uint8_t)(rnd() * 4.0f) & 3;
	std::vector<float> ps(cells);
	for (auto& v : ps) v = -40.0f * rnd();
	Model m = {logf(0.9f), logf(0.05f), logf(0.5f), logf(0.5f), logf(0.05f), logf(1.0f / H), 0.0f};
	// the same data read-major for the warp kernel
	std::vector<uint8_t> seqB(cells);
	std::vector<float> psB(cells);
	for (size_t i = 0; i < (size_t)(L + 2); i++)
		for (size_t r = 0; r < (size_t)READS; r++) { seqB[r * (L + 2) + i] = seq[i * READS + r]; psB[r * (L + 2) + i] = ps[i * READS + r]; }
	float *d_tab, *d_em, *d_ps, *d_cs, *d_psB;
	uint8_t *d_seq, *d_seqB;
	CK(cudaMalloc(&d_tab, tab.size() * 4)); CK(cudaMalloc(&d_em, em.size() * 4));
	CK(cudaMalloc(&d_ps, cells * 4)); CK(cudaMalloc(&d_cs, cells * 4)); CK(cudaMalloc(&d_seq, cells));
	CK(cudaMemcpy(d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_em, em.data(), em.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_ps, ps.data(), cells * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_seq, seq.data(), cells, cudaMemcpyHostToDevice));
	CK(cudaMalloc(&d_psB, cells * 4)); CK(cudaMalloc(&d_seqB, cells));
	CK(cudaMemcpy(d_psB, psB.data(), cells * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_seqB, seqB.data(), cells, cudaMemcpyHostToDevice));
	const int smA = (16000 + H * (NC + 1) * 4) * 4;
	const int smB = (16000 + H * (NC + 1) * 4 + 12 + WARPS * 64 * TP) * 4;
	CK(cudaFuncSetAttribute(k_thread, cudaFuncAttributeMaxDynamicSharedMemorySize, smA));
	CK(cudaFuncSetAttribute(k_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, smB));
	const int smC = (16000 + H * (NC + 1) * 4 + 12 + H * PB * 32) * 4;
	CK(cudaFuncSetAttribute(k_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, smC));
	std::vector<float> csC(cells);
	float msC = 0;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	std::vector<float> csA(cells), csB(cells);
	float msA = 0, msB = 0;
	for (int rep = 0; rep < 4; rep++) {
		std::vector<float> init(cells, NEG);
		CK(cudaMemcpy(d_cs, init.data(), cells * 4, cudaMemcpyHostToDevice));
		cudaEventRecord(e0);
		k_thread<<<148, 512, smA>>>(d_tab, d_em, d_seq, d_ps, d_cs, m);
		cudaEventRecord(e1);
		CK(cudaEventSynchronize(e1));
		cudaEventElapsedTime(&msA, e0, e1);
		CK(cudaMemcpy(csA.data(), d_cs, cells * 4, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(d_cs, init.data(), cells * 4, cudaMemcpyHostToDevice));
		cudaEventRecord(e0);
		k_warp<<<148, WARPS * 32, smB>>>(d_tab, d_em, d_seqB, d_psB, d_cs, m, READS);
		cudaEventRecord(e1);
		CK(cudaEventSynchronize(e1));
		cudaEventElapsedTime(&msB, e0, e1);
		CK(cudaMemcpy(csB.data(), d_cs, cells * 4, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(d_cs, init.data(), cells * 4, cudaMemcpyHostToDevice));
		cudaEventRecord(e0);
		k_cta<<<148, CW * 32, smC>>>(d_tab, d_em, d_seq, d_ps, d_cs, m, READS);
		cudaEventRecord(e1);
		CK(cudaEventSynchronize(e1));
		cudaEventElapsedTime(&msC, e0, e1);
		CK(cudaMemcpy(csC.data(), d_cs, cells * 4, cudaMemcpyDeviceToHost));
		size_t diffC = 0;
		for (size_t i = (size_t)READS; i < (size_t)(L + 1) * READS; i++) diffC += memcmp(&csA[i], &csC[i], 4) != 0;
		printf("rep %d: warp-per-HMM x 32 reads (one CTA per 32 reads) %.3f ms (x%.2f); silent rows differing: %zu\n", rep, msC, msC / msA, diffC);
		size_t diff = 0;
		for (size_t i = 1; i <= (size_t)L; i++)
			for (size_t r = 0; r < (size_t)READS; r++) diff += memcmp(&csA[i * READS + r], &csB[r * (L + 2) + i], 4) != 0;
		printf("rep %d: thread-per-read %.3f ms, warp-per-read %.3f ms (x%.2f) per wave of %d reads; silent rows differing: %zu of %zu\n", rep, msA, msB, msB / msA,
		       READS, diff, (size_t)L * READS);
	}
	return 0;
}
