// Microbenchmark (development): how fast can 200-byte (5x5 fp64, 8-byte aligned) blocks be gathered on
// B200 by (0) lane-per-row LDG.64, (1) the same behind prefetch.global.L2, (2) behind
// cp.async.bulk.prefetch.L2, (3) one TMA bulk copy per block into shared memory, (4) 25 lanes per block.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather5 gather5.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while(0)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pf_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ void bulk_pf_l2(const void *p, unsigned bytes) {
	asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}

template <int PF>
__global__ void __launch_bounds__(256, 4)
gather_rows(const long long M, const int *__restrict__ idx, const double *__restrict__ data, double *__restrict__ out)
{
	constexpr int BS = 5, GPW = 6, BS2 = 25;
	const int lane = threadIdx.x & 31, g = lane/BS, r = lane - g*BS;
	const long long warp = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const long long stride = (((long long)gridDim.x*blockDim.x) >> 5)*GPW;
	const bool lv = g < GPW;
	long long t = warp*GPW + g;
	int m0 = -1, m1 = -1, m2 = -1;
	if(lv && t < M) m0 = __ldg(idx + t);
	if(lv && t + stride < M) m1 = __ldg(idx + t + stride);
	const long long niter = (M + stride - 1)/stride;
	for(long long it = 0; it < niter; it++) {
		m2 = -1;
		if(lv && t + 2*stride < M) m2 = __ldg(idx + t + 2*stride);
		if(PF == 1 && m1 >= 0) {
			const char *b = (const char*)(data + (size_t)m1*BS2);
			pf_l2(b + 32*r);                    // bytes 0..160 in steps of 32
			if(r < 2) pf_l2(b + 160 + 32*r);    // 160, 192
		}
		if(PF == 2 && m1 >= 0 && r == 0) {
			const size_t a = (size_t)(data + (size_t)m1*BS2);
			bulk_pf_l2((const void*)(a & ~(size_t)15), 208);
		}
		double s = 0;
		if(m0 >= 0) {
			const double *b = data + (size_t)m0*BS2;
#pragma unroll
			for(int c = 0; c < BS; c++) s += __ldg(b + c*BS + r);
			out[t*BS + r] = s;
		}
		m0 = m1; m1 = m2; t += stride;
	}
}

// 25 lanes per block, one load instruction per block, 4 blocks in flight per warp
__global__ void __launch_bounds__(256, 6)
gather_flat(const long long M, const int *__restrict__ idx, const double *__restrict__ data, double *__restrict__ out)
{
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	const long long nw = ((long long)gridDim.x*blockDim.x) >> 5;
	for(long long t = warp*4; t < M; t += nw*4) {
		double v[4];
#pragma unroll
		for(int q = 0; q < 4; q++) {
			v[q] = 0;
			if(t + q < M && lane < 25) v[q] = __ldg(data + (size_t)__ldg(idx + t + q)*25 + lane);
		}
#pragma unroll
		for(int q = 0; q < 4; q++) {
			double s = v[q];
			s += __shfl_down_sync(0xffffffffu, s, 5); s += __shfl_down_sync(0xffffffffu, s, 10);   // crude row sums
			s += __shfl_down_sync(0xffffffffu, s, 20);
			if(lane < 5 && t + q < M) out[(t+q)*5 + lane] = s;
		}
	}
}

// one TMA bulk copy (BYTES per copy) into shared memory, S stages of NB copies per CTA; the index of
// the copy a thread will issue next is prefetched one iteration ahead (no dependent global load in
// the issue path), so that the measured rate is the TMA unit's, not the index latency's
template <int S, int NB, int BYTES>
__global__ void __launch_bounds__(256)
gather_tma(const long long M, const int *__restrict__ idx, const double *__restrict__ data, double *__restrict__ out)
{
	constexpr int SLOT = BYTES;
	constexpr int BPC = BYTES/208;                 // blocks per copy (runs of consecutive blocks)
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) unsigned long long full[S];
	__shared__ int soff[S][NB];
	const int tid = threadIdx.x;
	if(tid == 0) for(int s = 0; s < S; s++) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&full[s])), "r"(1));
	}
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncthreads();
	const long long ncopies = M/BPC;
	const long long stride = (long long)gridDim.x*NB;
	const long long base = (long long)blockIdx.x*NB;
	const long long niter = (ncopies + stride - 1)/stride;
	auto fetch = [&](long long it) -> int {
		const long long t = base + it*stride + tid;
		return (tid < NB && t < ncopies) ? __ldg(idx + t*BPC) : -1;
	};
	auto issue = [&](long long it, int s, int myidx) {
		const long long t0 = base + it*stride;
		const int cnt = (int)max(0LL, min((long long)NB, ncopies - t0));
		if(tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
		                          :: "r"(smem_u32(&full[s])), "r"(cnt*SLOT) : "memory");
		if(tid < cnt) {
			const size_t a = (size_t)(data + (size_t)myidx*25);
			soff[s][tid] = (int)(a & 15) >> 3;
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			             :: "r"(smem_u32(smem + ((size_t)s*NB + tid)*SLOT)), "l"(a & ~(size_t)15), "r"(SLOT),
			                "r"(smem_u32(&full[s])) : "memory");
		}
	};
	for(int s = 0; s < S && s < niter; s++) { const int i0 = fetch(s); if(tid < 64) issue(s, s, i0); }
	int nxt = fetch(S);
	const int lane = tid & 31, w = tid >> 5, g = lane/5, r = lane - g*5;
	for(long long it = 0; it < niter; it++) {
		const int s = (int)(it % S);
		const unsigned parity = (unsigned)((it / S) & 1);
		const int nxt2 = fetch(it + S + 1);
		asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n"
		             :: "r"(smem_u32(&full[s])), "r"(parity) : "memory");
		// consume: every group sums one block (the first of its copy when BPC > 1 - enough to keep the loads honest)
		const int slot = w*6 + g;
		const long long t = base + it*stride + slot;
		if(g < 6 && slot < NB && t < ncopies) {
			const double *b = (const double*)(smem + ((size_t)s*NB + slot)*SLOT) + soff[s][slot];
			double sum = 0;
#pragma unroll
			for(int q = 0; q < BPC; q++)
#pragma unroll
				for(int c = 0; c < 5; c++) sum += b[q*25 + c*5 + r];
			out[t*5 + r] = sum;
		}
		__syncthreads();
		if(it + S < niter && tid < 64) issue(it + S, s, nxt);
		nxt = nxt2;
	}
}

int main(int argc, char **argv)
{
	const long long N = 16LL << 20;                  // blocks in the array (3.4 GB)
	const long long M = 16LL << 20;                  // blocks gathered
	std::vector<int> h(M);
	double *data, *out; int *idx;
	CK(cudaMalloc(&data, N*200 + 256)); CK(cudaMalloc(&out, M*40)); CK(cudaMalloc(&idx, M*4));
	CK(cudaMemset(data, 0, N*200 + 256));
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	for(int pattern = 0; pattern < 3; pattern++) {
		if(pattern == 0) std::iota(h.begin(), h.end(), 0);                              // streaming
		else if(pattern == 1) for(long long i = 0; i < M; i++) h[i] = (int)((i*7) % N); // stride 7 blocks
		else { std::iota(h.begin(), h.end(), 0); std::mt19937 rng(1); std::shuffle(h.begin(), h.end(), rng); }
		CK(cudaMemcpy(idx, h.data(), M*4, cudaMemcpyHostToDevice));
		const char *pn[3] = {"sequential", "stride7", "random"};
		auto run = [&](const char *name, auto launch) {
			float best = 1e30f;
			for(int rep = 0; rep < 4; rep++) {
				cudaEventRecord(a); launch(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
				float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms);
			}
			CK(cudaGetLastError());
			printf("%-10s %-28s %8.3f ms  %7.0f GB/s (200 B per block + 40 B out)\n", pn[pattern], name, best, M*240.0/best/1e6);
		};
		run("rows LDG.64", [&] { gather_rows<0><<<148*4, 256>>>(M, idx, data, out); });
		run("rows + prefetch.global.L2", [&] { gather_rows<1><<<148*4, 256>>>(M, idx, data, out); });
		run("rows + bulk prefetch L2", [&] { gather_rows<2><<<148*4, 256>>>(M, idx, data, out); });
		run("flat 25 lanes/block", [&] { gather_flat<<<148*6, 256>>>(M, idx, data, out); });
#define TMA_RUN(SV, NBV, BYTESV, CTAS, LABEL) { auto k = gather_tma<SV,NBV,BYTESV>; const int sm = SV*NBV*BYTESV; \
		CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)); \
		run(LABEL, [&] { k<<<148*CTAS, 256, sm>>>(M, idx, data, out); }); }
		TMA_RUN(4, 48, 208, 4, "TMA 208B S=4 x4CTA (idx pf)")
		TMA_RUN(8, 48, 208, 2, "TMA 208B S=8 x2CTA (idx pf)")
		TMA_RUN(3, 48, 624, 2, "TMA 624B(3 blk) S=3 x2CTA")
		TMA_RUN(2, 48, 624, 3, "TMA 624B(3 blk) S=2 x3CTA")
	}
	return 0;
}
