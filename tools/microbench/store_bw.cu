// Store-bandwidth microbenchmark (diagnostic, not product).  Question: a library fill reaches ~7.4 TB/s on this
// part while persistent warps writing 38,016-byte observation tiles with STG.128 stay at 6.0-6.4 TB/s.  Which
// property of the store stream makes the difference?
//   mode 0  persistent warps, each writes whole 38,016-B tiles (the rollout kernel's pattern), .cs
//   mode 1  persistent warps, grid-wide linear front, .cs
//   mode 2  NON-persistent: one 128-thread CTA per 8 KB, default stores, varied data
//   mode 3  as 2 with .cs
//   mode 4  as 2 with constant data (what a fill writes)
//   mode 5  persistent warps, tiles written with TMA bulk stores (cp.async.bulk shared -> global), 4,752-B chunks
//   mode 6  NON-persistent: one 256-thread CTA per 38,016-B tile, .cs
//   mode 7  cudaMemsetAsync
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void tile_store(int4* out, long long ntiles, int mode) {
	const int lane = threadIdx.x & 31;
	const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
	if (mode == 0) {
		for (long long ti = gw; ti < ntiles; ti += nw) {
			int4* g = out + ti * 2376 + lane;
#pragma unroll 4
			for (int q = lane; q < 2376; q += 32, g += 32) __stcs(g, make_int4(q, q + 1, q + 2, q + 3));
		}
	} else {
		const long long total = ntiles * 2376;
		for (long long q = gw * 32 + lane; q < total; q += nw * 32) __stcs(out + q, make_int4((int)q, 1, 2, 3));
	}
}

template <int MODE>
__global__ void __launch_bounds__(128) small_cta_store(int4* out, long long total) {
	long long base = (long long)blockIdx.x * 512 + threadIdx.x;
#pragma unroll
	for (int i = 0; i < 4; i++) {
		long long q = base + i * 128;
		if (q < total) {
			int4 v = MODE == 4 ? make_int4(7, 7, 7, 7) : make_int4((int)q, 1, 2, 3);
			if (MODE == 3) __stcs(out + q, v);
			else out[q] = v;
		}
	}
}

__global__ void __launch_bounds__(256) tile_cta_store(int4* out, long long ntiles) {
	int4* g = out + (long long)blockIdx.x * 2376;
	for (int q = threadIdx.x; q < 2376; q += 256) __stcs(g + q, make_int4(q, q + 1, q + 2, q + 3));
}

// TMA bulk store: each warp owns 4,752 B of shared memory (1/8 tile), fills it once, then issues bulk stores.
__global__ void tma_store(int4* out, long long ntiles, int depth) {
	extern __shared__ __align__(128) unsigned char smem[];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
	unsigned char* my = smem + w * 4752 * 2;
	for (int i = lane; i < 4752 * 2 / 16; i += 32) ((int4*)my)[i] = make_int4(i, i + 1, i + 2, i + 3);
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	__syncwarp();
	for (long long ti = gw; ti < ntiles; ti += nw) {
		unsigned char* g = (unsigned char*)(out + ti * 2376);
		if (lane == 0) {
			for (int c = 0; c < 8; c++) {
				unsigned s = (unsigned)__cvta_generic_to_shared(my + (c & 1) * 4752);
				asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g + c * 4752), "r"(s), "r"(4752) : "memory");
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
				if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
				else asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
			}
		}
		__syncwarp();
	}
	if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}


// mode 8/9: persistent with an atomic work counter (in-order front like the hardware CTA scheduler)
__global__ void __launch_bounds__(128) dyn_chunk_store(int4* out, long long nchunks, unsigned long long* ctr) {
	__shared__ unsigned long long s;
	for (;;) {
		if (threadIdx.x == 0) s = atomicAdd(ctr, 1ull);
		__syncthreads();
		unsigned long long c = s;
		__syncthreads();
		if (c >= (unsigned long long)nchunks) break;
		int4* g = out + c * 512 + threadIdx.x;
#pragma unroll
		for (int i = 0; i < 4; i++) __stcs(g + i * 128, make_int4((int)c, i, 2, 3));
	}
}
__global__ void dyn_tile_store(int4* out, long long ntiles, unsigned long long* ctr) {
	const int lane = threadIdx.x & 31;
	for (;;) {
		unsigned long long t = 0;
		if (lane == 0) t = atomicAdd(ctr, 1ull);
		t = __shfl_sync(0xffffffffu, t, 0);
		if (t >= (unsigned long long)ntiles) break;
		int4* g = out + t * 2376 + lane;
#pragma unroll 4
		for (int q = lane; q < 2376; q += 32, g += 32) __stcs(g, make_int4(q, q + 1, q + 2, q + 3));
	}
}
// mode 10: non-persistent, one WARP-sized CTA per tile; mode 11: non-persistent small CTAs in a scrambled order
__global__ void __launch_bounds__(32) warp_cta_store(int4* out) {
	int4* g = out + (long long)blockIdx.x * 2376 + threadIdx.x;
#pragma unroll 4
	for (int q = threadIdx.x; q < 2376; q += 32, g += 32) __stcs(g, make_int4(q, q + 1, q + 2, q + 3));
}
__global__ void __launch_bounds__(128) scrambled_cta_store(int4* out, long long nchunks, unsigned long long mul) {
	unsigned long long c = ((unsigned long long)blockIdx.x * mul) % (unsigned long long)nchunks;
	int4* g = out + c * 512 + threadIdx.x;
#pragma unroll
	for (int i = 0; i < 4; i++) __stcs(g + i * 128, make_int4((int)c, i, 2, 3));
}
// mode 12: persistent warps, tile per warp, but the warps of the whole grid re-synchronise every tile (cooperative-free:
// spin on a global generation counter) -- bounds the drift between warps to one tile
__global__ void synced_tile_store(int4* out, long long ntiles, unsigned int* gen) {
	const int lane = threadIdx.x & 31;
	const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
	unsigned int it = 0;
	const long long iters = (ntiles + nw - 1) / nw;
	for (long long ti = gw; it < iters; ti += nw, it++) {
		int4* g = out + ti * 2376 + lane;
		if (ti < ntiles) {
#pragma unroll 4
			for (int q = lane; q < 2376; q += 32, g += 32) __stcs(g, make_int4(q, q + 1, q + 2, q + 3));
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			atomicAdd(gen, 1u);
			unsigned int target = (it + 1) * gridDim.x;
			while (*(volatile unsigned int*)gen < target) {}
		}
		__syncthreads();
	}
}


// mode 13: static tile assignment (as mode 0) but every warp delays its start by a pseudo-random time (time de-synchronisation)
// mode 14: static assignment, simultaneous start, but every warp starts its tile at a different chunk (address-phase de-sync)
// mode 15: static assignment + a pseudo-random compute delay before every tile (what the rules code does to the real kernel)
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void variant_tile_store(int4* out, long long ntiles, int mode, int delay_cycles) {
	const int lane = threadIdx.x & 31;
	const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
	if (mode == 13) {
		long long t0 = clock64(), d = hash32((unsigned)gw) % (unsigned)delay_cycles;
		while (clock64() - t0 < d) {}
	}
	unsigned it = 0;
	for (long long ti = gw; ti < ntiles; ti += nw, it++) {
		if (mode == 15) {
			long long t0 = clock64(), d = hash32((unsigned)gw * 7919u + it) % (unsigned)delay_cycles;
			while (clock64() - t0 < d) {}
		}
		int4* g = out + ti * 2376;
		if (mode == 14) {
			int k0 = (hash32((unsigned)gw) % 74u) * 32;
			for (int j = 0; j < 2368; j += 32) {
				int q = k0 + j; q -= q >= 2368 ? 2368 : 0;
				__stcs(g + q + lane, make_int4(q, q + 1, q + 2, q + 3));
			}
			if (lane < 8) __stcs(g + 2368 + lane, make_int4(1, 2, 3, 4));
		} else {
			g += lane;
#pragma unroll 4
			for (int q = lane; q < 2376; q += 32, g += 32) __stcs(g, make_int4(q, q + 1, q + 2, q + 3));
		}
	}
}


// ---- load-balance hypothesis: are some SMs slower at storing than others?
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long r; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(r)); return r; }
// mode 16: dynamic tiles (as 9), counts tiles per SM.  mode 18: dynamic, but the order inside each window of `win` tiles is scrambled
__global__ void dyn_tile_count(int4* out, long long ntiles, unsigned long long* ctr, unsigned* per_sm, int win, unsigned mul) {
	const int lane = threadIdx.x & 31;
	unsigned n = 0;
	for (;;) {
		unsigned long long t = 0;
		if (lane == 0) t = atomicAdd(ctr, 1ull);
		t = __shfl_sync(0xffffffffu, t, 0);
		if (t >= (unsigned long long)ntiles) break;
		if (win > 1) { unsigned long long w0 = t / win * win; unsigned r = (unsigned)(t - w0); if (w0 + win <= (unsigned long long)ntiles) t = w0 + (unsigned long long)r * mul % win; }
		int4* g = out + t * 2376 + lane;
#pragma unroll 4
		for (int q = lane; q < 2376; q += 32, g += 32) __stcs(g, make_int4(q, q + 1, q + 2, q + 3));
		n++;
	}
	if (lane == 0) atomicAdd(per_sm + smid(), n);
}
// mode 17: static tiles (as 0), records when each SM's last warp finished
__global__ void static_tile_time(int4* out, long long ntiles, unsigned long long* fin) {
	const int lane = threadIdx.x & 31;
	const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
	for (long long ti = gw; ti < ntiles; ti += nw) {
		int4* g = out + ti * 2376 + lane;
#pragma unroll 4
		for (int q = lane; q < 2376; q += 32, g += 32) __stcs(g, make_int4(q, q + 1, q + 2, q + 3));
	}
	if (lane == 0) atomicMax(fin + smid(), gtime());
	if (gw == 0 && lane == 0) {}
}
// mode 19: static, but every SM owns a contiguous range of tiles proportional to the quota measured by mode 16
__global__ void quota_tile_store(int4* out, const unsigned* start, unsigned* slot_ctr, int wps) {
	const int lane = threadIdx.x & 31;
	unsigned s = smid(), j = 0;
	if (lane == 0) j = atomicAdd(slot_ctr + s, 1u);
	j = __shfl_sync(0xffffffffu, j, 0);
	for (unsigned ti = start[s] + j; ti < start[s + 1]; ti += wps) {
		int4* g = out + (long long)ti * 2376 + lane;
#pragma unroll 4
		for (int q = lane; q < 2376; q += 32, g += 32) __stcs(g, make_int4(q, q + 1, q + 2, q + 3));
	}
}

static float best_of(void (*launch)(void*), void* ctx, int reps) {
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	launch(ctx);
	cudaDeviceSynchronize();
	float best = 1e9;
	for (int r = 0; r < reps; r++) {
		cudaEventRecord(e0);
		launch(ctx);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		best = ms < best ? ms : best;
	}
	return best;
}

struct Ctx {
	int4* out;
	long long ntiles;
	int mode, grid, tpb, depth;
};

int main(int argc, char** argv) {
	setvbuf(stdout, NULL, _IOLBF, 0);
	const long long ntiles = 262144;  // 9.96 GB
	const long long total = ntiles * 2376;
	const double bytes = total * 16.0;
	int4* out;
	cudaMalloc(&out, total * sizeof(int4));
	Ctx c{out, ntiles, 0, 0, 0, 0};
	for (int mode = 0; mode < 2; mode++)
		for (int wps : {8, 20, 64}) {
			c.mode = mode;
			c.tpb = 128;
			c.grid = 148 * wps * 32 / 128;
			float ms = best_of([](void* p) { Ctx* c = (Ctx*)p; tile_store<<<c->grid, c->tpb>>>(c->out, c->ntiles, c->mode); }, &c, 3);
			printf("mode=%d persistent warps/SM=%2d  %.0f GB/s\n", mode, wps, bytes / (ms * 1e-3) / 1e9);
		}
	c.grid = (int)((total + 511) / 512);
	float ms = best_of([](void* p) { Ctx* c = (Ctx*)p; small_cta_store<2><<<c->grid, 128>>>(c->out, c->ntiles * 2376); }, &c, 5);
	printf("mode=2 small CTAs default stores varied data  %.0f GB/s\n", bytes / (ms * 1e-3) / 1e9);
	ms = best_of([](void* p) { Ctx* c = (Ctx*)p; small_cta_store<3><<<c->grid, 128>>>(c->out, c->ntiles * 2376); }, &c, 5);
	printf("mode=3 small CTAs .cs stores varied data      %.0f GB/s\n", bytes / (ms * 1e-3) / 1e9);
	ms = best_of([](void* p) { Ctx* c = (Ctx*)p; small_cta_store<4><<<c->grid, 128>>>(c->out, c->ntiles * 2376); }, &c, 5);
	printf("mode=4 small CTAs default stores constant     %.0f GB/s\n", bytes / (ms * 1e-3) / 1e9);
	for (int depth : {1, 4})
		for (int wps : {4, 8, 16}) {
			c.depth = depth;
			c.tpb = 128;
			c.grid = 148 * wps * 32 / 128;
			cudaFuncSetAttribute(tma_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4752 * 2);
			ms = best_of([](void* p) { Ctx* c = (Ctx*)p; tma_store<<<c->grid, c->tpb, 4 * 4752 * 2>>>(c->out, c->ntiles, c->depth); }, &c, 3);
			printf("mode=5 TMA bulk stores warps/SM=%2d depth=%d  %.0f GB/s  (%s)\n", wps, depth, bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
		}
	c.grid = (int)ntiles;
	ms = best_of([](void* p) { Ctx* c = (Ctx*)p; tile_cta_store<<<c->grid, 256>>>(c->out, c->ntiles); }, &c, 5);
	printf("mode=6 one CTA per 38,016-B tile .cs          %.0f GB/s\n", bytes / (ms * 1e-3) / 1e9);
	ms = best_of([](void* p) { Ctx* c = (Ctx*)p; cudaMemsetAsync(c->out, 7, c->ntiles * 2376 * 16); }, &c, 5);
	printf("mode=7 cudaMemsetAsync                         %.0f GB/s\n", bytes / (ms * 1e-3) / 1e9);

	unsigned long long* ctr;
	cudaMalloc(&ctr, 8);
	static unsigned long long* g_ctr; g_ctr = ctr;
	for (int cps : {4, 8, 16}) {
		c.grid = 148 * cps;
		ms = best_of([](void* p) { Ctx* c = (Ctx*)p; cudaMemsetAsync(g_ctr, 0, 8); dyn_chunk_store<<<c->grid, 128>>>(c->out, c->ntiles * 2376 / 512, g_ctr); }, &c, 3);
		printf("mode=8 persistent, atomic counter, 8 KB chunks per 128-thread CTA, CTAs/SM=%2d  %.0f GB/s\n", cps, bytes / (ms * 1e-3) / 1e9);
	}
	for (int wps : {8, 20, 64}) {
		c.grid = 148 * wps / 4;
		ms = best_of([](void* p) { Ctx* c = (Ctx*)p; cudaMemsetAsync(g_ctr, 0, 8); dyn_tile_store<<<c->grid, 128>>>(c->out, c->ntiles, g_ctr); }, &c, 3);
		printf("mode=9 persistent, atomic counter, one tile per warp, warps/SM=%2d  %.0f GB/s\n", wps, bytes / (ms * 1e-3) / 1e9);
	}
	ms = best_of([](void* p) { Ctx* c = (Ctx*)p; warp_cta_store<<<(int)c->ntiles, 32>>>(c->out); }, &c, 3);
	printf("mode=10 non-persistent, one 32-thread CTA per tile  %.0f GB/s\n", bytes / (ms * 1e-3) / 1e9);
	c.grid = (int)(total / 512);
	for (unsigned long long mul : {1ull, 1000003ull, 2654435761ull}) {
		static unsigned long long g_mul; g_mul = mul;
		ms = best_of([](void* p) { Ctx* c = (Ctx*)p; scrambled_cta_store<<<c->grid, 128>>>(c->out, c->ntiles * 2376 / 512, g_mul); }, &c, 3);
		printf("mode=11 non-persistent small CTAs, order multiplier %llu  %.0f GB/s\n", mul, bytes / (ms * 1e-3) / 1e9);
	}
	for (int wps : {8, 20}) {
		c.grid = 148 * wps / 4;
		ms = best_of([](void* p) { Ctx* c = (Ctx*)p; cudaMemsetAsync(g_ctr, 0, 8); synced_tile_store<<<c->grid, 128>>>(c->out, c->ntiles, (unsigned int*)g_ctr); }, &c, 3);
		printf("mode=12 persistent, tile per warp, grid re-sync per tile, warps/SM=%2d  %.0f GB/s\n", wps, bytes / (ms * 1e-3) / 1e9);
	}

	for (int mode : {13, 14, 15})
		for (int wps : {8, 20})
			for (int dc : {8000, 30000}) {
				if (mode == 14 && dc != 8000) continue;
				c.mode = mode; c.depth = dc; c.grid = 148 * wps / 4;
				ms = best_of([](void* p) { Ctx* c = (Ctx*)p; variant_tile_store<<<c->grid, 128>>>(c->out, c->ntiles, c->mode, c->depth); }, &c, 3);
				printf("mode=%d static tiles, %s, warps/SM=%2d delay<=%d cycles  %.0f GB/s\n", mode, mode == 13 ? "staggered start" : mode == 14 ? "rotated chunk order" : "random delay per tile", wps, dc, bytes / (ms * 1e-3) / 1e9);
			}
	// the rollout kernel's shape at 65,536 envs: 2048 one-warp CTAs, tile (t, w) at t*2048 + w
	for (int mode : {0, 14, 15}) {
		c.mode = mode; c.depth = 8000; c.grid = 2048;
		ms = best_of([](void* p) { Ctx* c = (Ctx*)p; variant_tile_store<<<c->grid, 32>>>(c->out, c->ntiles, c->mode, c->depth); }, &c, 3);
		printf("mode=%d 2048 one-warp CTAs (65,536-env rollout shape)  %.0f GB/s\n", mode, bytes / (ms * 1e-3) / 1e9);
	}

	{
		unsigned* per_sm; cudaMalloc(&per_sm, 256 * 4);
		unsigned long long* fin; cudaMalloc(&fin, 256 * 8);
		unsigned h[256]; unsigned long long hf[256];
		cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
		for (int wps : {8, 20}) {
			int grid = 148 * wps / 4;
			for (int win : {1, 2960}) {
				cudaMemset(per_sm, 0, 1024); cudaMemset(ctr, 0, 8);
				cudaEventRecord(e0);
				dyn_tile_count<<<grid, 128>>>(out, ntiles, ctr, per_sm, win, 1103u);
				cudaEventRecord(e1); cudaEventSynchronize(e1);
				float m; cudaEventElapsedTime(&m, e0, e1);
				cudaMemcpy(h, per_sm, 1024, cudaMemcpyDeviceToHost);
				unsigned mn = ~0u, mx = 0; for (int i = 0; i < 148; i++) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; }
				printf("mode=%d dynamic tiles%s warps/SM=%2d  %.0f GB/s; tiles per SM min %u max %u mean %.0f\n  per SM:", win > 1 ? 18 : 16, win > 1 ? " (scrambled inside 2960-tile windows)" : "", wps, bytes / (m * 1e-3) / 1e9, mn, mx, ntiles / 148.0);
				for (int i = 0; i < 148; i++) printf(" %u", h[i]);
				printf("\n");
			}
			// static with timing
			cudaMemset(fin, 0, 2048);
			unsigned long long t_begin;
			cudaDeviceSynchronize();
			cudaEventRecord(e0);
			static_tile_time<<<grid, 128>>>(out, ntiles, fin);
			cudaEventRecord(e1); cudaEventSynchronize(e1);
			float m; cudaEventElapsedTime(&m, e0, e1);
			cudaMemcpy(hf, fin, 2048, cudaMemcpyDeviceToHost);
			unsigned long long fmin = ~0ull, fmax = 0; for (int i = 0; i < 148; i++) { fmin = hf[i] < fmin ? hf[i] : fmin; fmax = hf[i] > fmax ? hf[i] : fmax; }
			(void)t_begin;
			printf("mode=17 static tiles warps/SM=%2d  %.0f GB/s (%.3f ms); SM finish times relative to the last one (us before the end):\n ", wps, bytes / (m * 1e-3) / 1e9, m);
			for (int i = 0; i < 148; i++) printf(" %llu", (fmax - hf[i]) / 1000);
			printf("\n");
			// quota-weighted static
			unsigned start[257]; start[0] = 0; for (int i = 0; i < 148; i++) start[i + 1] = start[i] + h[i];
			// h currently holds the scrambled-window run; rerun plain dynamic for quotas
			cudaMemset(per_sm, 0, 1024); cudaMemset(ctr, 0, 8);
			dyn_tile_count<<<grid, 128>>>(out, ntiles, ctr, per_sm, 1, 1u);
			cudaMemcpy(h, per_sm, 1024, cudaMemcpyDeviceToHost);
			start[0] = 0; for (int i = 0; i < 148; i++) start[i + 1] = start[i] + h[i];
			unsigned* dstart; cudaMalloc(&dstart, 257 * 4); cudaMemcpy(dstart, start, 149 * 4, cudaMemcpyHostToDevice);
			float bestq = 1e9;
			for (int r = 0; r < 3; r++) {
				cudaMemset(per_sm, 0, 1024);
				cudaEventRecord(e0);
				quota_tile_store<<<grid, 128>>>(out, dstart, per_sm, wps);
				cudaEventRecord(e1); cudaEventSynchronize(e1);
				cudaEventElapsedTime(&m, e0, e1); bestq = m < bestq ? m : bestq;
			}
			printf("mode=19 static, contiguous per-SM ranges sized by the dynamic run's quotas, warps/SM=%2d  %.0f GB/s\n", wps, bytes / (bestq * 1e-3) / 1e9);
			// equal quotas, same per-SM contiguous layout (control)
			for (int i = 0; i <= 148; i++) start[i] = (unsigned)(ntiles * i / 148);
			cudaMemcpy(dstart, start, 149 * 4, cudaMemcpyHostToDevice);
			bestq = 1e9;
			for (int r = 0; r < 3; r++) {
				cudaMemset(per_sm, 0, 1024);
				cudaEventRecord(e0);
				quota_tile_store<<<grid, 128>>>(out, dstart, per_sm, wps);
				cudaEventRecord(e1); cudaEventSynchronize(e1);
				cudaEventElapsedTime(&m, e0, e1); bestq = m < bestq ? m : bestq;
			}
			printf("mode=19c control: contiguous per-SM ranges of EQUAL size, warps/SM=%2d  %.0f GB/s\n", wps, bytes / (bestq * 1e-3) / 1e9);
		}
	}
	printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
