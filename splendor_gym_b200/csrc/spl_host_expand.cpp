// spl_host_expand.cpp -- host half of the host-buffer path (include/splendor_b200.h, spl_host_step):
// widen the compact device->host records into the reference-typed arrays of SplendorEnv.step
// (envs/splendor_env.py:51-90: int32 observation, int8 action mask, float reward, bool terminated).
// Pure format conversion, no game logic: every value was computed by the CUDA step kernel.
//
// The work is memory-bound (1,243 B written per env-step), so: OpenMP over blocks of envs, AVX2 zero-extension
// with non-temporal stores for the observation (no read-for-ownership of the destination lines), a 256-entry
// bits->bytes table for the mask.
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <omp.h>
#include <sched.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/splendor_b200.h"

namespace {

const float kRewardOfCode[8] = {0.0f, 1.0f, -1.0f, -0.1f, -0.01f, 0.0f, 0.0f, 0.0f};

struct BitTable {
	uint64_t v[256];
	BitTable() {
		for (int b = 0; b < 256; b++) {
			uint64_t x = 0;
			for (int k = 0; k < 8; k++) x |= (uint64_t)((b >> k) & 1) << (8 * k);
			v[b] = x;
		}
	}
};
const BitTable kBits;

void widen_scalar(const uint8_t* src, int32_t* dst, size_t n) {
	for (size_t i = 0; i < n; i++) dst[i] = (int32_t)src[i];
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void widen_avx2(const uint8_t* src, int32_t* dst, size_t n) {
	size_t i = 0;
	// head: bring dst to a 32-byte boundary for the streaming stores
	while (i < n && ((uintptr_t)(dst + i) & 31)) {
		dst[i] = (int32_t)src[i];
		i++;
	}
	for (; i + 32 <= n; i += 32) {
		const __m128i lo = _mm_loadu_si128((const __m128i*)(src + i));
		const __m128i hi = _mm_loadu_si128((const __m128i*)(src + i + 16));
		_mm256_stream_si256((__m256i*)(dst + i), _mm256_cvtepu8_epi32(lo));
		_mm256_stream_si256((__m256i*)(dst + i + 8), _mm256_cvtepu8_epi32(_mm_srli_si128(lo, 8)));
		_mm256_stream_si256((__m256i*)(dst + i + 16), _mm256_cvtepu8_epi32(hi));
		_mm256_stream_si256((__m256i*)(dst + i + 24), _mm256_cvtepu8_epi32(_mm_srli_si128(hi, 8)));
	}
	for (; i < n; i++) dst[i] = (int32_t)src[i];
	_mm_sfence();
}
__attribute__((target("avx512f"))) void widen_avx512(const uint8_t* src, int32_t* dst, size_t n) {
	size_t i = 0;
	while (i < n && ((uintptr_t)(dst + i) & 63)) {  // head: 64-byte boundary, then one full cache line per streaming store
		dst[i] = (int32_t)src[i];
		i++;
	}
	for (; i + 64 <= n; i += 64) {
		const __m128i a = _mm_loadu_si128((const __m128i*)(src + i));
		const __m128i b = _mm_loadu_si128((const __m128i*)(src + i + 16));
		const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32));
		const __m128i d = _mm_loadu_si128((const __m128i*)(src + i + 48));
		_mm512_stream_si512((__m512i*)(dst + i), _mm512_cvtepu8_epi32(a));
		_mm512_stream_si512((__m512i*)(dst + i + 16), _mm512_cvtepu8_epi32(b));
		_mm512_stream_si512((__m512i*)(dst + i + 32), _mm512_cvtepu8_epi32(c));
		_mm512_stream_si512((__m512i*)(dst + i + 48), _mm512_cvtepu8_epi32(d));
	}
	for (; i < n; i++) dst[i] = (int32_t)src[i];
	_mm_sfence();
}
int simd_level() {  // 0 scalar, 2 AVX2, 5 AVX-512 (SPL_HOST_SIMD overrides downwards)
	static const int v = [] {
		int lvl = __builtin_cpu_supports("avx512f") ? 5 : (__builtin_cpu_supports("avx2") ? 2 : 0);
		const char* e = getenv("SPL_HOST_SIMD");
		if (e && atoi(e) < lvl) lvl = atoi(e);
		return lvl;
	}();
	return v;
}
#endif

inline void widen(const uint8_t* src, int32_t* dst, size_t n) {
#if defined(__x86_64__)
	if (simd_level() >= 5) return widen_avx512(src, dst, n);
	if (simd_level() >= 2) return widen_avx2(src, dst, n);
#endif
	widen_scalar(src, dst, n);
}

int g_threads = 0;

}  // namespace

// widen envs [lo, hi): obs_u8 / side are the staging arrays (indexed by env), outputs are the caller's arrays
static void expand_block(const uint8_t* obs_u8, const uint32_t* side, int64_t lo, int64_t hi, const spl_host_io_t* io) {
	if (io->obs) widen(obs_u8 + lo * 297, io->obs + lo * 297, (size_t)(hi - lo) * 297);
	if (io->obs_u8 && io->obs_u8 != obs_u8) memcpy(io->obs_u8 + lo * 297, obs_u8 + lo * 297, (size_t)(hi - lo) * 297);
	for (int64_t i = lo; i < hi; i++) {
		const uint32_t x = side[4 * i], y = side[4 * i + 1], z = side[4 * i + 2];
		if (io->mask) {
			const uint64_t m = (uint64_t)x | ((uint64_t)(y & 0x1FFFu) << 32);
			int8_t* row = io->mask + i * 45;
			uint64_t w;
			for (int q = 0; q < 5; q++) {
				w = kBits.v[(m >> (8 * q)) & 0xFF];
				memcpy(row + 8 * q, &w, 8);
			}
			w = kBits.v[(m >> 40) & 0x1F];
			memcpy(row + 40, &w, 5);
		}
		if (io->reward) io->reward[i] = kRewardOfCode[(y >> 16) & 7];
		if (io->terminated) io->terminated[i] = (uint8_t)((y >> 24) & 1);
		if (io->info) io->info[i] = (uint8_t)(z & 0xFF);
		if (io->next_action) io->next_action[i] = (int32_t)((z >> 8) & 0xFF);
	}
}

// All chunks of one lock-step in ONE parallel region (a fork/join per chunk costs more than widening a chunk):
// thread 0 waits for chunk c's copy (`wait(ctx, c)`, a cudaEventSynchronize) and publishes it; the others spin on the
// counter; then every thread widens its share of the chunk.  bounds[c]..bounds[c+1] = envs of chunk c.
void spl_expand_chunks(const uint8_t* obs_u8, const uint32_t* side, const int64_t* bounds, int chunks, const spl_host_io_t* io,
                       int (*wait)(void*, int), void* ctx, int* rc_out) {
	const int64_t blk = 64;  // envs per task: 19 KB of observation bytes in, 76 KB out
	const int nthreads = g_threads > 0 ? g_threads : omp_get_max_threads();
	int ready = 0, rc = 0;
#pragma omp parallel num_threads(nthreads)
	{
		const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
		for (int c = 0; c < chunks; c++) {
			if (tid == 0) {
				int r = wait ? wait(ctx, c) : 0;
				if (r) __atomic_store_n(&rc, r, __ATOMIC_RELAXED);
				__atomic_store_n(&ready, c + 1, __ATOMIC_RELEASE);
			} else {
				for (unsigned spins = 0; __atomic_load_n(&ready, __ATOMIC_ACQUIRE) <= c; spins++) {
#if defined(__x86_64__)
					_mm_pause();
#endif
					if ((spins & 1023u) == 1023u) sched_yield();  // oversubscribed hosts: let thread 0 run
				}
			}
			if (__atomic_load_n(&rc, __ATOMIC_RELAXED)) continue;
			const int64_t b = bounds[c], e = bounds[c + 1];
			const int64_t nblk = (e - b + blk - 1) / blk;
			// contiguous share per thread (streams well); thread 0 (which also waits on the device) gets the tail
			const int64_t per = (nblk + nt - 1) / nt;
			const int64_t k0 = per * ((tid + nt - 1) % nt), k1 = k0 + per < nblk ? k0 + per : nblk;
			for (int64_t k = k0; k < k1; k++) {
				const int64_t lo = b + k * blk, hi = lo + blk < e ? lo + blk : e;
				expand_block(obs_u8, side, lo, hi, io);
			}
		}
	}
	if (rc_out) *rc_out = rc;
}

void spl_expand_range(const uint8_t* obs_u8, const uint32_t* side, int64_t b, int64_t e, const spl_host_io_t* io) {
	const int64_t bounds[2] = {b, e};
	spl_expand_chunks(obs_u8, side, bounds, 1, io, nullptr, nullptr, nullptr);
}

void spl_parallel_copy(void* dst, const void* src, size_t bytes) {
	const size_t blk = 1 << 16;
	const int64_t nblk = (int64_t)((bytes + blk - 1) / blk);
	const int nthreads = g_threads > 0 ? g_threads : omp_get_max_threads();
	if (nblk <= 4) {
		memcpy(dst, src, bytes);
		return;
	}
#pragma omp parallel for schedule(static) num_threads(nthreads)
	for (int64_t k = 0; k < nblk; k++) {
		const size_t o = (size_t)k * blk;
		memcpy((char*)dst + o, (const char*)src + o, o + blk <= bytes ? blk : bytes - o);
	}
}

extern "C" int spl_host_expand(const uint8_t* obs_u8, const void* side, int64_t n, const spl_host_io_t* io) {
	if (!side || !io || n <= 0 || ((io->obs || io->obs_u8) && !obs_u8)) return SPL_E_BADARG;
	spl_expand_range(obs_u8, (const uint32_t*)side, 0, n, io);
	return 0;
}

extern "C" int spl_host_set_threads(int n) {
	if (n > 0) g_threads = n;
	return g_threads > 0 ? g_threads : omp_get_max_threads();
}
