// Host memory bandwidth probe (diagnostic): what can the box's cores write with streaming stores, and how close
// is the uint8 -> int32 widening of the host-buffer path (spl_host_expand.cpp) to that?
#include <immintrin.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char** argv) {
	const size_t n = (size_t)65536 * 297;  // one lock-step of observations at 65,536 envs
	int threads = argc > 1 ? atoi(argv[1]) : omp_get_max_threads();
	uint8_t* src = aligned_alloc(64, n);
	int32_t* dst = aligned_alloc(64, n * 4);
	memset(src, 3, n);
	memset(dst, 0, n * 4);
	for (int mode = 0; mode < 4; mode++) {
		double best = 1e9;
		for (int rep = 0; rep < 20; rep++) {
			double t0 = now();
#pragma omp parallel num_threads(threads)
			{
				int tid = omp_get_thread_num(), nt = omp_get_num_threads();
				size_t lo = n / nt * tid / 64 * 64, hi = tid == nt - 1 ? n : n / nt * (tid + 1) / 64 * 64;
				if (mode == 0) {  // streaming fill (pure write)
					__m256i v = _mm256_set1_epi32(7);
					for (size_t i = lo; i + 8 <= hi; i += 8) _mm256_stream_si256((__m256i*)(dst + i), v);
				} else if (mode == 1) {  // widen, AVX2 streaming stores
					for (size_t i = lo; i + 16 <= hi; i += 16) {
						__m128i a = _mm_loadu_si128((const __m128i*)(src + i));
						_mm256_stream_si256((__m256i*)(dst + i), _mm256_cvtepu8_epi32(a));
						_mm256_stream_si256((__m256i*)(dst + i + 8), _mm256_cvtepu8_epi32(_mm_srli_si128(a, 8)));
					}
				} else if (mode == 2) {  // widen, regular stores
					for (size_t i = lo; i + 16 <= hi; i += 16) {
						__m128i a = _mm_loadu_si128((const __m128i*)(src + i));
						_mm256_storeu_si256((__m256i*)(dst + i), _mm256_cvtepu8_epi32(a));
						_mm256_storeu_si256((__m256i*)(dst + i + 8), _mm256_cvtepu8_epi32(_mm_srli_si128(a, 8)));
					}
				} else {  // memcpy of the same number of output bytes (read + write)
					memcpy(dst + lo, (const char*)dst + (n * 2) + 0 * lo, 0);
					for (size_t i = lo; i + 8 <= hi; i += 8) _mm256_stream_si256((__m256i*)(dst + i), _mm256_loadu_si256((const __m256i*)(dst + ((i + n / 2) % (n - 8)))));
				}
				_mm_sfence();
			}
			double el = now() - t0;
			if (el < best) best = el;
		}
		const char* names[] = {"streaming fill", "widen u8->i32, streaming stores", "widen u8->i32, regular stores", "copy i32->i32, streaming stores"};
		printf("threads=%2d  %-34s %7.1f us  %6.1f GB/s written\n", threads, names[mode], best * 1e6, n * 4 / best / 1e9);
	}
	return 0;
}
