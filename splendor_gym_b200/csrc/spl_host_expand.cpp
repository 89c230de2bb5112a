// spl_host_expand.cpp -- host half of the host-buffer path (include/splendor_b200.h, spl_host_step):
// widen the compact device->host records into the reference-typed arrays of SplendorEnv.step
// (envs/splendor_env.py:51-90: int32 observation, int8 action mask, float reward, bool terminated).
// Pure format conversion, no game logic: every value was computed by the CUDA step kernel.
//
// The work is memory-bound (1,243 B written per env-step), so: a pool of pinned worker threads that own contiguous
// ranges of envs, AVX-512 / AVX2 zero-extension with non-temporal stores for the observation (no read-for-ownership
// of the destination lines), a 256-entry bits->bytes table for the mask.  Workers find their input by polling arrival
// flags the GPU writes into pinned host memory behind each 64-env group (spl_host.cu, spl_push_kernel).
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>

#include <atomic>

#include <sched.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/splendor_b200.h"
#include "spl_host_pool.h"

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
__attribute__((target("avx512f"))) void fill_avx512(void* dst, size_t bytes) {
	const __m512i v = _mm512_set1_epi32(1);
	char* p = (char*)dst;
	for (size_t i = 0; i + 64 <= bytes; i += 64) _mm512_stream_si512((__m512i*)(p + i), v);
	_mm_sfence();
}
__attribute__((target("avx2"))) void stream_lines(void* dst, const void* src, size_t bytes) {
	for (size_t i = 0; i < bytes; i += 32) _mm256_stream_si256((__m256i*)((char*)dst + i), _mm256_loadu_si256((const __m256i*)((const char*)src + i)));
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

inline void cpu_relax() {
#if defined(__x86_64__)
	_mm_pause();
#endif
}

}  // namespace

double spl_now_us() {
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec * 1e6 + (double)ts.tv_nsec * 1e-3;
}

// whole cache lines to a 64-byte aligned destination without reading them first; anything else through memcpy
static void stream_copy(void* dst, const void* src, size_t bytes) {
#if defined(__x86_64__)
	if ((((uintptr_t)dst | bytes) & 63) == 0 && simd_level() >= 2) {
		stream_lines(dst, src, bytes);
		return;
	}
#endif
	memcpy(dst, src, bytes);
}

// widen envs [lo, hi), hi - lo <= SPL_HOST_GROUP: obs_lo / side_lo point at the observation bytes / records of env `lo`,
// outputs are the caller's arrays (indexed by env).  The small outputs of a group are assembled in L1 and leave as whole cache lines too (a full
// group is 2,880 B of mask, 256 B of rewards, ...: no read-for-ownership of the destination when it is 64-byte aligned).
void spl_expand_block(const uint8_t* obs_lo, const uint32_t* side_lo, int64_t lo, int64_t hi, const spl_host_io_t* io) {
	if (io->obs) widen(obs_lo, io->obs + lo * 297, (size_t)(hi - lo) * 297);
	if (io->obs_u8 && io->obs_u8 + lo * 297 != obs_lo) stream_copy(io->obs_u8 + lo * 297, obs_lo, (size_t)(hi - lo) * 297);
	alignas(64) int8_t mask[SPL_HOST_GROUP * 45 + 8];
	alignas(64) float reward[SPL_HOST_GROUP];
	alignas(64) uint8_t term[SPL_HOST_GROUP], info[SPL_HOST_GROUP];
	alignas(64) int32_t next[SPL_HOST_GROUP];
	const int m = (int)(hi - lo);
	for (int i = 0; i < m; i++) {
		const uint32_t* rec = side_lo + 4 * i;
		const uint32_t x = rec[0], y = rec[1], z = rec[2];
		if (io->mask) {
			const uint64_t mk = (uint64_t)x | ((uint64_t)(y & 0x1FFFu) << 32);
			int8_t* row = mask + i * 45;
			for (int q = 0; q < 6; q++) {  // the last 8-byte store spills 3 bytes into the next row, which is written after it
				const uint64_t w = kBits.v[(mk >> (8 * q)) & 0xFF];
				memcpy(row + 8 * q, &w, 8);
			}
		}
		reward[i] = kRewardOfCode[(y >> 16) & 7];
		term[i] = (uint8_t)((y >> 24) & 1);
		info[i] = (uint8_t)(z & 0xFF);
		next[i] = (int32_t)((z >> 8) & 0xFF);
	}
	if (io->mask) stream_copy(io->mask + lo * 45, mask, (size_t)m * 45);
	if (io->reward) stream_copy(io->reward + lo, reward, (size_t)m * 4);
	if (io->terminated) stream_copy(io->terminated + lo, term, (size_t)m);
	if (io->info) stream_copy(io->info + lo, info, (size_t)m);
	if (io->next_action) stream_copy(io->next_action + lo, next, (size_t)m * 4);
}

// ------------------------------------------------------------------------------------------------
// Worker pool.  One job = one lock-step: worker j owns the contiguous groups [start_j, start_j + len_j) of the
// CPU share and widens group g as soon as flags[g] == seq.  Workers spin between jobs for a while (a vector loop
// calls spl_host_step back to back), then sleep on a condition variable.
// ------------------------------------------------------------------------------------------------
struct SplPool {
	int threads = 0;  // dedicated workers; the caller only waits (and watches the stream for errors)
	pthread_t tid[SPL_POOL_MAX];
	int index[SPL_POOL_MAX];
	std::atomic<uint64_t> generation{0};
	std::atomic<int> done{0};
	std::atomic<int> stop{0};
	std::atomic<int> sleepers{0};
	pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
	pthread_cond_t cv = PTHREAD_COND_INITIALIZER;
	SplHostJob job;
	int cpus[SPL_POOL_MAX];  // core of worker j (-1: not pinned)
};

static SplPool* g_pool = nullptr;
static int g_threads = 0;
static int g_pin = 1;

void spl_job_share(const SplHostJob* job, int j, int64_t* start, int64_t* len) {
	const int64_t base = job->cpu_groups / job->threads, rem = job->cpu_groups % job->threads;
	*start = (int64_t)j * base + (j < rem ? j : rem);
	*len = base + (j < rem ? 1 : 0);
}

// nibble-packed group (spl_push_kernel) -> observation bytes: entries 2i / 2i+1 are the low / high nibble of byte i,
// the 17 columns that can exceed 15 follow as whole bytes per env
static const int kWideCols[17] = {12, 13, 14, 15, 16, 17, 25, 26, 27, 28, 29, 30, 290, 291, 292, 293, 295};
static void unpack_group(const uint8_t* src, int m, uint8_t* out) {
	const int nbytes = m * 297;
	const int nn = (nbytes + 1) / 2;
	int i = 0;
#if defined(__x86_64__)
	const __m128i lowmask = _mm_set1_epi8(0x0F);
	for (; i + 16 <= nn; i += 16) {
		const __m128i v = _mm_loadu_si128((const __m128i*)(src + i));
		const __m128i lo = _mm_and_si128(v, lowmask), hi = _mm_and_si128(_mm_srli_epi16(v, 4), lowmask);
		_mm_storeu_si128((__m128i*)(out + 2 * i), _mm_unpacklo_epi8(lo, hi));
		_mm_storeu_si128((__m128i*)(out + 2 * i + 16), _mm_unpackhi_epi8(lo, hi));
	}
#endif
	for (; i < nn; i++) out[2 * i] = src[i] & 15, out[2 * i + 1] = src[i] >> 4;
	const uint8_t* wide = src + SPL_HOST_GROUP * 297 / 2;
	for (int e = 0; e < m; e++)
		for (int w = 0; w < 17; w++) out[e * 297 + kWideCols[w]] = wide[e * 17 + w];
}

static void run_share(SplHostJob* job, int j) {
	int64_t start, len;
	spl_job_share(job, j, &start, &len);
	const int64_t G = SPL_HOST_GROUP;
	double t_first = 0;
	alignas(64) uint8_t tmp[SPL_HOST_GROUP * 297 + 64];
	const bool want_obs = job->io.obs || job->io.obs_u8;
	for (int64_t r = 0; r < len; r++) {
		const int64_t g = start + r;
		const int64_t lo = g * G, hi = lo + G < job->n ? lo + G : job->n;
		if (job->ring == nullptr) {  // complete linear arrays (spl_host_expand)
			spl_expand_block(job->obs_u8 ? job->obs_u8 + lo * 297 : nullptr, job->side + 4 * lo, lo, hi, &job->io);
			continue;
		}
		const uint64_t want = job->tag_base + (uint64_t)r + 1u;
		const uint64_t* flag = job->ring_flags + (size_t)j * job->ring_slots + (size_t)(r % job->ring_slots);
		uint64_t f;
		while (((f = __atomic_load_n(flag, __ATOMIC_ACQUIRE)) & ~(1ull << 63)) != want) {
			if (job->abort.load(std::memory_order_relaxed)) return;
			cpu_relax();
		}
		if (r == 0) t_first = spl_now_us();
		const uint8_t* slot = job->ring + ((size_t)j * job->ring_slots + (size_t)(r % job->ring_slots)) * SPL_SLOT_BYTES;
		const uint32_t* side = reinterpret_cast<const uint32_t*>(slot);
		const uint8_t* obs = slot + SPL_SLOT_SIDE;
		if (want_obs && !(f >> 63)) {
			unpack_group(obs, (int)(hi - lo), tmp);
			obs = tmp;
		}
		spl_expand_block(obs, side, lo, hi, &job->io);
		// every read of the slot is done (x86: loads are not reordered after a later store): the link may overwrite it
		__atomic_store_n(job->consumed + 8 * (size_t)j, want, __ATOMIC_RELEASE);
	}
	job->t_first[j] = t_first;
	job->t_done[j] = spl_now_us();
}

static void* worker_main(void* arg) {
	SplPool* P = g_pool;
	const int j = *(int*)arg;
	if (P->cpus[j] >= 0) {
		cpu_set_t set;
		CPU_ZERO(&set);
		CPU_SET(P->cpus[j], &set);
		pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
	}
	uint64_t seen = 0;
	for (;;) {
		// wait for the next generation: spin ~0.3 ms, then sleep
		unsigned spins = 0;
		while (P->generation.load(std::memory_order_acquire) == seen && !P->stop.load(std::memory_order_relaxed)) {
			cpu_relax();
			if (++spins > 200000u) {
				pthread_mutex_lock(&P->mu);
				P->sleepers.fetch_add(1);
				while (P->generation.load(std::memory_order_acquire) == seen && !P->stop.load()) pthread_cond_wait(&P->cv, &P->mu);
				P->sleepers.fetch_sub(1);
				pthread_mutex_unlock(&P->mu);
				spins = 0;
			}
		}
		if (P->stop.load()) return nullptr;
		seen = P->generation.load(std::memory_order_acquire);
		if (P->job.custom) P->job.custom(P->job.custom_ctx, j);
		else if (j < P->job.threads) run_share(&P->job, j);
		P->done.fetch_add(1, std::memory_order_release);
	}
}

static void pool_shutdown() {
	SplPool* P = g_pool;
	if (!P) return;
	P->stop.store(1);
	pthread_mutex_lock(&P->mu);
	pthread_cond_broadcast(&P->cv);
	pthread_mutex_unlock(&P->mu);
	for (int j = 0; j < P->threads; j++) pthread_join(P->tid[j], nullptr);
	delete P;
	g_pool = nullptr;
}

// the cores this rank may use: the process's affinity mask, cut into LOCAL_WORLD_SIZE contiguous slices (torchrun:
// one process per GPU; contiguous core numbers share a socket, and so does the memory their threads touch first)
static int rank_cpus(int* out, int cap) {
	cpu_set_t set;
	if (sched_getaffinity(0, sizeof(set), &set) != 0) return 0;
	int avail[1024], na = 0;
	for (int c = 0; c < CPU_SETSIZE && na < 1024; c++)
		if (CPU_ISSET(c, &set)) avail[na++] = c;
	if (na == 0) return 0;
	const char* lr = getenv("LOCAL_RANK");
	const char* lws = getenv("LOCAL_WORLD_SIZE");
	const int rank = lr ? atoi(lr) : 0, ws = lws && atoi(lws) > 0 ? atoi(lws) : 1;
	const int per = na / ws > 0 ? na / ws : 1;
	const int first = ((rank % ws) * per) % na;
	int n = 0;
	for (int k = 0; k < per && n < cap; k++) out[n++] = avail[(first + k) % na];
	return n;
}

static void choose_cpus(SplPool* P) {
	for (int j = 0; j < SPL_POOL_MAX; j++) P->cpus[j] = -1;
	if (!g_pin) return;
	int mine[1024];
	const int per = rank_cpus(mine, 1024);
	if (per == 0) return;
	// workers take the rank's cores from the top; the caller's thread is left where the OS put it
	for (int j = 0; j < P->threads; j++) P->cpus[j] = mine[per - 1 - (j % per)];
}

// Run the calling thread on the rank's cores while it allocates and first-touches buffers (several ranks per node only:
// the pages then land on the memory of the socket whose cores will stream into them), then put its mask back.
SplRankAffinity::SplRankAffinity() : active(false) {
	const char* lws = getenv("LOCAL_WORLD_SIZE");
	if (!g_pin || !lws || atoi(lws) <= 1) return;
	if (sched_getaffinity(0, sizeof(saved), &saved) != 0) return;
	int mine[1024];
	const int per = rank_cpus(mine, 1024);
	if (per == 0) return;
	cpu_set_t set;
	CPU_ZERO(&set);
	for (int k = 0; k < per; k++) CPU_SET(mine[k], &set);
	active = sched_setaffinity(0, sizeof(set), &set) == 0;
}
SplRankAffinity::~SplRankAffinity() {
	if (active) sched_setaffinity(0, sizeof(saved), &saved);
}

// While a lock-step is in flight the pinned workers spin on their cores; the caller's thread still has kernels to launch and
// must not be time-sliced against one of them (a runnable thread sharing a core with a spinning worker waits for a scheduler
// tick: milliseconds).  For the duration of the call the caller is kept off the workers' cores; its mask is put back after.
SplCallerOffWorkers::SplCallerOffWorkers() : active(false) {
	SplPool* P = g_pool;
	if (!g_pin || P == nullptr) return;
	if (sched_getaffinity(0, sizeof(saved), &saved) != 0) return;
	// several ranks per node: stay inside this rank's own slice (other ranks' workers spin on theirs); else anywhere but on a worker
	cpu_set_t set;
	const char* lws = getenv("LOCAL_WORLD_SIZE");
	if (lws && atoi(lws) > 1) {
		int mine[1024];
		const int per = rank_cpus(mine, 1024);
		CPU_ZERO(&set);
		for (int k = 0; k < per; k++)
			if (CPU_ISSET(mine[k], &saved)) CPU_SET(mine[k], &set);
	} else {
		set = saved;
	}
	int removed = 0;
	for (int j = 0; j < P->threads; j++)
		if (P->cpus[j] >= 0 && CPU_ISSET(P->cpus[j], &set)) CPU_CLR(P->cpus[j], &set), removed++;
	if (CPU_COUNT(&set) == 0) {  // the workers own the whole slice: anywhere but on a worker of this rank
		set = saved;
		for (int j = 0; j < P->threads; j++)
			if (P->cpus[j] >= 0) CPU_CLR(P->cpus[j], &set);
	}
	if (removed == 0 || CPU_COUNT(&set) == 0) return;  // nothing pinned, or the workers own every core this thread may use
	active = sched_setaffinity(0, sizeof(set), &set) == 0;
}
SplCallerOffWorkers::~SplCallerOffWorkers() {
	if (active) sched_setaffinity(0, sizeof(saved), &saved);
}

// default size of the pool: the cores this rank may use minus two (one for the caller's thread, which launches the kernels and
// then only waits, one for everything else in the process), at least 3/4 of them.  Measured on a 16-vCPU box with the final
// path (tools/sweep_host.py, same box): 10 workers 811 us per lock-step, 12: 753-823, 13: 742, 14: 704, 15: 730.
static void default_threads() {
	cpu_set_t set;
	int n = 0;
	if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
	if (n < 1) n = 1;
	const char* lws = getenv("LOCAL_WORLD_SIZE");  // one process per GPU (torchrun): the ranks of a node share its cores
	if (lws && atoi(lws) > 1) n /= atoi(lws);
	n = n - 2 > (n * 3 + 3) / 4 ? n - 2 : (n * 3 + 3) / 4;
	const char* e = getenv("SPL_HOST_THREADS");
	if (e && atoi(e) > 0) n = atoi(e);
	if (n < 1) n = 1;
	if (n > 32) n = 32;
	g_threads = n;
	const char* pin = getenv("SPL_HOST_PIN");
	if (pin && atoi(pin) == 0) g_pin = 0;
}

static SplPool* pool_get() {
	if (g_threads <= 0) default_threads();
	const int want = g_threads > 0 ? g_threads : 1;
	if (g_pool && g_pool->threads == want) return g_pool;
	pool_shutdown();
	SplPool* P = new SplPool();
	P->threads = want > SPL_POOL_MAX ? SPL_POOL_MAX : want;
	choose_cpus(P);
	g_pool = P;
	for (int j = 0; j < P->threads; j++) {
		P->index[j] = j;
		pthread_create(&P->tid[j], nullptr, worker_main, &P->index[j]);
	}
	static bool registered = false;
	if (!registered) {
		registered = true;
		atexit(pool_shutdown);
	}
	return P;
}

SplHostJob* spl_pool_job() { return &pool_get()->job; }
int spl_pool_threads() { return pool_get()->threads; }

// Run the job that was filled into spl_pool_job() and return when every share is done.  The caller's thread does none of the
// widening: it may be held up inside the kernel launch (a profiler serialising launches, a slow driver call) while the push
// kernel already waits for ring slots to be emptied -- only threads that are free to run may own a share.  It waits for the
// GPU-written share (after_wait) and watches the stream for errors.
static void pool_dispatch(SplPool* P) {
	P->job.abort.store(0);
	P->done.store(0, std::memory_order_relaxed);
	P->generation.fetch_add(1, std::memory_order_release);
	if (P->sleepers.load() > 0) {
		pthread_mutex_lock(&P->mu);
		pthread_cond_broadcast(&P->cv);
		pthread_mutex_unlock(&P->mu);
	}
}

// Two halves: spl_pool_start() hands the job to the workers (they begin to poll their arrival tags), spl_pool_wait() returns
// when every share is done.  spl_host_step starts the workers BEFORE it launches the kernels: a launch that blocks until the
// kernel has finished (CUDA_LAUNCH_BLOCKING, a profiler serialising launches) would otherwise wait for ring slots that
// nobody has been asked to empty yet.
void spl_pool_start() {
	SplPool* P = pool_get();
	SplHostJob* job = &P->job;
	if (job->threads > P->threads) job->threads = P->threads;
	if (job->threads < 1) job->threads = 1;
	pool_dispatch(P);
}

void spl_pool_wait() {
	SplPool* P = pool_get();
	SplHostJob* job = &P->job;
	if (job->after_share0) job->after_share0(job);
	unsigned spins = 0;
	while (P->done.load(std::memory_order_acquire) < P->threads) {
		if (job->poll && (++spins & 0x3FFFu) == 0 && !job->abort.load(std::memory_order_relaxed) && job->poll(job->poll_ctx)) job->abort.store(1);
		cpu_relax();
	}
}

void spl_pool_abort() {  // a launch failed after spl_pool_start: release the workers
	SplPool* P = pool_get();
	P->job.abort.store(1);
	while (P->done.load(std::memory_order_acquire) < P->threads) cpu_relax();
}

void spl_pool_run() {
	spl_pool_start();
	spl_pool_wait();
}

// streaming-store rate of the pool (GB/s written) over `bytes` per thread, `reps` passes: the ceiling the widened
// results are reported against (bench.py e2e.host_store_gbs).  mode 0: fill, 1: widen u8 -> int32 from a 1/4-size source
extern "C" double spl_host_store_rate(int64_t bytes_per_thread, int reps, int mode) {
	SplPool* P = pool_get();
	const int T = P->threads;
	if (bytes_per_thread < 4096) bytes_per_thread = 4096;
	bytes_per_thread &= ~(int64_t)4095;
	// destination like the path's result arrays (spl_host_alloc): 2 MB-aligned anonymous memory, huge pages requested
	const size_t huge = (size_t)2 << 20;
	const size_t dst_len = ((size_t)bytes_per_thread * T + huge - 1) / huge * huge + huge;
	void* raw = mmap(nullptr, dst_len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	if (raw == MAP_FAILED) return 0.0;
	char* dst = (char*)(((uintptr_t)raw + huge - 1) / huge * huge);
	madvise(dst, dst_len - huge, MADV_HUGEPAGE);
	uint8_t* src = nullptr;
	if (posix_memalign((void**)&src, 4096, (size_t)bytes_per_thread / 4 * T)) {
		munmap(raw, dst_len);
		return 0.0;
	}
	memset(dst, 0, (size_t)bytes_per_thread * T);
	memset(src, 3, (size_t)bytes_per_thread / 4 * T);
	struct Ctx {
		char* dst;
		uint8_t* src;
		int64_t bytes;
		int mode;
	} ctx{dst, src, bytes_per_thread, mode};
	// a job whose "groups" are whole per-thread buffers: reuse the pool through a custom block function
	// sustained rate: one untimed pass, then at least `reps` passes and at least 40 ms (ranks of a node that measure at the
	// same time then overlap for the whole measurement, whatever their relative start)
	double total_us = 0.0, best = 0.0;
	int passes = 0;
	for (int r = 0; r < 1000 && (passes < reps || total_us < 40000.0); r++) {
		SplHostJob* job = &P->job;
		*job = SplHostJob();
		job->threads = T;
		job->custom = [](void* c, int j) {
			Ctx* x = (Ctx*)c;
#if defined(__x86_64__)
			if (x->mode == 0 && simd_level() >= 5) {
				fill_avx512(x->dst + (size_t)j * x->bytes, (size_t)x->bytes);
				return;
			}
#endif
			if (x->mode == 0) memset(x->dst + (size_t)j * x->bytes, 1, (size_t)x->bytes);
			else widen(x->src + (size_t)j * (x->bytes / 4), (int32_t*)(x->dst + (size_t)j * x->bytes), (size_t)x->bytes / 4);
		};
		job->custom_ctx = &ctx;
		const double t0 = spl_now_us();
		spl_pool_run_custom();
		const double dt = spl_now_us() - t0;
		if (r == 0) continue;
		total_us += dt;
		passes++;
	}
	best = passes > 0 ? (double)bytes_per_thread * T * passes / total_us * 1e-3 : 0.0;
	munmap(raw, dst_len);
	free(src);
	return best;
}

void spl_pool_run_custom() {
	SplPool* P = pool_get();
	P->job.cpu_groups = 0;
	pool_dispatch(P);
	while (P->done.load(std::memory_order_acquire) < P->threads) cpu_relax();
}

void spl_parallel_copy(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); }

extern "C" int spl_host_expand(const uint8_t* obs_u8, const void* side, int64_t n, const spl_host_io_t* io) {
	if (!side || !io || n <= 0 || ((io->obs || io->obs_u8) && !obs_u8)) return SPL_E_BADARG;
	SplHostJob* job = spl_pool_job();
	*job = SplHostJob();
	job->obs_u8 = obs_u8, job->side = (const uint32_t*)side, job->io = *io, job->n = n;  // (no ring: the arrays are complete)
	job->threads = spl_pool_threads();
	job->cpu_groups = (n + SPL_HOST_GROUP - 1) / SPL_HOST_GROUP;
	spl_pool_run();
	return 0;
}

extern "C" int spl_host_set_threads(int n) {
	if (n > 0) g_threads = n > SPL_POOL_MAX ? SPL_POOL_MAX : n;
	else if (g_threads <= 0) default_threads();
	return g_threads;
}

extern "C" int spl_host_set_pinning(int on) {
	g_pin = on ? 1 : 0;
	return g_pin;
}
