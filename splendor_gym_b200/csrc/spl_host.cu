// spl_host.cu -- host-buffer entry points of the C ABI (include/splendor_b200.h: spl_host_*).
//
// The reference's callers hold NumPy arrays on the host (SplendorEnv.step, envs/splendor_env.py:51-90; the vector
// loop of ppo_splendor.py:235-285).  For them a lock-step is: actions host->device, the step kernel, results
// device->host.  The reference-typed results are 1,243 B per env-step; PCIe alone would cap that at ~4.5e7
// env-steps/s, and a host core writes them at ~20 GB/s.  Two resources can deliver them -- the PCIe link writing
// host memory directly, and host cores widening a compact form -- so the path uses both at once:
//
//   stream:  H2D actions | step kernel (COMPACT: 297 observation bytes + one 16-byte record per env, in HBM)
//            | push kernel: a few warps walk the envs in groups of 64 and store, over PCIe, into pinned host memory
//                 * for the CPU share: the compact group into the staging buffers, then its arrival flag
//                 * for the direct share: the group ALREADY WIDENED (int32 observation, int8 mask, float reward, ...)
//                   straight into the caller's arrays (when those are pinned / registered, spl_host_alloc)
//   host  :  a pool of pinned worker threads, each owning a contiguous range of the CPU share, polls the arrival
//            flags and widens group after group with non-temporal stores (spl_host_expand.cpp)
//
// The push order serves the workers round-robin (group r of every worker before group r+1 of any), so every worker
// starts after ~one group's transfer time and the link is never idle.  The split between the two shares is a
// feedback loop on the two finish times (the link's last flag vs the slowest worker), see balance().
//
// There are no copy-engine calls and no events on the results path: one H2D copy and two kernel launches per
// lock-step.  The library owns the compact device buffers, the pinned staging and the flags of one `spl_host_t`;
// the game state stays in the caller's (PyTorch's) tensors as everywhere else.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include "../../include/splendor_b200.h"
#include "spl_host_pool.h"

#define SPL_OBS_DIM_ 297
#define SPL_GROUP_OBS_BYTES (SPL_HOST_GROUP * SPL_OBS_DIM_) /* 19,008 = 1,188 x 16 */

extern int64_t g_launches;
int spl_launch_compact(const spl_envs_t* e, const spl_step_io_t* io, bool do_step, uint8_t* obs_u8, void* side, cudaStream_t st);

struct PushParams {
	const uint8_t* d_obs;  // [n][297] HBM (+16 B pad)
	const uint4* d_side;   // [n] HBM
	int64_t n;
	int32_t groups, cpu_groups, threads;
	uint32_t seq;     // 31 bits
	int32_t nibbles;  // nibble-pack the observation bytes of the CPU share
	uint32_t* flags;  // [groups] pinned host: seq | bit 31 = this group's observation bytes are raw
	// CPU share: per-worker staging rings (pinned host, device aliases), see spl_host_pool.h
	uint8_t* ring;                        // [threads][SPL_RING_SLOTS][SPL_SLOT_BYTES]
	unsigned long long* ring_flags;       // [threads][SPL_RING_SLOTS]
	const unsigned long long* consumed;   // [threads][8], written by the workers
	unsigned long long tag_base;
	int32_t ring_slots;  // slots per worker
	int32_t want_obs;
	// direct share: the caller's arrays (device aliases of pinned host memory), each nullable
	int32_t* o_obs;
	uint8_t* o_obs_u8;
	int8_t* o_mask;
	float* o_reward;
	uint8_t* o_term;
	uint8_t* o_info;
	int32_t* o_next;
};

__device__ __forceinline__ float spl_reward_of_code(uint32_t c) {
	return c == SPL_REWARD_CODE_WIN ? 1.0f : c == SPL_REWARD_CODE_LOSS ? -1.0f : c == SPL_REWARD_CODE_LIMIT ? -0.1f
	     : c == SPL_REWARD_CODE_ILLEGAL ? -0.01f : 0.0f;
}

// Columns of the observation that can exceed 15 in legal play (bonuses, prestige, deck sizes, turn / move counters,
// engine/encode.py:113-183): the nibble-packed form carries them as whole bytes behind the nibbles.
#define SPL_WIDE_COLS 17
__constant__ int c_wide_cols[SPL_WIDE_COLS] = {12, 13, 14, 15, 16, 17, 25, 26, 27, 28, 29, 30, 290, 291, 292, 293, 295};
#define SPL_NIBBLE_BYTES (SPL_GROUP_OBS_BYTES / 2) /* 9,504 = 594 x 16 */

// One CTA per group of 64 envs, groups in push order.  Stores to pinned host memory are posted PCIe writes;
// __threadfence_system orders a group's payload before its flag.
//   CPU share   : side records as they are; observation bytes nibble-packed (two entries per byte + the 17 wide columns
//                 as bytes = 165.5 B per env instead of 297), or raw when a record says an entry does not fit (flag bit 31)
//   direct share: the reference-typed arrays themselves
__global__ void __launch_bounds__(256) spl_push_kernel(const PushParams p) {
	__shared__ int s_raw;
	__shared__ __align__(16) uint8_t s_wide[SPL_HOST_GROUP * SPL_WIDE_COLS];
	const int tid = threadIdx.x, nthr = blockDim.x;
	const int T = p.threads;
	const int base = p.cpu_groups / T, rem = p.cpu_groups % T;
	for (int k = blockIdx.x; k < p.groups; k += gridDim.x) {
		int g = k, rr = 0, jj = 0;
		const bool direct = k >= p.cpu_groups;
		if (!direct) {  // push order k -> (round rr, worker jj) -> group start_jj + rr   (spl_job_share)
			if (k < base * T) rr = k / T, jj = k - rr * T;
			else rr = base, jj = k - base * T;
			g = jj * base + min(jj, rem) + rr;
		}
		const int64_t lo = (int64_t)g * SPL_HOST_GROUP;
		const int m = (int)min((int64_t)SPL_HOST_GROUP, p.n - lo);
		const uint8_t* sb = p.d_obs + lo * SPL_OBS_DIM_;
		const uint4* so = reinterpret_cast<const uint4*>(sb);
		unsigned long long tag = 0;
		unsigned long long* flag64 = nullptr;
		if (!direct) {
			const int slot_i = jj * p.ring_slots + rr % p.ring_slots;
			uint8_t* slot = p.ring + (size_t)slot_i * SPL_SLOT_BYTES;
			tag = p.tag_base + (unsigned long long)rr + 1ull;
			flag64 = p.ring_flags + slot_i;
			if (tid == 0) {
				s_raw = p.nibbles ? 0 : 1;
				// flow control: the slot's previous occupant (round rr - SLOTS of this worker) must have been emptied.  The first
				// SLOTS rounds of a lock-step never wait (the previous call returned only after every worker was done).
				if (rr >= p.ring_slots) {
					const volatile unsigned long long* c = p.consumed + 8 * jj;
					while (*c + (unsigned long long)p.ring_slots < tag) __nanosleep(500);
				}
			}
			__syncthreads();
			for (int e = tid; e < m; e += nthr) {
				const uint4 rec = __ldcs(p.d_side + lo + e);
				if (rec.w & 1u) s_raw = 1;
				reinterpret_cast<uint4*>(slot)[e] = rec;
			}
			if (p.want_obs) {
				for (int q = tid; q < m * SPL_WIDE_COLS; q += nthr) {
					const int e = q / SPL_WIDE_COLS;
					s_wide[q] = sb[e * SPL_OBS_DIM_ + c_wide_cols[q - e * SPL_WIDE_COLS]];
				}
			}
			__syncthreads();
			const bool raw = s_raw != 0;
			if (raw) tag |= 1ull << 63;
			if (p.want_obs) {
				uint8_t* db = slot + SPL_SLOT_SIDE;
				if (raw) {
					const int nq = (m * SPL_OBS_DIM_) >> 4;
#pragma unroll 4
					for (int q = tid; q < nq; q += nthr) reinterpret_cast<uint4*>(db)[q] = __ldcs(so + q);
					for (int b = 16 * nq + tid; b < m * SPL_OBS_DIM_; b += nthr) db[b] = sb[b];
				} else {
					// entries 2i, 2i+1 -> byte i (low, high nibble).  32 source bytes -> one 16-byte store; a partial last
					// group reads a few bytes past its rows (the buffer is padded), the host ignores them
					const int nq = (m * SPL_OBS_DIM_ + 31) >> 5;
#pragma unroll 2
					for (int q = tid; q < nq; q += nthr) {
						const uint4 a = __ldcs(so + 2 * q), b = __ldcs(so + 2 * q + 1);
						auto pk = [](uint32_t x, uint32_t y) {
							x &= 0x0F0F0F0Fu, y &= 0x0F0F0F0Fu;
							x |= x >> 4, y |= y >> 4;
							return __byte_perm(x, y, 0x6420);
						};
						reinterpret_cast<uint4*>(db)[q] = make_uint4(pk(a.x, a.y), pk(a.z, a.w), pk(b.x, b.y), pk(b.z, b.w));
					}
					uint4* dw = reinterpret_cast<uint4*>(db + SPL_NIBBLE_BYTES);
					for (int q = tid; q < (m * SPL_WIDE_COLS + 15) >> 4; q += nthr) dw[q] = reinterpret_cast<const uint4*>(s_wide)[q];
				}
			}
		} else {
			if (p.o_obs_u8 != nullptr) {
				if (m == SPL_HOST_GROUP) {
					uint4* dob = reinterpret_cast<uint4*>(p.o_obs_u8 + lo * SPL_OBS_DIM_);
#pragma unroll 4
					for (int q = tid; q < SPL_GROUP_OBS_BYTES / 16; q += nthr) dob[q] = __ldcs(so + q);
				} else {
					for (int q = tid; q < m * SPL_OBS_DIM_; q += nthr) p.o_obs_u8[lo * SPL_OBS_DIM_ + q] = sb[q];
				}
			}
			if (p.o_obs != nullptr) {
				const uint32_t* sw = reinterpret_cast<const uint32_t*>(so);
				int4* dob = reinterpret_cast<int4*>(p.o_obs + lo * SPL_OBS_DIM_);
				const int nw = (m * SPL_OBS_DIM_) >> 2;  // whole 4-byte words of the group
#pragma unroll 4
				for (int q = tid; q < nw; q += nthr) {
					const uint32_t v = __ldcs(sw + q);
					dob[q] = make_int4((int)(v & 0xFFu), (int)((v >> 8) & 0xFFu), (int)((v >> 16) & 0xFFu), (int)(v >> 24));
				}
				for (int q = 4 * nw + tid; q < m * SPL_OBS_DIM_; q += nthr) p.o_obs[lo * SPL_OBS_DIM_ + q] = (int32_t)sb[q];
			}
			if (p.o_mask != nullptr) {
				int8_t* dm = p.o_mask + lo * 45;
				const int nb = m * 45;
				for (int q = tid; q < (nb >> 4); q += nthr) {
					uint32_t w[4];
#pragma unroll
					for (int c = 0; c < 4; c++) {
						uint32_t acc = 0;
#pragma unroll
						for (int b = 0; b < 4; b++) {
							const int byte = 16 * q + 4 * c + b;
							const int e = byte / 45, a = byte - e * 45;
							const uint4 rec = __ldg(p.d_side + lo + e);
							const uint64_t mk = (uint64_t)rec.x | ((uint64_t)(rec.y & 0x1FFFu) << 32);
							acc |= (uint32_t)((mk >> a) & 1ull) << (8 * b);
						}
						w[c] = acc;
					}
					reinterpret_cast<uint4*>(dm)[q] = make_uint4(w[0], w[1], w[2], w[3]);
				}
				for (int byte = (nb & ~15) + tid; byte < nb; byte += nthr) {
					const int e = byte / 45, a = byte - e * 45;
					const uint4 rec = __ldg(p.d_side + lo + e);
					const uint64_t mk = (uint64_t)rec.x | ((uint64_t)(rec.y & 0x1FFFu) << 32);
					dm[byte] = (int8_t)((mk >> a) & 1ull);
				}
			}
			for (int e = tid; e < m; e += nthr) {
				const uint4 rec = __ldg(p.d_side + lo + e);
				if (p.o_reward != nullptr) p.o_reward[lo + e] = spl_reward_of_code((rec.y >> 16) & 7u);
				if (p.o_term != nullptr) p.o_term[lo + e] = (uint8_t)((rec.y >> 24) & 1u);
				if (p.o_info != nullptr) p.o_info[lo + e] = (uint8_t)(rec.z & 0xFFu);
				if (p.o_next != nullptr) p.o_next[lo + e] = (int32_t)((rec.z >> 8) & 0xFFu);
			}
		}
		__threadfence_system();
		__syncthreads();
		if (tid == 0) {
			if (direct) *reinterpret_cast<volatile uint32_t*>(p.flags + g) = p.seq;
			else *reinterpret_cast<volatile unsigned long long*>(flag64) = tag;
		}
	}
}

struct spl_host {
	int64_t n;
	int32_t groups;
	uint8_t* d_obs;   // HBM: compact outputs of the step kernel
	uint4* d_side;
	// per-worker staging rings in pinned host memory + device aliases (spl_host_pool.h)
	int ring_threads;  // workers the rings were sized for
	int ring_slots;    // slots per worker (SPL_RING_SLOTS; the environment variable of that name overrides it, diagnostics)
	uint8_t *h_ring, *hd_ring;
	uint64_t *h_ring_flags, *h_consumed;
	unsigned long long *hd_ring_flags, *hd_consumed;
	uint64_t tag_base;
	int32_t *h_act, *hd_act;  // pinned copy of the actions | its device alias
	uint32_t *h_flags, *hd_flags;
	uint32_t seq;
	int device;
	int push_ctas, push_threads;
	int nibbles;
	// split between the link and the cores
	double direct_frac;      // share of the groups the GPU writes already widened
	int direct_fixed;        // SPL_HOST_DIRECT set: no feedback
	// device aliases of the caller's arrays, looked up on every call
	spl_host_io_t seen_io;
	PushParams seen_dev;
	int seen_ok;
	cudaStream_t poll_stream;
	double t_gpu;  // when the last flag of the GPU-written share was seen (this call)
	double stats[SPL_HOST_STATS];
};

#define SPL_CUDA(x)                                \
	do {                                           \
		cudaError_t e_ = (x);                      \
		if (e_ != cudaSuccess) return (int)e_;     \
	} while (0)

static int env_int(const char* name, int dflt) {
	const char* e = getenv(name);
	return e && *e ? atoi(e) : dflt;
}

// device alias of a host pointer the GPU can write (pinned / registered), or null
static void* dev_alias(const void* host) {
	if (!host) return nullptr;
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return nullptr;
	return at.devicePointer;
}

extern "C" {

int spl_host_destroy(spl_host_t* h);

int spl_host_create(int64_t n, int32_t chunks, spl_host_t** out) {
	(void)chunks;  // kept in the signature (ABI): arrival is tracked per 64-env group now
	if (n <= 0 || !out) return SPL_E_BADARG;
	SplRankAffinity on_my_cores;
	spl_host* h = (spl_host*)calloc(1, sizeof(spl_host));
	if (!h) return SPL_E_BADARG;
	h->n = n;
	h->groups = (int32_t)((n + SPL_HOST_GROUP - 1) / SPL_HOST_GROUP);
	h->push_ctas = env_int("SPL_PUSH_CTAS", 64);  // measured: 64 groups in flight hide the fence + flag latency of a group (~8 us)
	h->nibbles = env_int("SPL_HOST_NIBBLES", 1);
	h->push_threads = env_int("SPL_PUSH_THREADS", 64) / 32 * 32;
	if (h->push_threads < 32) h->push_threads = 32;
	if (h->push_threads > 256) h->push_threads = 256;
	const char* df = getenv("SPL_HOST_DIRECT");
	h->direct_fixed = df && *df;
	h->direct_frac = h->direct_fixed ? atof(df) : 0.10;
	if (h->direct_frac < 0) h->direct_frac = 0;
	if (h->direct_frac > 1) h->direct_frac = 1;
	const unsigned fl = cudaHostAllocMapped | cudaHostAllocPortable;
	cudaError_t e = cudaGetDevice(&h->device);
	if (e == cudaSuccess) e = cudaMalloc(&h->d_obs, (size_t)n * SPL_OBS_DIM_ + 64);
	if (e == cudaSuccess) e = cudaMalloc(&h->d_side, (size_t)n * 16);
	h->ring_threads = spl_pool_threads();
	h->ring_slots = env_int("SPL_RING_SLOTS", SPL_RING_SLOTS);
	if (h->ring_slots < 1) h->ring_slots = 1;
	const size_t slots = (size_t)h->ring_threads * h->ring_slots;
	if (e == cudaSuccess) e = cudaHostAlloc(&h->h_ring, slots * SPL_SLOT_BYTES, fl);
	if (e == cudaSuccess) e = cudaHostAlloc(&h->h_ring_flags, slots * 8, fl);
	if (e == cudaSuccess) e = cudaHostAlloc(&h->h_consumed, (size_t)h->ring_threads * 64, fl);
	if (e == cudaSuccess) e = cudaHostAlloc(&h->h_act, (size_t)n * 4, fl);
	if (e == cudaSuccess) e = cudaHostAlloc(&h->h_flags, (size_t)h->groups * 4 + 64, fl);
	if (e == cudaSuccess) e = cudaHostGetDevicePointer(&h->hd_ring, h->h_ring, 0);
	if (e == cudaSuccess) e = cudaHostGetDevicePointer(&h->hd_ring_flags, h->h_ring_flags, 0);
	if (e == cudaSuccess) e = cudaHostGetDevicePointer(&h->hd_consumed, h->h_consumed, 0);
	if (e == cudaSuccess) e = cudaHostGetDevicePointer(&h->hd_flags, h->h_flags, 0);
	if (e == cudaSuccess) e = cudaHostGetDevicePointer(&h->hd_act, h->h_act, 0);
	if (e != cudaSuccess) {  // nothing half-built is handed out (spl_host_destroy skips what was never created)
		spl_host_destroy(h);
		return (int)e;
	}
	memset(h->h_flags, 0, (size_t)h->groups * 4 + 64);
	memset(h->h_ring, 0, slots * SPL_SLOT_BYTES);
	memset(h->h_ring_flags, 0, slots * 8);
	memset(h->h_consumed, 0, (size_t)h->ring_threads * 64);
	h->tag_base = 0;
	*out = h;
	return 0;
}

int spl_host_destroy(spl_host_t* h) {
	if (!h) return 0;
	if (h->d_obs) cudaFree(h->d_obs);
	if (h->d_side) cudaFree(h->d_side);
	if (h->h_ring) cudaFreeHost(h->h_ring);
	if (h->h_ring_flags) cudaFreeHost(h->h_ring_flags);
	if (h->h_consumed) cudaFreeHost(h->h_consumed);
	if (h->h_act) cudaFreeHost(h->h_act);
	if (h->h_flags) cudaFreeHost(h->h_flags);
	free(h);
	return 0;
}

/* result arrays the GPU can write directly: anonymous memory with transparent huge pages requested, registered with
 * CUDA (pinned + mapped).  Plain host memory to the caller (NumPy / torch can wrap it). */
int spl_host_alloc(size_t bytes, void** out) {
	if (!out || bytes == 0) return SPL_E_BADARG;
	SplRankAffinity on_my_cores;
	const size_t huge = (size_t)2 << 20;
	const size_t len = (bytes + huge - 1) / huge * huge;
	void* p = mmap(nullptr, len + huge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	if (p == MAP_FAILED) return SPL_E_BADARG;
	// 2 MB-aligned start inside the mapping (the slack stays mapped; spl_host_free recovers the base from the header)
	uintptr_t a = ((uintptr_t)p + huge - 1) / huge * huge;
	if (a == (uintptr_t)p) a += huge;  // room for the header below the aligned block
	madvise((void*)a, len, MADV_HUGEPAGE);
	size_t* hdr = (size_t*)(a - 3 * sizeof(size_t));
	hdr[0] = (size_t)((uintptr_t)p), hdr[1] = len + huge;
	memset((void*)a, 0, len);  // fault the pages in (as huge pages where the kernel grants them) before pinning
	// A box that refuses to pin this much memory still gets its arrays: spl_host_step then sees that they are not
	// GPU-writable and lets host threads widen everything.
	cudaError_t e = cudaHostRegister((void*)a, len, cudaHostRegisterMapped | cudaHostRegisterPortable);
	if (e != cudaSuccess) cudaGetLastError();
	hdr[2] = e == cudaSuccess ? 1 : 0;
	*out = (void*)a;
	return 0;
}

int spl_host_free(void* ptr) {
	if (!ptr) return 0;
	size_t* hdr = (size_t*)((uintptr_t)ptr - 3 * sizeof(size_t));
	if (hdr[2] && cudaHostUnregister(ptr) != cudaSuccess) cudaGetLastError();
	munmap((void*)(uintptr_t)hdr[0], hdr[1]);
	return 0;
}

int spl_host_get_stats(const spl_host_t* h, double* out) {
	if (!h || !out) return SPL_E_BADARG;
	memcpy(out, h->stats, sizeof(h->stats));
	return 0;
}

}  // extern "C"

// caller's arrays -> device aliases for the direct share; all requested outputs must be GPU-writable and 16-byte aligned
static bool resolve_direct(spl_host* h, const spl_host_io_t* io) {
	// looked up on every call (~0.5 us per pointer): an address seen before may have been freed and reused since
	h->seen_io = *io;
	PushParams& d = h->seen_dev;
	memset(&d, 0, sizeof(d));
	d.o_obs = (int32_t*)dev_alias(io->obs), d.o_obs_u8 = (uint8_t*)dev_alias(io->obs_u8), d.o_mask = (int8_t*)dev_alias(io->mask);
	d.o_reward = (float*)dev_alias(io->reward), d.o_term = (uint8_t*)dev_alias(io->terminated);
	d.o_info = (uint8_t*)dev_alias(io->info), d.o_next = (int32_t*)dev_alias(io->next_action);
	bool ok = (!io->obs || d.o_obs) && (!io->obs_u8 || d.o_obs_u8) && (!io->mask || d.o_mask) && (!io->reward || d.o_reward) &&
	          (!io->terminated || d.o_term) && (!io->info || d.o_info) && (!io->next_action || d.o_next);
	const uintptr_t al = (uintptr_t)d.o_obs | (uintptr_t)d.o_obs_u8 | (uintptr_t)d.o_mask | (uintptr_t)d.o_reward | (uintptr_t)d.o_term |
	                     (uintptr_t)d.o_info | (uintptr_t)d.o_next;
	ok = ok && (al & 15) == 0;
	h->seen_ok = ok ? 1 : -1;
	return ok;
}

// Feedback on the split: the GPU-written share should land when the slowest worker finishes.  `gap` = link finish
// minus cores finish (us) over a step of `total` us: move the share by a fraction of the relative gap.
static void balance(spl_host* h, double gap, double total, bool direct_possible) {
	if (h->direct_fixed || !direct_possible || total <= 0) return;
	double f = h->direct_frac - 0.25 * gap / total;
	if (f < 0.02) f = 0.02;  // never zero: the link's finish time is only observable through a GPU-written share
	if (f > 0.6) f = 0.6;
	h->direct_frac = f;
}

static int host_run(spl_host_t* h, const spl_envs_t* envs, const spl_host_io_t* io, bool do_step, cudaStream_t st) {
	if (!h || !envs || !io || envs->n != h->n) return SPL_E_BADARG;
	if (do_step && !io->actions) return SPL_E_BADARG;
	const int64_t n = h->n;
	const double t0 = spl_now_us();
	spl_pool_threads();  // (creates the pool on first use)
	SplCallerOffWorkers off_the_worker_cores;
	spl_step_io_t dio;
	memset(&dio, 0, sizeof(dio));
	if (do_step) {
		// The step kernel reads its actions straight from host memory (one coalesced 128-byte PCIe read per warp, ~2 us of
		// latency once per launch -- a copy-engine hop in front of the kernel costs ~10 us): the caller's own array when
		// the GPU can see it, else a pinned copy
		const int32_t* dev_act = ((uintptr_t)io->actions & 3) ? nullptr : (const int32_t*)dev_alias(io->actions);  // (~1 us; not cached: the array may be freed)
		if (!dev_act) {
			memcpy(h->h_act, io->actions, (size_t)n * 4);
			dev_act = h->hd_act;
		}
		dio.actions = dev_act;
	}
	dio.stats = io->stats;
	dio.action_key = io->action_key, dio.action_t = io->action_t;
	dio.autoreset = io->autoreset;
	const bool want_obs = io->obs || io->obs_u8;
	const bool direct_ok = resolve_direct(h, io);
	PushParams p = h->seen_dev;
	p.d_obs = h->d_obs, p.d_side = h->d_side, p.n = n, p.groups = h->groups;
	h->seq = (h->seq + 1u) & 0x7FFFFFFFu;
	if (h->seq == 0) h->seq = 1;
	p.flags = h->hd_flags, p.seq = h->seq;
	p.nibbles = h->nibbles;
	p.ring = h->hd_ring, p.ring_flags = h->hd_ring_flags, p.consumed = h->hd_consumed, p.tag_base = h->tag_base;
	p.ring_slots = h->ring_slots;
	p.want_obs = want_obs ? 1 : 0;
	int T = spl_pool_threads();
	if (T > h->ring_threads) T = h->ring_threads;  // (the pool grew after this context was created)
	int64_t direct_groups = direct_ok ? (int64_t)(h->direct_frac * h->groups + 0.5) : 0;
	if (direct_groups > h->groups) direct_groups = h->groups;
	p.cpu_groups = (int32_t)(h->groups - direct_groups);
	if (T > p.cpu_groups) T = p.cpu_groups > 0 ? p.cpu_groups : 1;
	p.threads = T;
	for (int j = 0; j < T; j++) h->h_consumed[8 * j] = h->tag_base;  // every worker starts the lock-step with an empty ring
	SplHostJob* job = spl_pool_job();
	*job = SplHostJob();
	job->io = *io, job->n = n;
	job->ring = h->h_ring, job->ring_flags = h->h_ring_flags, job->consumed = h->h_consumed, job->tag_base = h->tag_base;
	job->ring_slots = h->ring_slots;
	h->tag_base += (uint64_t)(p.cpu_groups / T + 1);  // tags never repeat: a stale flag cannot match a later lock-step
	job->cpu_groups = p.cpu_groups, job->threads = T;
	job->flags = h->h_flags, job->seq = p.seq;
	h->poll_stream = st;
	job->poll = [](void* ctx) -> int {
		cudaError_t q = cudaStreamQuery(((spl_host*)ctx)->poll_stream);
		return q != cudaSuccess && q != cudaErrorNotReady;
	};
	job->poll_ctx = h;
	h->t_gpu = 0.0;
	job->after_share0 = [](SplHostJob* jb) {  // the GPU-written share: its flags follow the CPU share's in push order
		spl_host* hh = (spl_host*)jb->poll_ctx;
		unsigned spins = 0;
		for (int64_t g = hh->groups - 1; g >= jb->cpu_groups; g--) {
			while (__atomic_load_n(jb->flags + g, __ATOMIC_ACQUIRE) != jb->seq) {
				if (jb->abort.load(std::memory_order_relaxed)) return;
				if ((++spins & 0x3FFFu) == 0 && jb->poll(jb->poll_ctx)) {
					jb->abort.store(1);
					return;
				}
			}
		}
		hh->t_gpu = spl_now_us();
	};
	// The workers are started BEFORE the kernels are launched: a launch that only returns when its kernel has finished
	// (CUDA_LAUNCH_BLOCKING, a profiler serialising launches) must find somebody emptying the ring slots the push kernel
	// waits for.  Until the first tag arrives they only poll.
	spl_pool_start();
	int rc = spl_launch_compact(envs, &dio, do_step, h->d_obs, h->d_side, st);
	if (rc == 0) {
		spl_push_kernel<<<h->push_ctas, h->push_threads, 0, st>>>(p);
		g_launches++;
		rc = (int)cudaGetLastError();
	}
	if (rc) {
		spl_pool_abort();
		return rc;
	}
	const double t1 = spl_now_us();
	spl_pool_wait();
	const double t3 = spl_now_us();
	if (job->abort.load()) {
		cudaError_t q = cudaStreamSynchronize(st);
		return q != cudaSuccess ? (int)q : SPL_E_BADARG;
	}
	double t_cpu = t1, t_first = 0;
	for (int j = 0; j < T; j++) {
		if (job->t_done[j] > t_cpu) t_cpu = job->t_done[j];
		if (job->t_first[j] > t_first) t_first = job->t_first[j];
	}
	const double t_gpu = direct_groups > 0 ? h->t_gpu : t_cpu;
	h->stats[0] = t3 - t0;                        // whole call
	h->stats[1] = t1 - t0;                        // actions staged, copy + two kernels enqueued
	h->stats[2] = t_first > 0 ? t_first - t0 : 0; // last worker to see its first group
	h->stats[3] = t_cpu - t0;                     // slowest worker done
	h->stats[4] = t_gpu - t0;                     // GPU-written share landed
	h->stats[5] = (double)direct_groups / h->groups;
	h->stats[6] = T;
	h->stats[7] = direct_ok ? 1.0 : 0.0;
	if (direct_groups > 0 || direct_ok) balance(h, t_gpu - t_cpu, t3 - t0, direct_ok && p.cpu_groups > 0);
	return 0;
}

extern "C" {

int spl_host_step(spl_host_t* h, const spl_envs_t* envs, const spl_host_io_t* io, void* stream) {
	return host_run(h, envs, io, true, (cudaStream_t)stream);
}

int spl_host_observe(spl_host_t* h, const spl_envs_t* envs, const spl_host_io_t* io, void* stream) {
	return host_run(h, envs, io, false, (cudaStream_t)stream);
}

}  // extern "C"
