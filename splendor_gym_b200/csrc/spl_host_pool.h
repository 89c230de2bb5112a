// spl_host_pool.h -- internal interface between spl_host.cu (CUDA side of the host-buffer path) and
// spl_host_expand.cpp (worker pool + widening).  Not part of the C ABI.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <atomic>

#include <sched.h>

#include "../../include/splendor_b200.h"

#define SPL_HOST_GROUP 64  /* envs per arrival flag: 19,008 observation bytes + 1 KB of records */
#define SPL_POOL_MAX 64
/* Staging ring of one worker: SPL_RING_SLOTS slots of SPL_SLOT_BYTES.  A slot holds one 64-env group in transfer form:
 * bytes [0, 1024) the 16-byte records, then the observation bytes -- nibble-packed (9,504 B + 17 wide bytes per env) or raw
 * (19,008 B).  Rings bound the staging memory (32 x 20 KB per worker instead of a linear image of all envs) and keep the
 * link writing into the same few hundred KB; the price is flow control -- a worker publishes the tag of the last slot it
 * has emptied, and the push kernel does not reuse a slot before that.  (Measured: 8 ... 128 slots per worker are within
 * +-2 % of each other -- whether the staging bytes make a round trip through DRAM does not decide the speed of the
 * path; 4 slots throttle the link.) */
#define SPL_RING_SLOTS 32
#define SPL_SLOT_SIDE 1024
#define SPL_SLOT_BYTES (SPL_SLOT_SIDE + SPL_HOST_GROUP * 297 + 64) /* 20,096 = 314 x 64 */

struct SplHostJob {
	const uint8_t* obs_u8 = nullptr;   // [n][297] staging (or the caller's own uint8 array)
	const uint32_t* side = nullptr;    // [n][4] staging
	spl_host_io_t io = {};             // the caller's arrays
	int64_t n = 0;
	int64_t cpu_groups = 0;            // groups [0, cpu_groups) are widened by the pool, the rest arrive already widened
	int threads = 1;
	const uint32_t* flags = nullptr;   // [groups] arrival flags of the GPU-written share (pinned host memory)
	uint32_t seq = 0;                  // value such a flag takes when its group of this lock-step has landed
	// CPU share: per-worker staging rings (null: obs_u8 / side are complete linear arrays, nothing to wait for)
	const uint8_t* ring = nullptr;     // [threads][SPL_RING_SLOTS][SPL_SLOT_BYTES] pinned host memory
	const uint64_t* ring_flags = nullptr;  // [threads][SPL_RING_SLOTS]: tag_base + r + 1 (| bit 63: raw observation bytes) once round r landed
	uint64_t* consumed = nullptr;      // [threads][8] (one cache line each): tag of the last round the worker has emptied
	uint64_t tag_base = 0;
	int ring_slots = SPL_RING_SLOTS;
	int (*poll)(void*) = nullptr;      // worker 0, while waiting: non-zero = device error, give up
	void* poll_ctx = nullptr;
	void (*after_share0)(SplHostJob*) = nullptr;  // caller's thread, after its own share (waits for the GPU-written share)
	void (*custom)(void*, int) = nullptr;         // measurement jobs (spl_host_store_rate)
	void* custom_ctx = nullptr;
	std::atomic<int> abort{0};
	double t_first[SPL_POOL_MAX] = {};  // per worker: first group seen / share finished (us, CLOCK_MONOTONIC)
	double t_done[SPL_POOL_MAX] = {};

	SplHostJob() = default;
	SplHostJob& operator=(const SplHostJob& o) {
		obs_u8 = o.obs_u8, side = o.side, io = o.io, n = o.n, cpu_groups = o.cpu_groups, threads = o.threads;
		flags = o.flags, seq = o.seq, ring = o.ring, ring_flags = o.ring_flags, consumed = o.consumed, tag_base = o.tag_base, ring_slots = o.ring_slots, poll = o.poll, poll_ctx = o.poll_ctx, after_share0 = o.after_share0;
		custom = o.custom, custom_ctx = o.custom_ctx;
		abort.store(o.abort.load());
		return *this;
	}
};

struct SplRankAffinity {  // scope guard, see spl_host_expand.cpp
	cpu_set_t saved;
	bool active;
	SplRankAffinity();
	~SplRankAffinity();
};

struct SplCallerOffWorkers {  // scope guard, see spl_host_expand.cpp
	cpu_set_t saved;
	bool active;
	SplCallerOffWorkers();
	~SplCallerOffWorkers();
};

double spl_now_us();
SplHostJob* spl_pool_job();
int spl_pool_threads();
void spl_pool_run();
void spl_pool_start();
void spl_pool_wait();
void spl_pool_abort();
void spl_pool_run_custom();
void spl_job_share(const SplHostJob* job, int j, int64_t* start, int64_t* len);
void spl_expand_block(const uint8_t* obs_lo, const uint32_t* side_lo, int64_t lo, int64_t hi, const spl_host_io_t* io);
