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

struct SplHostJob {
	const uint8_t* obs_u8 = nullptr;   // [n][297] staging (or the caller's own uint8 array)
	const uint32_t* side = nullptr;    // [n][4] staging
	spl_host_io_t io = {};             // the caller's arrays
	int64_t n = 0;
	int64_t cpu_groups = 0;            // groups [0, cpu_groups) are widened by the pool, the rest arrive already widened
	int threads = 1;
	int packed = 0;                    // staged observation bytes are nibble-packed unless a group's flag has bit 31 set
	const uint32_t* flags = nullptr;   // [groups] arrival flags in pinned host memory (null: everything is already there)
	uint32_t seq = 0;                  // value a flag takes when its group of this lock-step has landed
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
		obs_u8 = o.obs_u8, side = o.side, io = o.io, n = o.n, cpu_groups = o.cpu_groups, threads = o.threads, packed = o.packed;
		flags = o.flags, seq = o.seq, poll = o.poll, poll_ctx = o.poll_ctx, after_share0 = o.after_share0;
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

double spl_now_us();
SplHostJob* spl_pool_job();
int spl_pool_threads();
void spl_pool_run();
void spl_pool_run_custom();
void spl_job_share(const SplHostJob* job, int j, int64_t* start, int64_t* len);
void spl_expand_block(const uint8_t* obs_u8, const uint32_t* side, int64_t lo, int64_t hi, const spl_host_io_t* io);
