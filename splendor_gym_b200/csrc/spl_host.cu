// spl_host.cu -- host-buffer entry points of the C ABI (include/splendor_b200.h: spl_host_*).
//
// The reference's callers hold NumPy arrays on the host (SplendorEnv.step, envs/splendor_env.py:51-90; the vector
// loop of ppo_splendor.py:235-285).  For them a lock-step is: actions host->device, the step kernel, results
// device->host.  The reference-typed results are 1,243 B per env-step, which PCIe caps at ~4e7 env-steps/s, so
// this path moves the COMPACT form (observation bytes + one 16-byte record per env = 313 B) and widens it on the
// host while the next chunk of the copy is still in flight:
//
//   stream:  H2D actions | step kernel (COMPACT) | D2H chunk 0 | ev0 | D2H chunk 1 | ev1 | ...
//   host  :                                         wait ev0 -> widen chunk 0 (OpenMP) | wait ev1 -> widen chunk 1 ...
//
// Measured on the round-1 box (16 vCPUs, PCIe D2H 51 GB/s): plain copies of the reference-typed arrays 4.3e7
// env-steps/s; this path 9.5e7 with int32 observations (bound by the HOST's memory bandwidth: 81 MB written per
// lock-step at 65,536 envs) and 1.3e8 with uint8 observations (bound by PCIe).
//
// The library owns the compact device buffers, the pinned staging and the events of one `spl_host_t`; the game
// state stays in the caller's (PyTorch's) tensors as everywhere else.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/splendor_b200.h"

#define SPL_OBS_DIM_ 297
#define SPL_HOST_MAX_CHUNKS 64

int spl_launch_compact(const spl_envs_t* e, const spl_step_io_t* io, bool do_step, uint8_t* obs_u8, void* side, cudaStream_t st);
void spl_expand_chunks(const uint8_t* obs_u8, const uint32_t* side, const int64_t* bounds, int chunks, const spl_host_io_t* io,
                       int (*wait)(void*, int), void* ctx, int* rc_out);
void spl_parallel_copy(void* dst, const void* src, size_t bytes);

struct spl_host {
	int64_t n;
	int chunks;
	uint8_t* d_obs;    // [n][297] device
	uint4* d_side;     // [n] device
	int32_t* d_act;    // [n] device
	uint8_t* h_obs;    // pinned
	uint32_t* h_side;  // pinned, 4 words per env
	int32_t* h_act;    // pinned
	cudaEvent_t ev[SPL_HOST_MAX_CHUNKS];
	int device;
};

#define SPL_CUDA(x)                                \
	do {                                           \
		cudaError_t e_ = (x);                      \
		if (e_ != cudaSuccess) return (int)e_;     \
	} while (0)

static bool g_threads_set = false;

static void default_threads() {
	if (g_threads_set) return;
	g_threads_set = true;
	cpu_set_t set;
	int n = 0;
	if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
	if (n < 1) n = 1;
	// one process per GPU (torchrun): the ranks of a node share its cores
	const char* lws = getenv("LOCAL_WORLD_SIZE");
	if (lws && atoi(lws) > 1) n /= atoi(lws);
	// measured (16 vCPUs, tools/microbench/host_bw.c): streaming stores peak at 8-12 threads (200 GB/s) and drop to
	// 150 GB/s at 16; the path also needs a core for the caller -> use 5/8 of the cores
	n = (n * 5 + 4) / 8;
	if (n < 1) n = 1;
	if (n > 32) n = 32;
	const char* e = getenv("SPL_HOST_THREADS");
	if (e && atoi(e) > 0) n = atoi(e);
	spl_host_set_threads(n);
}

extern "C" {

int spl_host_destroy(spl_host_t* h);

int spl_host_create(int64_t n, int32_t chunks, spl_host_t** out) {
	if (n <= 0 || !out) return SPL_E_BADARG;
	if (chunks <= 0) {
		// ~1.2 MB of observation bytes per chunk (measured best for the int32 widening at 65,536 envs): long enough for the copy engine to run at PCIe speed, short enough
		// that the widening of the last chunk (the only part that is not overlapped) stays small
		chunks = (int)((n + 4095) / 4096);
		if (chunks < 1) chunks = 1;
		if (chunks > 16) chunks = 16;
		const char* e = getenv("SPL_HOST_CHUNKS");
		if (e && atoi(e) > 0) chunks = atoi(e);
	}
	if (chunks > SPL_HOST_MAX_CHUNKS) chunks = SPL_HOST_MAX_CHUNKS;
	default_threads();
	spl_host* h = (spl_host*)calloc(1, sizeof(spl_host));
	if (!h) return SPL_E_BADARG;
	h->n = n, h->chunks = chunks;
	cudaError_t e = cudaGetDevice(&h->device);
	if (e == cudaSuccess) e = cudaMalloc(&h->d_obs, (size_t)n * SPL_OBS_DIM_ + 16);
	if (e == cudaSuccess) e = cudaMalloc(&h->d_side, (size_t)n * 16);
	if (e == cudaSuccess) e = cudaMalloc(&h->d_act, (size_t)n * 4);
	if (e == cudaSuccess) e = cudaHostAlloc(&h->h_obs, (size_t)n * SPL_OBS_DIM_ + 16, cudaHostAllocDefault);
	if (e == cudaSuccess) e = cudaHostAlloc(&h->h_side, (size_t)n * 16, cudaHostAllocDefault);
	if (e == cudaSuccess) e = cudaHostAlloc(&h->h_act, (size_t)n * 4, cudaHostAllocDefault);
	for (int c = 0; c < chunks && e == cudaSuccess; c++) e = cudaEventCreateWithFlags(&h->ev[c], cudaEventDisableTiming);
	if (e != cudaSuccess) {  // nothing half-built is handed out (spl_host_destroy skips what was never created)
		spl_host_destroy(h);
		return (int)e;
	}
	*out = h;
	return 0;
}

int spl_host_destroy(spl_host_t* h) {
	if (!h) return 0;
	if (h->d_obs) cudaFree(h->d_obs);
	if (h->d_side) cudaFree(h->d_side);
	if (h->d_act) cudaFree(h->d_act);
	if (h->h_obs) cudaFreeHost(h->h_obs);
	if (h->h_side) cudaFreeHost(h->h_side);
	if (h->h_act) cudaFreeHost(h->h_act);
	for (int c = 0; c < h->chunks; c++)
		if (h->ev[c]) cudaEventDestroy(h->ev[c]);
	free(h);
	return 0;
}

static int host_run(spl_host_t* h, const spl_envs_t* envs, const spl_host_io_t* io, bool do_step, cudaStream_t st) {
	if (!h || !envs || !io || envs->n != h->n) return SPL_E_BADARG;
	if (do_step && !io->actions) return SPL_E_BADARG;
	const int64_t n = h->n;
	spl_step_io_t dio;
	memset(&dio, 0, sizeof(dio));
	if (do_step) {
		spl_parallel_copy(h->h_act, io->actions, (size_t)n * 4);
		SPL_CUDA(cudaMemcpyAsync(h->d_act, h->h_act, (size_t)n * 4, cudaMemcpyHostToDevice, st));
		dio.actions = h->d_act;
	}
	dio.stats = io->stats;
	dio.action_key = io->action_key, dio.action_t = io->action_t;
	dio.autoreset = io->autoreset;
	int rc = spl_launch_compact(envs, &dio, do_step, h->d_obs, h->d_side, st);
	if (rc) return rc;
	// chunk boundaries on multiples of 32 envs (tile = 9,504 B, keeps every copy 16-byte aligned).  The first chunk is
	// half a share so that the host starts widening early; the widening (host memory bandwidth) is the critical path
	// Without int32 widening there is little host work to overlap: two chunks keep the copy engine streaming.
	const int chunks = io->obs ? h->chunks : (h->chunks < 2 ? h->chunks : 2);
	const int64_t tiles = (n + 31) / 32;
	int64_t bounds[SPL_HOST_MAX_CHUNKS + 1];
	bounds[0] = 0;
	for (int c = 1; c <= chunks; c++) {
		int64_t b = tiles * (2 * c - 1) / (2 * chunks - 1) * 32;
		bounds[c] = b < n ? b : n;
	}
	bounds[chunks] = n;
	const bool want_obs = io->obs || io->obs_u8;
	// a pinned / registered uint8 destination receives the observation bytes straight from the copy engine
	uint8_t* obs_dst = h->h_obs;
	spl_host_io_t xio = *io;
	if (io->obs_u8 && !io->obs) {
		cudaPointerAttributes at;
		if (cudaPointerGetAttributes(&at, io->obs_u8) == cudaSuccess && at.type == cudaMemoryTypeHost) obs_dst = io->obs_u8;
		else cudaGetLastError();
		if (obs_dst == io->obs_u8) xio.obs_u8 = nullptr;
	}
	for (int c = 0; c < chunks; c++) {
		const int64_t b = bounds[c], e = bounds[c + 1];
		if (e > b) {
			if (want_obs)
				SPL_CUDA(cudaMemcpyAsync(obs_dst + b * SPL_OBS_DIM_, h->d_obs + b * SPL_OBS_DIM_, (size_t)(e - b) * SPL_OBS_DIM_, cudaMemcpyDeviceToHost, st));
			SPL_CUDA(cudaMemcpyAsync(h->h_side + 4 * b, h->d_side + b, (size_t)(e - b) * 16, cudaMemcpyDeviceToHost, st));
		}
		SPL_CUDA(cudaEventRecord(h->ev[c], st));
	}
	rc = 0;
	spl_expand_chunks(obs_dst, h->h_side, bounds, chunks, &xio,
	                  [](void* ctx, int c) -> int { return (int)cudaEventSynchronize(((spl_host*)ctx)->ev[c]); }, h, &rc);
	return rc;
}

int spl_host_step(spl_host_t* h, const spl_envs_t* envs, const spl_host_io_t* io, void* stream) {
	return host_run(h, envs, io, true, (cudaStream_t)stream);
}

int spl_host_observe(spl_host_t* h, const spl_envs_t* envs, const spl_host_io_t* io, void* stream) {
	return host_run(h, envs, io, false, (cudaStream_t)stream);
}

}  // extern "C"
