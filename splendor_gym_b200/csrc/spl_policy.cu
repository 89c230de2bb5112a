// spl_policy.cu -- the callers on either side of the step path ("next" rows 1-2), kept on
// the device so that a rollout has no host round-trip:
//   * scripted opponents of scripts/eval_suite.py:10-128 over (obs, mask)            -> spl_scripted_action
//   * masked categorical sample / argmax of policy logits (ppo_splendor.py:27-38,54-59;
//     scripts/eval_suite.py:131-141)                                                  -> spl_masked_sample
//   * generalised advantage estimation of ppo_splendor.py:299-314                     -> spl_gae
// These are small elementwise kernels; they are not on the roofline-critical path.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/splendor_b200.h"

#define SPL_FULL 0xFFFFFFFFu

extern int64_t g_launches;  // spl_kernels.cu

__device__ __forceinline__ uint4 pol_philox(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll 2
	for (int r = 0; r < 10; r++) {
		uint32_t h0 = __umulhi(0xD2511F53u, c.x), l0 = 0xD2511F53u * c.x;
		uint32_t h1 = __umulhi(0xCD9E8D57u, c.z), l1 = 0xCD9E8D57u * c.z;
		c = make_uint4(h1 ^ c.y ^ k0, l1, h0 ^ c.w ^ k1, l0);
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	return c;
}

__device__ __forceinline__ uint4 pol_rand(uint64_t key, uint64_t genv, uint64_t t) {
	return pol_philox(make_uint4((uint32_t)genv, (uint32_t)(genv >> 32), (uint32_t)t, (uint32_t)(t >> 32)), (uint32_t)key,
	                  (uint32_t)(key >> 32));
}

// k-th (0-based) set bit of a 64-bit set
__device__ __forceinline__ int pol_kth(uint64_t m, uint32_t k) {
	for (uint32_t i = 0; i < k; i++) m &= m - 1;
	return __ffsll((long long)m) - 1;
}
__device__ __forceinline__ int pol_first(uint64_t m) { return __ffsll((long long)m) - 1; }
__device__ __forceinline__ int pol_random_of(uint64_t m, uint32_t r) { return pol_kth(m, r % (uint32_t)__popcll(m)); }

#define RANGE(lo, hi) ((((1ull << ((hi) - (lo) + 1)) - 1ull)) << (lo))
#define M_TAKE3 RANGE(0, 9)
#define M_TAKE2 RANGE(10, 14)
#define M_BUYVIS RANGE(15, 26)
#define M_RESERVE RANGE(27, 41)
#define M_BUYRES RANGE(42, 44)

// warp-cooperative load of the [32 x 45] int8 mask tile -> one 45-bit set per lane
__device__ __forceinline__ uint64_t pol_load_mask(const int8_t* __restrict__ mask, int64_t tile, int rows, uint8_t* smem, int lane) {
	const int8_t* g = mask + tile * 32 * SPL_NUM_ACTIONS;
	for (int b = lane; b < rows * SPL_NUM_ACTIONS; b += 32) smem[b] = (uint8_t)g[b];
	__syncwarp();
	uint64_t m = 0;
	if (lane < rows)
		for (int a = 0; a < SPL_NUM_ACTIONS; a++) m |= (uint64_t)(smem[lane * SPL_NUM_ACTIONS + a] != 0) << a;
	__syncwarp();
	return m;
}

// itertools.combinations(range(5),3) as colour bit sets, 5 bits per action (engine/encode.py:35)
#define POL_TAKE3_COMBOS 0x00039AB3B356CD67ull

__global__ void __launch_bounds__(128) spl_scripted_action_kernel(const int32_t* __restrict__ obs, const int8_t* __restrict__ mask, int64_t n,
                                                                int kind, uint64_t env_offset, uint64_t key, uint64_t t,
                                                                int32_t* __restrict__ actions) {
	__shared__ uint8_t rows_s[4][32 * SPL_NUM_ACTIONS];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t ntiles = (n + 31) >> 5;
	for (int64_t ti = (int64_t)blockIdx.x * 4 + warp; ti < ntiles; ti += (int64_t)gridDim.x * 4) {
		const int rows = (int)min((int64_t)32, n - ti * 32);
		const uint64_t m = pol_load_mask(mask, ti, rows, rows_s[warp], lane);
		if (lane >= rows) continue;
		const int64_t env = ti * 32 + lane;
		const int32_t* o = obs + env * SPL_OBS_DIM;
		int a = 0;
		if (m != 0) {
			const uint4 rnd = pol_rand(key, env_offset + (uint64_t)env, t);
			const uint64_t buys = m & (M_BUYVIS | M_BUYRES);
			if (kind == SPL_BOT_RANDOM) {  // wrappers/selfplay.py:66-73
				a = pol_random_of(m, rnd.x);
			} else if (kind == SPL_BOT_GREEDY_V1) {  // scripts/eval_suite.py:10-30
				if (buys) a = pol_first(buys);
				else if (m & M_TAKE2) a = pol_first(m & M_TAKE2);
				else if (m & M_TAKE3) a = pol_first(m & M_TAKE3);
				else if (m & M_RESERVE) a = pol_first(m & M_RESERVE);
				else a = pol_first(m);
			} else if (kind == SPL_BOT_BASIC_PRIORITY) {  // scripts/eval_suite.py:33-77 (np.random.choice -> Philox)
				if (m & M_BUYVIS) {
					int best = -1;
					uint64_t cand = 0;
					for (int s = 0; s < 12; s++) {
						if (!((m >> (15 + s)) & 1)) continue;
						int pt = o[32 + s * 13 + 2];  // points of the visible card in that slot
						if (pt > best) best = pt, cand = 0;
						if (pt == best) cand |= 1ull << (15 + s);
					}
					a = pol_random_of(cand, rnd.x);
				} else if (m & M_BUYRES) a = pol_random_of(m & M_BUYRES, rnd.x);
				else if (m & M_TAKE3) a = pol_random_of(m & M_TAKE3, rnd.x);
				else if (m & M_TAKE2) a = pol_random_of(m & M_TAKE2, rnd.x);
				else if (m & M_RESERVE) a = pol_random_of(m & M_RESERVE, rnd.x);
				else a = pol_first(m);
			} else {  // SPL_BOT_GREEDY_V2, scripts/eval_suite.py:80-128 with env_ref (bank = obs[0:5])
				if (buys) {
					a = pol_first(buys);
				} else if (m & M_TAKE2) {
					int best = 1 << 30;
					for (int c = 0; c < 5; c++)
						if (((m >> (10 + c)) & 1) && o[c] < best) best = o[c], a = 10 + c;  // min() keeps the first minimum
				} else if (m & M_TAKE3) {
					int best = 1 << 30;
					for (int k = 0; k < 10; k++) {
						if (!((m >> k) & 1)) continue;
						uint32_t combo = (uint32_t)(POL_TAKE3_COMBOS >> (5 * k)) & 31u;
						int sum = 0;
						for (int c = 0; c < 5; c++)
							if ((combo >> c) & 1) sum += o[c];
						if (sum < best) best = sum, a = k;
					}
				} else if (m & M_RESERVE) {
					a = 63 - __clzll((long long)(m & M_RESERVE));  // sorted(res, reverse=True)[0]
				} else a = pol_first(m);
			}
		}
		actions[env] = a;
	}
}

// masked categorical over 45 logits per env.  mode 0: sample (inverse CDF with one Philox uniform),
// mode 1: argmax (first maximum, like torch.argmax).  Rows without any legal action are left unmasked
// (ppo_splendor.py:27-38).  logprob (nullable) = log softmax of the masked logits at the chosen action,
// entropy (nullable) per env.
template <typename LT>
__global__ void __launch_bounds__(128) spl_masked_sample_kernel(const LT* __restrict__ logits, int64_t pitch, const int8_t* __restrict__ mask,
                                                              int64_t n, int mode, uint64_t env_offset, uint64_t key, uint64_t t,
                                                              int32_t* __restrict__ actions, float* __restrict__ logprob,
                                                              float* __restrict__ entropy) {
	__shared__ uint8_t rows_s[4][32 * SPL_NUM_ACTIONS];
	__shared__ float lg_s[4][32 * SPL_NUM_ACTIONS + 1];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t ntiles = (n + 31) >> 5;
	for (int64_t ti = (int64_t)blockIdx.x * 4 + warp; ti < ntiles; ti += (int64_t)gridDim.x * 4) {
		const int rows = (int)min((int64_t)32, n - ti * 32);
		uint64_t m = pol_load_mask(mask, ti, rows, rows_s[warp], lane);
		const LT* g = logits + ti * 32 * pitch;
		if (pitch == SPL_NUM_ACTIONS) {
			for (int b = lane; b < rows * SPL_NUM_ACTIONS; b += 32) lg_s[warp][b] = (float)g[b];  // coalesced tile load
		} else {  // padded rows (e.g. an output layer rounded up to 48 columns): skip the padding
			for (int b = lane; b < rows * SPL_NUM_ACTIONS; b += 32) {
				const int r = b / SPL_NUM_ACTIONS, a = b - r * SPL_NUM_ACTIONS;
				lg_s[warp][b] = (float)g[r * pitch + a];
			}
		}
		__syncwarp();
		if (lane < rows) {
			const int64_t env = ti * 32 + lane;
			const float* l = &lg_s[warp][lane * SPL_NUM_ACTIONS];  // stride 45 words: conflict-free
			if (m == 0) m = (1ull << SPL_NUM_ACTIONS) - 1ull;
			float mx = -INFINITY;
			int amax = 0;
			for (int a = 0; a < SPL_NUM_ACTIONS; a++)
				if (((m >> a) & 1) && l[a] > mx) mx = l[a], amax = a;
			float z = 0.0f;
			for (int a = 0; a < SPL_NUM_ACTIONS; a++)
				if ((m >> a) & 1) z += expf(l[a] - mx);
			int choice = amax;
			if (mode == 0) {
				const uint4 rnd = pol_rand(key, env_offset + (uint64_t)env, t);
				const float u = ((rnd.x >> 8) + 0.5f) * (1.0f / 16777216.0f) * z;  // uniform in (0, z)
				float acc = 0.0f;
				choice = -1;
				int last = amax;
				for (int a = 0; a < SPL_NUM_ACTIONS; a++) {
					if (!((m >> a) & 1)) continue;
					acc += expf(l[a] - mx);
					last = a;
					if (choice < 0 && u < acc) choice = a;
				}
				if (choice < 0) choice = last;
			}
			actions[env] = choice;
			const float logz = mx + logf(z);
			if (logprob) logprob[env] = l[choice] - logz;
			if (entropy) {
				float h = 0.0f;
				for (int a = 0; a < SPL_NUM_ACTIONS; a++)
					if ((m >> a) & 1) {
						float lp = l[a] - logz;
						h -= expf(lp) * lp;
					}
				entropy[env] = h;
			}
		}
		__syncwarp();
	}
}

// GAE (ppo_splendor.py:307-314): reverse scan over T per env, coalesced over envs.
__global__ void spl_gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ terminals,
                               const float* __restrict__ last_values, int T, int64_t n, float gamma, float lam,
                               float* __restrict__ advantages, float* __restrict__ returns) {
	int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n) return;
	float lastgaelam = 0.0f;
	float nextvalues = last_values[e];
	for (int t = T - 1; t >= 0; t--) {
		const int64_t i = (int64_t)t * n + e;
		const float nonterminal = terminals[i] ? 0.0f : 1.0f;
		const float v = values[i];
		const float delta = rewards[i] + gamma * nextvalues * nonterminal - v;
		lastgaelam = delta + gamma * lam * nonterminal * lastgaelam;
		advantages[i] = lastgaelam;
		returns[i] = lastgaelam + v;
		nextvalues = v;
	}
}

extern "C" {

int spl_scripted_action(const int32_t* obs, const int8_t* mask, int64_t n, int kind, uint64_t env_offset, uint64_t key, uint64_t t,
                        int32_t* actions, void* stream) {
	if (!mask || !actions || n <= 0 || kind < 0 || kind > SPL_BOT_GREEDY_V2) return SPL_E_BADARG;
	if (!obs && kind >= SPL_BOT_BASIC_PRIORITY) return SPL_E_BADARG;
	int64_t ctas = ((n + 31) / 32 + 3) / 4;
	spl_scripted_action_kernel<<<(int)(ctas < 148 * 8 ? ctas : 148 * 8), 128, 0, (cudaStream_t)stream>>>(obs, mask, n, kind, env_offset, key, t, actions);
	g_launches++;
	return (int)cudaGetLastError();
}

int spl_masked_sample(const float* logits, const int8_t* mask, int64_t n, int mode, uint64_t env_offset, uint64_t key, uint64_t t,
                      int32_t* actions, float* logprob, float* entropy, void* stream) {
	if (!logits || !mask || !actions || n <= 0 || (mode != 0 && mode != 1)) return SPL_E_BADARG;
	int64_t ctas = ((n + 31) / 32 + 3) / 4;
	spl_masked_sample_kernel<float><<<(int)(ctas < 148 * 8 ? ctas : 148 * 8), 128, 0, (cudaStream_t)stream>>>(
	    logits, SPL_NUM_ACTIONS, mask, n, mode, env_offset, key, t, actions, logprob, entropy);
	g_launches++;
	return (int)cudaGetLastError();
}

int spl_masked_sample_f16(const void* logits, int64_t pitch, const int8_t* mask, int64_t n, int mode, uint64_t env_offset, uint64_t key,
                          uint64_t t, int32_t* actions, float* logprob, float* entropy, void* stream) {
	if (!logits || !mask || !actions || n <= 0 || pitch < SPL_NUM_ACTIONS || (mode != 0 && mode != 1)) return SPL_E_BADARG;
	int64_t ctas = ((n + 31) / 32 + 3) / 4;
	spl_masked_sample_kernel<__half><<<(int)(ctas < 148 * 8 ? ctas : 148 * 8), 128, 0, (cudaStream_t)stream>>>(
	    (const __half*)logits, pitch, mask, n, mode, env_offset, key, t, actions, logprob, entropy);
	g_launches++;
	return (int)cudaGetLastError();
}

int spl_gae(const float* rewards, const float* values, const uint8_t* terminals, const float* last_values, int32_t T, int64_t n, float gamma,
            float lam, float* advantages, float* returns, void* stream) {
	if (!rewards || !values || !terminals || !last_values || !advantages || !returns || T <= 0 || n <= 0) return SPL_E_BADARG;
	spl_gae_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rewards, values, terminals, last_values, T, n, gamma, lam,
	                                                                           advantages, returns);
	g_launches++;
	return (int)cudaGetLastError();
}

}  // extern "C"
