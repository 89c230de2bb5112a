// spl_kernels.cu -- sm_100a kernels + C ABI of the batched Splendor engine (include/splendor_b200.h).
//
// Execution model: one environment per lane, one warp per tile of 32 environments, persistent CTAs
// looping over tiles.  The branchy integer rules run per lane on the packed 64-byte state
// (spl_core.cuh); everything that touches HBM is warp-cooperative:
//   * state   : 4 planes of uint4[N]  -> 4 fully coalesced LDG.128 / STG.128 per warp
//   * obs     : each lane encodes its 297 entries as BYTES into a flat [32 x 297]-byte shared-memory
//               tile (lane-dependent byte shift + neighbour shuffle so rows abut exactly); the warp then
//               streams the tile out as 2376 int4 = 38,016 contiguous, 16-B aligned bytes of int32
//   * mask    : 45-bit ballot-style bitset per lane, exchanged by shuffle and expanded to the contiguous
//               [32 x 45] int8 tile as 90 int4 stores
// No tensor cores: the path is integer/branch logic bounded by HBM writes (1,370 B per env-step).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "spl_core.cuh"
#include "spl_tables_host.h"

#define SPL_TILE_WORDS 2376 /* 32 envs * 297 bytes / 4 */
#define SPL_FULL 0xFFFFFFFFu
#define SPL_DECK_SMEM 100 /* per-lane deck row in shared memory: 25 words (odd) to spread banks */

__device__ SplTables g_tables;
#ifdef SPL_DEBUG_SM_UNITS
__device__ unsigned int g_sm_units[256];  // diagnostic build (tools/sm_units.py): warp-lock-steps processed per SM
#endif
__device__ uint64_t g_ret_table[SPL_RET_TABLE_LEN];
#ifdef SPL_DEBUG_PHASES
// diagnostic build (tools/step_phases.py): per-warp SM-clock stamps at the phase boundaries of the single-step kernel
#define SPL_PHASE_SLOTS 13
#define SPL_PHASE_WARPS 65536
__device__ unsigned long long g_phase[SPL_PHASE_SLOTS][SPL_PHASE_WARPS];
__device__ __forceinline__ void spl_phase_stamp(int k) {
	if ((threadIdx.x & 31) == 0) {
		const unsigned wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
		unsigned long long t;
		if (k == 10 || k == 12) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
		else if (k == 11) {
			unsigned sm;
			asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
			t = sm;
		} else asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
		if (wid < SPL_PHASE_WARPS) g_phase[k][wid] = t;
	}
}
#define SPL_PH(k) spl_phase_stamp(k)
#else
#define SPL_PH(k)
#endif

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) -- counter-based stream for action sampling and native shuffles
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 spl_philox(uint4 c, uint32_t k0, uint32_t k1) {
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

// index of the k-th (0-based) set bit of a 45-bit set
__device__ __forceinline__ uint32_t spl_kth_set_bit(uint64_t m, uint32_t k) {
	uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
	uint32_t nlo = __popc(lo);
	uint32_t word = lo, base = 0;
	if (k >= nlo) {
		k -= nlo;
		word = hi;
		base = 32;
	}
#pragma unroll
	for (int sh = 16; sh >= 1; sh >>= 1) {
		uint32_t c = __popc(word & ((1u << sh) - 1u));
		if (k >= c) {
			k -= c;
			word >>= sh;
			base += sh;
		}
	}
	return base;
}

__device__ __forceinline__ int32_t spl_sample_action(uint64_t m, uint64_t key, uint64_t genv, uint64_t t) {
	uint32_t n = __popcll(m);
	if (n == 0) return 0;  // wrappers/selfplay.py:66-73
	uint4 r = spl_philox(make_uint4((uint32_t)genv, (uint32_t)(genv >> 32), (uint32_t)t, (uint32_t)(t >> 32)),
	                     (uint32_t)key, (uint32_t)(key >> 32));
	return (int32_t)spl_kth_set_bit(m, r.x % n);
}

// ------------------------------------------------------------------------------------------------
// warp-cooperative tile output
// ------------------------------------------------------------------------------------------------
// [32 x 45] int8 action-mask tile from one 45-bit set per lane.  `rows` = valid envs in the tile.
__device__ __forceinline__ void spl_store_mask_tile(int8_t* gtile, uint64_t m, int lane, int rows) {
	const int tile_bytes = rows * SPL_NUM_ACTIONS;
#pragma unroll 1
	for (int it = 0; it < 3; it++) {
		int q = lane + 32 * it;  // int4 index, 90 per full tile
		int byte0 = 16 * q;
		int env_lo = byte0 / SPL_NUM_ACTIONS;
		int a0 = byte0 - env_lo * SPL_NUM_ACTIONS;
		uint64_t m_lo = __shfl_sync(SPL_FULL, m, env_lo & 31);
		uint64_t m_hi = __shfl_sync(SPL_FULL, m, (env_lo + 1) & 31);
		uint32_t bits = (uint32_t)((m_lo >> a0) | (m_hi << (SPL_NUM_ACTIONS - a0))) & 0xFFFFu;
		int4 v;
		v.x = (int)spl_spread4(bits);
		v.y = (int)spl_spread4(bits >> 4);
		v.z = (int)spl_spread4(bits >> 8);
		v.w = (int)spl_spread4(bits >> 12);
		if (byte0 + 16 <= tile_bytes) {
			__stcs(reinterpret_cast<int4*>(gtile) + q, v);
		} else if (byte0 < tile_bytes) {  // ragged last tile
			for (int j = 0; j < 16 && byte0 + j < tile_bytes; j++) gtile[byte0 + j] = (int8_t)((bits >> j) & 1u);
		}
	}
}

// Stage one lane's observation (75 words of bytes) into the flat byte tile.  Row r starts at byte 297 r,
// i.e. word (297 r)>>2 with a byte shift of r&3; a lane writes exactly the tile words that START inside
// its row, completing its last word with the first bytes of the next lane's row (= that lane's packed
// word 0, fetched by shuffle).
struct SplObsStager {
	uint32_t* dst;
	uint32_t shift8, prev, next_r0;
	__device__ __forceinline__ SplObsStager(uint32_t* tile, int lane, uint32_t w0)
	    : dst(tile + ((SPL_OBS_DIM * lane) >> 2)), shift8(8u * (lane & 3)), prev(0) {
		next_r0 = __shfl_down_sync(SPL_FULL, w0, 1);
	}
	__device__ __forceinline__ void first(uint32_t v) {
		if (shift8 == 0) dst[0] = v;
		prev = v;
	}
	__device__ __forceinline__ void put(int k, uint32_t v) {
		dst[k] = __funnelshift_l(prev, v, shift8);
		prev = v;
	}
	__device__ __forceinline__ void last(uint32_t v) { put(74, v | (next_r0 << 8)); }
};

// stream a staged tile to global memory as int32.  Full tile: 2376 x (LDS.32 -> 4 PRMT byte-extracts -> STG.128
// streaming store), the warp writing 512 contiguous bytes per instruction; otherwise (ragged last tile / unaligned
// caller buffer) per entry.  (256-bit stores, st.global.v8.b32 = STG.256 on sm_100a, were measured 8 % SLOWER here.)
__device__ __forceinline__ void spl_store_obs_tile(int32_t* gtile, const uint32_t* tile, int lane, int rows, bool vec) {
	if (rows == 32 && vec) {
		const uint32_t* t1 = tile + lane;
		int4* g4 = reinterpret_cast<int4*>(gtile) + lane;
#pragma unroll 2
		for (int q = lane; q < SPL_TILE_WORDS; q += 32, t1 += 32, g4 += 32) {
			const uint32_t v = *t1;
			__stcs(g4, make_int4((int)__byte_perm(v, 0, 0x4440), (int)__byte_perm(v, 0, 0x4441), (int)__byte_perm(v, 0, 0x4442),
			                     (int)__byte_perm(v, 0, 0x4443)));
		}
	} else {
		const uint8_t* tb = reinterpret_cast<const uint8_t*>(tile);
		for (int e = lane; e < rows * SPL_OBS_DIM; e += 32) gtile[e] = (int32_t)tb[e];
	}
}

// the staged tile IS the byte image of the [32][297] uint8 observation tile: 594 int4 per warp (host path)
__device__ __forceinline__ void spl_store_obs_tile_u8(uint8_t* gtile, const uint32_t* tile, int lane, int rows, bool vec) {
	if (rows == 32 && vec) {
		const int4* t4 = reinterpret_cast<const int4*>(tile);
		int4* g4 = reinterpret_cast<int4*>(gtile);
#pragma unroll 2
		for (int q = lane; q < SPL_TILE_WORDS / 4; q += 32) __stcs(g4 + q, t4[q]);
	} else {
		const uint8_t* tb = reinterpret_cast<const uint8_t*>(tile);
		for (int e = lane; e < rows * SPL_OBS_DIM; e += 32) gtile[e] = tb[e];
	}
}

// Policy-ready observation tile: fp16 [32][SPL_OBS_F16_PITCH] (the 297 entries + 7 zero columns, so that a row is 608
// bytes = 38 int4 and the first Linear of the policy sees a K that is a multiple of 8 -- K = 297 sends a half-precision
// GEMM down a misaligned slow path, 7x slower on B200).  Every entry is an integer < 256, exact in fp16: two bytes are
// merged with 0x64 into the fp16 pair (1024 + b0, 1024 + b1) by one PRMT and 1024 is subtracted by one HSUB2.
__device__ __forceinline__ uint32_t spl_bytes_to_half2(uint32_t v, uint32_t sel) {
	const uint32_t x = __byte_perm(v, 0x64646464u, sel), k = 0x64006400u;
	const __half2 h = __hsub2(*reinterpret_cast<const __half2*>(&x), *reinterpret_cast<const __half2*>(&k));
	return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void spl_store_obs_tile_f16(__half* gtile, const uint32_t* tile, int lane, int rows, bool vec) {
	constexpr int ROW4 = SPL_OBS_F16_PITCH / 8;  // int4 per row
	if (rows == 32 && vec) {
		int4* g4 = reinterpret_cast<int4*>(gtile);
#pragma unroll 2
		for (int q = lane; q < 32 * ROW4; q += 32) {
			const int r = q / ROW4, j = q - r * ROW4;
			const int b = SPL_OBS_DIM * r + 8 * j;  // first source byte in the flat byte tile
			const uint32_t* w = tile + (b >> 2);
			const bool last = j == ROW4 - 1;        // columns 296..303: one entry + the zero padding
			const uint32_t w0 = w[0], w1 = last ? 0u : w[1], w2 = last ? 0u : w[2];
			const uint32_t sh = 8u * (uint32_t)(b & 3);
			uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
			if (last) lo &= 0xFFu;
			__stcs(g4 + q, make_int4((int)spl_bytes_to_half2(lo, 0x5140), (int)spl_bytes_to_half2(lo, 0x7362),
			                         (int)spl_bytes_to_half2(hi, 0x5140), (int)spl_bytes_to_half2(hi, 0x7362)));
		}
	} else {
		const uint8_t* tb = reinterpret_cast<const uint8_t*>(tile);
		for (int e = lane; e < rows * SPL_OBS_F16_PITCH; e += 32) {
			const int r = e / SPL_OBS_F16_PITCH, c = e - r * SPL_OBS_F16_PITCH;
			gtile[e] = __ushort2half_rn(c < SPL_OBS_DIM ? (unsigned short)tb[SPL_OBS_DIM * r + c] : (unsigned short)0);
		}
	}
}

// reward of envs/splendor_env.py:61-80 as a small code (the host path moves one byte, not a float)
__device__ __forceinline__ uint32_t spl_reward_code(float r) {
	return r == 0.0f ? SPL_REWARD_CODE_ZERO : r == 1.0f ? SPL_REWARD_CODE_WIN : r == -1.0f ? SPL_REWARD_CODE_LOSS
	     : r == -0.1f ? SPL_REWARD_CODE_LIMIT : SPL_REWARD_CODE_ILLEGAL;
}

// ------------------------------------------------------------------------------------------------
// Native (Philox) deal: a uniformly random permutation of each deck and of the nobles by SORTING
// RANDOM KEYS, done by the whole warp for ONE environment (the rest of the warp would otherwise idle
// while a single lane ran a 100-step Fisher-Yates).  Element e (cards 0..89, nobles 90..99) gets the key
// philox4x32-10(ctr = {e/4, env_lo, env_hi, episode}, key = seed)[e%4] with its low 7 bits replaced by e; its
// position in its deck is the rank of that key among the elements of the same tier.  The deal is a pure function of
// (seed, global env id, episode), so it does not depend on the GPU count or on which kernel performs it.
// `scratch` = >= 132 words of shared memory private to the warp.  Writes the 96-byte deck row to `gdeck`.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void spl_coop_deal(uint64_t seed, uint64_t genv, uint32_t episode, uint32_t* scratch, int lane,
                                              uint8_t* gdeck, uint32_t board[3], uint32_t& nobles, uint32_t& tops) {
	uint32_t* keys = scratch;                                      // [100] (+4 pad)
	uint8_t* sdeck = reinterpret_cast<uint8_t*>(scratch + 104);    // [96]
	uint8_t* snob = reinterpret_cast<uint8_t*>(scratch + 128);     // [10] (+pad)
	if (lane < 25) {
		uint4 c = spl_philox(make_uint4((uint32_t)lane, (uint32_t)genv, (uint32_t)(genv >> 32), episode), (uint32_t)seed,
		                     (uint32_t)(seed >> 32));
		// low 7 bits <- element index: keys become unique, so a rank is a plain count of smaller keys
		// (25 random bits decide the order; the index only breaks the ~1e-4-probable ties)
		const uint32_t e = 4u * (uint32_t)lane;
		c.x = (c.x & ~127u) | e, c.y = (c.y & ~127u) | (e + 1), c.z = (c.z & ~127u) | (e + 2), c.w = (c.w & ~127u) | (e + 3);
		*reinterpret_cast<uint4*>(keys + 4 * lane) = c;
	}
	if (lane < 24) reinterpret_cast<uint32_t*>(sdeck)[lane] = 0xFFFFFFFFu;
	__syncwarp();
	{  // tier 1: elements 0..39, two per lane for lanes 0..7
		const uint32_t e0 = lane, e1 = 32 + lane;
		const uint32_t k0 = keys[e0], k1 = keys[e1 < 40 ? e1 : 0];
		uint32_t r0 = 0, r1 = 0;
#pragma unroll 4
		for (uint32_t i = 0; i < 40; i++) {
			uint32_t ki = keys[i];
			r0 += ki < k0;
			r1 += ki < k1;
		}
		sdeck[r0] = (uint8_t)e0;
		if (lane < 8) sdeck[r1] = (uint8_t)e1;
	}
	{  // tier 2: elements 40..69 | tier 3: 70..89 | nobles: 90..99 -- one element per lane each
		const uint32_t e2 = 40 + lane, e3 = 70 + lane, e4 = 90 + lane;
		const uint32_t k2 = keys[lane < 30 ? e2 : 40], k3 = keys[lane < 20 ? e3 : 70], k4 = keys[lane < 10 ? e4 : 90];
		uint32_t r2 = 0, r3 = 0, r4 = 0;
#pragma unroll 2
		for (uint32_t i = 40; i < 70; i++) {
			uint32_t ki = keys[i];
			r2 += ki < k2;
		}
#pragma unroll 2
		for (uint32_t i = 70; i < 90; i++) {
			uint32_t ki = keys[i];
			r3 += ki < k3;
		}
#pragma unroll 2
		for (uint32_t i = 90; i < 100; i++) {
			uint32_t ki = keys[i];
			r4 += ki < k4;
		}
		if (lane < 30) sdeck[40 + r2] = (uint8_t)e2;
		if (lane < 20) sdeck[70 + r3] = (uint8_t)e3;
		if (lane < 10) snob[r4] = (uint8_t)lane;
	}
	__syncwarp();
	// deal four cards per tier by pop() from the END of the deck (engine/state.py:190-191); first three nobles (:194-195)
	const uint32_t* d32 = reinterpret_cast<const uint32_t*>(sdeck);
	board[0] = __byte_perm(d32[9], 0, 0x0123);                          // deck1[39],[38],[37],[36]
	board[1] = __byte_perm(d32[16], d32[17], 0x2345);                   // deck2 = bytes 40..69: [69],[68],[67],[66]
	board[2] = __byte_perm(d32[21], d32[22], 0x2345);                   // deck3 = bytes 70..89: [89],[88],[87],[86]
	nobles = (uint32_t)snob[0] | ((uint32_t)snob[1] << 8) | ((uint32_t)snob[2] << 16);
	tops = (uint32_t)sdeck[35] | ((uint32_t)sdeck[40 + 25] << 8) | ((uint32_t)sdeck[70 + 15] << 16);  // tops after dealing 4
	if (lane < 24) reinterpret_cast<uint32_t*>(gdeck)[lane] = d32[lane];
	__syncwarp();
}

__device__ __forceinline__ void spl_state_from_deal(SplState& s, const uint32_t board[3], uint32_t nobles) {
	spl_fresh_state(s);
	s.board[0] = board[0], s.board[1] = board[1], s.board[2] = board[2];
	s.deckn = 36u | (26u << 8) | (16u << 16);
	s.nobles = nobles;
}

// ------------------------------------------------------------------------------------------------
// step / observe / rollout kernels
// ------------------------------------------------------------------------------------------------
#define SPL_RESET_NONE 0
#define SPL_RESET_WORKLIST 1 /* queue finished envs for spl_reset_kernel (MT19937 shuffles need 2.5 KB of state per env) */
#define SPL_RESET_FUSED 2    /* deal the new episode right here (Philox) */
#define SPL_RESET_SPARE 3    /* MT19937 with prefetched deals of the NEXT episode(s) per env (spl_envs_t.spare): take it right here */
#define SPL_RESET_SPARE_INLINE 4 /* the same inside the rollout kernel: a missing spare is dealt in place by one lane (slow, rare) */

// Engine seed of (env, episode) under SPL_SHUFFLE_MT19937: the library's own schedule, or -- replay of a reference
// run -- the caller's table of the seeds its SplendorEnv instances drew from their PCG64 streams on every auto-reset
// (envs/splendor_env.py:42-43; spl_envs_t.episode_seeds).  Episode 0 is the one spl_reset started (its seeds argument).
__device__ __forceinline__ uint64_t spl_episode_seed(uint64_t seed_base, uint64_t env_offset, int64_t env, uint32_t ep,
                                                      const uint64_t* table, int table_eps) {
	if (table != nullptr && ep >= 1u && ep <= (uint32_t)table_eps) return table[env * table_eps + (int64_t)(ep - 1u)];
	return (seed_base + 1000003ull * ep + env_offset + (uint64_t)env) % 2147483647ull;
}

struct StepParams {
	uint4* state;
	int64_t stride;
	uint8_t* decks;
	uint32_t* episode;
	int32_t* scratch;
	int64_t n;
	uint64_t env_offset, seed_base;
	const uint64_t* ep_seeds;  // spl_envs_t.episode_seeds [n][ep_seed_count] or nullptr
	int ep_seed_count;
	const int32_t* actions;
	const uint8_t* active;
	int32_t* obs;
	int8_t* mask;
	float* reward;
	uint8_t* terminated;
	uint8_t* info;
	unsigned long long* stats;
	int32_t* next_action;
	uint64_t action_key, action_t;
	const uint64_t* action_t_base;
	int reset_mode;
	uint8_t* spare;      // SPL_RESET_SPARE: [n][slots][96] prefetched deals, then the int32 refill list (header of 4 + n*slots entries)
	int spare_slots;     // deals kept ahead per env (ring indexed by episode % slots)
	int spare_async;     // a refill may be dealing into the ring while this launch runs: read a slot's flag word first (acquire)
	int vec_ok;  // obs / mask bases are 16-byte aligned (and, for step-major buffers, every step's slice is)
	int steps;   // rollout kernel: lock-steps per launch; outputs are [steps][n][...], next_action is [steps+1][n]
	int sync;    // rollout kernel: CTA barrier per lock-step (keeps the warps of a CTA in the same code region)
	uint8_t* obs_u8;  // compact outputs (host path, spl_host.cu): observation tile as bytes [n][297] ...
	__half* obs_f16;  // policy-ready observation, fp16 [n][SPL_OBS_F16_PITCH] (F16 output mode; obs_u8 may accompany it)
	uint4* side;      // ... and one 16-byte record per env: legal-mask bits, reward code, terminated, info, sampled action
	int chunk, nchunks;  // rollout kernel: lock-steps per work unit, chunks per tile group (spl_chunk_bounds)
};

struct SplTile {
	const SplTables* T;
	uint32_t* smem;  // SPL_TILE_WORDS words private to the warp
	int lane;
	int64_t ti;      // tile index; env = ti*32 + lane
	int rows;        // valid envs in the tile
};

__device__ __forceinline__ uint32_t spl_ld_acquire(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

// initial_state(seed) by ONE lane, MT19937 state in `smem` (>= 656 words), deck row to global memory (defined below);
// results in smem[648..652]: board rows, visible nobles, deck tops (no reference arguments: nothing of the caller is
// forced onto the stack)
__device__ __noinline__ void spl_mt_deal_in_place(uint64_t seed, uint32_t* smem, uint8_t* deck_row);

// one SplendorEnv.step for the lane's env + episode statistics + same-step auto-reset
template <bool KNOWN_MASK>
__device__ __forceinline__ void spl_tile_step(const StepParams& p, const SplTile& tl, SplState& s, bool act, int32_t action, int64_t env,
                                              SplStepResult& r, uint64_t cur_mask = 0, uint32_t* tops = nullptr) {
	const int lane = tl.lane;
	r.reward = 0.0f, r.terminated = 0, r.info = 0;
	if (act) spl_env_step_t<KNOWN_MASK>(s, action, p.decks + env * SPL_DECK_STRIDE, tl.T, g_ret_table, r, cur_mask, tops);
	const bool finished = r.terminated && !(r.info & SPL_INFO_ERROR);
	const bool do_reset = act && r.terminated && p.reset_mode != SPL_RESET_NONE;
	if (do_reset) r.info |= SPL_INFO_RESET;
	// episode statistics: warp-reduce, one atomic per counter per warp
	if (p.stats != nullptr && __any_sync(SPL_FULL, finished)) {
		uint32_t wcode = (s.flags & SPL_FLAG_WINNER_MASK) >> SPL_FLAG_WINNER_SHIFT;
		bool nolegal = r.info & SPL_INFO_NOLEGAL_DRAW, limit = r.info & SPL_INFO_TURN_LIMIT;
		uint32_t v[8];
		v[SPL_STAT_EPISODES] = finished;
		v[SPL_STAT_NOLEGAL_DRAWS] = finished && nolegal;
		v[SPL_STAT_LIMIT_DRAWS] = finished && !nolegal && limit;
		v[SPL_STAT_P0_WINS] = finished && !nolegal && !limit && wcode == 1;
		v[SPL_STAT_P1_WINS] = finished && !nolegal && !limit && wcode == 2;
		v[SPL_STAT_TIE_DRAWS] = finished && !nolegal && !limit && wcode == 0;
		v[SPL_STAT_SUM_MOVES] = finished ? s.move : 0u;
		// terminal => to_play == 0 => perspective index == player index
		v[SPL_STAT_SUM_WINNER_PRESTIGE] = (finished && wcode) ? (wcode == 1 ? s.prestige[0] : s.prestige[1]) : 0u;
#pragma unroll
		for (int k = 0; k < 8; k++) {
			uint32_t tot = __reduce_add_sync(SPL_FULL, v[k]);
			if (lane == 0 && tot) atomicAdd(p.stats + k, (unsigned long long)tot);
		}
	}
	uint32_t rb = __ballot_sync(SPL_FULL, do_reset);
	if (rb == 0) return;
	if (p.reset_mode == SPL_RESET_SPARE || p.reset_mode == SPL_RESET_SPARE_INLINE) {
		// The deals of every env's NEXT episodes were computed ahead of time, off the critical path (one lane's
		// random.Random(seed) chain takes ~35 us): episode e of an env sits in slot e % slots of its ring, tagged with
		// e.  Copy it in and clear the slot's ready flag (refills find their work by scanning the flags).  A spare that is not there (more finishes of
		// one env between refills than the ring holds: cannot happen in legal play with the refill cadence used, but
		// hand-built states may) falls back to the work list below / the in-place deal of the rollout kernel.
		const int64_t R = p.spare_slots;
		uint32_t late = 0;
		for (uint32_t todo = rb; todo;) {
			const int src = __ffs(todo) - 1;
			todo &= todo - 1;
			const int64_t e = __shfl_sync(SPL_FULL, env, src);
			const uint32_t ep = __ldcg(p.episode + e) + 1u;  // the episode that starts now (L2: another SM may have bumped it)
			const int64_t code = e * R + (int64_t)(ep % (uint32_t)R);
			uint32_t* srow = reinterpret_cast<uint32_t*>(p.spare + code * SPL_DECK_STRIDE);
			// bytes 92..95 of the row: third noble, episode tag (16 bits), ready flag.  The dealer publishes them last, so when
			// a refill may be running concurrently they are read first, with acquire semantics, and the deck order after
			uint32_t w23 = p.spare_async ? spl_ld_acquire(srow + 23) : 0u;
			if (p.spare_async && ((w23 >> 24) == 0u || ((w23 >> 8) & 0xFFFFu) != (ep & 0xFFFFu))) {
				late |= 1u << src;
				continue;
			}
			const uint32_t wv = lane < 24 ? __ldcg(srow + lane) : 0u;
			if (!p.spare_async) w23 = __shfl_sync(SPL_FULL, wv, 23);
			if ((w23 >> 24) == 0u || ((w23 >> 8) & 0xFFFFu) != (ep & 0xFFFFu)) {
				late |= 1u << src;  // nothing usable in the slot (the next refill scan finds it by its flag / tag)
				continue;
			}
			if (lane < 24)  // deck row: bytes 90.. are padding there
				reinterpret_cast<uint32_t*>(p.decks + e * SPL_DECK_STRIDE)[lane] = lane == 22 ? (wv | 0xFFFF0000u) : (lane == 23 ? 0xFFFFFFFFu : wv);
			const uint32_t w8 = __shfl_sync(SPL_FULL, wv, 8), w9 = __shfl_sync(SPL_FULL, wv, 9), w16 = __shfl_sync(SPL_FULL, wv, 16);
			const uint32_t w17 = __shfl_sync(SPL_FULL, wv, 17), w21 = __shfl_sync(SPL_FULL, wv, 21), w22 = __shfl_sync(SPL_FULL, wv, 22);
			if (lane == 23) srow[23] = 0u;  // taken: the refill scan (spl_spare_scan_kernel) picks the slot up by this flag
			if (lane == src) {
				p.episode[env] = ep;
				uint32_t board[3];
				board[0] = __byte_perm(w9, 0, 0x0123);     // deck1[39],[38],[37],[36] (engine/state.py:190-191: pop() from the end)
				board[1] = __byte_perm(w16, w17, 0x2345);  // deck2 = bytes 40..69
				board[2] = __byte_perm(w21, w22, 0x2345);  // deck3 = bytes 70..89
				spl_state_from_deal(s, board, (w22 >> 16) | ((w23 & 0xFFu) << 16));  // bytes 90..92: the three visible nobles
				if (tops != nullptr) *tops = (w8 >> 24) | (((w16 >> 8) & 0xFFu) << 8) | (((w21 >> 8) & 0xFFu) << 16);  // bytes 35, 65, 85
			}
		}
		__syncwarp();  // the deck rows written above (lanes 0..23) are popped by their owning lanes in later steps
		rb = late;
		if (rb == 0) return;
		if (p.reset_mode == SPL_RESET_SPARE_INLINE) {
			while (rb) {  // warp-uniform; the serial generator runs on ONE lane with its 624 words in the warp's (idle) tile
				const int src = __ffs(rb) - 1;
				rb &= rb - 1;
				if (lane == src) {
					const uint32_t ep = __ldcg(p.episode + env) + 1u;
					p.episode[env] = ep;
					spl_mt_deal_in_place(spl_episode_seed(p.seed_base, p.env_offset, env, ep, p.ep_seeds, p.ep_seed_count), tl.smem,
					                     p.decks + env * SPL_DECK_STRIDE);
					uint32_t board[3] = {tl.smem[648], tl.smem[649], tl.smem[650]};
					spl_state_from_deal(s, board, tl.smem[651]);
					if (tops != nullptr) *tops = tl.smem[652];
				}
				__syncwarp();
			}
			return;
		}
	}
	if (p.reset_mode == SPL_RESET_WORKLIST || p.reset_mode == SPL_RESET_SPARE) {
		const bool mine = (rb >> lane) & 1u;
		int base = 0;
		if (lane == 0) base = atomicAdd(p.scratch, __popc(rb));
		base = __shfl_sync(SPL_FULL, base, 0);
		if (mine) p.scratch[4 + base + __popc(rb & ((1u << lane) - 1u))] = (int32_t)env;
	} else {
		while (rb) {  // warp-uniform loop over the lanes whose episode just ended
			const int src = __ffs(rb) - 1;
			rb &= rb - 1;
			uint32_t ep = 0;
			if (lane == src) {
				ep = __ldcg(p.episode + env) + 1u;  // L2: the previous chunk of this game may have run on another SM
				p.episode[env] = ep;
			}
			ep = __shfl_sync(SPL_FULL, ep, src);
			const int64_t e = __shfl_sync(SPL_FULL, env, src);
			uint32_t board[3], nobles, new_tops;
			spl_coop_deal(p.seed_base, p.env_offset + (uint64_t)e, ep, tl.smem, lane, p.decks + e * SPL_DECK_STRIDE, board, nobles, new_tops);
			if (lane == src) {
				spl_state_from_deal(s, board, nobles);
				if (tops != nullptr) *tops = new_tops;
			}
		}
	}
}

// outputs of the lane's (possibly new) state: legal mask tile, sampled next action, observation tile
// `m` = legal_moves of that state as a bit set (all-zero once terminal, envs/splendor_env.py:81)
__device__ __forceinline__ int32_t spl_tile_emit(const StepParams& p, const SplTile& tl, const SplState& s, const uint32_t* w, int64_t env,
                                                 bool valid, int32_t* obs, int8_t* mask, int32_t* next_action, uint64_t t, uint64_t m) {
	const int lane = tl.lane;
	int32_t sampled = 0;
	if (mask != nullptr) {
		if (p.vec_ok) spl_store_mask_tile(mask + tl.ti * 32 * SPL_NUM_ACTIONS, m, lane, tl.rows);
		else if (valid)
			for (int a = 0; a < SPL_NUM_ACTIONS; a++) mask[env * SPL_NUM_ACTIONS + a] = (int8_t)((m >> a) & 1);
	}
	if (next_action != nullptr) {
		sampled = spl_sample_action(m, p.action_key, p.env_offset + (uint64_t)env, t);
		if (valid) next_action[env] = sampled;
	}
	SPL_PH(6);
	if (obs != nullptr) {
		SplObsStager stage(tl.smem, lane, w[0]);
		spl_encode_observation(w, s, tl.T, stage);
		__syncwarp();
		SPL_PH(7);
		spl_store_obs_tile(obs + tl.ti * 32 * SPL_OBS_DIM, tl.smem, lane, tl.rows, p.vec_ok);
		__syncwarp();
	}
	SPL_PH(8);
	return sampled;
}

__device__ __forceinline__ void spl_load_state(const StepParams& p, int64_t env, bool valid, uint32_t* w) {
	if (valid) {
#pragma unroll
		for (int pl = 0; pl < SPL_STATE_PLANES; pl++) {
			uint4 v = p.state[pl * p.stride + env];
			w[4 * pl + 0] = v.x, w[4 * pl + 1] = v.y, w[4 * pl + 2] = v.z, w[4 * pl + 3] = v.w;
		}
	} else {
#pragma unroll
		for (int k = 0; k < 16; k++) w[k] = 0;
	}
}

__device__ __forceinline__ void spl_store_state(const StepParams& p, int64_t env, const uint32_t* w) {
#pragma unroll
	for (int pl = 0; pl < SPL_STATE_PLANES; pl++)
		p.state[pl * p.stride + env] = make_uint4(w[4 * pl + 0], w[4 * pl + 1], w[4 * pl + 2], w[4 * pl + 3]);
}

// constant tables -> shared memory, once per CTA: 128-bit loads, all of a thread's loads in flight before its first
// store (a dependent load/store pair per iteration cost a 1-warp CTA 16 global-latency round trips = 6 us)
template <int THREADS>
__device__ __forceinline__ const SplTables* spl_stage_tables(SplTables* T) {
	static_assert(sizeof(SplTables) % 16 == 0, "SplTables is copied as uint4");
	constexpr int N4 = (int)(sizeof(SplTables) / 16), PER = (N4 + THREADS - 1) / THREADS;
	const uint4* src = reinterpret_cast<const uint4*>(&g_tables);
	uint4* dst = reinterpret_cast<uint4*>(T);
	uint4 v[PER];
#pragma unroll
	for (int k = 0; k < PER; k++)
		if (k * THREADS + (int)threadIdx.x < N4) v[k] = src[k * THREADS + threadIdx.x];
#pragma unroll
	for (int k = 0; k < PER; k++)
		if (k * THREADS + (int)threadIdx.x < N4) dst[k * THREADS + threadIdx.x] = v[k];
	__syncthreads();
	return T;
}

// one lock-step (DO_STEP) or just encode_observation + legal_moves of the current states.
// COMPACT (host path): outputs are the uint8 observation tile + one 16-byte record per env instead of the
// reference-typed int32 / int8 / float arrays -- 313 B instead of 1,243 B per env-step cross PCIe and the host
// widens them (spl_host_expand.cpp).
//   OUT = SPL_OUT_I32 reference-typed int32 observation | SPL_OUT_COMPACT | SPL_OUT_F16 (fp16 policy input and / or bytes)
#define SPL_OUT_I32 0
#define SPL_OUT_COMPACT 1
#define SPL_OUT_F16 2
template <bool DO_STEP, int WPC, int OUT>
__global__ void __launch_bounds__(WPC * 32) spl_step_kernel(const StepParams p) {
	constexpr bool COMPACT = OUT == SPL_OUT_COMPACT;
	__shared__ SplTables Ts;
	__shared__ __align__(16) uint32_t tiles[WPC][SPL_TILE_WORDS];
	SplTile tl;
	SPL_PH(0);
	SPL_PH(10);
	SPL_PH(11);
	tl.T = spl_stage_tables<WPC * 32>(&Ts);
	SPL_PH(1);
	tl.lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	tl.smem = tiles[warp];
	const int64_t ntiles = (p.n + 31) >> 5;
	const uint64_t t = p.action_t + (p.action_t_base ? *p.action_t_base : 0ull);

	for (tl.ti = (int64_t)blockIdx.x * WPC + warp; tl.ti < ntiles; tl.ti += (int64_t)gridDim.x * WPC) {
		const int64_t env = tl.ti * 32 + tl.lane;
		const bool valid = env < p.n;
		tl.rows = (int)min((int64_t)32, p.n - tl.ti * 32);
		uint32_t w[16];
		spl_load_state(p, env, valid, w);
		SplState s;
		spl_unpack(w, s);
		SplStepResult r;
		r.reward = 0.0f, r.terminated = 0, r.info = 0;
		if (DO_STEP) {
			const bool act = valid && (p.active == nullptr || p.active[env] != 0);
#ifdef SPL_DEBUG_PHASES
			if (__shfl_xor_sync(SPL_FULL, w[0] + (act ? p.actions[env] : 0), 1) == 0xdeadbeefu) return;  // forces the loads to have landed
			SPL_PH(2);
#endif
			spl_tile_step<false>(p, tl, s, act, act ? p.actions[env] : 0, env, r);
			SPL_PH(3);
			if (act) {
				spl_pack(s, w);
				spl_store_state(p, env, w);
			}
			if (valid && !COMPACT) {
				p.reward[env] = r.reward;
				p.terminated[env] = (uint8_t)r.terminated;
				p.info[env] = (uint8_t)r.info;
			}
		}
		SPL_PH(4);
		uint64_t m = 0;
		if (COMPACT || p.mask != nullptr || p.next_action != nullptr) m = spl_is_terminal(s) ? 0ull : spl_legal_mask(s, tl.T);
#ifdef SPL_DEBUG_PHASES
		if (__shfl_xor_sync(SPL_FULL, (uint32_t)m, 1) == 0xdeadbeefu && m == 0x123456789ull) return;  // forces the mask to be complete
		SPL_PH(5);
#endif
		if (COMPACT) {
			const int32_t sampled = spl_sample_action(m, p.action_key, p.env_offset + (uint64_t)env, t);
			SplObsStager stage(tl.smem, tl.lane, w[0]);
			spl_encode_observation(w, s, tl.T, stage);
			__syncwarp();
			spl_store_obs_tile_u8(p.obs_u8 + tl.ti * 32 * SPL_OBS_DIM, tl.smem, tl.lane, tl.rows, p.vec_ok);
			__syncwarp();
			if (valid)
				__stcs(p.side + env, make_uint4((uint32_t)m, (uint32_t)(m >> 32) | (spl_reward_code(r.reward) << 16) | ((uint32_t)r.terminated << 24),
				                               (uint32_t)r.info | ((uint32_t)sampled << 8),
				                               // bit 0: a token count >= 16 (hand-built states only): this env's observation does not fit
				                               // the nibble-packed transfer form of spl_push_kernel
				                               ((s.bank | s.tok[0] | s.tok[1]) & 0xF0F0F0F0F0F0ull) != 0 ? 1u : 0u));
		} else {
			spl_tile_emit(p, tl, s, w, env, valid, OUT == SPL_OUT_F16 ? nullptr : p.obs, p.mask, p.next_action, t, m);
			if (OUT == SPL_OUT_F16) {
				SplObsStager stage(tl.smem, tl.lane, w[0]);
				spl_encode_observation(w, s, tl.T, stage);
				__syncwarp();
				if (p.obs_f16 != nullptr)
					spl_store_obs_tile_f16(p.obs_f16 + tl.ti * 32 * SPL_OBS_F16_PITCH, tl.smem, tl.lane, tl.rows, p.vec_ok);
				if (p.obs_u8 != nullptr) spl_store_obs_tile_u8(p.obs_u8 + tl.ti * 32 * SPL_OBS_DIM, tl.smem, tl.lane, tl.rows, p.vec_ok);
				__syncwarp();
			}
		}
		SPL_PH(9);
		SPL_PH(12);
	}
}

// ------------------------------------------------------------------------------------------------
// Single-step kernel, two warps per tile (EXPERIMENT, off by default: knob SPL_STEP_PAIRED=1; int32 observation output).
// Measured on B200 at 65,536 envs, same box, alternating runs: 26.17 us per lock-step against 24.78 us for the one-warp-per-tile
// kernel below (-5.6 %), bit-exact.  Kept as the record of that experiment.
// At 65,536 envs there are 4.6 warp slots per tile, but the rules need ~96 registers, which caps residency at 20 one-warp
// CTAs per SM.  Here a CTA is two warpgroups handling four tiles: warps 0-3 run the rules (state load, step, fused
// auto-reset, state store, observation encode) with a raised register budget, warps 4-7 are their helpers (legal mask of
// the new state, sampler, mask tile, half of the observation-tile stores) with a lowered one (setmaxnreg), so that four CTAs
// = 16 tiles fit an SM in one wave and a tile's mask / encode / store phases overlap instead of queueing in one warp.
// Hand-over per tile through shared memory: the packed state (16 words per lane) behind named barrier 1+k, the staged
// observation tile behind named barrier 5+k.  Same lane-local code (spl_core.cuh), same results bit for bit.
// ------------------------------------------------------------------------------------------------
#define SPL_PAIR_TILES 4
#define SPL_PAIR_REGS_RULES 88
#define SPL_PAIR_REGS_HELPER 40

__device__ __forceinline__ void spl_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void spl_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// every other 512-byte store of the tile (the two warps of a tile take the even / the odd ones)
__device__ __forceinline__ void spl_store_obs_tile_half(int32_t* gtile, const uint32_t* tile, int lane, int half) {
#pragma unroll 2
	for (int q = lane + 32 * half; q < SPL_TILE_WORDS; q += 64) {
		const uint32_t v = tile[q];
		__stcs(reinterpret_cast<int4*>(gtile) + q, make_int4((int)__byte_perm(v, 0, 0x4440), (int)__byte_perm(v, 0, 0x4441),
		                                                     (int)__byte_perm(v, 0, 0x4442), (int)__byte_perm(v, 0, 0x4443)));
	}
}

template <bool DO_STEP>
__global__ void __launch_bounds__(SPL_PAIR_TILES * 64, 4) spl_step_paired_kernel(const StepParams p) {
	__shared__ SplTables Ts;
	__shared__ __align__(16) uint32_t tiles[SPL_PAIR_TILES][SPL_TILE_WORDS];
	__shared__ uint32_t packed[SPL_PAIR_TILES][16][32];  // word-major: lane-consecutive, conflict-free
	SplTile tl;
	tl.T = spl_stage_tables<SPL_PAIR_TILES * 64>(&Ts);
	tl.lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int k = warp & (SPL_PAIR_TILES - 1);
	const bool helper = warp >= SPL_PAIR_TILES;
	tl.smem = tiles[k];
	const int64_t ntiles = (p.n + 31) >> 5;
	const uint64_t t = p.action_t + (p.action_t_base ? *p.action_t_base : 0ull);
	tl.ti = (int64_t)blockIdx.x * SPL_PAIR_TILES + k;
	const bool live = tl.ti < ntiles;  // (uniform per tile: both warps of a dead tile skip the hand-over)
	const int64_t env = tl.ti * 32 + tl.lane;
	const bool valid = live && env < p.n;
	tl.rows = live ? (int)min((int64_t)32, p.n - tl.ti * 32) : 0;
	const bool want_mask = p.mask != nullptr || p.next_action != nullptr;
	if (!helper) {
		asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SPL_PAIR_REGS_RULES));
		if (!live) return;
		uint32_t w[16];
		spl_load_state(p, env, valid, w);
		SplState s;
		spl_unpack(w, s);
		SplStepResult r;
		r.reward = 0.0f, r.terminated = 0, r.info = 0;
		if (DO_STEP) {
			const bool act = valid && (p.active == nullptr || p.active[env] != 0);
			spl_tile_step<false>(p, tl, s, act, act ? p.actions[env] : 0, env, r);
			if (act) {
				spl_pack(s, w);
				spl_store_state(p, env, w);
			}
			if (valid) {
				p.reward[env] = r.reward;
				p.terminated[env] = (uint8_t)r.terminated;
				p.info[env] = (uint8_t)r.info;
			}
		}
		if (want_mask) {  // hand the new state to the helper warp
#pragma unroll
			for (int j = 0; j < 16; j++) packed[k][j][tl.lane] = w[j];
			__threadfence_block();
			spl_bar_arrive(1 + k, 64);
		}
		if (p.obs != nullptr) {
			SplObsStager stage(tl.smem, tl.lane, w[0]);
			spl_encode_observation(w, s, tl.T, stage);
			__threadfence_block();
			spl_bar_sync(5 + k, 64);  // tile staged: both warps stream it out
			int32_t* g = p.obs + tl.ti * 32 * SPL_OBS_DIM;
			if (tl.rows == 32 && p.vec_ok) spl_store_obs_tile_half(g, tl.smem, tl.lane, 0);
			else spl_store_obs_tile(g, tl.smem, tl.lane, tl.rows, false);
		}
	} else {
		asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SPL_PAIR_REGS_HELPER));
		if (!live) return;
		if (want_mask) {
			spl_bar_sync(1 + k, 64);
			uint32_t w[16];
#pragma unroll
			for (int j = 0; j < 16; j++) w[j] = packed[k][j][tl.lane];
			SplState s;
			spl_unpack(w, s);
			const uint64_t m = spl_is_terminal(s) ? 0ull : spl_legal_mask(s, tl.T);
			if (p.mask != nullptr) {
				if (p.vec_ok) spl_store_mask_tile(p.mask + tl.ti * 32 * SPL_NUM_ACTIONS, m, tl.lane, tl.rows);
				else if (valid)
					for (int a = 0; a < SPL_NUM_ACTIONS; a++) p.mask[env * SPL_NUM_ACTIONS + a] = (int8_t)((m >> a) & 1);
			}
			if (p.next_action != nullptr) {
				const int32_t sampled = spl_sample_action(m, p.action_key, p.env_offset + (uint64_t)env, t);
				if (valid) p.next_action[env] = sampled;
			}
		}
		if (p.obs != nullptr) {
			spl_bar_sync(5 + k, 64);
			if (tl.rows == 32 && p.vec_ok) spl_store_obs_tile_half(p.obs + tl.ti * 32 * SPL_OBS_DIM, tl.smem, tl.lane, 1);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Rollout work queue.  SM store bandwidth on B200 is NOT uniform: a store-only microbenchmark
// (tools/microbench/store_bw.cu) hands 8 SMs ~2,200 tiles, ~108 SMs ~1,850 and 32 SMs ~1,430 when tiles are
// taken from an atomic counter, and reaches 7.3-7.5 TB/s that way against 6.0-6.4 TB/s for any static split
// (the slow SMs finish last, the fast ones idle for 30 % of the run).  The rollout kernel therefore cuts the
// job into work units (tile group g, chunk c of lock-steps), numbers them chunk-major (u = c*G + g, so the
// write front still moves linearly through the step-major rollout buffers) and lets persistent CTAs pull
// units from an atomic counter.  A group's packed state travels between chunks through HBM (128 B per env per
// chunk); unit (g, c) waits until `done[g]` says chunk c-1 of the same group has been published.  Units are
// handed out in increasing order and a unit only depends on a smaller one, so the oldest unit in flight never
// waits: no deadlock as long as units are only held by resident CTAs (they are: a CTA takes one when it runs).
// Chunk lengths: `chunk` lock-steps each, then the last chunk..2*chunk steps are halved down to 2 so that the
// tail of the launch (CTAs finding the queue empty) is short.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void spl_st_release(uint32_t* p, uint32_t v) {
	asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// `steps` lock-steps of uniform-random-legal play with same-step auto-reset in ONE launch: a warp keeps its
// 32 games in registers for the lock-steps of a work unit and streams each step's observation / mask / reward /
// terminated / action into [steps][n][...] rollout buffers.  Bit-identical to `steps` calls of spl_step chained
// through next_action, whatever the schedule.  p.scratch: [0] = unit counter, [4 + g] = lock-steps completed by
// group g (zeroed by the launcher).
template <int WPC>
__global__ void __launch_bounds__(WPC * 32, 20 / WPC) spl_rollout_kernel(const StepParams p) {
	__shared__ SplTables Ts;
	__shared__ __align__(16) uint32_t tiles[WPC][SPL_TILE_WORDS];
	__shared__ uint32_t s_unit;
	SplTile tl;
	tl.T = spl_stage_tables<WPC * 32>(&Ts);
	tl.lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	tl.smem = tiles[warp];
	const int64_t ntiles = (p.n + 31) >> 5;
	const uint32_t groups = (uint32_t)((ntiles + WPC - 1) / WPC);
	const uint32_t total = groups * (uint32_t)p.nchunks;
	const uint64_t t0 = p.action_t + (p.action_t_base ? *p.action_t_base : 0ull);
	uint32_t* const queue = reinterpret_cast<uint32_t*>(p.scratch);
	uint32_t* const done = queue + 4;

	for (;;) {
		if (threadIdx.x == 0) s_unit = atomicAdd(queue, 1u);
		__syncthreads();
		const uint32_t u = s_unit;
		if (u >= total) break;  // CTA-uniform
		const uint32_t c = u / groups, tg = u - c * groups;
		int start, len;
		spl_chunk_bounds((int)c, p.steps, p.chunk, start, len);
		if (start > 0) {  // the group's previous chunk must have published its state (almost always long ago)
			if (threadIdx.x == 0)
				while (spl_ld_acquire(done + tg) < (uint32_t)start) __nanosleep(100);
			__syncthreads();
		}
		tl.ti = (int64_t)tg * WPC + warp;
		const int64_t env = tl.ti * 32 + tl.lane;
		const bool valid = env < p.n;
		tl.rows = (int)max((int64_t)0, min((int64_t)32, p.n - tl.ti * 32));
		uint32_t w[16];
		if (valid) {  // L2 loads: the rows may have been written by another SM a moment ago
#pragma unroll
			for (int pl = 0; pl < SPL_STATE_PLANES; pl++) {
				uint4 v = __ldcg(p.state + pl * p.stride + env);
				w[4 * pl + 0] = v.x, w[4 * pl + 1] = v.y, w[4 * pl + 2] = v.z, w[4 * pl + 3] = v.w;
			}
		} else {
#pragma unroll
			for (int k = 0; k < 16; k++) w[k] = 0;
		}
		SplState s;
		spl_unpack(w, s);
		int32_t action = 0;
		if (valid) action = start == 0 ? p.actions[env] : __ldcg(p.next_action + (int64_t)start * p.n + env);
		uint32_t tops = valid ? spl_deck_tops(s, p.decks + env * SPL_DECK_STRIDE) : 0xFFFFFFu;
		// rotated loop: [legal mask of the current state] -> [emit the outputs of the previous step] -> [step].
		// The mask is needed twice -- as the action mask returned by step t-1 and as the legality check of
		// step t (envs/splendor_env.py:55,64,81) -- and is computed once; one code copy keeps the hot loop small.
		for (int st = start; st <= start + len; st++) {
			if (p.sync) __syncthreads();
			const uint64_t m = spl_is_terminal(s) ? 0ull : spl_legal_mask(s, tl.T);
			if (st > start) {
				const int64_t o = (int64_t)(st - 1) * p.n;
				action = spl_tile_emit(p, tl, s, w, env, valid, p.obs ? p.obs + o * SPL_OBS_DIM : nullptr,
				                       p.mask ? p.mask + o * SPL_NUM_ACTIONS : nullptr, p.next_action + o + p.n, t0 + (uint64_t)(st - 1), m);
			}
			if (st == start + len) break;
			const int64_t o = (int64_t)st * p.n;
			SplStepResult r;
			spl_tile_step<true>(p, tl, s, valid, action, env, r, m, &tops);
			spl_pack(s, w);
			if (valid) {
				p.reward[o + env] = r.reward;
				p.terminated[o + env] = (uint8_t)r.terminated;
				if (p.info != nullptr) p.info[o + env] = (uint8_t)r.info;
			}
		}
#ifdef SPL_DEBUG_SM_UNITS
		if (tl.lane == 0) {
			unsigned sm;
			asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
			atomicAdd(&g_sm_units[sm], (unsigned)len);
		}
#endif
		if (valid) spl_store_state(p, env, w);
		__syncthreads();  // all state / next_action stores of the CTA are ordered before the publication below
		if (threadIdx.x == 0 && start + len < p.steps) {
			__threadfence();
			spl_st_release(done + tg, (uint32_t)(start + len));
		}
	}
}

// ------------------------------------------------------------------------------------------------
// reset kernel: initial_state (engine/state.py:181-211) for a list of environments.
// One warp per CTA, one environment per lane; the deck order is built in shared memory.
//   SPL_SHUFFLE_MT19937: CPython random.Random(seed): init_by_array + shuffle by _randbelow, bit-exact (one env per lane)
//   SPL_SHUFFLE_PHILOX : the warp-cooperative sort-by-key deal (spl_coop_deal), one env at a time per warp
// ------------------------------------------------------------------------------------------------
struct ResetParams {
	uint4* state;
	int64_t stride;
	uint8_t* decks;
	uint32_t* episode;
	const int32_t* list;  // list[0] = count, list[4..] = env indices; nullptr => all envs 0..n-1
	int64_t n;
	uint64_t env_offset, seed_base;
	const uint64_t* ep_seeds;  // spl_envs_t.episode_seeds [n][ep_seed_count] or nullptr
	int ep_seed_count;
	const uint64_t* seeds;  // explicit engine seeds [n] or nullptr
	int32_t* obs;
	int8_t* mask;
	int bump_episode;  // 0: episode <- 0 (full reset), 1: episode += 1
	int32_t* next_action;  // nullable: fused random-legal sampler for the fresh state
	uint64_t action_key, action_t;
	const uint64_t* action_t_base;
	uint8_t* spare_out;  // non-null: do NOT reset anything; deal upcoming episodes of the listed envs into their spare rows
	int spare_slots;     // ring slots per env; an item is a code env * slots + slot (refill list / all codes) ...
	int list_is_envs;    // ... or, for a list of env ids (masked reset), item / slots indexes the list and item % slots is the slot
	int max_outputs;     // batch dealer: generator outputs a deal may use before it is left to the consumer (227; tests lower it)
	int32_t* refill;     // with spare_out: the list is the refill list (codes); the last CTA to finish empties it
};

// MT19937 state of one lane, lane-interleaved in shared memory (conflict-free).  The generator sits on the critical
// path of every bit-exact auto-reset, and one lane cannot go faster than its dependent chain, so that chain is kept as
// short as it can be:
//   * random.Random(a) = init_genrand(19650218) + init_by_array(key): the init_genrand recurrence is carried in a
//     register next to the first init_by_array pass (624 fused iterations, one store each, no loads); the second pass
//     carries mt[i-1] in a register and only loads the old mt[i], whose address does not depend on the chain;
//   * the first 227 outputs depend on OLD state words only (k, k+1, k+397), so they are produced on demand without the
//     624-word regeneration; a deal consumes ~130 outputs.  Output 228 (never seen in 1e6 games, but reachable through
//     rejection sampling) triggers the regular in-place regeneration of the untouched state.
#ifndef SPL_MT_LAZY
#define SPL_MT_LAZY 227 /* a smaller value is still exact (tests build one to exercise the regeneration path) */
#endif
template <int STRIDE = 32>
struct SplMT {
	uint32_t* mt;
	int idx;
	bool twisted;
	__device__ __forceinline__ uint32_t& at(int i) { return mt[i * STRIDE]; }
	__device__ void seed(uint64_t a) {  // random.Random(a): init_by_array over the 32-bit words of a
		const uint32_t key0 = (uint32_t)a, key1 = (uint32_t)(a >> 32);
		const bool two = key1 != 0;
		uint32_t g = 19650218u, q = g, m1_1 = 0, j = 0;
#pragma unroll 4
		for (int i = 1; i < 624; i++) {  // init_genrand (g) and iterations 1..623 of the first init_by_array loop (q)
			g = 1812433253u * (g ^ (g >> 30)) + (uint32_t)i;
			q = (g ^ ((q ^ (q >> 30)) * 1664525u)) + (j ? key1 : key0) + j;
			at(i) = q;
			if (i == 1) m1_1 = q;
			j = two ? (j ^ 1u) : 0u;
		}
		// 624th iteration: the index wrapped (mt[0] <- mt[623] = q) and mt[1] is updated once more
		uint32_t p = (m1_1 ^ ((q ^ (q >> 30)) * 1664525u)) + (j ? key1 : key0) + j;
		const uint32_t m1p_1 = p;
#pragma unroll 4
		for (int i = 2; i < 624; i++) {  // second loop, iterations 1..622
			p = (at(i) ^ ((p ^ (p >> 30)) * 1566083941u)) - (uint32_t)i;
			at(i) = p;
		}
		// 623rd iteration after the wrap (mt[0] <- mt[623] = p): mt[1]; then mt[0] = 0x80000000
		at(1) = (m1p_1 ^ ((p ^ (p >> 30)) * 1566083941u)) - 1u;
		at(0) = 0x80000000u;
		idx = 0;
		twisted = false;
	}
	__device__ void twist() {
		for (int kk = 0; kk < 624; kk++) {
			uint32_t y = (at(kk) & 0x80000000u) | (at((kk + 1) % 624) & 0x7fffffffu);
			at(kk) = at((kk + 397) % 624) ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
		}
	}
	__device__ uint32_t next() {
		uint32_t y;
		if (!twisted && idx < SPL_MT_LAZY) {
			const uint32_t u = (at(idx) & 0x80000000u) | (at(idx + 1) & 0x7fffffffu);
			y = at(idx + 397) ^ (u >> 1) ^ ((u & 1u) ? 0x9908b0dfu : 0u);
			idx++;
		} else {
			if (!twisted || idx >= 624) {
				twist();
				if (twisted) idx = 0;
				twisted = true;
			}
			y = at(idx++);
		}
		y ^= y >> 11;
		y ^= (y << 7) & 0x9d2c5680u;
		y ^= (y << 15) & 0xefc60000u;
		y ^= y >> 18;
		return y;
	}
	__device__ uint32_t randbelow(uint32_t n) {  // Lib/random.py _randbelow_with_getrandbits
		int k = 32 - __clz(n);
		uint32_t r = next() >> (32 - k);
		while (r >= n) r = next() >> (32 - k);
		return r;
	}
};

template <class Rng>
__device__ __forceinline__ void spl_shuffle_bytes(Rng& rng, uint8_t* x, int len) {  // Lib/random.py shuffle
	for (int i = len - 1; i >= 1; i--) {
		uint32_t j = rng.randbelow((uint32_t)(i + 1));
		uint8_t t = x[i];
		x[i] = x[j];
		x[j] = t;
	}
}

template <class Rng>
__device__ __forceinline__ void spl_deal(Rng& rng, uint8_t* deck, SplState& s) {
	spl_fresh_state(s);
	const int off[3] = {0, 40, 70}, len[3] = {40, 30, 20};
	for (int k = 0; k < SPL_DECK_STRIDE; k++) deck[k] = k < 90 ? (uint8_t)k : (uint8_t)SPL_EMPTY;
	uint32_t dn = 0;
	for (int t = 0; t < 3; t++) {  // engine/state.py:188-191
		spl_shuffle_bytes(rng, deck + off[t], len[t]);
		uint32_t b = 0;
		for (int i = 0; i < 4; i++) b |= (uint32_t)deck[off[t] + len[t] - 1 - i] << (8 * i);  // pop() from the end
		s.board[t] = b;
		dn |= (uint32_t)(len[t] - 4) << (8 * t);
	}
	s.deckn = dn;
	uint8_t nob[10];
	for (int i = 0; i < 10; i++) nob[i] = (uint8_t)i;
	spl_shuffle_bytes(rng, nob, 10);  // engine/state.py:193-195
	s.nobles = (uint32_t)nob[0] | ((uint32_t)nob[1] << 8) | ((uint32_t)nob[2] << 16);
}

__device__ __noinline__ void spl_mt_deal_in_place(uint64_t seed, uint32_t* smem, uint8_t* deck_row) {
	SplMT<1> rng;
	rng.mt = smem;
	rng.seed(seed);
	uint8_t* deck = reinterpret_cast<uint8_t*>(smem + 624);
	SplState s;
	spl_deal(rng, deck, s);
	const uint32_t* d4 = reinterpret_cast<const uint32_t*>(deck);
	uint32_t* g4 = reinterpret_cast<uint32_t*>(deck_row);
	for (int k = 0; k < SPL_DECK_STRIDE / 4; k++) g4[k] = d4[k];
	smem[648] = s.board[0], smem[649] = s.board[1], smem[650] = s.board[2], smem[651] = s.nobles;
	smem[652] = (uint32_t)deck[35] | ((uint32_t)deck[65] << 8) | ((uint32_t)deck[85] << 16);
}

template <int SHUFFLE>
__global__ void __launch_bounds__(32) spl_reset_kernel(const ResetParams p) {
	extern __shared__ __align__(16) uint8_t smem[];
	SplTables* T = reinterpret_cast<SplTables*>(smem);
	uint32_t* tile = reinterpret_cast<uint32_t*>(smem + sizeof(SplTables));
	uint8_t* decks_s = smem + sizeof(SplTables) + SPL_TILE_WORDS * 4;
	uint32_t* mt_s = reinterpret_cast<uint32_t*>(decks_s + 32 * SPL_DECK_SMEM);
	const int lane = threadIdx.x;
	const bool spare_mode = SHUFFLE == SPL_SHUFFLE_MT19937 && p.spare_out != nullptr;
	const int64_t R = spare_mode ? p.spare_slots : 1;
	int64_t count = p.list ? (int64_t)p.list[0] * (p.list_is_envs ? R : 1) : p.n * R;
	if (count > p.n * R) count = p.n * R;  // capacity of the lists
	if (count == 0) return;
	spl_stage_tables<32>(T);
	// A work list (auto-reset of the envs that just finished) is short and sits on the critical path of the lock-step:
	// spread it over every warp of the grid -- `ipw` items per warp instead of 32 -- so that the serial parts (one
	// lane's MT19937 chain, one scattered output row per iteration) run side by side on all SMs.
	int ipw = 32;
	if (p.list != nullptr) {
		const int64_t per = (count + gridDim.x - 1) / gridDim.x;
		ipw = per < 1 ? 1 : (per > 32 ? 32 : (int)per);
	}
	for (int64_t g = blockIdx.x; g * ipw < count; g += gridDim.x) {
		const int64_t item = g * ipw + lane;
		const bool valid = lane < ipw && item < count;
		int64_t env = 0, slot = 0;
		if (valid) {
			if (!spare_mode) env = p.list ? (int64_t)p.list[4 + item] : item;
			else if (p.list != nullptr && p.list_is_envs) env = (int64_t)p.list[4 + item / R], slot = item % R;
			else {
				const int64_t code = p.list ? (int64_t)p.list[4 + item] : item;
				env = code / R, slot = code - env * R;
			}
		}
		uint8_t* deck = decks_s + lane * SPL_DECK_SMEM;
		SplState s;
		uint32_t w[16];
		uint32_t ep = 0;
		uint64_t seed = 0;
		if (valid) {
			if (spare_mode) {
				// the first episode AFTER the current one that lives in this slot; the counter itself moves when a spare is taken
				const uint32_t cur = p.episode[env];
				ep = cur + 1u + (uint32_t)(((uint32_t)slot + (uint32_t)R - (cur + 1u) % (uint32_t)R) % (uint32_t)R);
			} else {
				ep = p.bump_episode ? p.episode[env] + 1u : 0u;
				p.episode[env] = ep;
			}
			uint64_t genv = p.env_offset + (uint64_t)env;
			seed = (p.seeds && !spare_mode) ? p.seeds[env] : spl_episode_seed(p.seed_base, p.env_offset, env, ep, p.ep_seeds, p.ep_seed_count);
			(void)genv;
		}
		spl_fresh_state(s);
		if (SHUFFLE == SPL_SHUFFLE_MT19937) {
			if (valid) {
				SplMT<32> rng;
				rng.mt = mt_s + lane;
				rng.seed(seed);
				spl_deal(rng, deck, s);
				if (spare_mode) {  // bytes 90..92: visible nobles, 93..94: episode tag, 95: ready
					deck[90] = (uint8_t)(s.nobles & 0xFFu), deck[91] = (uint8_t)((s.nobles >> 8) & 0xFFu), deck[92] = (uint8_t)((s.nobles >> 16) & 0xFFu);
					deck[93] = (uint8_t)(ep & 0xFFu), deck[94] = (uint8_t)((ep >> 8) & 0xFFu), deck[95] = 1;
				}
				const uint32_t* d4 = reinterpret_cast<const uint32_t*>(deck);
				uint32_t* g4 = reinterpret_cast<uint32_t*>(spare_mode ? p.spare_out + (env * R + slot) * SPL_DECK_STRIDE : p.decks + env * SPL_DECK_STRIDE);
#pragma unroll
				for (int k = 0; k < SPL_DECK_STRIDE / 4; k++) g4[k] = d4[k];
			}
			if (spare_mode) continue;  // nothing else changes
		} else {
			// same warp-cooperative deal as the fused auto-reset of the step kernel, one environment at a time
			const uint32_t vb = __ballot_sync(SPL_FULL, valid);
			for (uint32_t rb = vb; rb;) {
				const int src = __ffs(rb) - 1;
				rb &= rb - 1;
				const int64_t e = __shfl_sync(SPL_FULL, env, src);
				const uint32_t epr = __shfl_sync(SPL_FULL, ep, src);
				const uint64_t key = p.seeds ? __shfl_sync(SPL_FULL, seed, src) : p.seed_base;
				uint32_t board[3], nobles, tops_unused;
				spl_coop_deal(key, p.env_offset + (uint64_t)e, epr, tile, lane, p.decks + e * SPL_DECK_STRIDE, board, nobles, tops_unused);
				if (lane == src) spl_state_from_deal(s, board, nobles);
			}
		}
		spl_pack(s, w);
		if (valid) {
#pragma unroll
			for (int pl = 0; pl < SPL_STATE_PLANES; pl++)
				p.state[pl * p.stride + env] = make_uint4(w[4 * pl + 0], w[4 * pl + 1], w[4 * pl + 2], w[4 * pl + 3]);
		}
		if (p.mask != nullptr || p.next_action != nullptr) {  // rows are scattered: one row per warp iteration, 45 coalesced bytes
			uint64_t m = spl_legal_mask(s, T);
			if (p.next_action != nullptr && valid) {
				uint64_t t = p.action_t + (p.action_t_base ? *p.action_t_base : 0ull);
				p.next_action[env] = spl_sample_action(m, p.action_key, p.env_offset + (uint64_t)env, t);
			}
			if (p.mask != nullptr)
			for (int r = 0; r < ipw; r++) {
				int64_t e = __shfl_sync(SPL_FULL, env, r);
				uint64_t mr = __shfl_sync(SPL_FULL, m, r);
				if (g * ipw + r < count) {
					p.mask[e * SPL_NUM_ACTIONS + lane] = (int8_t)((mr >> lane) & 1);
					if (lane < SPL_NUM_ACTIONS - 32) p.mask[e * SPL_NUM_ACTIONS + 32 + lane] = (int8_t)((mr >> (32 + lane)) & 1);
				}
			}
		}
		if (p.obs != nullptr) {
			SplObsStager stage(tile, lane, w[0]);
			spl_encode_observation(w, s, T, stage);
			__syncwarp();
			const uint8_t* tb = reinterpret_cast<const uint8_t*>(tile);
			for (int r = 0; r < ipw; r++) {
				int64_t e = __shfl_sync(SPL_FULL, env, r);
				if (g * ipw + r < count) {
#pragma unroll
					for (int c = lane; c < SPL_OBS_DIM; c += 32) p.obs[e * SPL_OBS_DIM + c] = (int32_t)tb[SPL_OBS_DIM * r + c];
				}
			}
			__syncwarp();
		}
	}
	if (p.refill != nullptr) {  // every CTA read the header before it got here: the last one to arrive empties the list
		__syncwarp();
		if (lane == 0) {
			__threadfence();
			if (atomicAdd(reinterpret_cast<unsigned int*>(p.refill + 2), 1u) == gridDim.x - 1u) {
				p.refill[0] = 0, p.refill[1] = 0, p.refill[2] = 0;
			}
		}
	}
}

// Refill scan: which (env, slot) rings entries need a deal?  Those whose ready flag is clear (taken since the last
// refill, or left undone by the dealer) and those whose tag is not the episode the slot is next needed for (an env that
// ran its ring dry was dealt in place and has moved past them).  Compacted into the list behind the spare rows; consumers
// never touch that list, so a scan + deal may run while step kernels take deals from other slots.
__global__ void __launch_bounds__(256) spl_spare_scan_kernel(const uint8_t* spare, const uint32_t* episode, int64_t n, int slots, int32_t* list) {
	const int64_t code = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool need = false;
	if (code < n * slots) {
		const int64_t env = code / slots;
		const uint32_t slot = (uint32_t)(code - env * slots), R = (uint32_t)slots;
		const uint32_t w23 = __ldcg(reinterpret_cast<const uint32_t*>(spare + code * SPL_DECK_STRIDE) + 23);
		const uint32_t cur = __ldcg(episode + env);
		const uint32_t ep = cur + 1u + (slot + R - (cur + 1u) % R) % R;  // the episode this slot is next needed for
		need = (w23 >> 24) == 0u || ((w23 >> 8) & 0xFFFFu) != (ep & 0xFFFFu);
	}
	const uint32_t b = __ballot_sync(SPL_FULL, need);
	if (b) {
		const int lane = threadIdx.x & 31;
		int base = 0;
		if (lane == 0) base = atomicAdd(list, __popc(b));
		base = __shfl_sync(SPL_FULL, base, 0);
		if (need) list[4 + base + __popc(b & ((1u << lane) - 1u))] = (int32_t)code;
	}
}

// ------------------------------------------------------------------------------------------------
// Batch dealer for the prefetched deals (spl_envs_t.spare): initial_state(seed)'s shuffles for MANY (env, slot) items at
// once, one item per thread at full occupancy, with the MT19937 generator entirely in REGISTERS.
//   * random.Random(seed) for a one-word key is two passes over the 624 state words (init_by_array); pass 1 needs no
//     memory at all (its inputs are the seed-independent init_genrand(19650218) words, a 2.5 KB table), so pass 2 can
//     re-run pass 1 next to itself instead of reading stored values.
//   * a deal consumes ~130 outputs, in order, and output k of the first generation is a function of the final words
//     k, k+1 and k+397 only.  After pass 2 has run once (word 1 is final only at its very end), two more copies of the
//     pass-2 chain are advanced on demand, one from word 2 and one from word 398 -- every output is produced from
//     registers, nothing of the state is ever stored.  227 outputs are available that way; a deal that needs more
//     (never observed in 1e7 games, but rejection sampling can) is left undone: its slot stays "not ready" and the
//     consumer deals that episode itself (work list / in-place deal).
//   * the four shuffles run as ONE flat loop per thread (a state machine over (deck, position)): a warp iterates
//     max-over-lanes of the draws of a whole deal (~155) instead of the sum of per-position maxima (~350) that
//     32 diverging _randbelow rejection loops would cost.
// spl_reset_kernel<MT19937> keeps its 624 words per lane in shared memory (2 warps per SM, ~35 us per chain); here
// ~108,000 deals (what a 65,536-env x 128-step rollout segment consumes) take ~0.1 ms.
// ------------------------------------------------------------------------------------------------
#define SPL_DEAL_THREADS 128
__device__ uint32_t g_mt_init[624];  // init_genrand(19650218): mt[0..623] before init_by_array (spl_init)

__global__ void __launch_bounds__(SPL_DEAL_THREADS) spl_spare_deal_kernel(const ResetParams p) {
	__shared__ uint32_t G[624];
	__shared__ uint8_t decks_s[SPL_DEAL_THREADS * SPL_DECK_SMEM];
	const int64_t R = p.spare_slots;
	int64_t count = p.list ? (int64_t)p.list[0] * (p.list_is_envs ? R : 1) : p.n * R;
	if (count > p.n * R) count = p.n * R;  // capacity of the refill list
	if (count == 0) return;
	const int64_t tid = (int64_t)blockIdx.x * SPL_DEAL_THREADS + threadIdx.x;
	if ((int64_t)blockIdx.x * SPL_DEAL_THREADS < count) {
		for (int k = threadIdx.x; k < 624; k += SPL_DEAL_THREADS) G[k] = g_mt_init[k];
		__syncthreads();
	}
	uint8_t* deck = decks_s + threadIdx.x * SPL_DECK_SMEM;
	for (int64_t item = tid; item < count; item += (int64_t)gridDim.x * SPL_DEAL_THREADS) {
		int64_t env, slot;
		if (p.list != nullptr && p.list_is_envs) env = (int64_t)p.list[4 + item / R], slot = item % R;
		else {
			const int64_t code = p.list ? (int64_t)p.list[4 + item] : item;
			env = code / R, slot = code - env * R;
		}
		// the first episode AFTER the current one that lives in this slot; the counter itself moves when a spare is taken
		const uint32_t cur = p.episode[env];
		const uint32_t ep = cur + 1u + (uint32_t)(((uint32_t)slot + (uint32_t)R - (cur + 1u) % (uint32_t)R) % (uint32_t)R);
		const uint64_t key64 = spl_episode_seed(p.seed_base, p.env_offset, env, ep, p.ep_seeds, p.ep_seed_count);
		if (key64 >> 32) continue;  // a table seed wider than one word: left to the consumer's in-place deal (full init_by_array)
		const uint32_t key = (uint32_t)key64;  // one-word key
		const bool overflow = !spl_mt_deal_stream(key, G, deck, (uint32_t)p.max_outputs);
		if (overflow) continue;
		// bytes 90..92: the three visible nobles (already there), 93..94: episode tag, 95: ready
		deck[93] = (uint8_t)(ep & 0xFFu), deck[94] = (uint8_t)((ep >> 8) & 0xFFu), deck[95] = 1;
		const uint32_t* d4 = reinterpret_cast<const uint32_t*>(deck);
		uint4* g4 = reinterpret_cast<uint4*>(p.spare_out + (env * R + slot) * SPL_DECK_STRIDE);
#pragma unroll
		for (int k = 0; k < SPL_DECK_STRIDE / 16 - 1; k++) g4[k] = make_uint4(d4[4 * k], d4[4 * k + 1], d4[4 * k + 2], d4[4 * k + 3]);
		__threadfence();  // the last 16 bytes carry the tag and the ready flag: published after the deck order
		g4[5] = make_uint4(d4[20], d4[21], d4[22], d4[23]);
	}
	if (p.refill != nullptr) {  // every CTA read the header before it got here: the last one to arrive empties the list
		__syncthreads();
		if (threadIdx.x == 0) {
			__threadfence();
			if (atomicAdd(reinterpret_cast<unsigned int*>(p.refill + 2), 1u) == gridDim.x - 1u) {
				p.refill[0] = 0, p.refill[1] = 0, p.refill[2] = 0;
			}
		}
	}
}

// reset_mask[n] -> work list (count in list[0], indices from list[4])
__global__ void spl_compact_kernel(const uint8_t* __restrict__ flags, int64_t n, int32_t* list) {
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool f = i < n && flags[i] != 0;
	uint32_t b = __ballot_sync(SPL_FULL, f);
	if (b) {
		int lane = threadIdx.x & 31, base = 0;
		if (lane == 0) base = atomicAdd(list, __popc(b));
		base = __shfl_sync(SPL_FULL, base, 0);
		if (f) list[4 + base + __popc(b & ((1u << lane) - 1u))] = (int32_t)i;
	}
}

// ------------------------------------------------------------------------------------------------
// small kernels: random legal action from an int8 mask, export / import, dual-step bookkeeping
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) spl_random_action_kernel(const int8_t* __restrict__ mask, int64_t n, uint64_t env_offset,
                                                              uint64_t key, uint64_t t, int32_t* __restrict__ actions) {
	__shared__ uint8_t rows[4][32 * SPL_NUM_ACTIONS];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t ntiles = (n + 31) >> 5;
	for (int64_t ti = (int64_t)blockIdx.x * 4 + warp; ti < ntiles; ti += (int64_t)gridDim.x * 4) {
		const int rowsv = (int)min((int64_t)32, n - ti * 32);
		const int8_t* g = mask + ti * 32 * SPL_NUM_ACTIONS;
		for (int b = lane; b < rowsv * SPL_NUM_ACTIONS; b += 32) rows[warp][b] = (uint8_t)g[b];
		__syncwarp();
		if (lane < rowsv) {
			uint64_t m = 0;
			for (int a = 0; a < SPL_NUM_ACTIONS; a++) m |= (uint64_t)(rows[warp][lane * SPL_NUM_ACTIONS + a] != 0) << a;
			int64_t env = ti * 32 + lane;
			actions[env] = spl_sample_action(m, key, env_offset + (uint64_t)env, t);
		}
		__syncwarp();
	}
}

__global__ void spl_export_kernel(const uint4* state, int64_t stride, const uint8_t* decks, int64_t n, int32_t* rows) {
	int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (env >= n) return;
	uint32_t w[16];
	for (int pl = 0; pl < SPL_STATE_PLANES; pl++) {
		uint4 v = state[pl * stride + env];
		w[4 * pl + 0] = v.x, w[4 * pl + 1] = v.y, w[4 * pl + 2] = v.z, w[4 * pl + 3] = v.w;
	}
	SplState s;
	spl_unpack(w, s);
	spl_export_row(s, decks + env * SPL_DECK_STRIDE, rows + env * SPL_ROW_LEN);
}

// A row outside the packed state's domain (spl_row_valid) is not imported: its env keeps its state and *bad is bumped.
__global__ void spl_import_kernel(uint4* state, int64_t stride, uint8_t* decks, int64_t n, const int32_t* rows, const uint8_t* which, int* bad) {
	int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (env >= n || (which && !which[env])) return;
	if (!spl_row_valid(rows + env * SPL_ROW_LEN)) {
		atomicAdd(bad, 1);
		return;
	}
	SplState s;
	uint32_t w[16];
	spl_import_row(rows + env * SPL_ROW_LEN, s, decks + env * SPL_DECK_STRIDE);
	spl_pack(s, w);
	for (int pl = 0; pl < SPL_STATE_PLANES; pl++)
		state[pl * stride + env] = make_uint4(w[4 * pl + 0], w[4 * pl + 1], w[4 * pl + 2], w[4 * pl + 3]);
}

// spl_load_deals: caller-supplied deals -> ring slots.  One thread per (env, k); validation first (all rows), then the copy.
__global__ void spl_check_deals_kernel(const uint8_t* __restrict__ deals, int64_t items, int* bad) {
	const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (it >= items) return;
	const uint8_t* d = deals + it * SPL_DECK_STRIDE;
	const int lo[4] = {0, 40, 70, 90};
	bool ok = true;
	for (int t = 0; t < 3; t++) {  // each tier: a permutation of its own card ids
		uint64_t seen = 0;
		for (int k = lo[t]; k < lo[t + 1]; k++) {
			const int id = (int)d[k] - lo[t];
			ok = ok && id >= 0 && id < lo[t + 1] - lo[t];
			if (ok) seen |= 1ull << id;
		}
		ok = ok && seen == ((1ull << (lo[t + 1] - lo[t])) - 1ull);
	}
	ok = ok && d[90] < 10 && d[91] < 10 && d[92] < 10 && d[90] != d[91] && d[90] != d[92] && d[91] != d[92];
	if (!ok) atomicAdd(bad, 1);
}

__global__ void spl_load_deals_kernel(const uint8_t* __restrict__ deals, int64_t n, int count, const uint32_t* __restrict__ episode,
                                      uint8_t* spare, int slots, const int* bad) {
	const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (it >= n * count || *bad != 0) return;
	const int64_t env = it / count;
	const uint32_t ep = episode[env] + 1u + (uint32_t)(it - env * count);
	const uint32_t* src = reinterpret_cast<const uint32_t*>(deals + it * SPL_DECK_STRIDE);
	uint32_t* dst = reinterpret_cast<uint32_t*>(spare + (env * slots + (int64_t)(ep % (uint32_t)slots)) * SPL_DECK_STRIDE);
	for (int k = 0; k < 23; k++) dst[k] = src[k];
	__threadfence();  // tag + ready flag last, as every other writer of the ring does
	dst[23] = (src[23] & 0xFFu) | ((ep & 0xFFFFu) << 8) | 0x01000000u;
}

// final_rewards[player] (envs/splendor_env.py:92-115) from an info byte; 0 when absent
__device__ __forceinline__ float spl_final_reward(uint32_t info, uint32_t player) {
	if (!(info & SPL_INFO_TERMINATED) || (info & (SPL_INFO_NOLEGAL_DRAW | SPL_INFO_ERROR))) return 0.0f;
	uint32_t wcode = (info & SPL_INFO_WINNER_MASK) >> SPL_INFO_WINNER_SHIFT;
	if (wcode == 0) return (info & SPL_INFO_TURN_LIMIT) ? -0.1f : 0.0f;
	return (wcode - 1 == player) ? 1.0f : -1.0f;
}

// wrappers/dual_step_native.py:132-167: phase 1 = agent (player 0), phase 2 = opponent (player 1)
// mode 0: DualStepNativeWrapper / DualStepSelfPlayWrapper (agent reward = final_rewards[0]);
// mode 1: SelfPlayWrapper (agent reward = -opponent reward when the opponent's move ends the game, selfplay.py:55-57)
__global__ void spl_dual_combine_kernel(const float* r1, const uint8_t* t1, const uint8_t* i1, const float* r2, const uint8_t* t2,
                                        const uint8_t* i2, int64_t n, float* agent_reward, float* opp_reward, uint8_t* done, int mode) {
	int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n) return;
	if (t1[e]) {  // game ended on the agent's move (:132-145)
		agent_reward[e] = r1[e];
		opp_reward[e] = spl_final_reward(i1[e], 1);
		done[e] = 1;
	} else if (i1[e] & (SPL_INFO_ILLEGAL | SPL_INFO_ERROR)) {  // agent move rejected: opponent does not move
		agent_reward[e] = r1[e];
		opp_reward[e] = 0.0f;
		done[e] = 0;
	} else {  // (:157-167)
		agent_reward[e] = t2[e] ? (mode == 1 ? -r2[e] : spl_final_reward(i2[e], 0)) : 0.0f;
		opp_reward[e] = r2[e];
		done[e] = t2[e];
	}
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static bool g_inited[64];  // per device ordinal: constant tables uploaded, kernel attributes set
static bool device_ready() {
	int dev = -1;
	return cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && g_inited[dev];
}
static int g_num_sms = 0;
static int g_occ[3][2];  // [kernel: step, observe, rollout][wpc 1 / 4] resident CTAs per SM
static uint64_t g_host_ret[SPL_RET_TABLE_LEN];
static bool g_host_ret_built = false;
int64_t g_launches = 0;  // also bumped by spl_policy.cu
// optional per-launch timing of the step kernel (bench.py roofline leg): CUDA events recorded on the
// caller's stream right around the kernel
#define SPL_TIMING_POOL 4096
static cudaEvent_t g_ev[2 * SPL_TIMING_POOL];
static int g_ev_created = 0, g_ev_used = 0, g_timing = 0;

#define SPL_MT_SMEM (sizeof(SplTables) + SPL_TILE_WORDS * 4 + 32 * SPL_DECK_SMEM + 624 * 32 * 4)
#define SPL_PHILOX_SMEM (sizeof(SplTables) + SPL_TILE_WORDS * 4 + 32 * SPL_DECK_SMEM)

#define SPL_CUDA(x)                      \
	do {                                 \
		cudaError_t e_ = (x);            \
		if (e_ != cudaSuccess) return (int)e_; \
	} while (0)

// Tuning knobs (environment variables, diagnostics only): read once per spl_init() call, not per launch.
enum { K_STEP_PAIRED, K_DEAL_BATCH, K_DEAL_MAX_OUTPUTS, K_ROLLOUT_CHUNK, K_ROLLOUT_CTAS_PER_SM, K_ROLLOUT_SYNC, K_SPARE_REFILL_AGE, K_SPARE_WORKLIST, K_STEP_PERSISTENT, K_STEP_WPC, K_WPC, K_COUNT };
static const char* const g_knob_names[K_COUNT] = {"SPL_STEP_PAIRED", "SPL_DEAL_BATCH", "SPL_DEAL_MAX_OUTPUTS", "SPL_ROLLOUT_CHUNK", "SPL_ROLLOUT_CTAS_PER_SM", "SPL_ROLLOUT_SYNC", "SPL_SPARE_REFILL_AGE", "SPL_SPARE_WORKLIST", "SPL_STEP_PERSISTENT", "SPL_STEP_WPC", "SPL_WPC"};
static int g_knobs[K_COUNT];
static bool g_knob_set[K_COUNT];
static void load_knobs() {
	for (int k = 0; k < K_COUNT; k++) {
		const char* e = getenv(g_knob_names[k]);
		g_knob_set[k] = e && *e;
		g_knobs[k] = g_knob_set[k] ? atoi(e) : 0;
	}
}
static inline int knob(int k, int dflt) { return g_knob_set[k] ? g_knobs[k] : dflt; }

extern "C" {

int spl_version(void) { return 120; }  // 120: spl_envs.episode_seeds, spl_load_deals, spl_host_alloc / push path, SPL_E_BADROW

int spl_timing_enable(int on) {
	if (on && !g_ev_created) {
		for (int i = 0; i < 2 * SPL_TIMING_POOL; i++) SPL_CUDA(cudaEventCreate(&g_ev[i]));
		g_ev_created = 1;
	}
	g_timing = on;
	g_ev_used = 0;
	return 0;
}

// synchronises the device; returns the summed duration and the number of timed step-kernel launches
int spl_timing_read(double* total_ms, int64_t* count) {
	SPL_CUDA(cudaDeviceSynchronize());
	double tot = 0;
	for (int i = 0; i < g_ev_used; i++) {
		float ms = 0;
		SPL_CUDA(cudaEventElapsedTime(&ms, g_ev[2 * i], g_ev[2 * i + 1]));
		tot += ms;
	}
	*total_ms = tot;
	*count = g_ev_used;
	g_ev_used = 0;
	return 0;
}

int64_t spl_launch_count(void) { return g_launches; }

#ifdef SPL_DEBUG_PHASES
int spl_debug_phases(unsigned long long* out, int warps) {  // out[SPL_PHASE_SLOTS][warps]
	SPL_CUDA(cudaDeviceSynchronize());
	for (int k = 0; k < SPL_PHASE_SLOTS; k++)
		SPL_CUDA(cudaMemcpyFromSymbol(out + (size_t)k * warps, g_phase, sizeof(unsigned long long) * warps, sizeof(unsigned long long) * SPL_PHASE_WARPS * k));
	return 0;
}
#endif


#ifdef SPL_DEBUG_SM_UNITS
int spl_debug_sm_units(unsigned int* out, int reset) {
	SPL_CUDA(cudaMemcpyFromSymbol(out, g_sm_units, sizeof(unsigned int) * 256));
	if (reset) {
		unsigned int z[256] = {0};
		SPL_CUDA(cudaMemcpyToSymbol(g_sm_units, z, sizeof(z)));
	}
	return 0;
}
#endif

const char* spl_error_string(int code) {
	if (code == 0) return "ok";
	if (code == SPL_E_BADARG) return "splendor_b200: bad argument";
	if (code == SPL_E_NOTINIT) return "splendor_b200: spl_init() has not been called on this device";
	if (code == SPL_E_ALIGN) return "splendor_b200: state pointer must be 16-byte aligned";
	if (code == SPL_E_BADROW) return "splendor_b200: state row outside the engine's domain (counter >= 128, id out of range or list too long); rejected rows were not imported";
	if (code > 0) return cudaGetErrorString((cudaError_t)code);
	return "splendor_b200: unknown error";
}

int spl_host_ret_table(uint64_t* out) {
	if (!g_host_ret_built) {
		spl_build_ret_table(g_host_ret);
		g_host_ret_built = true;
	}
	memcpy(out, g_host_ret, sizeof(g_host_ret));
	return 0;
}

int spl_init(void) {
	int dev = 0;
	SPL_CUDA(cudaGetDevice(&dev));
	load_knobs();
	if (dev < 0 || dev >= 64) return SPL_E_BADARG;
	if (g_inited[dev]) return 0;
	SplTables T;
	spl_build_tables(&T);
	if (!g_host_ret_built) {
		spl_build_ret_table(g_host_ret);
		g_host_ret_built = true;
	}
	SPL_CUDA(cudaMemcpyToSymbol(g_tables, &T, sizeof(T)));
	SPL_CUDA(cudaMemcpyToSymbol(g_ret_table, g_host_ret, sizeof(g_host_ret)));
	{
		uint32_t g[624];
		spl_mt_init_table(g);
		SPL_CUDA(cudaMemcpyToSymbol(g_mt_init, g, sizeof(g)));
	}
	SPL_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
	SPL_CUDA(cudaFuncSetAttribute(spl_reset_kernel<SPL_SHUFFLE_MT19937>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPL_MT_SMEM));
	SPL_CUDA(cudaFuncSetAttribute(spl_reset_kernel<SPL_SHUFFLE_PHILOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPL_PHILOX_SMEM));
#define SPL_SETUP(K, kid, slot, threads)                                                                     \
	SPL_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
	SPL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_occ[kid][slot], K, threads, 0));                       \
	if (g_occ[kid][slot] < 1) g_occ[kid][slot] = 1;
	SPL_SETUP((spl_step_kernel<true, 4, SPL_OUT_I32>), 0, 1, 128)
	SPL_SETUP((spl_step_kernel<false, 4, SPL_OUT_I32>), 1, 1, 128)
	SPL_SETUP((spl_step_kernel<true, 1, SPL_OUT_I32>), 0, 0, 32)
	SPL_SETUP((spl_step_kernel<false, 1, SPL_OUT_I32>), 1, 0, 32)
	SPL_CUDA(cudaFuncSetAttribute(spl_step_kernel<true, 1, SPL_OUT_COMPACT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	SPL_CUDA(cudaFuncSetAttribute(spl_step_kernel<false, 1, SPL_OUT_COMPACT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	SPL_CUDA(cudaFuncSetAttribute(spl_step_kernel<true, 1, SPL_OUT_F16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	SPL_CUDA(cudaFuncSetAttribute(spl_step_kernel<false, 1, SPL_OUT_F16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	SPL_CUDA(cudaFuncSetAttribute(spl_step_kernel<true, 4, SPL_OUT_COMPACT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	SPL_CUDA(cudaFuncSetAttribute(spl_step_kernel<false, 4, SPL_OUT_COMPACT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	SPL_CUDA(cudaFuncSetAttribute(spl_step_kernel<true, 4, SPL_OUT_F16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	SPL_CUDA(cudaFuncSetAttribute(spl_step_kernel<false, 4, SPL_OUT_F16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	SPL_CUDA(cudaFuncSetAttribute(spl_step_paired_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	SPL_CUDA(cudaFuncSetAttribute(spl_step_paired_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	SPL_SETUP((spl_rollout_kernel<1>), 2, 0, 32)
	SPL_SETUP((spl_rollout_kernel<4>), 2, 1, 128)
#undef SPL_SETUP
	SPL_CUDA(cudaDeviceSynchronize());
	g_inited[dev] = true;
	return 0;
}

static int check_envs(const spl_envs_t* e) {
	if (!e || !e->state || !e->decks || !e->episode || !e->scratch || e->n <= 0 || e->stride < e->n) return SPL_E_BADARG;
	if (((uintptr_t)e->state & 15) || ((uintptr_t)e->decks & 15) || ((uintptr_t)e->spare & 15)) return SPL_E_ALIGN;
	if (!device_ready()) return SPL_E_NOTINIT;
	return 0;
}


// what a reset launch does: reset the envs themselves, or (prefetched deals) only deal upcoming episodes into ring slots
#define SPL_RESET_NORMAL 0
#define SPL_RESET_SPARE_FILL 1 /* every slot of every env (after a full reset) */
#define SPL_RESET_SPARE_REFILL_NOW 3 /* the refill list (slots taken since the last refill) */
#define SPL_RESET_SPARE_FILL_ENVS 4  /* every slot of the envs in a list of env ids (masked reset) */
#define SPL_SPARE_REFILL_AGE 16 /* lock-steps per ring slot; a game lasts >= 17 moves, so a deal is back before its env can need it */

static int spare_slots(const spl_envs_t* e) { return e->spare_slots > 1 ? (e->spare_slots > SPL_MAX_SPARE_SLOTS ? SPL_MAX_SPARE_SLOTS : e->spare_slots) : 1; }
static int32_t* spare_list(const spl_envs_t* e) { return reinterpret_cast<int32_t*>(e->spare + e->n * spare_slots(e) * SPL_DECK_STRIDE); }

static int launch_reset(const spl_envs_t* e, const int32_t* list, const uint64_t* seeds, int bump, int32_t* obs, int8_t* mask,
                        cudaStream_t st, const spl_step_io_t* io = nullptr, int kind = SPL_RESET_NORMAL) {
	ResetParams p;
	p.next_action = io ? io->next_action : nullptr;
	p.action_key = io ? io->action_key : 0, p.action_t = io ? io->action_t : 0;
	p.action_t_base = io ? io->action_t_base : nullptr;
	p.state = (uint4*)e->state, p.stride = e->stride, p.decks = e->decks, p.episode = e->episode, p.list = list;
	p.n = e->n, p.env_offset = e->env_offset, p.seed_base = e->seed_base, p.seeds = seeds, p.obs = obs, p.mask = mask;
	p.ep_seeds = e->episode_seeds, p.ep_seed_count = e->episode_seeds ? e->episode_seed_count : 0;
	p.bump_episode = bump;
	p.spare_out = nullptr, p.refill = nullptr;
	p.spare_slots = 1, p.list_is_envs = 0, p.max_outputs = knob(K_DEAL_MAX_OUTPUTS, 227);
	int64_t groups = (e->n + 31) / 32;
	if (kind != SPL_RESET_NORMAL) {
		if (e->shuffle_mode != SPL_SHUFFLE_MT19937 || e->spare == nullptr) return SPL_E_BADARG;
		p.spare_out = e->spare, p.obs = nullptr, p.mask = nullptr, p.next_action = nullptr, p.seeds = nullptr;
		p.spare_slots = spare_slots(e);
		p.list_is_envs = kind == SPL_RESET_SPARE_FILL_ENVS;
		groups *= p.spare_slots;
		if (kind == SPL_RESET_SPARE_REFILL_NOW) {  // build the list of slots to deal: scan the ready flags / tags
			p.list = spare_list(e), p.refill = spare_list(e);
			SPL_CUDA(cudaMemsetAsync(spare_list(e), 0, 16, st));
			const int64_t codes = e->n * p.spare_slots;
			spl_spare_scan_kernel<<<(unsigned)((codes + 255) / 256), 256, 0, st>>>(e->spare, e->episode, e->n, p.spare_slots, spare_list(e));
			g_launches++;
			SPL_CUDA(cudaGetLastError());
		}
	}
	if (kind != SPL_RESET_NORMAL && knob(K_DEAL_BATCH, 1) != 0) {
		// items: every (env, slot) / every slot of the listed envs / the refill list (length known on the device only)
		int64_t ctas = (e->n * p.spare_slots + SPL_DEAL_THREADS - 1) / SPL_DEAL_THREADS;
		const int64_t cap = (int64_t)g_num_sms * 12;
		if (ctas > cap) ctas = cap;
		spl_spare_deal_kernel<<<(int)ctas, SPL_DEAL_THREADS, 0, st>>>(p);
	} else if (e->shuffle_mode == SPL_SHUFFLE_MT19937) {
		int grid = (int)(groups < (int64_t)g_num_sms * 2 ? groups : (int64_t)g_num_sms * 2);
		spl_reset_kernel<SPL_SHUFFLE_MT19937><<<grid, 32, SPL_MT_SMEM, st>>>(p);
	} else if (e->shuffle_mode == SPL_SHUFFLE_PHILOX) {
		int grid = (int)(groups < (int64_t)g_num_sms * 12 ? groups : (int64_t)g_num_sms * 12);
		spl_reset_kernel<SPL_SHUFFLE_PHILOX><<<grid, 32, SPL_PHILOX_SMEM, st>>>(p);
	} else {
		return SPL_E_BADARG;
	}
	g_launches++;
	return (int)cudaGetLastError();
}

int spl_reset(const spl_envs_t* envs, const uint64_t* seeds, const uint8_t* reset_mask, int32_t* obs, int8_t* mask, void* stream) {
	int rc = check_envs(envs);
	if (rc) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	const bool spares = envs->spare != nullptr && envs->shuffle_mode == SPL_SHUFFLE_MT19937;
	if (reset_mask == nullptr) {
		rc = launch_reset(envs, nullptr, seeds, 0, obs, mask, st);
		if (rc || !spares) return rc;
		SPL_CUDA(cudaMemsetAsync(spare_list(envs), 0, 16, st));  // nothing waits for a refill any more
		return launch_reset(envs, nullptr, nullptr, 0, nullptr, nullptr, st, nullptr, SPL_RESET_SPARE_FILL);
	}
	SPL_CUDA(cudaMemsetAsync(envs->scratch, 0, 16, st));
	spl_compact_kernel<<<(unsigned)((envs->n + 255) / 256), 256, 0, st>>>(reset_mask, envs->n, envs->scratch);
	g_launches++;
	SPL_CUDA(cudaGetLastError());
	rc = launch_reset(envs, envs->scratch, seeds, 1, obs, mask, st);
	if (rc || !spares) return rc;
	return launch_reset(envs, envs->scratch, nullptr, 0, nullptr, nullptr, st, nullptr, SPL_RESET_SPARE_FILL_ENVS);
}

static void fill_step_params(StepParams& p, const spl_envs_t* e, const spl_step_io_t* io, int32_t* obs, int8_t* mask) {
	p.state = (uint4*)e->state, p.stride = e->stride, p.decks = e->decks, p.episode = e->episode, p.scratch = e->scratch;
	p.n = e->n, p.env_offset = e->env_offset, p.seed_base = e->seed_base;
	p.ep_seeds = e->episode_seeds, p.ep_seed_count = e->episode_seeds ? e->episode_seed_count : 0;
	p.actions = io ? io->actions : nullptr, p.active = io ? io->active : nullptr;
	p.obs = obs, p.mask = mask;
	p.reward = io ? io->reward : nullptr, p.terminated = io ? io->terminated : nullptr, p.info = io ? io->info : nullptr;
	p.stats = io ? (unsigned long long*)io->stats : nullptr;
	p.next_action = io ? io->next_action : nullptr;
	p.action_key = io ? io->action_key : 0, p.action_t = io ? io->action_t : 0;
	p.action_t_base = io ? io->action_t_base : nullptr;
	p.reset_mode = SPL_RESET_NONE;
	p.spare = e->spare;
	p.spare_slots = spare_slots(e);
	p.spare_async = io && (io->flags & SPL_IO_ASYNC_REFILL) ? 1 : 0;
	if (io && io->autoreset)
		p.reset_mode = e->shuffle_mode == SPL_SHUFFLE_PHILOX ? SPL_RESET_FUSED
		             : (e->spare ? (knob(K_SPARE_WORKLIST, 0) ? SPL_RESET_SPARE : SPL_RESET_SPARE_INLINE) : SPL_RESET_WORKLIST);
	p.vec_ok = (((uintptr_t)obs | (uintptr_t)mask) & 15) == 0;  // 128-bit tile stores
	p.steps = 1;
	p.sync = 0;
	p.chunk = 1, p.nchunks = 1;
	p.obs_u8 = nullptr, p.side = nullptr, p.obs_f16 = nullptr;
}

// Launch shape.  Single-step kernels: one tile (1-warp CTA) per CTA, non-persistent (the 2 KB table staging is one
// 128-bit load per lane x 4).  Rollout kernel: persistent CTAs pulling (tile group, step chunk) units from the work queue;
// 1-warp CTAs when every tile can be resident at once (e.g. 65,536 envs = 2,048 tiles: one CTA per tile spreads
// them evenly and the balancing comes from tiles migrating between SMs from chunk to chunk), 4-warp CTAs at full
// occupancy otherwise.  Measured on B200 (tools/sweep_rollout.py): fewer CTAs than tile groups is slower (the
// fast SMs are latency-bound and need the warps), a CTA barrier per lock-step no longer pays once the queue
// balances the SMs, chunk lengths between 8 and 16 lock-steps are equivalent within 2 %.
struct LaunchShape {
	int wpc, grid, sync, chunk;
};

static LaunchShape launch_shape(int64_t n, int kernel) {
	const int64_t ntiles = (n + 31) / 32;
	LaunchShape L;
	int64_t resident4 = (int64_t)g_num_sms * g_occ[kernel][1] * 4;
	L.wpc = kernel != 2 ? knob(K_STEP_WPC, 1) : knob(K_WPC, ntiles <= resident4 ? 1 : 4);
	if (L.wpc != 1) L.wpc = 4;
	int64_t ctas = (ntiles + L.wpc - 1) / L.wpc;
	int64_t cap = (int64_t)g_num_sms * g_occ[kernel][L.wpc == 1 ? 0 : 1];
	L.grid = (int)(ctas < cap ? ctas : cap);
	if (kernel != 2 && knob(K_STEP_PERSISTENT, 0) == 0) L.grid = (int)ctas;  // one tile group per CTA: the hardware
	                                                                                  // CTA scheduler balances the SMs
	L.sync = 0;
	L.chunk = 1;
	if (kernel == 2) {
		L.sync = knob(K_ROLLOUT_SYNC, 0);
		L.chunk = knob(K_ROLLOUT_CHUNK, ctas <= cap ? 8 : 12);  // measured: flat between 8 and 16
		if (L.chunk < 1) L.chunk = 1;
		const int per_sm = knob(K_ROLLOUT_CTAS_PER_SM, 0);
		if (per_sm > 0) {
			cap = (int64_t)g_num_sms * per_sm;
			L.grid = (int)(ctas < cap ? ctas : cap);
		}
	}
	return L;
}

// single-step kernels: 1-warp CTAs by default (measured 3-7 % faster than 4-warp CTAs once the table staging is
// vectorised: finer-grained balancing by the hardware CTA scheduler); SPL_STEP_WPC=4 selects the 4-warp shape
#define SPL_LAUNCH_STEP(OUT)                                                                 \
	{                                                                                        \
		if (L.wpc == 1) {                                                                    \
			if (do_step) spl_step_kernel<true, 1, OUT><<<L.grid, 32, 0, st>>>(p);             \
			else spl_step_kernel<false, 1, OUT><<<L.grid, 32, 0, st>>>(p);                    \
		} else {                                                                             \
			if (do_step) spl_step_kernel<true, 4, OUT><<<L.grid, 128, 0, st>>>(p);            \
			else spl_step_kernel<false, 4, OUT><<<L.grid, 128, 0, st>>>(p);                   \
		}                                                                                    \
	}

// The deals the step kernels took are replaced in batches, every (16 x slots)-th lock-step by the caller's lock-step
// counter io->action_t: a slot taken at step t is needed again after `slots` more games, i.e. >= 17 x slots moves later.
// The batch costs the latency of one generator chain (~50 us) whatever its size, so a deeper ring makes the bit-exact
// lock-step cheaper: 1 slot ~3.5 us per lock-step, 4 slots < 1 us.
static int refill_spares_if_due(const spl_envs_t* envs, const spl_step_io_t* io, cudaStream_t st) {
	if (io->flags & SPL_IO_ASYNC_REFILL) return 0;  // the caller refills (spl_refill_spares)
	// (capped at 64 lock-steps: callers that re-base action_t per rollout segment, e.g. per replayed CUDA graph, then still
	// reach the cadence inside every segment of 64 x k lock-steps)
	const int slots = spare_slots(envs) < 4 ? spare_slots(envs) : 4;
	const int age = knob(K_SPARE_REFILL_AGE, SPL_SPARE_REFILL_AGE * slots);
	if (age <= 1 || io->action_t % (uint64_t)age == 0)
		return launch_reset(envs, nullptr, nullptr, 0, nullptr, nullptr, st, io, SPL_RESET_SPARE_REFILL_NOW);
	return 0;
}

static int launch_step(const spl_envs_t* e, const spl_step_io_t* io, bool do_step, int32_t* obs, int8_t* mask, cudaStream_t st,
                       void* obs_f16 = nullptr, uint8_t* obs_u8 = nullptr) {
	StepParams p;
	fill_step_params(p, e, io, obs, mask);
	const bool f16 = obs_f16 != nullptr || obs_u8 != nullptr;
	if (f16) {
		if (obs != nullptr) return SPL_E_BADARG;  // one observation format per call
		p.obs_f16 = (__half*)obs_f16, p.obs_u8 = obs_u8;
		p.vec_ok = (((uintptr_t)obs_f16 | (uintptr_t)obs_u8 | (uintptr_t)mask) & 15) == 0;
	}
	LaunchShape L = launch_shape(e->n, do_step ? 0 : 1);
	const bool timed = do_step && g_timing && g_ev_used < SPL_TIMING_POOL;
	if (timed) SPL_CUDA(cudaEventRecord(g_ev[2 * g_ev_used], st));
	const int64_t ntiles = (e->n + 31) / 32;
	// two warps per tile (spl_step_paired_kernel) while every tile still fits the GPU in one wave of 4-tile CTAs, 4 per SM
	const bool paired = !f16 && knob(K_STEP_PAIRED, 0) != 0 && (ntiles + SPL_PAIR_TILES - 1) / SPL_PAIR_TILES <= (int64_t)g_num_sms * 4;
	if (paired) {
		const unsigned grid = (unsigned)((ntiles + SPL_PAIR_TILES - 1) / SPL_PAIR_TILES);
		if (do_step) spl_step_paired_kernel<true><<<grid, SPL_PAIR_TILES * 64, 0, st>>>(p);
		else spl_step_paired_kernel<false><<<grid, SPL_PAIR_TILES * 64, 0, st>>>(p);
	} else if (f16) SPL_LAUNCH_STEP(SPL_OUT_F16)
	else SPL_LAUNCH_STEP(SPL_OUT_I32)
	if (timed) SPL_CUDA(cudaEventRecord(g_ev[2 * g_ev_used++ + 1], st));
	g_launches++;
	return (int)cudaGetLastError();
}

}  // extern "C"

// internal (spl_host.cu): one lock-step (or observe) with COMPACT outputs into obs_u8 [n][297] / side [n]
int spl_launch_compact(const spl_envs_t* e, const spl_step_io_t* io, bool do_step, uint8_t* obs_u8, void* side, cudaStream_t st) {
	int rc = check_envs(e);
	if (rc) return rc;
	// resets must happen inside the step kernel: the native deal, or MT19937 decks with prefetched deals (e->spare)
	const bool spares = e->shuffle_mode == SPL_SHUFFLE_MT19937 && e->spare != nullptr && do_step && io && io->autoreset;
	if (e->shuffle_mode != SPL_SHUFFLE_PHILOX && do_step && io && io->autoreset && !spares) return SPL_E_BADARG;
	StepParams p;
	fill_step_params(p, e, io, nullptr, nullptr);
	p.obs_u8 = obs_u8, p.side = (uint4*)side;
	p.vec_ok = ((uintptr_t)obs_u8 & 15) == 0;
	LaunchShape L = launch_shape(e->n, do_step ? 0 : 1);
	const bool timed = do_step && g_timing && g_ev_used < SPL_TIMING_POOL;
	if (timed) SPL_CUDA(cudaEventRecord(g_ev[2 * g_ev_used], st));
	if (spares) p.reset_mode = SPL_RESET_SPARE_INLINE;
	SPL_LAUNCH_STEP(SPL_OUT_COMPACT)
	if (timed) SPL_CUDA(cudaEventRecord(g_ev[2 * g_ev_used++ + 1], st));
	g_launches++;
	SPL_CUDA(cudaGetLastError());
	if (spares) return refill_spares_if_due(e, io, st);
	return 0;
}

extern "C" {

int spl_step(const spl_envs_t* envs, const spl_step_io_t* io, void* stream) {
	int rc = check_envs(envs);
	if (rc) return rc;
	if (!io || !io->actions || !io->reward || !io->terminated || !io->info) return SPL_E_BADARG;
	cudaStream_t st = (cudaStream_t)stream;
	// Philox decks are dealt inside the step kernel.  MT19937 decks (bit-exact with the reference) come from the env's
	// prefetched deals when it has them (spl_envs_t.spare; an env whose ring is empty is dealt in place by one lane of
	// the step kernel), else finished envs are queued for the reset kernel that follows (624 generator words per env)
	const bool mt = io->autoreset && envs->shuffle_mode != SPL_SHUFFLE_PHILOX;
	const bool spares = mt && envs->spare != nullptr;
	const bool worklist = mt && (!spares || knob(K_SPARE_WORKLIST, 0) != 0);
	if (worklist) SPL_CUDA(cudaMemsetAsync(envs->scratch, 0, 16, st));
	if ((io->obs_f16 || io->obs_u8) && worklist) return SPL_E_BADARG;  // the MT19937 reset kernel writes int32 observations only
	rc = launch_step(envs, io, true, io->obs, io->mask, st, io->obs_f16, io->obs_u8);
	if (rc) return rc;
	if (worklist) {
		// the reset kernel also re-samples next_action for the envs whose mask it replaces
		rc = launch_reset(envs, envs->scratch, nullptr, 1, io->obs, io->mask, st, io);
		if (rc) return rc;
	}
	if (spares) return refill_spares_if_due(envs, io, st);
	return 0;
}

int spl_rollout_random(const spl_envs_t* envs, const spl_step_io_t* io, int32_t steps, void* stream) {
	int rc = check_envs(envs);
	if (rc) return rc;
	if (!io || !io->actions || !io->reward || !io->terminated || !io->next_action || steps <= 0 || io->active) return SPL_E_BADARG;
	// in-kernel resets need the native deal or, for the reference's own decks, deals prepared ahead of time
	const bool mt = envs->shuffle_mode == SPL_SHUFFLE_MT19937;
	if (!io->autoreset || (mt && envs->spare == nullptr) || (!mt && envs->shuffle_mode != SPL_SHUFFLE_PHILOX)) return SPL_E_BADARG;
	StepParams p;
	fill_step_params(p, envs, io, io->obs, io->mask);
	if (mt) p.reset_mode = SPL_RESET_SPARE_INLINE;
	p.steps = steps;
	// every step's tile must stay 16-byte aligned: n*297*4 and n*45 are multiples of 16 iff n % 16 == 0
	if (envs->n % 16 != 0) p.vec_ok = 0;
	LaunchShape L = launch_shape(envs->n, 2);
	p.sync = L.sync;
	p.chunk = L.chunk;
	p.nchunks = spl_num_chunks(steps, L.chunk);
	cudaStream_t st = (cudaStream_t)stream;
	// work queue: unit counter + per-group progress, in the scratch list (n + 4 words >= 4 + groups)
	const int64_t groups = ((envs->n + 31) / 32 + L.wpc - 1) / L.wpc;
	if (groups * (int64_t)p.nchunks >= (int64_t)1 << 32) return SPL_E_BADARG;
	SPL_CUDA(cudaMemsetAsync(envs->scratch, 0, (size_t)(4 + groups) * sizeof(int32_t), st));
	const bool timed = g_timing && g_ev_used < SPL_TIMING_POOL;
	if (timed) SPL_CUDA(cudaEventRecord(g_ev[2 * g_ev_used], st));
	if (L.wpc == 1) spl_rollout_kernel<1><<<L.grid, 32, 0, st>>>(p);
	else spl_rollout_kernel<4><<<L.grid, 128, 0, st>>>(p);
	if (timed) SPL_CUDA(cudaEventRecord(g_ev[2 * g_ev_used++ + 1], st));
	g_launches++;
	SPL_CUDA(cudaGetLastError());
	// the deals the launch consumed are replaced right behind it (the launch itself never waits for a generator)
	if (mt && !(io->flags & SPL_IO_ASYNC_REFILL)) return launch_reset(envs, nullptr, nullptr, 0, nullptr, nullptr, st, io, SPL_RESET_SPARE_REFILL_NOW);
	return 0;
}

int spl_refill_spares(const spl_envs_t* envs, void* stream) {
	int rc = check_envs(envs);
	if (rc) return rc;
	if (envs->shuffle_mode != SPL_SHUFFLE_MT19937 || envs->spare == nullptr) return SPL_E_BADARG;
	return launch_reset(envs, nullptr, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream, nullptr, SPL_RESET_SPARE_REFILL_NOW);
}

int spl_rollout_plan(int64_t n, int32_t steps, int32_t* out) {
	if (n <= 0 || steps <= 0 || !out) return SPL_E_BADARG;
	if (!device_ready()) return SPL_E_NOTINIT;
	LaunchShape L = launch_shape(n, 2);
	const int64_t groups = ((n + 31) / 32 + L.wpc - 1) / L.wpc;
	int chunk = L.chunk, nchunks = spl_num_chunks(steps, L.chunk);
	if (4 + groups > n + 4) chunk = steps, nchunks = 1;
	out[0] = L.wpc, out[1] = L.grid, out[2] = chunk, out[3] = nchunks, out[4] = L.sync, out[5] = (int32_t)groups;
	return 0;
}

int spl_observe(const spl_envs_t* envs, int32_t* obs, int8_t* mask, void* stream) {
	int rc = check_envs(envs);
	if (rc) return rc;
	return launch_step(envs, nullptr, false, obs, mask, (cudaStream_t)stream);
}

int spl_observe_policy(const spl_envs_t* envs, void* obs_f16, uint8_t* obs_u8, int8_t* mask, void* stream) {
	int rc = check_envs(envs);
	if (rc) return rc;
	if (!obs_f16 && !obs_u8) return SPL_E_BADARG;
	return launch_step(envs, nullptr, false, nullptr, mask, (cudaStream_t)stream, obs_f16, obs_u8);
}

int spl_random_action(const int8_t* mask, int64_t n, uint64_t env_offset, uint64_t key, uint64_t t, int32_t* actions, void* stream) {
	if (!mask || !actions || n <= 0) return SPL_E_BADARG;
	if (!device_ready()) return SPL_E_NOTINIT;
	int64_t ctas = ((n + 31) / 32 + 3) / 4;
	int64_t cap = (int64_t)g_num_sms * 8;
	spl_random_action_kernel<<<(int)(ctas < cap ? ctas : cap), 128, 0, (cudaStream_t)stream>>>(mask, n, env_offset, key, t, actions);
	g_launches++;
	return (int)cudaGetLastError();
}

int spl_export_state(const spl_envs_t* envs, int32_t* rows, void* stream) {
	int rc = check_envs(envs);
	if (rc) return rc;
	if (!rows) return SPL_E_BADARG;
	spl_export_kernel<<<(unsigned)((envs->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const uint4*)envs->state, envs->stride,
	                                                                                    envs->decks, envs->n, rows);
	g_launches++;
	return (int)cudaGetLastError();
}

int spl_import_state(const spl_envs_t* envs, const int32_t* rows, const uint8_t* which, void* stream) {
	int rc = check_envs(envs);
	if (rc) return rc;
	if (!rows) return SPL_E_BADARG;
	// import is not on the step path (tests, debugging, the single-env facade): it waits for the kernel and reports rows it
	// rejected (counters >= 128, card / noble ids out of range, list lengths > 3 ...) instead of truncating them silently
	int* bad = nullptr;
	SPL_CUDA(cudaMallocAsync(&bad, sizeof(int), (cudaStream_t)stream));
	SPL_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), (cudaStream_t)stream));
	spl_import_kernel<<<(unsigned)((envs->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>((uint4*)envs->state, envs->stride, envs->decks,
	                                                                                    envs->n, rows, which, bad);
	g_launches++;
	SPL_CUDA(cudaGetLastError());
	int nbad = 0;
	SPL_CUDA(cudaMemcpyAsync(&nbad, bad, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
	SPL_CUDA(cudaFreeAsync(bad, (cudaStream_t)stream));
	SPL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
	return nbad ? SPL_E_BADROW : 0;
}

int spl_load_deals(const spl_envs_t* envs, const uint8_t* deals, int32_t count, void* stream) {
	int rc = check_envs(envs);
	if (rc) return rc;
	if (!deals || count <= 0 || envs->shuffle_mode != SPL_SHUFFLE_MT19937 || envs->spare == nullptr || count > spare_slots(envs)) return SPL_E_BADARG;
	cudaStream_t st = (cudaStream_t)stream;
	const int64_t items = envs->n * count;
	const size_t bytes = (size_t)items * SPL_DECK_STRIDE;
	cudaPointerAttributes at;
	const bool on_device = cudaPointerGetAttributes(&at, deals) == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
	cudaGetLastError();
	uint8_t* staged = nullptr;
	if (!on_device) {
		SPL_CUDA(cudaMallocAsync(&staged, bytes, st));
		SPL_CUDA(cudaMemcpyAsync(staged, deals, bytes, cudaMemcpyHostToDevice, st));
		deals = staged;
	}
	int* bad = nullptr;
	SPL_CUDA(cudaMallocAsync(&bad, sizeof(int), st));
	SPL_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
	const unsigned grid = (unsigned)((items + 127) / 128);
	spl_check_deals_kernel<<<grid, 128, 0, st>>>(deals, items, bad);
	spl_load_deals_kernel<<<grid, 128, 0, st>>>(deals, envs->n, count, envs->episode, envs->spare, spare_slots(envs), bad);
	g_launches += 2;
	SPL_CUDA(cudaGetLastError());
	int nbad = 0;
	SPL_CUDA(cudaMemcpyAsync(&nbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
	SPL_CUDA(cudaFreeAsync(bad, st));
	if (staged) SPL_CUDA(cudaFreeAsync(staged, st));
	SPL_CUDA(cudaStreamSynchronize(st));
	return nbad ? SPL_E_BADROW : 0;
}

int spl_dual_combine(const float* r1, const uint8_t* term1, const uint8_t* info1, const float* r2, const uint8_t* term2,
                     const uint8_t* info2, int64_t n, float* agent_reward, float* opp_reward, uint8_t* done, int mode, void* stream) {
	if (!r1 || !term1 || !info1 || !r2 || !term2 || !info2 || !agent_reward || !opp_reward || !done || n <= 0) return SPL_E_BADARG;
	spl_dual_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(r1, term1, info1, r2, term2, info2, n,
	                                                                                    agent_reward, opp_reward, done, mode);
	g_launches++;
	return (int)cudaGetLastError();
}

}  // extern "C"
