// spl_core.cuh -- lane-local Splendor rules on the packed 64-byte state (one environment per lane).
//
// Everything here is per-thread register code with no warp-level or shared-memory dependence, so the
// same source also compiles for the host (tests/emu builds it with g++ to check the logic on a
// machine without a GPU; the product never runs it on the CPU).  The warp-cooperative parts
// (coalesced tile I/O, mask/obs staging) live in spl_kernels.cu.
//
// Reference semantics (YiyangShao/splendor-gym), cited per function:
//   engine/state.py:61-71 can_afford | engine/rules.py:40-93 legal_moves | :101-147 pay/refill/noble
//   :150-193 token return | :196-308 apply_action / compute_winner / is_terminal
//   engine/encode.py:124-187 encode_observation | envs/splendor_env.py:51-90 SplendorEnv.step
#pragma once
#include <stdint.h>

#include "../../include/splendor_b200.h"

#if defined(__CUDACC__)
#define SPL_HD __host__ __device__ __forceinline__
#define SPL_HD_NOINLINE __host__ __device__ __noinline__
#define SPL_HD_MEMBER __host__ __device__ __forceinline__
#else
#define SPL_HD static inline
#define SPL_HD_NOINLINE static
#define SPL_HD_MEMBER inline
#endif

// ------------------------------------------------------------------------------------------------
// small intrinsics with host equivalents
// ------------------------------------------------------------------------------------------------
SPL_HD uint32_t spl_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
	return __popc(x);
#else
	return (uint32_t)__builtin_popcount(x);
#endif
}
SPL_HD uint32_t spl_popcll(uint64_t x) {
#if defined(__CUDA_ARCH__)
	return __popcll(x);
#else
	return (uint32_t)__builtin_popcountll(x);
#endif
}
SPL_HD uint32_t spl_clz(uint32_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
	return (uint32_t)__clz((int)x);
#else
	return (uint32_t)__builtin_clz(x);
#endif
}

SPL_HD uint32_t spl_ffs(uint32_t x) {  // 1-based index of lowest set bit, 0 if none
#if defined(__CUDA_ARCH__)
	return __ffs(x);
#else
	return (uint32_t)__builtin_ffs((int)x);
#endif
}
// sum of the 4 bytes of x (+c): one IDP.4A on the device
SPL_HD uint32_t spl_bytesum4(uint32_t x, uint32_t c) {
#if defined(__CUDA_ARCH__)
	return __dp4a(x, 0x01010101u, c);
#else
	return (x & 0xFF) + ((x >> 8) & 0xFF) + ((x >> 16) & 0xFF) + (x >> 24) + c;
#endif
}
SPL_HD uint32_t spl_min(uint32_t a, uint32_t b) { return a < b ? a : b; }
SPL_HD uint32_t spl_lo(uint64_t x) { return (uint32_t)x; }
SPL_HD uint32_t spl_hi(uint64_t x) { return (uint32_t)(x >> 32); }
SPL_HD uint64_t spl_u64(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }
SPL_HD uint32_t spl_byte(uint64_t v, uint32_t i) { return (uint32_t)(v >> (8 * i)) & 0xFFu; }

// ------------------------------------------------------------------------------------------------
// Constant tables (built on the host by spl_build_tables(), copied to __constant__ then to shared
// memory once per CTA).  Index 90 / 91 of the card tables and 10 / 11 of the noble tables are the
// "absent" entries so that ids are clamped with one min().
// ------------------------------------------------------------------------------------------------
struct alignas(16) SplTables {
	uint32_t card_feat[92][4];  // obs bytes of a card: present,tier,points,onehot5,cost5,revealed(=1),0,0
	uint32_t card_info[92];     // cost nibbles [0:20) | colour [20:23) | points [24:27) | tier [28:30)
	uint32_t noble_feat[12][2]; // obs bytes of a noble: present, req5, 0, 0
	uint32_t noble_req[12];     // req nibbles [0:20) | points [24:28)
	uint16_t take3_lut[32];     // legal take-3 actions (10 bits) per availability set of the 5 colours
};

#define SPL_EMPTY 0xFFu
#define SPL_FLAG_TO_PLAY 1u
#define SPL_FLAG_GAME_OVER 2u
#define SPL_FLAG_TURN_LIMIT 4u
#define SPL_FLAG_WINNER_SHIFT 3 /* 2 bits: 0 none, 1 p0, 2 p1 */
#define SPL_FLAG_WINNER_MASK 0x18u

// ------------------------------------------------------------------------------------------------
// Working (register) form of one environment.  Players are stored in PERSPECTIVE order: index 0 is
// the player to move ("me"), index 1 the opponent -- the same order encode_observation uses
// (engine/encode.py:131-142), so bytes 0..31 of the packed row ARE observation entries 0..31.
// ------------------------------------------------------------------------------------------------
struct SplState {
	uint64_t bank;       // 6 bytes: white, blue, green, red, black, gold
	uint64_t tok[2];     // 6 bytes each
	uint64_t bon[2];     // 5 bytes each
	uint32_t prestige[2];
	uint32_t nres[2];
	uint32_t res[2];     // reserved card ids, bytes 0..2 (0xFF beyond nres), byte 3 = 0
	uint32_t rev[2];     // revealed flags, bits 0..2
	uint32_t nlist[2];   // nobles owned: nibbles 0..2 = noble index (0xF none), nibble 3 = count
	uint32_t board[3];   // 4 card ids per tier (0xFF = empty slot)
	uint32_t deckn;      // bytes 0..2: cards left in each deck
	uint32_t nobles;     // bytes 0..2: visible noble index (0xFF = taken)
	uint32_t flags;      // SPL_FLAG_*
	uint32_t turn, move; // turn_count, move_count
};

// Packed row (16 x u32, little-endian bytes):
//   0-5 bank | 6-11 me.tokens | 12-16 me.bonuses | 17 me.prestige | 18 me.nres
//   19-24 opp.tokens | 25-29 opp.bonuses | 30 opp.prestige | 31 opp.nres
//   32-43 board | 44-46 deck sizes | 47 flags
//   48-50 me.reserved | 51-53 opp.reserved | 54 revealed (bits 0-2 me, 4-6 opp) | 55-57 nobles
//   58 turn_count | 59 move_count | 60-61 me.noble list | 62-63 opp.noble list
SPL_HD void spl_unpack(const uint32_t* w, SplState& s) {
	s.bank = spl_u64(w[0], w[1] & 0xFFFFu);
	s.tok[0] = (uint64_t)(w[1] >> 16) | ((uint64_t)w[2] << 16);
	s.bon[0] = spl_u64(w[3], w[4] & 0xFFu);
	s.prestige[0] = (w[4] >> 8) & 0xFFu;
	s.nres[0] = (w[4] >> 16) & 0xFFu;
	s.tok[1] = (uint64_t)(w[4] >> 24) | ((uint64_t)w[5] << 8) | ((uint64_t)(w[6] & 0xFFu) << 40);
	s.bon[1] = (uint64_t)(w[6] >> 8) | ((uint64_t)(w[7] & 0xFFFFu) << 24);
	s.prestige[1] = (w[7] >> 16) & 0xFFu;
	s.nres[1] = w[7] >> 24;
	s.board[0] = w[8];
	s.board[1] = w[9];
	s.board[2] = w[10];
	s.deckn = w[11] & 0xFFFFFFu;
	s.flags = w[11] >> 24;
	s.res[0] = w[12] & 0xFFFFFFu;
	s.res[1] = (w[12] >> 24) | ((w[13] & 0xFFFFu) << 8);
	s.rev[0] = (w[13] >> 16) & 0x7u;
	s.rev[1] = (w[13] >> 20) & 0x7u;
	s.nobles = (w[13] >> 24) | ((w[14] & 0xFFFFu) << 8);
	s.turn = (w[14] >> 16) & 0xFFu;
	s.move = w[14] >> 24;
	s.nlist[0] = w[15] & 0xFFFFu;
	s.nlist[1] = w[15] >> 16;
}

SPL_HD void spl_pack(const SplState& s, uint32_t* w) {
	w[0] = spl_lo(s.bank);
	w[1] = (spl_hi(s.bank) & 0xFFFFu) | (spl_lo(s.tok[0]) << 16);
	w[2] = (uint32_t)(s.tok[0] >> 16);
	w[3] = spl_lo(s.bon[0]);
	w[4] = (spl_hi(s.bon[0]) & 0xFFu) | (s.prestige[0] << 8) | (s.nres[0] << 16) | (spl_lo(s.tok[1]) << 24);
	w[5] = (uint32_t)(s.tok[1] >> 8);
	w[6] = ((uint32_t)(s.tok[1] >> 40) & 0xFFu) | (spl_lo(s.bon[1]) << 8);
	w[7] = ((uint32_t)(s.bon[1] >> 24) & 0xFFFFu) | (s.prestige[1] << 16) | (s.nres[1] << 24);
	w[8] = s.board[0];
	w[9] = s.board[1];
	w[10] = s.board[2];
	w[11] = s.deckn | (s.flags << 24);
	w[12] = s.res[0] | (s.res[1] << 24);
	w[13] = (s.res[1] >> 8) | (s.rev[0] << 16) | (s.rev[1] << 20) | (s.nobles << 24);
	w[14] = (s.nobles >> 8) | (s.turn << 16) | (s.move << 24);
	w[15] = s.nlist[0] | (s.nlist[1] << 16);
}

SPL_HD void spl_swap_players(SplState& s) {
	uint64_t t64;
	uint32_t t32;
	t64 = s.tok[0], s.tok[0] = s.tok[1], s.tok[1] = t64;
	t64 = s.bon[0], s.bon[0] = s.bon[1], s.bon[1] = t64;
	t32 = s.prestige[0], s.prestige[0] = s.prestige[1], s.prestige[1] = t32;
	t32 = s.nres[0], s.nres[0] = s.nres[1], s.nres[1] = t32;
	t32 = s.res[0], s.res[0] = s.res[1], s.res[1] = t32;
	t32 = s.rev[0], s.rev[0] = s.rev[1], s.rev[1] = t32;
	t32 = s.nlist[0], s.nlist[0] = s.nlist[1], s.nlist[1] = t32;
}

// fresh game (engine/state.py:196-210): bank 4/4/4/4/4/5, empty hands; board / decks / nobles are
// filled in by the reset kernel from the shuffled deck order.
SPL_HD void spl_fresh_state(SplState& s) {
	s.bank = 0x050404040404ull;
	s.tok[0] = s.tok[1] = 0;
	s.bon[0] = s.bon[1] = 0;
	s.prestige[0] = s.prestige[1] = 0;
	s.nres[0] = s.nres[1] = 0;
	s.res[0] = s.res[1] = 0xFFFFFFu;
	s.rev[0] = s.rev[1] = 0;
	s.nlist[0] = s.nlist[1] = 0x0FFFu;
	s.board[0] = s.board[1] = s.board[2] = 0xFFFFFFFFu;
	s.deckn = 0;
	s.nobles = 0xFFFFFFu;
	s.flags = 0;
	s.turn = 1;
	s.move = 0;
}

SPL_HD bool spl_is_terminal(const SplState& s) {  // engine/rules.py:306-308
	return (s.flags & (SPL_FLAG_GAME_OVER | SPL_FLAG_TO_PLAY)) == SPL_FLAG_GAME_OVER;
}

// ------------------------------------------------------------------------------------------------
// SWAR helpers on byte vectors (domain: every counter < 128)
// ------------------------------------------------------------------------------------------------
// per-byte (x >= k) flags at bit 7 of each byte, 1 <= k <= 128
SPL_HD uint32_t spl_ge_flags4(uint32_t x, uint32_t k) {
	return (((x & 0x7F7F7F7Fu) + (0x80u - k) * 0x01010101u) | x) & 0x80808080u;
}
// bit-7 flags of 4 bytes -> 4 low bits
SPL_HD uint32_t spl_compress_flags4(uint32_t f) { return (((f >> 7) * 0x01020408u) >> 24) & 0xFu; }
// 4 low bits -> 4 bytes of 0/1
SPL_HD uint32_t spl_spread4(uint32_t b) { return ((b & 0xFu) * 0x00204081u) & 0x01010101u; }

// colours (0..4) of a 6-byte token vector whose count is >= k, as 5 bits
SPL_HD uint32_t spl_colours_ge(uint64_t v, uint32_t k) {
	uint32_t lo = spl_compress_flags4(spl_ge_flags4(spl_lo(v), k));
	return lo | (((spl_hi(v) & 0xFFu) >= k) ? 16u : 0u);
}
SPL_HD uint32_t spl_sum6(uint64_t v) { return spl_bytesum4(spl_lo(v), spl_bytesum4(spl_hi(v) & 0xFFFFu, 0)); }
SPL_HD uint32_t spl_sum5(uint64_t v) { return spl_bytesum4(spl_lo(v), spl_hi(v) & 0xFFu); }

// 5 bytes -> 5 nibbles of min(byte, 7)
SPL_HD uint32_t spl_nib5_clamp7(uint64_t v) {
	uint32_t x = spl_lo(v);
	uint32_t big = spl_ge_flags4(x, 8) >> 7;           // 0x01 per byte >= 8
	uint32_t c = (x & 0x07070707u) | (big * 7u);       // min(byte, 7)
	c = (c | (c >> 4)) & 0x00FF00FFu;
	c = (c | (c >> 8)) & 0xFFFFu;
	uint32_t h = spl_min(spl_hi(v) & 0xFFu, 7u);
	return c | (h << 16);
}

// engine/state.py:61-71 can_afford, restated on "wealth" w_c = min(7, tokens_c + bonuses_c):
//   gold_needed = sum_c max(0, max(0, cost_c - bonus_c) - tokens_c) = sum_c max(0, cost_c - w_c)
// (all terms are non-negative integers and cost_c <= 7).  cost / wealth are 5 nibbles each.
SPL_HD uint32_t spl_shortfall(uint32_t cost_nib, uint32_t wealth_nib) {
	uint32_t d = ((cost_nib & 0xFFFFFu) | 0x88888u) - wealth_nib;  // nibble = 8 + cost - wealth, no borrows
	uint32_t pos = (d >> 3) & 0x11111u;                            // cost >= wealth
	uint32_t need = d & (pos * 7u);
	uint32_t s = (need & 0x0F0F0Fu) + ((need >> 4) & 0x0F0Fu);
	return ((s * 0x010101u) >> 16) & 0xFFu;
}

SPL_HD uint32_t spl_card_at(const SplState& s, uint32_t slot) {  // board slot 0..11 -> card id
	uint32_t w = slot < 4 ? s.board[0] : (slot < 8 ? s.board[1] : s.board[2]);
	return (w >> (8 * (slot & 3))) & 0xFFu;
}
SPL_HD void spl_set_card_at(SplState& s, uint32_t slot, uint32_t id) {
	uint32_t sh = 8 * (slot & 3);
	uint32_t clr = ~(0xFFu << sh), val = id << sh;
	if (slot < 4) s.board[0] = (s.board[0] & clr) | val;
	else if (slot < 8) s.board[1] = (s.board[1] & clr) | val;
	else s.board[2] = (s.board[2] & clr) | val;
}

// itertools.combinations(range(5),3) as colour bit sets, 5 bits per action (engine/encode.py:35)
#define SPL_TAKE3_COMBOS 0x00039AB3B356CD67ull
SPL_HD uint32_t spl_take3_combo(uint32_t a) { return (uint32_t)(SPL_TAKE3_COMBOS >> (5 * a)) & 31u; }

// ------------------------------------------------------------------------------------------------
// legal_moves (engine/rules.py:40-93) as a 45-bit set for the player to move
// ------------------------------------------------------------------------------------------------
// The 15 card slots (12 board + my 3 reserved) can be evaluated one at a time (begin / slot / finish), so that a caller may
// interleave them with other work; spl_legal_mask below is the plain sequence every kernel uses today.  (Round 2 ran the
// slots inside the observation-tile store loop of the step kernels: no gain)
struct SplMaskBuilder {
	uint32_t wealth, gold, afford, present;
	SPL_HD_MEMBER void begin(const SplState& s) {
		wealth = spl_nib5_clamp7((s.tok[0] & 0xFFFFFFFFFFull) + s.bon[0]);
		gold = spl_byte(s.tok[0], 5);
		afford = 0, present = 0;
	}
	// board slots 0..11 (:66-80), my reserved cards 12..14 (:89-91)
	SPL_HD_MEMBER void slot(const SplState& s, const SplTables* T, uint32_t k) {
		uint32_t word = k < 4 ? s.board[0] : (k < 8 ? s.board[1] : (k < 12 ? s.board[2] : s.res[0]));
		uint32_t id = (word >> (8 * (k & 3))) & 0xFFu;
		uint32_t info = T->card_info[spl_min(id, 90u)];
		bool here = id != SPL_EMPTY;
		present |= here ? (1u << k) : 0u;
		afford |= (here && spl_shortfall(info, wealth) <= gold) ? (1u << k) : 0u;
	}
	SPL_HD_MEMBER uint64_t finish(const SplState& s, const SplTables* T) const {
		uint32_t avail = spl_colours_ge(s.bank, 1);
		uint32_t m_lo = T->take3_lut[avail];                 // bits 0..9   (:45-58)
		m_lo |= spl_colours_ge(s.bank, 4) << 10;             // bits 10..14 (:61-63)
		const uint32_t buy = afford & 0xFFFu, buyres = afford >> 12;
		const uint32_t onboard = present & 0xFFFu;
		bool can_reserve = s.nres[0] < 3;                    // (:74, :83)
		uint32_t blind = spl_compress_flags4(spl_ge_flags4(s.deckn, 1)) & 7u;
		m_lo |= buy << 15;                                    // bits 15..26
		uint32_t resv = can_reserve ? onboard : 0u;           // bits 27..38
		uint32_t rb = can_reserve ? blind : 0u;               // bits 39..41
		m_lo |= resv << 27;
		uint32_t m_hi = (resv >> 5) | (rb << 7) | (buyres << 10);
		return spl_u64(m_lo, m_hi);
	}
};

SPL_HD uint64_t spl_legal_mask(const SplState& s, const SplTables* T) {
	SplMaskBuilder b;
	b.begin(s);
	// one loop body, kept rolled on purpose -- the step kernels are instruction-cache bound, not ALU bound
#pragma unroll 3
	for (uint32_t k = 0; k < 15; k++) b.slot(s, T, k);
	return b.finish(s, T);
}

// legality of one action without building the whole mask (what `mask[action] != 1` checks,
// envs/splendor_env.py:64)
SPL_HD bool spl_action_legal(const SplState& s, uint32_t a, uint32_t avail, const SplTables* T) {
	if (a < 10) return (T->take3_lut[avail] >> a) & 1u;
	if (a < 15) return spl_byte(s.bank, a - 10) >= 4;
	bool can_reserve = s.nres[0] < 3;
	if (a >= 27 && a < 39) return can_reserve && spl_card_at(s, a - 27) != SPL_EMPTY;
	if (a >= 39 && a < 42) return can_reserve && ((s.deckn >> (8 * (a - 39))) & 0xFFu) > 0;
	uint32_t id = (a < 27) ? spl_card_at(s, a - 15) : ((s.res[0] >> (8 * (a - 42))) & 0xFFu);
	if (id == SPL_EMPTY) return false;
	uint32_t wealth = spl_nib5_clamp7((s.tok[0] & 0xFFFFFFFFFFull) + s.bon[0]);
	return spl_shortfall(T->card_info[id], wealth) <= spl_byte(s.tok[0], 5);
}

// ------------------------------------------------------------------------------------------------
// CPython random.Random(seed) restated with O(1) storage (engine/rules.py:166,173 calls it with a
// state-derived seed).  Returns the top 3 bits of MT19937 outputs [21*blk, 21*blk+21) packed 3 bits
// each -- all that getrandbits(k<=3) consumes.  init_by_array is a pair of sequential recurrences over
// the 624-word state; instead of materialising the state we re-run the first recurrence alongside the
// second and keep only the 43 words the requested outputs depend on (mt[i], mt[i+1], mt[i+397]).
// Used only for seeds outside the tabulated domain (hand-built states); blk <= 9.
// ------------------------------------------------------------------------------------------------
SPL_HD_NOINLINE uint64_t spl_mt_top3_block(uint64_t seed, uint32_t blk) {
	const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
	const uint32_t klen = key[1] ? 2u : 1u;
	const uint32_t lo = 21u * blk;
	uint32_t a[22], b[21];  // final mt[lo .. lo+21], mt[lo+397 .. lo+417]
	// pass 1 of init_by_array: m1[i], i = 1..623 (mt[0] is the init_genrand seed constant)
	uint32_t g = 19650218u, q = g, m1_1 = 0;
	for (uint32_t i = 1; i < 624; i++) {
		g = 1812433253u * (g ^ (g >> 30)) + i;
		uint32_t j = (i - 1) % klen;
		q = (g ^ ((q ^ (q >> 30)) * 1664525u)) + key[j] + j;
		if (i == 1) m1_1 = q;
	}
	// 624th iteration wraps: mt[0] = m1[623]; mt[1] is updated once more
	uint32_t j623 = 623u % klen;
	uint32_t m1p_1 = (m1_1 ^ ((q ^ (q >> 30)) * 1664525u)) + key[j623] + j623;
	// pass 2 over i = 2..623, re-running pass 1 in lock-step for the m1[i] it consumes
	uint32_t p = m1p_1;
	g = 1812433253u * (19650218u ^ (19650218u >> 30)) + 1u;  // g_1
	q = m1_1;
	for (uint32_t i = 2; i < 624; i++) {
		g = 1812433253u * (g ^ (g >> 30)) + i;
		uint32_t j = (i - 1) % klen;
		q = (g ^ ((q ^ (q >> 30)) * 1664525u)) + key[j] + j;
		p = (q ^ ((p ^ (p >> 30)) * 1566083941u)) - i;
		if (i >= lo && i <= lo + 21) a[i - lo] = p;
		if (i >= lo + 397 && i <= lo + 417) b[i - lo - 397] = p;
	}
	// wrap: mt[0] = mt[623]; final mt[1]; then mt[0] = 0x80000000
	uint32_t mt1 = (m1p_1 ^ ((p ^ (p >> 30)) * 1566083941u)) - 1u;
	if (lo == 0) {
		a[0] = 0x80000000u;
		a[1] = mt1;
	}
	uint64_t out = 0;
	for (uint32_t k = 0; k < 21; k++) {
		uint32_t y = (a[k] & 0x80000000u) | (a[k + 1] & 0x7fffffffu);
		uint32_t v = b[k] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
		v ^= v >> 11;
		v ^= (v << 7) & 0x9d2c5680u;
		v ^= (v << 15) & 0xefc60000u;
		v ^= v >> 18;
		out |= (uint64_t)(v >> 29) << (3 * k);
	}
	return out;
}

// ------------------------------------------------------------------------------------------------
// initial_state(seed)'s four shuffles (engine/state.py:186-195) with CPython's MT19937 held entirely in registers --
// the per-thread body of the batch dealer (spl_spare_deal_kernel); see the notes there.  `G` = init_genrand(19650218)
// (624 words, spl_mt_init_table), `key` = the seed as a ONE-word init_by_array key (seeds < 2^32), `deck` = 100 bytes:
// on return bytes 0..39 / 40..69 / 70..89 hold the shuffled tiers (top of a deck = its end) and 90..99 the shuffled
// noble ids.  Returns false (deck contents then meaningless) when the deal would need more than min(227, max_outputs)
// generator outputs: output k of the first generation is a function of the final state words k, k+1, k+397 only, which
// two restarted copies of the seeding recurrence deliver in order, but output 227 needs the next generation.
// ------------------------------------------------------------------------------------------------
SPL_HD void spl_mt_init_table(uint32_t* g) {  // Modules/_randommodule.c init_genrand(19650218)
	g[0] = 19650218u;
	for (uint32_t i = 1; i < 624; i++) g[i] = 1812433253u * (g[i - 1] ^ (g[i - 1] >> 30)) + i;
}

struct SplMTChain {  // one copy of the (pass 1, pass 2) recurrences of init_by_array at index i: q = pass-1 word i, p = final word i
	uint32_t q, p, i;
	SPL_HD_MEMBER uint32_t step(const uint32_t* G, uint32_t key) {
		i++;
		q = (G[i] ^ ((q ^ (q >> 30)) * 1664525u)) + key;
		p = (q ^ ((p ^ (p >> 30)) * 1566083941u)) - i;
		return p;
	}
};

SPL_HD bool spl_mt_deal_stream(uint32_t key, const uint32_t* G, uint8_t* deck, uint32_t max_outputs) {
	// ---- random.Random(key): init_by_array pass 1 (for its last word), then pass 2 with pass 1 re-run next to it
	uint32_t q = (G[1] ^ ((G[0] ^ (G[0] >> 30)) * 1664525u)) + key;
	const uint32_t q1 = q;
#pragma unroll 8
	for (int i = 2; i < 624; i++) q = (G[i] ^ ((q ^ (q >> 30)) * 1664525u)) + key;
	const uint32_t p1 = (q1 ^ ((q ^ (q >> 30)) * 1664525u)) + key;  // 624th iteration: word 1 once more, word 0 <- word 623
	SplMTChain c;
	c.q = q1, c.p = p1, c.i = 1;
#pragma unroll 8
	for (int i = 2; i <= 397; i++) c.step(G, key);
	SplMTChain hi = c;  // at word 397
	const uint32_t mt397 = c.p;
#pragma unroll 8
	for (int i = 398; i < 624; i++) c.step(G, key);
	const uint32_t mt1 = (p1 ^ ((c.p ^ (c.p >> 30)) * 1566083941u)) - 1u;
	SplMTChain lo;  // at word 1 again: produces words 2, 3, ...
	lo.q = q1, lo.p = p1, lo.i = 1;
	// ---- shuffle(deck1), shuffle(deck2), shuffle(deck3), shuffle(nobles); Lib/random.py shuffle: for i in
	// reversed(range(1, len)): j = _randbelow(i + 1); x[i], x[j] = x[j], x[i] -- as ONE flat loop over (deck, position)
	{  // deck[k] = k for the 90 cards, then the nobles 0..9 at bytes 90..99: 25 word stores
		uint32_t* dw = reinterpret_cast<uint32_t*>(deck);
#pragma unroll
		for (int k = 0; k < 22; k++) dw[k] = 0x03020100u + 0x04040404u * (uint32_t)k;
		dw[22] = 0x01005958u, dw[23] = 0x05040302u, dw[24] = 0x09080706u;
	}
	uint32_t a = 0x80000000u, nout = 0;  // state word 0 after init_by_array
	int seg = 0, base = 0, i = 39;
	while (seg < 4) {
		uint32_t b, cw;
		if (nout == 0) b = mt1, cw = mt397;
		else {
			if (hi.i >= 623u || nout >= max_outputs) return false;
			b = lo.step(G, key), cw = hi.step(G, key);
		}
		nout++;
		const uint32_t u = (a & 0x80000000u) | (b & 0x7fffffffu);
		uint32_t y = cw ^ (u >> 1) ^ ((u & 1u) ? 0x9908b0dfu : 0u);
		a = b;
		y ^= y >> 11;
		y ^= (y << 7) & 0x9d2c5680u;
		y ^= (y << 15) & 0xefc60000u;
		y ^= y >> 18;
		const uint32_t r = y >> spl_clz((uint32_t)i + 1u);  // getrandbits((i + 1).bit_length())
		if (r <= (uint32_t)i) {
			const uint8_t t = deck[base + i];
			deck[base + i] = deck[base + r];
			deck[base + r] = t;
			if (--i < 1) {
				seg++;
				base = seg == 1 ? 40 : (seg == 2 ? 70 : 90);
				i = seg == 1 ? 29 : (seg == 2 ? 19 : 9);
			}
		}
	}
	return true;
}

// auto_return_tokens / _enforce_token_limit (engine/rules.py:150-193) for the mover (index 0).
// `tp` = actual to_play, `turn` = turn_count, both BEFORE the end-of-turn increment (:160-165).
// index of the tabulated token-return stream for (turn_count, to_play, hand total, bank total), or 0xFFFFFFFF outside
// the domain that legal play reaches (turn 1..99, hand 11..13, bank 0..14)
SPL_HD uint32_t spl_ret_index(uint32_t turn, uint32_t tp, uint32_t total, uint32_t bank_sum) {
	const bool tabulated = (turn - 1u < 99u) && (total - 11u < 3u) && (bank_sum < 15u);
	return tabulated ? (((turn - 1u) * 2u + tp) * 3u + (total - 11u)) * 15u + bank_sum : 0xFFFFFFFFu;
}
SPL_HD uint64_t spl_ret_load(const uint64_t* ret_table, uint32_t idx) {
#if defined(__CUDA_ARCH__)
	return __ldg(reinterpret_cast<const unsigned long long*>(ret_table) + idx);
#else
	return ret_table[idx];
#endif
}

// `early_idx` / `early_val`: a table entry the caller loaded ahead of time (spl_env_step_t predicts the index from the
// action); used when it is the entry needed, otherwise the entry is loaded here.
SPL_HD void spl_enforce_token_limit(SplState& s, uint32_t tp, uint32_t turn, const uint64_t* ret_table, uint32_t& err,
                                    uint32_t early_idx = 0xFFFFFFFFu, uint64_t early_val = 0) {
	uint32_t total = spl_sum6(s.tok[0]);
	if (total <= 10) return;
	uint32_t remaining = total - 10;
	uint32_t bank_sum = spl_sum6(s.bank);
	// 21 three-bit draws per 64-bit block; block 0 of every seed reachable in legitimate play is tabulated,
	// anything else (hand-built states, or > 21 draws) is recomputed by spl_mt_top3_block -- one call site.
	const uint64_t seed = ((uint64_t)turn * 1315423911ull) ^ ((uint64_t)tp * 2654435761ull) ^ ((uint64_t)total * 97531ull) ^
	                      ((uint64_t)bank_sum * 31337ull);
	const uint32_t idx = spl_ret_index(turn, tp, total, bank_sum);
	uint64_t stream = 0;
	uint32_t pos = 21, next_blk = 0;
	if (idx != 0xFFFFFFFFu) {
		stream = idx == early_idx ? early_val : spl_ret_load(ret_table, idx);
		pos = 0;
		next_blk = 1;
	}
	while (remaining > 0) {
		uint32_t choices = spl_colours_ge(s.tok[0], 1);  // ascending colour order (:170)
		uint32_t n = spl_popc(choices);
		if (n == 0) break;
		uint32_t drop = 3u - (n >= 4 ? 3u : (n >= 2 ? 2u : 1u));  // 3 - n.bit_length()
		uint32_t r;
		do {  // Lib/random.py _randbelow_with_getrandbits
			if (pos == 21) {
				if (next_blk > 9) {  // > 210 MT outputs: outside what the O(1)-storage restatement covers
					err = 1;
					return;
				}
				stream = spl_mt_top3_block(seed, next_blk++);
				pos = 0;
			}
			r = ((uint32_t)(stream >> (3 * pos)) & 7u) >> drop;
			pos++;
		} while (r >= n);
		uint32_t cset = choices;
		for (uint32_t k = 0; k < r; k++) cset &= cset - 1;  // drop the r lowest choices
		uint32_t c = spl_ffs(cset) - 1;
		s.tok[0] -= 1ull << (8 * c);
		s.bank += 1ull << (8 * c);
		remaining--;
	}
	if (remaining > 0) {  // gold only as a last resort (:179-184)
		uint32_t gold = spl_byte(s.tok[0], 5);
		uint32_t give = spl_min(remaining, gold);
		s.tok[0] -= (uint64_t)give << 40;
		s.bank += (uint64_t)give << 40;
	}
}

struct SplStepResult {
	float reward;
	uint32_t terminated;
	uint32_t info;
};

// ------------------------------------------------------------------------------------------------
// SplendorEnv.step (envs/splendor_env.py:51-90) + apply_action (engine/rules.py:196-287) in place.
// `deck` = this env's deck-order row (top of tier t = deck[off_t + deckn_t - 1]).
// ------------------------------------------------------------------------------------------------
// KNOWN_MASK: the caller already holds legal_moves(state) as a bit set (`cur_mask`) -- the reference computes
// exactly that at the top of step() (:55) -- otherwise legality is evaluated for the one action only.
SPL_HD uint32_t spl_deck_byte(const uint8_t* deck, uint32_t idx) {
#if defined(__CUDA_ARCH__)
	return __ldcg(deck + idx);  // L2-coherent: the row may have been written by other lanes (fused reset)
#else
	return deck[idx];
#endif
}

// top card of each deck (bytes 0..2; 0xFF when the deck is empty) -- see `tops` below
SPL_HD uint32_t spl_deck_tops(const SplState& s, const uint8_t* deck) {
	uint32_t tops = 0xFFFFFFu;
	const uint32_t off[3] = {0u, 40u, 70u};
#pragma unroll
	for (int t = 0; t < 3; t++) {
		uint32_t dn = (s.deckn >> (8 * t)) & 0xFFu;
		if (dn > 0) tops = (tops & ~(0xFFu << (8 * t))) | (spl_deck_byte(deck, off[t] + dn - 1) << (8 * t));
	}
	return tops;
}

// KNOWN_MASK: see above.  `tops` (nullable): the caller keeps the top card of each deck in a register across
// steps; a pop then takes the card from the register and issues the load of the NEW top, which nobody waits for
// until that deck is popped again -- the dependent global load leaves the critical path of the step.
template <bool KNOWN_MASK>
SPL_HD void spl_env_step_t(SplState& s, int32_t action, const uint8_t* deck, const SplTables* T, const uint64_t* ret_table,
                           SplStepResult& out, uint64_t cur_mask, uint32_t* tops = nullptr) {
	out.reward = 0.0f;
	out.terminated = 0;
	out.info = 0;
	const uint32_t tp = s.flags & SPL_FLAG_TO_PLAY;
	if (spl_is_terminal(s)) {  // RuntimeError in the reference (:53-54)
		out.terminated = 1;
		out.info = SPL_INFO_ERROR | SPL_INFO_TERMINATED;
		return;
	}
	const uint32_t avail = spl_colours_ge(s.bank, 1);
	// any legal move? a non-empty bank always allows a take-3 (:45-58), so the full mask is only
	// needed when all five colours are exhausted
	bool any;
	if (KNOWN_MASK) {
		any = cur_mask != 0;
	} else {
		any = avail != 0;
		if (!any) any = spl_legal_mask(s, T) != 0;
	}
	if (!any) {  // no-legal-move draw (:55-61): game_over, winner None, to_play = 0; counters untouched
		s.flags = (s.flags | SPL_FLAG_GAME_OVER) & ~(SPL_FLAG_WINNER_MASK | SPL_FLAG_TO_PLAY);
		if (tp) spl_swap_players(s);
		out.terminated = 1;
		out.info = SPL_INFO_NOLEGAL_DRAW | SPL_INFO_TERMINATED;
		return;
	}
	const uint32_t a = (uint32_t)action;
	if (a >= SPL_NUM_ACTIONS) {  // ValueError (:62-63)
		out.info = SPL_INFO_ERROR;
		return;
	}
	const bool legal = KNOWN_MASK ? (((cur_mask >> a) & 1ull) != 0) : spl_action_legal(s, a, avail, T);
	if (!legal) {  // (:64-66)
		out.reward = -0.01f;
		out.info = SPL_INFO_ILLEGAL;
		return;
	}

	// ---- apply_action: decode the action into orthogonal effects (engine/rules.py:201-257) ----
	uint64_t take = 0;             // tokens bank -> me
	uint32_t pay_card = SPL_EMPTY; // card bought
	uint32_t res_card = SPL_EMPTY; // card reserved from the board
	bool reserve = false, revealed = false;
	int32_t refill = -1;           // board slot to refill
	int32_t pop_tier = -1;         // deck to pop
	int32_t remove_res = -1;       // reserved index bought
	if (a < 10) {                  // take-3: colours in the combo that the bank still has (:201-210)
		uint32_t got = spl_take3_combo(a) & avail;
		take = spl_u64(spl_spread4(got), got >> 4);
	} else if (a < 15) {           // take-2 (:211-215)
		take = 2ull << (8 * (a - 10));
	} else if (a < 27) {           // buy visible (:216-225)
		refill = (int32_t)(a - 15);
		pay_card = spl_card_at(s, (uint32_t)refill);
	} else if (a < 39) {           // reserve visible (:226-240)
		refill = (int32_t)(a - 27);
		res_card = spl_card_at(s, (uint32_t)refill);
		reserve = true;
		revealed = true;
	} else if (a < 42) {           // reserve blind (:241-249)
		pop_tier = (int32_t)(a - 39);
		reserve = true;
	} else {                       // buy reserved (:250-255)
		remove_res = (int32_t)(a - 42);
		pay_card = (s.res[0] >> (8 * remove_res)) & 0xFFu;
	}
	if (refill >= 0) pop_tier = refill >> 2;

	// A token return (:150-193) reads ONE table entry, indexed by the hand and bank totals after the action.  Only take and
	// reserve actions can push a hand over 10, and for them both totals are known right here: issue the load now, ~200
	// instructions before spl_enforce_token_limit consumes it (which checks the index and reloads if a hand-built state
	// broke the prediction).
	uint32_t early_idx = 0xFFFFFFFFu;
	uint64_t early_val = 0;
	if (pay_card == SPL_EMPTY && remove_res < 0) {
		const uint32_t ntake = reserve ? (spl_byte(s.bank, 5) > 0 ? 1u : 0u) : spl_sum6(take);
		const uint32_t total = spl_sum6(s.tok[0]) + ntake;
		if (total > 10) {
			early_idx = spl_ret_index(s.turn, tp, total, spl_sum6(s.bank) - ntake);
			if (early_idx != 0xFFFFFFFFu) early_val = spl_ret_load(ret_table, early_idx);
		}
	}

	// deck.pop() from the END of the list (:127, :244)
	uint32_t popped = SPL_EMPTY;
	if (pop_tier >= 0) {
		uint32_t dn = (s.deckn >> (8 * pop_tier)) & 0xFFu;
		if (dn > 0) {
			uint32_t off = pop_tier == 0 ? 0u : (pop_tier == 1 ? 40u : 70u);
			if (tops != nullptr) {
				const uint32_t sh = 8u * (uint32_t)pop_tier;
				popped = (*tops >> sh) & 0xFFu;
				const uint32_t next = dn > 1 ? spl_deck_byte(deck, off + dn - 2) : SPL_EMPTY;
				*tops = (*tops & ~(0xFFu << sh)) | (next << sh);
			} else {
				popped = spl_deck_byte(deck, off + dn - 1);
			}
			s.deckn -= 1u << (8 * pop_tier);
		}
	}
	if (refill >= 0) spl_set_card_at(s, (uint32_t)refill, popped);  // _refill_slot (:125-129)
	else if (pop_tier >= 0) res_card = popped;

	if (pay_card != SPL_EMPTY) {  // _pay_for_card (:101-122)
		uint32_t info = T->card_info[pay_card];
		uint64_t spend = 0;
		uint32_t gold_spent = 0;
#pragma unroll
		for (int c = 0; c < 5; c++) {
			uint32_t cost = (info >> (4 * c)) & 0xFu;
			uint32_t bonus = spl_byte(s.bon[0], c);
			uint32_t disc = cost > bonus ? cost - bonus : 0u;
			uint32_t pay = spl_min(spl_byte(s.tok[0], c), disc);
			spend |= (uint64_t)pay << (8 * c);
			gold_spent += disc - pay;  // legal => total <= gold held, so min(remaining, gold left) == remaining
		}
		spend |= (uint64_t)gold_spent << 40;
		s.tok[0] -= spend;
		s.bank += spend;
		s.bon[0] += 1ull << (8 * ((info >> 20) & 7u));
		s.prestige[0] += (info >> 24) & 7u;
	}
	if (remove_res >= 0) {  // reserved.pop(idx) shifts later cards down (:253-254)
		uint32_t lowb = (1u << (8 * remove_res)) - 1u;
		s.res[0] = (s.res[0] & lowb) | ((s.res[0] >> 8) & ~lowb) | 0xFF0000u;
		uint32_t lowr = (1u << remove_res) - 1u;
		s.rev[0] = (s.rev[0] & lowr) | ((s.rev[0] >> 1) & ~lowr);
		s.nres[0] -= 1;
	}
	if (reserve) {  // reserved.append(card) + one gold if the bank has any (:234-239, :245-249)
		uint32_t k = s.nres[0];
		s.res[0] = (s.res[0] & ~(0xFFu << (8 * k))) | (res_card << (8 * k));
		s.rev[0] |= (revealed ? 1u : 0u) << k;
		s.nres[0] = k + 1;
		if (spl_byte(s.bank, 5) > 0) take = 1ull << 40;
	}
	s.bank -= take;
	s.tok[0] += take;

	// _grant_noble_if_applicable (:132-147): first visible noble whose requirements are met, at most one
	{
		uint32_t bon = spl_nib5_clamp7(s.bon[0]) | 0x88888u;
		bool done = false;
#pragma unroll
		for (int i = 0; i < 3; i++) {
			uint32_t nb = (s.nobles >> (8 * i)) & 0xFFu;
			uint32_t rq = T->noble_req[spl_min(nb, 10u)];
			bool meets = nb != SPL_EMPTY && (((bon - (rq & 0xFFFFFu)) & 0x88888u) == 0x88888u);
			if (meets && !done) {
				done = true;
				s.prestige[0] += (rq >> 24) & 0xFu;
				s.nobles |= 0xFFu << (8 * i);
				uint32_t cnt = s.nlist[0] >> 12;
				if (cnt < 3) s.nlist[0] = (s.nlist[0] & ~(0xFu << (4 * cnt))) | (nb << (4 * cnt));
				s.nlist[0] = (s.nlist[0] & 0x0FFFu) | ((cnt + 1) << 12);
			}
		}
	}

	uint32_t err = 0;
	spl_enforce_token_limit(s, tp, s.turn, ret_table, err, early_idx, early_val);  // (:261)

	// end of turn (:263-285)
	uint32_t flags = s.flags;
	if (s.prestige[0] >= 15) flags |= SPL_FLAG_GAME_OVER;
	s.move += 1;
	flags ^= SPL_FLAG_TO_PLAY;
	s.turn = (s.move >> 1) + 1;
	spl_swap_players(s);  // perspective order follows to_play: the mover is now index 1
	if (s.turn >= 100) {  // turn limit: draw, even if somebody reached 15 (:275-279)
		flags = (flags | SPL_FLAG_GAME_OVER | SPL_FLAG_TURN_LIMIT) & ~SPL_FLAG_WINNER_MASK;
	} else if ((flags & SPL_FLAG_GAME_OVER) && !(flags & SPL_FLAG_TO_PLAY)) {
		// compute_winner (:290-303): key (prestige, -cards, -reserved); exact tie -> None.
		// index 0 is actual player 0 here because to_play == 0
		uint32_t c0 = spl_sum5(s.bon[0]), c1 = spl_sum5(s.bon[1]);
		int32_t w = -1;
		if (s.prestige[0] != s.prestige[1]) w = s.prestige[0] > s.prestige[1] ? 0 : 1;
		else if (c0 != c1) w = c0 < c1 ? 0 : 1;
		else if (s.nres[0] != s.nres[1]) w = s.nres[0] < s.nres[1] ? 0 : 1;
		flags = (flags & ~SPL_FLAG_WINNER_MASK) | ((uint32_t)(w + 1) << SPL_FLAG_WINNER_SHIFT);
	}
	s.flags = flags;

	// reward from the mover's perspective (envs/splendor_env.py:68-88)
	if ((flags & (SPL_FLAG_GAME_OVER | SPL_FLAG_TO_PLAY)) == SPL_FLAG_GAME_OVER) {
		uint32_t wcode = (flags & SPL_FLAG_WINNER_MASK) >> SPL_FLAG_WINNER_SHIFT;
		bool limit = flags & SPL_FLAG_TURN_LIMIT;
		if (wcode == 0) out.reward = limit ? -0.1f : 0.0f;
		else out.reward = (wcode - 1 == tp) ? 1.0f : -1.0f;  // mover = old to_play
		out.terminated = 1;
		out.info = SPL_INFO_TERMINATED | (limit ? SPL_INFO_TURN_LIMIT : 0u) | (wcode << SPL_INFO_WINNER_SHIFT);
	}
	if (err) out.info |= SPL_INFO_ERROR;
}

SPL_HD void spl_env_step(SplState& s, int32_t action, const uint8_t* deck, const SplTables* T, const uint64_t* ret_table,
                         SplStepResult& out) {
	spl_env_step_t<false>(s, action, deck, T, ret_table, out, 0ull);
}

// ------------------------------------------------------------------------------------------------
// encode_observation (engine/encode.py:124-187) as 75 little-endian words of one byte per entry
// (entries 297..299 are zero padding).  `w` is the PACKED row of the same state: words 0..7 are
// entries 0..31 verbatim.  emit(k, word) is called for k = 0..74 in order.
// ------------------------------------------------------------------------------------------------
SPL_HD uint32_t spl_bp(uint32_t lo, uint32_t hi, uint32_t sel) {  // byte permute of the 8 bytes {hi:lo}
#if defined(__CUDA_ARCH__)
	return __byte_perm(lo, hi, sel);
#else
	uint64_t v = spl_u64(lo, hi);
	uint32_t r = 0;
	for (int i = 0; i < 4; i++) {
		uint32_t n = (sel >> (4 * i)) & 0xF;
		uint32_t b = (uint32_t)(v >> (8 * (n & 7))) & 0xFF;
		if (n & 8) b = (b & 0x80) ? 0xFF : 0x00;
		r |= b << (8 * i);
	}
	return r;
#endif
}

struct SplFeat {
	uint32_t x, y, z, w;
};
SPL_HD SplFeat spl_card_feat(const SplTables* T, uint32_t id) {
	const uint32_t* f = T->card_feat[spl_min(id, 90u)];
#if defined(__CUDA_ARCH__)
	uint4 v = *reinterpret_cast<const uint4*>(f);
	return SplFeat{v.x, v.y, v.z, v.w};
#else
	return SplFeat{f[0], f[1], f[2], f[3]};
#endif
}

// Sink interface: first(v) = word 0, put(k, v) for 1 <= k <= 73 (k may be a run-time value), last(v) = word 74.
template <class Sink>
SPL_HD void spl_encode_observation(const uint32_t* w, const SplState& s, const SplTables* T, Sink& sink) {
	// [0:32) bank, me, opponent
	sink.first(w[0]);
#pragma unroll
	for (int k = 1; k < 8; k++) sink.put(k, w[k]);
	// [32:188) board: 12 cards x 13 entries; four cards = 52 bytes = 13 words (records at byte 0,13,26,39)
#pragma unroll 1
	for (int t = 0; t < 3; t++) {
		uint32_t ids = t == 0 ? s.board[0] : (t == 1 ? s.board[1] : s.board[2]);
		SplFeat A = spl_card_feat(T, ids & 0xFF), B = spl_card_feat(T, (ids >> 8) & 0xFF);
		SplFeat C = spl_card_feat(T, (ids >> 16) & 0xFF), D = spl_card_feat(T, ids >> 24);
		int k = 8 + 13 * t;
		sink.put(k + 0, A.x);
		sink.put(k + 1, A.y);
		sink.put(k + 2, A.z);
		sink.put(k + 3, spl_bp(A.w, B.x, 0x6540));                          // A12 B0 B1 B2
		sink.put(k + 4, spl_bp(B.x, B.y, 0x6543));                          // B3..B6
		sink.put(k + 5, spl_bp(B.y, B.z, 0x6543));                          // B7..B10
		sink.put(k + 6, spl_bp(spl_bp(B.z, B.w, 0x0043), C.x, 0x5410));     // B11 B12 C0 C1
		sink.put(k + 7, spl_bp(C.x, C.y, 0x5432));                          // C2..C5
		sink.put(k + 8, spl_bp(C.y, C.z, 0x5432));                          // C6..C9
		sink.put(k + 9, spl_bp(spl_bp(C.z, C.w, 0x0432), D.x, 0x4210));     // C10 C11 C12 D0
		sink.put(k + 10, spl_bp(D.x, D.y, 0x4321));                         // D1..D4
		sink.put(k + 11, spl_bp(D.y, D.z, 0x4321));                         // D5..D8
		sink.put(k + 12, spl_bp(D.z, D.w, 0x4321));                         // D9..D12
	}
	// [188:272) reserved: my three (always revealed=1), then the opponent's (a hidden one is 14 zeros,
	// engine/encode.py:158-168).  Six 14-entry records = three aligned pairs of 7 words.
	const uint32_t rv = s.rev[1];
	const uint32_t theirs = s.res[1] | ((rv & 1u) ? 0u : 0xFFu) | ((rv & 2u) ? 0u : 0xFF00u) | ((rv & 4u) ? 0u : 0xFF0000u);
	const uint64_t six = (uint64_t)s.res[0] | ((uint64_t)theirs << 24);
#pragma unroll 1
	for (int pr = 0; pr < 3; pr++) {
		uint32_t two = (uint32_t)(six >> (16 * pr));
		SplFeat X = spl_card_feat(T, two & 0xFF), Y = spl_card_feat(T, (two >> 8) & 0xFF);
		int k = 47 + 7 * pr;
		sink.put(k + 0, X.x);
		sink.put(k + 1, X.y);
		sink.put(k + 2, X.z);
		sink.put(k + 3, spl_bp(X.w, Y.x, 0x5410));  // X12 X13 Y0 Y1
		sink.put(k + 4, spl_bp(Y.x, Y.y, 0x5432));  // Y2..Y5
		sink.put(k + 5, spl_bp(Y.y, Y.z, 0x5432));  // Y6..Y9
		sink.put(k + 6, spl_bp(Y.z, Y.w, 0x5432));  // Y10..Y13
	}
	// [272:290) nobles 3 x (present, req5); [290:293) deck sizes; turn_count, to_play, move_count, terminal
	{
		const uint32_t* X = T->noble_feat[spl_min(s.nobles & 0xFF, 10u)];
		const uint32_t* Y = T->noble_feat[spl_min((s.nobles >> 8) & 0xFF, 10u)];
		const uint32_t* Z = T->noble_feat[spl_min((s.nobles >> 16) & 0xFF, 10u)];
		sink.put(68, X[0]);
		sink.put(69, spl_bp(X[1], Y[0], 0x5410));
		sink.put(70, spl_bp(Y[0], Y[1], 0x5432));
		sink.put(71, Z[0]);
		sink.put(72, spl_bp(Z[1], s.deckn, 0x5410));
		sink.put(73, ((s.deckn >> 16) & 0xFFu) | (s.turn << 8) | ((s.flags & SPL_FLAG_TO_PLAY) << 16) | (s.move << 24));
		sink.last(spl_is_terminal(s) ? 1u : 0u);
	}
}

// ------------------------------------------------------------------------------------------------
// packed state <-> flat int32 row (include/splendor_b200.h SPL_ROW_*), absolute player order
// ------------------------------------------------------------------------------------------------
SPL_HD void spl_export_row(const SplState& s, const uint8_t* deck, int32_t* row) {
	for (int i = 0; i < SPL_ROW_LEN; i++) row[i] = -1;
	for (int i = 0; i < 6; i++) row[i] = (int32_t)spl_byte(s.bank, i);
	uint32_t tp = s.flags & SPL_FLAG_TO_PLAY;
	for (int q = 0; q < 2; q++) {       // q = perspective index, p = actual player
		int p = q ^ (int)tp;
		int o = SPL_ROW_PLAYER0 + SPL_ROW_PLAYER_STRIDE * p;
		for (int i = 0; i < 6; i++) row[o + i] = (int32_t)spl_byte(s.tok[q], i);
		for (int i = 0; i < 5; i++) row[o + 6 + i] = (int32_t)spl_byte(s.bon[q], i);
		row[o + 11] = (int32_t)s.prestige[q];
		row[o + 12] = (int32_t)s.nres[q];
		for (int i = 0; i < 3; i++) {
			uint32_t id = (s.res[q] >> (8 * i)) & 0xFFu;
			bool have = (uint32_t)i < s.nres[q];
			row[o + 13 + i] = have ? (int32_t)id : -1;
			row[o + 16 + i] = have ? (int32_t)((s.rev[q] >> i) & 1u) : 0;
		}
		uint32_t cnt = s.nlist[q] >> 12;
		row[o + 19] = (int32_t)cnt;
		for (int i = 0; i < 3; i++) row[o + 20 + i] = (uint32_t)i < cnt ? (int32_t)((s.nlist[q] >> (4 * i)) & 0xFu) : -1;
	}
	for (int k = 0; k < 12; k++) {
		uint32_t id = (s.board[k >> 2] >> (8 * (k & 3))) & 0xFFu;
		row[SPL_ROW_BOARD + k] = id == SPL_EMPTY ? -1 : (int32_t)id;
	}
	for (int t = 0; t < 3; t++) row[SPL_ROW_DECK_SIZES + t] = (int32_t)((s.deckn >> (8 * t)) & 0xFFu);
	for (int i = 0; i < 3; i++) {
		uint32_t nb = (s.nobles >> (8 * i)) & 0xFFu;
		row[SPL_ROW_NOBLES + i] = nb == SPL_EMPTY ? -1 : (int32_t)nb;
	}
	row[SPL_ROW_TO_PLAY] = (int32_t)tp;
	row[SPL_ROW_TURN_COUNT] = (int32_t)s.turn;
	row[SPL_ROW_MOVE_COUNT] = (int32_t)s.move;
	row[SPL_ROW_GAME_OVER] = (s.flags & SPL_FLAG_GAME_OVER) ? 1 : 0;
	row[SPL_ROW_WINNER] = (int32_t)((s.flags & SPL_FLAG_WINNER_MASK) >> SPL_FLAG_WINNER_SHIFT) - 1;
	row[SPL_ROW_TURN_LIMIT] = (s.flags & SPL_FLAG_TURN_LIMIT) ? 1 : 0;
	const int off[3] = {SPL_ROW_DECK1, SPL_ROW_DECK2, SPL_ROW_DECK3};
	const int doff[3] = {0, 40, 70};
	for (int t = 0; t < 3; t++) {
		int dn = (int)((s.deckn >> (8 * t)) & 0xFFu);
		for (int k = 0; k < dn; k++) row[off[t] + k] = (int32_t)deck[doff[t] + k];
	}
}

// Is a flat row inside the domain of the packed state?  Counters are bytes whose SWAR arithmetic needs values < 128,
// ids index the 90-card / 10-noble tables, list lengths are at most 3.  spl_import_state rejects rows that are not
// (nothing is truncated silently).
SPL_HD bool spl_row_valid(const int32_t* row) {
	bool ok = true;
	for (int i = 0; i < 6; i++) ok = ok && (uint32_t)row[i] < 128u;
	for (int p = 0; p < 2; p++) {
		const int o = SPL_ROW_PLAYER0 + SPL_ROW_PLAYER_STRIDE * p;
		for (int i = 0; i < 12; i++) ok = ok && (uint32_t)row[o + i] < 128u;  // tokens, bonuses, prestige
		const int32_t nres = row[o + 12], nnob = row[o + 19];
		ok = ok && nres >= 0 && nres <= 3 && nnob >= 0 && nnob <= 3;
		for (int i = 0; i < 3; i++) {
			if (i < nres) ok = ok && (uint32_t)row[o + 13 + i] < 90u;
			if (i < nnob) ok = ok && (uint32_t)row[o + 20 + i] < 10u;
		}
	}
	for (int k = 0; k < 12; k++) ok = ok && row[SPL_ROW_BOARD + k] >= -1 && row[SPL_ROW_BOARD + k] < 90;
	const int off[3] = {SPL_ROW_DECK1, SPL_ROW_DECK2, SPL_ROW_DECK3};
	const int dlen[3] = {40, 30, 20};
	for (int t = 0; t < 3; t++) {
		const int32_t dn = row[SPL_ROW_DECK_SIZES + t];
		ok = ok && dn >= 0 && dn <= dlen[t];
		for (int k = 0; k < dlen[t]; k++)
			if (k < dn) ok = ok && (uint32_t)row[off[t] + k] < 90u;
	}
	for (int i = 0; i < 3; i++) ok = ok && row[SPL_ROW_NOBLES + i] >= -1 && row[SPL_ROW_NOBLES + i] < 10;
	ok = ok && (uint32_t)row[SPL_ROW_TO_PLAY] < 2u && (uint32_t)row[SPL_ROW_TURN_COUNT] < 256u && (uint32_t)row[SPL_ROW_MOVE_COUNT] < 256u;
	ok = ok && row[SPL_ROW_WINNER] >= -1 && row[SPL_ROW_WINNER] < 2;
	return ok;
}

SPL_HD void spl_import_row(const int32_t* row, SplState& s, uint8_t* deck) {
	s.bank = 0;
	for (int i = 0; i < 6; i++) s.bank |= (uint64_t)(row[i] & 0xFF) << (8 * i);
	uint32_t tp = (uint32_t)row[SPL_ROW_TO_PLAY] & 1u;
	for (int q = 0; q < 2; q++) {
		int p = q ^ (int)tp;
		int o = SPL_ROW_PLAYER0 + SPL_ROW_PLAYER_STRIDE * p;
		s.tok[q] = 0;
		s.bon[q] = 0;
		for (int i = 0; i < 6; i++) s.tok[q] |= (uint64_t)(row[o + i] & 0xFF) << (8 * i);
		for (int i = 0; i < 5; i++) s.bon[q] |= (uint64_t)(row[o + 6 + i] & 0xFF) << (8 * i);
		s.prestige[q] = (uint32_t)row[o + 11] & 0xFFu;
		s.nres[q] = (uint32_t)row[o + 12] & 0xFFu;
		s.res[q] = 0;
		s.rev[q] = 0;
		for (int i = 0; i < 3; i++) {
			bool have = (uint32_t)i < s.nres[q];
			s.res[q] |= (have ? ((uint32_t)row[o + 13 + i] & 0xFFu) : SPL_EMPTY) << (8 * i);
			s.rev[q] |= (have && row[o + 16 + i]) ? (1u << i) : 0u;
		}
		uint32_t cnt = (uint32_t)row[o + 19] & 0xFu;
		s.nlist[q] = cnt << 12;
		for (int i = 0; i < 3; i++) s.nlist[q] |= ((uint32_t)i < cnt ? ((uint32_t)row[o + 20 + i] & 0xFu) : 0xFu) << (4 * i);
	}
	s.board[0] = s.board[1] = s.board[2] = 0;
	for (int k = 0; k < 12; k++) {
		int32_t id = row[SPL_ROW_BOARD + k];
		s.board[k >> 2] |= (id < 0 ? SPL_EMPTY : (uint32_t)id & 0xFFu) << (8 * (k & 3));
	}
	s.deckn = 0;
	for (int t = 0; t < 3; t++) s.deckn |= ((uint32_t)row[SPL_ROW_DECK_SIZES + t] & 0xFFu) << (8 * t);
	s.nobles = 0;
	for (int i = 0; i < 3; i++) {
		int32_t nb = row[SPL_ROW_NOBLES + i];
		s.nobles |= (nb < 0 ? SPL_EMPTY : (uint32_t)nb & 0xFFu) << (8 * i);
	}
	s.turn = (uint32_t)row[SPL_ROW_TURN_COUNT] & 0xFFu;
	s.move = (uint32_t)row[SPL_ROW_MOVE_COUNT] & 0xFFu;
	s.flags = tp | (row[SPL_ROW_GAME_OVER] ? SPL_FLAG_GAME_OVER : 0u) | (row[SPL_ROW_TURN_LIMIT] ? SPL_FLAG_TURN_LIMIT : 0u) |
	          ((uint32_t)(row[SPL_ROW_WINNER] + 1) << SPL_FLAG_WINNER_SHIFT);
	const int off[3] = {SPL_ROW_DECK1, SPL_ROW_DECK2, SPL_ROW_DECK3};
	const int doff[3] = {0, 40, 70};
	const int dlen[3] = {40, 30, 20};
	for (int t = 0; t < 3; t++) {
		int dn = (int)((s.deckn >> (8 * t)) & 0xFFu);
		for (int k = 0; k < dlen[t]; k++) deck[doff[t] + k] = k < dn ? (uint8_t)row[off[t] + k] : (uint8_t)SPL_EMPTY;
	}
	for (int k = 90; k < SPL_DECK_STRIDE; k++) deck[k] = (uint8_t)SPL_EMPTY;
}

// ------------------------------------------------------------------------------------------------
// Rollout work units: how `steps` lock-steps are cut into chunks (see the work-queue comment in spl_kernels.cu)
// ------------------------------------------------------------------------------------------------
struct SplChunking {
	int full, tail_steps;  // `full` chunks of `chunk` steps, then tail_steps split by halving
};

SPL_HD SplChunking spl_chunking(int steps, int chunk) {
	SplChunking k;
	k.full = steps >= 2 * chunk ? (steps - chunk) / chunk : 0;
	k.tail_steps = steps - k.full * chunk;
	return k;
}

// chunk index -> [start, start+len) in lock-steps; returns false past the last chunk
SPL_HD bool spl_chunk_bounds(int c, int steps, int chunk, int& start, int& len) {
	const SplChunking k = spl_chunking(steps, chunk);
	if (c < k.full) {
		start = c * chunk, len = chunk;
		return true;
	}
	int rem = k.tail_steps, s0 = k.full * chunk;
	for (int j = k.full;; j++) {
		if (rem <= 0) return false;
		const int l = rem > 3 ? rem / 2 : rem;
		if (j == c) {
			start = s0, len = l;
			return true;
		}
		s0 += l, rem -= l;
	}
}

SPL_HD int spl_num_chunks(int steps, int chunk) {
	int c = 0, a, b;
	while (spl_chunk_bounds(c, steps, chunk, a, b)) c++;
	return c;
}

