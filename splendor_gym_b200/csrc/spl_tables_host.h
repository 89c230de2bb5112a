// spl_tables_host.h -- host-side construction of the device lookup tables from the plain card /
// noble data (include/spl_tables.h, generated from the reference's cards.json / nobles.json).
#pragma once
#include <string.h>

#include "../../include/spl_tables.h"
#include "spl_core.cuh"

static inline void spl_build_tables(SplTables* T) {
	memset(T, 0, sizeof(*T));
	for (int id = 0; id < SPL_NUM_CARDS; id++) {
		const unsigned char* c = SPL_CARD_TABLE[id];  // tier, colour, points, cost[5]
		unsigned char f[16] = {0};
		f[0] = 1;
		f[1] = c[0];
		f[2] = c[2];
		f[3 + c[1]] = 1;
		for (int k = 0; k < 5; k++) f[8 + k] = c[3 + k];
		f[13] = 1;  // "revealed" entry of a reserved-card record (engine/encode.py:118-119)
		memcpy(T->card_feat[id], f, 16);
		uint32_t info = 0;
		for (int k = 0; k < 5; k++) info |= (uint32_t)c[3 + k] << (4 * k);
		info |= (uint32_t)c[1] << 20 | (uint32_t)c[2] << 24 | (uint32_t)c[0] << 28;
		T->card_info[id] = info;
	}
	for (int n = 0; n < SPL_NUM_NOBLES; n++) {
		const unsigned char* r = SPL_NOBLE_TABLE[n];  // req[5], points
		unsigned char f[8] = {1, r[0], r[1], r[2], r[3], r[4], 0, 0};
		memcpy(T->noble_feat[n], f, 8);
		uint32_t rq = 0;
		for (int k = 0; k < 5; k++) rq |= (uint32_t)r[k] << (4 * k);
		T->noble_req[n] = rq | ((uint32_t)r[5] << 24);
	}
	// take-3 legality per availability set (engine/rules.py:45-58)
	for (uint32_t avail = 0; avail < 32; avail++) {
		uint32_t n = (uint32_t)__builtin_popcount(avail), bits = 0;
		for (uint32_t a = 0; a < 10; a++) {
			uint32_t combo = spl_take3_combo(a);
			bool ok = n >= 3 ? (combo & ~avail) == 0 : (n >= 1 ? (avail & ~combo) == 0 : false);
			bits |= ok ? (1u << a) : 0u;
		}
		T->take3_lut[avail] = (uint16_t)bits;
	}
}

// Token-return stream table (engine/rules.py:160-173): for every seed reachable in legitimate play
// -- turn_count 1..99, to_play 0..1, hand 11..13, bank total 0..14 -- the top 3 bits of the first 21
// MT19937 outputs of random.Random(seed).  Order: ((turn-1)*2 + to_play)*3 + (hand-11))*15 + bank.
static inline void spl_build_ret_table(uint64_t* out) {
	int idx = 0;
	for (uint64_t turn = 1; turn < 100; turn++)
		for (uint64_t tp = 0; tp < 2; tp++)
			for (uint64_t hand = 11; hand < 14; hand++)
				for (uint64_t bank = 0; bank < 15; bank++) {
					uint64_t seed = (turn * 1315423911ull) ^ (tp * 2654435761ull) ^ (hand * 97531ull) ^ (bank * 31337ull);
					out[idx++] = spl_mt_top3_block(seed, 0);
				}
}
