// Host build of the lane-local rules (splendor_gym_b200/csrc/spl_core.cuh) for logic checks on a
// machine without a GPU.  TEST-ONLY: the product package never loads this; it exists so that the
// per-lane code can be compared with the oracle before (and independently of) a GPU run.
#include <stdint.h>
#include <string.h>

#include "../../splendor_gym_b200/csrc/spl_tables_host.h"

struct ArraySink {
	uint32_t* R;
	void first(uint32_t v) { R[0] = v; }
	void put(int k, uint32_t v) { R[k] = v; }
	void last(uint32_t v) { R[74] = v; }
};

static SplTables g_T;
static uint64_t g_ret[SPL_RET_TABLE_LEN];
static bool g_init = false;
static void init() {
	if (g_init) return;
	spl_build_tables(&g_T);
	spl_build_ret_table(g_ret);
	g_init = true;
}

extern "C" {

void emu_ret_table(uint64_t* out) {
	init();
	memcpy(out, g_ret, sizeof(g_ret));
}

uint64_t emu_mt_block(uint64_t seed, uint32_t blk) { return spl_mt_top3_block(seed, blk); }

// the batch dealer's per-thread body: initial_state(seed)'s shuffles with the generator in registers
int emu_mt_deal_stream(uint32_t key, uint8_t* deck100, uint32_t max_outputs) {
	static uint32_t G[624];
	static bool have = false;
	if (!have) {
		spl_mt_init_table(G);
		have = true;
	}
	alignas(4) uint8_t deck[100];
	const bool ok = spl_mt_deal_stream(key, G, deck, max_outputs);
	memcpy(deck100, deck, 100);
	return ok ? 1 : 0;
}

// pack -> unpack round trip of a flat row
void emu_roundtrip(const int32_t* row, int32_t* row_out) {
	init();
	SplState s, s2;
	uint8_t deck[SPL_DECK_STRIDE];
	uint32_t w[16];
	spl_import_row(row, s, deck);
	spl_pack(s, w);
	spl_unpack(w, s2);
	spl_export_row(s2, deck, row_out);
}

void emu_observe(const int32_t* row, int32_t* obs, int8_t* mask) {
	init();
	SplState s;
	uint8_t deck[SPL_DECK_STRIDE];
	uint32_t w[16];
	spl_import_row(row, s, deck);
	spl_pack(s, w);
	uint64_t m = spl_is_terminal(s) ? 0 : spl_legal_mask(s, &g_T);
	for (int i = 0; i < SPL_NUM_ACTIONS; i++) mask[i] = (int8_t)((m >> i) & 1);
	uint32_t R[75];
	ArraySink sink{R};
	spl_encode_observation(w, s, &g_T, sink);
	for (int i = 0; i < SPL_OBS_DIM; i++) obs[i] = (int32_t)((R[i >> 2] >> (8 * (i & 3))) & 0xFF);
}

void emu_env_step(const int32_t* row, int32_t action, int32_t* row_out, int32_t* obs, int8_t* mask, float* reward,
                  uint8_t* terminated, uint8_t* info) {
	init();
	SplState s;
	uint8_t deck[SPL_DECK_STRIDE];
	uint32_t w[16];
	spl_import_row(row, s, deck);
	spl_pack(s, w);  // go through the packed form like the kernel does
	spl_unpack(w, s);
	SplStepResult r;
	spl_env_step(s, action, deck, &g_T, g_ret, r);
	spl_pack(s, w);
	uint64_t m = spl_is_terminal(s) ? 0 : spl_legal_mask(s, &g_T);
	for (int i = 0; i < SPL_NUM_ACTIONS; i++) mask[i] = (int8_t)((m >> i) & 1);
	uint32_t R[75];
	ArraySink sink{R};
	spl_encode_observation(w, s, &g_T, sink);
	for (int i = 0; i < SPL_OBS_DIM; i++) obs[i] = (int32_t)((R[i >> 2] >> (8 * (i & 3))) & 0xFF);
	*reward = r.reward;
	*terminated = (uint8_t)r.terminated;
	*info = (uint8_t)r.info;
	spl_export_row(s, deck, row_out);
}

// spl_import_state's admission rule for a flat row
int emu_row_valid(const int32_t* row) { return spl_row_valid(row) ? 1 : 0; }

// rollout work-unit chunking: chunk c of a `steps`-step rollout -> [start, start + len)
int emu_chunk_bounds(int c, int steps, int chunk, int* start, int* len) { return spl_chunk_bounds(c, steps, chunk, *start, *len) ? 1 : 0; }
int emu_num_chunks(int steps, int chunk) { return spl_num_chunks(steps, chunk); }
}
