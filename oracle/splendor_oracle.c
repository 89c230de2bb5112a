/*
 * splendor_oracle.c -- CPU restatement of the reference Splendor engine.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA path.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it; the product package never does.
 *
 * It restates, in plain scalar C, the algorithm of YiyangShao/splendor-gym (pure Python):
 *   engine/state.py:61-71     PlayerState.can_afford
 *   engine/state.py:181-211   initial_state (seeded MT19937 shuffles, deal by pop() from list end)
 *   engine/rules.py:40-93     legal_moves
 *   engine/rules.py:101-147   _pay_for_card, _refill_slot, _grant_noble_if_applicable
 *   engine/rules.py:150-193   auto_return_tokens / _enforce_token_limit (state-seeded MT19937)
 *   engine/rules.py:196-308   apply_action, compute_winner, is_terminal
 *   engine/encode.py:20-35    action layout, TAKE3_COMBOS
 *   engine/encode.py:124-187  encode_observation
 *   envs/splendor_env.py:41-115  SplendorEnv.reset / step / get_final_rewards
 *
 * Third-party arithmetic on the path that is NOT under /root/reference: CPython's `random`
 * module (MT19937; Modules/_randommodule.c and Lib/random.py of CPython 3.9-3.12: seed(int) =
 * init_by_array over the 32-bit little-endian words of abs(seed); getrandbits(k<=32) = genrand>>(32-k);
 * _randbelow(n) = rejection on getrandbits(n.bit_length()); shuffle = reversed Fisher-Yates;
 * choice = seq[_randbelow(len)]).  Restated below from the published MT19937 reference code
 * (Matsumoto & Nishimura 2002, mt19937ar.c).
 *
 * PARITY PINNING: this oracle is pinned against outputs of the reference itself, executed in the
 * build container (oracle/gen_golden.py -> tests/golden/*.json; tests/test_oracle_golden.py), and,
 * when /root/reference is present, live against the Python engine (tests/test_oracle_vs_pyref.py).
 *
 * It deliberately shares no code with splendor_gym_b200/csrc (no packed state, no SWAR, no tables
 * beyond the plain card/noble data in include/spl_tables.h).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/spl_tables.h"

#define NUM_ACTIONS 45
#define OBS_DIM 297
#define ROW_LEN 166

/* info bits: keep in sync with include/splendor_b200.h (SPL_INFO_*) */
#define INFO_ILLEGAL 1u
#define INFO_NOLEGAL_DRAW 2u
#define INFO_TURN_LIMIT 4u
#define INFO_TERMINATED 8u
#define INFO_WINNER_SHIFT 4
#define INFO_ERROR 64u
#define INFO_RESET 128u

/* ------------------------------------------------------------------------------------------
 * MT19937, CPython flavour
 * ------------------------------------------------------------------------------------------ */
typedef struct {
	uint32_t mt[624];
	int idx;
} MT;

static void mt_init_genrand(MT *m, uint32_t s) {
	m->mt[0] = s;
	for (int i = 1; i < 624; i++)
		m->mt[i] = 1812433253u * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
	m->idx = 624;
}

static void mt_init_by_array(MT *m, const uint32_t *key, int len) {
	mt_init_genrand(m, 19650218u);
	int i = 1, j = 0;
	int k = 624 > len ? 624 : len;
	for (; k; k--) {
		m->mt[i] = (m->mt[i] ^ ((m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
		i++;
		j++;
		if (i >= 624) {
			m->mt[0] = m->mt[623];
			i = 1;
		}
		if (j >= len) j = 0;
	}
	for (k = 623; k; k--) {
		m->mt[i] = (m->mt[i] ^ ((m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
		i++;
		if (i >= 624) {
			m->mt[0] = m->mt[623];
			i = 1;
		}
	}
	m->mt[0] = 0x80000000u;
}

/* random.Random(a) for a non-negative Python int that fits in 64 bits */
static void mt_seed_u64(MT *m, uint64_t a) {
	uint32_t key[2] = {(uint32_t)(a & 0xffffffffu), (uint32_t)(a >> 32)};
	mt_init_by_array(m, key, key[1] ? 2 : 1);
}

static uint32_t mt_next(MT *m) {
	if (m->idx >= 624) {
		uint32_t *mt = m->mt;
		int kk;
		for (kk = 0; kk < 624 - 397; kk++) {
			uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
			mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
		}
		for (; kk < 623; kk++) {
			uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
			mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
		}
		uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
		mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
		m->idx = 0;
	}
	uint32_t y = m->mt[m->idx++];
	y ^= (y >> 11);
	y ^= (y << 7) & 0x9d2c5680u;
	y ^= (y << 15) & 0xefc60000u;
	y ^= (y >> 18);
	return y;
}

static int bit_length(uint32_t n) {
	int k = 0;
	while (n) {
		k++;
		n >>= 1;
	}
	return k;
}

/* Lib/random.py _randbelow_with_getrandbits */
static uint32_t mt_randbelow(MT *m, uint32_t n) {
	int k = bit_length(n);
	uint32_t r = mt_next(m) >> (32 - k);
	while (r >= n) r = mt_next(m) >> (32 - k);
	return r;
}

/* Lib/random.py shuffle */
static void mt_shuffle(MT *m, int *x, int len) {
	for (int i = len - 1; i >= 1; i--) {
		int j = (int)mt_randbelow(m, (uint32_t)(i + 1));
		int t = x[i];
		x[i] = x[j];
		x[j] = t;
	}
}

/* ------------------------------------------------------------------------------------------
 * State (engine/state.py:36-104)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
	int tokens[6];
	int bonuses[5];
	int prestige;
	int n_reserved;
	int reserved[3];
	int revealed[3];
	int n_nobles;
	int nobles[3];
} OPlayer;

typedef struct {
	int bank[6];
	OPlayer pl[2];
	int board[3][4]; /* card id or -1 */
	int deck[3][40]; /* top of deck = deck[t][deck_n[t]-1] */
	int deck_n[3];
	int nobles[3]; /* noble index or -1 (taken) */
	int to_play, turn_count, move_count, game_over, winner, turn_limit;
} OState;

static const int TIER_OFF[3] = {0, 40, 70};
static const int TIER_LEN[3] = {40, 30, 20};
/* itertools.combinations(range(5), 3), engine/encode.py:35 */
static const int TAKE3[10][3] = {{0, 1, 2}, {0, 1, 3}, {0, 1, 4}, {0, 2, 3}, {0, 2, 4},
                                 {0, 3, 4}, {1, 2, 3}, {1, 2, 4}, {1, 3, 4}, {2, 3, 4}};

static int card_tier(int id) { return SPL_CARD_TABLE[id][0]; }
static int card_color(int id) { return SPL_CARD_TABLE[id][1]; }
static int card_points(int id) { return SPL_CARD_TABLE[id][2]; }
static int card_cost(int id, int c) { return SPL_CARD_TABLE[id][3 + c]; }

/* engine/state.py:181-211 */
static void o_initial_state(OState *s, uint64_t seed) {
	MT m;
	memset(s, 0, sizeof(*s));
	mt_seed_u64(&m, seed);
	for (int t = 0; t < 3; t++) {
		for (int i = 0; i < TIER_LEN[t]; i++) s->deck[t][i] = TIER_OFF[t] + i;
		s->deck_n[t] = TIER_LEN[t];
		mt_shuffle(&m, s->deck[t], TIER_LEN[t]);
		for (int i = 0; i < 4; i++) s->board[t][i] = s->deck_n[t] ? s->deck[t][--s->deck_n[t]] : -1;
	}
	int nob[10];
	for (int i = 0; i < 10; i++) nob[i] = i;
	mt_shuffle(&m, nob, 10);
	for (int i = 0; i < 3; i++) s->nobles[i] = nob[i];
	static const int bank0[6] = {4, 4, 4, 4, 4, 5};
	memcpy(s->bank, bank0, sizeof(bank0));
	for (int p = 0; p < 2; p++)
		for (int i = 0; i < 3; i++) s->pl[p].reserved[i] = s->pl[p].nobles[i] = -1;
	s->to_play = 0;
	s->turn_count = 1;
	s->move_count = 0;
	s->game_over = 0;
	s->winner = -1;
	s->turn_limit = 0;
}

/* engine/state.py:61-71 */
static int o_can_afford(const OPlayer *p, int card) {
	int gold_needed = 0;
	for (int i = 0; i < 5; i++) {
		int req = card_cost(card, i);
		int discounted = req - p->bonuses[i];
		if (discounted < 0) discounted = 0;
		int pay = p->tokens[i] < discounted ? p->tokens[i] : discounted;
		gold_needed += discounted - pay;
	}
	return p->tokens[5] >= gold_needed;
}

/* engine/rules.py:40-93 */
static void o_legal_moves(const OState *s, int8_t *mask) {
	memset(mask, 0, NUM_ACTIONS);
	const OPlayer *p = &s->pl[s->to_play];
	int avail[5], n_avail = 0;
	for (int i = 0; i < 5; i++) {
		avail[i] = s->bank[i] >= 1;
		n_avail += avail[i];
	}
	for (int idx = 0; idx < 10; idx++) {
		int in_combo[5] = {0, 0, 0, 0, 0};
		for (int k = 0; k < 3; k++) in_combo[TAKE3[idx][k]] = 1;
		if (n_avail >= 3) {
			int ok = 1; /* action colours subset of available */
			for (int c = 0; c < 5; c++)
				if (in_combo[c] && !avail[c]) ok = 0;
			if (ok) mask[idx] = 1;
		} else if (n_avail >= 1) {
			int ok = 1; /* available subset of action colours */
			for (int c = 0; c < 5; c++)
				if (avail[c] && !in_combo[c]) ok = 0;
			if (ok) mask[idx] = 1;
		}
	}
	for (int i = 0; i < 5; i++)
		if (s->bank[i] >= 4) mask[10 + i] = 1;
	for (int t = 0; t < 3; t++)
		for (int k = 0; k < 4; k++) {
			int card = s->board[t][k];
			if (card >= 0 && o_can_afford(p, card)) mask[15 + t * 4 + k] = 1;
		}
	if (p->n_reserved < 3) {
		for (int t = 0; t < 3; t++)
			for (int k = 0; k < 4; k++)
				if (s->board[t][k] >= 0) mask[27 + t * 4 + k] = 1;
		for (int t = 0; t < 3; t++)
			if (s->deck_n[t] > 0) mask[39 + t] = 1;
	}
	int nr = p->n_reserved < 3 ? p->n_reserved : 3;
	for (int i = 0; i < nr; i++)
		if (o_can_afford(p, p->reserved[i])) mask[42 + i] = 1;
}

/* engine/rules.py:101-122 */
static void o_pay_for_card(OPlayer *p, int *bank, int card) {
	int gold_available = p->tokens[5];
	int gold_spent = 0;
	for (int i = 0; i < 5; i++) {
		int req = card_cost(card, i);
		int discounted = req - p->bonuses[i];
		if (discounted < 0) discounted = 0;
		int spend = p->tokens[i] < discounted ? p->tokens[i] : discounted;
		p->tokens[i] -= spend;
		bank[i] += spend;
		int remaining = discounted - spend;
		if (remaining > 0) {
			int use_gold = gold_available - gold_spent;
			if (remaining < use_gold) use_gold = remaining;
			gold_spent += use_gold;
		}
	}
	p->tokens[5] -= gold_spent;
	bank[5] += gold_spent;
	p->bonuses[card_color(card)] += 1;
	p->prestige += card_points(card);
}

/* engine/rules.py:125-129 */
static void o_refill_slot(OState *s, int t, int k) {
	if (s->deck_n[t] > 0)
		s->board[t][k] = s->deck[t][--s->deck_n[t]];
	else
		s->board[t][k] = -1;
}

/* engine/rules.py:132-147 */
static void o_grant_noble(OPlayer *p, int *nobles) {
	for (int idx = 0; idx < 3; idx++) {
		int n = nobles[idx];
		if (n < 0) continue;
		int meets = 1;
		for (int i = 0; i < 5; i++)
			if (p->bonuses[i] < SPL_NOBLE_TABLE[n][i]) {
				meets = 0;
				break;
			}
		if (meets) {
			if (p->n_nobles < 3) p->nobles[p->n_nobles] = n;
			p->n_nobles++;
			p->prestige += SPL_NOBLE_TABLE[n][5];
			nobles[idx] = -1;
			break;
		}
	}
}

/* engine/rules.py:150-193 */
static void o_enforce_token_limit(OPlayer *p, OState *s) {
	int total = 0;
	for (int i = 0; i < 6; i++) total += p->tokens[i];
	if (total <= 10) return;
	int remaining = total - 10;
	int bank_sum = 0;
	for (int i = 0; i < 6; i++) bank_sum += s->bank[i];
	uint64_t seed = ((uint64_t)s->turn_count * 1315423911ull) ^ ((uint64_t)s->to_play * 2654435761ull) ^
	                ((uint64_t)total * 97531ull) ^ ((uint64_t)bank_sum * 31337ull);
	MT m;
	mt_seed_u64(&m, seed);
	while (remaining > 0) {
		int choices[5], n = 0;
		for (int i = 0; i < 5; i++)
			if (p->tokens[i] > 0) choices[n++] = i;
		if (n == 0) break;
		int idx = choices[mt_randbelow(&m, (uint32_t)n)];
		p->tokens[idx] -= 1;
		s->bank[idx] += 1;
		remaining -= 1;
	}
	if (remaining > 0 && p->tokens[5] > 0) {
		int give = remaining < p->tokens[5] ? remaining : p->tokens[5];
		p->tokens[5] -= give;
		s->bank[5] += give;
	}
}

/* engine/rules.py:290-303; returns -1 for None */
static int o_compute_winner(const OState *s) {
	int key[2][3];
	for (int p = 0; p < 2; p++) {
		int nb = 0;
		for (int i = 0; i < 5; i++) nb += s->pl[p].bonuses[i];
		key[p][0] = s->pl[p].prestige;
		key[p][1] = -nb;
		key[p][2] = -s->pl[p].n_reserved;
	}
	for (int i = 0; i < 3; i++) {
		if (key[0][i] > key[1][i]) return 0;
		if (key[0][i] < key[1][i]) return 1;
	}
	return -1;
}

/* engine/rules.py:306-308 */
static int o_is_terminal(const OState *s) { return s->game_over && s->to_play == 0; }

/* engine/rules.py:196-287; mutates in place (the reference copies first; callers here own the state).
 * Returns 0, or -1 for an invalid action index (ValueError in the reference). */
static int o_apply_action(OState *s, int action) {
	OPlayer *p = &s->pl[s->to_play];
	int *bank = s->bank;
	if (action >= 0 && action < 10) {
		for (int k = 0; k < 3; k++) {
			int c = TAKE3[action][k];
			if (bank[c] >= 1) {
				bank[c] -= 1;
				p->tokens[c] += 1;
			}
		}
	} else if (action >= 10 && action < 15) {
		int c = action - 10;
		bank[c] -= 2;
		p->tokens[c] += 2;
	} else if (action >= 15 && action < 27) {
		int t = (action - 15) / 4, k = (action - 15) % 4;
		int card = s->board[t][k];
		o_pay_for_card(p, bank, card);
		s->board[t][k] = -1;
		o_refill_slot(s, t, k);
	} else if (action >= 27 && action < 39) {
		int t = (action - 27) / 4, k = (action - 27) % 4;
		int card = s->board[t][k];
		s->board[t][k] = -1;
		p->reserved[p->n_reserved] = card;
		p->revealed[p->n_reserved] = 1;
		p->n_reserved++;
		if (bank[5] > 0) {
			bank[5] -= 1;
			p->tokens[5] += 1;
		}
		o_refill_slot(s, t, k);
	} else if (action >= 39 && action < 42) {
		int t = action - 39;
		int card = s->deck[t][--s->deck_n[t]];
		p->reserved[p->n_reserved] = card;
		p->revealed[p->n_reserved] = 0;
		p->n_reserved++;
		if (bank[5] > 0) {
			bank[5] -= 1;
			p->tokens[5] += 1;
		}
	} else if (action >= 42 && action < 45) {
		int idx = action - 42;
		int card = p->reserved[idx];
		for (int i = idx; i < p->n_reserved - 1; i++) { /* list.pop(idx) */
			p->reserved[i] = p->reserved[i + 1];
			p->revealed[i] = p->revealed[i + 1];
		}
		p->n_reserved--;
		p->reserved[p->n_reserved] = -1;
		p->revealed[p->n_reserved] = 0;
		o_pay_for_card(p, bank, card);
	} else {
		return -1;
	}
	o_grant_noble(p, s->nobles);
	o_enforce_token_limit(p, s);
	if (p->prestige >= 15) s->game_over = 1;
	s->move_count += 1;
	s->to_play = (s->to_play + 1) % 2;
	s->turn_count = s->move_count / 2 + 1;
	if (s->turn_count >= 100) {
		s->game_over = 1;
		s->turn_limit = 1;
		s->winner = -1;
		return 0;
	}
	if (s->game_over && s->to_play == 0) s->winner = o_compute_winner(s);
	return 0;
}

/* engine/encode.py:77-121 */
static int32_t *o_encode_card(int32_t *v, int card) {
	if (card < 0) {
		for (int i = 0; i < 13; i++) *v++ = 0;
		return v;
	}
	*v++ = 1;
	*v++ = card_tier(card);
	*v++ = card_points(card);
	for (int c = 0; c < 5; c++) *v++ = (card_color(card) == c);
	for (int c = 0; c < 5; c++) *v++ = card_cost(card, c);
	return v;
}

/* engine/encode.py:124-187 */
static void o_encode_observation(const OState *s, int32_t *obs) {
	int32_t *v = obs;
	for (int i = 0; i < 6; i++) *v++ = s->bank[i];
	const OPlayer *p = &s->pl[s->to_play];
	const OPlayer *opp = &s->pl[(s->to_play + 1) % 2];
	const OPlayer *both[2] = {p, opp};
	for (int w = 0; w < 2; w++) {
		const OPlayer *q = both[w];
		for (int i = 0; i < 6; i++) *v++ = q->tokens[i];
		for (int i = 0; i < 5; i++) *v++ = q->bonuses[i];
		*v++ = q->prestige;
		*v++ = q->n_reserved;
	}
	for (int t = 0; t < 3; t++)
		for (int k = 0; k < 4; k++) v = o_encode_card(v, s->board[t][k]);
	for (int i = 0; i < 3; i++) {
		if (i < p->n_reserved) {
			v = o_encode_card(v, p->reserved[i]);
			*v++ = 1;
		} else {
			for (int j = 0; j < 14; j++) *v++ = 0;
		}
	}
	for (int i = 0; i < 3; i++) {
		if (i < opp->n_reserved && opp->revealed[i]) {
			v = o_encode_card(v, opp->reserved[i]);
			*v++ = 1;
		} else {
			for (int j = 0; j < 14; j++) *v++ = 0;
		}
	}
	for (int i = 0; i < 3; i++) {
		int n = s->nobles[i];
		if (n >= 0) {
			*v++ = 1;
			for (int c = 0; c < 5; c++) *v++ = SPL_NOBLE_TABLE[n][c];
		} else {
			for (int j = 0; j < 6; j++) *v++ = 0;
		}
	}
	for (int t = 0; t < 3; t++) *v++ = s->deck_n[t];
	*v++ = s->turn_count;
	*v++ = s->to_play;
	*v++ = s->move_count;
	*v++ = o_is_terminal(s) ? 1 : 0;
}

/* envs/splendor_env.py:51-90.  obs/mask may be NULL. */
static void o_env_step(OState *s, int action, int32_t *obs, int8_t *mask, float *reward, uint8_t *terminated,
                       uint8_t *info) {
	int8_t m[NUM_ACTIONS];
	uint32_t inf = 0;
	float r = 0.0f;
	int term = 0;
	if (o_is_terminal(s)) { /* RuntimeError in the reference (:53-54): state untouched, flagged */
		inf = INFO_ERROR | INFO_TERMINATED;
		term = 1;
		if (obs) o_encode_observation(s, obs);
		if (mask) memset(mask, 0, NUM_ACTIONS);
		goto out;
	}
	o_legal_moves(s, m);
	int any = 0;
	for (int i = 0; i < NUM_ACTIONS; i++) any |= m[i];
	if (!any) { /* :55-61 */
		s->game_over = 1;
		s->winner = -1;
		s->to_play = 0;
		if (obs) o_encode_observation(s, obs);
		if (mask) memset(mask, 0, NUM_ACTIONS);
		inf = INFO_NOLEGAL_DRAW | INFO_TERMINATED;
		term = 1;
		goto out;
	}
	if (action < 0 || action >= NUM_ACTIONS) { /* ValueError :62-63 */
		inf = INFO_ERROR;
		if (obs) o_encode_observation(s, obs);
		if (mask) memcpy(mask, m, NUM_ACTIONS);
		goto out;
	}
	if (m[action] != 1) { /* :64-66 */
		if (obs) o_encode_observation(s, obs);
		if (mask) memcpy(mask, m, NUM_ACTIONS);
		r = -0.01f;
		inf = INFO_ILLEGAL;
		goto out;
	}
	o_apply_action(s, action);
	if (obs) o_encode_observation(s, obs);
	term = o_is_terminal(s);
	if (term) { /* :71-80 */
		int w = s->winner;
		if (w < 0 && s->turn_limit)
			r = -0.1f;
		else {
			int mover = (s->to_play + 1) % 2; /* (to_play - 1) % num_players */
			r = (w < 0) ? 0.0f : (w == mover ? 1.0f : -1.0f);
		}
		inf |= INFO_TERMINATED;
		if (s->turn_limit) inf |= INFO_TURN_LIMIT;
		inf |= (uint32_t)(s->winner + 1) << INFO_WINNER_SHIFT;
		if (mask) memset(mask, 0, NUM_ACTIONS);
	} else if (mask) {
		o_legal_moves(s, mask);
	}
out:
	*reward = r;
	*terminated = (uint8_t)term;
	*info = (uint8_t)inf;
}

/* canonical flat row (layout: include/splendor_b200.h SPL_ROW_*) */
static void o_export_row(const OState *s, int32_t *row) {
	for (int i = 0; i < ROW_LEN; i++) row[i] = -1;
	for (int i = 0; i < 6; i++) row[i] = s->bank[i];
	for (int p = 0; p < 2; p++) {
		const OPlayer *q = &s->pl[p];
		int o = 6 + 23 * p;
		for (int i = 0; i < 6; i++) row[o + i] = q->tokens[i];
		for (int i = 0; i < 5; i++) row[o + 6 + i] = q->bonuses[i];
		row[o + 11] = q->prestige;
		row[o + 12] = q->n_reserved;
		for (int i = 0; i < 3; i++) row[o + 13 + i] = i < q->n_reserved ? q->reserved[i] : -1;
		for (int i = 0; i < 3; i++) row[o + 16 + i] = i < q->n_reserved ? q->revealed[i] : 0;
		row[o + 19] = q->n_nobles;
		for (int i = 0; i < 3; i++) row[o + 20 + i] = i < q->n_nobles ? q->nobles[i] : -1;
	}
	for (int t = 0; t < 3; t++)
		for (int k = 0; k < 4; k++) row[52 + t * 4 + k] = s->board[t][k];
	for (int t = 0; t < 3; t++) row[64 + t] = s->deck_n[t];
	for (int i = 0; i < 3; i++) row[67 + i] = s->nobles[i];
	row[70] = s->to_play;
	row[71] = s->turn_count;
	row[72] = s->move_count;
	row[73] = s->game_over;
	row[74] = s->winner;
	row[75] = s->turn_limit;
	static const int off[3] = {76, 116, 146};
	for (int t = 0; t < 3; t++)
		for (int k = 0; k < s->deck_n[t]; k++) row[off[t] + k] = s->deck[t][k];
}

static void o_import_row(OState *s, const int32_t *row) {
	memset(s, 0, sizeof(*s));
	for (int i = 0; i < 6; i++) s->bank[i] = row[i];
	for (int p = 0; p < 2; p++) {
		OPlayer *q = &s->pl[p];
		int o = 6 + 23 * p;
		for (int i = 0; i < 6; i++) q->tokens[i] = row[o + i];
		for (int i = 0; i < 5; i++) q->bonuses[i] = row[o + 6 + i];
		q->prestige = row[o + 11];
		q->n_reserved = row[o + 12];
		for (int i = 0; i < 3; i++) q->reserved[i] = row[o + 13 + i];
		for (int i = 0; i < 3; i++) q->revealed[i] = row[o + 16 + i];
		q->n_nobles = row[o + 19];
		for (int i = 0; i < 3; i++) q->nobles[i] = row[o + 20 + i];
	}
	for (int t = 0; t < 3; t++)
		for (int k = 0; k < 4; k++) s->board[t][k] = row[52 + t * 4 + k];
	for (int t = 0; t < 3; t++) s->deck_n[t] = row[64 + t];
	for (int i = 0; i < 3; i++) s->nobles[i] = row[67 + i];
	s->to_play = row[70];
	s->turn_count = row[71];
	s->move_count = row[72];
	s->game_over = row[73];
	s->winner = row[74];
	s->turn_limit = row[75];
	static const int off[3] = {76, 116, 146};
	for (int t = 0; t < 3; t++)
		for (int k = 0; k < s->deck_n[t]; k++) s->deck[t][k] = row[off[t] + k];
}

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11, "Parallel random numbers: as easy as 1, 2, 3"), restated
 * from the paper.  Used ONLY to reproduce the benchmark's uniform-random-legal action stream
 * (SURVEY.md section 8d, config 2): a = kth_set_bit(mask, philox(key,ctr=(env,t)).x % popcount).
 * ------------------------------------------------------------------------------------------ */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
	for (int r = 0; r < 10; r++) {
		uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
		uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
		uint32_t n1 = (uint32_t)p1;
		uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
		uint32_t n3 = (uint32_t)p0;
		c[0] = n0;
		c[1] = n1;
		c[2] = n2;
		c[3] = n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
}

static int o_random_legal_action(const int8_t *mask, uint64_t key, uint64_t env, uint64_t t) {
	int n = 0;
	for (int i = 0; i < NUM_ACTIONS; i++) n += mask[i] != 0;
	if (n == 0) return 0; /* wrappers/selfplay.py:66-73: 0 when nothing is legal */
	uint32_t c[4] = {(uint32_t)env, (uint32_t)(env >> 32), (uint32_t)t, (uint32_t)(t >> 32)};
	philox4x32_10(c, (uint32_t)key, (uint32_t)(key >> 32));
	int k = (int)(c[0] % (uint32_t)n);
	for (int i = 0; i < NUM_ACTIONS; i++)
		if (mask[i]) {
			if (k == 0) return i;
			k--;
		}
	return 0;
}

/* ------------------------------------------------------------------------------------------
 * Exported C interface (ctypes)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
	int64_t n;
	OState *st;
	uint32_t *episode;
	int8_t *cur_mask; /* mask of the current state (what the caller would hold) */
	uint64_t seed_base;
	uint64_t env_offset;
	int64_t stats[8];
} OVec;

/* per-episode engine seed schedule shared with the CUDA path */
uint64_t orc_engine_seed(uint64_t seed_base, uint64_t global_env, uint64_t episode) {
	return (seed_base + 1000003ull * episode + global_env) % 2147483647ull;
}

void *orc_vec_create(int64_t n, uint64_t seed_base, uint64_t env_offset) {
	OVec *v = (OVec *)calloc(1, sizeof(OVec));
	v->n = n;
	v->st = (OState *)calloc((size_t)n, sizeof(OState));
	v->episode = (uint32_t *)calloc((size_t)n, sizeof(uint32_t));
	v->cur_mask = (int8_t *)calloc((size_t)n * NUM_ACTIONS, 1);
	v->seed_base = seed_base;
	v->env_offset = env_offset;
	return v;
}

void orc_vec_destroy(void *h) {
	OVec *v = (OVec *)h;
	free(v->st);
	free(v->episode);
	free(v->cur_mask);
	free(v);
}

/* reset every env (reset_mask NULL) or the flagged ones; engine seed = explicit seeds[i] if seeds
 * is non-NULL, else the schedule orc_engine_seed(seed_base, env_offset+i, episode[i]). */
void orc_vec_reset(void *h, const uint64_t *seeds, const uint8_t *reset_mask, int32_t *obs, int8_t *mask) {
	OVec *v = (OVec *)h;
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < v->n; i++) {
		if (reset_mask && !reset_mask[i]) continue;
		uint64_t sd = seeds ? seeds[i] : orc_engine_seed(v->seed_base, v->env_offset + (uint64_t)i, v->episode[i]);
		o_initial_state(&v->st[i], sd);
		o_legal_moves(&v->st[i], v->cur_mask + i * NUM_ACTIONS);
		if (obs) o_encode_observation(&v->st[i], obs + i * OBS_DIM);
		if (mask) memcpy(mask + i * NUM_ACTIONS, v->cur_mask + i * NUM_ACTIONS, NUM_ACTIONS);
	}
}

/* one lock-step of SplendorEnv.step over all envs; autoreset != 0 gives the same-step auto-reset of
 * ppo_splendor.py:245-250 (reward/terminated/info of the finished episode, obs/mask of the new one). */
void orc_vec_step(void *h, const int32_t *actions, const uint8_t *active, int autoreset, int32_t *obs, int8_t *mask,
                  float *reward, uint8_t *terminated, uint8_t *info) {
	OVec *v = (OVec *)h;
	int64_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma omp parallel for schedule(static) reduction(+ : st[:8])
	for (int64_t i = 0; i < v->n; i++) {
		if (active && !active[i]) continue;
		OState *s = &v->st[i];
		int8_t *cm = v->cur_mask + i * NUM_ACTIONS;
		uint8_t inf;
		o_env_step(s, actions[i], obs ? obs + i * OBS_DIM : NULL, cm, &reward[i], &terminated[i], &inf);
		if (terminated[i] && !(inf & INFO_ERROR)) {
			st[0] += 1;
			if (inf & INFO_NOLEGAL_DRAW)
				st[5] += 1;
			else if (inf & INFO_TURN_LIMIT)
				st[4] += 1;
			else if (s->winner == 0)
				st[1] += 1;
			else if (s->winner == 1)
				st[2] += 1;
			else
				st[3] += 1;
			st[6] += s->move_count;
			if (s->winner >= 0) st[7] += s->pl[s->winner].prestige;
		}
		if (terminated[i] && autoreset) {
			v->episode[i] += 1;
			o_initial_state(s, orc_engine_seed(v->seed_base, v->env_offset + (uint64_t)i, v->episode[i]));
			o_legal_moves(s, cm);
			if (obs) o_encode_observation(s, obs + i * OBS_DIM);
			inf |= INFO_RESET;
		}
		info[i] = inf;
		if (mask) memcpy(mask + i * NUM_ACTIONS, cm, NUM_ACTIONS);
	}
	for (int k = 0; k < 8; k++) v->stats[k] += st[k];
}

void orc_vec_stats(void *h, int64_t *out) { memcpy(out, ((OVec *)h)->stats, sizeof(int64_t) * 8); }

void orc_vec_observe(void *h, int32_t *obs, int8_t *mask) {
	OVec *v = (OVec *)h;
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < v->n; i++) {
		if (obs) o_encode_observation(&v->st[i], obs + i * OBS_DIM);
		if (mask) {
			if (o_is_terminal(&v->st[i]))
				memset(mask + i * NUM_ACTIONS, 0, NUM_ACTIONS);
			else
				o_legal_moves(&v->st[i], mask + i * NUM_ACTIONS);
		}
	}
}

void orc_vec_export(void *h, int32_t *rows) {
	OVec *v = (OVec *)h;
	for (int64_t i = 0; i < v->n; i++) o_export_row(&v->st[i], rows + i * ROW_LEN);
}

void orc_vec_import(void *h, const int32_t *rows, const uint8_t *which) {
	OVec *v = (OVec *)h;
	for (int64_t i = 0; i < v->n; i++) {
		if (which && !which[i]) continue;
		o_import_row(&v->st[i], rows + i * ROW_LEN);
		if (o_is_terminal(&v->st[i]))
			memset(v->cur_mask + i * NUM_ACTIONS, 0, NUM_ACTIONS);
		else
			o_legal_moves(&v->st[i], v->cur_mask + i * NUM_ACTIONS);
	}
}

void orc_vec_episodes(void *h, uint32_t *out) { memcpy(out, ((OVec *)h)->episode, sizeof(uint32_t) * ((OVec *)h)->n); }

/* uniform-random-legal action per env for lock-step index t (same stream as spl_random_action) */
void orc_vec_random_actions(void *h, uint64_t key, uint64_t t, int32_t *actions) {
	OVec *v = (OVec *)h;
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < v->n; i++)
		actions[i] = o_random_legal_action(v->cur_mask + i * NUM_ACTIONS, key, v->env_offset + (uint64_t)i, t);
}

/* CPU baseline driver: `steps` lock-steps of random-legal play with same-step auto-reset, full
 * step+mask+obs work per env-step (obs is encoded into a per-thread scratch row).  Returns the number
 * of env-steps executed. */
int64_t orc_vec_rollout_random(void *h, uint64_t key, uint64_t t0, int64_t steps) {
	OVec *v = (OVec *)h;
	int64_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma omp parallel for schedule(static) reduction(+ : st[:8])
	for (int64_t i = 0; i < v->n; i++) {
		int32_t obs[OBS_DIM];
		OState *s = &v->st[i];
		int8_t *cm = v->cur_mask + i * NUM_ACTIONS;
		for (int64_t t = 0; t < steps; t++) {
			int a = o_random_legal_action(cm, key, v->env_offset + (uint64_t)i, (uint64_t)(t0 + t));
			float r;
			uint8_t term, inf;
			o_env_step(s, a, obs, cm, &r, &term, &inf);
			if (term) {
				st[0] += 1;
				if (inf & INFO_NOLEGAL_DRAW)
					st[5] += 1;
				else if (inf & INFO_TURN_LIMIT)
					st[4] += 1;
				else if (s->winner == 0)
					st[1] += 1;
				else if (s->winner == 1)
					st[2] += 1;
				else
					st[3] += 1;
				st[6] += s->move_count;
				if (s->winner >= 0) st[7] += s->pl[s->winner].prestige;
				v->episode[i] += 1;
				o_initial_state(s, orc_engine_seed(v->seed_base, v->env_offset + (uint64_t)i, v->episode[i]));
				o_legal_moves(s, cm);
				o_encode_observation(s, obs);
			}
		}
	}
	for (int k = 0; k < 8; k++) v->stats[k] += st[k];
	return v->n * steps;
}

/* --- single-function entry points for known-answer tests --- */
void orc_initial_row(uint64_t seed, int32_t *row) {
	OState s;
	o_initial_state(&s, seed);
	o_export_row(&s, row);
}

void orc_row_legal_moves(const int32_t *row, int8_t *mask) {
	OState s;
	o_import_row(&s, row);
	o_legal_moves(&s, mask);
}

void orc_row_encode(const int32_t *row, int32_t *obs) {
	OState s;
	o_import_row(&s, row);
	o_encode_observation(&s, obs);
}

int orc_row_apply(const int32_t *row, int action, int32_t *row_out) {
	OState s;
	o_import_row(&s, row);
	int rc = o_apply_action(&s, action);
	o_export_row(&s, row_out);
	return rc;
}

void orc_row_env_step(const int32_t *row, int action, int32_t *row_out, int32_t *obs, int8_t *mask, float *reward,
                      uint8_t *terminated, uint8_t *info) {
	OState s;
	o_import_row(&s, row);
	o_env_step(&s, action, obs, mask, reward, terminated, info);
	o_export_row(&s, row_out);
}

/* first `n` MT19937 outputs of random.Random(seed) -- used to pin the MT restatement */
void orc_mt_outputs(uint64_t seed, int n, uint32_t *out) {
	MT m;
	mt_seed_u64(&m, seed);
	for (int i = 0; i < n; i++) out[i] = mt_next(&m);
}

void orc_philox(uint32_t *ctr4, uint32_t k0, uint32_t k1) { philox4x32_10(ctr4, k0, k1); }

int orc_num_threads(void) {
#ifdef _OPENMP
	extern int omp_get_max_threads(void);
	return omp_get_max_threads();
#else
	return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
	extern void omp_set_num_threads(int);
	omp_set_num_threads(n);
#else
	(void)n;
#endif
}
