/*
 * splendor_b200.h -- C ABI of the B200-native batched Splendor engine (libsplendor_b200.so).
 *
 * The reference (YiyangShao/splendor-gym) is pure Python and has no FFI layer; the boundary this
 * library replaces is the pair of Python APIs
 *     splendor_gym/envs/splendor_env.py:41-90   SplendorEnv.reset / SplendorEnv.step
 *     splendor_gym/engine/rules.py:40-93,196-287 legal_moves / apply_action
 *     splendor_gym/engine/encode.py:124-187      encode_observation
 * evaluated for N independent environments in lock-step (what gym.vector.SyncVectorEnv and the
 * per-env loop of ppo_splendor.py:235-250 do serially).  Each entry point below names the
 * reference interface it stands in for.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates; the library keeps
 *    only per-device constant tables);  all work is enqueued on `stream` and is asynchronous;
 *  - return value 0 = ok, >0 = cudaError_t, <0 = SPL_E_* argument error; nothing throws or exits;
 *  - one process per GPU, one calling thread: the entry points keep no per-call state besides the caller's buffers, but the
 *    diagnostics (launch counter, spl_timing_*), the host worker pool and spl_init's per-device tables are process-wide and
 *    unsynchronised.  spl_init() must have run on the CURRENT device (checked on every call).
 */
#ifndef SPLENDOR_B200_H
#define SPLENDOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPL_NUM_ACTIONS 45 /* engine/encode.py:32 TOTAL_ACTIONS */
#define SPL_OBS_DIM 297    /* engine/encode.py:74 OBSERVATION_DIM */
#define SPL_OBS_F16_PITCH 304 /* row pitch of the fp16 policy-input observation (297 rounded up to a multiple of 8) */
#define SPL_ROW_LEN 166    /* flat int32 debug row, see SPL_ROW_* below */
#define SPL_STATE_PLANES 4 /* packed hot state: 4 planes of 16 B per env = 64 B */
#define SPL_DECK_STRIDE 96 /* bytes per env of deck order (90 used: tier1[40] tier2[30] tier3[20]) */
#define SPL_RET_TABLE_LEN 8910
#define SPL_MAX_SPARE_SLOTS 16

/* info byte written per env by spl_step (envs/splendor_env.py:56-88 info dict, flattened) */
#define SPL_INFO_ILLEGAL 1u      /* info["illegal_action"]: reward -0.01, state unchanged (:64-66) */
#define SPL_INFO_NOLEGAL_DRAW 2u /* info["draw"]: no legal move -> draw (:55-61) */
#define SPL_INFO_TURN_LIMIT 4u   /* info["turn_limit"] (:84-85) */
#define SPL_INFO_TERMINATED 8u
#define SPL_INFO_WINNER_SHIFT 4 /* bits 4-5: 0 = None, 1 = player 0, 2 = player 1 (final_rewards derive from it) */
#define SPL_INFO_WINNER_MASK 0x30u
#define SPL_INFO_ERROR 64u /* the reference would raise: action out of range (ValueError :62-63) or, with
                              TERMINATED also set, step after termination (RuntimeError :53-54); state unchanged */
#define SPL_INFO_RESET 128u /* same-step auto-reset happened: obs/mask belong to the new episode */

/* deck shuffles at reset (engine/state.py:181-195) */
#define SPL_SHUFFLE_MT19937 0 /* bit-exact with initial_state(seed): CPython random.Random(seed).shuffle */
#define SPL_SHUFFLE_PHILOX 1  /* native counter-based stream (distribution-equivalent, not bit-equal) */

/* stats[8] accumulated (int64 atomics) per finished episode */
#define SPL_STAT_EPISODES 0
#define SPL_STAT_P0_WINS 1
#define SPL_STAT_P1_WINS 2
#define SPL_STAT_TIE_DRAWS 3
#define SPL_STAT_LIMIT_DRAWS 4
#define SPL_STAT_NOLEGAL_DRAWS 5
#define SPL_STAT_SUM_MOVES 6
#define SPL_STAT_SUM_WINNER_PRESTIGE 7

/* reward constants of envs/splendor_env.py:61-80 as codes (compact device->host record, spl_host_step) */
#define SPL_REWARD_CODE_ZERO 0u    /*  0.0  */
#define SPL_REWARD_CODE_WIN 1u     /* +1.0  */
#define SPL_REWARD_CODE_LOSS 2u    /* -1.0  */
#define SPL_REWARD_CODE_LIMIT 3u   /* -0.1  turn-limit draw (:75-76) */
#define SPL_REWARD_CODE_ILLEGAL 4u /* -0.01 illegal action (:64-66) */

#define SPL_E_BADARG (-1)
#define SPL_E_NOTINIT (-2)
#define SPL_E_ALIGN (-3)
#define SPL_E_BADROW (-4) /* spl_import_state: a row outside the packed state's domain; such rows are skipped, not truncated */

/* Flat int32 state row (export/import, parity checks): mirrors the reference dataclasses
 * (engine/state.py:52-87) field by field; -1 = None / absent.
 *   [0:6] bank | player p at 6+23p: [+0:6] tokens [+6:11] bonuses [+11] prestige [+12] len(reserved)
 *   [+13:16] reserved card ids [+16:19] revealed flags [+19] len(nobles) [+20:23] noble indices |
 *   [52:64] board ids tier-major | [64:67] deck sizes | [67:70] visible nobles | [70] to_play
 *   [71] turn_count [72] move_count [73] game_over [74] winner_index [75] turn_limit_reached
 *   [76:116] deck tier 1, [116:146] tier 2, [146:166] tier 3 (bottom..top, -1 beyond the size) */
#define SPL_ROW_BANK 0
#define SPL_ROW_PLAYER0 6
#define SPL_ROW_PLAYER_STRIDE 23
#define SPL_ROW_BOARD 52
#define SPL_ROW_DECK_SIZES 64
#define SPL_ROW_NOBLES 67
#define SPL_ROW_TO_PLAY 70
#define SPL_ROW_TURN_COUNT 71
#define SPL_ROW_MOVE_COUNT 72
#define SPL_ROW_GAME_OVER 73
#define SPL_ROW_WINNER 74
#define SPL_ROW_TURN_LIMIT 75
#define SPL_ROW_DECK1 76
#define SPL_ROW_DECK2 116
#define SPL_ROW_DECK3 146

/* The environments of one shard: structure-of-arrays game state in HBM (replaces the per-env
 * SplendorState dataclass, engine/state.py:74-104). */
typedef struct spl_envs {
	void *state;         /* uint4[SPL_STATE_PLANES][stride]: packed 64-B hot rows, plane-major, 16-B aligned */
	uint8_t *decks;      /* [n][SPL_DECK_STRIDE] deck order (cold; written at reset, 1 B read per refill) */
	uint32_t *episode;   /* [n] episodes started per env (drives the per-episode seed schedule) */
	int32_t *scratch;    /* [n + 4] work list of envs to auto-reset (library-internal use) */
	int64_t stride;      /* envs per plane (>= n) */
	int64_t n;           /* environments in this shard */
	uint64_t env_offset; /* global id of env 0 (multi-GPU sharding: seeds depend on the global id only) */
	uint64_t seed_base;  /* engine seed of (global env g, episode e) = (seed_base + 1000003 e + g) mod (2^31-1) */
	int32_t shuffle_mode; /* SPL_SHUFFLE_* */
	int32_t spare_slots; /* S = deals kept ahead per env in `spare` (0 or 1 = one; at most SPL_MAX_SPARE_SLOTS) */
	uint8_t *spare;      /* nullable, SPL_SHUFFLE_MT19937 only: [n][S][SPL_DECK_STRIDE] prefetched deals of every env's NEXT S
	                        episodes (episode e in slot e % S, tagged with e) followed by int32[n * S + 4] (refill list), i.e.
	                        n * S * 96 + (n * S + 4) * 4 bytes, 16-byte aligned.  With it the bit-exact auto-reset no longer
	                        waits for one lane's random.Random(seed) chain (~35 us): spl_step takes the prepared deal and
	                        refills the spares in batches; spl_rollout_random (one launch for many lock-steps) takes up to S
	                        per env and refills behind the launch -- a game lasts >= 17 moves, so S = 8 covers 128 lock-steps
	                        (an env that runs out is dealt in place, slowly).  Zero it once. */
	const uint64_t *episode_seeds; /* nullable, SPL_SHUFFLE_MT19937: DEVICE [n][episode_seed_count] engine seeds of every env's
	                        episodes 1 .. episode_seed_count (episode 0 = the one spl_reset started).  Replay of a reference run:
	                        the reference's SplendorEnv draws a fresh engine seed from its own PCG64 stream on every reset
	                        (envs/splendor_env.py:42-43, re-drawn per auto-reset by the vector env, ppo_splendor.py:246-247); with
	                        this table the auto-resets deal exactly those decks.  Later episodes fall back to the schedule. */
	int32_t episode_seed_count;
	int32_t reserved_;
} spl_envs_t;

/* spl_step_io.flags */
#define SPL_IO_ASYNC_REFILL 1 /* prefetched deals (envs->spare): this call does not refill the rings; the caller runs
                                 spl_refill_spares itself -- possibly on another stream, CONCURRENTLY with later step /
                                 rollout launches that carry this flag (they then read a slot's ready flag before its
                                 deck order).  The refill of the deals taken by launch k must have completed before launch
                                 k + 2 starts if no env is to run its ring dry (16 slots cover two 128-step launches). */

/* Outputs of one lock-step (the 5-tuple of SplendorEnv.step, batched). Nullable members are skipped. */
typedef struct spl_step_io {
	const int32_t *actions; /* [n] */
	const uint8_t *active;  /* [n] nullable: 0 = leave this env untouched this call (dual_step phase 2) */
	int32_t *obs;           /* [n][297] nullable */
	int8_t *mask;           /* [n][45]  nullable: info["action_mask"] of the returned observation */
	float *reward;          /* [n] */
	uint8_t *terminated;    /* [n] */
	uint8_t *info;          /* [n] SPL_INFO_* */
	int64_t *stats;         /* [8] nullable */
	int32_t *next_action;   /* [n] nullable: uniform random legal action for the returned mask (fused sampler) */
	uint64_t action_key;    /* Philox key / lock-step counter for next_action */
	uint64_t action_t;
	const uint64_t *action_t_base; /* nullable device scalar added to action_t (lets a captured CUDA graph advance
	                                  the sampler's counter between replays) */
	int32_t autoreset; /* same-step auto-reset (ppo_splendor.py:245-250) */
	int32_t flags;     /* SPL_IO_* */
	/* policy-ready observation formats (spl_step only; `obs` must then be NULL and auto-reset needs SPL_SHUFFLE_PHILOX or
	 * SPL_SHUFFLE_MT19937 with prefetched deals, envs->spare).
	 * The reference's caller casts the int32 observation to float for the MLP (ppo_splendor.py:221); here the cast is
	 * fused into the step kernel: */
	void *obs_f16;   /* [n][SPL_OBS_F16_PITCH] fp16, nullable: entries 0..296 = the observation (exact: all < 256),
	                    297..303 = 0, so that a row is 16-byte aligned and the first Linear has K % 8 == 0 */
	uint8_t *obs_u8; /* [n][297] nullable: the same observation as bytes (compact rollout buffers) */
} spl_step_io_t;

/* Upload the card / noble / token-return tables to the current device. Idempotent. */
int spl_init(void);

/* SplendorEnv.reset (envs/splendor_env.py:41-49) for every env, or those with reset_mask[i] != 0.
 * Engine seeds: explicit `seeds[i]` if non-NULL, else the schedule above (episode[i] is zeroed on a
 * full reset and incremented on a masked one).  Writes obs / mask if non-NULL. */
int spl_reset(const spl_envs_t *envs, const uint64_t *seeds, const uint8_t *reset_mask, int32_t *obs, int8_t *mask,
              void *stream);

/* SplendorEnv.step (envs/splendor_env.py:51-90) for every env in lock-step:
 * legality check -> apply_action -> encode_observation -> legal_moves [-> same-step auto-reset]. */
int spl_step(const spl_envs_t *envs, const spl_step_io_t *io, void *stream);

/* `steps` lock-steps of uniform-random-legal play with same-step auto-reset in ONE launch -- the batched
 * equivalent of scripts/random_rollout.py:13-30 (and of a random-policy rollout collection,
 * ppo_splendor.py:219-297).  Outputs are step-major rollout buffers: obs [steps][n][297], mask
 * [steps][n][45], reward / terminated / info [steps][n]; io->actions = actions of the first step [n];
 * io->next_action = [steps+1][n] (row 0 is left untouched, row t+1 = action chosen after step t).
 * Results are bit-identical to `steps` chained spl_step calls with action_t, action_t+1, ...
 * Requires autoreset and either SPL_SHUFFLE_PHILOX or SPL_SHUFFLE_MT19937 with prefetched deals (envs->spare). */
int spl_rollout_random(const spl_envs_t *envs, const spl_step_io_t *io, int32_t steps, void *stream);

/* Replace, now, every prefetched deal that has been taken since the last refill (envs->spare, SPL_SHUFFLE_MT19937).
 * spl_step / spl_host_step do this themselves on every call whose io->action_t is a multiple of 16 x min(spare_slots, 4)
 * (pass the lock-step counter there), spl_rollout_random behind every launch; a caller that cannot keep that cadence -- e.g.
 * one that replays a captured single-step CUDA graph, whose action_t is frozen -- calls this every <= 16 x spare_slots
 * lock-steps instead.  It may run concurrently with step / rollout launches that carry SPL_IO_ASYNC_REFILL (the rows are
 * published flag-last and taken flag-first); otherwise order it with them (same stream, or an event).  An env that finds
 * its ring empty is dealt in place by the step kernel, bit-identically but slowly (one lane's random.Random(seed), ~35 us). */
int spl_refill_spares(const spl_envs_t *envs, void *stream);

/* Replay of caller-supplied deck permutations (north_star: "a replay mode accepts the reference's deck permutations").
 * deals: [n][count][SPL_DECK_STRIDE] bytes, HOST or DEVICE memory; deals[e][k] is the deal of env e's (current episode + 1
 * + k)-th episode, count <= spare_slots.  A deal row is what initial_state() shuffles (engine/state.py:181-211):
 *   bytes 0..39 the tier-1 card ids in list order (the LAST four are dealt to board slots 0..3: byte 39 -> slot 0, ...,
 *   byte 36 -> slot 3; deck.pop() continues from byte 35 downwards), 40..69 tier 2, 70..89 tier 3, 90..92 the three visible
 *   noble indices; 93..95 are ignored.  Rows are validated (each tier a permutation of its own cards, nobles distinct);
 * returns SPL_E_BADROW if any is not, having loaded none.  The rows go into the ring of prefetched deals (envs->spare):
 * the next `count` auto-resets of every env use them; refills afterwards follow episode_seeds / the seed schedule. */
int spl_load_deals(const spl_envs_t *envs, const uint8_t *deals, int32_t count, void *stream);

/* How spl_rollout_random would run `steps` lock-steps of n envs on the current device (measurement aid):
 * out[0] warps per CTA, [1] CTAs, [2] lock-steps per work unit, [3] work units per tile group, [4] CTA barrier per
 * lock-step, [5] tile groups.  Each work unit reloads / stores the packed state of its envs (128 B per env). */
int spl_rollout_plan(int64_t n, int32_t steps, int32_t *out);

/* encode_observation (engine/encode.py:124-187) + legal_moves (engine/rules.py:40-93) of the current
 * states, without stepping (mask is all-zero for terminal states, as in envs/splendor_env.py:81). */
int spl_observe(const spl_envs_t *envs, int32_t *obs, int8_t *mask, void *stream);
/* the same with the policy-ready formats of spl_step_io (obs_f16 [n][SPL_OBS_F16_PITCH] fp16 and / or obs_u8 [n][297]) */
int spl_observe_policy(const spl_envs_t *envs, void *obs_f16, uint8_t *obs_u8, int8_t *mask, void *stream);

/* random_opponent (wrappers/selfplay.py:66-73) batched: uniform over the legal actions of mask[n][45],
 * a = k-th set bit with k = philox4x32-10(key, ctr=(global env, t)).x mod popcount; 0 if none legal. */
int spl_random_action(const int8_t *mask, int64_t n, uint64_t env_offset, uint64_t key, uint64_t t, int32_t *actions,
                      void *stream);

/* packed state <-> flat int32 rows [n][SPL_ROW_LEN] (device pointers). Import recomputes nothing else; it validates every
 * row (counters < 128, card ids < 90, noble ids < 10, list lengths <= 3, deck sizes within their tier), skips the rows
 * that fail, waits for the stream and returns SPL_E_BADROW if there were any. */
int spl_export_state(const spl_envs_t *envs, int32_t *rows, void *stream);
int spl_import_state(const spl_envs_t *envs, const int32_t *rows, const uint8_t *which, void *stream);

/* Self-play turn bookkeeping: combines phase-1 (agent = player 0) and phase-2 (opponent) step results into
 * agent_reward / opp_reward / done.  mode 0 = DualStepNativeWrapper.dual_step (wrappers/dual_step_native.py:132-167,
 * 195-201; also DualStepSelfPlayWrapper, wrappers/dual_step_selfplay.py:138-152): agent reward = final_rewards[0];
 * mode 1 = SelfPlayWrapper.step (wrappers/selfplay.py:42-63): agent reward = -opponent reward. */
int spl_dual_combine(const float *r1, const uint8_t *term1, const uint8_t *info1, const float *r2, const uint8_t *term2,
                     const uint8_t *info2, int64_t n, float *agent_reward, float *opp_reward, uint8_t *done, int mode,
                     void *stream);

/* ---- callers on either side of the step path, kept on the device ---- */

/* scripted opponents of scripts/eval_suite.py over (obs [n][297], mask [n][45]) */
#define SPL_BOT_RANDOM 0         /* wrappers/selfplay.py:66-73 random_opponent */
#define SPL_BOT_GREEDY_V1 1      /* scripts/eval_suite.py:10-30 greedy_opponent_v1 */
#define SPL_BOT_BASIC_PRIORITY 2 /* scripts/eval_suite.py:33-77 basic_priority_opponent (np.random.choice -> Philox stream) */
#define SPL_BOT_GREEDY_V2 3      /* scripts/eval_suite.py:80-128 greedy_opponent_v2_factory(env_ref) (bank read from obs[0:5]) */
int spl_scripted_action(const int32_t *obs, const int8_t *mask, int64_t n, int kind, uint64_t env_offset, uint64_t key,
                        uint64_t t, int32_t *actions, void *stream);

/* masked categorical over logits [n][45] (ppo_splendor.py:27-38,54-59): mode 0 = sample (Philox inverse CDF),
 * mode 1 = argmax (scripts/eval_suite.py:131-141 model_greedy_policy_from).  logprob / entropy nullable. */
int spl_masked_sample(const float *logits, const int8_t *mask, int64_t n, int mode, uint64_t env_offset, uint64_t key,
                      uint64_t t, int32_t *actions, float *logprob, float *entropy, void *stream);
/* the same over fp16 logits with a row pitch (in elements) >= 45, e.g. the output of a half-precision head padded to 48
 * columns: no float copy of the logits is needed; the arithmetic is done in float on the exactly converted values */
int spl_masked_sample_f16(const void *logits, int64_t pitch, const int8_t *mask, int64_t n, int mode, uint64_t env_offset,
                          uint64_t key, uint64_t t, int32_t *actions, float *logprob, float *entropy, void *stream);

/* generalised advantage estimation over step-major [T][n] buffers (ppo_splendor.py:299-314) */
int spl_gae(const float *rewards, const float *values, const uint8_t *terminals, const float *last_values, int32_t T,
            int64_t n, float gamma, float lam, float *advantages, float *returns, void *stream);

/* ---- host-buffer entry points ---------------------------------------------------------------------
 * The reference's callers live on the host: SplendorEnv.step returns NumPy arrays (envs/splendor_env.py:51-90)
 * and the vector loop stacks them (ppo_splendor.py:235-285).  These calls take HOST pointers for actions and
 * results.  Per lock-step the step kernel reads 4 B/env of actions from pinned host memory and a push kernel stores the
 * results over PCIe in 64-env groups: in a transfer form of 181.5 B/env (observation entries as nibbles + the 17 columns
 * that can exceed 15 as bytes + one 16-byte record: 45 legal-mask bits, reward code, terminated, info, sampled next action)
 * for the share that pinned host threads widen into the reference-typed arrays (uint8 -> int32 observation, bits -> int8
 * mask, code -> float reward), and ALREADY WIDENED (1,243 B/env) for the rest when the result arrays are GPU-writable
 * (spl_host_alloc).  Results in the caller's buffers are exactly those of spl_step / spl_observe. */
typedef struct spl_host spl_host_t; /* opaque: compact device buffers, pinned staging rings + arrival flags, the split's state */

typedef struct spl_host_io {
	const int32_t *actions; /* [n] host (ignored by spl_host_observe) */
	int32_t *obs;           /* [n][297] host, nullable */
	uint8_t *obs_u8;        /* [n][297] host, nullable: the same observation as bytes (all entries < 256) */
	int8_t *mask;           /* [n][45] host, nullable */
	float *reward;          /* [n] host, nullable */
	uint8_t *terminated;    /* [n] host, nullable */
	uint8_t *info;          /* [n] host, nullable: SPL_INFO_* */
	int32_t *next_action;   /* [n] host, nullable: uniform random legal action for the returned mask */
	int64_t *stats;         /* [8] DEVICE, nullable */
	uint64_t action_key;    /* Philox key / lock-step counter of next_action */
	uint64_t action_t;
	int32_t autoreset;
	int32_t reserved_;
} spl_host_io_t;

int spl_host_create(int64_t n, int32_t chunks, spl_host_t **out); /* chunks: ignored since ABI 120 (arrival is tracked per 64-env group) */
int spl_host_destroy(spl_host_t *h);
/* SplendorEnv.step for every env with host buffers; returns when the caller's buffers are filled */
int spl_host_step(spl_host_t *h, const spl_envs_t *envs, const spl_host_io_t *io, void *stream);
/* observation / mask (/ sampled action) of the current states into host buffers (after spl_reset) */
int spl_host_observe(spl_host_t *h, const spl_envs_t *envs, const spl_host_io_t *io, void *stream);
/* the widening step alone, for callers that move the compact records themselves: obs_u8 [n][297] and side [n][16 B]
 * (word 0 = mask bits 0-31; word 1 = mask bits 32-44 | reward code << 16 | terminated << 24; word 2 = info |
 * next action << 8) in HOST memory -> the non-NULL host arrays of `io`.  Pure CPU; needs no device. */
int spl_host_expand(const uint8_t *obs_u8, const void *side, int64_t n, const spl_host_io_t *io);
/* host threads used for widening (n <= 0: query; default: the cores this rank may run on minus two, at least 3/4 of them; SPL_HOST_THREADS
 * overrides); returns the count.  Workers are pinned to distinct cores unless spl_host_set_pinning(0). */
int spl_host_set_threads(int n);
int spl_host_set_pinning(int on);
/* Result arrays the GPU can write directly (pinned + mapped, transparent huge pages requested).  When every non-NULL
 * result pointer of spl_host_io_t lies in such memory (or in any cudaHostAlloc / cudaHostRegister'ed memory) and is
 * 16-byte aligned, spl_host_step lets the GPU write a share of the envs ALREADY WIDENED over PCIe while host threads
 * widen the rest; otherwise host threads widen everything.  Same values either way. */
int spl_host_alloc(size_t bytes, void **out);
int spl_host_free(void *ptr);
/* timings of the last spl_host_step / spl_host_observe of `h` (microseconds from the start of the call):
 * [0] whole call, [1] work enqueued, [2] last worker saw its first group, [3] slowest worker done, [4] GPU-written
 * share landed, [5] share of the envs the GPU wrote widened, [6] worker threads, [7] 1 = result arrays GPU-writable */
#define SPL_HOST_STATS 8
int spl_host_get_stats(const spl_host_t *h, double *out);
/* sustained streaming-store rate of the worker pool in GB/s written (mode 0: fill, 1: uint8 -> int32 widening) over at
 * least `reps` passes and 40 ms: the ceiling the host-buffer path is reported against (bench.py e2e.host_store_gbs).
 * Pure CPU; needs no device. */
double spl_host_store_rate(int64_t bytes_per_thread, int reps, int mode);

const char *spl_error_string(int code);
int spl_version(void);
/* host copy of the token-return table (SPL_RET_TABLE_LEN x u64) for tests */
int spl_host_ret_table(uint64_t *out);
/* number of kernels this library has launched since load (bench.py "gpu_launches") */
int64_t spl_launch_count(void);
/* measurement aid: when enabled, spl_step records CUDA events on the caller's stream around its main
 * kernel; spl_timing_read synchronises and returns the summed duration [ms] and launch count */
int spl_timing_enable(int on);
int spl_timing_read(double *total_ms, int64_t *count);

#ifdef __cplusplus
}
#endif
#endif
