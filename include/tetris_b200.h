/*
 * tetris_b200.h -- C ABI of the B200-native batched Tetris simulator (libtetris_b200.so).
 *
 * This is the drop-in boundary for the hot path of Max-We/Tetris-Gymnasium.  The reference has
 * no FFI of its own (its plugin surface is Python: the gymnasium registry entry
 * "tetris_gymnasium/Tetris" -> tetris_gymnasium/envs/__init__.py:10-14, the constructor kwargs
 * tetris_gymnasium/envs/tetris.py:77-91, the wrapper constructors wrappers/grouped.py:44-49 and
 * wrappers/observation.py:18,140-147, and the functional signatures envs/tetris_fn.py:276-324),
 * so every entry point below names the reference interface it replaces.  The Python host in
 * tetris_gymnasium_b200/ binds these symbols with ctypes and mirrors the reference classes.
 *
 * Conventions
 *   - plain C types only; all `d_*` / state / obs pointers are DEVICE pointers owned by the caller
 *     (PyTorch allocates them; the library never frees them); 16-byte aligned.
 *   - every call is asynchronous on the passed `stream` (a cudaStream_t passed as void*), no
 *     implicit synchronisation; the *_host entry points synchronise before returning.
 *     The step / wrapper kernels are launched with programmatic stream serialization and execute griddepcontrol.wait before
 *     their first access to global memory: stream-order semantics are unchanged, only their prologue may overlap the tail
 *     of the preceding kernel on the stream (TG_NO_PDL=1 selects plain launches).
 *   - return value: 0 = OK, otherwise a TG_ERR_* code; tg_last_error() gives the text.
 *   - invalid *actions* are data errors: the env treats them like the reference's unmatched
 *     elif-chain (no move), it does not abort (reference asserts, envs/tetris.py:215).
 *   - a tg_env handle is not re-entrant; distinct handles are independent and thread-safe.
 *   - every entry point selects the env's device for the duration of the call and restores the caller's current device.
 *   - there is NO CPU fallback: without a CUDA device tg_create fails with TG_ERR_CUDA.
 */
#ifndef TETRIS_B200_H
#define TETRIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_VERSION 2
#define TG_PADDING 4   /* = max tetromino matrix dim, envs/tetris.py:130 */
#define TG_MAX_QUEUE 16
#define TG_N_PIECES 7

enum {
    TG_OK = 0,
    TG_ERR_CONFIG = 1,   /* unsupported width/height/queue_size/... */
    TG_ERR_POINTER = 2,  /* NULL or misaligned pointer */
    TG_ERR_CUDA = 3,     /* CUDA runtime error (text in tg_last_error) */
    TG_ERR_ARG = 4
};

/* autoreset policy of the batched env (gymnasium.vector AutoresetMode; the reference NumPy env
 * itself does not guard stepping after game over, SURVEY Q7 = TG_AUTORESET_DISABLED) */
enum { TG_AUTORESET_DISABLED = 0, TG_AUTORESET_NEXT_STEP = 1, TG_AUTORESET_SAME_STEP = 2 };

/* piece-stream source (replaces components/tetromino_randomizer.py) */
enum {
    TG_RNG_PHILOX = 0,   /* device-native 7-bag: Philox4x32-10 keyed by (seed, global env id) */
    TG_RNG_SEQUENCE = 1, /* injected stream piece_seq[env][cursor++ % seq_len] (parity tests) */
    TG_RNG_NUMPY = 2     /* numpy-exact 7-bag: PCG64 + masked-rejection Fisher-Yates, in-place
                            reshuffle (components/tetromino_randomizer.py:67-91) */
};

/* randomizer kind (components/tetromino_randomizer.py): BagRandomizer :49-102 | TrueRandomizer :105-136
 * (`rng.integers(0, 7)` per draw; with TG_RNG_NUMPY bit-exact through numpy's Lemire bounded integers on
 * PCG64's buffered 32-bit halves, with TG_RNG_PHILOX one Philox4x32-10 block per draw) */
enum { TG_RANDOMIZER_BAG = 0, TG_RANDOMIZER_TRUE = 1 };

/* Constructor options of the reference env (envs/tetris.py:77-91) + mappings
 * (mappings/actions.py:12-19, mappings/rewards.py:12-15) + wrapper options. */
typedef struct tg_config {
    int32_t width;        /* playfield width  W (reference default 10); W + 8 <= 32 */
    int32_t height;       /* playfield height H (reference default 20); H + 4 <= 64 */
    int32_t queue_size;   /* visible queue length (reference TetrominoQueue default 4); 1..16 */
    int32_t gravity;      /* envs/tetris.py:82 */
    int32_t autoreset;    /* TG_AUTORESET_* */
    int32_t rng_mode;     /* TG_RNG_* */
    /* ActionsMapping values in field order: move_left, move_right, move_down, rotate_clockwise,
     * rotate_counterclockwise, hard_drop, swap, no_op.  The elif order of envs/tetris.py:223-256
     * (left,right,down,cw,ccw,swap,hard_drop,no_op; first match wins) is applied by the library. */
    int32_t action_map[8];
    int32_t terminate_on_illegal; /* GroupedActionsObservations(terminate_on_illegal_action) */
    int32_t randomizer;   /* TG_RANDOMIZER_*: 7-bag (BagRandomizer, the reference default) or uniform (TrueRandomizer) */
    double reward_alife;          /* RewardsMapping.alife          (default 1)    */
    double reward_clear_line;     /* RewardsMapping.clear_line     (unused by the env, kept) */
    double reward_game_over;      /* RewardsMapping.game_over      (default 0)    */
    double reward_invalid_action; /* RewardsMapping.invalid_action (default -0.1) */
    int64_t seq_len;              /* TG_RNG_SEQUENCE: entries per env in piece_seq */
    uint64_t env_id_offset;       /* global id of local env 0 (multi-GPU sharding: Philox streams
                                     are keyed by global id so results do not depend on #GPUs) */
    int32_t holder_size;          /* TetrominoHolder(size) (components/tetromino_holder.py:14-21): 1..4; 0 = 1 (the reference default).
                                     A FIFO: a swap stores the active piece and, once the holder is full, hands back the oldest one. */
    /* Tetris(tetrominoes=[...]) (envs/tetris.py:88-89, 113-132): 0 = the reference's seven pieces; else 1..7 custom pieces, piece i
     * an n x n matrix (piece_n[i] <= 4, row-major in piece_matrix[i], non-zero = cell) with EXACTLY FOUR cells and colour
     * piece_color[i].  The reference derives the padding from the set (largest matrix dimension); this library is built on
     * padding 4, so the largest matrix of the set must be 4 x 4.  Cell values on the board are i + 2 as in the reference. */
    int32_t n_pieces;
    uint8_t piece_n[8];
    uint8_t piece_matrix[7][16];
    uint8_t piece_color[7][3];
    uint8_t reserved1[3];
} tg_config;

/* Sizes of the per-env state arrays the caller must allocate (bytes per env). */
typedef struct tg_layout {
    int32_t width_padded;   /* W + 2P (envs/tetris.py:131) */
    int32_t height_padded;  /* H + P  (envs/tetris.py:132) */
    int32_t hot_stride;     /* bytes/env of state.hot   (position, rotation, holder, queue, bag, episode stats) */
    int32_t board_stride;   /* bytes/env of state.board (column occupancy bitboards + nibble-packed piece ids) */
    int32_t rng_stride;     /* bytes/env of state.rng   */
    int32_t obs_board_bytes;  /* H_pad * W_pad          (u8) */
    int32_t obs_holder_bytes; /* P * P                  (u8) */
    int32_t obs_queue_bytes;  /* P * P * queue_size     (u8) */
    int32_t n_placements;   /* 4 * W   (grouped action space, wrappers/grouped.py:57) */
    int32_t n_features;     /* W + 3   (wrappers/observation.py:149-160) */
    int32_t rgb_width;      /* W_pad + max(queue_size, 1) * P (wrappers/observation.py:25-33) */
    int32_t host_record_bytes; /* bytes/env TG_HOST_COMPACT moves over the link: 12 (hot words 0, 2, 3) + nibble id plane, 16-byte rounded */
} tg_layout;

/* Device state of n envs, structure-of-records resident in HBM (caller-allocated). */
typedef struct tg_state {
    void *hot;                /* n * hot_stride   */
    void *board;              /* n * board_stride */
    void *rng;                /* n * rng_stride   */
    const uint8_t *piece_seq; /* TG_RNG_SEQUENCE: u8[n][seq_len], values 0..6; else NULL */
} tg_state;

/* Observation dict of Tetris._get_obs (envs/tetris.py:566-615), all uint8, env-major. */
typedef struct tg_obs {
    uint8_t *board;  /* [n][H_pad][W_pad]  locked board + active piece ids          */
    uint8_t *mask;   /* [n][H_pad][W_pad]  active_tetromino_mask (n x n bounding box) */
    uint8_t *holder; /* [n][P][P*holder_size]  held pieces oldest first, ones in the empty slots (holder_size 1: the reference's
                        array; > 1: fixed shape where the reference's np.hstack of the held pieces is ragged)            */
    uint8_t *queue;  /* [n][P][P*queue_size]                                         */
} tg_obs;

/* Per-step scalar outputs (the 5-tuple of Tetris.step, envs/tetris.py:266-272). */
typedef struct tg_step_out {
    float *reward;       /* [n] */
    uint8_t *terminated; /* [n] */
    uint8_t *truncated;  /* [n] always 0 (envs/tetris.py:219) */
    int32_t *lines;      /* [n] info["lines_cleared"] */
} tg_step_out;

/* Episode statistics accumulated on device (RecordEpisodeStatistics-style; examples/train_lin_grouped.py:148).
 * 4 x double: finished episodes, sum of returns, sum of lengths, sum of cleared lines. */
typedef struct tg_stats {
    double episodes, sum_return, sum_length, sum_lines;
} tg_stats;

typedef struct tg_env tg_env;

/* ---- lifecycle --------------------------------------------------------------------------- */
/* replaces Tetris.__init__ (envs/tetris.py:77-201): validates the config, derives padding and
 * padded sizes, uploads the constant piece tables to the device. */
int tg_create(const tg_config *cfg, int device, tg_env **out);
int tg_destroy(tg_env *env);
int tg_get_layout(const tg_env *env, tg_layout *out);
/* options that may change after construction: GroupedActionsObservations(terminate_on_illegal_action) is a WRAPPER option in the
 * reference (wrappers/grouped.py:44-49), so the wrapper sets it on the handle it wraps */
enum { TG_OPT_TERMINATE_ON_ILLEGAL = 1, TG_OPT_HOST_THREADS = 2 };
int tg_set_option(tg_env *env, int32_t option, int64_t value);
const char *tg_last_error(const tg_env *env); /* env may be NULL: error of the last failed tg_create */
int tg_version(void);

/* ---- base env ---------------------------------------------------------------------------- */
/* replaces Tetris.reset (envs/tetris.py:274-307) + TetrominoQueue.reset + BagRandomizer.reset.
 *   seeds      : NULL, or u64[n]  (TG_RNG_PHILOX: stream key; TG_RNG_NUMPY: ignored, use tg_seed_numpy)
 *   reset_mask : NULL (= all), or u8[n]; envs with 0 keep their state (their obs is still written)
 *   obs        : observation dict written for all n envs                                        */
int tg_reset(tg_env *env, tg_state st, int64_t n, const uint64_t *d_seeds, const uint8_t *d_reset_mask,
             tg_obs obs, void *stream);

/* TG_RNG_NUMPY: load PCG64 states; d_pcg = u64[n][4] {state_hi, state_lo, inc_hi, inc_lo} of
 * PCG64(SeedSequence(seed)) (Randomizer.reset, components/tetromino_randomizer.py:40-43). */
int tg_seed_numpy(tg_env *env, tg_state st, int64_t n, const uint64_t *d_pcg, const uint8_t *d_mask, void *stream);

/* Same seeding from the raw seeds: d_seeds = u64[n]; env i gets PCG64(SeedSequence(d_seeds[i])) -- numpy's SeedSequence hash
 * and pcg_setseq_128_srandom_r run on the device (`np.random.default_rng(seed)`, components/tetromino_randomizer.py:40-43). */
int tg_seed_numpy_seeds(tg_env *env, tg_state st, int64_t n, const uint64_t *d_seeds, const uint8_t *d_mask, void *stream);

/* replaces Tetris.step (envs/tetris.py:203-272) for n envs, observation dict written every call
 * (pass an all-NULL tg_obs to skip the dict, e.g. when only tg_render_rgb / tg_features output is consumed).
 * d_stats may be NULL. */
int tg_step(tg_env *env, tg_state st, int64_t n, const int32_t *d_actions, tg_obs obs, tg_step_out out,
            tg_stats *d_stats, void *stream);

/* K consecutive Tetris.step calls in one launch: d_actions = i32[k_steps][n].  The observation dict / 5-tuple arrays of step k
 * start `k * obs_stride` / `k * out_stride` ENVS behind the passed pointers (stride n = [K][n] rollout storage; stride 0 = one
 * set of arrays: the 5-tuple of every step overwrites the previous one and only the last step's dict is written).  Batches whose
 * packed records fit in shared memory (up to ~10^5 envs at 10x20) run as ONE persistent launch with the records resident on chip
 * for all K steps; larger ones as K back-to-back launches.  Same results as K tg_step calls. */
int tg_step_n(tg_env *env, tg_state st, int64_t n, int32_t k_steps, const int32_t *d_actions, tg_obs obs, int64_t obs_stride,
              tg_step_out out, int64_t out_stride, tg_stats *d_stats, void *stream);

/* Same call with HOST buffers (the reference's own calling convention: numpy arrays in, numpy arrays out --
 * Tetris.step returns the dict of envs/tetris.py:566-615 in host memory).  Actions are copied H2D, the envs are stepped, the
 * observation dict and the 5-tuple arrive in the caller's host arrays; synchronous.  `stream` is the stream the caller's
 * earlier work on this state (tg_reset, tg_step, tg_set_state ...) was enqueued on: the internal copy streams wait for it.
 *   TG_HOST_DMA     the dict is produced on the device and DMA-copied (2*Hp*Wp + 16 + 16Q bytes per env over the link);
 *   TG_HOST_COMPACT the step runs without the dict; what the dict is a function of -- per env hot words 0, 2, 3 and the nibble
 *                   id plane, tg_layout.host_record_bytes (112 B at 10x20 against 992 B of dict) -- is packed on the device
 *                   and crosses the link chunk by chunk, and a pool of host threads rebuilds the dict in the caller's arrays
 *                   with streaming stores while later chunks are still stepping (format conversion only -- no game logic runs
 *                   on the host).  Arrays aligned to 64 bytes take the fast path.  Both modes produce identical bytes. */
enum { TG_HOST_DMA = 0, TG_HOST_COMPACT = 1 };
int tg_step_host(tg_env *env, tg_state st, int64_t n, const int32_t *h_actions, tg_obs h_obs, tg_step_out h_out,
                 int32_t mode, void *stream);
/* host threads of the TG_HOST_COMPACT expansion; 0 = the cores this process may run on / LOCAL_WORLD_SIZE (TG_HOST_THREADS) */
int tg_set_host_threads(tg_env *env, int32_t threads);
/* timing of the last tg_step_host call: seconds total, waiting for the device, expanding; number of chunks */
int tg_host_stats(tg_env *env, double *out4);
/* The expansion alone, no device involved: packed records in HOST memory (hot u8[n][32], board u8[n][board_stride] as read
 * back from tg_state) -> observation dict in host arrays. */
int tg_host_expand(const tg_config *cfg, int64_t n, const void *h_hot, const void *h_board, tg_obs h_obs, int32_t threads);
/* measurement aid: GB/s of `threads` host threads filling h_dst (64-byte aligned) with streaming stores, `reps` passes --
 * the ceiling of the TG_HOST_COMPACT expansion's output traffic on this host */
int tg_host_membw(void *h_dst, int64_t bytes, int32_t threads, int32_t reps, double *gb_per_s);

/* ---- wrappers ---------------------------------------------------------------------------- */
/* replaces FeatureVectorObservation.observation (wrappers/observation.py:238-278) applied to the
 * base env's observation: d_feats = u8[n][W+3] = heights(W), max height, holes, bumpiness. */
int tg_features(tg_env *env, tg_state st, int64_t n, uint8_t *d_feats, void *stream);

/* replaces RgbObservation.observation (wrappers/observation.py:38-74): u8[n][H_pad][rgb_width][3] */
int tg_render_rgb(tg_env *env, tg_state st, int64_t n, uint8_t *d_img, void *stream);

/* Fused image adapter of the CNN trainer (examples/train_cnn.py:127-147):
 *     RgbObservation -> gym.wrappers.ResizeObservation((out_h, out_w)) -> gym.wrappers.GrayscaleObservation
 * (cv2.resize INTER_AREA with an enlarged axis = OpenCV's fixed-point bilinear emulation; grey = floor of the float64
 * weighted sum).  Writes the grey frame u8[out_h][out_w] of env i at d_frames + i * env_stride.  FrameStackObservation
 * support: for envs with d_fill_mask[i] != 0 (just reset) the frame is also written to the `fill_count` preceding
 * frames (d_frames + i * env_stride - k * out_h * out_w, k = 1..fill_count), so a caller that advances d_frames by one
 * frame per step keeps a sliding stack window per env.  d_fill_mask may be NULL.  Not both axes may shrink. */
int tg_cnn_observe(tg_env *env, tg_state st, int64_t n, int32_t out_h, int32_t out_w, uint8_t *d_frames,
                   int64_t env_stride, const uint8_t *d_fill_mask, int32_t fill_count, void *stream);

/* replaces GroupedActionsObservations.observation (wrappers/grouped.py:124-207):
 *   d_feats  : NULL or u8[n][4W][W+3]      (observation_wrappers=[FeatureVectorObservation])
 *   d_boards : NULL or u8[n][4W][H_pad][W_pad] (no observation wrappers)
 *   d_legal  : u8[n][4W]                   (legal_actions_mask)                                 */
int tg_grouped_observe(tg_env *env, tg_state st, int64_t n, uint8_t *d_feats, uint8_t *d_boards,
                       uint8_t *d_legal, void *stream);

/* replaces GroupedActionsObservations.step (wrappers/grouped.py:209-269): executes placement
 * d_actions[i] in [0, 4W) (illegal ones per cfg.terminate_on_illegal), then re-enumerates.
 *   d_legal      : in = mask from the previous observe/step, out = new mask
 *   d_info_board : NULL or u8[n][W+3] = info["board"] (features of the real observation)
 *   obs          : base observation dict after the step (any pointer may be NULL)              */
int tg_grouped_step(tg_env *env, tg_state st, int64_t n, const int32_t *d_actions, uint8_t *d_legal,
                    uint8_t *d_feats, uint8_t *d_boards, uint8_t *d_info_board, tg_obs obs,
                    tg_step_out out, tg_stats *d_stats, void *stream);

/* Fused K-step rollout of an integer linear placement policy (BASELINE config 4): per step
 * enumerate the 4W placements, score = w[0]*sum(heights) + w[1]*lines + w[2]*holes + w[3]*bumpiness
 * (int32), take the lowest-index maximum over legal placements, execute it; boards stay on chip
 * for K steps.  d_stats accumulates episode statistics. */
int tg_rollout(tg_env *env, tg_state st, int64_t n, const int32_t weights[4], int32_t k_steps,
               tg_stats *d_stats, void *stream);

/* ---- functional env facade (envs/tetris_fn.py:276-367: reset / step on an explicit State) ----------
 * The functional env is a different game from the NumPy env (7 actions, no holder, queue == bag, score-delta
 * rewards).  State arrays are caller-owned and passed in AND out (aliasing allowed):
 *   board   i8[n][H_pad][W_pad]
 *   scalars i32[n][TG_FN_SCALARS + queue_size] = active_tetromino, rotation, x, y, queue_index, game_over,
 *           score (float32 bits), rng_key[0], rng_key[1], then the queue
 *   obs     i8[n][H][W] in {-1, 0, 1}   (get_observation, envs/tetris_fn.py:137-158)
 * d_actions == NULL performs reset (scalars_in then only provides rng_key).  d_piece_seq (nullable, u8[n][seq_len])
 * injects the bags: bag k of env e = seq[e][k*Q .. k*Q+Q); otherwise bags come from Philox(rng_key): permutations
 * (queue.create_bag_queue, functional/queue.py:20-35) when seq_len >= 0, the uniform queue (queue.create_uniform_queue,
 * functional/queue.py:71-87: Q draws from [0, Q - 1), maxval exclusive as in the reference) when seq_len < 0. */
#define TG_FN_SCALARS 9
int tg_fn_step(int32_t width, int32_t height, int32_t queue_size, int32_t gravity, int64_t n,
               const int8_t *d_board_in, const int32_t *d_scalars_in, const int32_t *d_actions,
               const uint8_t *d_piece_seq, int64_t seq_len,
               int8_t *d_board_out, int32_t *d_scalars_out, int8_t *d_obs, float *d_reward,
               uint8_t *d_terminated, int32_t *d_lines, void *stream);

/* test hook: i32[n] device buffer receiving the placement chosen at the last step of tg_rollout (NULL = off) */
int tg_debug_set_rollout_trace(tg_env *env, int32_t *d_last_action);

/* ---- state access (replaces Tetris.get_state/set_state, envs/tetris.py:681-708, and the direct
 * env.unwrapped.board/x/y/active_tetromino pokes of the reference tests) ---------------------- */
/* canonical, unpacked views: board u8[n][H_pad][W_pad] (locked cells, bedrock = 1);
 * scalars i32[n][TG_SCALARS + queue_size (+ 2 * holder_size if holder_size > 1)] = x, y, piece (0..6), rotation (0..3 rot90(k=+1)
 * presses), holder piece (-1 = empty), holder rotation, has_swapped, game_over, then the queue; with holder_size > 1 columns 4 / 5
 * hold the number of held pieces / 0, and the held (piece, rotation) pairs follow the queue, oldest first (-1 = empty slot). */
#define TG_SCALARS 8
int tg_get_state(tg_env *env, tg_state st, int64_t n, uint8_t *d_board, int32_t *d_scalars, void *stream);
int tg_set_state(tg_env *env, tg_state st, int64_t n, const uint8_t *d_board, const int32_t *d_scalars,
                 const uint8_t *d_mask, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TETRIS_B200_H */
