/*
 * tetris_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference NumPy Tetris env, its components and its
 * wrappers, written to follow the reference's *algorithm* (byte boards, n x n piece
 * matrices rotated with rot90, top-down drop loops, row filtering) -- deliberately NOT
 * the bitboard formulation the CUDA product path uses, so that the two are independent.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product package never links or calls it.
 *
 * Parity pinning: this file is checked (tests/test_oracle_golden.py) against
 *   - the reference's own golden vectors (tests/test_grouped_env/expected_result_i_placement.csv,
 *     the legal-mask table, the mock-board features, the 4-line-clear reward 161, shapes), and
 *   - trajectories recorded from the UNMODIFIED reference imported in the build container
 *     (oracle/make_golden.py -> the .npz files under tests/golden/),
 * and, in the build container only, step-for-step against the live reference
 * (oracle/validate_against_reference.py).
 *
 * Reference citations are `file:line` relative to /root/reference/tetris_gymnasium/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXQ 16
#define ORC_P 4 /* padding = max tetromino matrix dim (envs/tetris.py:130) */

typedef struct {
    int idx; /* index into TETROMINOES, -1 = None */
    int id;  /* cell value = idx + 2 (offset_tetromino_id, envs/tetris.py:658-679) */
    int n;   /* matrix is n x n */
    uint8_t m[16];
} orc_piece;

typedef struct {
    int32_t width, height, gravity, queue_size;
    /* ActionsMapping (mappings/actions.py:12-19): left,right,down,cw,ccw,hard_drop,swap,no_op */
    int32_t act[8];
    /* RewardsMapping (mappings/rewards.py:12-15) */
    double r_alife, r_clear_line, r_game_over, r_invalid;
} orc_config;

typedef struct {
    orc_config c;
    int W, H, Wp, Hp, Q;
    uint8_t *board; /* Hp x Wp, row major */
    orc_piece active;
    int x, y;
    int queue[ORC_MAXQ]; /* deque of piece indices, front = [0] (components/tetromino_queue.py) */
    int holder_len;      /* TetrominoHolder (components/tetromino_holder.py:14-21): deque(maxlen = size), oldest first */
    int holder_size;     /* 1 (reference default) .. 4 */
    orc_piece held[4];
    int has_swapped, game_over;
    /* randomizer: scripted stream or numpy-exact 7-bag */
    int rng_mode; /* 0 scripted, 1 numpy PCG64 bag, 2 numpy PCG64 TrueRandomizer */
    const uint8_t *seq;
    int64_t seq_len, seq_cur;
    int8_t bag[7];
    int bag_index;
    /* the tetromino set (Tetris(tetrominoes=...), envs/tetris.py:117-127): defaults to TETROMINOES */
    int np;
    int base_n[7];
    uint8_t base_m[7][16];
    uint8_t colors[16][3];
    uint64_t pcg_state_hi, pcg_state_lo, pcg_inc_hi, pcg_inc_lo;
    int pcg_has32;
    uint32_t pcg_u32;
} orc_env;

/* ---- Tetris.TETROMINOES (envs/tetris.py:47-75): I O T S Z J L ------------------------- */
static const int BASE_N[7] = {4, 2, 3, 3, 3, 3, 3};
static const uint8_t BASE_M[7][16] = {
    {0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0}, /* I */
    {1, 1, 1, 1},                                     /* O */
    {0, 1, 0, 1, 1, 1, 0, 0, 0},                      /* T */
    {0, 1, 1, 1, 1, 0, 0, 0, 0},                      /* S */
    {1, 1, 0, 0, 1, 1, 0, 0, 0},                      /* Z */
    {1, 0, 0, 1, 1, 1, 0, 0, 0},                      /* J */
    {0, 0, 1, 1, 1, 1, 0, 0, 0},                      /* L */
};
/* colours: BASE_PIXELS (envs/tetris.py:45) then the tetromino colours (envs/tetris.py:47-75) */
static const uint8_t COLORS[9][3] = {{0, 0, 0},     {128, 128, 128}, {0, 240, 240},
                                     {240, 240, 0}, {160, 0, 240},   {0, 240, 0},
                                     {240, 0, 0},   {0, 0, 240},     {240, 160, 0}};

static void make_piece(const orc_env *e, orc_piece *p, int idx) {
    /* tetrominoes[i].matrix * (i + offset), offset = len(base_pixels) = 2 (envs/tetris.py:675-677) */
    p->idx = idx;
    p->id = idx + 2;
    p->n = e->base_n[idx];
    memset(p->m, 0, 16);
    for (int k = 0; k < p->n * p->n; k++) p->m[k] = (uint8_t)(e->base_m[idx][k] * (idx + 2));
}

/* Tetris.rotate (envs/tetris.py:429-443): np.rot90(matrix, k = 1 if clockwise else -1).
 * np.rot90(m, 1)[i][j] = m[j][n-1-i];  np.rot90(m, -1)[i][j] = m[n-1-j][i]. */
static void rotate_piece(orc_piece *dst, const orc_piece *src, int clockwise) {
    orc_piece t = *src;
    int n = src->n;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            t.m[i * n + j] = clockwise ? src->m[j * n + (n - 1 - i)] : src->m[(n - 1 - j) * n + i];
    *dst = t;
}

/* ---- randomizer ----------------------------------------------------------------------- */
/* numpy PCG64 (pcg64.h: pcg_setseq_128_xsl_rr_64_random_r): state = state*MULT + inc, then
 * out = rotr64(hi ^ lo, hi >> 58).  numpy is a third-party dependency of the reference
 * (poetry.lock pins numpy 2.2.4); restated from its published algorithm and pinned against
 * numpy itself in tests/test_oracle_golden.py::test_numpy_bag_stream. */
static uint64_t pcg64_next64(orc_env *e) {
    const unsigned __int128 MULT =
        ((unsigned __int128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
    unsigned __int128 st = ((unsigned __int128)e->pcg_state_hi << 64) | e->pcg_state_lo;
    unsigned __int128 inc = ((unsigned __int128)e->pcg_inc_hi << 64) | e->pcg_inc_lo;
    st = st * MULT + inc;
    e->pcg_state_hi = (uint64_t)(st >> 64);
    e->pcg_state_lo = (uint64_t)st;
    uint64_t x = e->pcg_state_hi ^ e->pcg_state_lo;
    unsigned rot = (unsigned)(e->pcg_state_hi >> 58);
    return (x >> rot) | (x << ((64 - rot) & 63));
}
static uint32_t pcg64_next32(orc_env *e) {
    if (e->pcg_has32) {
        e->pcg_has32 = 0;
        return e->pcg_u32;
    }
    uint64_t v = pcg64_next64(e);
    e->pcg_has32 = 1;
    e->pcg_u32 = (uint32_t)(v >> 32);
    return (uint32_t)v;
}
/* numpy random_interval (distributions.c): masked rejection on next_uint32 */
static uint32_t np_random_interval(orc_env *e, uint32_t max) {
    if (max == 0) return 0;
    uint32_t mask = max, v;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    while ((v = (pcg64_next32(e) & mask)) > max) {
    }
    return v;
}
/* BagRandomizer.shuffle_bag (components/tetromino_randomizer.py:82-85): rng.shuffle(bag) in
 * place (Generator.shuffle -> _shuffle_raw: for i = n-1 .. 1: j = random_interval(i); swap). */
static void shuffle_bag(orc_env *e) {
    for (int i = e->np - 1; i >= 1; i--) {
        int j = (int)np_random_interval(e, (uint32_t)i);
        int8_t t = e->bag[i];
        e->bag[i] = e->bag[j];
        e->bag[j] = t;
    }
    e->bag_index = 0;
}
/* numpy Generator.integers(0, n) for a scalar draw (numpy/random/_bounded_integers.pyx: _rand_int64 ->
 * random_bounded_uint64_fill -> buffered_bounded_lemire_uint32, distributions.c): Lemire's
 * multiply-shift with rejection on next_uint32.  Third-party arithmetic (numpy, poetry.lock pin 2.2.4);
 * pinned against numpy itself in tests/test_oracle_golden.py::test_numpy_true_randomizer_stream. */
static uint32_t np_bounded_lemire32(orc_env *e, uint32_t rng) {
    const uint32_t rng_excl = rng + 1;
    uint64_t m = (uint64_t)pcg64_next32(e) * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
        const uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
        while (leftover < threshold) {
            m = (uint64_t)pcg64_next32(e) * rng_excl;
            leftover = (uint32_t)m;
        }
    }
    return (uint32_t)(m >> 32);
}
/* Randomizer.get_next_tetromino */
static int rnd_next(orc_env *e) {
    /* TrueRandomizer.get_next_tetromino (components/tetromino_randomizer.py:119-121): rng.integers(0, size) */
    if (e->rng_mode == 2) return e->np > 1 ? (int)np_bounded_lemire32(e, (uint32_t)(e->np - 1)) : 0; /* empty range: numpy returns `low` without drawing */
    if (e->rng_mode == 0) { /* scripted stream (test injection hook) */
        int v = e->seq[e->seq_cur % e->seq_len];
        e->seq_cur++;
        return v;
    }
    /* BagRandomizer.get_next_tetromino (components/tetromino_randomizer.py:67-80) */
    int v = e->bag[e->bag_index];
    e->bag_index++;
    if (e->bag_index >= e->np) shuffle_bag(e);
    return v;
}
/* BagRandomizer.reset (components/tetromino_randomizer.py:87-91); the reseed itself
 * (Randomizer.reset :34-46, "if seed and seed > 0") is done by the caller via orc_seed_numpy. */
static void rnd_reset(orc_env *e) {
    if (e->rng_mode != 1) return; /* scripted; TrueRandomizer.reset only reseeds (:123-127) */
    for (int i = 0; i < e->np; i++) e->bag[i] = (int8_t)i;
    shuffle_bag(e);
}

/* TetrominoQueue.reset / get_next_tetromino (components/tetromino_queue.py:24-42) */
static void queue_reset(orc_env *e) {
    rnd_reset(e);
    for (int i = 0; i < e->Q; i++) e->queue[i] = rnd_next(e);
}
static int queue_pop(orc_env *e) {
    int t = e->queue[0];
    for (int i = 0; i + 1 < e->Q; i++) e->queue[i] = e->queue[i + 1];
    e->queue[e->Q - 1] = rnd_next(e);
    return t;
}

/* ---- board helpers ---------------------------------------------------------------------- */
/* Tetris.create_board (envs/tetris.py:632-641) */
static void create_board(const orc_env *e, uint8_t *b) {
    for (int r = 0; r < e->Hp; r++)
        for (int c = 0; c < e->Wp; c++)
            b[r * e->Wp + c] = (r < e->H && c >= ORC_P && c < ORC_P + e->W) ? 0 : 1;
}
/* Tetris.collision (envs/tetris.py:408-427): any(board_subsection[matrix > 0] > 0) */
static int collision(const orc_env *e, const uint8_t *b, const orc_piece *t, int x, int y) {
    for (int i = 0; i < t->n; i++)
        for (int j = 0; j < t->n; j++)
            if (t->m[i * t->n + j] > 0 && b[(y + i) * e->Wp + (x + j)] > 0) return 1;
    return 0;
}
/* GroupedActionsObservations.collision_with_frame (wrappers/grouped.py:101-122): == 1 */
static int collision_with_frame(const orc_env *e, const uint8_t *b, const orc_piece *t, int x,
                                int y) {
    for (int i = 0; i < t->n; i++)
        for (int j = 0; j < t->n; j++)
            if (t->m[i * t->n + j] > 0 && b[(y + i) * e->Wp + (x + j)] == 1) return 1;
    return 0;
}
/* Tetris.project_tetromino (envs/tetris.py:543-564): copy; unchanged if colliding; else += */
static void project(const orc_env *e, const uint8_t *b, const orc_piece *t, int x, int y,
                    uint8_t *out) {
    memcpy(out, b, (size_t)e->Hp * e->Wp);
    if (collision(e, b, t, x, y)) return;
    for (int i = 0; i < t->n; i++)
        for (int j = 0; j < t->n; j++) out[(y + i) * e->Wp + (x + j)] += t->m[i * t->n + j];
}
/* Tetris.clear_filled_rows (envs/tetris.py:481-512) */
static int clear_filled_rows(const orc_env *e, uint8_t *b) {
    int Wp = e->Wp, Hp = e->Hp;
    uint8_t filled[128];
    int n_filled = 0;
    for (int r = 0; r < Hp; r++) {
        int any0 = 0, all1 = 1;
        for (int c = 0; c < Wp; c++) {
            if (b[r * Wp + c] == 0) any0 = 1;
            if (b[r * Wp + c] != 1) all1 = 0;
        }
        filled[r] = (uint8_t)((!any0) && (!all1));
        n_filled += filled[r];
    }
    if (n_filled > 0) {
        uint8_t *tmp = (uint8_t *)malloc((size_t)Hp * Wp);
        int o = 0;
        for (int r = 0; r < n_filled; r++, o++) /* free_space rows padded with 1 */
            for (int c = 0; c < Wp; c++) tmp[o * Wp + c] = (c >= ORC_P && c < ORC_P + e->W) ? 0 : 1;
        for (int r = 0; r < Hp; r++)
            if (!filled[r]) {
                memcpy(tmp + (size_t)o * Wp, b + (size_t)r * Wp, (size_t)Wp);
                o++;
            }
        memcpy(b, tmp, (size_t)Hp * Wp);
        free(tmp);
    }
    return n_filled;
}
/* Tetris.reset_tetromino_position (envs/tetris.py:536-541) */
static void reset_position(orc_env *e) {
    e->x = e->Wp / 2 - e->active.n / 2;
    e->y = 0;
}
/* Tetris.spawn_tetromino (envs/tetris.py:393-401) */
static int spawn(orc_env *e) {
    make_piece(e, &e->active, queue_pop(e));
    reset_position(e);
    return !collision(e, e->board, &e->active, e->x, e->y);
}
/* Tetris.commit_active_tetromino (envs/tetris.py:450-479) */
static void commit(orc_env *e, double *reward, int *lines) {
    *lines = 0;
    if (collision(e, e->board, &e->active, e->x, e->y)) {
        *reward = e->c.r_game_over;
        e->game_over = 1;
    } else {
        while (!collision(e, e->board, &e->active, e->x, e->y + 1)) e->y++; /* :445-448 */
        uint8_t *tmp = (uint8_t *)malloc((size_t)e->Hp * e->Wp);
        project(e, e->board, &e->active, e->x, e->y, tmp); /* place_active_tetromino :403-406 */
        memcpy(e->board, tmp, (size_t)e->Hp * e->Wp);
        free(tmp);
        *lines = clear_filled_rows(e, e->board);
        *reward = (double)((*lines) * (*lines) * e->W); /* score :621-630 */
        e->game_over = !spawn(e);
        *reward += e->c.r_alife;
        if (e->game_over) *reward = e->c.r_game_over;
        e->has_swapped = 0;
    }
}

/* ---- public API ------------------------------------------------------------------------- */
orc_env *orc_create(const orc_config *c) {
    if (c->queue_size < 1 || c->queue_size > ORC_MAXQ || c->height + ORC_P > 128) return NULL;
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    e->c = *c;
    e->W = c->width;
    e->H = c->height;
    e->Wp = c->width + 2 * ORC_P; /* envs/tetris.py:131-132 */
    e->Hp = c->height + ORC_P;
    e->Q = c->queue_size;
    e->board = (uint8_t *)malloc((size_t)e->Hp * e->Wp);
    create_board(e, e->board);
    e->active.idx = -1;
    e->rng_mode = 0;
    e->holder_size = 1;
    e->np = 7;
    for (int p = 0; p < 7; p++) {
        e->base_n[p] = BASE_N[p];
        memcpy(e->base_m[p], BASE_M[p], 16);
    }
    memcpy(e->colors, COLORS, sizeof COLORS);
    return e;
}
void orc_destroy(orc_env *e) {
    if (!e) return;
    free(e->board);
    free(e);
}
/* Tetris(tetrominoes=[...]) (envs/tetris.py:117-132): np pieces, piece i an n[i] x n[i] binary matrix (row-major in m[i]) with
 * colour rgb[i]; board values are i + 2 (offset_tetromino_id multiplies the matrix by index + offset).  Call before reset. */
void orc_set_tetrominoes(orc_env *e, int np, const int32_t *n, const uint8_t *m, const uint8_t *rgb) {
    e->np = np;
    for (int p = 0; p < np; p++) {
        e->base_n[p] = n[p];
        for (int k = 0; k < 16; k++) e->base_m[p][k] = m[p * 16 + k] != 0;
        memcpy(e->colors[p + 2], rgb + 3 * p, 3);
    }
}
/* TetrominoHolder(size): the reference can only get a bigger holder by assigning env.holder after construction */
void orc_set_holder_size(orc_env *e, int size) {
    e->holder_size = size < 1 ? 1 : (size > 4 ? 4 : size);
    e->holder_len = 0;
}
int orc_get_holder_len(const orc_env *e) { return e->holder_len; }
/* slot-th held piece (0 = oldest): returns its index or -1, matrix as n x n in m16 */
int orc_get_held_slot(const orc_env *e, int slot, int32_t *n, uint8_t *m16) {
    if (slot < 0 || slot >= e->holder_len) { *n = 0; return -1; }
    *n = e->held[slot].n;
    memcpy(m16, e->held[slot].m, 16);
    return e->held[slot].idx;
}
/* scripted randomizer: stream of piece indices consumed in order (wraps at len) */
void orc_set_sequence(orc_env *e, const uint8_t *seq, int64_t len, int64_t cursor) {
    e->rng_mode = 0;
    e->seq = seq;
    e->seq_len = len;
    e->seq_cur = cursor;
}
/* numpy-exact 7-bag: st = {state_hi, state_lo, inc_hi, inc_lo} of PCG64(SeedSequence(seed)) */
void orc_set_true_randomizer(orc_env *e, int on) { e->rng_mode = on ? 2 : 1; }
void orc_seed_numpy(orc_env *e, const uint64_t *st) {
    if (e->rng_mode == 0) e->rng_mode = 1;
    e->pcg_state_hi = st[0];
    e->pcg_state_lo = st[1];
    e->pcg_inc_hi = st[2];
    e->pcg_inc_lo = st[3];
    e->pcg_has32 = 0;
    e->pcg_u32 = 0;
}

/* Tetris._get_obs (envs/tetris.py:566-615); any output pointer may be NULL */
void orc_get_obs(const orc_env *e, uint8_t *board, uint8_t *mask, uint8_t *holder,
                 uint8_t *queue) {
    int Wp = e->Wp, Hp = e->Hp;
    if (board) project(e, e->board, &e->active, e->x, e->y, board);
    if (mask) {
        memset(mask, 0, (size_t)Hp * Wp);
        for (int i = 0; i < e->active.n; i++)
            for (int j = 0; j < e->active.n; j++) mask[(e->y + i) * Wp + (e->x + j)] = 1;
    }
    if (holder) {
        /* :578-594.  size 1: the held piece padded to P x P, or np.ones((P, P * size)) when the holder is empty.  size > 1: the
         * reference hstacks the held pieces only -- an array whose WIDTH depends on how many pieces are held; here the array
         * has the fixed shape (P, P * size): the held pieces oldest first, ones in the empty slots (identical when the holder
         * is empty or full; RgbObservation pads the ragged array with ones up to the queue's width, :49-58, which gives
         * exactly this layout). */
        int HW = ORC_P * e->holder_size;
        memset(holder, 1, (size_t)ORC_P * HW);
        for (int s = 0; s < e->holder_len; s++) {
            const orc_piece *t = &e->held[s];
            for (int i = 0; i < ORC_P; i++)
                for (int j = 0; j < ORC_P; j++) holder[i * HW + s * ORC_P + j] = (i < t->n && j < t->n) ? t->m[i * t->n + j] : 0;
        }
    }
    if (queue) {
        int QW = ORC_P * e->Q;
        memset(queue, 0, (size_t)ORC_P * QW);
        for (int q = 0; q < e->Q; q++) {
            orc_piece t;
            make_piece(e, &t, e->queue[q]);
            for (int i = 0; i < t.n; i++)
                for (int j = 0; j < t.n; j++) queue[i * QW + q * ORC_P + j] = t.m[i * t.n + j];
        }
    }
}

/* Tetris.reset (envs/tetris.py:274-307) */
void orc_reset(orc_env *e) {
    create_board(e, e->board);
    e->game_over = 0;
    queue_reset(e);
    make_piece(e, &e->active, queue_pop(e));
    reset_position(e);
    e->holder_len = 0;
    e->has_swapped = 0;
}

/* Tetris.step (envs/tetris.py:203-272); returns -1 if the action is outside Discrete(8) */
int orc_step(orc_env *e, int action, double *reward, int *terminated, int *lines) {
    if (action < 0 || action >= 8) return -1; /* assert action_space.contains :215 */
    const int32_t *A = e->c.act;
    double rew = 0.0;
    int ln = 0;
    orc_piece rot;
    if (action == A[0]) {
        if (!collision(e, e->board, &e->active, e->x - 1, e->y)) e->x -= 1;
    } else if (action == A[1]) {
        if (!collision(e, e->board, &e->active, e->x + 1, e->y)) e->x += 1;
    } else if (action == A[2]) {
        if (!collision(e, e->board, &e->active, e->x, e->y + 1)) e->y += 1;
    } else if (action == A[3]) {
        rotate_piece(&rot, &e->active, 1);
        if (!collision(e, e->board, &rot, e->x, e->y)) e->active = rot;
    } else if (action == A[4]) {
        rotate_piece(&rot, &e->active, 0);
        if (!collision(e, e->board, &rot, e->x, e->y)) e->active = rot;
    } else if (action == A[6]) { /* swap :242-252 */
        if (!e->has_swapped) {
            /* TetrominoHolder.swap (components/tetromino_holder.py:31-49) */
            orc_piece cur = e->active;
            if (e->holder_len < e->holder_size) { /* not full: store, nothing comes back */
                e->held[e->holder_len++] = cur;
                e->has_swapped = 1;
                (void)spawn(e); /* result ignored :250 */
            } else { /* full: the oldest piece comes back (popleft), the active one is appended */
                e->active = e->held[0];
                for (int s = 1; s < e->holder_size; s++) e->held[s - 1] = e->held[s];
                e->held[e->holder_size - 1] = cur;
                e->has_swapped = 1;
                reset_position(e);
            }
        }
    } else if (action == A[5]) { /* hard drop :253-254 */
        commit(e, &rew, &ln);
    } else if (action == A[7]) {
    }
    if (e->c.gravity && action != A[5]) { /* :259-264 */
        if (!collision(e, e->board, &e->active, e->x, e->y + 1))
            e->y += 1;
        else
            commit(e, &rew, &ln);
    }
    *reward = rew;
    *terminated = e->game_over;
    *lines = ln;
    return 0;
}

/* ---- FeatureVectorObservation (wrappers/observation.py:177-278) --------------------------- */
/* `board` is an observation board (Hp x Wp) and is MUTATED like the reference mutates
 * obs["board"]: `board_obs[active_tetromino_mask] = 0` indexes ROWS with the uint8 mask values
 * (integer fancy indexing): row 0 is zeroed always, row 1 too iff the mask contains a 1. */
void orc_features(const orc_env *e, uint8_t *board, const uint8_t *mask, uint8_t *out) {
    int Wp = e->Wp, Hp = e->Hp, W = e->W, H = e->H;
    int has1 = 0;
    for (int k = 0; k < Hp * Wp; k++) has1 |= (mask[k] == 1);
    memset(board, 0, (size_t)Wp);
    if (has1) memset(board + Wp, 0, (size_t)Wp);
    int heights[64];
    int holes = 0, bump = 0, maxh = 0;
    for (int c = 0; c < W; c++) {
        /* calc_height :177-193: H - argmax(col != 0); 0 for an all-empty column */
        int first = -1;
        for (int r = 0; r < H; r++)
            if (board[r * Wp + ORC_P + c] != 0) {
                first = r;
                break;
            }
        heights[c] = first < 0 ? 0 : H - first;
        if (heights[c] > maxh) maxh = heights[c];
        /* calc_holes :222-236: (cell == 0) & (cumsum(filled) > 0) */
        int seen = 0;
        for (int r = 0; r < H; r++) {
            if (board[r * Wp + ORC_P + c] != 0)
                seen = 1;
            else if (seen)
                holes++;
        }
    }
    for (int c = 0; c + 1 < W; c++) bump += abs(heights[c + 1] - heights[c]); /* :207-220 */
    for (int c = 0; c < W; c++) out[c] = (uint8_t)heights[c];
    out[W] = (uint8_t)maxh;
    out[W + 1] = (uint8_t)holes; /* np.array(features, dtype=np.uint8) wraps mod 256 :277 */
    out[W + 2] = (uint8_t)bump;
}

/* ---- GroupedActionsObservations.observation (wrappers/grouped.py:124-207) ----------------- */
/* boards: NULL or u8[4W][Hp][Wp]; feats: NULL or u8[4W][W+3]; legal: u8[4W];
 * lines: NULL or i32[4W] = rows cleared by a regular placement, -1 for a game-over placement, -2 illegal
 * (not part of the reference observation; used by the heuristic-rollout parity test) */
void orc_grouped_observe_ex(const orc_env *e, uint8_t *boards, uint8_t *feats, uint8_t *legal, int32_t *lines) {
    int Wp = e->Wp, Hp = e->Hp, W = e->W;
    size_t bsz = (size_t)Hp * Wp;
    uint8_t *tmp = (uint8_t *)malloc(bsz), *zeros = (uint8_t *)calloc(bsz, 1);
    orc_piece t = e->active;
    for (int xb = 0; xb < W; xb++) {
        for (int r = 0; r < 4; r++) {
            int y = 0;
            if (r > 0) rotate_piece(&t, &t, 1);
            int x = xb + ORC_P - t.n / 2;
            while (!collision(e, e->board, &t, x, y + 1)) y++; /* no test at y = 0 */
            int a = xb * 4 + r;
            legal[a] = 1;
            if (collision_with_frame(e, e->board, &t, x, y)) {
                legal[a] = 0;
                memset(tmp, 1, bsz);
                if (lines) lines[a] = -2;
            } else if (collision(e, e->board, &t, x, y)) {
                memset(tmp, 0, bsz);
                if (lines) lines[a] = -1;
            } else {
                project(e, e->board, &t, x, y, tmp);
                int nl = clear_filled_rows(e, tmp);
                if (lines) lines[a] = nl;
            }
            if (boards) memcpy(boards + (size_t)a * bsz, tmp, bsz);
            if (feats) orc_features(e, tmp, zeros, feats + (size_t)a * (W + 3));
        }
        rotate_piece(&t, &t, 1);
    }
    free(tmp);
    free(zeros);
}

void orc_grouped_observe(const orc_env *e, uint8_t *boards, uint8_t *feats, uint8_t *legal) {
    orc_grouped_observe_ex(e, boards, feats, legal, NULL);
}

/* GroupedActionsObservations.step (wrappers/grouped.py:209-269).
 * `legal` is the mask produced by the previous observe().  Returns
 *   0 legal placement executed, 1 illegal + terminate (env untouched), 2 illegal + no-op step. */
int orc_grouped_step(orc_env *e, int action, const uint8_t *legal, int terminate_on_illegal,
                     double *reward, int *terminated, int *lines) {
    int xb = action / 4, r = action % 4;
    if (legal[action] == 0) {
        int code = 1;
        if (terminate_on_illegal) {
            *terminated = 1;
            *lines = 0;
        } else {
            orc_step(e, e->c.act[7], reward, terminated, lines);
            code = 2;
        }
        *reward = e->c.r_invalid;
        return code;
    }
    orc_piece nt = e->active;
    int x = xb + ORC_P - e->active.n / 2;
    for (int k = 0; k < r; k++) rotate_piece(&nt, &nt, 1);
    e->x = x;
    e->active = nt;
    orc_step(e, e->c.act[5], reward, terminated, lines);
    return 0;
}

/* ---- RgbObservation.observation (wrappers/observation.py:38-74) --------------------------- */
/* out: u8[Hp][Wp + max(Q, holder)*P][3] */
void orc_rgb(const orc_env *e, uint8_t *out) {
    int Wp = e->Wp, Hp = e->Hp, P = ORC_P;
    int max_len = P * (e->Q > 1 ? e->Q : 1);
    int OW = Wp + max_len;
    uint8_t *board = (uint8_t *)malloc((size_t)Hp * Wp);
    uint8_t holder[16 * 4], queue[4 * 4 * ORC_MAXQ];
    orc_get_obs(e, board, NULL, holder, queue);
    for (int r = 0; r < Hp; r++)
        for (int c = 0; c < OW; c++) {
            int v;
            if (c < Wp)
                v = board[r * Wp + c];
            else {
                int cc = c - Wp;
                if (r < P)
                    v = queue[r * (P * e->Q) + cc]; /* queue is max_len wide (Q >= holder) */
                else if (r >= Hp - P)
                    v = cc < P * e->holder_size ? holder[(r - (Hp - P)) * (P * e->holder_size) + cc] : 1;
                else
                    v = 1;
            }
            memcpy(out + ((size_t)r * OW + c) * 3, e->colors[v & 15], 3);
        }
    free(board);
}

/* ---- state access (the reference tests poke env.unwrapped.* directly) ---------------------- */
/* ints: x, y, active idx, active rot-normalised? (not tracked: matrices), holder idx (-1 none),
 * has_swapped, game_over */
void orc_get_board(const orc_env *e, uint8_t *out) { memcpy(out, e->board, (size_t)e->Hp * e->Wp); }
void orc_set_board(orc_env *e, const uint8_t *in) { memcpy(e->board, in, (size_t)e->Hp * e->Wp); }
void orc_get_scalars(const orc_env *e, int32_t *out) {
    out[0] = e->x;
    out[1] = e->y;
    out[2] = e->active.idx;
    out[3] = e->holder_len ? e->held[0].idx : -1;
    out[4] = e->has_swapped;
    out[5] = e->game_over;
    for (int q = 0; q < e->Q; q++) out[6 + q] = e->queue[q];
}
void orc_get_active_matrix(const orc_env *e, int32_t *n, uint8_t *m16) {
    *n = e->active.n;
    memcpy(m16, e->active.m, 16);
}
void orc_get_held_matrix(const orc_env *e, int32_t *n, uint8_t *m16) {
    *n = e->holder_len ? e->held[0].n : 0;
    memcpy(m16, e->held[0].m, 16);
}
/* set the active piece to TETROMINOES[idx] rotated `rot` times with rot90(k=+1) */
void orc_set_active(orc_env *e, int idx, int rot, int x, int y) {
    make_piece(e, &e->active, idx);
    for (int k = 0; k < (rot & 3); k++) rotate_piece(&e->active, &e->active, 1);
    e->x = x;
    e->y = y;
}
void orc_set_flags(orc_env *e, int has_swapped, int game_over) {
    e->has_swapped = has_swapped;
    e->game_over = game_over;
}
void orc_set_holder(orc_env *e, int idx, int rot) {
    if (idx < 0) {
        e->holder_len = 0;
        return;
    }
    make_piece(e, &e->held[0], idx);
    for (int k = 0; k < (rot & 3); k++) rotate_piece(&e->held[0], &e->held[0], 1);
    e->holder_len = 1;
}
void orc_set_queue(orc_env *e, const int32_t *q) {
    for (int i = 0; i < e->Q; i++) e->queue[i] = q[i];
}

/* ---- batched drivers (cpu_baseline / --impl reference legs of bench.py) ------------------- */
/* One vector-env step over n envs with gymnasium's NEXT_STEP autoreset (SyncVectorEnv.step,
 * gymnasium 1.1.1): an env that terminated on the previous call is reset instead of stepped.
 * obs pointers may be NULL (then that part of the observation is not produced). */
void orc_vec_step(orc_env **envs, int64_t n, const int32_t *actions, uint8_t *autoreset,
                  uint8_t *board, uint8_t *mask, uint8_t *holder, uint8_t *queue, float *reward,
                  uint8_t *terminated, int32_t *lines, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < n; i++) {
        orc_env *e = envs[i];
        size_t bsz = (size_t)e->Hp * e->Wp;
        double r = 0.0;
        int t = 0, l = 0;
        if (autoreset && autoreset[i]) {
            orc_reset(e);
        } else {
            orc_step(e, actions[i], &r, &t, &l);
        }
        orc_get_obs(e, board ? board + i * bsz : NULL, mask ? mask + i * bsz : NULL,
                    holder ? holder + i * (16 * (size_t)e->holder_size) : NULL, queue ? queue + i * (16 * (size_t)e->Q) : NULL);
        reward[i] = (float)r;
        terminated[i] = (uint8_t)t;
        lines[i] = l;
        if (autoreset) autoreset[i] = (uint8_t)t;
    }
}

/* grouped vector step: execute placement actions[i] then re-enumerate (features + mask) */
void orc_vec_grouped_step(orc_env **envs, int64_t n, const int32_t *actions, uint8_t *autoreset,
                          uint8_t *feats, uint8_t *legal, float *reward, uint8_t *terminated,
                          int32_t *lines, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < n; i++) {
        orc_env *e = envs[i];
        int A = 4 * e->W, F = e->W + 3;
        double r = 0.0;
        int t = 0, l = 0;
        if (autoreset && autoreset[i]) {
            orc_reset(e);
        } else {
            orc_grouped_step(e, actions[i], legal + i * A, 1, &r, &t, &l);
        }
        orc_grouped_observe(e, NULL, feats + i * (size_t)A * F, legal + i * A);
        reward[i] = (float)r;
        terminated[i] = (uint8_t)t;
        lines[i] = l;
        if (autoreset) autoreset[i] = (uint8_t)t;
    }
}

/* ---- bulk helpers of the bench's CPU arm: a million Python-side OracleEnv objects would take minutes to build -------- */
void orc_vec_create(const orc_config *c, int64_t n, orc_env **out) {
    for (int64_t i = 0; i < n; i++) out[i] = orc_create(c);
}
void orc_vec_destroy(orc_env **envs, int64_t n) {
    for (int64_t i = 0; i < n; i++) orc_destroy(envs[i]);
}
/* words = SeedSequence(seed).generate_state(4, uint64) per env (oracle/np_seed.py, vectorised); this is
 * pcg_setseq_128_srandom_r(state = w0:w1, seq = w2:w3) of numpy/random/src/pcg64/pcg64.h */
void orc_vec_seed_words(orc_env **envs, int64_t n, const uint64_t *words) {
    const unsigned __int128 MULT = ((unsigned __int128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
    for (int64_t i = 0; i < n; i++) {
        const uint64_t *w = words + 4 * i;
        unsigned __int128 initstate = ((unsigned __int128)w[0] << 64) | w[1], initseq = ((unsigned __int128)w[2] << 64) | w[3];
        unsigned __int128 inc = (initseq << 1) | 1, state = inc;
        state += initstate;
        state = state * MULT + inc;
        uint64_t st[4] = {(uint64_t)(state >> 64), (uint64_t)state, (uint64_t)(inc >> 64), (uint64_t)inc};
        orc_seed_numpy(envs[i], st);
    }
}
void orc_vec_reset(orc_env **envs, int64_t n, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < n; i++) orc_reset(envs[i]);
}

/* randomizer.reset() followed by n draws (pins the numpy-exact 7-bag against numpy itself) */
void orc_rnd_stream(orc_env *e, int n, uint8_t *out) {
    rnd_reset(e);
    for (int i = 0; i < n; i++) out[i] = (uint8_t)rnd_next(e);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
