// tg_api.cu -- host side of libtetris_b200.so: the C ABI declared in include/tetris_b200.h.
// No torch types, no CPU fallback: every entry point launches sm_100a kernels on the caller's stream.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include <string>
#include <vector>

#include "../../include/tetris_b200.h"
#include "tg_device.cuh"
#include "tg_step.cuh"
#include "tg_stepn.cuh"
#include "tg_aux.cuh"
#include "tg_gfeats.cuh"
#include "tg_rollout.cuh"
#include "tg_fn.cuh"
#include "tg_host_expand.h"

using namespace tg;

// piece tables on the host: uploaded to the device by upload_tables, read by the host-side dict expansion
struct HostTables {
    int np;                       // pieces of the set (7 for the reference's)
    int n[7];                     // matrix sizes
    unsigned char base[7][16];    // base matrices, n x n row-major, 0 / 1
    unsigned short cells[7][4];
    uint2 ptab[7][4];
    uint4 prec[7][4];
    unsigned int bot4[7][4];      // byte j = row offset of the lowest cell of matrix column j (0 where the column is empty)
    unsigned int rowbytes[7][4][4];
    unsigned char colors[16][4];
    uint64_t hash;
};

struct StepPlan {
    bool valid = false;
    StepParams p;          // tile size, warp roles, shared-memory offsets (pointers unset)
    void* kern = nullptr;
    int threads = 0;
    size_t smem = 0;
    int64_t grid_max = 0;  // resident CTAs of the whole device
    bool ws = false, pdl = false;
};

struct tg_env {
    StepPlan plans[6];     // [mode * 2 + with dict]
    tg_config cfg;
    DevCfg dev;
    tg_layout layout;
    int device;
    int num_sms;
    int col64;
    int tile;             // envs per CTA tile of the step kernel
    int threads_per_env;  // CTA threads = tile * threads_per_env (logic uses one thread per env, image fill uses all)
    int warp_specialized; // 1: k_step_ws (logic warp runs a tile ahead of the image warps)
    int fill_warps;       // image/store warps per CTA of k_step_ws
    int logic_warps;      // game-logic warps per CTA of k_step_ws (each runs every logic_warps-th tile of the CTA)
    int logic_warps_set, fill_warps_set;   // TG_NL / TG_NF given: use them for every launch
    void* rollout_last_action;
    int cnn_h, cnn_w;     // output size the tg_cnn_observe tables in stage[4] were built for
    int cnn_nax = 0, cnn_axv[4] = {-1, -1, -1, -1};   // classes of the x coefficient sums (k_cnn_obs2)
    bool cnn_v3_ok = false;                           // the word-wise kernel k_cnn_obs3 applies to the current tables
    std::string err;
    // tg_step_host staging
    cudaStream_t hs[3];
    bool hs_init;
    void* stage[16];
    int64_t fused_grid_max = 0;   // resident CTAs of k_grouped_step_feats (0 = not yet queried)
    size_t stage_bytes[16];
    // compact host step: pinned ring of packed records, events, expansion pool
    HostTables tabs;
    tgh::ExpandCfg xcfg;       // device record layout (tg_host_expand)
    tgh::ExpandCfg xcfg_pk;    // packed link records (tg_step_host, TG_HOST_COMPACT)
    int pk_bytes;              // bytes per env of a packed link record: 12 + id plane, rounded up to 16
    tgh::Pool* pool;
    int host_threads;          // 0 = tgh::default_threads()
    void* hring;               // pinned host memory
    size_t hring_bytes;
    cudaEvent_t hev[64];
    bool hev_init;
    cudaEvent_t caller_ev;
    double host_stats[4];      // last tg_step_host call: seconds total, waiting for the device, expanding; chunks
};

static std::string g_create_err;

// Every entry point runs on the env's device and leaves the caller's current device as it found it (a single-process
// multi-GPU program, or a destructor run by the garbage collector, must not see its current device change).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) err = cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
struct tg_env;
static int ensure_env_tables(tg_env* env);
#define ON_DEVICE(env) DeviceGuard _guard((env)->device); CUDA_TRY(env, _guard.err); { int _t = ensure_env_tables(env); if (_t) return _t; }

static int fail(tg_env* env, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (env) env->err = buf; else g_create_err = buf;
    return code;
}
#define CUDA_TRY(env, call)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess) return fail(env, TG_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_e)); \
    } while (0)

// ---- constant tables: rotate the reference's base matrices (envs/tetris.py:47-75) with rot90 -----
static const int kN[7] = {4, 2, 3, 3, 3, 3, 3};
static const unsigned char kBase[7][16] = {
    {0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0}, {1, 1, 1, 1},
    {0, 1, 0, 1, 1, 1, 0, 0, 0}, {0, 1, 1, 1, 1, 0, 0, 0, 0}, {1, 1, 0, 0, 1, 1, 0, 0, 0},
    {1, 0, 0, 1, 1, 1, 0, 0, 0}, {0, 0, 1, 1, 1, 1, 0, 0, 0}};
static const unsigned char kColors[9][3] = {{0, 0, 0}, {128, 128, 128}, {0, 240, 240}, {240, 240, 0}, {160, 0, 240},
                                            {0, 240, 0}, {240, 0, 0}, {0, 0, 240}, {240, 160, 0}};

// the piece set of a config: the reference's seven tetrominoes, or Tetris(tetrominoes=[...]) (cfg->n_pieces > 0)
static void piece_set(const tg_config* cfg, HostTables& T) {
    memset(&T, 0, sizeof T);
    for (int v = 0; v < 2; v++) for (int k = 0; k < 3; k++) T.colors[v][k] = kColors[v][k];      // BASE_PIXELS: empty, bedrock
    if (!cfg || cfg->n_pieces <= 0) {
        T.np = 7;
        for (int p = 0; p < 7; p++) {
            T.n[p] = kN[p];
            memcpy(T.base[p], kBase[p], 16);
            for (int k = 0; k < 3; k++) T.colors[p + 2][k] = kColors[p + 2][k];
        }
        return;
    }
    T.np = cfg->n_pieces;
    for (int p = 0; p < T.np; p++) {
        T.n[p] = cfg->piece_n[p];
        for (int k = 0; k < 16; k++) T.base[p][k] = cfg->piece_matrix[p][k] != 0;
        for (int k = 0; k < 3; k++) T.colors[p + 2][k] = cfg->piece_color[p][k];
    }
}

// fills the derived tables of T from T.np / T.n / T.base (set by piece_set)
static int build_tables(tg_env* env, HostTables& T) {
    auto& cells = T.cells; auto& ptab = T.ptab; auto& prec = T.prec; auto& rowbytes = T.rowbytes;
    memset(cells, 0, sizeof cells); memset(ptab, 0, sizeof ptab); memset(prec, 0, sizeof prec); memset(rowbytes, 0, sizeof rowbytes); memset(T.bot4, 0, sizeof T.bot4);
    for (int p = 0; p < T.np; p++) {
        int n = T.n[p];
        unsigned char m[16], t[16];
        memcpy(m, T.base[p], 16);
        for (int r = 0; r < 4; r++) {
            if (r > 0) {  // np.rot90(m, k=1): out[i][j] = m[j][n-1-i]  (Tetris.rotate, envs/tetris.py:429-443)
                for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) t[i * n + j] = m[j * n + (n - 1 - i)];
                memcpy(m, t, 16);
            }
            unsigned short c = 0;
            int k = 0;
            for (int i = 0; i < 4; i++) {
                unsigned int w = 0;
                for (int j = 0; j < 4; j++)
                    if (i < n && j < n && m[i * n + j]) {
                        c |= (unsigned short)(((i << 2) | j) << (4 * k));
                        k++;
                        w |= (unsigned int)(p + 2) << (8 * j);
                    }
                rowbytes[p][r][i] = w;
            }
            if (k != 4) return fail(env, TG_ERR_CONFIG, "piece table: %d cells", k);
            cells[p][r] = c;
            // column profile (place_fast): row mask, top row, cell count per matrix column; first / last used column
            unsigned int px = 0, py = 0;
            int jmin = 4, jmax = -1;
            for (int j = 0; j < 4; j++) {
                int mask = 0, top = -1, cnt = 0;
                for (int i = 0; i < 4; i++)
                    if (i < n && j < n && m[i * n + j]) { mask |= 1 << i; if (top < 0) top = i; cnt++; }
                if (cnt) { if (j < jmin) jmin = j; if (j > jmax) jmax = j; }
                px |= (unsigned)mask << (4 * j);
                px |= (unsigned)(top < 0 ? 0 : top) << (16 + 2 * j);
                py |= (unsigned)cnt << (3 * j);
            }
            for (int j = jmin; j <= jmax; j++)
                if (!((px >> (4 * j)) & 15u)) return fail(env, TG_ERR_CONFIG, "piece table: empty column inside a piece");
            px |= (unsigned)jmin << 24 | (unsigned)jmax << 26;
            ptab[p][r] = make_uint2(px, py);
            // packed-byte profile of k_grouped_feats_x (tg_gfeats.cuh)
            unsigned int m4 = 0, top4 = 0, mintop = 3, bot4 = 0;
            for (int j = 0; j < 4; j++) {
                unsigned mask = (px >> (4 * j)) & 15u, top = (px >> (16 + 2 * j)) & 3u;
                if (mask) {
                    m4 |= 0xFFu << (8 * j); top4 |= top << (8 * j); if (top < mintop) mintop = top;
                    unsigned bot = 3; while (!((mask >> bot) & 1u)) bot--;
                    bot4 |= bot << (8 * j);
                }
            }
            T.bot4[p][r] = bot4;
            prec[p][r] = make_uint4((unsigned)c | ((px & 0xFFFFu) << 16), m4, top4, (unsigned)jmin | (unsigned)jmax << 2 | mintop << 4);
        }
    }
    // identity of the set: the constant tables are per DEVICE, shared by every handle
    uint64_t hsh = 1469598103934665603ull;
    auto mix = [&](const void* d, size_t nbytes) { for (size_t i = 0; i < nbytes; i++) { hsh ^= ((const unsigned char*)d)[i]; hsh *= 1099511628211ull; } };
    mix(&T.np, sizeof T.np); mix(T.n, sizeof T.n); mix(T.base, sizeof T.base); mix(T.colors, sizeof T.colors);
    T.hash = hsh;
    return TG_OK;
}

// The piece tables live in __constant__ memory: one copy per device, shared by every handle.  Handles with different piece sets
// may coexist; whichever set the next launch needs is made resident first (after the device has drained: kernels in flight
// still read the old tables).  Switching is rare -- it happens only when envs with different sets alternate on one device.
static uint64_t g_resident_tables[64];
static int ensure_tables(tg_env* env, const HostTables& T, int device) {
    if (g_resident_tables[device & 63] == T.hash) return TG_OK;
    CUDA_TRY(env, cudaDeviceSynchronize());
    CUDA_TRY(env, cudaMemcpyToSymbol(c_cells, T.cells, sizeof T.cells));
    CUDA_TRY(env, cudaMemcpyToSymbol(c_ptab, T.ptab, sizeof T.ptab));
    CUDA_TRY(env, cudaMemcpyToSymbol(c_prec, T.prec, sizeof T.prec));
    CUDA_TRY(env, cudaMemcpyToSymbol(c_bot4, T.bot4, sizeof T.bot4));
    CUDA_TRY(env, cudaMemcpyToSymbol(c_rowbytes, T.rowbytes, sizeof T.rowbytes));
    CUDA_TRY(env, cudaMemcpyToSymbol(c_n, T.n, sizeof T.n));
    CUDA_TRY(env, cudaMemcpyToSymbol(c_colors, T.colors, sizeof T.colors));
    g_resident_tables[device & 63] = T.hash;
    return TG_OK;
}

// ExpandCfg of the host-side dict expansion from the device config + the piece tables
static void make_expand_cfg(const DevCfg& d, const HostTables& T, tgh::ExpandCfg& x) {
    memset(&x, 0, sizeof x);
    x.W = d.W; x.H = d.H; x.Wp = d.Wp; x.Hp = d.Hp; x.Q = d.Q; x.OB = d.OB; x.OQ = d.OQ;
    x.holder_size = d.holder_size; x.OH = d.OH; x.hdr = 12;
    x.board_stride = d.board_stride; x.ids_off = d.ids_off; x.ids_bytes = (d.H * d.W + 1) / 2;
    for (int p = 0; p < 7; p++) {
        x.n[p] = p < T.np ? T.n[p] : 3;
        for (int r = 0; r < 4; r++) {
            x.cells[p][r] = T.cells[p][r];
            for (int i = 0; i < 4; i++) x.rowbytes[p][r][i] = T.rowbytes[p][r][i];
        }
    }
    x.n[7] = 3;
    tgh::build_vector_tables(x);
}

extern "C" int tg_set_host_threads(tg_env* env, int32_t threads);

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

static int ensure_env_tables(tg_env* env) { return ensure_tables(env, env->tabs, env->device); }

extern "C" int tg_version(void) { return TG_VERSION; }

extern "C" const char* tg_last_error(const tg_env* env) { return env ? env->err.c_str() : g_create_err.c_str(); }

static int validate_cfg(const tg_config* cfg) {
    if (cfg->width < 4 || cfg->width + 2 * TG_PADDING > 32)
        return fail(nullptr, TG_ERR_CONFIG, "width %d unsupported (4 <= W <= 24: a padded row must fit 32 bits)", cfg->width);
    if (cfg->height < 4 || cfg->height + TG_PADDING > 64)
        return fail(nullptr, TG_ERR_CONFIG, "height %d unsupported (4 <= H <= 60)", cfg->height);
    if (cfg->queue_size < 1 || cfg->queue_size > TG_MAX_QUEUE)
        return fail(nullptr, TG_ERR_CONFIG, "queue_size %d unsupported (1..16)", cfg->queue_size);
    if (cfg->rng_mode < 0 || cfg->rng_mode > 2) return fail(nullptr, TG_ERR_CONFIG, "rng_mode %d", cfg->rng_mode);
    if (cfg->randomizer < 0 || cfg->randomizer > 1) return fail(nullptr, TG_ERR_CONFIG, "randomizer %d", cfg->randomizer);
    if (cfg->autoreset < 0 || cfg->autoreset > 2) return fail(nullptr, TG_ERR_CONFIG, "autoreset %d", cfg->autoreset);
    if (cfg->holder_size < 0 || cfg->holder_size > 4) return fail(nullptr, TG_ERR_CONFIG, "holder_size %d unsupported (1..4)", cfg->holder_size);
    if (cfg->rng_mode == TG_RNG_SEQUENCE && cfg->seq_len < 1) return fail(nullptr, TG_ERR_CONFIG, "seq_len must be >= 1");
    for (int i = 0; i < 8; i++)
        if (cfg->action_map[i] < 0 || cfg->action_map[i] >= 8)
            return fail(nullptr, TG_ERR_CONFIG, "action_map[%d] = %d outside Discrete(8)", i, cfg->action_map[i]);
    if (cfg->n_pieces < 0 || cfg->n_pieces > 7) return fail(nullptr, TG_ERR_CONFIG, "n_pieces %d unsupported (custom sets hold 1..7 pieces)", cfg->n_pieces);
    int max_n = cfg->n_pieces ? 0 : 4;
    for (int p = 0; p < cfg->n_pieces; p++) {
        const int n = cfg->piece_n[p];
        if (n < 1 || n > 4) return fail(nullptr, TG_ERR_CONFIG, "tetromino %d: matrix size %d unsupported (1..4)", p, n);
        int cells = 0;
        for (int k = 0; k < n * n; k++) cells += cfg->piece_matrix[p][k] != 0;
        if (cells != 4) return fail(nullptr, TG_ERR_CONFIG, "tetromino %d has %d cells: the kernels are built on four-cell pieces", p, cells);
        if (n > max_n) max_n = n;
    }
    if (max_n != TG_PADDING)
        return fail(nullptr, TG_ERR_CONFIG, "the tetromino set's largest matrix is %d x %d: the reference would derive padding %d from it "
                    "(envs/tetris.py:130), this library is built on padding 4", max_n, max_n, max_n);
    return TG_OK;
}

// device config (sizes, strides, LUTs) of a validated tg_config; pure host arithmetic
static void derive_cfg(const tg_config* cfg, const HostTables& T, DevCfg& d) {
    memset(&d, 0, sizeof d);
    d.W = cfg->width; d.H = cfg->height; d.Wp = d.W + 2 * TG_PADDING; d.Hp = d.H + TG_PADDING; d.Q = cfg->queue_size;
    d.gravity = cfg->gravity != 0; d.autoreset = cfg->autoreset; d.rng_mode = cfg->rng_mode;
    d.terminate_on_illegal = cfg->terminate_on_illegal != 0;
    d.rand_kind = cfg->randomizer;
    const int col_bytes = d.Hp > 32 ? 8 : 4;
    d.ids_off = d.W * col_bytes;
    d.ids_words = (d.H * d.W + 7) / 8;
    int bs = round_up(d.ids_off + d.ids_words * 4, 16);
    if (((bs / 16) & 1) == 0) bs += 16;  // odd multiple of 16 B: conflict-free 128-bit shared-memory access
    d.board_stride = bs;
    d.rng_stride = cfg->rng_mode == TG_RNG_NUMPY ? 48 : 16;
    d.OB = d.Hp * d.Wp; d.OQ = 16 * d.Q; d.A = 4 * d.W; d.F = d.W + 3;
    d.inv_q = 65536u / (unsigned)d.Q + 1u;
    d.holder_size = cfg->holder_size > 0 ? cfg->holder_size : 1;
    d.OH = 16 * d.holder_size;
    d.rgb_w = d.Wp + 4 * (d.Q > d.holder_size ? d.Q : d.holder_size);   // max(holder, queue) pieces wide (wrappers/observation.py:49-58)
    d.NPC = T.np;
    d.nhalf3 = 0;
    for (int p = 0; p < 7; p++) {
        const int n = p < T.np ? T.n[p] : 3;
        d.spawn_x[p] = d.Wp / 2 - n / 2;
        d.nhalf3 |= (unsigned)(n / 2) << (3 * p);
    }
    // elif chain of Tetris.step (envs/tetris.py:223-256): first matching name wins
    static const int order[8] = {0, 1, 2, 3, 4, 6, 5, 7};  // left,right,down,cw,ccw,swap,hard_drop,no_op
    static const int ops[8] = {OP_LEFT, OP_RIGHT, OP_DOWN, OP_CW, OP_CCW, OP_HARD, OP_SWAP, OP_NOOP};
    for (int a = 0; a < 8; a++) {
        d.op_lut[a] = OP_NOOP;
        for (int k = 0; k < 8; k++)
            if (cfg->action_map[order[k]] == a) { d.op_lut[a] = (unsigned char)ops[order[k]]; break; }
        d.skipgrav[a] = (unsigned char)(a == cfg->action_map[5]);
    }
    d.act_hard = cfg->action_map[5]; d.act_noop = cfg->action_map[7];
    d.r_alife = cfg->reward_alife; d.r_go = cfg->reward_game_over; d.r_invalid = cfg->reward_invalid_action;
    d.seq_len = cfg->seq_len; d.env_id_offset = cfg->env_id_offset;
}

extern "C" int tg_create(const tg_config* cfg, int device, tg_env** out) {
    if (!cfg || !out) return fail(nullptr, TG_ERR_POINTER, "tg_create: NULL argument");
    *out = nullptr;
    int vrc = validate_cfg(cfg);
    if (vrc) return vrc;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(nullptr, TG_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return fail(nullptr, TG_ERR_ARG, "device %d of %d", device, ndev);
    tg_env* env = new tg_env();
    env->cfg = *cfg;
    env->device = device;
    env->hs_init = false;
    env->rollout_last_action = nullptr;
    env->cnn_h = env->cnn_w = 0;
    env->pool = nullptr; env->host_threads = 0; env->hring = nullptr; env->hring_bytes = 0; env->hev_init = false;
    memset(env->host_stats, 0, sizeof env->host_stats);
    memset(env->stage, 0, sizeof env->stage);
    memset(env->stage_bytes, 0, sizeof env->stage_bytes);
    cudaError_t e1 = cudaSetDevice(device);
    if (e1 != cudaSuccess) { int rc = fail(nullptr, TG_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e1)); delete env; return rc; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    env->num_sms = prop.multiProcessorCount;

    piece_set(cfg, env->tabs);
    { int trc = build_tables(nullptr, env->tabs); if (trc) { delete env; return trc; } }
    DevCfg& d = env->dev;
    derive_cfg(cfg, env->tabs, d);
    env->col64 = d.Hp > 32;

    tg_layout& L = env->layout;
    memset(&L, 0, sizeof L);
    L.width_padded = d.Wp; L.height_padded = d.Hp; L.hot_stride = 32; L.board_stride = d.board_stride;
    L.rng_stride = d.rng_stride; L.obs_board_bytes = d.OB; L.obs_holder_bytes = d.OH; L.obs_queue_bytes = d.OQ;
    L.n_placements = d.A; L.n_features = d.F; L.rgb_width = d.rgb_w;
    L.host_record_bytes = ((d.holder_size > 1 ? 16 : 12) + d.ids_words * 4 + 15) / 16 * 16;

    env->tile = 32;
    env->threads_per_env = 4;
    if (const char* t = getenv("TG_TILE")) { int v = atoi(t); if (v == 32 || v == 64 || v == 96 || v == 128) env->tile = v; }
    if (const char* t = getenv("TG_TPE")) { int v = atoi(t); if (v >= 1 && v <= 8) env->threads_per_env = v; }
    env->warp_specialized = 1;
    env->fill_warps = 4;
    if (const char* t = getenv("TG_WS")) env->warp_specialized = atoi(t) != 0;
    env->logic_warps = 2;
    env->logic_warps_set = env->fill_warps_set = 0;
    if (const char* t = getenv("TG_NF")) { int v = atoi(t); if (v >= 1 && v <= 6) { env->fill_warps = v; env->fill_warps_set = 1; } }
    if (const char* t = getenv("TG_NL")) { int v = atoi(t); if (v >= 1 && v <= 6) { env->logic_warps = v; env->logic_warps_set = 1; } }
    if (env->logic_warps + env->fill_warps > 8) env->fill_warps = 8 - env->logic_warps;
    int rc = ensure_tables(env, env->tabs, device);
    if (rc != TG_OK) { g_create_err = env->err; delete env; return rc; }
    make_expand_cfg(env->dev, env->tabs, env->xcfg);
    const int hdr = env->dev.holder_size > 1 ? 16 : 12;          // hot words 0, 2, 3 (+ the holder FIFO word)
    env->pk_bytes = (hdr + env->dev.ids_words * 4 + 15) / 16 * 16;
    env->xcfg_pk = env->xcfg;
    env->xcfg_pk.board_stride = env->pk_bytes;
    env->xcfg_pk.ids_off = hdr;
    env->xcfg_pk.hdr = hdr;
    *out = env;
    return TG_OK;
}

extern "C" int tg_destroy(tg_env* env) {
    if (!env) return TG_OK;
    DeviceGuard guard(env->device);
    if (env->hs_init) for (int i = 0; i < 3; i++) cudaStreamDestroy(env->hs[i]);
    if (env->hev_init) { for (int i = 0; i < 64; i++) cudaEventDestroy(env->hev[i]); cudaEventDestroy(env->caller_ev); }
    if (env->hring) cudaFreeHost(env->hring);
    delete env->pool;
    for (int i = 0; i < 16; i++) if (env->stage[i]) cudaFree(env->stage[i]);
    delete env;
    return TG_OK;
}

extern "C" int tg_set_option(tg_env* env, int32_t option, int64_t value) {
    if (!env) return TG_ERR_POINTER;
    switch (option) {
    case TG_OPT_TERMINATE_ON_ILLEGAL: env->dev.terminate_on_illegal = value != 0; env->cfg.terminate_on_illegal = value != 0; return TG_OK;
    case TG_OPT_HOST_THREADS: return tg_set_host_threads(env, (int32_t)value);
    default: return fail(env, TG_ERR_ARG, "tg_set_option: unknown option %d", option);
    }
}

extern "C" int tg_get_layout(const tg_env* env, tg_layout* out) {
    if (!env || !out) return TG_ERR_POINTER;
    *out = env->layout;
    return TG_OK;
}

static bool misaligned(const void* p) { return ((uintptr_t)p & 15u) != 0; }
static int check_state(tg_env* env, const tg_state& st) {
    if (!st.hot || !st.board || !st.rng) return fail(env, TG_ERR_POINTER, "state pointer is NULL");
    if (misaligned(st.hot) || misaligned(st.board) || misaligned(st.rng)) return fail(env, TG_ERR_POINTER, "state pointer not 16-byte aligned");
    if (env->cfg.rng_mode == TG_RNG_SEQUENCE && !st.piece_seq) return fail(env, TG_ERR_POINTER, "piece_seq is NULL in TG_RNG_SEQUENCE mode");
    return TG_OK;
}

// The dynamic shared-memory limit is an attribute of the KERNEL (per device), shared by every plan and every handle that
// launches the instantiation: it is only ever raised.
#include <mutex>
static int raise_smem_limit(tg_env* env, const void* kern, size_t bytes) {
    static std::mutex mu;
    static std::vector<std::pair<std::pair<int, const void*>, size_t>> seen;
    std::lock_guard<std::mutex> lk(mu);
    for (auto& e : seen)
        if (e.first.first == env->device && e.first.second == kern) {
            if (e.second >= bytes) return TG_OK;
            CUDA_TRY(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            e.second = bytes;
            return TG_OK;
        }
    CUDA_TRY(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    seen.push_back({{env->device, kern}, bytes});
    return TG_OK;
}

// ---- step / reset launcher ----------------------------------------------------------------------
// Everything that does not depend on the call (tile size, warp roles, shared-memory carve-up, occupancy, kernel instantiation,
// the diagnostic environment switches) is worked out once per (mode, with / without dict, kernel family) and cached in the
// handle: a small batch's step is bound by the host-side cost of the call, and getenv + occupancy queries were half of it.
typedef void (*step_kernel_t)(const StepParams);
template <int WT, int HT, class COLT, bool XT>
static step_kernel_t pick_kernel_x(int mode) {
    return mode == 2 ? (step_kernel_t)k_step_ws<WT, HT, COLT, 2, XT> : mode == 1 ? (step_kernel_t)k_step_ws<WT, HT, COLT, 1, XT> : (step_kernel_t)k_step_ws<WT, HT, COLT, 0, XT>;
}
// xt: custom tetromino set or holder FIFO -- the reference configuration runs the leaner instantiation (tg_device.cuh: XT)
template <int WT, int HT, class COLT>
static step_kernel_t pick_kernel(bool ws, int mode, bool xt) {
    if (!ws) return (step_kernel_t)k_step<WT, HT, COLT>;
    return xt ? pick_kernel_x<WT, HT, COLT, true>(mode) : pick_kernel_x<WT, HT, COLT, false>(mode);
}

static int build_plan(tg_env* env, StepPlan& pl, int mode, bool want_obs, int force_plain) {
    const DevCfg& d = env->dev;
    const bool ws = env->warp_specialized && !force_plain;
    int E = ws ? 32 : env->tile;
    if (ws) if (const char* t = getenv("TG_E")) { int v = atoi(t); if (v >= 8 && v <= 32 && (v & 1) == 0) E = v; }
    // without the observation dict (image / feature / grouped-feature wrappers) the image warps only store records:
    // the kernel is bound by the game logic, so run more logic warps and drop the image buffers
    int NL = ws ? (want_obs || env->logic_warps_set ? env->logic_warps : 4) : 0;
    const int NF = ws ? (want_obs || env->fill_warps_set ? env->fill_warps : 2) : 0;
    int nsx = 2;                                   // stages beyond the ones the logic warps are working on
    if (const char* t = getenv("TG_NSX")) { int v = atoi(t); if (v >= 2 && v <= 8) nsx = v; }
    int NS = ws ? NL + nsx : 2;
    StepParams& p = pl.p;
    memset(&p, 0, sizeof p);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 127) / 128 * 128; return (int)o; };
    for (;;) {
        off = 0;
        NS = ws ? NL + nsx : 2;
        // NS must be a multiple of NL: logic warp w visits the stages (w + i * NL) % NS, and it tells a stage's reload from the
        // data of the tile before only by the PARITY of the stage's mbarrier phase.  With NS % NL != 0 (it was 4 + 2) a warp comes
        // back to a stage two phases later without having waited on the phase in between; if that phase's load has not landed yet
        // (loads are issued in order but need not complete in order -- seen with a second grouped env on another stream) the
        // parity test passes on the OLD phase and the warp steps stale records.  4 logic warps -> 8 stages, 3 -> 6, 2 -> 4.
        if (ws && NL > 0) NS = (NS + NL - 1) / NL * NL;
        p.st_hot = (int)(((size_t)E * 32 + 127) / 128 * 128);
        p.st_brd = (int)(((size_t)E * d.board_stride + 16 + 127) / 128 * 128);
        p.st_rng = (int)(((size_t)E * d.rng_stride + 127) / 128 * 128);
        p.off_hot = take((size_t)NS * p.st_hot);
        p.off_brd = take((size_t)NS * p.st_brd);
        p.off_rng = take((size_t)NS * p.st_rng);
        const size_t img_e = (ws && !want_obs) ? 0 : (size_t)E;   // image buffers only when the dict is written
        p.off_iboard = take(img_e * d.OB + 16);
        p.off_imask = take(img_e * d.OB + 16);
        p.off_iholder = take(img_e * d.OH + 16);
        p.off_iqueue = take(img_e * d.OQ + 16);
        p.off_bar = take(8 * 16);
        p.off_box = take((size_t)(2 * NS + 1) * E * 4);
        p.off_tab = take(112 * 4 + 64 + 32);
        p.off_feat = take((size_t)E * 64);
        // large boards: fewer logic warps (= fewer state stages), then half tiles, while that buys another resident CTA
        // (dict-less step of the 20x40 board: 3 logic warps x 6 stages of 24-env tiles -- with 32-env tiles only 2 logic warps fit
        // next to a second CTA: 0.72 vs 0.52 G env-steps/s with the RGB image kernel behind it)
        if (ws && !want_obs && NL == 3 && E == 32 && (227 * 1024) / (off + 1024) < 2 && !getenv("TG_E")) { E = 24; continue; }
        if (ws && NL > 1 && (227 * 1024) / (off + 1024) < 2) { NL--; continue; }
        if (ws && E == 32 && (227 * 1024) / (off + 1024) < 2 && !getenv("TG_E32")) { E = 16; NL = env->logic_warps; continue; }
        if (ws || off <= 100 * 1024 || E == 32) break;
        E -= 32;
    }
    if (off > 227 * 1024) {
        if (ws) return build_plan(env, pl, mode, want_obs, 1);   // three state stages do not fit: two-stage kernel
        return fail(env, TG_ERR_CONFIG, "board too large for the shared-memory tile (%zu B)", off);
    }
    p.cfg = d;
    p.E = E;
    p.NL = NL;
    p.NS = NS;
    p.mode = mode;
    p.whole_tile_min = E / 2;
    if (const char* t = getenv("TG_WHOLE")) p.whole_tile_min = atoi(t);
    int T = ws ? 32 * (NL + NF) : E * env->threads_per_env;
    if (T > 256) T = 256;
    pl.threads = T; pl.smem = off; pl.ws = ws;
    pl.pdl = ws && !getenv("TG_NO_PDL");
    const bool xt = d.NPC != 7 || d.holder_size > 1;
    if (d.W == 10 && d.H == 20) pl.kern = (void*)pick_kernel<10, 20, uint32_t>(ws, mode, xt);
    else if (d.W == 20 && d.H == 40) pl.kern = (void*)pick_kernel<20, 40, uint64_t>(ws, mode, xt);
    else if (env->col64) pl.kern = (void*)pick_kernel<0, 0, uint64_t>(ws, mode, xt);
    else pl.kern = (void*)pick_kernel<0, 0, uint32_t>(ws, mode, xt);
    { int rc = raise_smem_limit(env, pl.kern, off); if (rc) return rc; }
    int per_sm = 0;
    CUDA_TRY(env, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pl.kern, T, off));
    if (per_sm < 1) return fail(env, TG_ERR_CONFIG, "step kernel does not fit: %zu B shared memory per CTA", off);
    pl.grid_max = (int64_t)env->num_sms * per_sm;
    pl.valid = true;
    return TG_OK;
}

// `p` carries the per-call pointers (and p.mode); the plan supplies the rest
static int launch_step(tg_env* env, StepParams& p, cudaStream_t s) {
    const bool want_obs = p.o_board != nullptr;
    StepPlan& pl = env->plans[p.mode * 2 + (want_obs ? 1 : 0)];
    if (!pl.valid) { int rc = build_plan(env, pl, p.mode, want_obs, 0); if (rc) return rc; }
    const StepParams& q = pl.p;
    p.cfg = env->dev;           // (env_id_offset may differ per call: tg_step_host steps chunks)
    p.E = q.E; p.NL = q.NL; p.NS = q.NS; p.whole_tile_min = q.whole_tile_min;
    p.off_hot = q.off_hot; p.off_brd = q.off_brd; p.off_rng = q.off_rng; p.off_iboard = q.off_iboard; p.off_imask = q.off_imask;
    p.off_iholder = q.off_iholder; p.off_iqueue = q.off_iqueue; p.off_bar = q.off_bar; p.off_box = q.off_box; p.off_tab = q.off_tab;
    p.off_feat = q.off_feat; p.st_hot = q.st_hot; p.st_brd = q.st_brd; p.st_rng = q.st_rng;
    const int64_t ntiles = (p.n + p.E - 1) / p.E;
    const int64_t grid = pl.grid_max < ntiles ? pl.grid_max : ntiles;
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3((unsigned)pl.threads); lc.dynamicSmemBytes = pl.smem; lc.stream = s;
    cudaLaunchAttribute at[1];
    if (pl.pdl) {
        // programmatic stream serialization: this grid's prologue may overlap the tail of the previous kernel on the stream
        // (k_step_ws waits with griddepcontrol.wait before it touches global memory)
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = 1;
    }
    void* args[1] = {(void*)&p};
    CUDA_TRY(env, cudaLaunchKernelExC(&lc, pl.kern, args));
    return TG_OK;
}

static int check_obs(tg_env* env, const tg_obs& o) {
    if (!o.board || !o.mask || !o.holder || !o.queue) return fail(env, TG_ERR_POINTER, "observation pointer is NULL");
    if (misaligned(o.board) || misaligned(o.mask) || misaligned(o.holder) || misaligned(o.queue))
        return fail(env, TG_ERR_POINTER, "observation pointer not 16-byte aligned");
    return TG_OK;
}

extern "C" int tg_reset(tg_env* env, tg_state st, int64_t n, const uint64_t* d_seeds, const uint8_t* d_reset_mask,
                        tg_obs obs, void* stream) {
    if (!env) return TG_ERR_POINTER;
    if (n <= 0) return fail(env, TG_ERR_ARG, "n must be positive");
    int rc = check_state(env, st); if (rc) return rc;
    rc = check_obs(env, obs); if (rc) return rc;
    ON_DEVICE(env);
    StepParams p;
    memset(&p, 0, sizeof p);
    p.n = n; p.hot = (uint8_t*)st.hot; p.board = (uint8_t*)st.board; p.rng = (uint8_t*)st.rng; p.seq = st.piece_seq;
    p.seeds = d_seeds; p.reset_mask = d_reset_mask;
    p.o_board = obs.board; p.o_mask = obs.mask; p.o_holder = obs.holder; p.o_queue = obs.queue;
    p.mode = 1;
    return launch_step(env, p, (cudaStream_t)stream);
}

extern "C" int tg_step(tg_env* env, tg_state st, int64_t n, const int32_t* d_actions, tg_obs obs, tg_step_out out,
                       tg_stats* d_stats, void* stream) {
    if (!env) return TG_ERR_POINTER;
    if (n <= 0) return fail(env, TG_ERR_ARG, "n must be positive");
    int rc = check_state(env, st); if (rc) return rc;
    if (obs.board || obs.mask || obs.holder || obs.queue) { rc = check_obs(env, obs); if (rc) return rc; }  // all NULL = no dict
    if (!d_actions || !out.reward || !out.terminated || !out.truncated || !out.lines)
        return fail(env, TG_ERR_POINTER, "actions / step outputs pointer is NULL");
    ON_DEVICE(env);
    StepParams p;
    memset(&p, 0, sizeof p);
    p.n = n; p.hot = (uint8_t*)st.hot; p.board = (uint8_t*)st.board; p.rng = (uint8_t*)st.rng; p.seq = st.piece_seq;
    p.actions = d_actions;
    p.o_board = obs.board; p.o_mask = obs.mask; p.o_holder = obs.holder; p.o_queue = obs.queue;
    p.reward = out.reward; p.terminated = out.terminated; p.truncated = out.truncated; p.lines = out.lines;
    p.stats = (double*)d_stats;
    p.mode = 0;
    return launch_step(env, p, (cudaStream_t)stream);
}

// ---- K steps per call (small batches: records resident in shared memory for all K steps) ------------------------------------------
typedef void (*stepn_kernel_t)(const StepNParams);
extern "C" int tg_step_n(tg_env* env, tg_state st, int64_t n, int32_t k_steps, const int32_t* d_actions, tg_obs obs, int64_t obs_stride,
                         tg_step_out out, int64_t out_stride, tg_stats* d_stats, void* stream) {
    if (!env) return TG_ERR_POINTER;
    if (n <= 0 || k_steps <= 0) return fail(env, TG_ERR_ARG, "n and k_steps must be positive");
    if ((obs_stride != 0 && obs_stride < n) || (out_stride != 0 && out_stride < n)) return fail(env, TG_ERR_ARG, "step strides must be 0 or >= n");
    int rc = check_state(env, st); if (rc) return rc;
    rc = check_obs(env, obs); if (rc) return rc;
    if (!d_actions || !out.reward || !out.terminated || !out.truncated || !out.lines)
        return fail(env, TG_ERR_POINTER, "actions / step outputs pointer is NULL");
    ON_DEVICE(env);
    const DevCfg& d = env->dev;
    cudaStream_t s = (cudaStream_t)stream;
    // ---- resident plan: the whole batch's records stay in shared memory ----
    const int E = 32, NF = 4, NLR = 2, T = 32 * (NLR + NF);
    const int64_t ntiles = (n + E - 1) / E;
    StepNParams q;
    memset(&q, 0, sizeof q);
    StepParams& p = q.sp;
    size_t smem = 0;
    int64_t grid = 0;
    bool resident = env->warp_specialized && (((uintptr_t)obs.board | (uintptr_t)obs.mask) & 15) == 0 && (d.OB & 15) == 0 &&
                    (obs_stride * d.OB) % 16 == 0 && !getenv("TG_NO_RESIDENT");
    if (resident) {
        resident = false;
        // (fewer CTAs per SM would make room for more resident tiles, but the logic warps of one CTA then serialise them: beyond
        // ~10^5 envs at 10x20 one launch per step is faster)
        for (int per_sm = 3; per_sm >= 2 && !resident; per_sm--) {
            grid = (int64_t)env->num_sms * per_sm;
            if (grid > ntiles) grid = ntiles;
            const int64_t TL = (ntiles + grid - 1) / grid;
            if (TL > 8) continue;
            size_t off = 0;
            auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 127) / 128 * 128; return (int)o; };
            p.st_hot = (int)(((size_t)E * 32 + 127) / 128 * 128);
            p.st_brd = (int)(((size_t)E * d.board_stride + 16 + 127) / 128 * 128);
            p.st_rng = (int)(((size_t)E * d.rng_stride + 127) / 128 * 128);
            p.off_hot = take((size_t)TL * p.st_hot);
            p.off_brd = take((size_t)TL * p.st_brd);
            p.off_rng = take((size_t)TL * p.st_rng);
            p.off_iboard = take((size_t)E * d.OB + 16);
            p.off_imask = take((size_t)E * d.OB + 16);
            p.off_iholder = take((size_t)E * d.OH + 16);
            p.off_iqueue = take((size_t)E * d.OQ + 16);
            p.off_bar = take(16);
            p.off_box = take((size_t)(TL + 1) * E * 4);
            p.off_tab = take(112 * 4 + 64 + 32);
            q.off_cnt = take(16);
            q.off_dirty = take((size_t)TL * E * 4);
            if (off + 1024 <= (size_t)(227 * 1024) / per_sm) { resident = true; q.TL = (int)TL; smem = off; }
        }
    }
    if (resident) {
        p.cfg = d; p.n = n; p.E = E; p.mode = 0;
        p.hot = (uint8_t*)st.hot; p.board = (uint8_t*)st.board; p.rng = (uint8_t*)st.rng; p.seq = st.piece_seq;
        p.actions = d_actions;
        p.o_board = obs.board; p.o_mask = obs.mask; p.o_holder = obs.holder; p.o_queue = obs.queue;
        p.reward = out.reward; p.terminated = out.terminated; p.truncated = out.truncated; p.lines = out.lines;
        p.stats = (double*)d_stats;
        q.K = k_steps; q.NL = NLR; q.obs_stride = obs_stride; q.out_stride = out_stride;
        stepn_kernel_t kern;
        if (d.W == 10 && d.H == 20) kern = k_step_resident<10, 20, uint32_t>;
        else if (d.W == 20 && d.H == 40) kern = k_step_resident<20, 40, uint64_t>;
        else if (env->col64) kern = k_step_resident<0, 0, uint64_t>;
        else kern = k_step_resident<0, 0, uint32_t>;
        rc = raise_smem_limit(env, (const void*)kern, smem); if (rc) return rc;
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof lc);
        lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3((unsigned)T); lc.dynamicSmemBytes = smem; lc.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = 1;
        CUDA_TRY(env, cudaLaunchKernelEx(&lc, kern, q));
        return TG_OK;
    }
    // ---- large batches: one launch per step (back to back, programmatic dependent launches) ----
    for (int k = 0; k < k_steps; k++) {
        StepParams pk;
        memset(&pk, 0, sizeof pk);
        pk.n = n; pk.hot = (uint8_t*)st.hot; pk.board = (uint8_t*)st.board; pk.rng = (uint8_t*)st.rng; pk.seq = st.piece_seq;
        pk.actions = d_actions + (int64_t)k * n;
        const int64_t ob = (int64_t)k * obs_stride, oo = (int64_t)k * out_stride;
        pk.o_board = obs.board + ob * d.OB; pk.o_mask = obs.mask + ob * d.OB; pk.o_holder = obs.holder + ob * d.OH; pk.o_queue = obs.queue + ob * d.OQ;
        pk.reward = out.reward + oo; pk.terminated = out.terminated + oo; pk.truncated = out.truncated + oo; pk.lines = out.lines + oo;
        pk.stats = (double*)d_stats;
        pk.mode = 0;
        rc = launch_step(env, pk, s); if (rc) return rc;
    }
    return TG_OK;
}

// ---- host-buffer step (e2e path): H2D actions, step, D2H observation dict + 5-tuple -----------------
static int ensure_stage(tg_env* env, int slot, size_t bytes) {
    if (env->stage_bytes[slot] >= bytes) return TG_OK;
    if (env->stage[slot]) cudaFree(env->stage[slot]);
    env->stage[slot] = nullptr; env->stage_bytes[slot] = 0;
    CUDA_TRY(env, cudaMalloc(&env->stage[slot], bytes));
    env->stage_bytes[slot] = bytes;
    return TG_OK;
}

// make the internal copy streams wait for the work the caller enqueued on `stream` before this call (reset / step / set_state)
static int host_streams_begin(tg_env* env, cudaStream_t caller) {
    if (!env->hs_init) {
        for (int i = 0; i < 3; i++) CUDA_TRY(env, cudaStreamCreateWithFlags(&env->hs[i], cudaStreamNonBlocking));
        env->hs_init = true;
    }
    if (!env->hev_init) {
        for (int i = 0; i < 64; i++) CUDA_TRY(env, cudaEventCreateWithFlags(&env->hev[i], cudaEventDisableTiming));
        CUDA_TRY(env, cudaEventCreateWithFlags(&env->caller_ev, cudaEventDisableTiming));
        env->hev_init = true;
    }
    CUDA_TRY(env, cudaEventRecord(env->caller_ev, caller));
    for (int i = 0; i < 3; i++) CUDA_TRY(env, cudaStreamWaitEvent(env->hs[i], env->caller_ev, 0));
    return TG_OK;
}

static double wall_now() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

// one chunk [b, b + m) of the env range on stream s: actions H2D, step kernel (obs = all-NULL -> no dict), scalars D2H
static int host_chunk_step(tg_env* env, const tg_state& st, int64_t b, int64_t m, const int32_t* h_actions, int32_t* s_act,
                           const tg_obs& d_obs, float* s_rew, int32_t* s_lines, uint8_t* s_term, uint8_t* s_trunc, cudaStream_t s) {
    const DevCfg& d = env->dev;
    CUDA_TRY(env, cudaMemcpyAsync(s_act + b, h_actions + b, (size_t)m * 4, cudaMemcpyHostToDevice, s));
    StepParams p;
    memset(&p, 0, sizeof p);
    p.n = m;
    p.hot = (uint8_t*)st.hot + b * 32; p.board = (uint8_t*)st.board + b * d.board_stride; p.rng = (uint8_t*)st.rng + b * d.rng_stride;
    p.seq = st.piece_seq ? st.piece_seq + b * d.seq_len : nullptr;
    p.actions = s_act + b;
    if (d_obs.board) { p.o_board = d_obs.board + b * d.OB; p.o_mask = d_obs.mask + b * d.OB; p.o_holder = d_obs.holder + b * d.OH; p.o_queue = d_obs.queue + b * d.OQ; }
    p.reward = s_rew + b; p.terminated = s_term + b; p.truncated = s_trunc + b; p.lines = s_lines + b;
    p.mode = 0;
    DevCfg saved = env->dev;
    env->dev.env_id_offset += (unsigned long long)b;
    int rc = launch_step(env, p, s);
    env->dev = saved;
    return rc;
}

struct ExpandJob {
    const tgh::ExpandCfg* cfg;
    tgh::ExpandArgs args;     // pointers biased so that env index = global env index
    int64_t base, count, per; // items cover [base + i * per, base + min((i + 1) * per, count))
};
static void expand_item(void* ctx, int64_t i) {
    const ExpandJob& j = *(const ExpandJob*)ctx;
    const int64_t e0 = j.base + i * j.per, e1 = j.base + (((i + 1) * j.per < j.count) ? (i + 1) * j.per : j.count);
    tgh::expand_range(*j.cfg, j.args, e0, e1);
}
static void run_expand(tgh::Pool& pool, const tgh::ExpandCfg& cfg, const tgh::ExpandArgs& args, int64_t base, int64_t count) {
    ExpandJob j;
    j.cfg = &cfg; j.args = args; j.base = base; j.count = count;
    int64_t items = (int64_t)pool.size() * 2;
    int64_t per = (count + items - 1) / items;
    per = (per + 127) / 128 * 128;      // whole cache lines in every output array
    if (per < 128) per = 128;
    j.per = per;
    pool.run((count + per - 1) / per, expand_item, &j);
}

extern "C" int tg_set_host_threads(tg_env* env, int32_t threads) {
    if (!env) return TG_ERR_POINTER;
    if (threads < 0 || threads > 1024) return fail(env, TG_ERR_ARG, "host threads %d", threads);
    if (env->pool && env->pool->size() != (threads ? threads : tgh::default_threads())) { delete env->pool; env->pool = nullptr; }
    env->host_threads = threads;
    return TG_OK;
}

extern "C" int tg_host_stats(tg_env* env, double* out4) {
    if (!env || !out4) return TG_ERR_POINTER;
    for (int i = 0; i < 4; i++) out4[i] = env->host_stats[i];
    return TG_OK;
}

extern "C" int tg_step_host(tg_env* env, tg_state st, int64_t n, const int32_t* h_actions, tg_obs h_obs, tg_step_out h_out,
                            int32_t mode, void* stream) {
    if (!env) return TG_ERR_POINTER;
    if (n <= 0) return fail(env, TG_ERR_ARG, "n must be positive");
    if (mode != TG_HOST_DMA && mode != TG_HOST_COMPACT) return fail(env, TG_ERR_ARG, "tg_step_host: mode %d", mode);
    if (!h_actions || !h_obs.board || !h_obs.mask || !h_obs.holder || !h_obs.queue || !h_out.reward || !h_out.terminated ||
        !h_out.truncated || !h_out.lines)
        return fail(env, TG_ERR_POINTER, "host buffer is NULL");
    int rc = check_state(env, st); if (rc) return rc;
    ON_DEVICE(env);
    const double t_begin = wall_now();
    rc = host_streams_begin(env, (cudaStream_t)stream); if (rc) return rc;
    const DevCfg& d = env->dev;
    auto r16 = [](size_t v) { return (v + 15) / 16 * 16; };
    const size_t o_lines = r16((size_t)n * 4), o_term = o_lines + r16((size_t)n * 4), o_trunc = o_term + r16((size_t)n);
    const size_t out_total = o_trunc + r16((size_t)n);
    rc = ensure_stage(env, 0, (size_t)n * 4); if (rc) return rc;   // actions
    rc = ensure_stage(env, 2, out_total); if (rc) return rc;       // reward, lines, terminated, truncated
    uint8_t* so = (uint8_t*)env->stage[2];
    float* s_rew = (float*)so;
    int32_t* s_lines = (int32_t*)(so + o_lines);
    uint8_t* s_term = so + o_term;
    uint8_t* s_trunc = so + o_trunc;
    int32_t* s_act = (int32_t*)env->stage[0];

    if (mode == TG_HOST_DMA) {
        // the whole dict is produced on the device and copied: chunk the env range so that the D2H of chunk k overlaps the
        // kernel of chunk k+1
        int NCH = 3;
        if (const char* t = getenv("TG_HOST_CHUNKS")) { int v = atoi(t); if (v >= 1 && v <= 64) NCH = v; }
        int64_t chunk = (n + NCH - 1) / NCH;
        chunk = (chunk + 127) / 128 * 128;  // keeps every chunk's base 16-byte aligned in all arrays
        const size_t o_mask = r16((size_t)n * d.OB), o_holder = o_mask + r16((size_t)n * d.OB), o_queue = o_holder + (size_t)n * d.OH;
        const size_t obs_total = o_queue + (size_t)n * d.OQ;
        rc = ensure_stage(env, 1, obs_total); if (rc) return rc;       // observation dict
        tg_obs d_obs;
        d_obs.board = (uint8_t*)env->stage[1]; d_obs.mask = d_obs.board + o_mask; d_obs.holder = d_obs.board + o_holder; d_obs.queue = d_obs.board + o_queue;
        int k = 0;
        for (int64_t b = 0; b < n; b += chunk, k++) {
            int64_t m = n - b < chunk ? n - b : chunk;
            cudaStream_t s = env->hs[k % 3];
            rc = host_chunk_step(env, st, b, m, h_actions, s_act, d_obs, s_rew, s_lines, s_term, s_trunc, s); if (rc) return rc;
            CUDA_TRY(env, cudaMemcpyAsync(h_obs.board + b * d.OB, d_obs.board + b * d.OB, (size_t)m * d.OB, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(env, cudaMemcpyAsync(h_obs.mask + b * d.OB, d_obs.mask + b * d.OB, (size_t)m * d.OB, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(env, cudaMemcpyAsync(h_obs.holder + b * d.OH, d_obs.holder + b * d.OH, (size_t)m * d.OH, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(env, cudaMemcpyAsync(h_obs.queue + b * d.OQ, d_obs.queue + b * d.OQ, (size_t)m * d.OQ, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(env, cudaMemcpyAsync(h_out.reward + b, s_rew + b, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(env, cudaMemcpyAsync(h_out.lines + b, s_lines + b, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(env, cudaMemcpyAsync(h_out.terminated + b, s_term + b, (size_t)m, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(env, cudaMemcpyAsync(h_out.truncated + b, s_trunc + b, (size_t)m, cudaMemcpyDeviceToHost, s));
        }
        const double t_w = wall_now();
        for (int i = 0; i < 3; i++) CUDA_TRY(env, cudaStreamSynchronize(env->hs[i]));
        const double t_end = wall_now();
        env->host_stats[0] = t_end - t_begin; env->host_stats[1] = t_end - t_w; env->host_stats[2] = 0; env->host_stats[3] = k;
        return TG_OK;
    }

    // ---- compact: the step runs WITHOUT the observation dict; the packed records (hot + board) cross the link and the dict is
    // rebuilt in the caller's arrays by the host pool while later chunks are still stepping / copying -------------------------
    if (!env->pool) env->pool = new tgh::Pool(env->host_threads ? env->host_threads : tgh::default_threads());
    int64_t chunk = 131072;
    if (const char* t = getenv("TG_HOST_CHUNK")) { long v = atol(t); if (v >= 128) chunk = v; }
    chunk = (chunk + 127) / 128 * 128;
    if (chunk > n) chunk = (n + 127) / 128 * 128;
    const int64_t NCH = (n + chunk - 1) / chunk;
    int NB = 4;                                   // ring slots of packed records in pinned host memory
    if (const char* t = getenv("TG_HOST_RING")) { int v = atoi(t); if (v >= 2 && v <= 32) NB = v; }
    if (NB > NCH) NB = (int)NCH;
    const int pk = env->pk_bytes;
    const size_t slot = ((size_t)chunk * pk + 64 + 63) / 64 * 64;
    if (env->hring_bytes < slot * NB) {
        if (env->hring) cudaFreeHost(env->hring);
        env->hring = nullptr; env->hring_bytes = 0;
        CUDA_TRY(env, cudaHostAlloc(&env->hring, slot * NB, cudaHostAllocDefault));
        env->hring_bytes = slot * NB;
    }
    rc = ensure_stage(env, 5, slot * NB); if (rc) return rc;      // packed records of the chunks in flight (device side of the ring)
    memset(h_out.truncated, 0, (size_t)n);        // always False (envs/tetris.py:219)
    tg_obs no_obs;
    memset(&no_obs, 0, sizeof no_obs);
    double t_wait = 0, t_exp = 0;
    int64_t next_enq = 0;
    for (int64_t c = 0; c < NCH; c++) {
        for (; next_enq < NCH && next_enq < c + NB; next_enq++) {
            const int64_t b = next_enq * chunk, m = n - b < chunk ? n - b : chunk;
            cudaStream_t s = env->hs[next_enq % 3];
            uint8_t* ring = (uint8_t*)env->hring + (size_t)(next_enq % NB) * slot;
            rc = host_chunk_step(env, st, b, m, h_actions, s_act, no_obs, s_rew, s_lines, s_term, s_trunc, s); if (rc) return rc;
            uint8_t* dpk = (uint8_t*)env->stage[5] + (size_t)(next_enq % NB) * slot;
            {
                const int64_t words = m * (pk / 4);
                k_pack_host<<<(unsigned)((words + 255) / 256), 256, 0, s>>>((const uint8_t*)st.hot + b * 32, (const uint8_t*)st.board + b * d.board_stride,
                                                                          d.board_stride, d.ids_off, d.ids_words, pk / 4, env->xcfg_pk.hdr / 4, m, (uint32_t*)dpk);
                CUDA_TRY(env, cudaGetLastError());
            }
            CUDA_TRY(env, cudaMemcpyAsync(ring, dpk, (size_t)m * pk, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(env, cudaMemcpyAsync(h_out.reward + b, s_rew + b, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(env, cudaMemcpyAsync(h_out.lines + b, s_lines + b, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(env, cudaMemcpyAsync(h_out.terminated + b, s_term + b, (size_t)m, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(env, cudaEventRecord(env->hev[next_enq % NB], s));
        }
        const double t0 = wall_now();
        CUDA_TRY(env, cudaEventSynchronize(env->hev[c % NB]));
        const double t1 = wall_now();
        const int64_t b = c * chunk, m = n - b < chunk ? n - b : chunk;
        const uint8_t* ring = (const uint8_t*)env->hring + (size_t)(c % NB) * slot;
        tgh::ExpandArgs xa;
        xa.hot = nullptr; xa.board = ring - b * (int64_t)pk;   // biased: indexed by the global env id
        xa.board_end = ring + slot;
        xa.o_board = h_obs.board; xa.o_mask = h_obs.mask; xa.o_holder = h_obs.holder; xa.o_queue = h_obs.queue;
        run_expand(*env->pool, env->xcfg_pk, xa, b, m);
        t_wait += t1 - t0; t_exp += wall_now() - t1;
    }
    for (int i = 0; i < 3; i++) CUDA_TRY(env, cudaStreamSynchronize(env->hs[i]));
    env->host_stats[0] = wall_now() - t_begin; env->host_stats[1] = t_wait; env->host_stats[2] = t_exp; env->host_stats[3] = (double)NCH;
    return TG_OK;
}

// Host-memory write ceiling: `threads` host threads fill h_dst (64-byte aligned, `bytes` long) with streaming stores `reps` times;
// returns GB/s in *gb_per_s.  The output traffic of the TG_HOST_COMPACT expansion is bounded by this number (bench.py: e2e.roofline).
struct FillJob { uint8_t* dst; size_t per, bytes; int value; };
static void fill_item(void* ctx, int64_t i) {
    const FillJob& j = *(const FillJob*)ctx;
    const size_t o = (size_t)i * j.per, m = o + j.per <= j.bytes ? j.per : j.bytes - o;
    tgh::stream_fill(j.dst + o, m, j.value);
}
extern "C" int tg_host_membw(void* h_dst, int64_t bytes, int32_t threads, int32_t reps, double* gb_per_s) {
    if (!h_dst || !gb_per_s || bytes < 4096 || ((uintptr_t)h_dst & 63)) return fail(nullptr, TG_ERR_POINTER, "tg_host_membw: bad buffer");
    tgh::Pool pool(threads > 0 ? threads : tgh::default_threads());
    FillJob j;
    j.dst = (uint8_t*)h_dst; j.bytes = (size_t)bytes & ~(size_t)63;
    j.per = (j.bytes / ((size_t)pool.size() * 4) + 4095) & ~(size_t)4095;
    const int64_t items = (int64_t)((j.bytes + j.per - 1) / j.per);
    j.value = 0;
    pool.run(items, fill_item, &j);                 // warm-up: page faults, thread start
    const double t0 = wall_now();
    for (int r = 0; r < (reps > 0 ? reps : 1); r++) { j.value = r; pool.run(items, fill_item, &j); }
    *gb_per_s = (double)j.bytes * (reps > 0 ? reps : 1) / (wall_now() - t0) / 1e9;
    return TG_OK;
}

// Pure host entry point of the same expansion: packed records in host memory -> observation dict (no device involved; used by
// callers that fetch the packed state themselves, and by the CPU tests of the expansion).
extern "C" int tg_host_expand(const tg_config* cfg, int64_t n, const void* h_hot, const void* h_board, tg_obs h_obs, int32_t threads) {
    if (!cfg || !h_hot || !h_board || !h_obs.board || !h_obs.mask || !h_obs.holder || !h_obs.queue)
        return fail(nullptr, TG_ERR_POINTER, "tg_host_expand: NULL argument");
    if (n <= 0) return fail(nullptr, TG_ERR_ARG, "n must be positive");
    int rc = validate_cfg(cfg); if (rc) return rc;
    DevCfg d;
    HostTables T;
    piece_set(cfg, T);
    rc = build_tables(nullptr, T); if (rc) return rc;
    derive_cfg(cfg, T, d);
    tgh::ExpandCfg x;
    make_expand_cfg(d, T, x);
    tgh::ExpandArgs xa;
    xa.hot = (const uint8_t*)h_hot; xa.board = (const uint8_t*)h_board;
    xa.board_end = xa.board + (size_t)n * d.board_stride;
    xa.o_board = h_obs.board; xa.o_mask = h_obs.mask; xa.o_holder = h_obs.holder; xa.o_queue = h_obs.queue;
    tgh::Pool pool(threads > 0 ? threads : tgh::default_threads());
    run_expand(pool, x, xa, 0, n);
    return TG_OK;
}

// ---- numpy-exact seeding ------------------------------------------------------------------------
extern "C" int tg_seed_numpy(tg_env* env, tg_state st, int64_t n, const uint64_t* d_pcg, const uint8_t* d_mask, void* stream) {
    if (!env) return TG_ERR_POINTER;
    if (env->cfg.rng_mode != TG_RNG_NUMPY) return fail(env, TG_ERR_ARG, "tg_seed_numpy needs rng_mode = TG_RNG_NUMPY");
    if (!d_pcg || !st.rng) return fail(env, TG_ERR_POINTER, "NULL pointer");
    ON_DEVICE(env);
    int T = 256;
    k_seed_numpy<<<(unsigned)((n + T - 1) / T), T, 0, (cudaStream_t)stream>>>((uint8_t*)st.rng, env->dev.rng_stride, n, d_pcg, d_mask);
    CUDA_TRY(env, cudaGetLastError());
    return TG_OK;
}

extern "C" int tg_seed_numpy_seeds(tg_env* env, tg_state st, int64_t n, const uint64_t* d_seeds, const uint8_t* d_mask, void* stream) {
    if (!env) return TG_ERR_POINTER;
    if (env->cfg.rng_mode != TG_RNG_NUMPY) return fail(env, TG_ERR_ARG, "tg_seed_numpy_seeds needs rng_mode = TG_RNG_NUMPY");
    if (!d_seeds || !st.rng) return fail(env, TG_ERR_POINTER, "NULL pointer");
    ON_DEVICE(env);
    int T = 256;
    k_seed_numpy_seeds<<<(unsigned)((n + T - 1) / T), T, 0, (cudaStream_t)stream>>>((uint8_t*)st.rng, env->dev.rng_stride, n, d_seeds, d_mask);
    CUDA_TRY(env, cudaGetLastError());
    return TG_OK;
}

// ---- state access -----------------------------------------------------------------------------------
extern "C" int tg_get_state(tg_env* env, tg_state st, int64_t n, uint8_t* d_board, int32_t* d_scalars, void* stream) {
    if (!env) return TG_ERR_POINTER;
    int rc = check_state(env, st); if (rc) return rc;
    ON_DEVICE(env);
    int T = 128;
    if (env->col64) k_get_state<uint64_t><<<(unsigned)((n + T - 1) / T), T, 0, (cudaStream_t)stream>>>(env->dev, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_board, d_scalars);
    else k_get_state<uint32_t><<<(unsigned)((n + T - 1) / T), T, 0, (cudaStream_t)stream>>>(env->dev, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_board, d_scalars);
    CUDA_TRY(env, cudaGetLastError());
    return TG_OK;
}

extern "C" int tg_set_state(tg_env* env, tg_state st, int64_t n, const uint8_t* d_board, const int32_t* d_scalars,
                            const uint8_t* d_mask, void* stream) {
    if (!env) return TG_ERR_POINTER;
    int rc = check_state(env, st); if (rc) return rc;
    ON_DEVICE(env);
    int T = 128;
    if (env->col64) k_set_state<uint64_t><<<(unsigned)((n + T - 1) / T), T, 0, (cudaStream_t)stream>>>(env->dev, n, (uint8_t*)st.hot, (uint8_t*)st.board, d_board, d_scalars, d_mask);
    else k_set_state<uint32_t><<<(unsigned)((n + T - 1) / T), T, 0, (cudaStream_t)stream>>>(env->dev, n, (uint8_t*)st.hot, (uint8_t*)st.board, d_board, d_scalars, d_mask);
    CUDA_TRY(env, cudaGetLastError());
    return TG_OK;
}

// ---- functional facade ------------------------------------------------------------------------------------
extern "C" int tg_fn_step(int32_t width, int32_t height, int32_t queue_size, int32_t gravity, int64_t n,
                          const int8_t* d_board_in, const int32_t* d_scalars_in, const int32_t* d_actions,
                          const uint8_t* d_piece_seq, int64_t seq_len, int8_t* d_board_out, int32_t* d_scalars_out,
                          int8_t* d_obs, float* d_reward, uint8_t* d_terminated, int32_t* d_lines, void* stream) {
    if (width < 4 || width + 2 * TG_PADDING > 32 || height < 4 || height + TG_PADDING > 64 || queue_size < 1 || queue_size > 16)
        return fail(nullptr, TG_ERR_CONFIG, "tg_fn_step: unsupported width/height/queue_size");
    if (queue_size > 7 && !d_piece_seq) return fail(nullptr, TG_ERR_CONFIG, "tg_fn_step: queue_size doubles as the number of piece types (<= 7)");
    if (n <= 0 || !d_board_in || !d_board_out || !d_scalars_in || !d_scalars_out) return fail(nullptr, TG_ERR_POINTER, "tg_fn_step: NULL pointer");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(nullptr, TG_ERR_CUDA, "tg_fn_step: no CUDA device (no CPU fallback)");
    {   // the functional env always plays the reference's seven pieces: make that set resident on the current device
        static HostTables std_tabs;
        static bool built = false;
        if (!built) { piece_set(nullptr, std_tabs); build_tables(nullptr, std_tabs); built = true; }
        tg_env tmp;
        int rc = ensure_tables(&tmp, std_tabs, dev);
        if (rc) return fail(nullptr, rc, "tg_fn_step: %s", tmp.err.c_str());
    }
    FnParams p;
    memset(&p, 0, sizeof p);
    p.W = width; p.H = height; p.Wp = width + 2 * TG_PADDING; p.Hp = height + TG_PADDING; p.Q = queue_size; p.gravity = gravity != 0;
    p.invW = 65536u / (uint32_t)width + 1u;
    p.n = n; p.board_in = d_board_in; p.board_out = d_board_out; p.sc_in = d_scalars_in; p.sc_out = d_scalars_out;
    p.actions = d_actions; p.seq = d_piece_seq; p.seq_len = seq_len;
    p.obs = d_obs; p.reward = d_reward; p.terminated = d_terminated; p.lines = d_lines;
    // the tile kernel (bulk copies of whole tiles) needs 16-byte aligned arrays; TG_FN_V1=1 keeps the thread-per-env kernel
    const bool aligned = (((uintptr_t)d_board_in | (uintptr_t)d_board_out | (uintptr_t)d_scalars_in | (uintptr_t)d_scalars_out | (uintptr_t)d_obs) & 15) == 0;
    if (aligned && p.H <= 64 && !getenv("TG_FN_V1")) {
        // envs per tile: 16 (sixteen 128-thread CTAs per SM) unless TG_FN_E=32 / 8
        static const int fn_e = [] { const char* v = getenv("TG_FN_E"); const int e = v ? atoi(v) : 16; return (e == 8 || e == 32) ? e : 16; }();
        const FnTileSmem m = fn_tile_smem(p.Hp * p.Wp, p.H * p.W, FN_S + p.Q, fn_e);
        auto kern = fn_e == 32 ? k_fn_step_tile<32> : (fn_e == 8 ? k_fn_step_tile<8> : k_fn_step_tile<16>);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, m.bytes) != cudaSuccess)
            return fail(nullptr, TG_ERR_CONFIG, "tg_fn_step: board too large for the shared-memory tile (%d B)", m.bytes);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        kern<<<(unsigned)((n + fn_e - 1) / fn_e), fn_e * 8, (size_t)m.bytes, (cudaStream_t)stream>>>(p);
        cudaError_t e2 = cudaGetLastError();
        if (e2 != cudaSuccess) return fail(nullptr, TG_ERR_CUDA, "k_fn_step_tile: %s", cudaGetErrorString(e2));
        return TG_OK;
    }
    const int T = 64;
    int bstr = (p.Hp * p.Wp + 3) / 4 * 4;
    if (((bstr / 4) & 1) == 0) bstr += 4;                                  // odd word stride: conflict-free per-env access
    const size_t smem = ((size_t)T * bstr + 15) / 16 * 16 + (size_t)T * p.H * p.W;
    if (cudaFuncSetAttribute(k_fn_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return fail(nullptr, TG_ERR_CONFIG, "tg_fn_step: board too large for the shared-memory tile (%zu B)", smem);
    k_fn_step<<<(unsigned)((n + T - 1) / T), T, smem, (cudaStream_t)stream>>>(p, bstr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(nullptr, TG_ERR_CUDA, "k_fn_step: %s", cudaGetErrorString(e));
    return TG_OK;
}

// Launch with programmatic stream serialization (the kernel must execute griddepcontrol.wait before it touches global
// memory): its CTAs may be scheduled while the previous kernel on the stream drains.  TG_NO_PDL=1: plain launch.
template <class K, class... A>
static cudaError_t launch_pdl(K kern, unsigned grid, unsigned block, size_t smem, cudaStream_t s, A... args) {
    if (getenv("TG_NO_PDL")) {
        kern<<<grid, block, smem, s>>>(args...);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3(grid); lc.blockDim = dim3(block); lc.dynamicSmemBytes = smem; lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    return cudaLaunchKernelEx(&lc, kern, args...);
}

// ---- wrappers (tg_wrappers.cuh) ------------------------------------------------------------------------
#include "tg_wrappers.cuh"
#include "tg_cnn.cuh"
