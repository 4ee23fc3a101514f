// tg_device.cuh -- device-side building blocks of the B200 Tetris simulator.
//
// Data layout in HBM (per env, structure-of-records; all strides are multiples of 16 B so a tile
// of consecutive envs is one contiguous TMA bulk copy):
//   hot   [32 B]  w0 = x|y|piece|rot|holder|flags, w1 = 7-bag nibbles + index, w2:w3 = queue nibbles,
//                 w4..w6 = episode return / length / lines, w7 spare
//   board [BS B]  W column bitboards (bit r = row r occupied, floor rows H..H_pad-1 always set;
//                 u32 when H_pad <= 32 else u64)  followed by the piece-id plane, nibble-packed
//                 row-major (row r = nibbles [r*W, (r+1)*W))
//   rng   [16|48 B] Philox key+counter | stream cursor | PCG64 state (numpy-exact mode)
//
// Bitboard formulation (ours; the reference works on byte matrices, envs/tetris.py:408-564):
//   B(piece,rot,x) = OR over the 4 cells (i,j) of  col[x+j-P] >> i      (walls read as all-ones)
//   collision at y  <=>  bit y of B          (Tetris.collision, envs/tetris.py:408-427)
//   hard drop from y = y + ctz(B >> (y+1))   (Tetris.drop_active_tetromino, envs/tetris.py:445-448)
//   full rows       = AND over columns       (Tetris.clear_filled_rows, envs/tetris.py:481-512)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tg {

constexpr int P = 4;  // padding (envs/tetris.py:130)

enum Op : int { OP_LEFT = 0, OP_RIGHT, OP_DOWN, OP_CW, OP_CCW, OP_SWAP, OP_HARD, OP_NOOP };

// Everything a kernel needs to know about the env configuration (passed by value).
struct DevCfg {
    int W, H, Wp, Hp, Q;
    int gravity, autoreset, rng_mode, terminate_on_illegal;
    int rand_kind;     // 0 = 7-bag, 1 = uniform (TrueRandomizer)
    int board_stride;  // bytes per env in state.board
    int ids_off;       // byte offset of the nibble id plane inside a board record
    int ids_words;     // u32 words of the id plane (ceil(H*W/8))
    int rng_stride;
    int OB;            // obs board bytes = Hp*Wp
    int OQ;            // obs queue bytes = 16*Q
    unsigned int inv_q; // 65536 / Q + 1: it / Q == (it * inv_q) >> 16 for it < 4096 (host-computed: the image warps divided once per tile)
    int A, F;          // placements 4W, features W+3
    int rgb_w;         // Wp + 4*max(Q, holder_size, 1)
    int NPC;           // pieces of the tetromino set (7 for the reference's; Tetris(tetrominoes=[...]): 1..7)
    unsigned nhalf3;   // n // 2 of piece p in bits 3p..3p+2 (GroupedActionsObservations: x = column + padding - n // 2)
    int holder_size;   // TetrominoHolder(size): 1..4
    int OH;            // obs holder bytes = 16 * holder_size
    int spawn_x[7];    // W_pad//2 - n//2  (Tetris.reset_tetromino_position, envs/tetris.py:536-541)
    unsigned char op_lut[8];    // action id -> Op following the elif order of envs/tetris.py:223-256
    unsigned char skipgrav[8];  // action == actions.hard_drop (envs/tetris.py:259)
    int act_hard, act_noop;
    double r_alife, r_go, r_invalid;
    long long seq_len;
    unsigned long long env_id_offset;
};

// Constant piece tables, generated on the host by literally rotating the reference's base
// matrices with rot90 (tg_api.cu: build_tables) and uploaded once per device.
//   c_cells[p][r]  : 4 cells, nibble k = (i << 2) | j   (row i, column j inside the n x n box)
//   c_rowbytes[p][r][i] : row i of the matrix zero-padded to 4x4, one id-valued byte per cell
//   c_n[p]         : matrix size n
//   c_colors[v]    : RGB of cell value v (envs/tetris.py:45-75)
__constant__ unsigned short c_cells[7][4];
__constant__ unsigned int c_rowbytes[7][4][4];
__constant__ int c_n[7];
__constant__ uint2 c_ptab[7][4];   // column profile per (piece, rotation), see place_fast
__constant__ unsigned char c_colors[16][4];

// Piece tables as seen by device functions: the hot kernels copy them to shared memory first
// (per-thread piece indices make constant-cache reads serialise), the others read the __constant__ copy.
struct Tabs {
    const uint2* ptab;            // [p * 4 + r]  column profiles (place_fast); nullptr = read c_ptab
    const unsigned short* cells;  // [p * 4 + r]
    const unsigned int* rowbytes; // [(p * 4 + r) * 4 + i]
    const int* n;                 // [p]
};
__device__ __forceinline__ Tabs const_tabs() {
    Tabs t;
    t.ptab = &c_ptab[0][0]; t.cells = &c_cells[0][0]; t.rowbytes = &c_rowbytes[0][0][0]; t.n = &c_n[0];
    return t;
}

// ---- hot record ------------------------------------------------------------------------------
struct Hot {
    int x, y, p, r, hold, hold_r, swapped, over, pending;
    uint32_t bag;  // 7 nibbles + index in bits 28..30
    uint32_t hq;   // holder_size > 1: FIFO of held pieces -- bits 0..2 count, slot k (0 = oldest) at bits 3 + 5k: piece | rotation << 3
    uint64_t queue;
    float ep_ret;
    uint32_t ep_len, ep_lines;
};
// XT ("extended", template flag of the step path): the env has a custom tetromino set (NPC != 7) or a holder FIFO (size > 1).
// The XT = false instantiations of the per-call step kernel keep the reference configuration's code exactly as small as it was
// before those options existed (bag of 7 unrolled, no FIFO word, 16-byte holder image): its speed depends on its code footprint.
template <bool XT = true>
__device__ __forceinline__ void hot_load(Hot& h, const uint32_t* w) {
    uint32_t a = w[0];
    h.x = a & 63; h.y = (a >> 6) & 127; h.p = (a >> 13) & 7; h.r = (a >> 16) & 3;
    h.hold = (a >> 18) & 15; h.hold_r = (a >> 22) & 3;
    h.swapped = (a >> 24) & 1; h.over = (a >> 25) & 1; h.pending = (a >> 26) & 1;
    h.bag = w[1];
    h.queue = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
    h.ep_ret = __uint_as_float(w[4]); h.ep_len = w[5]; h.ep_lines = w[6];
    h.hq = XT ? w[7] : 0u;
}
template <bool XT = true>
__device__ __forceinline__ void hot_store(const Hot& h, uint32_t* w) {
    w[0] = (uint32_t)h.x | ((uint32_t)h.y << 6) | ((uint32_t)h.p << 13) | ((uint32_t)h.r << 16) |
           ((uint32_t)h.hold << 18) | ((uint32_t)h.hold_r << 22) | ((uint32_t)h.swapped << 24) |
           ((uint32_t)h.over << 25) | ((uint32_t)h.pending << 26);
    w[1] = h.bag;
    w[2] = (uint32_t)h.queue; w[3] = (uint32_t)(h.queue >> 32);
    w[4] = __float_as_uint(h.ep_ret); w[5] = h.ep_len; w[6] = h.ep_lines; w[7] = XT ? h.hq : 0u;
}

// ---- column bitboards -----------------------------------------------------------------------
template <class COLT> __device__ __forceinline__ int ctz_t(COLT v);
template <> __device__ __forceinline__ int ctz_t<uint32_t>(uint32_t v) { return __ffs((int)v) - 1; }
template <> __device__ __forceinline__ int ctz_t<uint64_t>(uint64_t v) { return __ffsll((long long)v) - 1; }
template <class COLT> __device__ __forceinline__ int popc_t(COLT v);
template <> __device__ __forceinline__ int popc_t<uint32_t>(uint32_t v) { return __popc(v); }
template <> __device__ __forceinline__ int popc_t<uint64_t>(uint64_t v) { return __popcll(v); }

template <class COLT>
__device__ __forceinline__ COLT col_at(const COLT* cols, int cx, int W) {
    return (unsigned)cx < (unsigned)W ? cols[cx] : ~COLT(0);  // bedrock walls
}
// B mask of a piece orientation at padded x (see header comment)
template <class COLT>
__device__ __forceinline__ COLT bmask(const COLT* cols, int W, uint32_t cells, int x) {
    COLT B = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int c = (cells >> (4 * k)) & 15;
        B |= col_at(cols, x + (c & 3) - P, W) >> (c >> 2);
    }
    return B;
}
template <class COLT>
__device__ __forceinline__ COLT floor_bits(int H, int Hp) {
    return (Hp >= (int)(8 * sizeof(COLT)) ? ~COLT(0) : ((COLT(1) << Hp) - 1)) & ~((COLT(1) << H) - 1);
}

// ---- nibble-packed id plane ---------------------------------------------------------------
__device__ __forceinline__ uint32_t ids_get8(const uint32_t* ids, int nib_off) {
    int b = nib_off * 4;
    return __funnelshift_r(ids[b >> 5], ids[(b >> 5) + 1], b & 31);
}
__device__ __forceinline__ void ids_put(uint32_t* ids, int nib_off, int cnt, uint32_t v) {
    int b = nib_off * 4, wi = b >> 5, sh = b & 31;
    uint64_t m = (cnt >= 8 ? 0xFFFFFFFFull : ((1ull << (4 * cnt)) - 1)) << sh;
    uint64_t vv = (uint64_t)v << sh;
    uint32_t lm = (uint32_t)m, hm = (uint32_t)(m >> 32);
    ids[wi] = (ids[wi] & ~lm) | ((uint32_t)vv & lm);
    if (hm) ids[wi + 1] = (ids[wi + 1] & ~hm) | ((uint32_t)(vv >> 32) & hm);
}
__device__ __forceinline__ void ids_set1(uint32_t* ids, int nib_off, uint32_t v) {
    int wi = nib_off >> 3, sh = (nib_off & 7) * 4;
    ids[wi] = (ids[wi] & ~(0xFu << sh)) | (v << sh);
}
__device__ __forceinline__ uint32_t ids_get1(const uint32_t* ids, int nib_off) {
    return (ids[nib_off >> 3] >> ((nib_off & 7) * 4)) & 15u;
}

// ---- randomizers ------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: one call -> 4 x u32.
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0];
        uint32_t h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = l1; c[2] = n2; c[3] = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

struct Rng {
    // view over one env's rng record (staged in shared memory by the step kernel)
    uint32_t* rec;
    bool dirty;          // record modified -> must be written back
    const uint8_t* seq;  // this env's injected stream (TG_RNG_SEQUENCE)
    uint64_t gid;        // global env id
};

// PCG64 (numpy): state = state * MULT + inc; out = rotr64(hi ^ lo, hi >> 58)
__device__ __forceinline__ uint64_t pcg64_next64(uint32_t* rec) {
    unsigned __int128 st = ((unsigned __int128)(((uint64_t*)rec)[0]) << 64) | ((uint64_t*)rec)[1];
    unsigned __int128 inc = ((unsigned __int128)(((uint64_t*)rec)[2]) << 64) | ((uint64_t*)rec)[3];
    const unsigned __int128 MULT = ((unsigned __int128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
    st = st * MULT + inc;
    uint64_t hi = (uint64_t)(st >> 64), lo = (uint64_t)st;
    ((uint64_t*)rec)[0] = hi; ((uint64_t*)rec)[1] = lo;
    uint64_t x = hi ^ lo;
    unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((64 - rot) & 63));
}
__device__ __forceinline__ uint32_t pcg64_next32(uint32_t* rec) {
    if (rec[8]) { rec[8] = 0; return rec[9]; }
    uint64_t v = pcg64_next64(rec);
    rec[8] = 1; rec[9] = (uint32_t)(v >> 32);
    return (uint32_t)v;
}

// in-place Fisher-Yates of the 7-bag (BagRandomizer.shuffle_bag, components/tetromino_randomizer.py:82-85)
// (plain values in, value out: a reference to the caller's Rng / Hot would pin them in local memory around the hot loop)
// (npc = pieces in the bag: 7 for the reference's set)
__device__ __noinline__ uint32_t shuffle_bag_raw(int rng_mode, uint32_t* rec, uint64_t gid, uint32_t bag, int npc) {
    uint32_t j6[6] = {0, 0, 0, 0, 0, 0};
    const int top = npc - 1;
    if (rng_mode == 2) {
        // numpy Generator.shuffle: for i = n-1..1: j = random_interval(i) (masked rejection on next_uint32)
        for (int i = top; i >= 1; i--) {
            uint32_t mask = i | (i >> 1); mask |= mask >> 2;
            uint32_t v;
            do { v = pcg64_next32(rec) & mask; } while (v > (uint32_t)i);
            j6[top - i] = v;
        }
    } else {
        uint64_t seed = ((uint64_t*)rec)[0];
        uint32_t ctr = rec[2];
        rec[2] = ctr + 1;
        // ONE Philox call per shuffle: draws 0..3 take a word each (umulhi(u, i + 1)); draws 4 and 5 reuse the fractions the first
        // two left over (u * m mod 2^32 is uniform again up to 2^-29: the multiply-shift chain of a mixed-radix expansion)
        uint32_t c[4] = {ctr, (uint32_t)gid, (uint32_t)(gid >> 32), 0u};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t u[6] = {c[0], c[1], c[2], c[3], top >= 1 ? c[0] * (uint32_t)(top + 1) : 0u, top >= 2 ? c[1] * (uint32_t)top : 0u};
        for (int i = top; i >= 1; i--) j6[top - i] = __umulhi(u[top - i], (uint32_t)(i + 1));   // 7 pieces: umulhi(u0, 7), (u1, 6), ... (u5, 2)
    }
    for (int i = top; i >= 1; i--) {
        uint32_t j = j6[top - i];
        uint32_t vi = (bag >> (4 * i)) & 15u, vj = (bag >> (4 * j)) & 15u;
        bag = (bag & ~(15u << (4 * i))) | (vj << (4 * i));
        bag = (bag & ~(15u << (4 * j))) | (vi << (4 * j));
    }
    return bag & 0x0FFFFFFFu;  // index = 0
}
// the reference's seven pieces: same draws, loops unrolled over constants
__device__ __noinline__ uint32_t shuffle_bag7_raw(int rng_mode, uint32_t* rec, uint64_t gid, uint32_t bag) {
    uint32_t j6[6];
    if (rng_mode == 2) {
        for (int i = 6; i >= 1; i--) {
            uint32_t mask = i | (i >> 1); mask |= mask >> 2;
            uint32_t v;
            do { v = pcg64_next32(rec) & mask; } while (v > (uint32_t)i);
            j6[6 - i] = v;
        }
    } else {
        uint64_t seed = ((uint64_t*)rec)[0];
        uint32_t ctr = rec[2];
        rec[2] = ctr + 1;
        // ONE Philox call per shuffle (see shuffle_bag_raw): the last two draws reuse the fractions left over by the first two
        uint32_t c[4] = {ctr, (uint32_t)gid, (uint32_t)(gid >> 32), 0u};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        j6[0] = __umulhi(c[0], 7u); j6[1] = __umulhi(c[1], 6u); j6[2] = __umulhi(c[2], 5u);
        j6[3] = __umulhi(c[3], 4u); j6[4] = __umulhi(c[0] * 7u, 3u); j6[5] = __umulhi(c[1] * 6u, 2u);
    }
#pragma unroll
    for (int i = 6; i >= 1; i--) {
        uint32_t j = j6[6 - i];
        uint32_t vi = (bag >> (4 * i)) & 15u, vj = (bag >> (4 * j)) & 15u;
        bag = (bag & ~(15u << (4 * i))) | (vj << (4 * i));
        bag = (bag & ~(15u << (4 * j))) | (vi << (4 * j));
    }
    return bag & 0x0FFFFFFFu;  // index = 0
}
template <bool XT = true>
__device__ __forceinline__ uint32_t shuffle_bag(const DevCfg& cfg, Rng& g, uint32_t bag) {
    g.dirty = true;
    if (!XT) return shuffle_bag7_raw(cfg.rng_mode, g.rec, g.gid, bag);
    return shuffle_bag_raw(cfg.rng_mode, g.rec, g.gid, bag, cfg.NPC);
}

// Draws that are not the 7-bag: the injected stream (TG_RNG_SEQUENCE) and TrueRandomizer.  Out of line: the call sites
// (commit, swap, reset) stay small -- inlined four times the PCG64 / Philox code made up a quarter of the step kernel and
// pushed it past the instruction cache.  Returns the piece; the caller marks the rng record dirty.
__device__ __noinline__ int draw_other(int rng_mode, long long seq_len, uint32_t* rec, const uint8_t* seq, uint64_t gid, int npc) {
    if (rng_mode == 1) {
        uint64_t cur = ((uint64_t*)rec)[0];
        ((uint64_t*)rec)[0] = cur + 1;
        return seq[cur % (uint64_t)seq_len];
    }
    // TrueRandomizer.get_next_tetromino (components/tetromino_randomizer.py:119-121): rng.integers(0, size)
    if (npc <= 1) return 0;                     // numpy returns `low` for an empty range without drawing
    const uint32_t un = (uint32_t)npc;
    if (rng_mode == 2) {
        // numpy: Lemire multiply-shift with rejection on next_uint32 (buffered_bounded_lemire_uint32, rng = size - 1)
        uint64_t m = (uint64_t)pcg64_next32(rec) * un;
        if ((uint32_t)m < un) {
            const uint32_t threshold = (0xFFFFFFFFu - (un - 1u)) % un;
            while ((uint32_t)m < threshold) m = (uint64_t)pcg64_next32(rec) * un;
        }
        return (int)(m >> 32);
    }
    uint64_t seed = ((uint64_t*)rec)[0];
    uint32_t ctr = rec[2];
    rec[2] = ctr + 1;
    uint32_t c[4] = {ctr, (uint32_t)gid, (uint32_t)(gid >> 32), 2u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (int)__umulhi(c[0], un);
}

// Randomizer.get_next_tetromino
template <bool XT = true>
__device__ __forceinline__ int draw_piece(const DevCfg& cfg, Rng& g, Hot& h) {
    if (cfg.rng_mode == 1 || cfg.rand_kind == 1) {
        g.dirty = true;
        return draw_other(cfg.rng_mode, cfg.seq_len, g.rec, g.seq, g.gid, XT ? cfg.NPC : 7);
    }
    // BagRandomizer.get_next_tetromino (components/tetromino_randomizer.py:67-80)
    int idx = (h.bag >> 28) & 7;
    int v = (h.bag >> (4 * idx)) & 15;
    idx++;
    if (idx >= (XT ? cfg.NPC : 7)) h.bag = shuffle_bag<XT>(cfg, g, h.bag);
    else h.bag = (h.bag & 0x0FFFFFFFu) | ((uint32_t)idx << 28);
    return v;
}
// TetrominoQueue.get_next_tetromino (components/tetromino_queue.py:35-42)
template <bool XT = true>
__device__ __forceinline__ int queue_pop(const DevCfg& cfg, Rng& g, Hot& h) {
    int v = (int)(h.queue & 15u);
    uint64_t nv = (uint64_t)draw_piece<XT>(cfg, g, h);
    h.queue = (h.queue >> 4) | (nv << (4 * (cfg.Q - 1)));
    return v;
}

// ---- env logic (one thread = one env; board record in shared memory) --------------------------
// Tetris.reset (envs/tetris.py:274-307) + TetrominoQueue.reset + BagRandomizer.reset
template <class COLT, bool XT = true>
__device__ __forceinline__ void env_reset(const DevCfg& cfg, Hot& h, uint32_t* rec, Rng& g) {
    COLT* cols = (COLT*)rec;
    uint32_t* ids = rec + cfg.ids_off / 4;
    COLT fl = floor_bits<COLT>(cfg.H, cfg.Hp);
    for (int c = 0; c < cfg.W; c++) cols[c] = fl;
    for (int i = 0; i < cfg.ids_words; i++) ids[i] = 0;
    h.over = 0; h.pending = 0;
    if (cfg.rng_mode != 1 && cfg.rand_kind == 0) h.bag = shuffle_bag<XT>(cfg, g, 0x06543210u);   // TrueRandomizer.reset only reseeds
    h.queue = 0;
    for (int i = 0; i < cfg.Q; i++) h.queue |= (uint64_t)draw_piece<XT>(cfg, g, h) << (4 * i);
    h.p = queue_pop<XT>(cfg, g, h);
    h.r = 0; h.x = cfg.spawn_x[h.p]; h.y = 0;
    h.hold = 0; h.hold_r = 0; h.swapped = 0; h.hq = 0;
    h.ep_ret = 0.f; h.ep_len = 0; h.ep_lines = 0;
}

// row compaction after a line clear (Tetris.clear_filled_rows, envs/tetris.py:481-512)
template <class COLT>
__device__ __noinline__ void clear_rows(const DevCfg& cfg, COLT* cols, uint32_t* ids, COLT full) {
    const int W = cfg.W, H = cfg.H;
    COLT any = 0;
    for (int c = 0; c < W; c++) any |= cols[c];
    int top = ctz_t<COLT>(any);  // rows above `top` are empty (floor bits guarantee any != 0)
    // id plane: surviving rows slide down, keeping order
    int dst = H - 1;
    for (int src = H - 1; src >= top; --src) {
        if ((full >> src) & 1) continue;
        if (dst != src)
            for (int k = 0; k < W; k += 8) {
                int cnt = min(8, W - k);
                ids_put(ids, dst * W + k, cnt, ids_get8(ids, src * W + k));
            }
        dst--;
    }
    for (; dst >= top; --dst)
        for (int k = 0; k < W; k += 8) ids_put(ids, dst * W + k, min(8, W - k), 0u);
    // column bitboards: delete the full rows' bits, upper bits move down (towards higher row index)
    for (int c = 0; c < W; c++) {
        COLT v = cols[c], f = full;
        while (f) {
            int r = ctz_t<COLT>(f);
            f &= f - 1;
            COLT below = (COLT(1) << r) - 1;  // rows above r (smaller index)
            v = (v & ~((below << 1) | 1)) | ((v & below) << 1);
        }
        cols[c] = v;
    }
}

// TetrominoHolder.swap for size > 1 (components/tetromino_holder.py:31-49): a FIFO -- while it is not full the piece is stored
// and nothing comes back; once full the oldest piece comes back.  Values in, values out (out of line: rare, and the step
// kernel's speed depends on its code footprint).  Returns new hq | (1 + piece + 8 * rotation of the piece handed back, 0 = none) << 32.
__device__ __noinline__ uint64_t holder_fifo_swap(uint32_t hq, int size, int p, int r) {
    int cnt = (int)(hq & 7u);
    uint32_t slots = hq >> 3, back = 0;
    if (cnt >= size) {
        back = 1u + (slots & 31u);          // oldest: piece | rotation << 3
        slots >>= 5;
        cnt--;
    }
    slots |= (uint32_t)(p | (r << 3)) << (5 * cnt);
    cnt++;
    return (uint64_t)((slots << 3) | (uint32_t)cnt) | ((uint64_t)back << 32);
}

// Row i of the holder image's slot s (Tetris._get_obs, envs/tetris.py:578-592): the held piece's matrix row as four id bytes,
// ones for an empty slot.  rowbytes = [(piece * 4 + rotation) * 4 + row].
__device__ __forceinline__ uint32_t holder_row(const DevCfg& cfg, const Hot& h, const uint32_t* rowbytes, int s, int i) {
    if (cfg.holder_size <= 1) return h.hold ? rowbytes[((h.hold - 1) * 4 + h.hold_r) * 4 + i] : 0x01010101u;
    const uint32_t sl = (h.hq >> (3 + 5 * s)) & 31u;
    return s < (int)(h.hq & 7u) ? rowbytes[((int)(sl & 7u) * 4 + (int)(sl >> 3)) * 4 + i] : 0x01010101u;
}

struct StepResult {
    double reward;
    int lines;
    int terminated;
    int dirty;  // board record changed
    int did_reset;
    unsigned long long Bfin;   // env_step: collision mask (bmask) of the piece that is active AFTER the step, at its x / rotation
};

// Tetris.commit_active_tetromino (envs/tetris.py:450-479); B = bmask of the active piece at h.x
template <class COLT, bool XT = true>
__device__ __forceinline__ void env_commit(const DevCfg& cfg, const Tabs& tb, Hot& h, uint32_t* rec, Rng& g, COLT B,
                                           StepResult& res) {
    COLT* cols = (COLT*)rec;
    uint32_t* ids = rec + cfg.ids_off / 4;
    if ((B >> h.y) & 1) {
        res.reward = cfg.r_go;
        h.over = 1;
        res.Bfin = (unsigned long long)B;   // the piece stays where it is
        return;
    }
    h.y += ctz_t<COLT>(B >> (h.y + 1));  // drop_active_tetromino
    uint32_t cells = tb.cells[h.p * 4 + h.r];
#pragma unroll
    for (int k = 0; k < 4; k++) {  // place_active_tetromino / project_tetromino
        int c = (cells >> (4 * k)) & 15;
        int row = h.y + (c >> 2), col = h.x + (c & 3) - P;
        cols[col] |= COLT(1) << row;
        ids_set1(ids, row * cfg.W + col, (uint32_t)(h.p + 2));
    }
    // every filled playfield row is cleared, also ones the piece did not touch (a poked board may hold them;
    // Tetris.clear_filled_rows scans the whole board, envs/tetris.py:481-512)
    COLT full = (COLT(1) << cfg.H) - 1;
    for (int c = 0; c < cfg.W; c++) full &= cols[c];
    int lines = popc_t<COLT>(full);
    if (lines) clear_rows<COLT>(cfg, cols, ids, full);
    res.lines = lines;
    res.reward = (double)(lines * lines * cfg.W);  // Tetris.score (envs/tetris.py:621-630)
    // spawn_tetromino (envs/tetris.py:393-401)
    h.p = queue_pop<XT>(cfg, g, h);
    h.r = 0; h.x = cfg.spawn_x[h.p]; h.y = 0;
    COLT Bn = bmask<COLT>(cols, cfg.W, tb.cells[h.p * 4], h.x);
    res.Bfin = (unsigned long long)Bn;
    h.over = (int)(Bn & 1);
    res.reward += cfg.r_alife;
    if (h.over) res.reward = cfg.r_go;
    h.swapped = 0;
    res.dirty = 1;
}

// Tetris.step (envs/tetris.py:203-272) with the vector-env autoreset policy around it.
// `force_x/force_r` >= 0: grouped placement (GroupedActionsObservations.step sets env.x and the
// rotated piece, y untouched, then base hard_drop; wrappers/grouped.py:241-259).
template <class COLT, bool XT = true>
__device__ __forceinline__ void env_step(const DevCfg& cfg, const Tabs& tb, Hot& h, uint32_t* rec, Rng& g, int action,
                                         StepResult& res) {
    COLT* cols = (COLT*)rec;
    res.reward = 0.0; res.lines = 0; res.dirty = 0; res.did_reset = 0;
    int op = OP_NOOP, skipgrav = 0;
    if ((unsigned)action < 8u) { op = cfg.op_lut[action]; skipgrav = cfg.skipgrav[action]; }
    if (op == OP_SWAP && !h.swapped) {  // envs/tetris.py:242-252 + TetrominoHolder.swap
        int np, nr;
        if (XT && cfg.holder_size > 1) {
            const uint64_t res = holder_fifo_swap(h.hq, cfg.holder_size, h.p, h.r);
            h.hq = (uint32_t)res;
            const uint32_t back = (uint32_t)(res >> 32);
            if (back == 0) { np = queue_pop<XT>(cfg, g, h); nr = 0; }
            else { np = (int)((back - 1) & 7u); nr = (int)((back - 1) >> 3); }
        } else {
            if (h.hold == 0) { np = queue_pop<XT>(cfg, g, h); nr = 0; }
            else { np = h.hold - 1; nr = h.hold_r; }
            h.hold = h.p + 1; h.hold_r = h.r;
        }
        h.p = np; h.r = nr; h.swapped = 1;
        h.x = cfg.spawn_x[np]; h.y = 0;
    }
    int dx = (op == OP_LEFT) ? -1 : (op == OP_RIGHT ? 1 : 0);
    int dr = (op == OP_CW) ? 1 : (op == OP_CCW ? 3 : 0);
    int dy = (op == OP_DOWN) ? 1 : 0;
    int cx = h.x + dx, cr = (h.r + dr) & 3, cy = h.y + dy;
    COLT B = bmask<COLT>(cols, cfg.W, tb.cells[h.p * 4 + cr], cx);
    if (!((B >> cy) & 1)) { h.x = cx; h.r = cr; h.y = cy; }
    else if (dx | dr) B = bmask<COLT>(cols, cfg.W, tb.cells[h.p * 4 + h.r], h.x);
    bool do_commit = (op == OP_HARD);
    if (cfg.gravity && !skipgrav) {  // envs/tetris.py:259-264
        if (!((B >> (h.y + 1)) & 1)) h.y += 1;
        else do_commit = true;
    }
    res.Bfin = (unsigned long long)B;       // B belongs to (h.p, h.r, h.x) on every path above; a commit replaces it
    if (do_commit) env_commit<COLT, XT>(cfg, tb, h, rec, g, B, res);
    res.terminated = h.over;
}

// ---- features from column bitboards (FeatureVectorObservation, wrappers/observation.py:177-278) --
// `cols` already has the rows the reference zeroes (row 0, or rows 0-1) cleared by the caller.
template <class COLT>
__device__ __forceinline__ void col_features(COLT col, int H, int& height, int& holes) {
    COLT v = col & ((COLT(1) << H) - 1);
    if (v == 0) { height = 0; holes = 0; return; }
    int first = ctz_t<COLT>(v);
    height = H - first;                               // calc_height
    holes = height - popc_t<COLT>(v);                 // empty cells below the first filled one
}

// Features of (cols | optional piece cells), after deleting `full` rows and zeroing `rowzero` rows.
// Q1: the reference zeroes padded rows 0 (mask all zero) or 0-1 (mask has a 1) through
// integer fancy indexing, it does NOT mask the active piece (wrappers/observation.py:252).
struct FeatSum { int sum_h, max_h, holes, bump, lines; };
template <class COLT>
__device__ __forceinline__ FeatSum placement_eval(const DevCfg& cfg, const COLT* cols, uint32_t cells, int x, int y,
                                                  bool place, bool do_clear, COLT rowzero, uint8_t* out) {
    const int W = cfg.W, H = cfg.H;
    int crow[4], ccol[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int c = (cells >> (4 * k)) & 15;
        crow[k] = y + (c >> 2);
        ccol[k] = place ? x + (c & 3) - P : -1;
    }
    COLT full = 0;
    if (place && do_clear) {
        full = (COLT(1) << H) - 1;   // all playfield rows are candidates (clear_filled_rows scans the whole board)
        for (int c = 0; c < W; c++) {
            COLT v = cols[c];
#pragma unroll
            for (int k = 0; k < 4; k++) if (ccol[k] == c) v |= COLT(1) << crow[k];
            full &= v;
        }
        full &= (COLT(1) << H) - 1;
    }
    FeatSum fs;
    fs.lines = popc_t<COLT>(full);
    int prev = 0, maxh = 0, holes = 0, bump = 0, sumh = 0;
    for (int c = 0; c < W; c++) {
        COLT v = cols[c];
#pragma unroll
        for (int k = 0; k < 4; k++) if (ccol[k] == c) v |= COLT(1) << crow[k];
        COLT f = full;
        while (f) {  // Tetris.clear_filled_rows on the projected copy (wrappers/grouped.py:171-177)
            int r = ctz_t<COLT>(f);
            f &= f - 1;
            COLT below = (COLT(1) << r) - 1;
            v = (v & ~((below << 1) | 1)) | ((v & below) << 1);
        }
        v &= ~rowzero;
        int hgt, hol;
        col_features<COLT>(v, H, hgt, hol);
        if (out) out[c] = (uint8_t)hgt;
        holes += hol;
        sumh += hgt;
        maxh = max(maxh, hgt);
        if (c > 0) bump += abs(hgt - prev);
        prev = hgt;
    }
    if (out) {
        out[W] = (uint8_t)maxh;
        out[W + 1] = (uint8_t)holes;  // uint8 wrap (wrappers/observation.py:277)
        out[W + 2] = (uint8_t)bump;
    }
    fs.sum_h = sumh; fs.max_h = maxh; fs.holes = holes; fs.bump = bump;
    return fs;
}
template <class COLT>
__device__ __forceinline__ void placement_features(const DevCfg& cfg, const COLT* cols, uint32_t cells, int x, int y,
                                                   bool place, bool do_clear, COLT rowzero, uint8_t* out, int& lines_out) {
    lines_out = placement_eval<COLT>(cfg, cols, cells, x, y, place, do_clear, rowzero, out).lines;
}

// ---- incremental placement evaluation -------------------------------------------------------------------
// Per env (once per enumeration): column heights / holes of the current board with the wrapper's row zeroing,
// their sums, and prefix / suffix ANDs of the occupancy columns.  A placement touches <= 4 adjacent columns, so
//   full rows = touched & pre[c0] & suf[c1] & AND_{c0..c1}(col | piece bits)
// and when no row is cleared (the common case) only the touched columns' features change; the full
// recomputation (placement_eval) is needed only when rows are cleared.
template <class COLT>
struct EnvBase {
    uint8_t* h;       // [W]   heights
    uint8_t* ho;      // [W]   holes per column
    COLT* pre;        // [W]   AND of columns < c
    COLT* suf;        // [W]   AND of columns > c
    uint16_t* bs;     // [W]   bumpiness prefix: bs[c] = sum over k < c of |h[k+1] - h[k]|
    int sum_h, holes, bump, max_h;
};
template <class COLT>
__device__ __forceinline__ void env_base_compute(const DevCfg& cfg, const COLT* cols, COLT rowzero, EnvBase<COLT>& b) {
    const int W = cfg.W, H = cfg.H;
    COLT acc = ~COLT(0);
    int prev = 0;
    b.sum_h = 0; b.holes = 0; b.bump = 0; b.max_h = 0;
    for (int c = 0; c < W; c++) {
        b.pre[c] = acc;
        acc &= cols[c];
        int hgt, hol;
        col_features<COLT>(cols[c] & ~rowzero, H, hgt, hol);
        b.h[c] = (uint8_t)hgt; b.ho[c] = (uint8_t)hol;
        b.sum_h += hgt; b.holes += hol; b.max_h = max(b.max_h, hgt);
        if (c > 0) b.bump += abs(hgt - prev);
        b.bs[c] = (uint16_t)b.bump;
        prev = hgt;
    }
    acc = ~COLT(0);
    for (int c = W - 1; c >= 0; c--) { b.suf[c] = acc; acc &= cols[c]; }
}
// One placement of GroupedActionsObservations.observation (wrappers/grouped.py:148-181).
struct Placement { int x, y, rot, kind; };  // kind: 0 regular, 1 illegal (frame), 2 game over
template <class COLT>
__device__ __forceinline__ Placement eval_placement(const DevCfg& cfg, const Tabs& tb, const COLT* cols, int piece, int rot0, int a, COLT& Bout) {
    Placement pl;
    int xb = a >> 2, rr = a & 3;
    pl.rot = (rot0 + rr) & 3;              // cumulative rot90 presses (wrappers/grouped.py:153-154)
    pl.x = xb + P - tb.n[piece] / 2;        // wrappers/grouped.py:157-158
    uint32_t cells = tb.cells[piece * 4 + pl.rot];
    COLT B = bmask<COLT>(cols, cfg.W, cells, pl.x);
    pl.y = ctz_t<COLT>(B >> 1);            // while !collision(y+1): y++  from y = 0, no test at y = 0 (Q3)
    bool frame = false;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int c = (cells >> (4 * k)) & 15;
        frame |= (unsigned)(pl.x + (c & 3) - P) >= (unsigned)cfg.W;
    }
    pl.kind = frame ? 1 : (((B >> pl.y) & 1) ? 2 : 0);
    Bout = B;
    return pl;
}

// ---- placement evaluation from column profiles -----------------------------------------------------------------------
// c_ptab[p][r] describes the 4x4 matrix of a piece orientation by COLUMN j = 0..3:
//   .x bits  4j..4j+3  row mask of column j          .x bits 16+2j..17+2j  row offset of its top cell
//   .x bits 24-25 / 26-27  first / last column holding cells          .y bits 3j..3j+2  cells in column j
// With the per-env base (heights h[], holes ho[], prefix / suffix column ANDs) a regular placement that clears no row
// needs no bit scan at all: in a touched column the new height is max(h, H - top) (the piece came down from above, but a
// poked board may hold cells above it) and popc grows by the column's cell count, so holes' = h' - (h - ho) - cells.
// `colp` is the column array padded with P all-ones wall columns on both sides (colp[c + P] = column c).
// Row clears stay on this path too: with F = the cleared rows, a column keeps its cells u = v & ~F in order, packed
// towards the floor, so its top cell (row t = ctz(u)) ends at row t + popc(F >> (t + 1)) and popc(u) cells remain; the
// result's row 0 is empty after a clear, so the wrapper's row zeroing (Q1) has no effect there.
// Returns 0 = regular (fs / out filled), 1 = frame (illegal), 2 = game over, 3 = regular, no row cleared, but a cell of the
// piece lands in playfield row 0, the row the feature wrapper zeroes: the caller runs the exact placement_eval (rare);
// 4 = rows get cleared and `defer_clear` was set.
template <class COLT>
__device__ __forceinline__ int place_fast(const DevCfg& cfg, const EnvBase<COLT>& b, const COLT* colp, uint32_t cells, uint2 pt, int x,
                                          FeatSum& fs, int& y_out, uint8_t* out, bool defer_clear = false) {
    const int W = cfg.W, H = cfg.H;
    COLT B = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int c = (cells >> (4 * k)) & 15;
        B |= colp[x + (c & 3)] >> (c >> 2);
    }
    const int y = ctz_t<COLT>(B >> 1);   // while !collision(y+1): y++ from y = 0, no test at y = 0
    y_out = y;
    const int jmin = (pt.x >> 24) & 3, jmax = (pt.x >> 26) & 3;
    const int c0 = x + jmin - P, c1 = x + jmax - P;
    if (c0 < 0 || c1 >= W) return 1;
    if ((B >> y) & 1) return 2;
    COLT full = b.pre[c0] & b.suf[c1] & ((COLT(1) << H) - 1);
    int nh[4], sumh = b.sum_h, holes = b.holes, maxh = b.max_h, bump = b.bump;
    bool top0 = false;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const int j = jmin + t, c = c0 + t;
        nh[t] = 0;
        if (j <= jmax) {
            full &= colp[c + P] | ((COLT)((pt.x >> (4 * j)) & 15u) << y);
            const int top = y + (int)((pt.x >> (16 + 2 * j)) & 3u);
            top0 |= top == 0;
            const int oh = (int)b.h[c], oho = (int)b.ho[c];
            const int hn = max(oh, H - top);
            nh[t] = hn;
            sumh += hn - oh;
            holes += hn - (oh - oho) - (int)((pt.y >> (3 * j)) & 7u) - oho;
            maxh = max(maxh, hn);
        }
    }
    if (full != 0) {
        if (defer_clear) return 4;   // the caller batches row-clearing placements (warp divergence) and calls again without the flag
        // Tetris.clear_filled_rows on the projected copy (wrappers/grouped.py:171-177), column by column
        const COLT keep = ~full & ((COLT(1) << H) - 1);
        int s_sum = 0, s_max = 0, s_hol = 0, s_bmp = 0, prev = 0;
        for (int c = 0; c < W; c++) {
            COLT v = colp[c + P];
            const int t = c - c0;
            if ((unsigned)t <= (unsigned)(c1 - c0)) v |= (COLT)((pt.x >> (4 * (jmin + t))) & 15u) << y;
            const COLT u = v & keep;
            int hgt = 0, hol = 0;
            if (u != 0) {
                const int tp = ctz_t<COLT>(u);
                hgt = H - tp - popc_t<COLT>((full >> tp) >> 1);
                hol = hgt - popc_t<COLT>(u);
            }
            if (out) out[c] = (uint8_t)hgt;
            s_sum += hgt; s_hol += hol; s_max = max(s_max, hgt);
            if (c > 0) s_bmp += abs(hgt - prev);
            prev = hgt;
        }
        if (out) { out[W] = (uint8_t)s_max; out[W + 1] = (uint8_t)s_hol; out[W + 2] = (uint8_t)s_bmp; }
        fs.lines = popc_t<COLT>(full); fs.sum_h = s_sum; fs.max_h = s_max; fs.holes = s_hol; fs.bump = s_bmp;
        return 0;
    }
    if (top0) return 3;
    // bumpiness: only the pairs (c, c+1), c = c0-1 .. c1, change; their old sum comes from the prefix array
    {
        const int wdt = c1 - c0;       // touched columns - 1
        const int lo = max(c0 - 1, 0), hi1 = min(c1 + 1, W - 1);
        int nb = 0;
        if (c0 > 0) nb += abs(nh[0] - (int)b.h[c0 - 1]);
        if (wdt >= 1) nb += abs(nh[1] - nh[0]);
        if (wdt >= 2) nb += abs(nh[2] - nh[1]);
        if (wdt >= 3) nb += abs(nh[3] - nh[2]);
        const int last = wdt == 0 ? nh[0] : (wdt == 1 ? nh[1] : (wdt == 2 ? nh[2] : nh[3]));
        if (c1 < W - 1) nb += abs((int)b.h[c1 + 1] - last);
        bump += nb - ((int)b.bs[hi1] - (int)b.bs[lo]);
    }
    if (out) {
        for (int c = 0; c < W; c++) out[c] = b.h[c];
#pragma unroll
        for (int t = 0; t < 4; t++) if (c0 + t <= c1) out[c0 + t] = (uint8_t)nh[t];
        out[W] = (uint8_t)maxh; out[W + 1] = (uint8_t)holes; out[W + 2] = (uint8_t)bump;
    }
    fs.lines = 0; fs.sum_h = sumh; fs.max_h = maxh; fs.holes = holes; fs.bump = bump;
    return 0;
}

}  // namespace tg
