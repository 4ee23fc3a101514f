// tg_fn.cuh -- the functional env facade (tetris_gymnasium/envs/tetris_fn.py + functional/core.py, queue.py).
//
// A different game from the NumPy env: 7 actions (L0 R1 D2 CCW3 CW4 NOOP5 HARD6), no holder,
// the queue IS the bag (a permutation of arange(queue_size)), every piece is a 4x4 matrix spawned at
// x = W_pad//2 - 2, reward = score delta (soft drop +1, hard drop 2/cell, lines 100/300/500/800),
// lock when gravity could not move the piece or on hard drop, game over only on a blocked spawn, and a
// finished game is frozen.  The state is explicit and functional: (board i8[n,Hp,Wp], scalars i32[n,FN_S+Q])
// in, new arrays out (in == out aliases are allowed).  One thread per env on byte boards staged in shared memory.
#pragma once
#include "tg_device.cuh"
#include "tg_step.cuh"   // mbarrier / bulk-copy wrappers

namespace tg {

// scalar columns of the functional State
enum { FN_ACTIVE = 0, FN_ROT, FN_X, FN_Y, FN_QIDX, FN_OVER, FN_SCORE /* float bits */, FN_KEY0, FN_KEY1, FN_S };

struct FnParams {
    int W, H, Wp, Hp, Q, gravity;
    uint32_t invW;                // 65536 / W + 1: i / W == (i * invW) >> 16 for i < 4096 (host-computed: a division per thread otherwise)
    int64_t n;
    const int8_t* board_in; int8_t* board_out;
    const int32_t* sc_in; int32_t* sc_out;
    const int32_t* actions;       // NULL = reset
    const uint8_t* seq; int64_t seq_len;   // injected bags: bag k of env e = seq[e][k*Q .. k*Q+Q); NULL = Philox bags (seq_len >= 0) or the uniform queue (seq_len < 0)
    int8_t* obs; float* reward; uint8_t* terminated; int32_t* lines;
};

// core.collision (functional/core.py:86-100) on the 4x4 zero-padded matrix of (piece, rot)
__device__ __forceinline__ bool fn_collision(const FnParams& p, const int8_t* b, uint32_t cells, int x, int y) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int c = (cells >> (4 * k)) & 15;
        if (b[(y + (c >> 2)) * p.Wp + x + (c & 3)] > 0) return true;
    }
    return false;
}

// queue.create_bag_queue (functional/queue.py:20-35): a permutation of arange(Q).  The reference draws it with
// jax.random.permutation (threefry); ours comes from Philox(key) or from the injected stream -- the facade's
// piece sequences are not JAX-bit-compatible (parity unpinned for key-derived sequences).
__device__ __noinline__ void fn_new_bag(const FnParams& p, int64_t e, int32_t* sc) {
    int32_t* q = sc + FN_S;
    uint32_t bagno = (uint32_t)sc[FN_KEY1];
    if (p.seq) {
        for (int i = 0; i < p.Q; i++) q[i] = p.seq[e * p.seq_len + ((int64_t)bagno * p.Q + i) % p.seq_len];
    } else if (p.seq_len < 0) {
        // queue.create_uniform_queue (functional/queue.py:71-87): randint(key, (Q,), 0, Q - 1) -- maxval is exclusive, so the
        // reference never draws piece Q - 1; kept.  Values from Philox(key), not threefry (sequences are not JAX-compatible).
        for (int i = 0; i < p.Q; i++) {
            uint32_t c[4] = {bagno, (uint32_t)i, (uint32_t)e, (uint32_t)(e >> 32)};
            philox4x32_10(c, (uint32_t)sc[FN_KEY0], 0x0a11f02du);
            q[i] = p.Q > 1 ? (int)__umulhi(c[0], (uint32_t)(p.Q - 1)) : 0;
        }
    } else {
        for (int i = 0; i < p.Q; i++) q[i] = i;
        for (int i = p.Q - 1; i >= 1; i--) {
            uint32_t c[4] = {bagno, (uint32_t)i, (uint32_t)e, (uint32_t)(e >> 32)};
            philox4x32_10(c, (uint32_t)sc[FN_KEY0], 0x7e7215u);
            int j = (int)__umulhi(c[0], (uint32_t)(i + 1));
            int t = q[i]; q[i] = q[j]; q[j] = t;
        }
    }
    sc[FN_KEY1] = (int32_t)(bagno + 1);
}

// One CTA = T envs.  The boards of the tile are staged in shared memory (coalesced word copies in and out, one padded slot
// per env with an odd word stride), the game logic runs thread-per-env on the staged bytes, the observation is built per env
// into a shared tile and leaves coalesced.  (The first version worked on the byte boards in global memory and spent most of
// its time in per-byte index divisions of the observation loop: 1.2 ms per 1 M envs.)
__global__ void k_fn_step(const __grid_constant__ FnParams p, int bstr /* bytes per staged board, multiple of 4, odd word count */) {
    extern __shared__ __align__(16) uint8_t fsm[];
    const int tid = threadIdx.x, T = blockDim.x;
    const int64_t base = (int64_t)blockIdx.x * T;
    const int nv = (int)min((int64_t)T, p.n - base);
    const int OB = p.Hp * p.Wp, NS = FN_S + p.Q, HW = p.H * p.W;
    int8_t* s_board = (int8_t*)fsm;                                   // [T][bstr]
    int8_t* s_obs = (int8_t*)fsm + (size_t)T * bstr;                  // [T][HW], tile layout = output layout
    // 1. stage the boards (new_board = board: the update is functional, in == out aliases are allowed): asynchronous
    //    4-byte copies (LDGSTS), all of a thread's copies in flight at once
    {
        const int8_t* src = p.board_in + base * OB;
        if ((OB & 3) == 0 && ((uintptr_t)src & 3) == 0) {
            const int obw = OB >> 2;
            int el = 0, k = tid;
            while (k >= obw) { k -= obw; el++; }
            while (el < nv) {
                const uint32_t d = (uint32_t)__cvta_generic_to_shared(s_board + (size_t)el * bstr + 4 * k);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src + (size_t)el * OB + 4 * k) : "memory");
                k += T;
                while (k >= obw) { k -= obw; el++; }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else {
            for (int el = 0; el < nv; el++)
                for (int k = tid; k < OB; k += T) s_board[(size_t)el * bstr + k] = src[(size_t)el * OB + k];
        }
    }
    __syncthreads();
    if (tid < nv) {
        const int64_t e = base + tid;
        int8_t* b = s_board + (size_t)tid * bstr;
        int32_t sc[FN_S + 16];
        for (int i = 0; i < NS; i++) sc[i] = p.sc_in[e * NS + i];
        float old_score = __int_as_float(sc[FN_SCORE]);
        int lines = 0;
        const int spawn_x = p.Wp / 2 - 2;   // core.get_initial_x_y: 4x4 matrices (functional/core.py:66-83)
        if (!p.actions) {
            // tetris_fn.reset (envs/tetris_fn.py:318-367)
            for (int r = 0; r < p.Hp; r++)
                for (int c = 0; c < p.Wp; c++) b[r * p.Wp + c] = (r < p.H && c >= P && c < P + p.W) ? 0 : 1;
            sc[FN_KEY1] = 0;
            fn_new_bag(p, e, sc);
            sc[FN_ACTIVE] = sc[FN_S]; sc[FN_QIDX] = 1;
            sc[FN_ROT] = 0; sc[FN_X] = spawn_x; sc[FN_Y] = 0; sc[FN_OVER] = 0; sc[FN_SCORE] = __float_as_int(0.f);
            old_score = 0.f;
        } else if (!sc[FN_OVER]) {
            // tetris_fn.update_state (envs/tetris_fn.py:161-273)
            const int a = p.actions[e];
            int piece = sc[FN_ACTIVE], rot = sc[FN_ROT], x = sc[FN_X], y = sc[FN_Y];
            uint32_t cells = c_cells[piece][rot];
            int drop_reward = 0;
            if (a == 0) { if (!fn_collision(p, b, cells, x - 1, y)) x -= 1; }
            else if (a == 1) { if (!fn_collision(p, b, cells, x + 1, y)) x += 1; }
            else if (a == 2) { if (!fn_collision(p, b, cells, x, y + 1)) { y += 1; drop_reward = 1; } }
            else if (a == 3 || a == 4) {
                int nr = (rot + (a == 4 ? 1 : 3)) & 3;   // 3 = counter-clockwise, 4 = clockwise (envs/tetris_fn.py:470-478)
                if (!fn_collision(p, b, c_cells[piece][nr], x, y)) { rot = nr; cells = c_cells[piece][nr]; }
            } else if (a == 6) {                          // core.hard_drop (functional/core.py:230-251)
                int ny = y;
                while (!fn_collision(p, b, cells, x, ny + 1)) ny++;
                drop_reward = 2 * (ny - y);
                y = ny;
            }
            int yg = y;
            if (p.gravity && !fn_collision(p, b, cells, x, y + 1)) yg = y + 1;   // core.graviy_step
            bool should_lock = (yg == y) && p.gravity;
            y = yg;
            int lock_reward = 0;
            if (should_lock || a == 6) {
                // place_active_tetromino (envs/tetris_fn.py:370-413) / core.lock_active_tetromino
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    int c = (cells >> (4 * k)) & 15;
                    b[(y + (c >> 2)) * p.Wp + x + (c & 3)] += (int8_t)(piece + 2);
                }
                // core.clear_filled_rows (functional/core.py:185-227): all(sub_board > 0); survivors keep order
                int dst = p.H - 1;
                for (int r = p.H - 1; r >= 0; r--) {
                    bool full = true;
                    for (int c = 0; c < p.W; c++) full &= b[r * p.Wp + P + c] > 0;
                    if (full) { lines++; continue; }
                    if (dst != r) for (int c = 0; c < p.W; c++) b[dst * p.Wp + P + c] = b[r * p.Wp + P + c];
                    dst--;
                }
                // the n new top rows: the reference gathers them with jnp.take(sub_board, -H, fill_value=0); jnp.take's default
                // mode "fill" wraps negative indices numpy-style first, so -H is row 0 (in bounds): COPIES OF THE OLD ROW 0, not
                // zeros -- identical whenever row 0 is empty.  Row 0 itself is still in place here (rows are only moved downwards).
                for (; dst >= 1 && lines > 0; dst--) for (int c = 0; c < p.W; c++) b[dst * p.Wp + P + c] = b[P + c];
                lock_reward = lines == 0 ? 0 : (lines == 4 ? 800 : lines * 200 - 100);   // core.score
                // next piece: queue.bag_queue_get_next_element (functional/queue.py:38-67)
                if (sc[FN_QIDX] >= p.Q) { fn_new_bag(p, e, sc); piece = sc[FN_S]; sc[FN_QIDX] = 1; }
                else { piece = sc[FN_S + sc[FN_QIDX]]; sc[FN_QIDX] += 1; }
                rot = 0; x = spawn_x; y = 0;
                sc[FN_OVER] = fn_collision(p, b, c_cells[piece][0], x, y) ? 1 : 0;   // core.check_game_over
            }
            sc[FN_ACTIVE] = piece; sc[FN_ROT] = rot; sc[FN_X] = x; sc[FN_Y] = y;
            sc[FN_SCORE] = __float_as_int(old_score + (float)(drop_reward + lock_reward));
        }
        for (int i = 0; i < NS; i++) p.sc_out[e * NS + i] = sc[i];
        if (p.reward) p.reward[e] = __int_as_float(sc[FN_SCORE]) - old_score;
        if (p.terminated) p.terminated[e] = (uint8_t)sc[FN_OVER];
        if (p.lines) p.lines[e] = lines;
        // 3. get_observation (envs/tetris_fn.py:137-158): (board > 0) + active piece * (-1), cropped to H x W
        if (p.obs) {
            int8_t* o = s_obs + (size_t)tid * HW;
            for (int r = 0; r < p.H; r++)
                for (int c = 0; c < p.W; c++) o[r * p.W + c] = b[r * p.Wp + P + c] > 0 ? 1 : 0;
            if (!sc[FN_OVER]) {
                const uint32_t cells = c_cells[sc[FN_ACTIVE]][sc[FN_ROT]];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int c = (cells >> (4 * k)) & 15;
                    const int r = sc[FN_Y] + (c >> 2), col = sc[FN_X] + (c & 3) - P;
                    if ((unsigned)r < (unsigned)p.H && (unsigned)col < (unsigned)p.W) o[r * p.W + col] -= 1;
                }
            }
        }
    }
    __syncthreads();
    // 4. coalesced copies out: boards, observation tile
    {
        int8_t* dst = p.board_out + base * OB;
        if ((OB & 3) == 0 && ((uintptr_t)dst & 3) == 0) {
            const int obw = OB >> 2;
            int el = 0, k = tid;
            while (k >= obw) { k -= obw; el++; }
#pragma unroll 4
            while (el < nv) {
                ((uint32_t*)(dst + (size_t)el * OB))[k] = ((const uint32_t*)(s_board + (size_t)el * bstr))[k];
                k += T;
                while (k >= obw) { k -= obw; el++; }
            }
        } else {
            for (int el = 0; el < nv; el++)
                for (int k = tid; k < OB; k += T) dst[(size_t)el * OB + k] = s_board[(size_t)el * bstr + k];
        }
        if (p.obs) {
            int8_t* og = p.obs + base * HW;
            const int bytes = nv * HW;
            if ((bytes & 3) == 0 && ((uintptr_t)og & 3) == 0)
                for (int k = tid; k < (bytes >> 2); k += T) ((uint32_t*)og)[k] = ((const uint32_t*)s_obs)[k];
            else
                for (int k = tid; k < bytes; k += T) og[k] = s_obs[k];
        }
    }
}

// ---- tile variant (the default): CTA = 32 envs x 8 threads ---------------------------------------------------------------
// The tile's boards, scalars and observations are contiguous in global memory, so they move with 1-D bulk TMA copies (one
// thread issues them; full tiles) and sit in shared memory in the SAME layout.  Two thread mappings:
//   owner:  warp 0, lane = env -- the game logic (32 lanes on the divergent per-action branches instead of 8 x 4);
//   coop:   thread (e = tid / 8, t = tid % 8) -- the row scan of core.clear_filled_rows and the observation, a word at a time.
//   A  owner: action, gravity, lock decision, place the piece's cells            -> s_lock[e]
//   B  coop:  full-row mask of the envs that locked (rows t, t + 8, ...), OR-combined with three shuffles -> s_fm[e]
//   C  owner: row compaction when a row is full (rare), score, next piece, game-over test, scalars and 5-tuple out
//   D  coop:  observation words ((board > 0), cropped);  owner: active piece overlay
//   E  bulk stores of the three tiles
struct FnTileSmem {
    int off_sc, off_obs, off_misc, off_bar, bytes;
};
__host__ __device__ inline FnTileSmem fn_tile_smem(int OB, int HW, int NS, int E = 32) {
    FnTileSmem m;
    int o = (E * OB + 15) & ~15;
    m.off_sc = o; o += E * NS * 4;
    m.off_obs = o; o += (E * HW + 15) & ~15;
    m.off_misc = o; o += E * 4 * 3;           // lock flags, full-row masks (lo, hi)
    m.off_bar = o; o += 16;
    m.bytes = o;
    return m;
}

// bit 7 of byte i set <=> byte i of v is > 0 as int8 (non-zero, sign bit clear)
__device__ __forceinline__ uint32_t fn_pos4(uint32_t v) {
    return ((((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & ~v) & 0x80808080u;
}
// four consecutive bytes at shared-window byte address `a` (any alignment; may read up to 3 bytes past them)
__device__ __forceinline__ uint32_t fn_ld4(uint32_t a) {
    uint32_t w0, w1;
    const uint32_t al = a & ~3u;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(al));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(al + 4u));
    return __funnelshift_r(w0, w1, (a & 3u) * 8u);
}

// 32 registers (100 B of spills in the owner logic): 2048 threads = 64 warps per SM; 5 CTAs of 256 threads at 48 registers
// measured 18 % slower.  E = envs per tile (32 or 16): the CTA is a chain of barrier-separated phases with ONE warp on the game
// logic, so smaller tiles put more independent chains on an SM (E = 16: sixteen 128-thread CTAs per SM, owner lanes 0..15).
template <int E>
__global__ void __launch_bounds__(E * 8, 2048 / (E * 8)) k_fn_step_tile(const __grid_constant__ FnParams p) {
    extern __shared__ __align__(128) uint8_t fsm[];
    constexpr int TPE = 8, T = E * TPE;
    const int tid = threadIdx.x, e_l = tid >> 3, t = tid & 7;
    const int64_t base = (int64_t)blockIdx.x * E;
    const int nv = (int)min((int64_t)E, p.n - base);
    const int OB = p.Hp * p.Wp, NS = FN_S + p.Q, HW = p.H * p.W;
    const FnTileSmem m = fn_tile_smem(OB, HW, NS, E);
    int8_t* s_board = (int8_t*)fsm;                       // [E][OB] (+ 16 bytes of slack: word reads past a row stay inside)
    int32_t* s_sc = (int32_t*)(fsm + m.off_sc);           // [E][NS]
    int8_t* s_obs = (int8_t*)(fsm + m.off_obs);           // [E][HW]
    uint32_t* s_lock = (uint32_t*)(fsm + m.off_misc);     // [E]
    uint32_t* s_fm = s_lock + E;                          // [2][E]
    uint64_t* bar = (uint64_t*)(fsm + m.off_bar);
    const bool full_tile = nv == E;
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < E) s_lock[tid] = 0;
    __syncthreads();
    if (full_tile) {
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t)(E * (OB + NS * 4)));
            bulk_g2s(s_board, p.board_in + base * OB, (uint32_t)(E * OB), bar);
            bulk_g2s(s_sc, p.sc_in + base * NS, (uint32_t)(E * NS * 4), bar);
        }
        mbar_wait(bar, 0);
    } else {
        for (int k = tid; k < nv * OB; k += T) s_board[k] = p.board_in[base * OB + k];
        for (int k = tid; k < nv * NS; k += T) s_sc[k] = p.sc_in[base * NS + k];
        __syncthreads();
    }
    const bool live = e_l < nv;                 // coop mapping
    int8_t* b = s_board + (size_t)e_l * OB;
    const bool owner = tid < nv;                // owner mapping: warp 0, lane = env
    const int64_t e = base + tid;
    int8_t* ob = s_board + (size_t)(tid & (E - 1)) * OB;
    int32_t* sc = s_sc + (tid & (E - 1)) * NS;
    const int spawn_x = p.Wp / 2 - 2;   // core.get_initial_x_y: 4x4 matrices (functional/core.py:66-83)
    float old_score = 0.f;
    int piece = 0, rot = 0, x = 0, y = 0, drop_reward = 0;
    bool locked = false;
    // ---- A
    if (!p.actions) {
        // tetris_fn.reset (envs/tetris_fn.py:318-367): the board by all eight threads, the scalars by the owner
        if (live)
            for (int k = t; k < OB; k += TPE) {
                const int r = k / p.Wp, c = k - r * p.Wp;
                b[k] = (r < p.H && c >= P && c < P + p.W) ? 0 : 1;
            }
        if (owner) {
            sc[FN_KEY1] = 0;
            fn_new_bag(p, e, sc);
            sc[FN_ACTIVE] = sc[FN_S]; sc[FN_QIDX] = 1;
            sc[FN_ROT] = 0; sc[FN_X] = spawn_x; sc[FN_Y] = 0; sc[FN_OVER] = 0; sc[FN_SCORE] = __float_as_int(0.f);
        }
    } else if (owner) {
        old_score = __int_as_float(sc[FN_SCORE]);
        if (!sc[FN_OVER]) {
            // tetris_fn.update_state (envs/tetris_fn.py:161-273)
            const int a = p.actions[e];
            piece = sc[FN_ACTIVE]; rot = sc[FN_ROT]; x = sc[FN_X]; y = sc[FN_Y];
            uint32_t cells = c_cells[piece][rot];
            if ((unsigned)a <= 4u) {
                // left / right / down / rotate: ONE collision test of the candidate (the per-action branches, each with its own
                // inlined test, ran one after the other on the owner warp)
                const int dx = (a == 1) - (a == 0), dy = (a == 2);
                const int nr = (rot + (a == 4 ? 1 : (a == 3 ? 3 : 0))) & 3;   // 3 = counter-clockwise, 4 = clockwise (envs/tetris_fn.py:470-478)
                const uint32_t ncells = c_cells[piece][nr];
                if (!fn_collision(p, ob, ncells, x + dx, y + dy)) { x += dx; y += dy; rot = nr; cells = ncells; drop_reward = dy; }
            } else if (a == 6) {                          // core.hard_drop (functional/core.py:230-251): while !collision(y + 1): y++
                // column by column: the cells of a tetromino column are contiguous (rows top .. bot of the matrix), so the column first
                // collides at y' = max(y + 1, g - bot), g = the first filled board row at or below y + 1 + top (bedrock ends the scan)
                const uint4 pr = c_prec[piece][rot];
                const uint32_t bot4 = c_bot4[piece][rot];
                int nyp = 1 << 20;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if ((pr.y >> (8 * j)) & 1u) {
                        const int8_t* q = ob + (y + 1 + (int)((pr.z >> (8 * j)) & 3u)) * p.Wp + x + j;
                        int g = 0;
                        while (q[g * p.Wp] <= 0) g++;
                        nyp = min(nyp, max(y + 1, y + 1 + (int)((pr.z >> (8 * j)) & 3u) + g - (int)((bot4 >> (8 * j)) & 3u)));
                    }
                }
                drop_reward = 2 * (nyp - 1 - y);
                y = nyp - 1;
            }
            int yg = y;
            if (p.gravity && !fn_collision(p, ob, cells, x, y + 1)) yg = y + 1;   // core.graviy_step
            const bool should_lock = (yg == y) && p.gravity;
            y = yg;
            locked = should_lock || a == 6;
            if (locked) {
                // place_active_tetromino (envs/tetris_fn.py:370-413) / core.lock_active_tetromino
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    int c = (cells >> (4 * k)) & 15;
                    ob[(y + (c >> 2)) * p.Wp + x + (c & 3)] += (int8_t)(piece + 2);
                }
                s_lock[tid] = 1;
            }
        }
    }
    __syncthreads();
    // ---- B: core.clear_filled_rows' row test (functional/core.py:185-227: all(sub_board > 0)), every row of the envs that locked
    uint32_t fmask_lo = 0, fmask_hi = 0;
    if (live && s_lock[e_l]) {
        const uint32_t b0 = smem_u32(b) + P;
        for (int r = t; r < p.H; r += TPE) {
            const uint32_t a0 = b0 + r * p.Wp, al = a0 & ~3u, lead = a0 & 3u;
            const int nw = (int)(lead + p.W + 3) >> 2;
            uint32_t ok = 0x80808080u;
            for (int i = 0; i < nw; i++) {
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(al + 4u * i));
                if (i == 0) { const uint32_t lm = (1u << (8 * lead)) - 1u; v = (v & ~lm) | (0x01010101u & lm); }   // bytes in front of the row
                if (i == nw - 1) {
                    const uint32_t endb = (lead + p.W) & 3u;                                     // bytes of the last word inside the row
                    if (endb) v = (v & ((1u << (8 * endb)) - 1u)) | (0x01010101u & ~((1u << (8 * endb)) - 1u));
                }
                ok &= fn_pos4(v);
            }
            if (ok == 0x80808080u) { if (r < 32) fmask_lo |= 1u << r; else fmask_hi |= 1u << (r - 32); }
        }
    }
#pragma unroll
    for (int o = 1; o < TPE; o <<= 1) {
        fmask_lo |= __shfl_xor_sync(0xffffffffu, fmask_lo, o);
        fmask_hi |= __shfl_xor_sync(0xffffffffu, fmask_hi, o);
    }
    if (t == 0) { s_fm[e_l] = fmask_lo; s_fm[E + e_l] = fmask_hi; }
    __syncthreads();
    // ---- C
    int lines = 0;
    if (owner && p.actions) {
        if (locked) {
            fmask_lo = s_fm[tid]; fmask_hi = s_fm[E + tid];
            lines = __popc(fmask_lo) + __popc(fmask_hi);
            if (lines) {
                // survivors keep their order, packed towards the floor
                int dst = p.H - 1;
                for (int r = p.H - 1; r >= 0; r--) {
                    const bool full = r < 32 ? (fmask_lo >> r) & 1u : (fmask_hi >> (r - 32)) & 1u;
                    if (full) continue;
                    if (dst != r) for (int c = 0; c < p.W; c++) ob[dst * p.Wp + P + c] = ob[r * p.Wp + P + c];
                    dst--;
                }
                // the n new top rows: the reference gathers them with jnp.take(sub_board, -H, fill_value=0); jnp.take's default
                // mode "fill" wraps negative indices numpy-style first, so -H is row 0 (in bounds): COPIES OF THE OLD ROW 0, not
                // zeros -- identical whenever row 0 is empty.  Row 0 itself is still in place here (rows are only moved downwards).
                for (; dst >= 1; dst--) for (int c = 0; c < p.W; c++) ob[dst * p.Wp + P + c] = ob[P + c];
            }
            const int lock_reward = lines == 0 ? 0 : (lines == 4 ? 800 : lines * 200 - 100);   // core.score
            drop_reward += lock_reward;
            // next piece: queue.bag_queue_get_next_element (functional/queue.py:38-67)
            if (sc[FN_QIDX] >= p.Q) { fn_new_bag(p, e, sc); piece = sc[FN_S]; sc[FN_QIDX] = 1; }
            else { piece = sc[FN_S + sc[FN_QIDX]]; sc[FN_QIDX] += 1; }
            rot = 0; x = spawn_x; y = 0;
            sc[FN_OVER] = fn_collision(p, ob, c_cells[piece][0], x, y) ? 1 : 0;   // core.check_game_over
        }
        if (locked || !sc[FN_OVER]) {   // (a finished game is frozen: nothing is written)
            sc[FN_ACTIVE] = piece; sc[FN_ROT] = rot; sc[FN_X] = x; sc[FN_Y] = y;
            sc[FN_SCORE] = __float_as_int(old_score + (float)drop_reward);
        }
    }
    if (owner) {
        if (p.reward) p.reward[e] = __int_as_float(sc[FN_SCORE]) - old_score;
        if (p.terminated) p.terminated[e] = (uint8_t)sc[FN_OVER];
        if (p.lines) p.lines[e] = lines;
    }
    __syncthreads();
    // ---- D: get_observation (envs/tetris_fn.py:137-158): (board > 0) + active piece * (-1), cropped to H x W
    if (p.obs) {
        if (live) {
            int8_t* o = s_obs + (size_t)e_l * HW;
            const uint32_t invW = p.invW;                        // i / W for i < 4096
            if ((HW & 3) == 0 && (OB & 3) == 0 && p.W >= 4) {
                // Wp = W + 2 P and P = 4: cell (r, c) sits at board byte r Wp + 4 + c = r W + c (mod 4), the alignment of its
                // observation byte r W + c.  Observation word k (cells 4k .. 4k + 3) is therefore ONE aligned board word when it
                // lies in one row, and a byte-wise select of two aligned board words when it continues in the next row.
                const uint32_t b0 = smem_u32(b) + P;
                for (int k = t; k < (HW >> 2); k += TPE) {
                    const uint32_t i = 4u * k, r = (i * invW) >> 16, n1 = (r + 1u) * p.W - i;   // cells left in row r
                    // board byte of cell i: r Wp + c = i + 8 r; the cells of the next row sit Wp - W = 8 bytes further, same byte positions
                    const uint32_t a = b0 + i + 8u * r;
                    uint32_t v, v2;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
                    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(v2) : "r"(a));        // (masked out when the word stays in its row)
                    const uint32_t keep = n1 >= 4u ? 0xFFFFFFFFu : (1u << (8 * n1)) - 1u;
                    v = (v & keep) | (v2 & ~keep);
                    ((uint32_t*)o)[k] = fn_pos4(v) >> 7;
                }
            } else {
                for (int i = t; i < HW; i += TPE) {
                    const uint32_t r = ((uint32_t)i * invW) >> 16, c = i - r * p.W;
                    o[i] = b[r * p.Wp + P + c] > 0;
                }
            }
        }
        __syncthreads();
        if (owner && !sc[FN_OVER]) {
            int8_t* o = s_obs + (size_t)tid * HW;
            const uint32_t cells = c_cells[sc[FN_ACTIVE]][sc[FN_ROT]];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int c = (cells >> (4 * k)) & 15;
                const int r = sc[FN_Y] + (c >> 2), col = sc[FN_X] + (c & 3) - P;
                if ((unsigned)r < (unsigned)p.H && (unsigned)col < (unsigned)p.W) o[r * p.W + col] -= 1;
            }
        }
    }
    // ---- E
    if (full_tile) {
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(p.board_out + base * OB, s_board, (uint32_t)(E * OB));
            bulk_s2g(p.sc_out + base * NS, s_sc, (uint32_t)(E * NS * 4));
            if (p.obs) bulk_s2g(p.obs + base * HW, s_obs, (uint32_t)(E * HW));
            bulk_commit();
            bulk_wait_read();
        }
    } else {
        __syncthreads();
        for (int k = tid; k < nv * OB; k += T) p.board_out[base * OB + k] = s_board[k];
        for (int k = tid; k < nv * NS; k += T) p.sc_out[base * NS + k] = s_sc[k];
        if (p.obs) for (int k = tid; k < nv * HW; k += T) p.obs[base * HW + k] = s_obs[k];
    }
}

}  // namespace tg
