// tg_aux.cuh -- state conversion kernels (canonical unpacked views <-> packed HBM records) and seeding.
// Not on the hot path: they back tg_get_state / tg_set_state (the reference tests poke
// env.unwrapped.board / x / y / active_tetromino directly; Tetris.get_state/set_state, envs/tetris.py:681-708).
#pragma once
#include "tg_device.cuh"

namespace tg {

__global__ void k_seed_numpy(uint8_t* rng, int stride, int64_t n, const uint64_t* pcg, const uint8_t* mask) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n || (mask && !mask[e])) return;
    uint64_t* r = (uint64_t*)(rng + e * stride);
    r[0] = pcg[e * 4 + 0]; r[1] = pcg[e * 4 + 1]; r[2] = pcg[e * 4 + 2]; r[3] = pcg[e * 4 + 3];
    r[4] = 0;  // has_uint32 = 0, uinteger = 0
    r[5] = 0;
}

template <class COLT>
__global__ void k_get_state(const DevCfg cfg, int64_t n, const uint8_t* hot, const uint8_t* board, uint8_t* o_board,
                            int32_t* o_scalars) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    Hot h;
    hot_load(h, (const uint32_t*)(hot + e * 32));
    if (o_scalars) {
        int32_t* s = o_scalars + e * (8 + cfg.Q);
        s[0] = h.x; s[1] = h.y; s[2] = h.p; s[3] = h.r; s[4] = h.hold ? h.hold - 1 : -1; s[5] = h.hold_r;
        s[6] = h.swapped; s[7] = h.over;
        for (int q = 0; q < cfg.Q; q++) s[8 + q] = (int)((h.queue >> (4 * q)) & 15u);
    }
    if (o_board) {
        const uint8_t* rec = board + e * cfg.board_stride;
        const COLT* cols = (const COLT*)rec;
        const uint32_t* ids = (const uint32_t*)(rec + cfg.ids_off);
        uint8_t* ob = o_board + e * cfg.OB;
        for (int r = 0; r < cfg.Hp; r++)
            for (int c = 0; c < cfg.Wp; c++) {
                uint8_t v = 1;
                if (r < cfg.H && c >= P && c < P + cfg.W) {
                    v = (uint8_t)ids_get1(ids, r * cfg.W + (c - P));
                    // occupancy and ids must agree (debug aid: 15 marks a mismatch)
                    if (((cols[c - P] >> r) & 1) != (COLT)(v != 0)) v = 15;
                }
                ob[r * cfg.Wp + c] = v;
            }
    }
}

template <class COLT>
__global__ void k_set_state(const DevCfg cfg, int64_t n, uint8_t* hot, uint8_t* board, const uint8_t* i_board,
                            const int32_t* i_scalars, const uint8_t* mask) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n || (mask && !mask[e])) return;
    if (i_scalars) {
        Hot h;
        hot_load(h, (const uint32_t*)(hot + e * 32));
        const int32_t* s = i_scalars + e * (8 + cfg.Q);
        h.x = s[0]; h.y = s[1]; h.p = s[2]; h.r = s[3] & 3; h.hold = s[4] < 0 ? 0 : s[4] + 1; h.hold_r = s[5] & 3;
        h.swapped = s[6] != 0; h.over = s[7] != 0;
        h.pending = 0;
        h.queue = 0;
        for (int q = 0; q < cfg.Q; q++) h.queue |= (uint64_t)(s[8 + q] & 15) << (4 * q);
        hot_store(h, (uint32_t*)(hot + e * 32));
    }
    if (i_board) {
        uint8_t* rec = board + e * cfg.board_stride;
        COLT* cols = (COLT*)rec;
        uint32_t* ids = (uint32_t*)(rec + cfg.ids_off);
        const uint8_t* ib = i_board + e * cfg.OB;
        COLT fl = floor_bits<COLT>(cfg.H, cfg.Hp);
        for (int i = 0; i < cfg.ids_words; i++) ids[i] = 0;
        for (int c = 0; c < cfg.W; c++) {
            COLT v = fl;
            for (int r = 0; r < cfg.H; r++) {
                uint8_t b = ib[r * cfg.Wp + c + P];
                if (b) { v |= COLT(1) << r; ids_set1(ids, r * cfg.W + c, b & 15u); }
            }
            cols[c] = v;
        }
    }
}

}  // namespace tg
