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

// numpy's SeedSequence(seed) -> PCG64 seeding on the device (Randomizer.reset, components/tetromino_randomizer.py:40-43:
// `np.random.default_rng(seed)`): the seed's two 32-bit words are hashed into the 4-word pool (hashmix / mix), 8 words are
// generated (generate_state(4, uint64)) and PCG64 is seeded with pcg_setseq_128_srandom_r(state = w0:w1, seq = w2:w3).
// Published algorithm of numpy/random/bit_generator.pyx + src/pcg64/pcg64.h; pinned against numpy for 10^5 seeds
// (tests/test_gpu_seeding.py; the same arithmetic in numpy form: oracle/np_seed.py, tests/test_oracle_seeding.py).
__device__ __forceinline__ uint32_t ss_hashmix(uint32_t v, uint32_t& hc) {
    v ^= hc; hc *= 0x931E8875u; v *= hc;
    return v ^ (v >> 16);
}
__device__ __forceinline__ uint32_t ss_mix(uint32_t x, uint32_t y) {
    const uint32_t r = 0xCA01F9DDu * x - 0x4973F715u * y;
    return r ^ (r >> 16);
}
__global__ void k_seed_numpy_seeds(uint8_t* rng, int stride, int64_t n, const uint64_t* seeds, const uint8_t* mask) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n || (mask && !mask[e])) return;
    const uint64_t seed = seeds[e];
    uint32_t hc = 0x43B0D7E5u, pool[4];
    pool[0] = ss_hashmix((uint32_t)seed, hc);
    pool[1] = ss_hashmix((uint32_t)(seed >> 32), hc);   // an absent word hashes like 0
    pool[2] = ss_hashmix(0u, hc);
    pool[3] = ss_hashmix(0u, hc);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (i != j) pool[j] = ss_mix(pool[j], ss_hashmix(pool[i], hc));
    uint32_t hb = 0x8B51F9DDu, w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t d = pool[i & 3] ^ hb;
        hb *= 0x58F38DEDu;
        d *= hb;
        w[i] = d ^ (d >> 16);
    }
    const unsigned __int128 MULT = ((unsigned __int128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
    const unsigned __int128 initstate = ((unsigned __int128)((uint64_t)w[0] | ((uint64_t)w[1] << 32)) << 64) | ((uint64_t)w[2] | ((uint64_t)w[3] << 32));
    const unsigned __int128 initseq = ((unsigned __int128)((uint64_t)w[4] | ((uint64_t)w[5] << 32)) << 64) | ((uint64_t)w[6] | ((uint64_t)w[7] << 32));
    const unsigned __int128 inc = (initseq << 1) | 1;
    unsigned __int128 state = inc;          // (0 * MULT + inc)
    state += initstate;
    state = state * MULT + inc;
    uint64_t* r = (uint64_t*)(rng + e * stride);
    r[0] = (uint64_t)(state >> 64); r[1] = (uint64_t)state; r[2] = (uint64_t)(inc >> 64); r[3] = (uint64_t)inc;
    r[4] = 0; r[5] = 0;
}

// Compact host step (tg_step_host, TG_HOST_COMPACT): what the observation dict is a function of, packed for the PCIe link --
// per env `pk` bytes = hot words 0, 2, 3 (position / piece / rotation / holder, queue; plus word 7, the holder FIFO, when
// holder_size > 1) followed by the nibble id plane.  Thread = (env, word), coalesced word stores.
__global__ void k_pack_host(const uint8_t* __restrict__ hot, const uint8_t* __restrict__ board, int board_stride, int ids_off,
                            int ids_words, int pk_words, int hdr_words, int64_t n, uint32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * pk_words) return;
    const int64_t e = i / pk_words;
    const int w = (int)(i - e * pk_words);
    uint32_t v = 0;
    if (w < hdr_words) v = ((const uint32_t*)(hot + e * 32))[w == 0 ? 0 : (w == 3 ? 7 : w + 1)];     // hot words 0, 2, 3 (, 7)
    else if (w - hdr_words < ids_words) v = ((const uint32_t*)(board + e * board_stride + ids_off))[w - hdr_words];
    out[i] = v;
}

template <class COLT>
__global__ void k_get_state(const DevCfg cfg, int64_t n, const uint8_t* hot, const uint8_t* board, uint8_t* o_board,
                            int32_t* o_scalars) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    Hot h;
    hot_load(h, (const uint32_t*)(hot + e * 32));
    if (o_scalars) {
        const int HS = cfg.holder_size, NSC = 8 + cfg.Q + (HS > 1 ? 2 * HS : 0);
        int32_t* s = o_scalars + e * NSC;
        s[0] = h.x; s[1] = h.y; s[2] = h.p; s[3] = h.r; s[4] = h.hold ? h.hold - 1 : -1; s[5] = h.hold_r;
        if (HS > 1) {   // columns 4 / 5: number of held pieces / 0; the (piece, rotation) pairs follow the queue, oldest first
            const int cnt = (int)(h.hq & 7u);
            s[4] = cnt; s[5] = 0;
            for (int k = 0; k < HS; k++) {
                const uint32_t sl = (h.hq >> (3 + 5 * k)) & 31u;
                s[8 + cfg.Q + 2 * k] = k < cnt ? (int)(sl & 7u) : -1;
                s[8 + cfg.Q + 2 * k + 1] = k < cnt ? (int)(sl >> 3) : 0;
            }
        }
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
        const int HS = cfg.holder_size, NSC = 8 + cfg.Q + (HS > 1 ? 2 * HS : 0);
        const int32_t* s = i_scalars + e * NSC;
        // out-of-range pokes are clamped into the record's bit fields (x: 6 bits inside the padded width, y: 7 bits inside the
        // padded height, piece 0..6): a bad value must not spill into the neighbouring fields or index past the piece tables
        h.x = min(max(s[0], 0), cfg.Wp - 1); h.y = min(max(s[1], 0), cfg.Hp - 1); h.p = min(max(s[2], 0), 6); h.r = s[3] & 3;
        h.hold = s[4] < 0 ? 0 : min(s[4], 6) + 1; h.hold_r = s[5] & 3;
        if (HS > 1) {
            h.hold = 0; h.hold_r = 0;
            uint32_t slots = 0;
            int cnt = 0;
            for (int k = 0; k < HS; k++) {
                const int pc = s[8 + cfg.Q + 2 * k];
                if (pc < 0) break;
                slots |= (uint32_t)(min(pc, 6) | ((s[8 + cfg.Q + 2 * k + 1] & 3) << 3)) << (5 * cnt);
                cnt++;
            }
            h.hq = (slots << 3) | (uint32_t)cnt;
        }
        h.swapped = s[6] != 0; h.over = s[7] != 0;
        h.pending = 0;
        h.queue = 0;
        for (int q = 0; q < cfg.Q; q++) h.queue |= (uint64_t)min(max(s[8 + q], 0), 6) << (4 * q);
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
