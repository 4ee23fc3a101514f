// tg_step.cuh -- the per-call batched step / reset kernel (BASELINE config 2).
//
// One persistent CTA loops over tiles of E = blockDim.x consecutive envs:
//   1. TMA bulk loads (cp.async.bulk, mbarrier completion) bring the tile's hot records and board
//      records HBM -> shared memory as two contiguous copies;
//   2. thread e runs the game logic of env e on shared memory (bitboard collision / drop / commit);
//   3. all threads expand the nibble id planes into the padded uint8 board image, the mask image,
//      the holder and the queue images, which live in shared memory with their constant parts
//      (bedrock, zeros) written once per CTA;
//   4. TMA bulk stores write the four observation tiles and the hot tile back as contiguous,
//      fully coalesced copies; board records are written back only for envs that committed a piece.
// HBM traffic per env-step = hot 32 R + 32 W, board record R (+ W on commit), action 4,
// outputs 10, observation dict Hp*Wp*2 + 16 + 16Q.
#pragma once
#include "tg_device.cuh"

namespace tg {

// ---- PTX wrappers: mbarrier + bulk async copies (TMA, 1-D) ----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// contiguous smem -> global copy: TMA when size/alignment allow, cooperative stores otherwise
__device__ __forceinline__ void tile_store(uint8_t* g, const uint8_t* s, uint32_t bytes, bool leader, int tid, int nthreads) {
    if ((bytes & 15u) == 0 && ((uintptr_t)g & 15u) == 0) {
        if (leader && bytes) bulk_s2g(g, s, bytes);
    } else {
        for (uint32_t i = tid; i < bytes; i += nthreads) g[i] = s[i];
    }
}

struct StepParams {
    DevCfg cfg;
    int64_t n;
    uint8_t* hot; uint8_t* board; uint8_t* rng; const uint8_t* seq;
    const int32_t* actions;          // step mode
    const uint64_t* seeds;           // reset mode (nullable)
    const uint8_t* reset_mask;       // reset mode (nullable)
    uint8_t *o_board, *o_mask, *o_holder, *o_queue;
    float* reward; uint8_t* terminated; uint8_t* truncated; int32_t* lines;
    double* stats;                   // nullable: episodes, sum_return, sum_length, sum_lines
    const uint8_t* legal;            // grouped mode: legal mask of the previous observation u8[n][A]
    uint8_t* info_board;             // grouped mode (nullable): features of the real observation u8[n][F]
    uint8_t* fill_high;              // grouped mode: u8[n], 1 = illegal action terminated the episode
    int mode;                        // 0 = step, 1 = reset, 2 = grouped placement step
    // shared-memory carve-up (bytes from the 128-aligned base)
    int off_hot, off_brd, off_iboard, off_imask, off_iholder, off_iqueue, off_bar;
};

// expand 8 nibbles -> 8 id bytes (two words)
__device__ __forceinline__ void nib8_to_bytes(uint32_t x, uint32_t& b0, uint32_t& b1) {
    uint32_t lo = x & 0x0F0F0F0Fu, hi = (x >> 4) & 0x0F0F0F0Fu;
    b0 = __byte_perm(lo, hi, 0x5140);
    b1 = __byte_perm(lo, hi, 0x7362);
}

// Writes the W cell bytes of playfield row `row` of one env into its padded board image.
// Only words that contain cell bytes are touched; the spill-over bytes are bedrock (1).
template <int WT>
__device__ __forceinline__ void fill_board_row(const DevCfg& cfg, const uint32_t* ids, uint8_t* tile, int env_off, int row) {
    const int W = WT ? WT : cfg.W;
    const int Wp = W + 2 * P;
    constexpr int MAXC = WT ? (WT + 7) / 8 : 3;       // 8-nibble chunks per row (W <= 24)
    constexpr int MAXW = 2 * MAXC + 1;
    uint32_t cw[MAXW + 1];
    const int nchunk = (W + 7) >> 3;
#pragma unroll
    for (int k = 0; k < MAXC; k++) {
        if (k < nchunk) {
            uint32_t x = ids_get8(ids, row * W + 8 * k);
            nib8_to_bytes(x, cw[2 * k], cw[2 * k + 1]);
        } else { cw[2 * k] = 0x01010101u; cw[2 * k + 1] = 0x01010101u; }
    }
    cw[2 * MAXC] = 0x01010101u; cw[MAXW] = 0x01010101u;
    // cells beyond W inside the last words are bedrock
    const int nw = (W + 3) >> 2;  // words holding cells
#pragma unroll
    for (int j = 0; j < MAXW; j++) {
        int valid = W - 4 * j;  // cell bytes in word j
        if (valid <= 0) cw[j] = 0x01010101u;
        else if (valid < 4) { uint32_t m = (1u << (8 * valid)) - 1; cw[j] = (cw[j] & m) | (0x01010101u & ~m); }
    }
    // byte offset of the first cell inside the image TILE (env images are packed back to back, so
    // the bytes just before an unaligned start are the previous row's / previous env's bedrock)
    const int cbase = env_off + row * Wp + P;
    const int a = cbase & 3;
    uint32_t* out = (uint32_t*)(tile + (cbase - a));
    const uint32_t sel = 0x7654u - 0x1111u * (uint32_t)a;
    const int nout = (a + W + 3) >> 2;
    uint32_t prev = 0x01010101u;
#pragma unroll
    for (int j = 0; j < MAXW; j++) {
        if (j < nout) {
            uint32_t cur = (j < nw) ? cw[j] : 0x01010101u;
            out[j] = __byte_perm(prev, cur, sel);
            prev = cur;
        }
    }
}

template <int WT, int HT, class COLT>
__global__ void k_step(const __grid_constant__ StepParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const DevCfg& cfg = p.cfg;
    const int E = blockDim.x, tid = threadIdx.x;
    const int W = WT ? WT : cfg.W, H = HT ? HT : cfg.H;
    const int Wp = W + 2 * P, Hp = H + P;
    const int OB = Hp * Wp, OQ = cfg.OQ, BS = cfg.board_stride;

    uint32_t* s_hot = (uint32_t*)(smem + p.off_hot);
    uint8_t* s_brd = smem + p.off_brd;
    uint8_t* i_board = smem + p.off_iboard;
    uint8_t* i_mask = smem + p.off_imask;
    uint8_t* i_holder = smem + p.off_iholder;
    uint8_t* i_queue = smem + p.off_iqueue;
    uint64_t* bar = (uint64_t*)(smem + p.off_bar);

    // constant parts of the images: bedrock frame, empty mask (written once per CTA)
    for (int i = tid; i < E * OB; i += E) {
        int b = i % OB, r = b / Wp, c = b - r * Wp;
        i_board[i] = (r < H && c >= P && c < P + W) ? 0 : 1;
        i_mask[i] = 0;
    }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();

    const int64_t ntiles = (p.n + E - 1) / E;
    uint32_t parity = 0;
    int pm_x = 0, pm_y = 0, pm_n = 0;  // bounding box this thread drew into the mask image last tile
    double st_ep = 0, st_ret = 0, st_len = 0, st_lines = 0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * E;
        const int nv = (int)min((int64_t)E, p.n - base);
        // (A) previous tile's stores must have finished reading shared memory
        bulk_wait_read();
        __syncthreads();
        for (int i = 0; i < pm_n; i++)
            for (int j = 0; j < pm_n; j++) i_mask[tid * OB + (pm_y + i) * Wp + pm_x + j] = 0;
        // (B) bulk loads of the tile's state
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t)(nv * 32 + nv * BS));
            bulk_g2s(s_hot, p.hot + base * 32, (uint32_t)(nv * 32), bar);
            bulk_g2s(s_brd, p.board + base * BS, (uint32_t)(nv * BS), bar);
        }
        int action = 0;
        if (p.mode != 1 && tid < nv) action = p.actions[base + tid];
        mbar_wait(bar, parity);
        parity ^= 1;

        // (C) game logic, one thread per env
        Hot h;
        StepResult res;
        res.dirty = 0;
        uint32_t* rec = (uint32_t*)(s_brd + tid * BS);
        COLT Bact = 0;
        if (tid < nv) {
            const int64_t e = base + tid;
            hot_load(h, s_hot + tid * 8);
            Rng g;
            g.rec = (uint32_t*)(p.rng + e * cfg.rng_stride);
            g.seq = p.seq ? p.seq + e * cfg.seq_len : nullptr;
            g.gid = cfg.env_id_offset + (uint64_t)e;
            res.reward = 0; res.lines = 0; res.terminated = 0;
            bool need_reset = false;
            if (p.mode == 1) {
                need_reset = (!p.reset_mask || p.reset_mask[e]);
                if (need_reset && p.seeds && cfg.rng_mode == 0) { ((uint64_t*)g.rec)[0] = p.seeds[e]; g.rec[2] = 0; }
            } else if (cfg.autoreset == 1 && h.pending) {
                need_reset = true;  // gymnasium NEXT_STEP autoreset: the action is ignored, the env is reset
            } else {
                bool stepped = true;
                if (p.mode == 2) {
                    // GroupedActionsObservations.step (wrappers/grouped.py:209-269)
                    bool ok = (unsigned)action < (unsigned)cfg.A && p.legal[e * cfg.A + action] != 0;
                    p.fill_high[e] = (uint8_t)(!ok && cfg.terminate_on_illegal);
                    if (ok) {
                        h.x = (action >> 2) + P - c_n[h.p] / 2;   // y untouched (wrappers/grouped.py:244-254)
                        h.r = (h.r + (action & 3)) & 3;
                        env_step<COLT>(cfg, h, rec, g, cfg.act_hard, res);
                    } else if (cfg.terminate_on_illegal) {
                        stepped = false;                           // env untouched, episode ends
                        res.reward = cfg.r_invalid; res.terminated = 1;
                    } else {
                        env_step<COLT>(cfg, h, rec, g, cfg.act_noop, res);
                        res.reward = cfg.r_invalid;
                    }
                } else {
                    env_step<COLT>(cfg, h, rec, g, action, res);
                }
                (void)stepped;
                h.ep_ret += (float)res.reward; h.ep_len += 1; h.ep_lines += res.lines;
                if (res.terminated) {
                    st_ep += 1; st_ret += h.ep_ret; st_len += h.ep_len; st_lines += h.ep_lines;
                    h.ep_ret = 0; h.ep_len = 0; h.ep_lines = 0;
                    if (cfg.autoreset == 1) h.pending = 1;
                    else if (cfg.autoreset == 2) need_reset = true;
                }
            }
            if (need_reset) { env_reset<COLT>(cfg, h, rec, g); res.dirty = 1; }
            hot_store(h, s_hot + tid * 8);
            if (p.mode == 2 && need_reset) p.fill_high[e] = 0;
            if (p.mode != 1) {
                p.reward[e] = (float)res.reward;
                p.terminated[e] = (uint8_t)res.terminated;
                p.truncated[e] = 0;
                p.lines[e] = res.lines;
            }
            Bact = bmask<COLT>((const COLT*)rec, W, c_cells[h.p][h.r], h.x);
            if (p.mode == 2 && p.info_board) {
                // info["board"]: FeatureVectorObservation of the real observation (wrappers/grouped.py:260-264)
                uint8_t f[32];
                int ln;
                placement_features<COLT>(cfg, (const COLT*)rec, c_cells[h.p][h.r], h.x, h.y, !((Bact >> h.y) & 1), false,
                                         COLT(3), f, ln);
                for (int i = 0; i < cfg.F; i++) p.info_board[e * cfg.F + i] = f[i];
            }
        }
        __syncthreads();
        const bool want_obs = p.o_board != nullptr;

        // (D) observation images: board rows, queue, holder
        if (want_obs)
        for (int it = tid; it < nv * H; it += E) {
            int e = it / H, row = it - e * H;
            fill_board_row<WT>(cfg, (const uint32_t*)(s_brd + e * BS + cfg.ids_off), i_board, e * OB, row);
        }
        if (want_obs) {
            const int Q = cfg.Q;
            for (int it = tid; it < nv * 4 * Q; it += E) {
                int e = it / (4 * Q), rem = it - e * 4 * Q, i = rem / Q, q = rem - i * Q;
                uint32_t w2 = s_hot[e * 8 + 2], w3 = s_hot[e * 8 + 3];
                uint64_t queue = (uint64_t)w2 | ((uint64_t)w3 << 32);
                int pc = (int)((queue >> (4 * q)) & 15u);
                ((uint32_t*)i_queue)[e * 4 * Q + i * Q + q] = c_rowbytes[pc][0][i];
            }
            for (int it = tid; it < nv * 4; it += E) {
                int e = it >> 2, i = it & 3;
                uint32_t a = s_hot[e * 8];
                int hold = (a >> 18) & 15, hr = (a >> 22) & 3;
                ((uint32_t*)i_holder)[e * 4 + i] = hold ? c_rowbytes[hold - 1][hr][i] : 0x01010101u;
            }
        }
        __syncthreads();
        // (E) active piece overlay + bounding-box mask (Tetris._get_obs, envs/tetris.py:566-576)
        if (tid < nv && want_obs) {
            uint8_t* ib = i_board + tid * OB;
            if (!((Bact >> h.y) & 1)) {
                uint32_t cells = c_cells[h.p][h.r];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    int c = (cells >> (4 * k)) & 15;
                    ib[(h.y + (c >> 2)) * Wp + h.x + (c & 3)] = (uint8_t)(h.p + 2);
                }
            }
            pm_n = c_n[h.p]; pm_x = h.x; pm_y = h.y;
            for (int i = 0; i < pm_n; i++)
                for (int j = 0; j < pm_n; j++) i_mask[tid * OB + (pm_y + i) * Wp + pm_x + j] = 1;
        } else pm_n = 0;
        // (F) stores
        fence_async_smem();
        __syncthreads();
        const bool leader = (tid == 0);
        if (want_obs) {
            tile_store(p.o_board + base * OB, i_board, (uint32_t)(nv * OB), leader, tid, E);
            tile_store(p.o_mask + base * OB, i_mask, (uint32_t)(nv * OB), leader, tid, E);
            if (leader) {
                bulk_s2g(p.o_holder + base * 16, i_holder, (uint32_t)(nv * 16));
                bulk_s2g(p.o_queue + base * OQ, i_queue, (uint32_t)(nv * OQ));
            }
        }
        if (leader) bulk_s2g(p.hot + base * 32, s_hot, (uint32_t)(nv * 32));
        if (res.dirty && tid < nv) bulk_s2g(p.board + (base + tid) * BS, s_brd + tid * BS, (uint32_t)BS);
        bulk_commit();
    }
    bulk_wait_all();
    if (p.stats) {
        for (int o = 16; o > 0; o >>= 1) {
            st_ep += __shfl_xor_sync(0xffffffffu, st_ep, o);
            st_ret += __shfl_xor_sync(0xffffffffu, st_ret, o);
            st_len += __shfl_xor_sync(0xffffffffu, st_len, o);
            st_lines += __shfl_xor_sync(0xffffffffu, st_lines, o);
        }
        if ((tid & 31) == 0 && st_ep > 0) {
            atomicAdd(p.stats + 0, st_ep); atomicAdd(p.stats + 1, st_ret);
            atomicAdd(p.stats + 2, st_len); atomicAdd(p.stats + 3, st_lines);
        }
    }
}

}  // namespace tg
