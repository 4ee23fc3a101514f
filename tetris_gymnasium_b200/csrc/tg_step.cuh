// tg_step.cuh -- the per-call batched step / reset kernels (BASELINE config 2).
//
// k_step_ws (the default): persistent, warp-specialised CTAs loop over tiles of E = 32 consecutive envs
//   * NL logic warps (lane = env) run the game logic on TMA-staged records in shared memory, NS >= NL + 2 state stages in flight
//     (cp.async.bulk + mbarrier), each logic warp takes every NL-th tile of its CTA;
//   * the image / store warps expand the nibble id planes into the padded uint8 board image, the mask image, the holder
//     and the queue images (constant parts -- bedrock, zeros -- written once per CTA) and issue the TMA bulk stores:
//     four observation tiles + the hot tile as contiguous copies, board records only for envs that committed a piece;
//   * one instantiation per mode (step / reset / grouped placement step): the step instantiation's speed depends on its
//     code footprint, so it carries neither the reset-mode nor the grouped code;
//   * launched with programmatic stream serialization: the prologue overlaps the previous kernel's tail.
// k_step: the two-stage variant without warp specialisation (very large boards, TG_WS=0).
// HBM traffic per env-step = hot 32 R + 32 W, board record R (+ W on commit), rng record R, action 4,
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
// smem -> global bulk copy of the observation images; -DTG_L2_HINT adds an L2 evict_first hint (createpolicy).  Measured
// (4 M envs, A/B on one box, 4 runs each): selected at run time inside one binary the hint gained 3 %, but the run-time
// switch cost 4 % (code size); compiled in unconditionally it LOST 1.7 % against no hint (4.18 vs 4.25 G) -> off.
__device__ __forceinline__ void bulk_s2g_stream(void* dst_gmem, const void* src_smem, uint32_t bytes) {
#ifndef TG_L2_HINT
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
#else
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes), "l"(pol)
                 : "memory");
#endif
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }   // all but the newest group
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// contiguous smem -> global copy: TMA when size/alignment allow, cooperative stores otherwise
template <bool STREAM = false>
__device__ __forceinline__ void tile_store(uint8_t* g, const uint8_t* s, uint32_t bytes, bool leader, int tid, int nthreads) {
    if ((bytes & 15u) == 0 && ((uintptr_t)g & 15u) == 0) {
        if (leader && bytes) { if (STREAM) bulk_s2g_stream(g, s, bytes); else bulk_s2g(g, s, bytes); }
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
    int E;                           // envs per tile
    int NL;                          // k_step_ws: logic warps per CTA
    int NS;                          // k_step_ws: state stages in flight (>= NL + 2 and a multiple of NL: tg_api.cu build_plan)
    int whole_tile_min;              // k_step_ws: dirty envs in a tile from which the board records leave as one bulk copy
    // shared-memory carve-up (bytes from the 128-aligned base)
    int off_hot, off_brd, off_rng, off_iboard, off_imask, off_iholder, off_iqueue, off_bar, off_box, off_tab, off_feat;
    int st_hot, st_brd, st_rng;      // bytes between the two pipeline stages of each state buffer
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
__device__ __forceinline__ void fill_board_row(const DevCfg& cfg, const uint32_t* ids, uint8_t* tile, int env_off, int row,
                                               int row_stride = 0) {
    const int W = WT ? WT : cfg.W;
    const int Wp = row_stride ? row_stride : W + 2 * P;
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


// ---- specialised row expansion -------------------------------------------------------------------
// W = 10: 4 rows = 160 bits = 5 words of the id plane -> 72 output bytes (8-byte aligned); only the
// 12 words that contain cells are written (4 x STS.64 + 4 x STS.32), the bedrock words persist.
__device__ __forceinline__ uint32_t w10_tail(uint32_t hi8) { return (hi8 & 15u) | ((hi8 & 0xF0u) << 4); }
__device__ __forceinline__ void fill_rows4_w10(const uint32_t* ids5, uint8_t* out72) {
    uint32_t w0 = ids5[0], w1 = ids5[1], w2 = ids5[2], w3 = ids5[3], w4 = ids5[4];
    uint32_t a0, a1, b0, b1, c0, c1, d0, d1;
    nib8_to_bytes(w0, a0, a1);
    nib8_to_bytes(__funnelshift_r(w1, w2, 8), b0, b1);
    nib8_to_bytes(__funnelshift_r(w2, w3, 16), c0, c1);
    nib8_to_bytes(__funnelshift_r(w3, w4, 24), d0, d1);
    uint32_t at = w10_tail(w1 & 0xFFu), bt = w10_tail((w2 >> 8) & 0xFFu), ct = w10_tail((w3 >> 16) & 0xFFu), dt = w10_tail(w4 >> 24);
    uint32_t* o = (uint32_t*)out72;
    o[1] = a0;
    *(uint2*)(o + 2) = make_uint2(a1, at | 0x01010000u);
    o[5] = 0x0101u | (b0 << 16);
    *(uint2*)(o + 6) = make_uint2(__funnelshift_r(b0, b1, 16), (b1 >> 16) | (bt << 16));
    *(uint2*)(o + 10) = make_uint2(c0, c1);
    o[12] = ct | 0x01010000u;
    *(uint2*)(o + 14) = make_uint2(0x0101u | (d0 << 16), __funnelshift_r(d0, d1, 16));
    o[16] = (d1 >> 16) | (dt << 16);
}
// W = 20: 2 rows = 160 bits = 5 words -> 56 output bytes (8-byte aligned); 10 cell words written.
__device__ __forceinline__ void fill_rows2_w20(const uint32_t* ids5, uint8_t* out56) {
    uint32_t w0 = ids5[0], w1 = ids5[1], w2 = ids5[2], w3 = ids5[3], w4 = ids5[4];
    uint32_t a0, a1, a2, a3, a4, ax, b0, b1, b2, b3, b4, bx;
    nib8_to_bytes(w0, a0, a1);
    nib8_to_bytes(w1, a2, a3);
    nib8_to_bytes(w2 & 0xFFFFu, a4, ax);
    nib8_to_bytes(__funnelshift_r(w2, w3, 16), b0, b1);
    nib8_to_bytes(__funnelshift_r(w3, w4, 16), b2, b3);
    nib8_to_bytes(w4 >> 16, b4, bx);
    (void)ax; (void)bx;
    uint32_t* o = (uint32_t*)out56;
    o[1] = a0;
    *(uint2*)(o + 2) = make_uint2(a1, a2);
    *(uint2*)(o + 4) = make_uint2(a3, a4);
    *(uint2*)(o + 8) = make_uint2(b0, b1);
    *(uint2*)(o + 10) = make_uint2(b2, b3);
    o[12] = b4;
}

// ---- per-env logic of one tile slot (shared by both step kernels) -----------------------------------
struct TileStats { double ep, ret, len, lines; };

// warp reduction + atomic accumulation of the episode statistics, once per warp at the end of the kernel: out of line (320
// instructions that have no business in the step kernel's instruction-cache footprint)
__device__ __noinline__ void flush_stats(double* stats, double ep, double ret, double len, double lines) {
    for (int o = 16; o > 0; o >>= 1) {
        ep += __shfl_xor_sync(0xffffffffu, ep, o);
        ret += __shfl_xor_sync(0xffffffffu, ret, o);
        len += __shfl_xor_sync(0xffffffffu, len, o);
        lines += __shfl_xor_sync(0xffffffffu, lines, o);
    }
    if ((threadIdx.x & 31) == 0 && ep > 0) {
        atomicAdd(stats + 0, ep); atomicAdd(stats + 1, ret);
        atomicAdd(stats + 2, len); atomicAdd(stats + 3, lines);
    }
}

// Runs reset / step / grouped placement for env `e` whose records sit at slot `slot` of the staged tile.
// Returns bit0 = board record dirty, bit1 = rng record dirty.  Writes the 5-tuple scalars and s_box[slot].
// MODE >= 0: the kernel instantiation's fixed mode (0 step, 1 reset, 2 grouped step); -1: p.mode at run time (k_step)
// `eo`: index of the env's 5-tuple outputs (== e except in the multi-step kernel, whose outputs are [step][env])
template <class COLT, bool INFO = true, int MODE = -1, bool XT = true>
__device__ __forceinline__ uint32_t logic_one_env(const StepParams& p, const Tabs& tb, int64_t e, int slot, int action,
                                                  uint32_t* s_hot, uint8_t* s_brd, uint8_t* s_rng, uint32_t* s_box,
                                                  TileStats& st, int64_t eo) {
    const DevCfg& cfg = p.cfg;
    const int BS = cfg.board_stride, RS = cfg.rng_stride;
    StepResult res;
    res.dirty = 0; res.reward = 0; res.lines = 0; res.terminated = 0;
    Hot h;
    uint32_t* rec = (uint32_t*)(s_brd + slot * BS);
    hot_load<XT>(h, s_hot + slot * 8);
    Rng g;
    g.rec = (uint32_t*)(s_rng + slot * RS);
    g.seq = p.seq ? p.seq + e * cfg.seq_len : nullptr;
    g.gid = cfg.env_id_offset + (uint64_t)e;
    g.dirty = false;
    bool need_reset = false, run_ok = true, have_B = false;
    const int mode = MODE >= 0 ? MODE : p.mode;
    if (mode == 1) {
        need_reset = (!p.reset_mask || p.reset_mask[e]);
        if (need_reset && p.seeds && cfg.rng_mode == 0) { ((uint64_t*)g.rec)[0] = p.seeds[e]; g.rec[2] = 0; g.dirty = true; }
    } else if (cfg.autoreset == 1 && h.pending) {
        need_reset = true;  // gymnasium NEXT_STEP autoreset: the action is ignored, the env is reset
    } else {
        // one env_step call site for all modes (it is the bulk of the kernel's code: instruction cache)
        int act = action;
        bool run = true, invalid = false;
        if (mode == 2) {
            // GroupedActionsObservations.step (wrappers/grouped.py:209-269)
            bool ok = (unsigned)action < (unsigned)cfg.A && p.legal[e * cfg.A + action] != 0;
            p.fill_high[e] = (uint8_t)(!ok && cfg.terminate_on_illegal);
            if (ok) {
                h.x = (action >> 2) + P - tb.n[h.p] / 2;   // y untouched (wrappers/grouped.py:244-254)
                h.r = (h.r + (action & 3)) & 3;
                act = cfg.act_hard;
            } else if (cfg.terminate_on_illegal) {
                res.reward = cfg.r_invalid; res.terminated = 1;   // env untouched, episode ends
                run = false; run_ok = false;
            } else {
                act = cfg.act_noop;
                invalid = true;
            }
        }
        if (run) {
            env_step<COLT, XT>(cfg, tb, h, rec, g, act, res);
            have_B = true;
            if (invalid) res.reward = cfg.r_invalid;
        }
        h.ep_ret += (float)res.reward; h.ep_len += 1; h.ep_lines += res.lines;
        if (res.terminated) {
            st.ep += 1; st.ret += h.ep_ret; st.len += h.ep_len; st.lines += h.ep_lines;
            h.ep_ret = 0; h.ep_len = 0; h.ep_lines = 0;
            if (cfg.autoreset == 1) h.pending = 1;
            else if (cfg.autoreset == 2) need_reset = true;
        }
    }
    // (inline: an out-of-line reset forces the hot record into local memory and cost 6 % on the 4 M-env step)
    if (need_reset) { env_reset<COLT, XT>(cfg, h, rec, g); res.dirty = 1; }
    if (mode == 2 && need_reset) p.fill_high[e] = 0;
    hot_store<XT>(h, s_hot + slot * 8);
    if (mode != 1) {
        p.reward[eo] = (float)res.reward;
        p.terminated[eo] = (uint8_t)res.terminated;
        p.truncated[eo] = 0;
        p.lines[eo] = res.lines;
    }
    // the active piece is drawn unless it collides where it stands: env_step leaves the collision mask of the piece that is
    // active after the step (one bmask = 4 shared loads + ~30 instructions per env less on the logic warps)
    const COLT Bact = (have_B && !need_reset) ? (COLT)res.Bfin : bmask<COLT>((const COLT*)rec, cfg.W, tb.cells[h.p * 4 + h.r], h.x);
    const uint32_t show = !((Bact >> h.y) & 1);
    s_box[slot] = (uint32_t)h.x | ((uint32_t)h.y << 8) | ((uint32_t)tb.n[h.p] << 16) | (show << 20) |
                  ((uint32_t)h.p << 24) | ((uint32_t)h.r << 28);
    if (INFO && mode == 2 && p.info_board) {
        // info["board"]: FeatureVectorObservation of the real observation (wrappers/grouped.py:260-264)
        uint8_t f[32];
        int ln;
        placement_features<COLT>(cfg, (const COLT*)rec, tb.cells[h.p * 4 + h.r], h.x, h.y, show != 0, false, COLT(3), f, ln);
        for (int i = 0; i < cfg.F; i++) p.info_board[e * cfg.F + i] = f[i];
    }
    // bit 2 (grouped mode): an illegal action ended the episode -- the observation is filled with `high` (= p.fill_high[e])
    return (uint32_t)res.dirty | ((uint32_t)g.dirty << 1) | ((mode == 2 && !need_reset && res.terminated && !run_ok) ? 4u : 0u);
}

// ---- observation images of one tile (called by `nt` cooperating threads, `t` = index among them) ------
__device__ __forceinline__ void mask_clear_boxes(const uint32_t* boxes, int n_env, uint8_t* i_mask, int OB, int Wp, int t, int nt) {
    for (int it = t; it < n_env * 4; it += nt) {
        int e = it >> 2, i = it & 3;
        uint32_t bx = boxes[e];
        int n = (bx >> 16) & 15;
        if (i < n) {
            int addr = e * OB + (((bx >> 8) & 255) + i) * Wp + (bx & 255), a = addr & 3;
            *(uint32_t*)(i_mask + addr - a) = 0;
            if (a + n > 4) *(uint32_t*)(i_mask + addr - a + 4) = 0;
        }
    }
}
template <int WT, int HT, bool XT = true>
__device__ __forceinline__ void fill_images(const DevCfg& cfg, int nv, const uint32_t* s_hot, const uint8_t* s_brd,
                                            const uint32_t* s_rowbytes, uint8_t* i_board, uint8_t* i_holder, uint8_t* i_queue,
                                            int t, int nt) {
    const int W = WT ? WT : cfg.W, H = HT ? HT : cfg.H;
    const int OB = (H + P) * (W + 2 * P), BS = cfg.board_stride, Q = cfg.Q;
    if (WT == 10 && (HT % 4) == 0) {
        constexpr int G = HT ? HT / 4 : 1;
        for (int it = t; it < nv * G; it += nt) {
            int e = it / G, g4 = it - e * G;
            fill_rows4_w10((const uint32_t*)(s_brd + e * BS + 40) + 5 * g4, i_board + e * OB + g4 * 72);
        }
    } else if (WT == 20 && (HT % 2) == 0) {
        constexpr int G = HT ? HT / 2 : 1;
        for (int it = t; it < nv * G; it += nt) {
            int e = it / G, g2 = it - e * G;
            fill_rows2_w20((const uint32_t*)(s_brd + e * BS + cfg.ids_off) + 5 * g2, i_board + e * OB + g2 * 56);
        }
    } else {
        for (int it = t; it < nv * H; it += nt) {
            int e = it / H, row = it - e * H;
            fill_board_row<WT>(cfg, (const uint32_t*)(s_brd + e * BS + cfg.ids_off), i_board, e * OB, row);
        }
    }
    const uint32_t invQ = cfg.inv_q;                     // it / Q for it < 4096 without a division
    for (int it = t; it < nv * Q; it += nt) {          // queue: one piece per item, its 4 matrix rows in one 128-bit read
        int e = (int)(((uint32_t)it * invQ) >> 16), q = it - e * Q;
        uint64_t queue = (uint64_t)s_hot[e * 8 + 2] | ((uint64_t)s_hot[e * 8 + 3] << 32);
        uint4 rb = *(const uint4*)(s_rowbytes + ((int)((queue >> (4 * q)) & 15u)) * 16);
        uint32_t* qo = (uint32_t*)i_queue + e * 4 * Q + q;
        qo[0] = rb.x; qo[Q] = rb.y; qo[2 * Q] = rb.z; qo[3 * Q] = rb.w;
    }
    if (!XT || cfg.holder_size <= 1) {
        for (int it = t; it < nv * 4; it += nt) {
            int e = it >> 2, i = it & 3;
            uint32_t w0 = s_hot[e * 8];
            int hold = (w0 >> 18) & 15, hr = (w0 >> 22) & 3;
            ((uint32_t*)i_holder)[it] = hold ? s_rowbytes[((hold - 1) * 4 + hr) * 4 + i] : 0x01010101u;
        }
    } else {   // holder_size > 1: the held pieces side by side, oldest first; ones in the empty slots
        const int S = cfg.holder_size;
        for (int it = t; it < nv * 4; it += nt) {
            int e = it >> 2, i = it & 3;
            const uint32_t hq = s_hot[e * 8 + 7];
            const int cnt = (int)(hq & 7u);
            for (int s = 0; s < S; s++) {
                const uint32_t sl = (hq >> (3 + 5 * s)) & 31u;
                ((uint32_t*)i_holder)[it * S + s] = s < cnt ? s_rowbytes[((int)(sl & 7u) * 4 + (int)(sl >> 3)) * 4 + i] : 0x01010101u;
            }
        }
    }
}
// active piece overlay + bounding-box mask (Tetris._get_obs, envs/tetris.py:566-576)
__device__ __forceinline__ void mask_set_and_overlay(const uint32_t* boxes, int nv, const unsigned short* s_cells, uint8_t* i_board,
                                                     uint8_t* i_mask, int OB, int Wp, int t, int nt) {
    for (int it = t; it < nv * 4; it += nt) {
        int e = it >> 2, i = it & 3;
        uint32_t bx = boxes[e];
        int n = (bx >> 16) & 15, x = bx & 255, y = (bx >> 8) & 255;
        if (i < n) {
            int addr = e * OB + (y + i) * Wp + x, a = addr & 3;
            uint64_t v = (uint64_t)(0x01010101u >> (8 * (4 - n))) << (8 * a);
            *(uint32_t*)(i_mask + addr - a) = (uint32_t)v;
            if (a + n > 4) *(uint32_t*)(i_mask + addr - a + 4) = (uint32_t)(v >> 32);
        }
        if ((bx >> 20) & 1) {
            int pc = (bx >> 24) & 7, c = (s_cells[pc * 4 + (bx >> 28)] >> (4 * i)) & 15;
            i_board[e * OB + (y + (c >> 2)) * Wp + x + (c & 3)] = (uint8_t)(pc + 2);
        }
    }
}
// once per CTA: piece tables to shared memory; constant parts of the images (bedrock frame, empty mask)
__device__ __forceinline__ void init_cta(int E, int W, int H, uint32_t* s_rowbytes, unsigned short* s_cells, int* s_n,
                                         uint8_t* i_board, uint8_t* i_mask, int tid, int T) {
    const int Wp = W + 2 * P, OB = (H + P) * Wp;
    for (int i = tid; i < 112; i += T) s_rowbytes[i] = (&c_rowbytes[0][0][0])[i];
    for (int i = tid; i < 28; i += T) s_cells[i] = (&c_cells[0][0])[i];
    for (int i = tid; i < 7; i += T) s_n[i] = c_n[i];
    for (int i = tid; i < (E ? OB : 0); i += T) {  // env 0's template ...
        int r = i / Wp, c = i - r * Wp;
        i_board[i] = (r < H && c >= P && c < P + W) ? 0 : 1;
    }
    for (int i = tid; i < (E * OB + 3) / 4; i += T) ((uint32_t*)i_mask)[i] = 0;
    __syncthreads();
    if ((OB & 3) == 0) {
        for (int e = 1; e < E; e++)      // ... replicated to the other env slots
            for (int i = tid; i < OB / 4; i += T) ((uint32_t*)(i_board + e * OB))[i] = ((const uint32_t*)i_board)[i];
    } else {
        for (int e = 1; e < E; e++)
            for (int i = tid; i < OB; i += T) i_board[e * OB + i] = i_board[i];
    }
}

template <int WT, int HT, class COLT>
__global__ void __launch_bounds__(256) k_step(const __grid_constant__ StepParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const DevCfg& cfg = p.cfg;
    const int E = p.E, T = blockDim.x, tid = threadIdx.x;
    const int W = WT ? WT : cfg.W, H = HT ? HT : cfg.H;
    const int Wp = W + 2 * P, Hp = H + P;
    const int OB = Hp * Wp, OQ = cfg.OQ, BS = cfg.board_stride, RS = cfg.rng_stride;

    uint8_t* i_board = smem + p.off_iboard;
    uint8_t* i_mask = smem + p.off_imask;
    uint8_t* i_holder = smem + p.off_iholder;
    uint8_t* i_queue = smem + p.off_iqueue;
    uint64_t* bar = (uint64_t*)(smem + p.off_bar);     // bar[0], bar[1]: one per state stage
    uint32_t* s_box2 = (uint32_t*)(smem + p.off_box);  // [2][E]: x | y<<8 | n<<16 | show<<20 | piece<<24 | rot<<28
    uint32_t* s_rowbytes = (uint32_t*)(smem + p.off_tab);            // 112 words
    unsigned short* s_cells = (unsigned short*)(s_rowbytes + 112);   // 28 halves
    int* s_n = (int*)(s_rowbytes + 112 + 16);                        // 7 ints
    Tabs tb;
    tb.cells = s_cells; tb.rowbytes = s_rowbytes; tb.n = s_n;

    for (int i = tid; i < 2 * E; i += T) s_box2[i] = 0;
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    init_cta(E, W, H, s_rowbytes, s_cells, s_n, i_board, i_mask, tid, T);

    const int64_t ntiles = (p.n + E - 1) / E;
    TileStats st = {0, 0, 0, 0};
    const bool want_obs = p.o_board != nullptr;
    int nv_prev = 0;

    auto issue_load = [&](int64_t tile, int b) {   // TMA bulk loads of one tile's state into stage b (one thread)
        const int64_t base = tile * E;
        const int nv = (int)min((int64_t)E, p.n - base);
        mbar_expect_tx(bar + b, (uint32_t)(nv * (32 + BS + RS)));
        bulk_g2s(smem + p.off_hot + b * p.st_hot, p.hot + base * 32, (uint32_t)(nv * 32), bar + b);
        bulk_g2s(smem + p.off_brd + b * p.st_brd, p.board + base * BS, (uint32_t)(nv * BS), bar + b);
        bulk_g2s(smem + p.off_rng + b * p.st_rng, p.rng + base * RS, (uint32_t)(nv * RS), bar + b);
    };
    if (tid == 0 && (int64_t)blockIdx.x < ntiles) issue_load(blockIdx.x, 0);
    __syncthreads();

    int k = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, k++) {
        const int b = k & 1;
        const int64_t base = tile * E;
        const int nv = (int)min((int64_t)E, p.n - base);
        uint32_t* s_hot = (uint32_t*)(smem + p.off_hot + b * p.st_hot);
        uint8_t* s_brd = smem + p.off_brd + b * p.st_brd;
        uint8_t* s_rng = smem + p.off_rng + b * p.st_rng;
        uint32_t* s_box = s_box2 + b * E;
        const uint32_t* s_box_prev = s_box2 + (b ^ 1) * E;

        // (C) game logic, one thread per env, on the state that was prefetched into stage b
        uint32_t dirty = 0;
        if (tid < nv) {
            const int64_t e = base + tid;
            int action = 0;
            if (p.mode != 1) action = p.actions[e];
            mbar_wait(bar + b, (uint32_t)((k >> 1) & 1));
            dirty = logic_one_env<COLT>(p, tb, e, tid, action, s_hot, s_brd, s_rng, s_box, st, e);
        }
        // the previous tile's stores must have finished reading shared memory (images + the other state stage)
        bulk_wait_read();
        __syncthreads();
        if (tid >= nv) mbar_wait(bar + b, (uint32_t)((k >> 1) & 1));  // phase already complete: acquire the TMA writes
        // prefetch the next tile's state into the other stage while this tile's images are produced
        if (tid == 0 && tile + gridDim.x < ntiles) issue_load(tile + gridDim.x, b ^ 1);

        if (want_obs) {
            // (D) erase last tile's bounding boxes; board rows, queue, holder images (all threads)
            mask_clear_boxes(s_box_prev, nv_prev, i_mask, OB, Wp, tid, T);
            fill_images<WT, HT>(cfg, nv, s_hot, s_brd, s_rowbytes, i_board, i_holder, i_queue, tid, T);
            __syncthreads();
            // (E) active piece overlay + bounding-box mask
            mask_set_and_overlay(s_box, nv, s_cells, i_board, i_mask, OB, Wp, tid, T);
        }
        // (F) stores
        fence_async_smem();
        __syncthreads();
        const bool leader = (tid == 0);
        if (want_obs) {
            tile_store(p.o_board + base * OB, i_board, (uint32_t)(nv * OB), leader, tid, T);
            tile_store(p.o_mask + base * OB, i_mask, (uint32_t)(nv * OB), leader, tid, T);
            if (leader) {
                bulk_s2g(p.o_holder + base * cfg.OH, i_holder, (uint32_t)(nv * cfg.OH));
                bulk_s2g(p.o_queue + base * OQ, i_queue, (uint32_t)(nv * OQ));
            }
        }
        if (leader) bulk_s2g(p.hot + base * 32, s_hot, (uint32_t)(nv * 32));
        if (tid < nv) {
            if (dirty & 1) bulk_s2g(p.board + (base + tid) * BS, s_brd + tid * BS, (uint32_t)BS);
            if (dirty & 2) bulk_s2g(p.rng + (base + tid) * RS, s_rng + tid * RS, (uint32_t)RS);
        }
        bulk_commit();
        nv_prev = want_obs ? nv : 0;
    }
    bulk_wait_all();
    if (p.stats) {
        for (int o = 16; o > 0; o >>= 1) {
            st.ep += __shfl_xor_sync(0xffffffffu, st.ep, o);
            st.ret += __shfl_xor_sync(0xffffffffu, st.ret, o);
            st.len += __shfl_xor_sync(0xffffffffu, st.len, o);
            st.lines += __shfl_xor_sync(0xffffffffu, st.lines, o);
        }
        if ((tid & 31) == 0 && st.ep > 0) {
            atomicAdd(p.stats + 0, st.ep); atomicAdd(p.stats + 1, st.ret);
            atomicAdd(p.stats + 2, st.len); atomicAdd(p.stats + 3, st.lines);
        }
    }
}

// ---- warp-specialised variant: the logic warps run ahead of the warps that produce and store the observation
// images.  NS state stages (hot + board + rng records of 32 envs each; NS >= NL + 2, NS % NL == 0) are kept in flight by TMA; the roles
// meet on named barriers (ready[stage]) and mbarriers (full[stage]). ----------
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// MODE = 0 step, 1 reset, 2 grouped placement step: one instantiation each, so that the step instantiation carries neither
// the reset-mode code nor the grouped code (legal-mask test, info board, whole-tile write-back) -- its speed depends on the
// code footprint (instruction cache).
// XT: custom tetromino set or holder FIFO (tg_device.cuh); the reference configuration runs the XT = false instantiation.
template <int WT, int HT, class COLT, int MODE, bool XT>
__global__ void __launch_bounds__(256) k_step_ws(const __grid_constant__ StepParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const DevCfg& cfg = p.cfg;
    const int E = p.E;                              // envs per tile (<= 32: one logic-warp lane per env)
    const int NL = p.NL, NS = p.NS;                 // logic warps; state stages in flight (>= NL + 2, a multiple of NL)
    const int T = blockDim.x, tid = threadIdx.x;
    const int FT = T - 32 * NL, ft = tid - 32 * NL; // fill threads
    const int W = WT ? WT : cfg.W, H = HT ? HT : cfg.H;
    const int Wp = W + 2 * P, Hp = H + P;
    const int OB = Hp * Wp, OQ = cfg.OQ, BS = cfg.board_stride, RS = cfg.rng_stride;
    const int OH = XT ? cfg.OH : 16;
    const int BAR_FILL = 1 + NS;                    // named barriers: 1 + s = ready[s], 1 + NS = fill warps only

    uint8_t* i_board = smem + p.off_iboard;
    uint8_t* i_mask = smem + p.off_imask;
    uint8_t* i_holder = smem + p.off_iholder;
    uint8_t* i_queue = smem + p.off_iqueue;
    uint64_t* bar = (uint64_t*)(smem + p.off_bar);     // full[NS]
    uint32_t* s_boxes = (uint32_t*)(smem + p.off_box); // [NS][E] boxes, [NS][E] dirty flags, [E] boxes of the previous tile (fill-private)
    uint32_t* s_flags = s_boxes + NS * E;
    uint32_t* s_boxprev = s_flags + NS * E;
    uint32_t* s_rowbytes = (uint32_t*)(smem + p.off_tab);
    unsigned short* s_cells = (unsigned short*)(s_rowbytes + 112);
    int* s_n = (int*)(s_rowbytes + 112 + 16);
    Tabs tb;
    tb.cells = s_cells; tb.rowbytes = s_rowbytes; tb.n = s_n;

    for (int i = tid; i < (2 * NS + 1) * E; i += T) s_boxes[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < NS; s++) mbar_init(bar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr bool GROUPED = MODE == 2;
    const bool want_obs = p.o_board != nullptr;
    init_cta(want_obs ? E : 0, W, H, s_rowbytes, s_cells, s_n, i_board, i_mask, tid, T);   // no image buffers without the obs dict

    const int64_t ntiles = (p.n + E - 1) / E;
    const int64_t G = gridDim.x;
    auto issue_load = [&](int64_t tile, int s) {
        const int64_t base = tile * E;
        const int nv = (int)min((int64_t)E, p.n - base);
        mbar_expect_tx(bar + s, (uint32_t)(nv * (32 + BS + RS)));
        bulk_g2s(smem + p.off_hot + s * p.st_hot, p.hot + base * 32, (uint32_t)(nv * 32), bar + s);
        bulk_g2s(smem + p.off_brd + s * p.st_brd, p.board + base * BS, (uint32_t)(nv * BS), bar + s);
        bulk_g2s(smem + p.off_rng + s * p.st_rng, p.rng + base * RS, (uint32_t)(nv * RS), bar + s);
    };
    // Programmatic dependent launch: the next launch on the stream (usually the next step) may start its prologue (tables,
    // image templates, barriers -- everything above) while this grid drains; nothing of the state or of the caller's buffers
    // is touched before the preceding grid has completed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (ft == 0) {   // the fill leader owns all TMA traffic: tiles 0 .. NS-2 of this CTA go to stages 0 .. NS-2
        for (int j = 0; j < NS - 1; j++)
            if ((int64_t)blockIdx.x + j * G < ntiles) issue_load((int64_t)blockIdx.x + j * G, j);
    }
    __syncthreads();

    if (tid < 32 * NL) {
        // ===== logic warps: warp lw runs the game logic of this CTA's tiles k = lw, lw + NL, ... (lane = env) =====
        const int lw = tid >> 5, lane = tid & 31;
        TileStats st = {0, 0, 0, 0};
        // stage s = k % NS and mbarrier phase (k / NS) & 1 of this warp's tiles k = lw, lw + NL, ...: kept incrementally (NL <= NS;
        // with a 64-bit k both were calls of the 64-bit division routine, once per tile)
        int s = lw;
        uint32_t ph = 0;
        for (int64_t k = lw; blockIdx.x + k * G < ntiles; k += NL, s += NL) {
            if (s >= NS) { s -= NS; ph ^= 1u; }
            const int64_t tile = blockIdx.x + k * G;
            const int64_t base = tile * E;
            const int nv = (int)min((int64_t)E, p.n - base);
            int action = 0;
            if (MODE != 1 && lane < nv) action = p.actions[base + lane];
            mbar_wait(bar + s, ph);
            uint32_t dirty = 0;
            if (lane < nv)
                dirty = logic_one_env<COLT, false, MODE, XT>(p, tb, base + lane, lane, action, (uint32_t*)(smem + p.off_hot + s * p.st_hot),
                                                   smem + p.off_brd + s * p.st_brd, smem + p.off_rng + s * p.st_rng, s_boxes + s * E, st, base + lane);
            // grouped mode: tiles where most envs committed (nearly always) write their board records back as ONE bulk copy
            const int ndirty = GROUPED ? __popc(__ballot_sync(0xffffffffu, (dirty & 1u) != 0)) : 0;
            if (lane < E) s_flags[s * E + lane] = dirty | ((uint32_t)ndirty << 8);
            // the records this warp just wrote leave through bulk (async-proxy) stores issued by the fill warps: the WRITER orders its
            // generic-proxy writes before them (without it two grouped envs on two streams corrupted board records: the dict-less
            // fill warps issue the whole-tile store right behind the barrier)
            fence_async_smem();
            __syncwarp();
            named_arrive(1 + s, 32 + FT);   // ready[s]: the fill warps may consume stage s
        }
        if (p.stats) flush_stats(p.stats, st.ep, st.ret, st.len, st.lines);
    } else {
        // ===== image / store warps =====
        const bool leader = (ft == 0);
        int nv_prev = 0, s = 0;
        uint32_t ph = 0;                                // mbarrier phase (k / NS) & 1 of tile k
        for (int64_t k = 0; blockIdx.x + k * G < ntiles; k++, ph ^= (s + 1 == NS ? 1u : 0u), s = (s + 1 == NS ? 0 : s + 1)) {
            const int64_t tile = blockIdx.x + k * G;
            const int64_t base = tile * E;
            const int nv = (int)min((int64_t)E, p.n - base);
            uint32_t* s_hot = (uint32_t*)(smem + p.off_hot + s * p.st_hot);
            uint8_t* s_brd = smem + p.off_brd + s * p.st_brd;
            uint8_t* s_rng = smem + p.off_rng + s * p.st_rng;
            named_sync(1 + s, 32 + FT);                 // logic of this tile is done, stage s is final
            mbar_wait(bar + s, ph);                     // (already complete) acquire the TMA writes
            bulk_wait_read();                           // stores of the previous tile have left shared memory
            named_sync(BAR_FILL, FT);
            // the stage of tile k-1 is free again: prefetch NS-1 tiles ahead into it
            if (leader && tile + (NS - 1) * G < ntiles) issue_load(tile + (NS - 1) * G, s == 0 ? NS - 1 : s - 1);
            if (want_obs) {
                mask_clear_boxes(s_boxprev, nv_prev, i_mask, OB, Wp, ft, FT);
                fill_images<WT, HT, XT>(cfg, nv, s_hot, s_brd, s_rowbytes, i_board, i_holder, i_queue, ft, FT);
                named_sync(BAR_FILL, FT);
                mask_set_and_overlay(s_boxes + s * E, nv, s_cells, i_board, i_mask, OB, Wp, ft, FT);
                for (int i = ft; i < nv; i += FT) s_boxprev[i] = s_boxes[s * E + i];
            }
            if (GROUPED && p.info_board) {
                // info["board"] = FeatureVectorObservation of the real observation (wrappers/grouped.py:260-264): rows 0-1 zeroed,
                // active piece projected when it does not collide.  One thread per (env, column), then one thread per env.
                uint8_t* f_h = smem + p.off_feat;          // [E][32] heights
                uint8_t* f_o = f_h + E * 32;               // [E][32] holes
                for (int it = ft; it < nv * W; it += FT) {
                    const int i = it / W, c = it - i * W;
                    const uint32_t bx = s_boxes[s * E + i];
                    COLT v = ((const COLT*)(s_brd + i * BS))[c];
                    if ((bx >> 20) & 1) {
                        const uint32_t cells = s_cells[((bx >> 24) & 7) * 4 + (bx >> 28)];
                        const int x = bx & 255, y = (bx >> 8) & 255;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int cc = (cells >> (4 * k)) & 15;
                            if (x + (cc & 3) - P == c) v |= COLT(1) << (y + (cc >> 2));
                        }
                    }
                    int hgt, hol;
                    col_features<COLT>(v & ~COLT(3), H, hgt, hol);
                    f_h[i * 32 + c] = (uint8_t)hgt; f_o[i * 32 + c] = (uint8_t)hol;
                }
                named_sync(BAR_FILL, FT);
                for (int i = ft; i < nv; i += FT) {
                    int maxh = 0, holes = 0, bump = 0, prev = 0;
                    uint8_t* o = p.info_board + (base + i) * cfg.F;
                    for (int c = 0; c < W; c++) {
                        const int hgt = f_h[i * 32 + c];
                        o[c] = (uint8_t)hgt;
                        holes += f_o[i * 32 + c];
                        maxh = max(maxh, hgt);
                        if (c > 0) bump += abs(hgt - prev);
                        prev = hgt;
                    }
                    o[W] = (uint8_t)maxh; o[W + 1] = (uint8_t)holes; o[W + 2] = (uint8_t)bump;   // uint8 wrap
                }
            }
            fence_async_smem();
            named_sync(BAR_FILL, FT);
            if (want_obs) {
                tile_store<true>(p.o_board + base * OB, i_board, (uint32_t)(nv * OB), leader, ft, FT);
                tile_store<true>(p.o_mask + base * OB, i_mask, (uint32_t)(nv * OB), leader, ft, FT);
                if (leader) {
                    bulk_s2g_stream(p.o_holder + base * OH, i_holder, (uint32_t)(nv * OH));
                    bulk_s2g_stream(p.o_queue + base * OQ, i_queue, (uint32_t)(nv * OQ));
                }
            }
            const bool whole = GROUPED && (int)(s_flags[s * E] >> 8) >= p.whole_tile_min;   // unchanged records are rewritten with the same bytes
            if (leader) bulk_s2g(p.hot + base * 32, s_hot, (uint32_t)(nv * 32));
            if (leader && whole) bulk_s2g(p.board + base * BS, s_brd, (uint32_t)(nv * BS));
            for (int i = ft; i < nv; i += FT) {
                uint32_t d = s_flags[s * E + i];
                if ((d & 1) && !whole) bulk_s2g(p.board + (base + i) * BS, s_brd + i * BS, (uint32_t)BS);
                if (d & 2) bulk_s2g(p.rng + (base + i) * RS, s_rng + i * RS, (uint32_t)RS);
            }
            bulk_commit();
            nv_prev = want_obs ? nv : 0;
        }
        bulk_wait_all();
    }
}

}  // namespace tg
