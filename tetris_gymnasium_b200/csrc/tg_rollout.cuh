// tg_rollout.cuh -- fused K-step rollout of an integer linear placement policy (BASELINE config 4).
//
// One thread owns one env for the whole launch: its hot record lives in registers, its board record
// (column bitboards + id plane) in shared memory with an odd word stride (bank-conflict free), so the
// K steps touch HBM only once on the way in and once on the way out.  Per step the thread enumerates the
// 4W placements exactly like GroupedActionsObservations.observation (wrappers/grouped.py:124-207; landing
// row, frame/game-over classification, line clear, FeatureVectorObservation quirks Q1/Q3/Q4), scores them
//     score = w0 * sum(heights) + w1 * lines + w2 * holes + w3 * bumpiness      (int32, on the uint8 feature values)
// picks the lowest-index maximum over legal non-game-over placements (lowest legal index if all lose),
// and executes it like GroupedActionsObservations.step (base hard drop).  NEXT_STEP / SAME_STEP / disabled
// autoreset follow the env config; episode statistics are reduced per CTA and added atomically.
#pragma once
#include "tg_device.cuh"

namespace tg {

struct RolloutParams {
    DevCfg cfg;
    int64_t n;
    uint8_t* hot; uint8_t* board; uint8_t* rng; const uint8_t* seq;
    int w[4];
    int k_steps;
    double* stats;
    int rec_words;   // shared-memory words per env record (odd / 8-byte friendly stride)
    int base_off;    // word offset of the per-thread EnvBase arrays (pre[W], suf[W], h[32 B], ho[32 B]) inside the record slot
    int32_t* last_action;   // nullable: action chosen at the last step (tests)
};

template <class COLT>
__global__ void __launch_bounds__(128) k_rollout(const __grid_constant__ RolloutParams p) {
    extern __shared__ __align__(16) uint32_t rsm[];
    const DevCfg& cfg = p.cfg;
    const int tid = threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + tid;
    __shared__ unsigned short s_cells[28];
    __shared__ int s_n[8];
    if (tid < 28) s_cells[tid] = (&c_cells[0][0])[tid];
    if (tid < 7) s_n[tid] = c_n[tid];
    __syncthreads();
    Tabs tb;
    tb.cells = s_cells; tb.rowbytes = &c_rowbytes[0][0][0]; tb.n = s_n;
    TileStats st = {0, 0, 0, 0};
    if (e < p.n) {
        uint32_t* rec = rsm + (size_t)tid * p.rec_words;
        const uint32_t* grec = (const uint32_t*)(p.board + e * cfg.board_stride);
        const int nw = cfg.board_stride / 4;
        for (int i = 0; i < nw; i++) rec[i] = grec[i];
        Hot h;
        hot_load(h, (const uint32_t*)(p.hot + e * 32));
        Rng g;
        g.rec = (uint32_t*)(p.rng + e * cfg.rng_stride);
        g.seq = p.seq ? p.seq + e * cfg.seq_len : nullptr;
        g.gid = cfg.env_id_offset + (uint64_t)e;
        g.dirty = false;
        const COLT* cols = (const COLT*)rec;
        const int A = cfg.A;
        EnvBase<COLT> eb;
        eb.pre = (COLT*)(rec + p.base_off);
        eb.suf = eb.pre + cfg.W;
        eb.h = (uint8_t*)(eb.suf + cfg.W);
        eb.ho = eb.h + 32;
        int last = -1;
        for (int step = 0; step < p.k_steps; step++) {
            if (cfg.autoreset == 1 && h.pending) { env_reset<COLT>(cfg, h, rec, g); last = -1; continue; }
            // ---- enumerate + score ----
            int best = -1, best_score = 0, first_legal = -1;
            env_base_compute<COLT>(cfg, cols, COLT(1), eb);
            uint32_t slow[3] = {0u, 0u, 0u};   // placements that clear rows: evaluated in a second, short loop
            for (int a = 0; a < A; a++) {
                COLT B;
                Placement pl = eval_placement<COLT>(cfg, tb, cols, h.p, h.r, a, B);
                if (pl.kind == 1) continue;
                if (first_legal < 0) first_legal = a;
                if (pl.kind == 2) continue;
                FeatSum fs = placement_eval_fast<COLT>(cfg, cols, eb, tb.cells[h.p * 4 + pl.rot], pl.x, pl.y, COLT(1), nullptr, true);
                if (fs.lines < 0) { slow[a >> 5] |= 1u << (a & 31); continue; }
                int score = p.w[0] * fs.sum_h + p.w[2] * (int)(uint8_t)fs.holes + p.w[3] * (int)(uint8_t)fs.bump;
                if (best < 0 || score > best_score) { best = a; best_score = score; }
            }
#pragma unroll
            for (int wi = 0; wi < 3; wi++) {
                uint32_t m = slow[wi];
                while (m) {
                    int a = wi * 32 + __ffs((int)m) - 1;
                    m &= m - 1;
                    COLT B;
                    Placement pl = eval_placement<COLT>(cfg, tb, cols, h.p, h.r, a, B);
                    FeatSum fs = placement_eval<COLT>(cfg, cols, tb.cells[h.p * 4 + pl.rot], pl.x, pl.y, true, true, COLT(1), nullptr);
                    int score = p.w[0] * fs.sum_h + p.w[1] * fs.lines + p.w[2] * (int)(uint8_t)fs.holes + p.w[3] * (int)(uint8_t)fs.bump;
                    // lowest index among the maxima: a later candidate wins only if strictly better, an earlier one on ties
                    if (best < 0 || score > best_score || (score == best_score && a < best)) { best = a; best_score = score; }
                }
            }
            int action = best >= 0 ? best : first_legal;
            last = action;
            // ---- execute (GroupedActionsObservations.step, wrappers/grouped.py:241-259) ----
            StepResult res;
            h.x = (action >> 2) + P - tb.n[h.p] / 2;
            h.r = (h.r + (action & 3)) & 3;
            env_step<COLT>(cfg, tb, h, rec, g, cfg.act_hard, res);
            h.ep_ret += (float)res.reward; h.ep_len += 1; h.ep_lines += res.lines;
            if (res.terminated) {
                st.ep += 1; st.ret += h.ep_ret; st.len += h.ep_len; st.lines += h.ep_lines;
                h.ep_ret = 0; h.ep_len = 0; h.ep_lines = 0;
                if (cfg.autoreset == 1) h.pending = 1;
                else if (cfg.autoreset == 2) env_reset<COLT>(cfg, h, rec, g);
            }
        }
        hot_store(h, (uint32_t*)(p.hot + e * 32));
        uint32_t* wrec = (uint32_t*)(p.board + e * cfg.board_stride);
        for (int i = 0; i < nw; i++) wrec[i] = rec[i];
        if (p.last_action) p.last_action[e] = last;
    }
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

}  // namespace tg
