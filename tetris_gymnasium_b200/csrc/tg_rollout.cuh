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
    // shared-memory slot of one env (words): [P wall columns][W columns][P wall columns][id plane (+1 word)][pre W][suf W][h][ho][bs]
    // -- the record is stored with its columns framed by all-ones walls, so place_fast reads them without bounds checks
    int rec_words;   // words per slot (odd / 8-byte friendly stride)
    int ids_off_g;   // byte offset of the id plane inside the HBM record (cfg.ids_off holds the in-slot offset (W + P) * sizeof(COLT))
    int base_off;    // word offset (from the first column) of pre[W]; suf[W], h[hb bytes], ho[hb bytes], bs[2 hb bytes] follow
    int hb;          // bytes of the h / ho arrays (W rounded up to 4)
    int32_t* last_action;   // nullable: action chosen at the last step (tests)
};

template <class COLT>
__global__ void __launch_bounds__(128, 6) k_rollout(const __grid_constant__ RolloutParams p) {
    extern __shared__ __align__(16) uint32_t rsm[];
    const DevCfg& cfg = p.cfg;
    const int tid = threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + tid;
    __shared__ unsigned short s_cells[28];
    __shared__ int s_n[8];
    __shared__ uint2 s_ptab[28];
    if (tid < 28) { s_cells[tid] = (&c_cells[0][0])[tid]; s_ptab[tid] = (&c_ptab[0][0])[tid]; }
    if (tid < 7) s_n[tid] = c_n[tid];
    __syncthreads();
    Tabs tb;
    tb.ptab = s_ptab; tb.cells = s_cells; tb.rowbytes = &c_rowbytes[0][0][0]; tb.n = s_n;
    TileStats st = {0, 0, 0, 0};
    if (e < p.n) {
        // cfg.ids_off is the IN-SLOT offset (P wall columns sit between the columns and the id plane); p.ids_off_g the HBM one
        COLT* colp = (COLT*)(rsm + (size_t)tid * p.rec_words);   // colp[c + P] = column c, walls on both sides
        uint32_t* rec = (uint32_t*)(colp + P);
        const uint32_t* grec = (const uint32_t*)(p.board + e * cfg.board_stride);
        const int ncw = p.ids_off_g / 4, nw = cfg.board_stride / 4;     // column words, record words in HBM
        uint32_t* ids = rec + cfg.ids_off / 4;
        for (int i = 0; i < ncw; i++) rec[i] = grec[i];
        for (int i = ncw; i < nw; i++) ids[i - ncw] = grec[i];
        for (int c = 0; c < P; c++) { colp[c] = ~COLT(0); colp[P + cfg.W + c] = ~COLT(0); }
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
        eb.ho = eb.h + p.hb;
        eb.bs = (uint16_t*)(eb.ho + p.hb);
        int last = -1;
        for (int step = 0; step < p.k_steps; step++) {
            if (cfg.autoreset == 1 && h.pending) { env_reset<COLT>(cfg, h, rec, g); last = -1; continue; }
            // ---- enumerate + score ----
            int best = -1, best_score = 0, first_legal = -1;
            env_base_compute<COLT>(cfg, cols, COLT(1), eb);
            uint32_t slow[3] = {0u, 0u, 0u};    // placements with a piece cell in the zeroed row 0 (rare): exact evaluation in a second loop
            const int xoff = P - tb.n[h.p] / 2;   // wrappers/grouped.py:157-158
            for (int a = 0; a < A; a++) {
                const int rot = (h.r + (a & 3)) & 3;   // cumulative rot90 presses (wrappers/grouped.py:153-154)
                FeatSum fs;
                int y;
                const int kind = place_fast<COLT>(cfg, eb, colp, tb.cells[h.p * 4 + rot], tb.ptab[h.p * 4 + rot], (a >> 2) + xoff, fs, y, nullptr, false);
                if (kind == 1) continue;
                if (first_legal < 0) first_legal = a;
                if (kind == 2) continue;
                if (kind >= 3) { slow[a >> 5] |= 1u << (a & 31); continue; }
                int score = p.w[0] * fs.sum_h + p.w[1] * fs.lines + p.w[2] * (int)(uint8_t)fs.holes + p.w[3] * (int)(uint8_t)fs.bump;
                if (best < 0 || score > best_score) { best = a; best_score = score; }
            }
#pragma unroll
            for (int wi = 0; wi < 3; wi++) {
                uint32_t m = slow[wi];
                while (m) {
                    int a = wi * 32 + __ffs((int)m) - 1;
                    m &= m - 1;
                    const int rot = (h.r + (a & 3)) & 3;
                    FeatSum fs;
                    int y;
                    const int kind = place_fast<COLT>(cfg, eb, colp, tb.cells[h.p * 4 + rot], tb.ptab[h.p * 4 + rot], (a >> 2) + xoff, fs, y, nullptr, false);
                    if (kind == 3)   // a piece cell in the zeroed row 0, no clear: exact evaluation
                        fs = placement_eval<COLT>(cfg, cols, tb.cells[h.p * 4 + rot], (a >> 2) + xoff, y, true, true, COLT(1), nullptr);
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
        for (int i = 0; i < ncw; i++) wrec[i] = rec[i];
        for (int i = ncw; i < nw; i++) wrec[i] = ids[i - ncw];
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
