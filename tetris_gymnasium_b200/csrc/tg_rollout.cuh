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
#include "tg_gfeats.cuh"

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

// out-of-line reset: keeps the rollout loop's code footprint (instruction cache) small, resets are rare
template <class COLT>
__device__ __noinline__ void env_reset_ni(const DevCfg& cfg, Hot& h, uint32_t* rec, Rng& g) { env_reset<COLT>(cfg, h, rec, g); }

// ---- packed-byte variant for W = 10 / W = 20 (same arithmetic as k_grouped_feats_x, tg_gfeats.cuh) ----------------------
// The thread keeps the column heights of its env as packed bytes (registers + a copy in its shared-memory slot for the
// dynamic 4-byte window at byte x), evaluates the four rotations of every column branch-free (bytewise max under the piece's
// column mask, IDP.4A for the height / hole sums, VABSDIFF4 for the bumpiness) and defers the rare placements that clear
// rows or touch the zeroed row 0 to an exact per-column pass.  Same slot layout as k_rollout (the packed heights take the
// place of its h / ho / bs arrays).
template <int W, class COLT>
__global__ void __launch_bounds__(128, 6) k_rollout_x(const __grid_constant__ RolloutParams p) {
    constexpr int WP = W + 2 * P, HW = (WP + 3) / 4, NH = (W + 3) / 4, VL = W - 4 * (NH - 1), A = 4 * W;
    extern __shared__ __align__(16) uint32_t rsm[];
    const DevCfg& cfg = p.cfg;
    const int tid = threadIdx.x, H = cfg.H;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + tid;
    __shared__ unsigned short s_cells[28];
    __shared__ int s_n[8];
    __shared__ uint4 s_prec[28];
    __shared__ __align__(16) uint32_t s_sel[(W + P) * 8];          // window merge selectors per x (tg_gfeats.cuh)
    if (tid < 28) {
        s_cells[tid] = (&c_cells[0][0])[tid];
        uint4 v = (&c_prec[0][0])[tid];
        v.w = prec_w_for_width<W>(v.w);   // x-legality mask | smallest top offset << 28 | first column << 30 (tg_gfeats.cuh)
        s_prec[tid] = v;
    }
    if (tid < 7) s_n[tid] = c_n[tid];
    build_sel_table<NH>(s_sel, W + P, tid, (int)blockDim.x);
    __syncthreads();
    Tabs tb;
    tb.ptab = nullptr; tb.cells = s_cells; tb.rowbytes = &c_rowbytes[0][0][0]; tb.n = s_n;
    TileStats st = {0, 0, 0, 0};
    if (e < p.n) {
        COLT* colp = (COLT*)(rsm + (size_t)tid * p.rec_words);   // colp[c + P] = column c, walls on both sides
        uint32_t* rec = (uint32_t*)(colp + P);
        const uint32_t* grec = (const uint32_t*)(p.board + e * cfg.board_stride);
        const int ncw = p.ids_off_g / 4, nw = cfg.board_stride / 4;
        uint32_t* ids = rec + cfg.ids_off / 4;
        for (int i = 0; i < ncw; i++) rec[i] = grec[i];
        for (int i = ncw; i < nw; i++) ids[i - ncw] = grec[i];
        for (int c = 0; c < P; c++) { colp[c] = ~COLT(0); colp[P + W + c] = ~COLT(0); }
        Hot h;
        hot_load(h, (const uint32_t*)(p.hot + e * 32));
        Rng g;
        g.rec = (uint32_t*)(p.rng + e * cfg.rng_stride);
        g.seq = p.seq ? p.seq + e * cfg.seq_len : nullptr;
        g.gid = cfg.env_id_offset + (uint64_t)e;
        g.dirty = false;
        const COLT* cols = (const COLT*)rec;
        COLT* pre = (COLT*)(rec + p.base_off);
        COLT* suf = pre + W;
        uint32_t* hv = (uint32_t*)(suf + W);                     // HW words: P pad bytes | W heights | pad
        const COLT field = (COLT(1) << H) - 1;
        int last = -1;
        for (int step = 0; step < p.k_steps; step++) {
            if (cfg.autoreset == 1 && h.pending) { env_reset_ni<COLT>(cfg, h, rec, g); last = -1; continue; }
            // ---- per-step base: prefix / suffix column ANDs, heights with row 0 zeroed (Q1), their sums ----
            uint32_t V[HW];
#pragma unroll
            for (int k = 0; k < HW; k++) V[k] = 0;
            int holes0 = 0, sumh0 = 0;
            {
                COLT acc = ~COLT(0);
#pragma unroll
                for (int c = 0; c < W; c++) {
                    const COLT col = cols[c];
                    pre[c] = acc; acc &= col;
                    int hgt, hol;
                    col_features<COLT>(col & ~COLT(1), H, hgt, hol);
                    V[(P + c) >> 2] |= (uint32_t)hgt << (8 * ((P + c) & 3));
                    holes0 += hol; sumh0 += hgt;
                }
                acc = ~COLT(0);
#pragma unroll
                for (int c = W - 1; c >= 0; c--) { suf[c] = acc; acc &= cols[c]; }
#pragma unroll
                for (int k = 0; k < HW; k++) hv[k] = V[k];
            }
            const uint4* prow = s_prec + h.p * 4;
            const int rot0 = h.r;
            const int xoff = P - (int)((cfg.nhalf3 >> (3 * h.p)) & 7u);               // wrappers/grouped.py:157-158 (n // 2)
            int best = -1, best_score = 0, first_legal = -1;
            unsigned long long slow_lo = 0;   // placements 0..63 / 64.. that need the exact pass
            uint32_t slow_hi = 0;
#pragma unroll 1
            for (int xb = 0; xb < W; xb++) {
                const int x = xb + xoff, wi = x >> 2, sh = (x & 3) * 8;
                const uint32_t O4 = __funnelshift_r(hv[wi], hv[wi + 1], sh);
                uint32_t sel[8];
                {
                    const uint4 sa = ((const uint4*)s_sel)[2 * x];
                    sel[0] = sa.x; sel[1] = sa.y; sel[2] = sa.z; sel[3] = sa.w;
                    if (NH > 4) { const uint4 sb = ((const uint4*)s_sel)[2 * x + 1]; sel[4] = sb.x; sel[5] = sb.y; sel[6] = sb.z; sel[7] = sb.w; }
                }
                COLT cj[4];
#pragma unroll
                for (int j = 0; j < 4; j++) cj[j] = colp[x + j];
                // columns left / right of the 4-column window: shared by the four rotations (see tg_gfeats.cuh)
                const COLT LR = pre[max(x - P, 0)] & suf[min(x - P + 3, W - 1)] & field;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int a = 4 * xb + r;
                    const uint4 q = prow[(rot0 + r) & 3];            // cumulative rot90 presses (wrappers/grouped.py:153-154)
                    COLT B = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int c = (q.x >> (4 * k)) & 15;
                        B |= colp[x + (c & 3)] >> (c >> 2);
                    }
                    const int y = ctz_t<COLT>(B >> 1);               // while !collision(y+1): y++ from y = 0
                    const int mintop = (q.w >> 28) & 3;
                    const bool legal = ((q.w >> x) & 1u) != 0;       // the piece stays inside the field (collision_with_frame)
                    if (legal && first_legal < 0) first_legal = a;
                    const bool lands = legal && !((B >> y) & 1);
                    COLT full = LR;
#pragma unroll
                    for (int j = 0; j < 4; j++) full &= cj[j] | ((COLT)((q.x >> (16 + 4 * j)) & 15u) << y);
                    const bool exact = lands && (full != 0 || y + mintop == 0);
                    if (exact) { if (a < 64) slow_lo |= 1ull << a; else slow_hi |= 1u << (a - 64); }
                    const uint32_t M4 = q.y;
                    // new window: columns under piece cells rise to H - y - top offset, the others keep their height (tg_gfeats.cuh)
                    const uint32_t T4 = ((uint32_t)(H - y) * 0x01010101u - q.z) & M4;
                    const uint32_t N4 = bytemax_lt128(O4, T4);
                    const int delta = __dp4a((int)O4, (int)0xFFFFFFFFu, __dp4a((int)N4, 0x01010101, 0));   // sum(new) - sum(old)
                    uint32_t hw[NH];
#pragma unroll
                    for (int k = 0; k < NH; k++) hw[k] = prmt_raw(V[k + 1], N4, sel[k]);
                    uint32_t bump = 0;
#pragma unroll
                    for (int k = 0; k < NH - 1; k++) bump = __vsadu4(hw[k], __funnelshift_r(hw[k], hw[k + 1], 8)) + bump;
                    {
                        const uint32_t l = hw[NH - 1];
                        const uint32_t aa = VL == 4 ? l : __byte_perm(l, 0, VL == 1 ? 0x4444 : (VL == 2 ? 0x4410 : 0x4210));
                        const uint32_t bb = __byte_perm(l, 0, VL == 1 ? 0x4444 : (VL == 2 ? 0x4411 : (VL == 3 ? 0x4221 : 0x3321)));
                        bump = __vsadu4(aa, bb) + bump;
                    }
                    // score on the uint8 feature values (holes / bumpiness wrap); no row is cleared on this path
                    const int score = p.w[0] * (sumh0 + delta) + p.w[2] * ((holes0 - 4 + delta) & 255) + p.w[3] * (int)(bump & 255u);
                    const bool take = lands && !exact && (best < 0 || score > best_score);
                    best = take ? a : best;
                    best_score = take ? score : best_score;
                }
            }
            // ---- exact pass: placements that clear rows / put a cell into the zeroed row 0 (per column, see tg_gfeats.cuh) ----
            while (slow_lo | slow_hi) {
                int a;
                if (slow_lo) { a = __ffsll((long long)slow_lo) - 1; slow_lo &= slow_lo - 1; }
                else { a = 64 + __ffs((int)slow_hi) - 1; slow_hi &= slow_hi - 1; }
                const uint4 q = s_prec[h.p * 4 + ((h.r + (a & 3)) & 3)];
                const int x = (a >> 2) + xoff;
                COLT B = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int c = (q.x >> (4 * k)) & 15;
                    B |= colp[x + (c & 3)] >> (c >> 2);
                }
                const int y = ctz_t<COLT>(B >> 1);
                const int jmin = q.w >> 30, c0 = x + jmin - P, c1 = x + ((31 - __clz((int)q.y)) >> 3) - P;
                COLT full = pre[c0] & suf[c1] & field;
#pragma unroll
                for (int j = 0; j < 4; j++) full &= colp[x + j] | ((COLT)((q.x >> (16 + 4 * j)) & 15u) << y);
                const COLT keep = (full != 0 ? ~full : ~COLT(1)) & field;
                int s_sum = 0, s_hol = 0, s_bmp = 0, prev = 0;
#pragma unroll 1
                for (int c = 0; c < W; c++) {
                    COLT v = colp[c + P];
                    const int t = c - c0;
                    if ((unsigned)t <= (unsigned)(c1 - c0)) v |= (COLT)((q.x >> (16 + 4 * (jmin + t))) & 15u) << y;
                    const COLT u = v & keep;
                    int hgt = 0, hol = 0;
                    if (u != 0) {
                        const int tp = ctz_t<COLT>(u);
                        hgt = H - tp - popc_t<COLT>((full >> tp) >> 1);
                        hol = hgt - popc_t<COLT>(u);
                    }
                    s_sum += hgt; s_hol += hol;
                    if (c > 0) s_bmp += abs(hgt - prev);
                    prev = hgt;
                }
                const int score = p.w[0] * s_sum + p.w[1] * popc_t<COLT>(full) + p.w[2] * (s_hol & 255) + p.w[3] * (s_bmp & 255);
                // lowest index among the maxima: a later candidate wins only if strictly better, an earlier one on ties
                if (best < 0 || score > best_score || (score == best_score && a < best)) { best = a; best_score = score; }
            }
            const int action = best >= 0 ? best : first_legal;
            last = action;
            // ---- execute (GroupedActionsObservations.step, wrappers/grouped.py:241-259) ----
            StepResult res;
            h.x = (action >> 2) + xoff;
            h.r = (h.r + (action & 3)) & 3;
            env_step<COLT>(cfg, tb, h, rec, g, cfg.act_hard, res);
            h.ep_ret += (float)res.reward; h.ep_len += 1; h.ep_lines += res.lines;
            if (res.terminated) {
                st.ep += 1; st.ret += h.ep_ret; st.len += h.ep_len; st.lines += h.ep_lines;
                h.ep_ret = 0; h.ep_len = 0; h.ep_lines = 0;
                if (cfg.autoreset == 1) h.pending = 1;
                else if (cfg.autoreset == 2) env_reset_ni<COLT>(cfg, h, rec, g);
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
