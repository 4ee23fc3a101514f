// tg_wrappers.cuh -- observation wrappers and the grouped (placement) path.
//   FeatureVectorObservation  (wrappers/observation.py:118-278)  -> k_features / placement_features
//   RgbObservation            (wrappers/observation.py:11-74)    -> k_rgb
//   GroupedActionsObservations(wrappers/grouped.py:16-294)       -> k_grouped_feats / k_grouped_boards
//                                                                   (+ k_step mode 2 executes the placement)
// Included at the end of tg_api.cu (uses tg_env, fail, CUDA_TRY, check_state).
#pragma once

namespace tg {

// FeatureVectorObservation applied to the base env's observation (rows 0-1 zeroed, active piece projected)
template <class COLT>
__global__ void k_features(const DevCfg cfg, int64_t n, const uint8_t* hot, const uint8_t* board, uint8_t* feats) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    Hot h;
    hot_load(h, (const uint32_t*)(hot + e * 32));
    const COLT* cols = (const COLT*)(board + e * cfg.board_stride);
    uint32_t cells = c_cells[h.p][h.r];
    COLT B = bmask<COLT>(cols, cfg.W, cells, h.x);
    int lines;
    uint8_t f[32];
    placement_features<COLT>(cfg, cols, cells, h.x, h.y, !((B >> h.y) & 1), false, COLT(3), f, lines);
    for (int i = 0; i < cfg.F; i++) feats[e * cfg.F + i] = f[i];
}

// grouped observation with FeatureVectorObservation: feats u8[n][A][F], legal u8[n][A]
// CTA = EPB envs; one thread per env precomputes the EnvBase, then one thread per (env, placement).
template <class COLT>
__global__ void __launch_bounds__(256) k_grouped_feats(const DevCfg cfg, int64_t n, const uint8_t* hot, const uint8_t* board, uint8_t* feats,
                                                       uint8_t* legal, const uint8_t* fill_high, int EPB) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int W = cfg.W, A = cfg.A, F = cfg.F;
    COLT* s_cols = (COLT*)sm;                      // [EPB][W]
    COLT* s_pre = s_cols + (size_t)EPB * W;        // [EPB][W]
    COLT* s_suf = s_pre + (size_t)EPB * W;         // [EPB][W]
    int* s_sum = (int*)(s_suf + (size_t)EPB * W);  // [EPB][4]: sum_h, holes, bump, max_h
    uint32_t* s_w0 = (uint32_t*)(s_sum + EPB * 4); // [EPB]
    uint8_t* s_h = (uint8_t*)(s_w0 + EPB);         // [EPB][32]
    uint8_t* s_ho = s_h + EPB * 32;                // [EPB][32]
    uint8_t* s_feats = s_ho + EPB * 32;            // [EPB][A][F]   (16-aligned: every block above is a multiple of 16 for EPB % 4 == 0)
    uint8_t* s_legal = s_feats + (size_t)EPB * A * F;
    __shared__ unsigned short s_cells[28];
    __shared__ int s_n[8];
    __shared__ unsigned short s_slow[16 * 96];   // EPB <= 16, A <= 96
    __shared__ int s_nslow;
    if (threadIdx.x < 28) s_cells[threadIdx.x] = (&c_cells[0][0])[threadIdx.x];
    if (threadIdx.x < 7) s_n[threadIdx.x] = c_n[threadIdx.x];
    if (threadIdx.x == 0) s_nslow = 0;
    Tabs tb;
    tb.cells = s_cells; tb.rowbytes = &c_rowbytes[0][0][0]; tb.n = s_n;
    const int64_t base = (int64_t)blockIdx.x * EPB;
    const int nv = (int)min((int64_t)EPB, n - base);
    for (int i = threadIdx.x; i < nv * W; i += blockDim.x) {
        int e = i / W, c = i - e * W;
        s_cols[i] = ((const COLT*)(board + (base + e) * cfg.board_stride))[c];
    }
    for (int i = threadIdx.x; i < nv; i += blockDim.x) s_w0[i] = *(const uint32_t*)(hot + (base + i) * 32);
    __syncthreads();
    if (threadIdx.x < nv) {
        int e = threadIdx.x;
        EnvBase<COLT> eb;
        eb.h = s_h + e * 32; eb.ho = s_ho + e * 32; eb.pre = s_pre + e * W; eb.suf = s_suf + e * W;
        env_base_compute<COLT>(cfg, s_cols + e * W, COLT(1), eb);
        s_sum[e * 4] = eb.sum_h; s_sum[e * 4 + 1] = eb.holes; s_sum[e * 4 + 2] = eb.bump; s_sum[e * 4 + 3] = eb.max_h;
    }
    __syncthreads();
    for (int it = threadIdx.x; it < nv * A; it += blockDim.x) {
        int e = it / A, a = it - e * A;
        uint32_t w0 = s_w0[e];
        int piece = (w0 >> 13) & 7, rot0 = (w0 >> 16) & 3;
        const COLT* cols = s_cols + e * W;
        uint8_t* out = s_feats + (size_t)it * F;
        if (fill_high && fill_high[base + e]) {
            // illegal action + terminate: obs = ones * high (wrappers/grouped.py:221-226); legal mask unchanged
            for (int i = 0; i < F; i++) out[i] = (uint8_t)(cfg.H * cfg.W);
            s_legal[it] = legal[(base + e) * A + a];
            continue;
        }
        COLT B;
        Placement pl = eval_placement<COLT>(cfg, tb, cols, piece, rot0, a, B);
        s_legal[it] = pl.kind != 1;
        if (pl.kind == 1) {          // ones board, row 0 zeroed -> heights H-1
            for (int i = 0; i <= W; i++) out[i] = (uint8_t)(cfg.H - 1);
            out[W + 1] = 0; out[W + 2] = 0;
        } else if (pl.kind == 2) {   // zeros board
            for (int i = 0; i < F; i++) out[i] = 0;
        } else {
            EnvBase<COLT> eb;
            eb.h = s_h + e * 32; eb.ho = s_ho + e * 32; eb.pre = s_pre + e * W; eb.suf = s_suf + e * W;
            eb.sum_h = s_sum[e * 4]; eb.holes = s_sum[e * 4 + 1]; eb.bump = s_sum[e * 4 + 2]; eb.max_h = s_sum[e * 4 + 3];
            FeatSum fs = placement_eval_fast<COLT>(cfg, cols, eb, tb.cells[piece * 4 + pl.rot], pl.x, pl.y, COLT(1), out, true);
            if (fs.lines < 0) s_slow[atomicAdd(&s_nslow, 1)] = (unsigned short)it;   // rows get cleared: batch the full evaluation
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < s_nslow; k += blockDim.x) {
        int it = s_slow[k], e = it / A, a = it - e * A;
        uint32_t w0 = s_w0[e];
        int piece = (w0 >> 13) & 7, rot0 = (w0 >> 16) & 3;
        COLT B;
        Placement pl = eval_placement<COLT>(cfg, tb, s_cols + e * W, piece, rot0, a, B);
        placement_eval<COLT>(cfg, s_cols + e * W, tb.cells[piece * 4 + pl.rot], pl.x, pl.y, true, true, COLT(1), s_feats + (size_t)it * F);
    }
    __syncthreads();
    // coalesced copy-out of the tile (contiguous in global memory)
    {
        size_t bytes = (size_t)nv * A * F;
        uint8_t* g = feats + (size_t)base * A * F;
        if ((bytes & 15) == 0 && (((uintptr_t)g) & 15) == 0) {
            for (size_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) ((uint4*)g)[i] = ((const uint4*)s_feats)[i];
        } else {
            for (size_t i = threadIdx.x; i < bytes; i += blockDim.x) g[i] = s_feats[i];
        }
        size_t lb = (size_t)nv * A;
        uint8_t* gl = legal + (size_t)base * A;
        if ((lb & 3) == 0 && (((uintptr_t)gl) & 3) == 0) {
            for (size_t i = threadIdx.x; i < lb / 4; i += blockDim.x) ((uint32_t*)gl)[i] = ((const uint32_t*)s_legal)[i];
        } else {
            for (size_t i = threadIdx.x; i < lb; i += blockDim.x) gl[i] = s_legal[i];
        }
    }
}

// grouped observation without wrappers: boards u8[n][A][Hp][Wp]; one warp per (env, placement)
template <class COLT>
__global__ void k_grouped_boards(const DevCfg cfg, int64_t n, const uint8_t* hot, const uint8_t* board, uint8_t* boards,
                                 uint8_t* legal, const uint8_t* fill_high) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int W = cfg.W, H = cfg.H, Wp = cfg.Wp, A = cfg.A, OB = cfg.OB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int OBr = (OB + 15) & ~15;
    uint8_t* buf = sm + (size_t)warp * OBr;
    const int64_t items = n * A;
    for (int64_t it = (int64_t)blockIdx.x * nwarps + warp; it < items; it += (int64_t)gridDim.x * nwarps) {
        int64_t e = it / A;
        int a = (int)(it - e * A);
        const uint8_t* rec = board + e * cfg.board_stride;
        const COLT* cols = (const COLT*)rec;
        const uint32_t* ids = (const uint32_t*)(rec + cfg.ids_off);
        uint32_t w0 = *(const uint32_t*)(hot + e * 32);
        int piece = (w0 >> 13) & 7, rot0 = (w0 >> 16) & 3;
        uint8_t* g = boards + (size_t)it * OB;
        int fillv = -1;
        Placement pl;
        COLT B;
        if (fill_high && fill_high[e]) fillv = (uint8_t)(cfg.H * cfg.W);
        else {
            pl = eval_placement<COLT>(cfg, const_tabs(), cols, piece, rot0, a, B);
            if (lane == 0) legal[it] = pl.kind != 1;
            if (pl.kind == 1) fillv = 1;
            else if (pl.kind == 2) fillv = 0;
        }
        __syncwarp();
        if (fillv >= 0) {
            for (int i = lane; i < OB; i += 32) buf[i] = (uint8_t)fillv;
        } else {
            uint32_t cells = c_cells[piece][pl.rot];
            int crow[4], ccol[4];
            COLT full = ~COLT(0);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int c = (cells >> (4 * k)) & 15;
                crow[k] = pl.y + (c >> 2); ccol[k] = pl.x + (c & 3) - P;
            }
            for (int c = 0; c < W; c++) {
                COLT v = cols[c];
#pragma unroll
                for (int k = 0; k < 4; k++) if (ccol[k] == c) v |= COLT(1) << crow[k];
                full &= v;
            }
            full &= (COLT(1) << H) - 1;
            int nclr = popc_t<COLT>(full);
            for (int r = lane; r < cfg.Hp; r += 32) {
                uint8_t* row = buf + r * Wp;
                if (r >= H) { for (int c = 0; c < Wp; c++) row[c] = 1; continue; }
                for (int c = 0; c < P; c++) { row[c] = 1; row[P + W + c] = 1; }
                if (r < nclr) { for (int c = 0; c < W; c++) row[P + c] = 0; continue; }
                int s = r - nclr;  // (r - nclr)-th surviving source row
                COLT f = full;
                while (f) { int fr = ctz_t<COLT>(f); f &= f - 1; if (fr <= s) s++; }
                for (int c = 0; c < W; c++) row[P + c] = (uint8_t)ids_get1(ids, s * W + c);
#pragma unroll
                for (int k = 0; k < 4; k++) if (crow[k] == s) row[P + ccol[k]] = (uint8_t)(piece + 2);
            }
        }
        __syncwarp();
        if ((OB & 15) == 0 && (((uintptr_t)g) & 15) == 0) {
            for (int i = lane; i < OB / 16; i += 32) ((uint4*)g)[i] = ((const uint4*)buf)[i];
        } else {
            for (int i = lane; i < OB; i += 32) g[i] = buf[i];
        }
        __syncwarp();
    }
}

// grouped observation without wrappers, streaming variant (OB % 16 == 0): one warp per ENV.
//   1. the env's record is staged in the warp's shared memory and its id plane expanded ONCE into the padded byte image
//      (bedrock frame persists in shared memory); every lane keeps its 16-byte slices of that image in registers;
//   2. lane a evaluates placement a (landing row, frame / game-over class, full-row mask) -- 32 placements per round;
//   3. per placement the warp streams the base image (or a constant fill) with one 128-bit store per lane: the output
//      of one env is 4W * OB contiguous bytes (17.3 KB at 10x20), written exactly once;
//   4. after a __syncwarp (orders the stores of the warp), the lane that owns a regular placement drops the four piece
//      cells on top of its board image with byte stores (they merge in L2);
//   5. placements that clear rows (rare) are composed row by row in a scratch image and stored from there.
// HBM bytes per env-step: 4W * OB written + record read; no re-reads.  Bound: HBM write bandwidth.
template <class COLT, int NV>
__global__ void __launch_bounds__(256) k_grouped_boards_stream(const DevCfg cfg, int64_t n, const uint8_t* hot, const uint8_t* board,
                                                               uint8_t* boards, uint8_t* legal, const uint8_t* fill_high, int rec_bytes,
                                                               int img_bytes) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ unsigned short s_cells[28];
    __shared__ int s_n[8];
    const int W = cfg.W, H = cfg.H, Wp = cfg.Wp, A = cfg.A, OB = cfg.OB, BS = cfg.board_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int NQ = OB >> 4;
    uint8_t* wbase = sm + (size_t)warp * (rec_bytes + 2 * img_bytes);
    uint32_t* rec = (uint32_t*)wbase;
    uint8_t* img = wbase + rec_bytes;
    uint8_t* scr = img + img_bytes;
    if (threadIdx.x < 28) s_cells[threadIdx.x] = (&c_cells[0][0])[threadIdx.x];
    if (threadIdx.x < 7) s_n[threadIdx.x] = c_n[threadIdx.x];
    Tabs tb;
    tb.cells = s_cells; tb.rowbytes = &c_rowbytes[0][0][0]; tb.n = s_n;
    for (int i = lane; i < OB; i += 32) {   // bedrock frame of the base and scratch images, once per warp
        int r = i / Wp, c = i - r * Wp;
        uint8_t v = (r < H && c >= P && c < P + W) ? 0 : 1;
        img[i] = v; scr[i] = v;
    }
    for (int i = BS / 4 + lane; i < rec_bytes / 4; i += 32) rec[i] = 0;   // ids_get8 may read one word past the id plane
    __syncthreads();
    const COLT* cols = (const COLT*)rec;
    const uint32_t* ids = rec + cfg.ids_off / 4;
    const COLT playfield = (COLT(1) << H) - 1;
    for (int64_t e = (int64_t)blockIdx.x * nwarps + warp; e < n; e += (int64_t)gridDim.x * nwarps) {
        const uint32_t* grec = (const uint32_t*)(board + e * BS);
        for (int i = lane; i < BS / 4; i += 32) rec[i] = grec[i];
        const uint32_t w0 = *(const uint32_t*)(hot + e * 32);
        const int piece = (w0 >> 13) & 7, rot0 = (w0 >> 16) & 3;
        const bool fh = fill_high && fill_high[e];
        __syncwarp();
        if (W == 10 && (H & 3) == 0) {
            for (int g4 = lane; g4 < (H >> 2); g4 += 32) fill_rows4_w10(ids + 5 * g4, img + g4 * 72);
        } else if (W == 20 && (H & 1) == 0) {
            for (int g2 = lane; g2 < (H >> 1); g2 += 32) fill_rows2_w20(ids + 5 * g2, img + g2 * 56);
        } else {
            for (int r = lane; r < H; r += 32) fill_board_row<0>(cfg, ids, img, 0, r);
        }
        __syncwarp();
        uint4 basev[NV];
#pragma unroll
        for (int j = 0; j < NV; j++) {
            int q = lane + 32 * j;
            basev[j] = q < NQ ? ((const uint4*)img)[q] : make_uint4(0, 0, 0, 0);
        }
        // placements of this env: lane `l` owns a = 32 * round + l
        uint32_t info[3];     // kind (bits 0-1: 0 regular, 1 frame -> ones, 2 game over -> zeros, 3 constant fill), bit 2 = rows get cleared
        uint32_t offlo[3], offhi[3];   // byte offsets of the 4 piece cells inside the board image (16 bits each)
#pragma unroll
        for (int rd = 0; rd < 3; rd++) {
            info[rd] = 3; offlo[rd] = 0; offhi[rd] = 0;
            const int a = rd * 32 + lane;
            if (rd * 32 < A && a < A && !fh) {
                COLT B;
                Placement pl = eval_placement<COLT>(cfg, tb, cols, piece, rot0, a, B);
                legal[e * A + a] = pl.kind != 1;
                uint32_t k = (uint32_t)pl.kind;
                if (pl.kind == 0) {
                    uint32_t cells = tb.cells[piece * 4 + pl.rot];
                    int crow[4], ccol[4];
                    COLT full = ~COLT(0);
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) {
                        int c = (cells >> (4 * c4)) & 15;
                        crow[c4] = pl.y + (c >> 2); ccol[c4] = pl.x + (c & 3) - P;
                    }
                    for (int c = 0; c < W; c++) {
                        COLT v = cols[c];
#pragma unroll
                        for (int c4 = 0; c4 < 4; c4++) if (ccol[c4] == c) v |= COLT(1) << crow[c4];
                        full &= v;
                    }
                    if (full & playfield) k |= 4u;
                    offlo[rd] = (uint32_t)(crow[0] * Wp + ccol[0] + P) | ((uint32_t)(crow[1] * Wp + ccol[1] + P) << 16);
                    offhi[rd] = (uint32_t)(crow[2] * Wp + ccol[2] + P) | ((uint32_t)(crow[3] * Wp + ccol[3] + P) << 16);
                }
                info[rd] = k;
            }
        }
        const uint32_t fhw = 0x01010101u * (uint32_t)(uint8_t)(cfg.H * cfg.W);
        uint8_t* genv = boards + (size_t)e * A * OB;
#pragma unroll
        for (int rd = 0; rd < 3; rd++) {
            if (rd * 32 < A) {
                const int na = min(32, A - rd * 32);
                for (int src = 0; src < na; src++) {
                    const uint32_t inf = __shfl_sync(0xffffffffu, info[rd], src);
                    const int a = rd * 32 + src;
                    uint4* g = (uint4*)(genv + (size_t)a * OB);
                    if (!(inf & 4u) || (inf & 3u) != 0) {
                        const uint32_t kind = inf & 3u;
                        const uint32_t fw = kind == 1 ? 0x01010101u : (kind == 2 ? 0u : fhw);
#pragma unroll
                        for (int j = 0; j < NV; j++) {
                            int q = lane + 32 * j;
                            if (q < NQ) g[q] = kind == 0 ? basev[j] : make_uint4(fw, fw, fw, fw);
                        }
                    } else {
                        // rows get cleared: project, compact (Tetris.clear_filled_rows on the copy, wrappers/grouped.py:171-177)
                        COLT B;
                        Placement pl = eval_placement<COLT>(cfg, tb, cols, piece, rot0, a, B);
                        uint32_t cells = tb.cells[piece * 4 + pl.rot];
                        int crow[4], ccol[4];
                        COLT full = ~COLT(0);
#pragma unroll
                        for (int c4 = 0; c4 < 4; c4++) {
                            int c = (cells >> (4 * c4)) & 15;
                            crow[c4] = pl.y + (c >> 2); ccol[c4] = pl.x + (c & 3) - P;
                        }
                        for (int c = 0; c < W; c++) {
                            COLT v = cols[c];
#pragma unroll
                            for (int c4 = 0; c4 < 4; c4++) if (ccol[c4] == c) v |= COLT(1) << crow[c4];
                            full &= v;
                        }
                        full &= playfield;
                        const int nclr = popc_t<COLT>(full);
                        for (int r = lane; r < H; r += 32) {
                            uint8_t* row = scr + r * Wp + P;
                            if (r < nclr) { for (int c = 0; c < W; c++) row[c] = 0; continue; }
                            int s = r - nclr;   // (r - nclr)-th surviving source row
                            COLT f = full;
                            while (f) { int fr = ctz_t<COLT>(f); f &= f - 1; if (fr <= s) s++; }
                            const uint8_t* srow = img + s * Wp + P;
                            for (int c = 0; c < W; c++) row[c] = srow[c];
#pragma unroll
                            for (int c4 = 0; c4 < 4; c4++) if (crow[c4] == s) row[ccol[c4]] = (uint8_t)(piece + 2);
                        }
                        __syncwarp();
                        for (int q = lane; q < NQ; q += 32) g[q] = ((const uint4*)scr)[q];
                        __syncwarp();
                    }
                }
            }
        }
        __syncwarp();   // orders this warp's image stores before the cell stores below
#pragma unroll
        for (int rd = 0; rd < 3; rd++) {
            const int a = rd * 32 + lane;
            if (rd * 32 < A && a < A && info[rd] == 0) {
                uint8_t* g = genv + (size_t)a * OB;
                const uint8_t v = (uint8_t)(piece + 2);
                g[offlo[rd] & 0xFFFFu] = v; g[offlo[rd] >> 16] = v; g[offhi[rd] & 0xFFFFu] = v; g[offhi[rd] >> 16] = v;
            }
        }
        __syncwarp();   // rec / img are rewritten by the next env
    }
}

// RgbObservation.observation (wrappers/observation.py:38-74): one warp per env, everything per-warp in shared memory:
//   record (cols + id plane) -> id image [Hp][RW] (bedrock / ones written once per warp, cells + queue + holder + active
//   piece per env) -> RGB bytes through a 16-entry colour LUT (4 pixels -> 3 words) -> one TMA bulk store per env.
template <class COLT>
__global__ void __launch_bounds__(256) k_rgb(const DevCfg cfg, int64_t n, const uint8_t* hot, const uint8_t* board, uint8_t* img,
                                             int rec_bytes, int pix_bytes, int rgb_bytes) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint32_t s_lut[16];
    __shared__ uint32_t s_rowbytes[112];
    const int W = cfg.W, H = cfg.H, Wp = cfg.Wp, Hp = cfg.Hp, RW = cfg.rgb_w, Q = cfg.Q;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int NP = Hp * RW;
    uint8_t* wbase = sm + (size_t)warp * (rec_bytes + pix_bytes + rgb_bytes);
    uint32_t* rec = (uint32_t*)wbase;
    uint8_t* pix = wbase + rec_bytes;
    uint8_t* rgb = pix + pix_bytes;
    if (threadIdx.x < 16) s_lut[threadIdx.x] = ((const uint32_t*)c_colors)[threadIdx.x];
    for (int i = threadIdx.x; i < 112; i += blockDim.x) s_rowbytes[i] = (&c_rowbytes[0][0][0])[i];
    // constant part of the id image: everything that is not a playfield cell, a queue cell or a holder cell is 1
    for (int i = lane; i < NP; i += 32) {
        int r = i / RW, c = i - r * RW;
        pix[i] = (c < Wp && r < H && c >= P && c < P + W) ? 0 : 1;
    }
    __syncthreads();
    const bool fast = (NP & 3) == 0;
    const bool tma = ((NP * 3) & 15) == 0 && (((uintptr_t)img) & 15) == 0;
    for (int64_t e = (int64_t)blockIdx.x * nwarps + warp; e < n; e += (int64_t)gridDim.x * nwarps) {
        // previous env's bulk store must have finished reading this warp's rgb buffer
        if (lane == 0) bulk_wait_read();
        __syncwarp();
        const uint32_t* grec = (const uint32_t*)(board + e * cfg.board_stride);
        for (int i = lane; i < cfg.board_stride / 4; i += 32) rec[i] = grec[i];
        Hot h;
        hot_load(h, (const uint32_t*)(hot + e * 32));
        __syncwarp();
        const COLT* cols = (const COLT*)rec;
        const uint32_t* ids = rec + cfg.ids_off / 4;
        // board rows (cells only; the frame persists)
        for (int r = lane; r < H; r += 32) fill_board_row<0>(cfg, ids, pix, 0, r, RW);
        // queue (top right) and holder (bottom right)
        for (int it = lane; it < 4 * Q; it += 32) {
            int i = it / Q, q = it - i * Q;
            uint32_t wv = s_rowbytes[((int)((h.queue >> (4 * q)) & 15u)) * 16 + i];
            uint8_t* d = pix + i * RW + Wp + 4 * q;
            d[0] = (uint8_t)wv; d[1] = (uint8_t)(wv >> 8); d[2] = (uint8_t)(wv >> 16); d[3] = (uint8_t)(wv >> 24);
        }
        if (lane < 4) {
            uint32_t wv = h.hold ? s_rowbytes[((h.hold - 1) * 4 + h.hold_r) * 4 + lane] : 0x01010101u;
            uint8_t* d = pix + (Hp - P + lane) * RW + Wp;
            d[0] = (uint8_t)wv; d[1] = (uint8_t)(wv >> 8); d[2] = (uint8_t)(wv >> 16); d[3] = (uint8_t)(wv >> 24);
        }
        __syncwarp();
        uint32_t cells = c_cells[h.p][h.r];
        COLT B = bmask<COLT>(cols, W, cells, h.x);
        if (!((B >> h.y) & 1) && lane < 4) {   // active piece on top (project_tetromino, envs/tetris.py:543-564)
            int c = (cells >> (4 * lane)) & 15;
            pix[(h.y + (c >> 2)) * RW + h.x + (c & 3)] = (uint8_t)(h.p + 2);
        }
        __syncwarp();
        uint8_t* g = img + (size_t)e * NP * 3;
        if (fast) {
            uint32_t* o = tma ? (uint32_t*)rgb : (uint32_t*)g;
            for (int q4 = lane; q4 < NP / 4; q4 += 32) {
                uint32_t pv = ((const uint32_t*)pix)[q4];
                uint32_t c0 = s_lut[pv & 15], c1 = s_lut[(pv >> 8) & 15], c2 = s_lut[(pv >> 16) & 15], c3 = s_lut[(pv >> 24) & 15];
                o[3 * q4] = c0 | (c1 << 24);
                o[3 * q4 + 1] = (c1 >> 8) | (c2 << 16);
                o[3 * q4 + 2] = (c2 >> 16) | (c3 << 8);
            }
            if (tma) {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) { bulk_s2g(g, rgb, (uint32_t)(NP * 3)); bulk_commit(); }
            }
        } else {
            for (int i = lane; i < NP * 3; i += 32) { int px = i / 3; g[i] = (uint8_t)(s_lut[pix[px]] >> (8 * (i - 3 * px))); }
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait_all();
}

}  // namespace tg

// ---- host entry points -------------------------------------------------------------------------------
extern "C" int tg_features(tg_env* env, tg_state st, int64_t n, uint8_t* d_feats, void* stream) {
    if (!env) return TG_ERR_POINTER;
    int rc = check_state(env, st); if (rc) return rc;
    if (!d_feats) return fail(env, TG_ERR_POINTER, "d_feats is NULL");
    CUDA_TRY(env, cudaSetDevice(env->device));
    int T = 128;
    unsigned g = (unsigned)((n + T - 1) / T);
    if (env->col64) k_features<uint64_t><<<g, T, 0, (cudaStream_t)stream>>>(env->dev, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_feats);
    else k_features<uint32_t><<<g, T, 0, (cudaStream_t)stream>>>(env->dev, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_feats);
    CUDA_TRY(env, cudaGetLastError());
    return TG_OK;
}

extern "C" int tg_render_rgb(tg_env* env, tg_state st, int64_t n, uint8_t* d_img, void* stream) {
    if (!env) return TG_ERR_POINTER;
    int rc = check_state(env, st); if (rc) return rc;
    if (!d_img) return fail(env, TG_ERR_POINTER, "d_img is NULL");
    CUDA_TRY(env, cudaSetDevice(env->device));
    const DevCfg& d = env->dev;
    auto r128 = [](size_t v) { return (int)((v + 127) / 128 * 128); };
    int rec_bytes = r128((size_t)d.board_stride + 16), pix_bytes = r128((size_t)d.Hp * d.rgb_w + 16), rgb_bytes = r128((size_t)d.Hp * d.rgb_w * 3);
    size_t per_warp = (size_t)rec_bytes + pix_bytes + rgb_bytes;
    int nw = 8;
    while (nw > 1 && per_warp * nw > 72 * 1024) nw >>= 1;
    size_t smem = per_warp * nw;
    if (smem > 227 * 1024) return fail(env, TG_ERR_CONFIG, "tg_render_rgb: image too large for shared memory");
    int T = nw * 32;
    auto launch = [&](auto kern) -> int {
        CUDA_TRY(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        CUDA_TRY(env, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem));
        int64_t blocks = (n + nw - 1) / nw, cap = (int64_t)env->num_sms * (per_sm > 0 ? per_sm : 1);
        if (blocks > cap) blocks = cap;
        kern<<<(unsigned)blocks, T, smem, (cudaStream_t)stream>>>(d, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_img, rec_bytes, pix_bytes, rgb_bytes);
        CUDA_TRY(env, cudaGetLastError());
        return TG_OK;
    };
    return env->col64 ? launch(k_rgb<uint64_t>) : launch(k_rgb<uint32_t>);
}

static int launch_grouped_observe(tg_env* env, tg_state st, int64_t n, uint8_t* d_feats, uint8_t* d_boards, uint8_t* d_legal,
                                  const uint8_t* fill_high, cudaStream_t s) {
    const DevCfg& d = env->dev;
    if (d_feats) {
        int EPB = 16, T = 256;
        size_t colb = env->col64 ? 8 : 4;
        size_t smem = (size_t)3 * EPB * d.W * colb + (size_t)EPB * 16 + (size_t)EPB * 4 + (size_t)EPB * 64 + (size_t)EPB * d.A * d.F + (size_t)EPB * d.A;
        if (smem > 48 * 1024) {
            if (env->col64) cudaFuncSetAttribute(k_grouped_feats<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            else cudaFuncSetAttribute(k_grouped_feats<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        unsigned g = (unsigned)((n + EPB - 1) / EPB);
        if (env->col64) k_grouped_feats<uint64_t><<<g, T, smem, s>>>(d, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_feats, d_legal, fill_high, EPB);
        else k_grouped_feats<uint32_t><<<g, T, smem, s>>>(d, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_feats, d_legal, fill_high, EPB);
        CUDA_TRY(env, cudaGetLastError());
    }
    const bool stream_ok = d_boards && (d.OB & 15) == 0 && (((uintptr_t)d_boards) & 15) == 0 && d.OB <= 4 * 32 * 16 && !getenv("TG_BOARDS_V1");
    if (stream_ok) {
        const int T = 256, nw = T / 32;
        const int NV = (d.OB / 16 + 31) / 32;
        auto r128 = [](size_t v) { return (int)((v + 127) / 128 * 128); };
        const int rec_bytes = r128((size_t)d.board_stride + 16), img_bytes = r128((size_t)d.OB);
        const size_t smem = (size_t)nw * (rec_bytes + 2 * img_bytes);
        auto launch = [&](auto kern) -> int {
            CUDA_TRY(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 1;
            CUDA_TRY(env, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem));
            int64_t blocks = (n + nw - 1) / nw, cap = (int64_t)env->num_sms * (per_sm > 0 ? per_sm : 1);
            if (blocks > cap) blocks = cap;
            kern<<<(unsigned)blocks, T, smem, s>>>(d, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_boards, d_legal, fill_high, rec_bytes, img_bytes);
            CUDA_TRY(env, cudaGetLastError());
            return TG_OK;
        };
        int rc;
        if (env->col64) rc = NV <= 1 ? launch(k_grouped_boards_stream<uint64_t, 1>) : NV == 2 ? launch(k_grouped_boards_stream<uint64_t, 2>)
                                   : NV == 3 ? launch(k_grouped_boards_stream<uint64_t, 3>) : launch(k_grouped_boards_stream<uint64_t, 4>);
        else rc = NV <= 1 ? launch(k_grouped_boards_stream<uint32_t, 1>) : NV == 2 ? launch(k_grouped_boards_stream<uint32_t, 2>)
                          : NV == 3 ? launch(k_grouped_boards_stream<uint32_t, 3>) : launch(k_grouped_boards_stream<uint32_t, 4>);
        if (rc) return rc;
    } else if (d_boards) {
        int T = 256, nw = T / 32;
        size_t smem = (size_t)nw * (((size_t)d.OB + 15) & ~(size_t)15);
        int64_t blocks = (n * d.A + nw - 1) / nw;
        int64_t cap = (int64_t)env->num_sms * 8;
        if (blocks > cap) blocks = cap;
        if (env->col64) k_grouped_boards<uint64_t><<<(unsigned)blocks, T, smem, s>>>(d, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_boards, d_legal, fill_high);
        else k_grouped_boards<uint32_t><<<(unsigned)blocks, T, smem, s>>>(d, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_boards, d_legal, fill_high);
        CUDA_TRY(env, cudaGetLastError());
    }
    return TG_OK;
}

extern "C" int tg_grouped_observe(tg_env* env, tg_state st, int64_t n, uint8_t* d_feats, uint8_t* d_boards, uint8_t* d_legal,
                                  void* stream) {
    if (!env) return TG_ERR_POINTER;
    int rc = check_state(env, st); if (rc) return rc;
    if (!d_legal || (!d_feats && !d_boards)) return fail(env, TG_ERR_POINTER, "grouped_observe: need d_legal and d_feats or d_boards");
    CUDA_TRY(env, cudaSetDevice(env->device));
    return launch_grouped_observe(env, st, n, d_feats, d_boards, d_legal, nullptr, (cudaStream_t)stream);
}

extern "C" int tg_grouped_step(tg_env* env, tg_state st, int64_t n, const int32_t* d_actions, uint8_t* d_legal, uint8_t* d_feats,
                               uint8_t* d_boards, uint8_t* d_info_board, tg_obs obs, tg_step_out out, tg_stats* d_stats,
                               void* stream) {
    if (!env) return TG_ERR_POINTER;
    if (n <= 0) return fail(env, TG_ERR_ARG, "n must be positive");
    int rc = check_state(env, st); if (rc) return rc;
    if (!d_actions || !d_legal || !out.reward || !out.terminated || !out.truncated || !out.lines)
        return fail(env, TG_ERR_POINTER, "grouped_step: NULL pointer");
    bool any_obs = obs.board || obs.mask || obs.holder || obs.queue;
    if (any_obs) { rc = check_obs(env, obs); if (rc) return rc; }
    CUDA_TRY(env, cudaSetDevice(env->device));
    rc = ensure_stage(env, 3, (size_t)n); if (rc) return rc;  // per-env "illegal + terminate" flags
    StepParams p;
    memset(&p, 0, sizeof p);
    p.n = n; p.hot = (uint8_t*)st.hot; p.board = (uint8_t*)st.board; p.rng = (uint8_t*)st.rng; p.seq = st.piece_seq;
    p.actions = d_actions;
    p.o_board = obs.board; p.o_mask = obs.mask; p.o_holder = obs.holder; p.o_queue = obs.queue;
    p.reward = out.reward; p.terminated = out.terminated; p.truncated = out.truncated; p.lines = out.lines;
    p.stats = (double*)d_stats;
    p.legal = d_legal; p.info_board = d_info_board; p.fill_high = (uint8_t*)env->stage[3];
    p.mode = 2;
    rc = launch_step(env, p, (cudaStream_t)stream); if (rc) return rc;
    if (d_feats || d_boards) return launch_grouped_observe(env, st, n, d_feats, d_boards, d_legal, (const uint8_t*)env->stage[3], (cudaStream_t)stream);
    return TG_OK;
}

extern "C" int tg_rollout(tg_env* env, tg_state st, int64_t n, const int32_t weights[4], int32_t k_steps, tg_stats* d_stats,
                          void* stream) {
    if (!env) return TG_ERR_POINTER;
    if (n <= 0 || k_steps < 0 || !weights) return fail(env, TG_ERR_ARG, "tg_rollout: bad argument");
    int rc = check_state(env, st); if (rc) return rc;
    CUDA_TRY(env, cudaSetDevice(env->device));
    const DevCfg& d = env->dev;
    RolloutParams p;
    memset(&p, 0, sizeof p);
    p.cfg = d; p.n = n; p.hot = (uint8_t*)st.hot; p.board = (uint8_t*)st.board; p.rng = (uint8_t*)st.rng; p.seq = st.piece_seq;
    for (int i = 0; i < 4; i++) p.w[i] = weights[i];
    p.k_steps = k_steps; p.stats = (double*)d_stats;
    p.last_action = (int32_t*)env->rollout_last_action;
    int words = (d.board_stride + 4) / 4;                 // +1 word: ids_get8 may read one word past the id plane
    if (env->col64) words = (words + 1) & ~1;             // keep the EnvBase arrays 8-byte aligned
    p.base_off = words;
    words += 2 * d.W * (env->col64 ? 2 : 1) + 16;         // pre[W], suf[W], h[32 B], ho[32 B]
    // u32 columns: odd word stride; u64 columns: stride = 2 (mod 4) words keeps 8-byte alignment and spreads the banks
    if (env->col64) { while ((words & 3) != 2) words += 1; } else { words |= 1; }
    p.rec_words = words;
    const int T = 128;
    size_t smem = (size_t)T * words * 4;
    auto launch = [&](auto kern) -> int {
        CUDA_TRY(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)((n + T - 1) / T), T, smem, (cudaStream_t)stream>>>(p);
        CUDA_TRY(env, cudaGetLastError());
        return TG_OK;
    };
    if (smem > 227 * 1024) return fail(env, TG_ERR_CONFIG, "tg_rollout: board record too large for shared memory");
    return env->col64 ? launch(k_rollout<uint64_t>) : launch(k_rollout<uint32_t>);
}

/* test hook: device buffer (i32[n]) that receives the action chosen at the last rollout step; NULL disables */
extern "C" int tg_debug_set_rollout_trace(tg_env* env, int32_t* d_last_action) {
    if (!env) return TG_ERR_POINTER;
    env->rollout_last_action = d_last_action;
    return TG_OK;
}
