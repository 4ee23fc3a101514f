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
// CTA = EPB envs; one thread per env precomputes the EnvBase (heights, holes, prefix / suffix column ANDs), then one
// thread per (env, placement) evaluates it from the piece's column profile (place_fast: no bit scans on the common
// path); placements that clear rows are batched into a second pass with the exact evaluation.
template <class COLT>
__global__ void __launch_bounds__(256) k_grouped_feats(const DevCfg cfg, int64_t n, const uint8_t* hot, const uint8_t* board, uint8_t* feats,
                                                       uint8_t* legal, const uint8_t* fill_high, int EPB, uint32_t magicA) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int W = cfg.W, A = cfg.A, F = cfg.F, WP = cfg.W + 2 * P;   // magicA = ceil(2^20 / A): it / A == (it * magicA) >> 20 for it < EPB * A
    COLT* s_colp = (COLT*)sm;                      // [EPB][W + 2P]  columns with P wall columns on both sides
    COLT* s_pre = s_colp + (size_t)EPB * WP;       // [EPB][W]
    COLT* s_suf = s_pre + (size_t)EPB * W;         // [EPB][W]
    int* s_sum = (int*)(s_suf + (size_t)EPB * W);  // [EPB][4]: sum_h, holes, bump, max_h
    uint32_t* s_w0 = (uint32_t*)(s_sum + EPB * 4); // [EPB]
    uint8_t* s_h = (uint8_t*)(s_w0 + EPB);         // [EPB][32]
    uint8_t* s_ho = s_h + EPB * 32;                // [EPB][32]
    uint16_t* s_bs = (uint16_t*)(s_ho + EPB * 32); // [EPB][32]
    uint8_t* s_feats = (uint8_t*)(s_bs + EPB * 32);   // [EPB][A][F]   (16-aligned: every block above is a multiple of 16 for EPB % 4 == 0)
    uint8_t* s_legal = s_feats + (size_t)EPB * A * F;
    __shared__ unsigned short s_cells[28];
    __shared__ uint2 s_ptab[28];
    __shared__ int s_n[8];
    __shared__ unsigned short s_slow[32 * 96];   // EPB <= 32, A <= 96
    __shared__ int s_nslow;
    if (threadIdx.x < 28) { s_cells[threadIdx.x] = (&c_cells[0][0])[threadIdx.x]; s_ptab[threadIdx.x] = (&c_ptab[0][0])[threadIdx.x]; }
    if (threadIdx.x < 7) s_n[threadIdx.x] = c_n[threadIdx.x];
    if (threadIdx.x == 0) s_nslow = 0;
    Tabs tb;
    tb.ptab = s_ptab; tb.cells = s_cells; tb.rowbytes = &c_rowbytes[0][0][0]; tb.n = s_n;
    const int64_t base = (int64_t)blockIdx.x * EPB;
    const int nv = (int)min((int64_t)EPB, n - base);
    for (int i = threadIdx.x; i < nv * WP; i += blockDim.x) {
        int e = i / WP, c = i - e * WP - P;
        s_colp[i] = (unsigned)c < (unsigned)W ? ((const COLT*)(board + (base + e) * cfg.board_stride))[c] : ~COLT(0);
    }
    for (int i = threadIdx.x; i < nv; i += blockDim.x)   // bit 31: illegal action + terminate -> the observation is filled with `high`
        s_w0[i] = (*(const uint32_t*)(hot + (base + i) * 32) & 0x7FFFFFFFu) | ((fill_high && fill_high[base + i]) ? 0x80000000u : 0u);
    __syncthreads();
    if (threadIdx.x < nv) {
        int e = threadIdx.x;
        EnvBase<COLT> eb;
        eb.h = s_h + e * 32; eb.ho = s_ho + e * 32; eb.bs = s_bs + e * 32; eb.pre = s_pre + e * W; eb.suf = s_suf + e * W;
        env_base_compute<COLT>(cfg, s_colp + e * WP + P, COLT(1), eb);
        s_sum[e * 4] = eb.sum_h; s_sum[e * 4 + 1] = eb.holes; s_sum[e * 4 + 2] = eb.bump; s_sum[e * 4 + 3] = eb.max_h;
    }
    __syncthreads();
    for (int it = threadIdx.x; it < nv * A; it += blockDim.x) {
        const int e = (int)(((uint32_t)it * magicA) >> 20), a = it - e * A;
        uint32_t w0 = s_w0[e];
        int piece = (w0 >> 13) & 7, rot0 = (w0 >> 16) & 3;
        uint8_t* out = s_feats + it * F;
        if (w0 >> 31) {
            // illegal action + terminate: obs = ones * high (wrappers/grouped.py:221-226); legal mask unchanged
            for (int i = 0; i < F; i++) out[i] = (uint8_t)min(255, cfg.H * cfg.W);   // observation_space.high = H * W, saturated to the uint8 range
            s_legal[it] = legal[(base + e) * A + a];
            continue;
        }
        EnvBase<COLT> eb;
        eb.h = s_h + e * 32; eb.ho = s_ho + e * 32; eb.bs = s_bs + e * 32; eb.pre = s_pre + e * W; eb.suf = s_suf + e * W;
        eb.sum_h = s_sum[e * 4]; eb.holes = s_sum[e * 4 + 1]; eb.bump = s_sum[e * 4 + 2]; eb.max_h = s_sum[e * 4 + 3];
        const int rot = (rot0 + (a & 3)) & 3;              // cumulative rot90 presses (wrappers/grouped.py:153-154)
        const int x = (a >> 2) + P - tb.n[piece] / 2;      // wrappers/grouped.py:157-158
        FeatSum fs;
        int y;
        const int kind = place_fast<COLT>(cfg, eb, s_colp + e * WP, tb.cells[piece * 4 + rot], tb.ptab[piece * 4 + rot], x, fs, y, out, true);
        s_legal[it] = kind != 1;
        if (kind == 1) {          // ones board, row 0 zeroed -> heights H-1
            for (int i = 0; i <= W; i++) out[i] = (uint8_t)(cfg.H - 1);
            out[W + 1] = 0; out[W + 2] = 0;
        } else if (kind == 2) {   // zeros board
            for (int i = 0; i < F; i++) out[i] = 0;
        } else if (kind >= 3) {
            s_slow[atomicAdd(&s_nslow, 1)] = (unsigned short)it;   // rows get cleared / a cell in the zeroed row 0: second pass
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < s_nslow; k += blockDim.x) {
        const int it = s_slow[k], e = (int)(((uint32_t)it * magicA) >> 20), a = it - e * A;
        uint32_t w0 = s_w0[e];
        int piece = (w0 >> 13) & 7, rot0 = (w0 >> 16) & 3;
        EnvBase<COLT> eb;
        eb.h = s_h + e * 32; eb.ho = s_ho + e * 32; eb.bs = s_bs + e * 32; eb.pre = s_pre + e * W; eb.suf = s_suf + e * W;
        eb.sum_h = s_sum[e * 4]; eb.holes = s_sum[e * 4 + 1]; eb.bump = s_sum[e * 4 + 2]; eb.max_h = s_sum[e * 4 + 3];
        const int rot = (rot0 + (a & 3)) & 3, x = (a >> 2) + P - tb.n[piece] / 2;
        FeatSum fs;
        int y;
        const int kind = place_fast<COLT>(cfg, eb, s_colp + e * WP, tb.cells[piece * 4 + rot], tb.ptab[piece * 4 + rot], x, fs, y, s_feats + it * F, false);
        if (kind == 3)
            placement_eval<COLT>(cfg, s_colp + e * WP + P, tb.cells[piece * 4 + rot], x, y, true, true, COLT(1), s_feats + it * F);
    }
    __syncthreads();
    // coalesced copy-out of the tile (contiguous in global memory)
    {
        const int bytes = nv * A * F;
        uint8_t* g = feats + (size_t)base * A * F;
        if ((bytes & 15) == 0 && (((uintptr_t)g) & 15) == 0) {
            for (int i = threadIdx.x; i < (bytes >> 4); i += blockDim.x) ((uint4*)g)[i] = ((const uint4*)s_feats)[i];
        } else {
            for (int i = threadIdx.x; i < bytes; i += blockDim.x) g[i] = s_feats[i];
        }
        const int lb = nv * A;
        uint8_t* gl = legal + (size_t)base * A;
        if ((lb & 3) == 0 && (((uintptr_t)gl) & 3) == 0) {
            for (int i = threadIdx.x; i < (lb >> 2); i += blockDim.x) ((uint32_t*)gl)[i] = ((const uint32_t*)s_legal)[i];
        } else {
            for (int i = threadIdx.x; i < lb; i += blockDim.x) gl[i] = s_legal[i];
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
        if (fill_high && fill_high[e]) fillv = (uint8_t)min(255, cfg.H * cfg.W);
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

// ---- per-warp record prefetch (Ampere-style cp.async, 16-byte chunks; each lane waits for its own copies) ----
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// record of env e (BS bytes) followed by the first 16 bytes of its hot record, into one warp-private buffer
__device__ __forceinline__ void warp_prefetch_env(uint8_t* dst, const uint8_t* board, const uint8_t* hot, int64_t e, int BS, int lane) {
    const uint8_t* src = board + e * BS;
    for (int i = lane; i < (BS >> 4); i += 32) cp_async16(dst + 16 * i, src + 16 * i);
    if (lane == 0) cp_async16(dst + BS, hot + e * 32);
    cp_async_commit();
}

// prefix / suffix ANDs of the occupancy columns of one env, computed by a warp (lane c holds column c):
//   pre[c] = AND of columns < c,  suf[c] = AND of columns > c      (see EnvBase in tg_device.cuh)
template <class COLT>
__device__ __forceinline__ void warp_pre_suf(const COLT* cols, int W, int lane, COLT* pre, COLT* suf) {
    COLT v = lane < W ? cols[lane] : ~COLT(0), u = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        COLT t = __shfl_up_sync(0xffffffffu, v, d), s = __shfl_down_sync(0xffffffffu, u, d);
        if (lane >= d) v &= t;
        if (lane + d < 32) u &= s;
    }
    COLT pe = __shfl_up_sync(0xffffffffu, v, 1), se = __shfl_down_sync(0xffffffffu, u, 1);
    if (lane < W) { pre[lane] = lane ? pe : ~COLT(0); suf[lane] = lane < 31 ? se : ~COLT(0); }
}

// grouped observation without wrappers, streaming variant (OB % 16 == 0): one warp per ENV.
//   1. the env's record is prefetched (cp.async, double buffered) into the warp's shared memory and its id plane expanded
//      ONCE into the padded byte image (bedrock frame persists); every lane keeps its 16-byte slices of it in registers;
//   2. lane a evaluates placement a (landing row, frame / game-over class, full-row mask via prefix / suffix column ANDs)
//      -- 32 placements per round; the classes are exchanged as ballot masks, so the composition loops are warp-uniform;
//   3. the placement images are composed G at a time in a double-buffered group buffer: base image (or a constant fill) from
//      registers with one 128-bit shared store per lane and slot, the four piece cells dropped in by the owning lane,
//      row-clearing placements (rare) recomposed row by row;
//   4. one TMA bulk store per group (G * OB contiguous bytes, 3.4 KB at 10x20); the next group is composed meanwhile.
// Every output byte is written exactly once, in full 16-byte pieces (measured: patching cells in global memory after the
// image stores costs 17 % -- partial sector writes).  HBM bytes per env-step: 4W * OB written + record read.
// Bound: HBM write bandwidth.
template <class COLT, int NV>
__global__ void __launch_bounds__(256, 3) k_grouped_boards_stream(const DevCfg cfg, int64_t n, const uint8_t* hot, const uint8_t* board,
                                                                  uint8_t* boards, uint8_t* legal, const uint8_t* fill_high, int rec_bytes,
                                                                  int img_bytes, int gbuf_bytes) {
    constexpr int G = NV == 1 ? 8 : (NV <= 3 ? 4 : 2);   // placements per bulk store (G * OB <= 6 KB)
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ unsigned short s_cells[28];
    __shared__ int s_n[8];
    const int W = cfg.W, H = cfg.H, Wp = cfg.Wp, A = cfg.A, OB = cfg.OB, BS = cfg.board_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int NQ = OB >> 4;
    uint8_t* wbase = sm + (size_t)warp * (2 * rec_bytes + img_bytes + 512 + 2 * gbuf_bytes);
    uint8_t* recbuf = wbase;                       // two record buffers (record + 16 B of the hot record)
    uint8_t* img = wbase + 2 * rec_bytes;
    COLT* s_pre = (COLT*)(img + img_bytes);        // [W] (<= 24 x 8 B)
    COLT* s_suf = s_pre + 32;
    uint8_t* gbuf = img + img_bytes + 512;         // two group buffers of G slots (OB bytes each, contiguous)
    if (threadIdx.x < 28) s_cells[threadIdx.x] = (&c_cells[0][0])[threadIdx.x];
    if (threadIdx.x < 7) s_n[threadIdx.x] = c_n[threadIdx.x];
    Tabs tb;
    tb.cells = s_cells; tb.rowbytes = &c_rowbytes[0][0][0]; tb.n = s_n;
    for (int i = lane; i < OB; i += 32) {   // bedrock frame of the base image, once per warp
        int r = i / Wp, c = i - r * Wp;
        img[i] = (r < H && c >= P && c < P + W) ? 0 : 1;
    }
    for (int i = lane; i < 2 * rec_bytes / 4; i += 32) ((uint32_t*)recbuf)[i] = 0;
    __syncthreads();
    const COLT playfield = (COLT(1) << H) - 1;
    const int64_t stride = (int64_t)gridDim.x * nwarps;
    int64_t e = (int64_t)blockIdx.x * nwarps + warp;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch, see k_step_ws
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (e < n) warp_prefetch_env(recbuf, board, hot, e, BS, lane);
    bool act[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) act[j] = lane + 32 * j < NQ;
    const uint32_t fhw = 0x01010101u * (uint32_t)min(255, cfg.H * cfg.W);
    uint32_t gcount = 0;   // groups issued by this warp (selects the group buffer)
    for (int it = 0; e < n; e += stride, it++) {
        const uint32_t* rec = (const uint32_t*)(recbuf + (it & 1) * rec_bytes);
        const COLT* cols = (const COLT*)rec;
        const uint32_t* ids = rec + cfg.ids_off / 4;
        cp_async_wait_all();
        __syncwarp();
        if (e + stride < n) warp_prefetch_env(recbuf + ((it + 1) & 1) * rec_bytes, board, hot, e + stride, BS, lane);
        const uint32_t w0 = rec[BS >> 2];
        const int piece = (w0 >> 13) & 7, rot0 = (w0 >> 16) & 3;
        const bool fh = fill_high && fill_high[e];
        if (W == 10 && (H & 3) == 0) {
            for (int g4 = lane; g4 < (H >> 2); g4 += 32) fill_rows4_w10(ids + 5 * g4, img + g4 * 72);
        } else if (W == 20 && (H & 1) == 0) {
            for (int g2 = lane; g2 < (H >> 1); g2 += 32) fill_rows2_w20(ids + 5 * g2, img + g2 * 56);
        } else {
            for (int r = lane; r < H; r += 32) fill_board_row<0>(cfg, ids, img, 0, r);
        }
        warp_pre_suf<COLT>(cols, W, lane, s_pre, s_suf);
        __syncwarp();
        uint4 basev[NV];
#pragma unroll
        for (int j = 0; j < NV; j++) basev[j] = act[j] ? ((const uint4*)img)[lane + 32 * j] : make_uint4(0, 0, 0, 0);
        // placements of this env: lane `l` owns a = 32 * round + l
        // class: 0 regular, 1 frame -> ones, 2 game over -> zeros, 3 constant fill (illegal action + terminate), 4 regular + rows cleared, 8 none
        uint8_t* genv = boards + (size_t)e * A * OB;
#pragma unroll 1
        for (int rd = 0; rd * 32 < A; rd++) {   // one round = 32 placements; code is NOT unrolled over rounds (instruction cache)
            {
                uint32_t info = 8, offlo = 0, offhi = 0;   // off*: byte offsets of the 4 piece cells inside the board image (16 bits each)
                const int a = rd * 32 + lane;
                if (a < A) {
                    uint32_t k = 3;
                    if (!fh) {
                        COLT B;
                        Placement pl = eval_placement<COLT>(cfg, tb, cols, piece, rot0, a, B);
                        legal[e * A + a] = pl.kind != 1;
                        k = (uint32_t)pl.kind;
                        if (pl.kind == 0) {
                            uint32_t cells = tb.cells[piece * 4 + pl.rot];
                            int crow[4], ccol[4], c0 = 64, c1 = -1;
#pragma unroll
                            for (int c4 = 0; c4 < 4; c4++) {
                                int c = (cells >> (4 * c4)) & 15;
                                crow[c4] = pl.y + (c >> 2); ccol[c4] = pl.x + (c & 3) - P;
                                c0 = min(c0, ccol[c4]); c1 = max(c1, ccol[c4]);
                            }
                            COLT full = s_pre[c0] & s_suf[c1] & playfield;
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                if (c0 + j <= c1) {
                                    COLT v = cols[c0 + j];
#pragma unroll
                                    for (int c4 = 0; c4 < 4; c4++) if (ccol[c4] == c0 + j) v |= COLT(1) << crow[c4];
                                    full &= v;
                                }
                            }
                            if (full) k = 4;
                            offlo = (uint32_t)(crow[0] * Wp + ccol[0] + P) | ((uint32_t)(crow[1] * Wp + ccol[1] + P) << 16);
                            offhi = (uint32_t)(crow[2] * Wp + ccol[2] + P) | ((uint32_t)(crow[3] * Wp + ccol[3] + P) << 16);
                        }
                    }
                    info = k;
                }
                const uint32_t m_base = __ballot_sync(0xffffffffu, info == 0 || info == 4);
                const uint32_t m_one = __ballot_sync(0xffffffffu, info == 1);
                const uint32_t m_high = __ballot_sync(0xffffffffu, info == 3);
                const uint32_t m_slow = __ballot_sync(0xffffffffu, info == 4);
                const int na = min(32, A - rd * 32);
                for (int l0 = 0; l0 < na; l0 += G, gcount++) {
                    uint8_t* buf = gbuf + (gcount & 1u) * gbuf_bytes;
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last used `buf` has read it
                    __syncwarp();
                    // base image into every slot of the group (one 128-bit store per lane and slot) ...
                    {
                        uint8_t* d = buf + 16 * lane;
#pragma unroll
                        for (int sl = 0; sl < G; sl++, d += OB) {
#pragma unroll
                            for (int j = 0; j < NV; j++) if (act[j]) ((uint4*)d)[32 * j] = basev[j];
                        }
                    }
                    // ... then the constant fills (frame -> ones, game over -> zeros, illegal + terminate -> high)
                    for (uint32_t m = (~m_base >> l0) & ((1u << G) - 1); m; m &= m - 1) {
                        const int sl = __ffs((int)m) - 1, bit = l0 + sl;
                        const uint32_t fw = ((m_one >> bit) & 1) ? 0x01010101u : (((m_high >> bit) & 1) ? fhw : 0u);
                        const uint4 fv = make_uint4(fw, fw, fw, fw);
                        uint4* d = (uint4*)(buf + sl * OB) + lane;
#pragma unroll
                        for (int j = 0; j < NV; j++) if (act[j]) d[32 * j] = fv;
                    }
                    __syncwarp();
                    if (lane >= l0 && lane < l0 + G && info == 0) {   // project_tetromino: the four cells of this lane's placement
                        uint8_t* d = buf + (lane - l0) * OB;
                        const uint8_t v = (uint8_t)(piece + 2);
                        d[offlo & 0xFFFFu] = v; d[offlo >> 16] = v; d[offhi & 0xFFFFu] = v; d[offhi >> 16] = v;
                    }
                    for (uint32_t m = (m_slow >> l0) & ((1u << G) - 1); m; m &= m - 1) {
                        // rows get cleared: project, compact (Tetris.clear_filled_rows on the copy, wrappers/grouped.py:171-177)
                        const int sl = __ffs((int)m) - 1, a = rd * 32 + l0 + sl;
                        COLT B;
                        Placement pl = eval_placement<COLT>(cfg, tb, cols, piece, rot0, a, B);
                        uint32_t cells = tb.cells[piece * 4 + pl.rot];
                        int crow[4], ccol[4];
                        COLT full = playfield;
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
                        const int nclr = popc_t<COLT>(full);
                        for (int r = lane; r < H; r += 32) {
                            uint8_t* row = buf + sl * OB + r * Wp + P;
                            if (r < nclr) { for (int c = 0; c < W; c++) row[c] = 0; continue; }
                            int s = r - nclr;   // (r - nclr)-th surviving source row
                            COLT f = full;
                            while (f) { int fr = ctz_t<COLT>(f); f &= f - 1; if (fr <= s) s++; }
                            const uint8_t* srow = img + s * Wp + P;
                            for (int c = 0; c < W; c++) row[c] = srow[c];
#pragma unroll
                            for (int c4 = 0; c4 < 4; c4++) if (crow[c4] == s) row[ccol[c4]] = (uint8_t)(piece + 2);
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        bulk_s2g(genv + (size_t)(rd * 32 + l0) * OB, buf, (uint32_t)(min(G, na - l0) * OB));
                        bulk_commit();
                    }
                }
            }
        }
        __syncwarp();   // img / pre / suf are rewritten by the next env
    }
    if (lane == 0) bulk_wait_all();
}

// RgbObservation.observation (wrappers/observation.py:38-74): one warp per env, everything per-warp in shared memory:
//   record (prefetched with cp.async, double buffered) -> id image [Hp][RW] (bedrock / ones written once per warp; cells +
//   queue + holder + active piece per env) -> RGB bytes through a 256-entry PAIR table (two ids -> six colour bytes; eight
//   pixels = one 64-bit read -> four table reads -> three 64-bit writes) -> one TMA bulk store per env.
// HBM bytes per env: record + hot read, Hp * RW * 3 written once.  Bound: HBM write bandwidth.
// 20-wide rows at a 4-byte aligned image row: 2 playfield rows (5 words of the id plane) per call
__device__ __forceinline__ void fill_rows2_w20_strided(const uint32_t* ids5, uint32_t* o0, uint32_t* o1) {
    uint32_t w0 = ids5[0], w1 = ids5[1], w2 = ids5[2], w3 = ids5[3], w4 = ids5[4];
    uint32_t a0, a1, a2, a3, a4, ax, b0, b1, b2, b3, b4, bx;
    nib8_to_bytes(w0, a0, a1);
    nib8_to_bytes(w1, a2, a3);
    nib8_to_bytes(w2 & 0xFFFFu, a4, ax);
    nib8_to_bytes(__funnelshift_r(w2, w3, 16), b0, b1);
    nib8_to_bytes(__funnelshift_r(w3, w4, 16), b2, b3);
    nib8_to_bytes(w4 >> 16, b4, bx);
    (void)ax; (void)bx;
    o0[0] = a0; o0[1] = a1; o0[2] = a2; o0[3] = a3; o0[4] = a4;
    o1[0] = b0; o1[1] = b1; o1[2] = b2; o1[3] = b3; o1[4] = b4;
}

template <class COLT>
__global__ void __launch_bounds__(256, 3) k_rgb(const DevCfg cfg, int64_t n, const uint8_t* hot, const uint8_t* board, uint8_t* img,
                                             int rec_bytes, int pix_bytes, int rgb_bytes) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint32_t s_lut[16];
    __shared__ __align__(8) uint2 s_pair[256];
    __shared__ uint32_t s_rowbytes[112];
    const int W = cfg.W, H = cfg.H, Wp = cfg.Wp, Hp = cfg.Hp, RW = cfg.rgb_w, Q = cfg.Q, BS = cfg.board_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int NP = Hp * RW;
    uint8_t* wbase = sm + (size_t)warp * (2 * rec_bytes + pix_bytes + rgb_bytes);
    uint8_t* recbuf = wbase;               // two buffers: record (BS bytes) + hot record (32 B)
    uint8_t* pix = wbase + 2 * rec_bytes;
    uint8_t* rgb = pix + pix_bytes;
    if (threadIdx.x < 16) s_lut[threadIdx.x] = ((const uint32_t*)c_colors)[threadIdx.x];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {   // pixel pair (lo nibble, hi nibble) -> R0 G0 B0 R1 | G1 B1
        uint32_t c0 = ((const uint32_t*)c_colors)[i & 15] & 0xFFFFFFu, c1 = ((const uint32_t*)c_colors)[i >> 4] & 0xFFFFFFu;
        s_pair[i] = make_uint2(c0 | (c1 << 24), c1 >> 8);
    }
    for (int i = threadIdx.x; i < 112; i += blockDim.x) s_rowbytes[i] = (&c_rowbytes[0][0][0])[i];
    // constant part of the id image: everything that is not a playfield cell, a queue cell or a holder cell is 1
    for (int i = lane; i < NP; i += 32) {
        int r = i / RW, c = i - r * RW;
        pix[i] = (c < Wp && r < H && c >= P && c < P + W) ? 0 : 1;
    }
    for (int i = lane; i < 2 * rec_bytes / 4; i += 32) ((uint32_t*)recbuf)[i] = 0;
    __syncthreads();
    const bool fast8 = (NP & 7) == 0, fast4 = (NP & 3) == 0;
    const bool tma = ((NP * 3) & 15) == 0 && (((uintptr_t)img) & 15) == 0;
    const bool rows20 = W == 20 && (H & 1) == 0 && (RW & 3) == 0;
    const int64_t stride = (int64_t)gridDim.x * nwarps;
    int64_t e = (int64_t)blockIdx.x * nwarps + warp;
    auto prefetch = [&](int64_t ee, uint8_t* dst) {
        const uint8_t* src = board + ee * BS;
        for (int i = lane; i < (BS >> 4); i += 32) cp_async16(dst + 16 * i, src + 16 * i);
        if (lane < 2) cp_async16(dst + BS + 16 * lane, hot + ee * 32 + 16 * lane);
        cp_async_commit();
    };
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch, see k_step_ws
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (e < n) prefetch(e, recbuf);
    for (int it = 0; e < n; e += stride, it++) {
        const uint32_t* rec = (const uint32_t*)(recbuf + (it & 1) * rec_bytes);
        cp_async_wait_all();
        __syncwarp();
        if (e + stride < n) prefetch(e + stride, recbuf + ((it + 1) & 1) * rec_bytes);
        Hot h;
        hot_load(h, rec + (BS >> 2));
        const COLT* cols = (const COLT*)rec;
        const uint32_t* ids = rec + cfg.ids_off / 4;
        // board rows (cells only; the frame persists)
        if (rows20) {
            for (int g2 = lane; g2 < (H >> 1); g2 += 32)
                fill_rows2_w20_strided(ids + 5 * g2, (uint32_t*)(pix + (2 * g2) * RW + P), (uint32_t*)(pix + (2 * g2 + 1) * RW + P));
        } else {
            for (int r = lane; r < H; r += 32) fill_board_row<0>(cfg, ids, pix, 0, r, RW);
        }
        // queue (top right) and holder (bottom right)
        for (int q = lane; q < Q; q += 32) {
            const uint4 rb = *(const uint4*)(s_rowbytes + ((int)((h.queue >> (4 * q)) & 15u)) * 16);
            const uint32_t wv[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint8_t* d = pix + i * RW + Wp + 4 * q;
                if ((RW & 3) == 0 && (Wp & 3) == 0) *(uint32_t*)d = wv[i];
                else { d[0] = (uint8_t)wv[i]; d[1] = (uint8_t)(wv[i] >> 8); d[2] = (uint8_t)(wv[i] >> 16); d[3] = (uint8_t)(wv[i] >> 24); }
            }
        }
        if (lane >= 16 && ((lane - 16) >> 2) < cfg.holder_size) {   // up to four held pieces side by side, four rows each
            const int s = (lane - 16) >> 2, i = lane & 3;
            const uint32_t wv = holder_row(cfg, h, s_rowbytes, s, i);
            uint8_t* d = pix + (Hp - P + i) * RW + Wp + 4 * s;
            d[0] = (uint8_t)wv; d[1] = (uint8_t)(wv >> 8); d[2] = (uint8_t)(wv >> 16); d[3] = (uint8_t)(wv >> 24);
        }
        __syncwarp();
        uint32_t cells = c_cells[h.p][h.r];
        COLT B = bmask<COLT>(cols, W, cells, h.x);
        if (!((B >> h.y) & 1) && lane < 4) {   // active piece on top (project_tetromino, envs/tetris.py:543-564)
            int c = (cells >> (4 * lane)) & 15;
            pix[(h.y + (c >> 2)) * RW + h.x + (c & 3)] = (uint8_t)(h.p + 2);
        }
        // previous env's bulk store must have finished reading this warp's rgb buffer
        if (lane == 0) bulk_wait_read();
        __syncwarp();
        uint8_t* g = img + (size_t)e * NP * 3;
        if (fast8 && tma) {
            for (int q8 = lane; q8 < (NP >> 3); q8 += 32) {
                const uint2 pv = ((const uint2*)pix)[q8];
                const uint32_t t0 = (pv.x | (pv.x >> 4)) << 3, t1 = (pv.y | (pv.y >> 4)) << 3;
                const uint2 A = *(const uint2*)((const uint8_t*)s_pair + (t0 & 0x7F8u));
                const uint2 Bp = *(const uint2*)((const uint8_t*)s_pair + ((t0 >> 16) & 0x7F8u));
                const uint2 C = *(const uint2*)((const uint8_t*)s_pair + (t1 & 0x7F8u));
                const uint2 D = *(const uint2*)((const uint8_t*)s_pair + ((t1 >> 16) & 0x7F8u));
                uint2* o = (uint2*)(rgb + 24 * q8);
                o[0] = make_uint2(A.x, __byte_perm(A.y, Bp.x, 0x5410));
                o[1] = make_uint2(__byte_perm(Bp.x, Bp.y, 0x5432), C.x);
                o[2] = make_uint2(__byte_perm(C.y, D.x, 0x5410), __byte_perm(D.x, D.y, 0x5432));
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { bulk_s2g(g, rgb, (uint32_t)(NP * 3)); bulk_commit(); }
        } else if (fast4) {
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
    ON_DEVICE(env);
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
    ON_DEVICE(env);
    const DevCfg& d = env->dev;
    auto r128 = [](size_t v) { return (int)((v + 127) / 128 * 128); };
    int rec_bytes = r128((size_t)d.board_stride + 48), pix_bytes = r128((size_t)d.Hp * d.rgb_w + 16), rgb_bytes = r128((size_t)d.Hp * d.rgb_w * 3);
    size_t per_warp = (size_t)2 * rec_bytes + pix_bytes + rgb_bytes;
    int nw = 8, best = 0;
    for (int c = 8; c >= 1; c >>= 1) {   // warps per CTA that keeps the most warps resident (227 KB shared memory, <= 24 warps by registers)
        int resident = (int)((227 * 1024) / (per_warp * c + 4096)) * c;
        if (resident > 24) resident = 24;
        if (resident > best) { best = resident; nw = c; }
    }
    size_t smem = per_warp * nw;
    if (smem > 227 * 1024) return fail(env, TG_ERR_CONFIG, "tg_render_rgb: image too large for shared memory");
    int T = nw * 32;
    auto launch = [&](auto kern) -> int {
        CUDA_TRY(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        CUDA_TRY(env, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem));
        int64_t blocks = (n + nw - 1) / nw, cap = (int64_t)env->num_sms * (per_sm > 0 ? per_sm : 1);
        if (blocks > cap) blocks = cap;
        CUDA_TRY(env, launch_pdl(kern, (unsigned)blocks, (unsigned)T, smem, (cudaStream_t)stream, d, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_img, rec_bytes, pix_bytes, rgb_bytes));
        return TG_OK;
    };
    return env->col64 ? launch(k_rgb<uint64_t>) : launch(k_rgb<uint32_t>);
}

// packed-byte kernel (tg_gfeats.cuh) for the two board widths of BASELINE.json; the generic kernel covers the rest
static bool gfeats_fast(const tg_env* env, const uint8_t* d_feats, const uint8_t* d_legal) {
    return d_feats && (env->dev.W == 10 || env->dev.W == 20) && (((uintptr_t)d_feats | (uintptr_t)d_legal) & 15) == 0 && !getenv("TG_GFEATS_V1");
}

// d_info_board (nullable): info["board"] computed by the packed-byte kernel (only passed when gfeats_fast() holds)
static int launch_grouped_observe(tg_env* env, tg_state st, int64_t n, uint8_t* d_feats, uint8_t* d_boards, uint8_t* d_legal,
                                  const uint8_t* fill_high, cudaStream_t s, uint8_t* d_info_board = nullptr) {
    const DevCfg& d = env->dev;
    const bool fast_x = gfeats_fast(env, d_feats, d_legal);
    if (fast_x) {
        auto launch = [&](auto kern, size_t smem, int T) -> int {
            CUDA_TRY(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CUDA_TRY(env, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            if (n >= 16384 && !getenv("TG_NO_PDL")) {   // (tiny batches are bound by the host-side launch cost, which the extended launch raises)
                cudaLaunchConfig_t lc;
                memset(&lc, 0, sizeof lc);
                lc.gridDim = dim3((unsigned)((n + 31) / 32)); lc.blockDim = dim3((unsigned)T); lc.dynamicSmemBytes = smem; lc.stream = s;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                lc.attrs = at; lc.numAttrs = 1;
                CUDA_TRY(env, cudaLaunchKernelEx(&lc, kern, d, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_feats, d_legal, fill_high, d_info_board));
            } else
            kern<<<(unsigned)((n + 31) / 32), T, smem, s>>>(d, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_feats, d_legal, fill_high, d_info_board);
            CUDA_TRY(env, cudaGetLastError());
            return TG_OK;
        };
        int rc;
        if (d.W == 10) rc = env->col64 ? launch(k_grouped_feats_x<10, uint64_t>, GFeatsSmem<10, uint64_t>::bytes, 320)
                                       : launch(k_grouped_feats_x<10, uint32_t>, GFeatsSmem<10, uint32_t>::bytes, 320);
        else rc = env->col64 ? launch(k_grouped_feats_x<20, uint64_t>, GFeatsSmem<20, uint64_t>::bytes, 640)
                             : launch(k_grouped_feats_x<20, uint32_t>, GFeatsSmem<20, uint32_t>::bytes, 640);
        if (rc) return rc;
    } else if (d_feats) {
        int EPB = 32, T = 256;             // 32 envs x 4W placements = a whole number of 256-thread rounds
        size_t colb = env->col64 ? 8 : 4;
        size_t smem = (size_t)EPB * (3 * d.W + 2 * TG_PADDING) * colb + (size_t)EPB * 16 + (size_t)EPB * 4 + (size_t)EPB * 128 + (size_t)EPB * d.A * d.F + (size_t)EPB * d.A;
        // always opt in: static (tables, slow list) + dynamic shared memory may exceed 48 KB even when the dynamic part does not
        if (env->col64) CUDA_TRY(env, cudaFuncSetAttribute(k_grouped_feats<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else CUDA_TRY(env, cudaFuncSetAttribute(k_grouped_feats<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        unsigned g = (unsigned)((n + EPB - 1) / EPB);
        const uint32_t magicA = ((1u << 20) + d.A - 1) / d.A;   // it / A == (it * magicA) >> 20, exact for it < 32 * A, A = 4W <= 96 (enumerated)
        if (env->col64) k_grouped_feats<uint64_t><<<g, T, smem, s>>>(d, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_feats, d_legal, fill_high, EPB, magicA);
        else k_grouped_feats<uint32_t><<<g, T, smem, s>>>(d, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_feats, d_legal, fill_high, EPB, magicA);
        CUDA_TRY(env, cudaGetLastError());
    }
    const bool stream_ok = d_boards && (d.OB & 15) == 0 && (((uintptr_t)d_boards) & 15) == 0 && d.OB <= 4 * 32 * 16 && !getenv("TG_BOARDS_V1");
    if (stream_ok) {
        const int NV = (d.OB / 16 + 31) / 32;
        auto r128 = [](size_t v) { return (int)((v + 127) / 128 * 128); };
        const int G = NV == 1 ? 8 : (NV <= 3 ? 4 : 2);   // placements per bulk store (matches the kernel template)
        const int rec_bytes = r128((size_t)d.board_stride + 32), img_bytes = r128((size_t)d.OB), gbuf_bytes = r128((size_t)G * d.OB);
        const size_t per_warp = (size_t)2 * rec_bytes + img_bytes + 512 + 2 * (size_t)gbuf_bytes;
        int nw = 8, best = 0;
        for (int c = 8; c >= 2; c >>= 1) {           // warps per CTA that keeps the most warps resident
            int resident = (int)((227 * 1024) / (per_warp * c + 2048)) * c;
            if (resident > 24) resident = 24;
            if (resident > best) { best = resident; nw = c; }
        }
        const int T = nw * 32;
        const size_t smem = (size_t)nw * per_warp;
        auto launch = [&](auto kern) -> int {
            CUDA_TRY(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 1;
            CUDA_TRY(env, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem));
            int64_t blocks = (n + nw - 1) / nw, cap = (int64_t)env->num_sms * (per_sm > 0 ? per_sm : 1);
            if (blocks > cap) blocks = cap;
            CUDA_TRY(env, launch_pdl(kern, (unsigned)blocks, (unsigned)T, smem, s, d, n, (const uint8_t*)st.hot, (const uint8_t*)st.board, d_boards, d_legal, fill_high, rec_bytes, img_bytes, gbuf_bytes));
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
    ON_DEVICE(env);
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
    ON_DEVICE(env);
    rc = ensure_stage(env, 3, (size_t)n); if (rc) return rc;  // per-env "illegal + terminate" flags
    StepParams p;
    memset(&p, 0, sizeof p);
    p.n = n; p.hot = (uint8_t*)st.hot; p.board = (uint8_t*)st.board; p.rng = (uint8_t*)st.rng; p.seq = st.piece_seq;
    p.actions = d_actions;
    p.o_board = obs.board; p.o_mask = obs.mask; p.o_holder = obs.holder; p.o_queue = obs.queue;
    p.reward = out.reward; p.terminated = out.terminated; p.truncated = out.truncated; p.lines = out.lines;
    p.stats = (double*)d_stats;
    // with the packed-byte feature kernel info["board"] is a by-product of its column pass, not of the step kernel
    const bool info_in_feats = d_info_board && ((uintptr_t)d_info_board & 15) == 0 && gfeats_fast(env, d_feats, d_legal) && !getenv("TG_INFO_IN_STEP");
    p.legal = d_legal; p.info_board = info_in_feats ? nullptr : d_info_board; p.fill_high = (uint8_t*)env->stage[3];
    p.mode = 2;
    // Feature observation of the reference board on 32-bit columns, small batches: ONE persistent kernel applies the placement and
    // enumerates the next state's placements (k_grouped_step_feats, tg_gfeats.cuh) -- one launch and one pass over the records
    // instead of two (4,096 envs: 16 instead of 20 us per step).  From 64 K envs on the two-kernel path is faster (1 M envs: 505
    // against 609 us: the fused kernel's code no longer fits the instruction cache next to 40 feature warps per SM).
    // TG_GROUPED_SPLIT=1 / TG_GROUPED_FUSED=1 force one or the other.
    const bool fused_ok = !any_obs && !d_boards && d_feats && env->dev.W == 10 && !env->col64 && gfeats_fast(env, d_feats, d_legal) &&
                          (!d_info_board || info_in_feats);
    if (fused_ok && !getenv("TG_GROUPED_SPLIT") && (n < 65536 || getenv("TG_GROUPED_FUSED"))) {
        const DevCfg& d = env->dev;
        const bool xt = d.NPC != 7 || d.holder_size > 1;
        typedef void (*fused_t)(const StepParams, uint8_t*, uint8_t*, uint8_t*);
        fused_t kern = xt ? (fused_t)k_grouped_step_feats<10, uint32_t, true> : (fused_t)k_grouped_step_feats<10, uint32_t, false>;
        const size_t smem = GFusedSmem<10, uint32_t>::bytes(d);
        rc = raise_smem_limit(env, (void*)kern, smem); if (rc) return rc;
        if (env->fused_grid_max <= 0) {
            int per_sm = 0;
            CUDA_TRY(env, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CUDA_TRY(env, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GFusedSmem<10, uint32_t>::threads, smem));
            if (per_sm < 1) return fail(env, TG_ERR_CONFIG, "fused grouped kernel does not fit: %zu B shared memory per CTA", smem);
            env->fused_grid_max = (int64_t)env->num_sms * per_sm;
        }
        p.cfg = d; p.E = 32; p.whole_tile_min = 16;
        if (const char* t = getenv("TG_WHOLE")) p.whole_tile_min = atoi(t);
        const int64_t ntiles = (n + 31) / 32;
        const int64_t grid = env->fused_grid_max < ntiles ? env->fused_grid_max : ntiles;
        CUDA_TRY(env, launch_pdl(kern, (unsigned)grid, (unsigned)GFusedSmem<10, uint32_t>::threads, smem, (cudaStream_t)stream, p, d_feats, d_legal, d_info_board));
        return TG_OK;
    }
    rc = launch_step(env, p, (cudaStream_t)stream); if (rc) return rc;
    if (d_feats || d_boards)
        return launch_grouped_observe(env, st, n, d_feats, d_boards, d_legal, (const uint8_t*)env->stage[3], (cudaStream_t)stream,
                                      info_in_feats ? d_info_board : nullptr);
    return TG_OK;
}

extern "C" int tg_rollout(tg_env* env, tg_state st, int64_t n, const int32_t weights[4], int32_t k_steps, tg_stats* d_stats,
                          void* stream) {
    if (!env) return TG_ERR_POINTER;
    if (n <= 0 || k_steps < 0 || !weights) return fail(env, TG_ERR_ARG, "tg_rollout: bad argument");
    int rc = check_state(env, st); if (rc) return rc;
    ON_DEVICE(env);
    const DevCfg& d = env->dev;
    RolloutParams p;
    memset(&p, 0, sizeof p);
    p.cfg = d; p.n = n; p.hot = (uint8_t*)st.hot; p.board = (uint8_t*)st.board; p.rng = (uint8_t*)st.rng; p.seq = st.piece_seq;
    for (int i = 0; i < 4; i++) p.w[i] = weights[i];
    p.k_steps = k_steps; p.stats = (double*)d_stats;
    p.last_action = (int32_t*)env->rollout_last_action;
    const int cw = env->col64 ? 2 : 1;                    // words per column
    int words = (d.W + 2 * TG_PADDING) * cw;              // P wall columns | W columns | P wall columns
    p.ids_off_g = d.ids_off;
    p.cfg.ids_off = (d.W + TG_PADDING) * cw * 4;   // in-slot offset of the id plane
    words += (d.board_stride - d.ids_off) / 4 + 1;        // id plane (+1 word: ids_get8 may read one word past it)
    if (env->col64) words = (words + 1) & ~1;             // keep the EnvBase arrays 8-byte aligned
    p.base_off = words - TG_PADDING * cw;                 // measured from the first column
    p.hb = (d.W + 3) & ~3;
    const bool packed = !getenv("TG_ROLLOUT_V1") && (d.W == 10 || d.W == 20);   // k_rollout_x: packed heights instead of h / ho / bs
    words += 2 * d.W * cw + (packed ? (d.W + 2 * TG_PADDING + 3) / 4 : 4 * p.hb / 4);   // pre[W], suf[W], then h, ho, bs (u16) | packed heights
    // u32 columns: odd word stride; u64 columns: stride = 2 (mod 4) words keeps 8-byte alignment and spreads the banks
    if (env->col64) { while ((words & 3) != 2) words += 1; } else { words |= 1; }
    p.rec_words = words;
    // CTA size: the largest of 128 / 64 / 32 threads that keeps the most env slots resident per SM (big boards: a 20x40
    // slot is ~1 KB, one 128-thread CTA would be alone on its SM)
    int T = 128;
    {
        size_t best = 0;
        for (int t = 128; t >= 32; t >>= 1) {
            const size_t per_cta = (size_t)t * words * 4 + 1024;
            size_t ctas = (227 * 1024) / per_cta;
            if (ctas > 32) ctas = 32;
            if (ctas * t > 768) ctas = 768 / t;          // 80 registers per thread: at most 768 threads per SM
            if (ctas * t > best) { best = ctas * t; T = t; }
        }
    }
    size_t smem = (size_t)T * words * 4;
    auto launch = [&](auto kern) -> int {
        CUDA_TRY(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)((n + T - 1) / T), T, smem, (cudaStream_t)stream>>>(p);
        CUDA_TRY(env, cudaGetLastError());
        return TG_OK;
    };
    if (smem > 227 * 1024) return fail(env, TG_ERR_CONFIG, "tg_rollout: board record too large for shared memory");
    if (packed) {   // packed-byte variant for the two board widths of BASELINE.json
        if (d.W == 10) return env->col64 ? launch(k_rollout_x<10, uint64_t>) : launch(k_rollout_x<10, uint32_t>);
        if (d.W == 20) return env->col64 ? launch(k_rollout_x<20, uint64_t>) : launch(k_rollout_x<20, uint32_t>);
    }
    return env->col64 ? launch(k_rollout<uint64_t>) : launch(k_rollout<uint32_t>);
}

/* test hook: device buffer (i32[n]) that receives the action chosen at the last rollout step; NULL disables */
extern "C" int tg_debug_set_rollout_trace(tg_env* env, int32_t* d_last_action) {
    if (!env) return TG_ERR_POINTER;
    env->rollout_last_action = d_last_action;
    return TG_OK;
}
