// tg_cnn.cuh -- fused image adapter of the CNN trainer (examples/train_cnn.py:127-147):
//     RgbObservation -> gym.wrappers.ResizeObservation((OH, OW)) -> gym.wrappers.GrayscaleObservation
// in ONE kernel: the RGB image never touches HBM, only the OH x OW grey frame (84 x 84 = 7 KB) is written, straight into
// the caller's frame-stack window (FrameStackObservation: the reset frame is replicated over the window).
//   * RgbObservation.observation            wrappers/observation.py:38-74 (id image -> colours)
//   * ResizeObservation = cv2.resize(INTER_AREA); with an enlarged axis OpenCV emulates it by a bilinear kernel with
//     `area_mode` coordinates, fixed point with 11-bit coefficients (imgproc/src/resize.cpp: HResizeLinear / VResizeLinear);
//     offsets / coefficients are computed on the host exactly like cv::hal::resize (tg_api.cu: cnn_axis_tables)
//   * GrayscaleObservation = floor(R * 0.2125 + G * 0.7154 + B * 0.0721) in float64, summed left to right.  Evaluated in
//     integers: N = 2125 R + 7154 G + 721 B, grey = N / 10000, minus one for the 63 (R, G, B) triples (all with
//     N % 10000 == 0) where the float64 sum lands just below the integer; the host finds them by enumerating all 2^24
//     triples with the float64 expression itself (tg_api.cu: cnn_gray_exceptions) -- at most two per quotient
// Third-party arithmetic (gymnasium 1.1.1, OpenCV): restated in oracle/cnn_obs_oracle.py and pinned against cv2 there.
// One warp per env: record prefetch (cp.async) -> id image in shared memory -> per lane fixed output columns (source
// offsets / coefficients in registers), horizontal pass kept in registers for the two live source rows, vertical pass +
// grey conversion per output row -> frame in shared memory -> TMA bulk store(s).
// Bound: integer issue (about 40 thread-instructions per output pixel), not HBM (7 KB written per env).
#pragma once

namespace tg {

struct CnnParams {
    DevCfg cfg;
    int64_t n;
    const uint8_t* hot; const uint8_t* board;
    const int32_t* xtab;      // [OW][4]: sx0, sx1 (pixel offsets inside an image row), a0, a1
    const int32_t* ytab;      // [OH][4]: sy0, sy1 (clamped source rows), b0, b1
    const uint32_t* gray_exc; // [256][2] by quotient: R | G << 8 | B << 16 of up to two exception triples (0xFFFFFFFF = none)
    uint8_t* frames;          // frame of env e at frames + e * env_stride
    int64_t env_stride;
    const uint8_t* fill_mask; // nullable: envs whose frame is replicated into the `fill_count` preceding frames
    int fill_count;
    int OH, OW;
    int rec_bytes, pix_bytes, out_bytes;
    // k_cnn_obs2: grey value of a pixel whose four source pixels hold the same id: gtab[(dy * 4 + class of a0 + a1) * 16 + id]
    const uint8_t* gtab;
    int axv[4];               // the (up to four) distinct sums a0 + a1 of the x coefficient pairs; class = index in here
    int list_bytes;           // per-warp list of the pixels that need the full interpolation
    int chunk_rows;           // k_cnn_obs2: output rows per staged chunk (two chunk buffers of out_bytes each per warp)
    int gtab_bytes;           // k_cnn_obs2: bytes of the table at the front of the dynamic shared memory
    // k_cnn_obs3 (word-wise first pass)
    const uint32_t* gtabT;    // [OH][16] words: byte c of gtabT[dy * 16 + id] = gtab[(dy * 4 + c) * 16 + id] (the four classes of a0 + a1)
    const int32_t* wtab;      // [OW / 4][4] per output word: first source column, byte mask of its source columns, class selector (PRMT), -
    const uint8_t* xcls;      // [OW] class of a0 + a1 per output column
    int wlist_bytes;          // per-warp list of the output words whose source pixels are not all one id
};

template <class COLT, int NX>
__global__ void __launch_bounds__(128) k_cnn_obs(const __grid_constant__ CnnParams p) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint32_t s_lut[16];
    __shared__ uint32_t s_rowbytes[112];
    __shared__ __align__(16) int4 s_y[128];
    __shared__ __align__(8) uint2 s_exc[256];
    const DevCfg& cfg = p.cfg;
    const int W = cfg.W, H = cfg.H, Wp = cfg.Wp, Hp = cfg.Hp, RW = cfg.rgb_w, Q = cfg.Q, BS = cfg.board_stride;
    const int OH = p.OH, OW = p.OW, FB = OH * OW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int NP = Hp * RW;
    uint8_t* wbase = sm + (size_t)warp * (2 * p.rec_bytes + p.pix_bytes + p.out_bytes);
    uint8_t* recbuf = wbase;
    uint8_t* pix = wbase + 2 * p.rec_bytes;
    uint8_t* out = pix + p.pix_bytes;
    if (threadIdx.x < 16) s_lut[threadIdx.x] = ((const uint32_t*)c_colors)[threadIdx.x];
    for (int i = threadIdx.x; i < 112; i += blockDim.x) s_rowbytes[i] = (&c_rowbytes[0][0][0])[i];
    for (int i = threadIdx.x; i < OH; i += blockDim.x) s_y[i] = ((const int4*)p.ytab)[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_exc[i] = ((const uint2*)p.gray_exc)[i];
    const uint32_t exc_addr = smem_u32(s_exc);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch, see k_step_ws
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = lane; i < NP; i += 32) {   // constant part of the id image (see k_rgb)
        int r = i / RW, c = i - r * RW;
        pix[i] = (c < Wp && r < H && c >= P && c < P + W) ? 0 : 1;
    }
    for (int i = lane; i < 2 * p.rec_bytes / 4; i += 32) ((uint32_t*)recbuf)[i] = 0;
    // this lane's output columns: dx = lane + 32 j
    int sx0[NX], sx1[NX];
    uint32_t acoef[NX];            // a0 | a1 << 16 (11-bit coefficients, 0 <= a <= 2048)
#pragma unroll
    for (int j = 0; j < NX; j++) {
        const int dx = lane + 32 * j;
        int4 t = dx < OW ? ((const int4*)p.xtab)[dx] : make_int4(0, 0, 0, 0);
        sx0[j] = t.x; sx1[j] = t.y; acoef[j] = (uint32_t)t.z | ((uint32_t)t.w << 16);
    }
    __syncthreads();
    const bool tma = (FB & 15) == 0 && (((uintptr_t)p.frames) & 15) == 0 && (p.env_stride & 15) == 0;
    const bool rows20 = W == 20 && (H & 1) == 0 && (RW & 3) == 0;
    const int64_t stride = (int64_t)gridDim.x * nwarps;
    int64_t e = (int64_t)blockIdx.x * nwarps + warp;
    auto prefetch = [&](int64_t ee, uint8_t* dst) {
        const uint8_t* src = p.board + ee * BS;
        for (int i = lane; i < (BS >> 4); i += 32) cp_async16(dst + 16 * i, src + 16 * i);
        if (lane < 2) cp_async16(dst + BS + 16 * lane, p.hot + ee * 32 + 16 * lane);
        cp_async_commit();
    };
    if (e < p.n) prefetch(e, recbuf);
    for (int it = 0; e < p.n; e += stride, it++) {
        const uint32_t* rec = (const uint32_t*)(recbuf + (it & 1) * p.rec_bytes);
        cp_async_wait_all();
        __syncwarp();
        if (e + stride < p.n) prefetch(e + stride, recbuf + ((it + 1) & 1) * p.rec_bytes);
        Hot h;
        hot_load(h, rec + (BS >> 2));
        const COLT* cols = (const COLT*)rec;
        const uint32_t* ids = rec + cfg.ids_off / 4;
        // ---- id image (RgbObservation layout: board | queue top right, holder bottom right) ----
        if (rows20) {
            for (int g2 = lane; g2 < (H >> 1); g2 += 32)
                fill_rows2_w20_strided(ids + 5 * g2, (uint32_t*)(pix + (2 * g2) * RW + P), (uint32_t*)(pix + (2 * g2 + 1) * RW + P));
        } else {
            for (int r = lane; r < H; r += 32) fill_board_row<0>(cfg, ids, pix, 0, r, RW);
        }
        for (int q = lane; q < Q; q += 32) {
            const uint4 rb = *(const uint4*)(s_rowbytes + ((int)((h.queue >> (4 * q)) & 15u)) * 16);
            const uint32_t wv[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint8_t* d = pix + i * RW + Wp + 4 * q;
                d[0] = (uint8_t)wv[i]; d[1] = (uint8_t)(wv[i] >> 8); d[2] = (uint8_t)(wv[i] >> 16); d[3] = (uint8_t)(wv[i] >> 24);
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
        bulk_wait_read();                  // this lane's stores of the previous env have read `out`
        __syncwarp();
        // ---- resize (bilinear kernel, area-mode coordinates) + grey, one output row at a time ----
        int hc[NX][3], hn[NX][3];          // horizontal pass of the two live source rows, this lane's columns
        int row_c = -1, row_n = -1;
        auto hpass = [&](int sy, int (&dst)[NX][3]) {
            const uint8_t* prow = pix + sy * RW;
#pragma unroll
            for (int j = 0; j < NX; j++) {
                const uint32_t c0 = s_lut[prow[sx0[j]]], c1 = s_lut[prow[sx1[j]]];
                // c0 * a0 + c1 * a1 per channel as 16-bit x 8-bit dot products (IDP.2A): bytes (R0, R1, G0, G1) / (B0, B1, -, -);
                // kept pre-shifted: the vertical pass only uses h >> 4
                const uint32_t rg = __byte_perm(c0, c1, 0x5140), bb = __byte_perm(c0, c1, 0x7762);
                dst[j][0] = (int)(__dp2a_lo(acoef[j], rg, 0u) >> 4);
                dst[j][1] = (int)(__dp2a_hi(acoef[j], rg, 0u) >> 4);
                dst[j][2] = (int)(__dp2a_lo(acoef[j], bb, 0u) >> 4);
            }
        };
        uint32_t orow = smem_u32(out) + lane;   // shared-window address of this lane's first pixel of the output row
        for (int dy = 0; dy < OH; dy++, orow += OW) {
            const int4 yt = s_y[dy];       // sy0, sy1, b0, b1 (warp-uniform)
            if (yt.x != row_c) {
                if (yt.x == row_n) {
#pragma unroll
                    for (int j = 0; j < NX; j++) { hc[j][0] = hn[j][0]; hc[j][1] = hn[j][1]; hc[j][2] = hn[j][2]; }
                } else hpass(yt.x, hc);
                row_c = yt.x;
            }
            if (yt.y != row_n) {
                if (yt.y == row_c) {
#pragma unroll
                    for (int j = 0; j < NX; j++) { hn[j][0] = hc[j][0]; hn[j][1] = hc[j][1]; hn[j][2] = hc[j][2]; }
                } else hpass(yt.y, hn);
                row_n = yt.y;
            }
#pragma unroll
            for (int j = 0; j < NX; j++) {
                // colours are <= 240 and the coefficient pairs sum to 2048 +- 1: 0 <= v <= 255 without clamping;
                // VResizeLinear: (((b0 * h0) >> 16) + ((b1 * h1) >> 16) + 2) >> 2
                const uint32_t r = (uint32_t)((((yt.z * hc[j][0] + 0x20000) >> 16) + ((yt.w * hn[j][0]) >> 16)) >> 2);
                const uint32_t g = (uint32_t)((((yt.z * hc[j][1] + 0x20000) >> 16) + ((yt.w * hn[j][1]) >> 16)) >> 2);
                const uint32_t b = (uint32_t)((((yt.z * hc[j][2] + 0x20000) >> 16) + ((yt.w * hn[j][2]) >> 16)) >> 2);
                const uint32_t N = r * 2125u + g * 7154u + b * 721u;
                uint32_t q = (uint32_t)(((uint64_t)N * 3518437209ull) >> 45);   // N / 10000 for N <= 2,550,000
                // (a branch on N == q * 10000 -- the only candidates -- measured 20 % slower than the unconditional lookup,
                //  and (b * h) >> 16 as IMAD.HI 15 % slower than multiply + shift)
                const uint32_t key = r | (g << 8) | (b << 16);
                uint32_t t0, t1;
                asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(t0), "=r"(t1) : "r"(exc_addr + q * 8u));
                q -= (uint32_t)(key == t0) | (uint32_t)(key == t1);
                if (lane + 32 * j < OW) asm volatile("st.shared.u8 [%0], %1;" ::"r"(orow + 32u * j), "r"(q) : "memory");
            }
        }
        // ---- store: the frame, plus (reset envs) the preceding frames of the stack window ----
        uint8_t* g = p.frames + e * p.env_stride;
        const int reps = 1 + ((p.fill_mask && p.fill_mask[e]) ? p.fill_count : 0);
        if (tma) {
            fence_async_smem();
            __syncwarp();
            if (lane < reps) { bulk_s2g(g - (size_t)lane * FB, out, (uint32_t)FB); bulk_commit(); }
            for (int r = 32 + lane; r < reps; r += 32) { bulk_s2g(g - (size_t)r * FB, out, (uint32_t)FB); bulk_commit(); }
        } else {
            __syncwarp();
            for (int r = 0; r < reps; r++)
                for (int i = lane; i < FB; i += 32) (g - (size_t)r * FB)[i] = out[i];
        }
        __syncwarp();
    }
    bulk_wait_all();
}

// ---- k_cnn_obs2: the same frame, two passes ---------------------------------------------------------------------------------
// An output pixel interpolates a 2 x 2 neighbourhood of the id image, and nearly all neighbourhoods hold ONE id (empty field,
// bedrock frame, the inside of a piece: 96 % of the pixels of a wide board in play).  For those the whole chain -- horizontal
// pass, vertical pass, grey conversion -- is a function of (id, a0 + a1 of the column, output row): a 5 KB table built on the
// host with the same integer / float64 expressions (tg_cnn_observe: cnn_uniform_table).  Pass 1 (lane = output columns, row by
// row): compare the ids, write the table value or append the pixel to the warp's list (ballot compaction, no atomics).
// Pass 2: the listed pixels, 32 at a time, through the full fixed-point interpolation of k_cnn_obs.  A neighbour with a zero
// coefficient does not count (a1 = 0 at the right border, b1 = 0 where an output row sits on a source row).
template <class COLT, int NX>
__global__ void __launch_bounds__(256, 4) k_cnn_obs2(const __grid_constant__ CnnParams p) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint32_t s_lut[16];
    __shared__ uint32_t s_rowbytes[112];
    __shared__ __align__(16) int4 s_y[128];
    __shared__ __align__(8) uint2 s_exc[256];
    const DevCfg& cfg = p.cfg;
    const int W = cfg.W, H = cfg.H, Wp = cfg.Wp, Hp = cfg.Hp, RW = cfg.rgb_w, Q = cfg.Q, BS = cfg.board_stride;
    const int OH = p.OH, OW = p.OW, FB = OH * OW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int NP = Hp * RW;
    uint8_t* s_g = sm;                     // [OH][4][16] uniform-neighbourhood table, then one region per warp
    const int4* s_x = (const int4*)p.xtab; // (pass 2 only: read through L1)
    uint8_t* wbase = sm + p.gtab_bytes + (size_t)warp * (2 * p.rec_bytes + p.pix_bytes + 2 * p.out_bytes + p.list_bytes);
    uint8_t* recbuf = wbase;
    uint8_t* pix = wbase + 2 * p.rec_bytes;
    uint8_t* out0 = pix + p.pix_bytes;     // two chunk buffers: one is filled while the bulk store of the other drains
    unsigned short* list = (unsigned short*)(out0 + 2 * p.out_bytes);
    const int CR = p.chunk_rows;
    uint32_t nchunk = 0;                   // chunks staged by this warp so far (buffer = nchunk & 1)
    const int LCAP = p.list_bytes / 2;
    if (threadIdx.x < 16) s_lut[threadIdx.x] = ((const uint32_t*)c_colors)[threadIdx.x];
    for (int i = threadIdx.x; i < 112; i += blockDim.x) s_rowbytes[i] = (&c_rowbytes[0][0][0])[i];
    for (int i = threadIdx.x; i < OH; i += blockDim.x) s_y[i] = ((const int4*)p.ytab)[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_exc[i] = ((const uint2*)p.gray_exc)[i];
    for (int i = threadIdx.x; i < OH * 4; i += blockDim.x) ((uint4*)s_g)[i] = ((const uint4*)p.gtab)[i];
    const uint32_t exc_addr = smem_u32(s_exc);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch, see k_step_ws
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = lane; i < NP; i += 32) {   // constant part of the id image (see k_rgb)
        int r = i / RW, c = i - r * RW;
        pix[i] = (c < Wp && r < H && c >= P && c < P + W) ? 0 : 1;
    }
    for (int i = lane; i < 2 * p.rec_bytes / 4; i += 32) ((uint32_t*)recbuf)[i] = 0;
    // this lane's output columns: dx = lane + 32 j
    int sx0[NX], sx1[NX], gofs[NX];
#pragma unroll
    for (int j = 0; j < NX; j++) {
        const int dx = lane + 32 * j;
        int4 t = dx < OW ? ((const int4*)p.xtab)[dx] : make_int4(0, 0, 2048, 0);
        sx0[j] = t.x; sx1[j] = t.w == 0 ? t.x : t.y;          // a neighbour with a zero coefficient does not count
        const int ax = t.z + t.w;
        gofs[j] = 16 * ((ax == p.axv[1]) + 2 * (ax == p.axv[2]) + 3 * (ax == p.axv[3]));
    }
    __syncthreads();
    const bool tma = (FB & 15) == 0 && (((uintptr_t)p.frames) & 15) == 0 && (p.env_stride & 15) == 0;
    const bool rows20 = W == 20 && (H & 1) == 0 && (RW & 3) == 0;
    const int64_t stride = (int64_t)gridDim.x * nwarps;
    int64_t e = (int64_t)blockIdx.x * nwarps + warp;
    auto prefetch = [&](int64_t ee, uint8_t* dst) {
        const uint8_t* src = p.board + ee * BS;
        for (int i = lane; i < (BS >> 4); i += 32) cp_async16(dst + 16 * i, src + 16 * i);
        if (lane < 2) cp_async16(dst + BS + 16 * lane, p.hot + ee * 32 + 16 * lane);
        cp_async_commit();
    };
    // pass 2: the full interpolation of the listed pixels (entry = dy | dx << 8), 32 at a time
    auto flush = [&](int cnt, uint8_t* out, int row0) {
        __syncwarp();
        for (int k = lane; k < cnt; k += 32) {
            const int ent = list[k], dy = ent & 255, dx = ent >> 8;
            const int4 yt = s_y[dy], xt = __ldg(s_x + dx);
            const uint32_t acoef = (uint32_t)xt.z | ((uint32_t)xt.w << 16);
            const uint8_t* r0 = pix + yt.x * RW;
            const uint8_t* r1 = pix + yt.y * RW;
            int hc[3], hn[3];
            {
                const uint32_t c0 = s_lut[r0[xt.x]], c1 = s_lut[r0[xt.y]];
                const uint32_t rg = __byte_perm(c0, c1, 0x5140), bb = __byte_perm(c0, c1, 0x7762);
                hc[0] = (int)(__dp2a_lo(acoef, rg, 0u) >> 4); hc[1] = (int)(__dp2a_hi(acoef, rg, 0u) >> 4); hc[2] = (int)(__dp2a_lo(acoef, bb, 0u) >> 4);
            }
            {
                const uint32_t c0 = s_lut[r1[xt.x]], c1 = s_lut[r1[xt.y]];
                const uint32_t rg = __byte_perm(c0, c1, 0x5140), bb = __byte_perm(c0, c1, 0x7762);
                hn[0] = (int)(__dp2a_lo(acoef, rg, 0u) >> 4); hn[1] = (int)(__dp2a_hi(acoef, rg, 0u) >> 4); hn[2] = (int)(__dp2a_lo(acoef, bb, 0u) >> 4);
            }
            // VResizeLinear: (((b0 * h0) >> 16) + ((b1 * h1) >> 16) + 2) >> 2 (colours <= 240, pairs sum to 2048 +- 1: no clamping)
            const uint32_t r = (uint32_t)((((yt.z * hc[0] + 0x20000) >> 16) + ((yt.w * hn[0]) >> 16)) >> 2);
            const uint32_t g = (uint32_t)((((yt.z * hc[1] + 0x20000) >> 16) + ((yt.w * hn[1]) >> 16)) >> 2);
            const uint32_t b = (uint32_t)((((yt.z * hc[2] + 0x20000) >> 16) + ((yt.w * hn[2]) >> 16)) >> 2);
            const uint32_t N = r * 2125u + g * 7154u + b * 721u;
            uint32_t q = (uint32_t)(((uint64_t)N * 3518437209ull) >> 45);   // N / 10000 for N <= 2,550,000
            const uint32_t key = r | (g << 8) | (b << 16);
            uint32_t t0, t1;
            asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(t0), "=r"(t1) : "r"(exc_addr + q * 8u));
            q -= (uint32_t)(key == t0) | (uint32_t)(key == t1);
            out[(dy - row0) * OW + dx] = (uint8_t)q;
        }
        __syncwarp();
    };
    if (e < p.n) prefetch(e, recbuf);
    for (int it = 0; e < p.n; e += stride, it++) {
        const uint32_t* rec = (const uint32_t*)(recbuf + (it & 1) * p.rec_bytes);
        cp_async_wait_all();
        __syncwarp();
        if (e + stride < p.n) prefetch(e + stride, recbuf + ((it + 1) & 1) * p.rec_bytes);
        Hot h;
        hot_load(h, rec + (BS >> 2));
        const COLT* cols = (const COLT*)rec;
        const uint32_t* ids = rec + cfg.ids_off / 4;
        // ---- id image (RgbObservation layout: board | queue top right, holder bottom right) ----
        if (rows20) {
            for (int g2 = lane; g2 < (H >> 1); g2 += 32)
                fill_rows2_w20_strided(ids + 5 * g2, (uint32_t*)(pix + (2 * g2) * RW + P), (uint32_t*)(pix + (2 * g2 + 1) * RW + P));
        } else {
            for (int r = lane; r < H; r += 32) fill_board_row<0>(cfg, ids, pix, 0, r, RW);
        }
        for (int q = lane; q < Q; q += 32) {
            const uint4 rb = *(const uint4*)(s_rowbytes + ((int)((h.queue >> (4 * q)) & 15u)) * 16);
            const uint32_t wv[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint8_t* d = pix + i * RW + Wp + 4 * q;
                d[0] = (uint8_t)wv[i]; d[1] = (uint8_t)(wv[i] >> 8); d[2] = (uint8_t)(wv[i] >> 16); d[3] = (uint8_t)(wv[i] >> 24);
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
        __syncwarp();
        // ---- pass 1: per column the id its source pixel pair shares in the current / next source row (0xFF = two ids);
        //      uniform neighbourhoods read the table (explicit shared-window addresses: no generic loads / stores) ----
        uint32_t uc[NX], un[NX];
        int row_c = -1, row_n = -1, cnt = 0;
        auto pairs = [&](int sy, uint32_t (&dst)[NX]) {
            const uint8_t* prow = pix + sy * RW;
#pragma unroll
            for (int j = 0; j < NX; j++) {
                const uint32_t a = prow[sx0[j]], b2 = prow[sx1[j]];
                dst[j] = a == b2 ? a : 0xFFu;
            }
        };
        const uint32_t g_addr = smem_u32(s_g);
        uint8_t* g = p.frames + e * p.env_stride;
        const int reps = 1 + ((p.fill_mask && p.fill_mask[e]) ? p.fill_count : 0);   // reset envs: the frame fills the stack window
        for (int c0 = 0; c0 < OH; c0 += CR, nchunk++) {
            const int c1 = min(OH, c0 + CR);
            uint8_t* out = out0 + (nchunk & 1u) * p.out_bytes;
            if (tma) { bulk_wait_read1(); __syncwarp(); }   // the store that read this buffer two chunks ago is done
            const uint32_t o_addr = smem_u32(out) + lane - c0 * OW;
            for (int dy = c0; dy < c1; dy++) {
                const int4 yt = s_y[dy];       // sy0, sy1, b0, b1 (warp-uniform)
                if (yt.x != row_c) {
                    if (yt.x == row_n) {
#pragma unroll
                        for (int j = 0; j < NX; j++) uc[j] = un[j];
                    } else pairs(yt.x, uc);
                    row_c = yt.x;
                }
                if (yt.y != row_n) {
                    if (yt.y == row_c) {
#pragma unroll
                        for (int j = 0; j < NX; j++) un[j] = uc[j];
                    } else pairs(yt.y, un);
                    row_n = yt.y;
                }
                const bool one_row = yt.w == 0;            // the output row sits on a source row: the next row does not count
                const uint32_t grow = g_addr + dy * 64, orow = o_addr + dy * OW;
                bool rest = false;
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    const bool in = lane + 32 * j < OW;
                    const bool uni = uc[j] != 0xFFu && (one_row || un[j] == uc[j]);
                    if (in && uni) {
                        uint32_t v;
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(grow + gofs[j] + uc[j]));
                        asm volatile("st.shared.u8 [%0], %1;" ::"r"(orow + 32u * j), "r"(v) : "memory");
                    }
                    rest |= in && !uni;
                }
                if (__any_sync(0xffffffffu, rest)) {       // (warp-uniform) append this row's other pixels to the list
#pragma unroll
                    for (int j = 0; j < NX; j++) {
                        const bool in = lane + 32 * j < OW;
                        const bool uni = uc[j] != 0xFFu && (one_row || un[j] == uc[j]);
                        const unsigned m = __ballot_sync(0xffffffffu, in && !uni);
                        if (in && !uni) list[cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(dy | ((lane + 32 * j) << 8));
                        cnt += __popc(m);
                    }
                    if (cnt > LCAP - 32 * NX) { flush(cnt, out, c0); cnt = 0; }
                }
            }
            flush(cnt, out, c0);
            cnt = 0;
            // ---- store the chunk (reset envs: also into the preceding frames of the stack window) ----
            const uint32_t cb = (uint32_t)((c1 - c0) * OW);
            uint8_t* gc = g + (size_t)c0 * OW;
            if (tma) {
                fence_async_smem();
                __syncwarp();
                for (int r = lane; r < reps; r += 32) bulk_s2g(gc - (size_t)r * FB, out, cb);
                bulk_commit();   // (every lane commits a group per chunk, empty for most: wait_group.read 1 counts groups)
            } else {
                __syncwarp();
                for (int r = 0; r < reps; r++)
                    for (int i = lane; i < (int)cb; i += 32) (gc - (size_t)r * FB)[i] = out[i];
                __syncwarp();
            }
        }
        __syncwarp();
    }
    bulk_wait_all();
}

// ---- k_cnn_obs3: the first pass a WORD (four output pixels) at a time ------------------------------------------------------------
// The four pixels of output word w read the source columns s_lo(w) .. s_lo(w) + len - 1 (len <= 4: an enlarged axis) of the
// two source rows of output row dy.  Nearly always all of those ids are equal (empty field, bedrock, queue background): then the
// word is ONE read of the transposed table (the grey values of (dy, id) for the four classes of a0 + a1, one per byte) and ONE
// PRMT with the word's class pattern.  Lane = output word (the per-word constants live in registers, dy is warp-uniform).
//   pass A  words: fast test + table word, the other words go to a list (ballot compaction);
//   pass B  the listed words pixel by pixel, 32 at a time (dense): 2 x 2 ids equal -> table byte, else -> pixel list;
//   pass C  the listed pixels through the full fixed-point interpolation (as k_cnn_obs2's pass 2).
// k_cnn_obs2 spent ~25 thread-instructions per pixel in its row loop whatever the pixel was; here a fast word costs ~6 per pixel.
// RW4: the id image's row stride is a multiple of 4 (W = 20: 48), so the alignment of a lane's source fetch is the same in every row
template <class COLT, bool RW4>
__global__ void __launch_bounds__(256, 4) k_cnn_obs3(const __grid_constant__ CnnParams p) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint32_t s_lut[16];
    __shared__ uint32_t s_rowbytes[112];
    __shared__ __align__(16) int4 s_y[128];
    __shared__ __align__(8) uint2 s_exc[256];
    __shared__ uint8_t s_cls[128];
    __shared__ __align__(16) int4 s_xt[128];   // sx0, sx1, a0, a1 per output column
    const DevCfg& cfg = p.cfg;
    const int W = cfg.W, H = cfg.H, Wp = cfg.Wp, Hp = cfg.Hp, RW = cfg.rgb_w, Q = cfg.Q, BS = cfg.board_stride;
    const int OH = p.OH, OW = p.OW, FB = OH * OW, NWD = OW >> 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int NP = Hp * RW;
    uint32_t* s_gt = (uint32_t*)sm;        // [OH][16] transposed uniform-neighbourhood table, then one region per warp
    uint8_t* wbase = sm + p.gtab_bytes + (size_t)warp * (2 * p.rec_bytes + p.pix_bytes + 2 * p.out_bytes + p.wlist_bytes + p.list_bytes);
    uint8_t* recbuf = wbase;
    uint8_t* pix = wbase + 2 * p.rec_bytes;
    uint8_t* out0 = pix + p.pix_bytes;     // two chunk buffers: one is filled while the bulk store of the other drains
    unsigned short* wlist = (unsigned short*)(out0 + 2 * p.out_bytes);
    unsigned short* list = (unsigned short*)((uint8_t*)wlist + p.wlist_bytes);
    const int CR = p.chunk_rows;
    uint32_t nchunk = 0;
    const int LCAP = p.list_bytes / 2;
    if (threadIdx.x < 16) s_lut[threadIdx.x] = ((const uint32_t*)c_colors)[threadIdx.x];
    for (int i = threadIdx.x; i < 112; i += blockDim.x) s_rowbytes[i] = (&c_rowbytes[0][0][0])[i];
    for (int i = threadIdx.x; i < OH; i += blockDim.x) s_y[i] = ((const int4*)p.ytab)[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_exc[i] = ((const uint2*)p.gray_exc)[i];
    for (int i = threadIdx.x; i < OH * 16; i += blockDim.x) s_gt[i] = p.gtabT[i];
    for (int i = threadIdx.x; i < OW; i += blockDim.x) { s_cls[i] = p.xcls[i]; s_xt[i] = ((const int4*)p.xtab)[i]; }
    const uint32_t exc_addr = smem_u32(s_exc);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch, see k_step_ws
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = lane; i < NP; i += 32) {   // constant part of the id image (see k_rgb)
        int r = i / RW, c = i - r * RW;
        pix[i] = (c < Wp && r < H && c >= P && c < P + W) ? 0 : 1;
    }
    for (int i = lane; i < 2 * p.rec_bytes / 4; i += 32) ((uint32_t*)recbuf)[i] = 0;
    // this lane's output word: first source column, byte mask of its source columns, class pattern
    const bool wlane = lane < NWD;
    const int4 wt = wlane ? ((const int4*)p.wtab)[lane] : make_int4(0, 0, 0x3210, 0);
    const uint32_t w_lo = (uint32_t)wt.x, w_mask = (uint32_t)wt.y, w_sel = (uint32_t)wt.z;
    __syncthreads();
    const bool tma = (FB & 15) == 0 && (((uintptr_t)p.frames) & 15) == 0 && (p.env_stride & 15) == 0;
    const bool rows20 = W == 20 && (H & 1) == 0 && (RW & 3) == 0;
    const int64_t stride = (int64_t)gridDim.x * nwarps;
    int64_t e = (int64_t)blockIdx.x * nwarps + warp;
    auto prefetch = [&](int64_t ee, uint8_t* dst) {
        const uint8_t* src = p.board + ee * BS;
        for (int i = lane; i < (BS >> 4); i += 32) cp_async16(dst + 16 * i, src + 16 * i);
        if (lane < 2) cp_async16(dst + BS + 16 * lane, p.hot + ee * 32 + 16 * lane);
        cp_async_commit();
    };
    // pass C: the full interpolation of the listed pixels (entry = dy | dx << 8), 32 at a time
    auto flush = [&](int cnt, uint8_t* out, int row0) {
        __syncwarp();
        for (int k = lane; k < cnt; k += 32) {
            const int ent = list[k], dy = ent & 255, dx = ent >> 8;
            const int4 yt = s_y[dy], xt = s_xt[dx];
            const uint32_t acoef = (uint32_t)xt.z | ((uint32_t)xt.w << 16);
            const uint8_t* r0 = pix + yt.x * RW;
            const uint8_t* r1 = pix + yt.y * RW;
            int hc[3], hn[3];
            {
                const uint32_t c0 = s_lut[r0[xt.x]], c1 = s_lut[r0[xt.y]];
                const uint32_t rg = __byte_perm(c0, c1, 0x5140), bb = __byte_perm(c0, c1, 0x7762);
                hc[0] = (int)(__dp2a_lo(acoef, rg, 0u) >> 4); hc[1] = (int)(__dp2a_hi(acoef, rg, 0u) >> 4); hc[2] = (int)(__dp2a_lo(acoef, bb, 0u) >> 4);
            }
            {
                const uint32_t c0 = s_lut[r1[xt.x]], c1 = s_lut[r1[xt.y]];
                const uint32_t rg = __byte_perm(c0, c1, 0x5140), bb = __byte_perm(c0, c1, 0x7762);
                hn[0] = (int)(__dp2a_lo(acoef, rg, 0u) >> 4); hn[1] = (int)(__dp2a_hi(acoef, rg, 0u) >> 4); hn[2] = (int)(__dp2a_lo(acoef, bb, 0u) >> 4);
            }
            // VResizeLinear: (((b0 * h0) >> 16) + ((b1 * h1) >> 16) + 2) >> 2 (colours <= 240, pairs sum to 2048 +- 1: no clamping)
            const uint32_t r = (uint32_t)((((yt.z * hc[0] + 0x20000) >> 16) + ((yt.w * hn[0]) >> 16)) >> 2);
            const uint32_t g = (uint32_t)((((yt.z * hc[1] + 0x20000) >> 16) + ((yt.w * hn[1]) >> 16)) >> 2);
            const uint32_t b = (uint32_t)((((yt.z * hc[2] + 0x20000) >> 16) + ((yt.w * hn[2]) >> 16)) >> 2);
            const uint32_t N = r * 2125u + g * 7154u + b * 721u;
            uint32_t q = (uint32_t)(((uint64_t)N * 3518437209ull) >> 45);   // N / 10000 for N <= 2,550,000
            const uint32_t key = r | (g << 8) | (b << 16);
            uint32_t t0, t1;
            asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(t0), "=r"(t1) : "r"(exc_addr + q * 8u));
            q -= (uint32_t)(key == t0) | (uint32_t)(key == t1);
            out[(dy - row0) * OW + dx] = (uint8_t)q;
        }
        __syncwarp();
    };
    if (e < p.n) prefetch(e, recbuf);
    for (int it = 0; e < p.n; e += stride, it++) {
        const uint32_t* rec = (const uint32_t*)(recbuf + (it & 1) * p.rec_bytes);
        cp_async_wait_all();
        __syncwarp();
        if (e + stride < p.n) prefetch(e + stride, recbuf + ((it + 1) & 1) * p.rec_bytes);
        Hot h;
        hot_load(h, rec + (BS >> 2));
        const COLT* cols = (const COLT*)rec;
        const uint32_t* ids = rec + cfg.ids_off / 4;
        // ---- id image (RgbObservation layout: board | queue top right, holder bottom right) ----
        if (rows20) {
            for (int g2 = lane; g2 < (H >> 1); g2 += 32)
                fill_rows2_w20_strided(ids + 5 * g2, (uint32_t*)(pix + (2 * g2) * RW + P), (uint32_t*)(pix + (2 * g2 + 1) * RW + P));
        } else {
            for (int r = lane; r < H; r += 32) fill_board_row<0>(cfg, ids, pix, 0, r, RW);
        }
        for (int q = lane; q < Q; q += 32) {
            const uint4 rb = *(const uint4*)(s_rowbytes + ((int)((h.queue >> (4 * q)) & 15u)) * 16);
            const uint32_t wv[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint8_t* d = pix + i * RW + Wp + 4 * q;
                d[0] = (uint8_t)wv[i]; d[1] = (uint8_t)(wv[i] >> 8); d[2] = (uint8_t)(wv[i] >> 16); d[3] = (uint8_t)(wv[i] >> 24);
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
        __syncwarp();
        uint32_t pix_addr = smem_u32(pix), gt_addr = smem_u32(s_gt);
        // (opaque copies: under register pressure the compiler re-derives the shared-window addresses -- S2R SR_CgaCtaId + five
        //  instructions -- inside the row loop instead of keeping them)
        asm volatile("mov.u32 %0, %0;" : "+r"(pix_addr));
        asm volatile("mov.u32 %0, %0;" : "+r"(gt_addr));
        const uint32_t f_al = (pix_addr + w_lo) & ~3u, f_sh = ((pix_addr + w_lo) & 3u) * 8u;   // RW4: aligned word / shift of this lane's fetch
        uint8_t* g = p.frames + e * p.env_stride;
        const int reps = 1 + ((p.fill_mask && p.fill_mask[e]) ? p.fill_count : 0);   // reset envs: the frame fills the stack window
        for (int c0 = 0; c0 < OH; c0 += CR, nchunk++) {
            const int c1 = min(OH, c0 + CR);
            uint8_t* out = out0 + (nchunk & 1u) * p.out_bytes;
            if (tma) { bulk_wait_read1(); __syncwarp(); }   // the store that read this buffer two chunks ago is done
            const uint32_t ob_addr = smem_u32(out), o_addr = ob_addr + 4u * lane - c0 * OW;
            // ---- pass A: one output word per lane and row (explicit shared-window addresses: no generic loads / stores) ----
            int wcnt = 0;
            for (int dy = c0; dy < c1; dy++) {
                const int4 yt = s_y[dy];                       // sy0, sy1, b0, b1 (warp-uniform)
                const int sy1 = yt.w == 0 ? yt.x : yt.y;       // the output row sits on a source row: the next row does not count
                uint32_t f0, f1;
                {
                    const uint32_t a = pix_addr + yt.x * RW + w_lo, al = RW4 ? f_al + yt.x * RW : (a & ~3u);
                    uint32_t lo, hi;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(al));
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"(al + 4u));
                    f0 = __funnelshift_r(lo, hi, RW4 ? f_sh : (a & 3u) * 8u);
                }
                {
                    const uint32_t a = pix_addr + sy1 * RW + w_lo, al = RW4 ? f_al + sy1 * RW : (a & ~3u);
                    uint32_t lo, hi;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(al));
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"(al + 4u));
                    f1 = __funnelshift_r(lo, hi, RW4 ? f_sh : (a & 3u) * 8u);
                }
                const uint32_t id = f0 & 255u, rep4 = id * 0x01010101u;
                const bool fast = (((f0 ^ rep4) | (f1 ^ rep4)) & w_mask) == 0;
                if (wlane && fast) {
                    uint32_t t;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(gt_addr + (uint32_t)(dy * 16 + (int)id) * 4u));
                    const uint32_t v = prmt_raw(t, 0u, w_sel);
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(o_addr + (uint32_t)(dy * OW)), "r"(v) : "memory");
                }
                // (a per-lane bit mask of the chunk's rows + one scan per chunk instead of a ballot per row measured slower: the lanes on
                //  the field's walls append in every row and the others wait for their loop)
                const bool slow = wlane && !fast;
                const unsigned m = __ballot_sync(0xffffffffu, slow);
                if (slow) wlist[wcnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(dy | (lane << 8));
                wcnt += __popc(m);
            }
            __syncwarp();
            // ---- pass B: the pixels of the listed words, 32 at a time ----
            int cnt = 0;
            for (int k0 = 0; k0 < 4 * wcnt; k0 += 32) {
                const int k = k0 + lane;
                const bool valid = k < 4 * wcnt;
                bool rest = false;
                int dy = 0, dx = 0;
                if (valid) {
                    const int ent = wlist[k >> 2];
                    dy = ent & 255; dx = 4 * (ent >> 8) + (k & 3);
                    const int4 yt = s_y[dy], xt = s_xt[dx];
                    const uint32_t sxb = (uint32_t)(xt.w == 0 ? xt.x : xt.y);   // a neighbour with a zero coefficient does not count
                    const uint32_t r0 = pix_addr + (uint32_t)(yt.x * RW), r1 = pix_addr + (uint32_t)((yt.w == 0 ? yt.x : yt.y) * RW);
                    uint32_t a, b2, c2, d2;
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(a) : "r"(r0 + (uint32_t)xt.x));
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b2) : "r"(r0 + sxb));
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c2) : "r"(r1 + (uint32_t)xt.x));
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(d2) : "r"(r1 + sxb));
                    if (a == b2 && a == c2 && a == d2) {
                        uint32_t v;   // byte (class of a0 + a1) of the table word of (dy, id)
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(gt_addr + (uint32_t)(dy * 16 + (int)a) * 4u + (uint32_t)s_cls[dx]));
                        asm volatile("st.shared.u8 [%0], %1;" ::"r"(ob_addr + (uint32_t)((dy - c0) * OW + dx)), "r"(v) : "memory");
                    } else rest = true;
                }
                const unsigned m = __ballot_sync(0xffffffffu, rest);
                if (rest) list[cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(dy | (dx << 8));
                cnt += __popc(m);
                if (cnt > LCAP - 32) { flush(cnt, out, c0); cnt = 0; }
            }
            flush(cnt, out, c0);
            // ---- store the chunk (reset envs: also into the preceding frames of the stack window) ----
            const uint32_t cb = (uint32_t)((c1 - c0) * OW);
            uint8_t* gc = g + (size_t)c0 * OW;
            if (tma) {
                fence_async_smem();
                __syncwarp();
                for (int r = lane; r < reps; r += 32) bulk_s2g(gc - (size_t)r * FB, out, cb);
                bulk_commit();   // (every lane commits a group per chunk, empty for most: wait_group.read 1 counts groups)
            } else {
                __syncwarp();
                for (int r = 0; r < reps; r++)
                    for (int i = lane; i < (int)cb; i += 32) (gc - (size_t)r * FB)[i] = out[i];
                __syncwarp();
            }
        }
        __syncwarp();
    }
    bulk_wait_all();
}

}  // namespace tg


// ---- host side --------------------------------------------------------------------------------------------------
#include <math.h>
// offsets and 11-bit coefficient pairs of one axis, exactly as cv::hal::resize computes them for INTER_AREA when the
// kernel is the bilinear emulation (area_mode): double scale factors, float fractions, saturate_cast<short> (lrintf).
// tab[d] = {s0, s1, c0, c1}: source indices (s1 clamped; equal to s0 with c = {2048, 0} at the right / bottom border).
static void cnn_axis_tables(int ssize, int dsize, std::vector<int32_t>& tab, bool clamp_pair) {
    const double inv_scale = (double)dsize / ssize, scale = 1. / inv_scale;
    tab.assign((size_t)dsize * 4, 0);
    for (int d = 0; d < dsize; d++) {
        int s = (int)floor(d * scale);
        float f = (float)((d + 1) - (s + 1) * inv_scale);
        f = f <= 0 ? 0.f : f - floorf(f);
        if (s < 0) { f = 0.f; s = 0; }
        bool border = false;
        if (s + 1 >= ssize) {
            border = true;                       // HResizeLinear: dx >= xmax reads S[sx] * ONE
            if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
        }
        int c0 = (int)lrintf((1.f - f) * 2048.f), c1 = (int)lrintf(f * 2048.f);
        c0 = c0 > 32767 ? 32767 : (c0 < -32768 ? -32768 : c0);
        c1 = c1 > 32767 ? 32767 : (c1 < -32768 ? -32768 : c1);
        int s1 = s + 1 < ssize ? s + 1 : ssize - 1;
        if (clamp_pair && border) { s1 = s; c0 = 2048; c1 = 0; }
        tab[4 * d] = s; tab[4 * d + 1] = s1; tab[4 * d + 2] = c0; tab[4 * d + 3] = c1;
    }
}

// R | G << 8 | B << 16 of the triples whose float64 grey value is one below N / 10000, two slots per quotient
static int cnn_gray_exceptions(std::vector<uint32_t>& tab) {
    tab.assign(512, 0xFFFFFFFFu);
    std::vector<int> cnt(256, 0);
    for (int r = 0; r < 256; r++)
        for (int g = 0; g < 256; g++)
            for (int b = 0; b < 256; b++) {
                const int N = 2125 * r + 7154 * g + 721 * b, q = N / 10000;
                volatile double f = r * 0.2125 + g * 0.7154;   // left-to-right float64 sum, no contraction
                f = f + b * 0.0721;
                const int v = (int)f;
                if (v == q) continue;
                if (v != q - 1 || cnt[q] >= 2) return -1;   // the integer scheme would not hold
                tab[2 * q + cnt[q]++] = (uint32_t)(r | (g << 8) | (b << 16));
            }
    return 0;
}

// grey value of an output pixel whose source pixels all hold colour (R, G, B): the chain of k_cnn_obs with h0 = h1 = c * ax
static uint8_t cnn_uniform_value(const unsigned char* rgb, int ax, int b0, int b1) {
    int v[3];
    for (int k = 0; k < 3; k++) {
        const int hh = ((int)rgb[k] * ax) >> 4;
        v[k] = (((b0 * hh) >> 16) + ((b1 * hh) >> 16) + 2) >> 2;
        v[k] = v[k] < 0 ? 0 : (v[k] > 255 ? 255 : v[k]);
    }
    volatile double f = v[0] * 0.2125 + v[1] * 0.7154;   // GrayscaleObservation: float64, summed left to right
    f = f + v[2] * 0.0721;
    return (uint8_t)(int)f;
}

extern "C" int tg_cnn_observe(tg_env* env, tg_state st, int64_t n, int32_t out_h, int32_t out_w, uint8_t* d_frames,
                              int64_t env_stride, const uint8_t* d_fill_mask, int32_t fill_count, void* stream) {
    if (!env) return TG_ERR_POINTER;
    int rc = check_state(env, st); if (rc) return rc;
    if (!d_frames) return fail(env, TG_ERR_POINTER, "d_frames is NULL");
    const DevCfg& d = env->dev;
    if (out_h < 1 || out_w < 1 || out_h > 128 || out_w > 128) return fail(env, TG_ERR_ARG, "tg_cnn_observe: output size must be within 1..128");
    if (d.rgb_w >= out_w && d.Hp >= out_h)
        return fail(env, TG_ERR_CONFIG, "tg_cnn_observe: both axes shrink (true area interpolation) -- not supported");
    if (fill_count < 0 || env_stride < (int64_t)out_h * out_w) return fail(env, TG_ERR_ARG, "tg_cnn_observe: bad fill_count / env_stride");
    ON_DEVICE(env);
    if (env->cnn_h != out_h || env->cnn_w != out_w) {   // (re)build the coefficient tables for this output size
        std::vector<int32_t> xt, yt;
        cnn_axis_tables(d.rgb_w, out_w, xt, true);
        cnn_axis_tables(d.Hp, out_h, yt, false);
        std::vector<uint32_t> gray;
        if (cnn_gray_exceptions(gray)) return fail(env, TG_ERR_CONFIG, "tg_cnn_observe: float64 grey conversion does not follow the integer scheme on this host");
        gray.resize(1536, 0xFFFFFFFFu);
        // k_cnn_obs2: classes of a0 + a1 (at most four, else the one-pass kernel runs) and the uniform-neighbourhood table
        env->cnn_nax = 0;
        for (int dx = 0; dx < out_w; dx++) {
            const int ax = xt[4 * dx + 2] + xt[4 * dx + 3];
            int k = 0;
            while (k < env->cnn_nax && env->cnn_axv[k] != ax) k++;
            if (k == env->cnn_nax) { if (k < 4) env->cnn_axv[k] = ax; env->cnn_nax++; }
        }
        for (int k = env->cnn_nax; k < 4; k++) env->cnn_axv[k] = -1;
        std::vector<uint8_t> gt((size_t)out_h * 64, 0);
        if (env->cnn_nax <= 4)
            for (int dy = 0; dy < out_h; dy++)
                for (int k = 0; k < env->cnn_nax; k++)
                    for (int id = 0; id < 16; id++)
                        gt[((size_t)dy * 4 + k) * 16 + id] = cnn_uniform_value(env->tabs.colors[id], env->cnn_axv[k], yt[4 * dy + 2], yt[4 * dy + 3]);
        // k_cnn_obs3: the table transposed (the four classes of (dy, id) in one word), per output word its first source column,
        // the byte mask of its source columns and its class pattern as a PRMT selector, per output column its class
        std::vector<uint32_t> gtT((size_t)out_h * 16, 0);
        for (int dy = 0; dy < out_h; dy++)
            for (int id = 0; id < 16; id++)
                for (int k = 0; k < 4; k++) gtT[(size_t)dy * 16 + id] |= (uint32_t)gt[((size_t)dy * 4 + k) * 16 + id] << (8 * k);
        const int nwd = out_w / 4;
        std::vector<int32_t> wt((size_t)(nwd > 0 ? nwd : 1) * 4, 0);
        std::vector<uint8_t> xc(((size_t)out_w + 15) / 16 * 16, 0);
        env->cnn_v3_ok = env->cnn_nax <= 4 && out_w % 4 == 0 && nwd >= 1 && nwd <= 32 && out_h <= 128;
        for (int dx = 0; dx < out_w; dx++) {
            const int ax = xt[4 * dx + 2] + xt[4 * dx + 3];
            int k = 0;
            while (k < 4 && env->cnn_axv[k] != ax) k++;
            xc[dx] = (uint8_t)(k & 3);
        }
        for (int w = 0; w < nwd && env->cnn_v3_ok; w++) {
            int lo = 1 << 30, hi = -1;
            uint32_t sel = 0;
            for (int j = 0; j < 4; j++) {
                const int dx = 4 * w + j, s0 = xt[4 * dx], s1 = xt[4 * dx + 3] == 0 ? s0 : xt[4 * dx + 1];   // a neighbour with a zero coefficient does not count
                lo = std::min(lo, std::min(s0, s1)); hi = std::max(hi, std::max(s0, s1));
                sel |= (uint32_t)xc[dx] << (4 * j);
            }
            const int len = hi - lo + 1;
            if (len > 4) { env->cnn_v3_ok = false; break; }   // (an axis enlarged by less than 4/3: the one-row kernels run)
            wt[4 * w] = lo; wt[4 * w + 1] = (int32_t)(len == 4 ? 0xFFFFFFFFu : ((1u << (8 * len)) - 1u)); wt[4 * w + 2] = (int32_t)sel;
        }
        size_t bytes = (xt.size() + yt.size()) * 4 + 6144 + gt.size() + gtT.size() * 4 + wt.size() * 4 + xc.size();
        rc = ensure_stage(env, 4, bytes); if (rc) return rc;
        uint8_t* base = (uint8_t*)env->stage[4];
        CUDA_TRY(env, cudaMemcpy(base, gray.data(), 6144, cudaMemcpyHostToDevice));
        CUDA_TRY(env, cudaMemcpy(base + 6144, xt.data(), xt.size() * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(env, cudaMemcpy(base + 6144 + xt.size() * 4, yt.data(), yt.size() * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(env, cudaMemcpy(base + 6144 + (xt.size() + yt.size()) * 4, gt.data(), gt.size(), cudaMemcpyHostToDevice));
        uint8_t* b3 = base + 6144 + (xt.size() + yt.size()) * 4 + gt.size();
        CUDA_TRY(env, cudaMemcpy(b3, gtT.data(), gtT.size() * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(env, cudaMemcpy(b3 + gtT.size() * 4, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(env, cudaMemcpy(b3 + gtT.size() * 4 + wt.size() * 4, xc.data(), xc.size(), cudaMemcpyHostToDevice));
        env->cnn_h = out_h; env->cnn_w = out_w;
    }
    CnnParams p;
    memset(&p, 0, sizeof p);
    p.cfg = d; p.n = n; p.hot = (const uint8_t*)st.hot; p.board = (const uint8_t*)st.board;
    uint8_t* base = (uint8_t*)env->stage[4];
    p.gray_exc = (const uint32_t*)base; p.xtab = (const int32_t*)(base + 6144); p.ytab = p.xtab + (size_t)out_w * 4;
    p.frames = d_frames; p.env_stride = env_stride; p.fill_mask = d_fill_mask; p.fill_count = fill_count;
    p.OH = out_h; p.OW = out_w;
    auto r128 = [](size_t v) { return (int)((v + 127) / 128 * 128); };
    p.rec_bytes = r128((size_t)d.board_stride + 48); p.pix_bytes = r128((size_t)d.Hp * d.rgb_w + 16); p.out_bytes = r128((size_t)out_h * out_w);
    const bool two_pass = env->cnn_nax <= 4 && !getenv("TG_CNN_V1");   // TG_CNN_V1=1: the one-pass kernel
    p.gtab = base + 6144 + ((size_t)out_w + out_h) * 16;
    p.gtabT = (const uint32_t*)(p.gtab + (size_t)out_h * 64);
    p.wtab = (const int32_t*)(p.gtabT + (size_t)out_h * 16);
    p.xcls = (const uint8_t*)(p.wtab + (size_t)(out_w / 4 > 0 ? out_w / 4 : 1) * 4);
    for (int k = 0; k < 4; k++) p.axv[k] = env->cnn_axv[k];
    // TG_CNN_V2=1: the per-pixel two-pass kernel instead of the word-wise one
    bool word_wise = two_pass && env->cnn_v3_ok && !getenv("TG_CNN_V2");
    p.list_bytes = two_pass ? 512 : 0;
    p.gtab_bytes = two_pass ? r128((size_t)out_h * 64) : 0;
    int nw = two_pass ? 8 : 4;
    if (const char* t = getenv("TG_CNN_NW")) { int v = atoi(t); if (v >= 1 && v <= 8) nw = v; }
    if (two_pass) {
        // the frame leaves in chunks of whole rows whose byte count is a multiple of 16 (bulk copies): two small chunk buffers
        // per warp instead of a whole frame keep three 8-warp CTAs resident
        int cr = 8;
        if (((size_t)out_h * out_w) % 16 == 0) while (((size_t)cr * out_w) % 16) cr++;
        else cr = out_h;                       // (plain stores: any chunking works; keep one)
        if (const char* t = getenv("TG_CNN_CR")) { int v = atoi(t); if (v >= 1 && ((size_t)v * out_w) % 16 == 0) cr = v; }
        if (cr > out_h) cr = out_h;
        p.chunk_rows = cr;
        p.out_bytes = r128((size_t)cr * out_w);
        p.wlist_bytes = word_wise ? r128((size_t)2 * cr * (out_w / 4)) : 0;
        if (word_wise) p.list_bytes = 256;
    }
    const int T = nw * 32;
    const size_t smem = (size_t)p.gtab_bytes + (size_t)nw * (2 * p.rec_bytes + p.pix_bytes + (two_pass ? 2 : 1) * p.out_bytes + p.wlist_bytes + p.list_bytes);
    if (smem > 200 * 1024) return fail(env, TG_ERR_CONFIG, "tg_cnn_observe: image too large for shared memory");
    const int NX = (out_w + 31) / 32;
    auto launch = [&](auto kern) -> int {
        CUDA_TRY(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        CUDA_TRY(env, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem));
        int64_t blocks = (n + nw - 1) / nw, cap = (int64_t)env->num_sms * (per_sm > 0 ? per_sm : 1);
        if (blocks > cap) blocks = cap;
        CUDA_TRY(env, launch_pdl(kern, (unsigned)blocks, (unsigned)T, smem, (cudaStream_t)stream, p));
        return TG_OK;
    };
    if (word_wise) {
        if ((d.rgb_w & 3) == 0) return env->col64 ? launch(k_cnn_obs3<uint64_t, true>) : launch(k_cnn_obs3<uint32_t, true>);
        return env->col64 ? launch(k_cnn_obs3<uint64_t, false>) : launch(k_cnn_obs3<uint32_t, false>);
    }
    if (two_pass) {
        if (env->col64) return NX == 1 ? launch(k_cnn_obs2<uint64_t, 1>) : NX == 2 ? launch(k_cnn_obs2<uint64_t, 2>) : NX == 3 ? launch(k_cnn_obs2<uint64_t, 3>) : launch(k_cnn_obs2<uint64_t, 4>);
        return NX == 1 ? launch(k_cnn_obs2<uint32_t, 1>) : NX == 2 ? launch(k_cnn_obs2<uint32_t, 2>) : NX == 3 ? launch(k_cnn_obs2<uint32_t, 3>) : launch(k_cnn_obs2<uint32_t, 4>);
    }
    if (env->col64) return NX == 1 ? launch(k_cnn_obs<uint64_t, 1>) : NX == 2 ? launch(k_cnn_obs<uint64_t, 2>) : NX == 3 ? launch(k_cnn_obs<uint64_t, 3>) : launch(k_cnn_obs<uint64_t, 4>);
    return NX == 1 ? launch(k_cnn_obs<uint32_t, 1>) : NX == 2 ? launch(k_cnn_obs<uint32_t, 2>) : NX == 3 ? launch(k_cnn_obs<uint32_t, 3>) : launch(k_cnn_obs<uint32_t, 4>);
}
