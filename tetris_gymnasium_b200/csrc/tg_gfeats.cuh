// tg_gfeats.cuh -- GroupedActionsObservations + FeatureVectorObservation (wrappers/grouped.py:124-207 with
// wrappers/observation.py:238-278 applied to every placement board), packed-byte formulation for W = 10 / W = 20.
//
// CTA = 32 envs x W threads: warp = board column xb of the placement (a = 4 xb + r, wrappers/grouped.py:78-99), lane = env.
// One thread evaluates the FOUR rotations of its (env, column): they share x, the four board columns under the 4x4 piece
// matrix and the window of column heights, and their 4 x F feature bytes are exactly F aligned 32-bit words of the output.
//   * heights live as packed bytes (P pad bytes | W heights | pad): the <= 4 touched columns are a 4-byte window at byte x;
//     new heights = bytewise max(old, H - y - top offset) under the piece's column mask, re-inserted with two funnel shifts;
//   * holes' = holes + sum(new) - sum(old) - 4 (IDP.4A), max' = max(max, H - y - min top offset),
//     bumpiness = sum |h[c+1] - h[c]| over the packed vector (VABSDIFF4 with accumulate), no incremental bookkeeping;
//   * full rows = L & R & AND_j (col_j | piece column j << y) over the four window columns, L / R = prefix / suffix AND of
//     the field columns left / right of the window (the same for the four rotations; wall columns are all ones);
//   * placements that clear rows or put a cell into the row the feature wrapper zeroes are rare: they are
//     collected in shared memory and evaluated column by column (exact row-clear arithmetic) in a dense second pass.
// The tile [32][4W][F] is staged in shared memory in output layout and leaves with 128-bit stores.
#pragma once

namespace tg {

// per (piece, rotation): .x = cells (16 bit) | row masks of matrix columns 0..3 (nibbles) << 16,
// .y = byte j 0xFF if column j holds cells, .z = byte j = row offset of the top cell of column j,
// .w = first column | last column << 2 | smallest top offset << 4
// (the packed-byte kernels re-code .w for their width when they stage the table: prec_w_for_width)
__constant__ uint4 c_prec[7][4];
// per (piece, rotation): byte j = row offset of the LOWEST cell of matrix column j (0 where the column is empty)
__constant__ unsigned int c_bot4[7][4];

__device__ __forceinline__ uint32_t bytemax_lt128(uint32_t a, uint32_t b) {   // bytewise max, all bytes < 128
    const uint32_t d = (a | 0x80808080u) - b;                    // bit 7 of byte i = (a_i >= b_i), no borrow between bytes
    const uint32_t m = ((d >> 7) & 0x01010101u) * 0xFFu;
    return (a & m) | (b & ~m);
}

// .w of a staged c_prec entry for board width W: bit x (x < 28) = the placement at matrix position x keeps the piece inside the
// field (collision_with_frame, wrappers/grouped.py:101-122: x + first column >= P and x + last column < W + P),
// smallest top offset << 28, first column << 30 (the last column is the top byte of .y's column mask)
template <int W>
__device__ __forceinline__ uint32_t prec_w_for_width(uint32_t w) {
    static_assert(W + P <= 28, "x mask and the two small fields share one word");
    const int jmin = w & 3, jmax = (w >> 2) & 3, mintop = (w >> 4) & 3;
    const uint32_t xmask = ((1u << (W + P - jmax)) - 1u) & ~((1u << (P - jmin)) - 1u);
    return xmask | ((uint32_t)mintop << 28) | ((uint32_t)jmin << 30);
}
// max of the four bytes of v (__vmaxu4 is emulated on sm_100a: ~8 instructions per call)
__device__ __forceinline__ uint32_t max4bytes(uint32_t v) {
    return max(max(v & 255u, __byte_perm(v, 0u, 0x4441)), max(__byte_perm(v, 0u, 0x4442), v >> 24));
}
// PRMT with a selector whose nibbles are already 0..7 (__byte_perm masks its selector with 0x7777: one LOP3 per use)
__device__ __forceinline__ uint32_t prmt_raw(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// Window merge selectors: hw[k] = prmt_raw(V[k + 1], N4, sel[k]) replaces the bytes x .. x + 3 of the padded height vector by
// the window N4 (byte g of the vector = byte g - 4 (k + 1) of word k + 1); one row of 8 selectors per x.
template <int NH>
__device__ __forceinline__ void build_sel_table(uint32_t* s_sel, int nx, int tid, int nthreads) {
    for (int i = tid; i < nx * 8; i += nthreads) {
        const int x = i >> 3, k = i & 7;
        uint32_t sel = 0x3210u;
        if (k < NH) {
            const int d = 4 * (k + 1) - x;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int t = d + b;
                if (t >= 0 && t < 4) sel = (sel & ~(0xFu << (4 * b))) | ((uint32_t)(4 + t) << (4 * b));
            }
        }
        s_sel[i] = sel;
    }
}

// OR `nb` low bytes of v into the byte stream o[] at the compile-time byte offset OFF
template <int OFF, int NB, int NWORDS>
__device__ __forceinline__ void put_bytes(uint32_t (&o)[NWORDS], uint32_t v) {
    constexpr int w = OFF >> 2, s = (OFF & 3) * 8;
    const uint32_t vm = NB >= 4 ? v : (v & ((1u << (8 * (NB & 3))) - 1u));
    o[w] |= vm << s;
    if (s + 8 * NB > 32) o[w + 1] |= vm >> ((32 - s) & 31);
}

template <int W, int R, int NWORDS>
__device__ __forceinline__ void put_row(uint32_t (&o)[NWORDS], const uint32_t (&hw)[(W + 3) / 4], uint32_t summary) {
    constexpr int F = W + 3, NH = (W + 3) / 4, V = W - 4 * (NH - 1);   // V = heights in the last word (1..4)
#pragma unroll
    for (int k = 0; k < NH - 1; k++) {
        // (static offsets after unrolling)
        const int off = R * F + 4 * k;
        const int w = off >> 2, s = (off & 3) * 8;
        o[w] |= hw[k] << s;
        if (s) o[w + 1] |= hw[k] >> (32 - s);
    }
    put_bytes<R * F + 4 * (NH - 1), V, NWORDS>(o, hw[NH - 1]);
    put_bytes<R * F + W, 3, NWORDS>(o, summary);
}

#ifndef GF_REGS
#define GF_REGS 48   // 4 CTAs of 320 threads per SM, no spills (40 registers / 5 CTAs measured 3 % slower)
#endif
template <int W, class COLT>
struct GFeatsSmem {
    static constexpr int EPB = 32, A = 4 * W, F = W + 3, WP = W + 2 * P;
    static constexpr int CS = WP | 1;                 // column row stride (odd: lanes = envs hit distinct banks)
    static constexpr int HW = (WP + 3) / 4;           // words of the padded height vector
    static constexpr int HS = HW | 1;
    static constexpr size_t off_colp = 0;
    static constexpr size_t off_hv = off_colp + sizeof(COLT) * EPB * CS;
    static constexpr size_t off_hol = off_hv + 4 * EPB * HS;              // u8 [EPB][W padded to 4]
    static constexpr size_t off_w0 = off_hol + EPB * ((W + 3) & ~3);
    static constexpr size_t off_slow = off_w0 + 4 * EPB;                  // u16 [EPB * A]
    static constexpr size_t off_prec = (off_slow + 2 * EPB * A + 127) & ~size_t(127);
    // piece table, 8 replicas interleaved: entry i of replica g at [i * 8 + g], i.e. always in 16-byte bank group g.  A 128-bit
    // shared load is served a quarter warp at a time; lanes read the entry of THEIR env's piece, and with one copy the 28 entries
    // fall into two bank groups per rotation (index 4 * piece + rot): 3 - 4 wavefronts per quarter instead of one
    static constexpr size_t off_feats = off_prec + 16 * 28 * 8;
    static constexpr size_t off_legal = off_feats + (size_t)EPB * A * F;  // A * F is a multiple of 4
    static constexpr size_t off_ih = off_legal + (size_t)EPB * A;         // info["board"]: u8 [EPB][W4] heights, holes
    static constexpr size_t off_iho = off_ih + EPB * ((W + 3) & ~3);
    static constexpr size_t off_info = off_iho + EPB * ((W + 3) & ~3);    // u8 [EPB][F]
    static constexpr size_t bytes = off_info + (size_t)EPB * F;
};

// tile_sync<FUSED>: the W feature warps of a CTA meet; in the fused kernel a logic warp shares the CTA, so they use a named barrier
template <int W, bool FUSED>
__device__ __forceinline__ void gf_sync() {
    if (FUSED) asm volatile("bar.sync 1, %0;" ::"n"(32 * W) : "memory");
    else __syncthreads();
}

// One tile of 32 envs (thread = (column xb, env e), tid < 32 * W).  `hot_t` / `board_t` point at the tile's first record (global
// memory, or the shared-memory stage of the fused kernel), `fill_t` (nullable) at its "illegal action + terminate" flags.
// FUSED (persistent CTA, the tile buffers are reused): the caller's thread 0 waits for the previous tile's bulk stores before
// the first barrier; here the stores are only committed.
template <int W, class COLT, bool FUSED>
__device__ __forceinline__ void gfeats_tile(const DevCfg& cfg, uint8_t* sm, int* s_nslow, const uint32_t* s_bot, const uint32_t* s_sel, int tid, int64_t base, int nv,
                                            const uint8_t* hot_t, const uint8_t* board_t, const uint8_t* fill_t,
                                            uint8_t* __restrict__ feats, uint8_t* legal, uint8_t* __restrict__ info_board,
                                            int consumed_bar = 0) {
    using S = GFeatsSmem<W, COLT>;
    constexpr int EPB = S::EPB, A = S::A, F = S::F, CS = S::CS, HW = S::HW, HS = S::HS;
    constexpr int NH = (W + 3) / 4, VL = W - 4 * (NH - 1), T = 32 * W;
    COLT* s_colp = (COLT*)(sm + S::off_colp);
    uint32_t* s_hv = (uint32_t*)(sm + S::off_hv);
    uint8_t* s_hol = sm + S::off_hol;
    uint32_t* s_w0 = (uint32_t*)(sm + S::off_w0);
    unsigned short* s_slow = (unsigned short*)(sm + S::off_slow);
    uint4* s_prec = (uint4*)(sm + S::off_prec);
    uint32_t* s_featw = (uint32_t*)(sm + S::off_feats);
    uint32_t* s_legalw = (uint32_t*)(sm + S::off_legal);
    uint8_t* s_ih = sm + S::off_ih;
    uint8_t* s_iho = sm + S::off_iho;
    uint8_t* s_info = sm + S::off_info;
    constexpr int W4 = (W + 3) & ~3;

    const int H = cfg.H, e = tid & 31, xb = tid >> 5;
    const bool live = e < nv;
    const COLT field = (COLT(1) << H) - 1;

    // ---- phase 1: thread = (column xb, env e): column -> shared memory, its height / holes with row 0 zeroed (Q1) ----
    if (tid < 2) s_nslow[tid] = 0;
    if (live) {
        const COLT col = ((const COLT*)(board_t + (size_t)e * cfg.board_stride))[xb];
        s_colp[e * CS + P + xb] = col;
        if (xb < 2 * P) s_colp[e * CS + (xb < P ? xb : W + xb)] = ~COLT(0);          // bedrock wall columns
        int hgt, hol;
        col_features<COLT>(col & ~COLT(1), H, hgt, hol);
        uint8_t* hv = (uint8_t*)(s_hv + e * HS);
        hv[P + xb] = (uint8_t)hgt;
        if (xb < P) hv[xb] = 0;
        if (P + W + xb < 4 * HW) hv[P + W + xb] = 0;
        s_hol[e * ((W + 3) & ~3) + xb] = (uint8_t)hol;
        if (W + xb < ((W + 3) & ~3)) s_hol[e * ((W + 3) & ~3) + W + xb] = 0;
        if (xb == 0)   // bit 31: illegal action + terminate -> the observation is filled with `high`
            s_w0[e] = (*(const uint32_t*)(hot_t + (size_t)e * 32) & 0x7FFFFFFFu) | ((fill_t && fill_t[e]) ? 0x80000000u : 0u);
    }
    gf_sync<W, FUSED>();
    // fused kernel: the staged records are not read any more -- the logic warp may load the tile after next into this stage
    if (FUSED && consumed_bar) asm volatile("bar.arrive %0, %1;" ::"r"(consumed_bar), "n"(32 * W + 32) : "memory");
    // ---- phase 3: the four rotations of (env e, column xb) ----
    if (live && info_board) {
        // info["board"] = FeatureVectorObservation of the real observation (wrappers/grouped.py:260-264): rows 0-1 zeroed
        // (SURVEY Q1), the active piece projected when it does not collide.  This thread: column xb of env e.
        const uint32_t w0 = s_w0[e];
        const int xa = w0 & 63, ya = (w0 >> 6) & 127;
        const uint4 pa = s_prec[(((w0 >> 13) & 7) * 4 + ((w0 >> 16) & 3)) * 8 + (e & 7)];
        const COLT* colp = s_colp + e * CS;
        COLT Ba = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = (pa.x >> (4 * k)) & 15;
            Ba |= colp[xa + (c & 3)] >> (c >> 2);
        }
        COLT v = colp[P + xb];
        const int j = xb + P - xa;
        if (!((Ba >> ya) & 1) && (unsigned)j < 4u) v |= (COLT)((pa.x >> (16 + 4 * j)) & 15u) << ya;
        int hgt, hol;
        col_features<COLT>(v & ~COLT(3), H, hgt, hol);
        s_ih[e * W4 + xb] = (uint8_t)hgt; s_iho[e * W4 + xb] = (uint8_t)hol;
    }
    if (live) {
        const uint32_t w0 = s_w0[e];
        uint32_t o[F];
#pragma unroll
        for (int i = 0; i < F; i++) o[i] = 0;
        uint32_t legal_w = 0;
        if (w0 >> 31) {
            // illegal action + terminate: obs = ones * high (wrappers/grouped.py:221-226); legal mask unchanged
            const uint32_t hi = (uint32_t)min(255, H * W) * 0x01010101u;
#pragma unroll
            for (int i = 0; i < F; i++) o[i] = hi;
            legal_w = ((const uint32_t*)legal)[(base + e) * W + xb];
        } else {
            const int piece = (w0 >> 13) & 7, rot0 = (w0 >> 16) & 3;
            const int nhalf = (int)((cfg.nhalf3 >> (3 * piece)) & 7u);    // n // 2 (reference set: 2 for I, 1 otherwise)
            const int x = xb + P - nhalf;                               // wrappers/grouped.py:157-158
            const COLT* colp = s_colp + e * CS;
            COLT cj[4];
#pragma unroll
            for (int j = 0; j < 4; j++) cj[j] = colp[x + j];
            const uint32_t* hvw = s_hv + e * HS;
            uint32_t V[HW];
#pragma unroll
            for (int k = 0; k < HW; k++) V[k] = hvw[k];
            const int wi = x >> 2, sh = (x & 3) * 8;
            const uint32_t O4 = __funnelshift_r(hvw[wi], hvw[wi + 1], sh);   // the window: heights of board columns x - P .. x - P + 3 (pads 0)
            uint32_t sel[8];
            {
                const uint4 sa = ((const uint4*)s_sel)[2 * x];
                sel[0] = sa.x; sel[1] = sa.y; sel[2] = sa.z; sel[3] = sa.w;
                if (NH > 4) { const uint4 sb = ((const uint4*)s_sel)[2 * x + 1]; sel[4] = sb.x; sel[5] = sb.y; sel[6] = sb.z; sel[7] = sb.w; }
            }
            // holes and max height of the env's board: sums / maxima over the packed bytes (every thread of the env derives them
            // itself -- a per-env scan by three of the ten warps plus a CTA barrier cost more than the instructions here)
            int holes0 = 0;
            uint32_t mx4 = V[1];
            {
                const uint32_t* how = (const uint32_t*)(s_hol + e * W4);
#pragma unroll
                for (int k = 0; k < W4 / 4; k++) holes0 = __dp4a(how[k], 0x01010101u, (uint32_t)holes0);   // bytes beyond W are zero
#pragma unroll
                for (int k = 2; k <= NH; k++) mx4 = bytemax_lt128(mx4, V[k]);   // heights sit in words 1 .. NH (pads are zero)
            }
            const int maxh0 = (int)max4bytes(mx4);
            // full rows = AND over ALL field columns of (column | piece bits): the columns left / right of the 4-column window
            // are the same for the four rotations (window columns without piece cells contribute themselves, wall columns ones)
            COLT LR = field;
#pragma unroll
            for (int c = 0; c < W; c++) {
                const COLT v = colp[P + c];
                LR &= ((unsigned)(c + P - x) < 4u) ? ~COLT(0) : v;
            }
            // rows that are full whatever the piece does (a poked board may hold them), and cells in row 0 under the window
            const COLT preFull = LR & cj[0] & cj[1] & cj[2] & cj[3];
            COLT top0 = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) top0 |= ((unsigned)(x + j - P) < (unsigned)W) ? cj[j] : COLT(0);
            const bool odd = ((uint32_t)top0 & 1u) != 0 || preFull != 0;
            uint32_t slowmask = 0;   // rotations that go to the dense second pass
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int rot = (rot0 + r) & 3;                         // cumulative rot90 presses (wrappers/grouped.py:153-154)
                const uint4 pr = s_prec[(piece * 4 + rot) * 8 + (e & 7)];
                const uint32_t bot4 = s_bot[piece * 4 + rot];
                // Landing row from the column heights: the piece comes down from row 0 through empty rows, so the first collision
                // is its lowest cell of some column j meeting that column's top cell, at y* = min_j (H - h_j - bot_j); it rests at
                // y = y* - 1 (wrappers/grouped.py:160-161: while !collision(y+1): y++ from y = 0).  Exact when the heights
                // are the true column tops (no cell in row 0 under the window: the heights have row 0 zeroed, Q1) and y* >= 1;
                // every other placement is evaluated from the column bitboards in the dense second pass.
                const uint32_t M4 = pr.y;
                const uint32_t OM = O4 & M4;
                const uint32_t Tb = OM + bot4;                          // all bytes < 128
                const int y = H - 1 - (int)max4bytes(Tb);
                const int mintop = (pr.w >> 28) & 3;
                uint32_t hw[NH];
                uint32_t summary;
                if (!((pr.w >> x) & 1u)) {
                    // collision_with_frame: ones board, row 0 zeroed -> heights H - 1, max H - 1, no holes, no bumpiness
#pragma unroll
                    for (int k = 0; k < NH; k++) hw[k] = (uint32_t)(H - 1) * 0x01010101u;
                    summary = (uint32_t)(H - 1);
                } else {
                    legal_w |= 1u << (8 * r);
#pragma unroll
                    for (int k = 0; k < NH; k++) hw[k] = 0;
                    summary = 0;                                        // game over (decided in the second pass): zeros board
                    {
                        // a row can only fill up where the columns outside the window are all set: LR bits y .. y + 3
                        if (odd || y < 0 || y + mintop == 0 || ((uint32_t)(LR >> (y & (8 * (int)sizeof(COLT) - 1))) & 15u) != 0) {
                            slowmask |= 1u << r;
                        } else {
                            // new window: columns under piece cells rise to H - y - top offset, the others keep their height
                            const uint32_t T4 = ((uint32_t)(H - y) * 0x01010101u - pr.z) & M4;
                            const uint32_t N4 = bytemax_lt128(O4, T4);
                            // all bytes < 128: signed dot products; sum(new) - sum(old) - 4 cells
                            const int holes = __dp4a((int)O4, (int)0xFFFFFFFFu /* 4 x -1 */, __dp4a((int)N4, 0x01010101, holes0 - 4));
                            const int maxh = max(maxh0, H - y - mintop);
#pragma unroll
                            for (int k = 0; k < NH; k++) hw[k] = prmt_raw(V[k + 1], N4, sel[k]);
                            // bumpiness over the packed heights (pairs (c, c+1), c < W - 1)
                            uint32_t bump = 0;
#pragma unroll
                            for (int k = 0; k < NH - 1; k++) bump = __vsadu4(hw[k], __funnelshift_r(hw[k], hw[k + 1], 8)) + bump;
                            {
                                const uint32_t l = hw[NH - 1];
                                // last word: VL valid heights; compare (b0,b1), .. (b[VL-2], b[VL-1]) only
                                const uint32_t a = VL == 4 ? l : __byte_perm(l, 0, VL == 1 ? 0x4444 : (VL == 2 ? 0x4410 : 0x4210));
                                const uint32_t b = __byte_perm(l, 0, VL == 1 ? 0x4444 : (VL == 2 ? 0x4411 : (VL == 3 ? 0x4221 : 0x3321)));
                                bump = __vsadu4(a, b) + bump;
                            }
                            summary = (uint32_t)maxh | ((uint32_t)(holes & 255) << 8) | ((bump & 255u) << 16);
                        }
                    }
                }
                if (r == 0) put_row<W, 0, F>(o, hw, summary);
                else if (r == 1) put_row<W, 1, F>(o, hw, summary);
                else if (r == 2) put_row<W, 2, F>(o, hw, summary);
                else put_row<W, 3, F>(o, hw, summary);
            }
            if (slowmask) {
                // one append per thread (its staged rows are already the zeros board).  Two lists in one array: placements that cannot
                // complete a row whatever their landing row (LR == 0: nearly all of them under random play) grow from the front,
                // the others from the back -- the second pass runs the full-row scan only for the back list
                const int cnt = __popc(slowmask);
                const bool back = LR != 0;
                const int p0 = atomicAdd(s_nslow + (back ? 1 : 0), cnt);
                int pos = back ? EPB * A - 1 - p0 : p0, dp = back ? -1 : 1;
#pragma unroll
                for (int r = 0; r < 4; r++)
                    if ((slowmask >> r) & 1u) { s_slow[pos] = (unsigned short)(e * A + 4 * xb + r); pos += dp; }
            }
        }
        uint32_t* dst = s_featw + (size_t)e * (W * F) + xb * F;
#pragma unroll
        for (int i = 0; i < F; i++) dst[i] = o[i];
        s_legalw[e * W + xb] = legal_w;
    }
    gf_sync<W, FUSED>();
    if (info_board && xb == W - 1 && live) {
        int maxh = 0, holes = 0, bump = 0, prev = 0;
#pragma unroll
        for (int c = 0; c < W; c++) {
            const int hgt = s_ih[e * W4 + c];
            s_info[e * F + c] = (uint8_t)hgt;
            holes += s_iho[e * W4 + c];
            maxh = max(maxh, hgt);
            if (c > 0) bump += abs(hgt - prev);
            prev = hgt;
        }
        s_info[e * F + W] = (uint8_t)maxh; s_info[e * F + W + 1] = (uint8_t)holes; s_info[e * F + W + 2] = (uint8_t)bump;   // uint8 wrap (Q4)
    }
    // ---- phase 4: placements that clear rows / touch the zeroed row, column by column (dense list) ----
    // Tetris.clear_filled_rows on the projected copy (wrappers/grouped.py:171-177): with F = the cleared rows a column keeps
    // its cells u = v & ~F in order, packed towards the floor: the top cell (row t = ctz(u)) ends at t + popc(F >> (t+1)) and
    // popc(u) cells remain.  After a clear row 0 is empty, so the wrapper's row zeroing (Q1) only applies when F = 0.
    {
        const int ns0 = s_nslow[0], ns = ns0 + s_nslow[1];
        uint8_t* s_feats = sm + S::off_feats;
        for (int k = tid; k < ns; k += T) {
            const bool scan = k >= ns0;                                 // back list: a row may fill up
            const int it = s_slow[scan ? EPB * A - 1 - (k - ns0) : k], es = it / A, a = it - es * A;
            const uint32_t w0 = s_w0[es];
            const int piece = (w0 >> 13) & 7, rot = (int)(((w0 >> 16) & 3) + (a & 3)) & 3;
            const int x = (a >> 2) + P - (int)((cfg.nhalf3 >> (3 * piece)) & 7u);
            const uint4 pr = s_prec[(piece * 4 + rot) * 8 + (tid & 7)];
            const COLT* colp = s_colp + es * CS;
            COLT B = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int c = (pr.x >> (4 * q)) & 15;
                B |= colp[x + (c & 3)] >> (c >> 2);
            }
            const int y = ctz_t<COLT>(B >> 1);                          // while !collision(y+1): y++ from y = 0
            if ((B >> y) & 1) continue;                                 // game over: the staged row is already the zeros board
            const int jmin = pr.w >> 30, c0 = x + jmin - P, c1 = x + ((31 - __clz((int)pr.y)) >> 3) - P;
            COLT full = 0;
            if (scan) {
                full = field;
                for (int c = 0; c < W; c++) {
                    COLT v = colp[P + c];
                    const int j = c + P - x;
                    if ((unsigned)j < 4u) v |= (COLT)((pr.x >> (16 + 4 * j)) & 15u) << y;
                    full &= v;
                }
            }
            uint8_t* out = s_feats + (size_t)it * F;
            if (full == 0) {
                // no row is cleared (the usual reason to be here: the piece lands in the top rows): only the window's columns change.
                // Their heights / holes come from the bitboards with row 0 zeroed (Q1); the rest of the row is the env's base vector.
                const COLT keep0 = ~COLT(1) & field;
                uint32_t N4 = 0;
                int holes = 0;
                {
                    const uint32_t* how = (const uint32_t*)(s_hol + es * W4);
#pragma unroll
                    for (int q = 0; q < W4 / 4; q++) holes = __dp4a(how[q], 0x01010101u, (uint32_t)holes);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = x + j - P;
                    if ((unsigned)c < (unsigned)W) {
                        const COLT u = (colp[x + j] | ((COLT)((pr.x >> (16 + 4 * j)) & 15u) << y)) & keep0;
                        const int hgt = u != 0 ? H - ctz_t<COLT>(u) : 0;
                        N4 |= (uint32_t)hgt << (8 * j);
                        holes += hgt - popc_t<COLT>(u) - (int)s_hol[es * W4 + c];
                    }
                }
                const uint32_t* hvw = s_hv + es * HS;
                uint32_t hw[NH];
#pragma unroll
                for (int q = 0; q < NH; q++) hw[q] = prmt_raw(hvw[q + 1], N4, s_sel[8 * x + q]);   // (window bytes beside the field are pad zeros in both)
                uint32_t bump = 0, mx4 = 0;
#pragma unroll
                for (int q = 0; q < NH - 1; q++) bump = __vsadu4(hw[q], __funnelshift_r(hw[q], hw[q + 1], 8)) + bump;
                {
                    const uint32_t l = hw[NH - 1];
                    const uint32_t a2 = VL == 4 ? l : __byte_perm(l, 0, VL == 1 ? 0x4444 : (VL == 2 ? 0x4410 : 0x4210));
                    const uint32_t b2 = __byte_perm(l, 0, VL == 1 ? 0x4444 : (VL == 2 ? 0x4411 : (VL == 3 ? 0x4221 : 0x3321)));
                    bump = __vsadu4(a2, b2) + bump;
                }
#pragma unroll
                for (int q = 0; q < NH; q++) mx4 = bytemax_lt128(mx4, hw[q]);   // (bytes beyond W in the last word are pad zeros)
                const uint32_t maxh = max4bytes(mx4);
#pragma unroll
                for (int c = 0; c < W; c++) out[c] = (uint8_t)(hw[c >> 2] >> (8 * (c & 3)));
                out[W] = (uint8_t)maxh; out[W + 1] = (uint8_t)holes; out[W + 2] = (uint8_t)bump;
                continue;
            }
            const COLT keep = ~full & field;
            int t_max = 0, t_hol = 0, t_bmp = 0, prev = 0;
#pragma unroll 2
            for (int c = 0; c < W; c++) {
                COLT v = colp[c + P];
                const int t = c - c0;
                if ((unsigned)t <= (unsigned)(c1 - c0)) v |= (COLT)((pr.x >> (16 + 4 * (jmin + t))) & 15u) << y;
                const COLT u = v & keep;
                int hgt = 0, hol = 0;
                if (u != 0) {
                    const int tp = ctz_t<COLT>(u);
                    hgt = H - tp - popc_t<COLT>((full >> tp) >> 1);
                    hol = hgt - popc_t<COLT>(u);
                }
                out[c] = (uint8_t)hgt;
                t_hol += hol; t_max = max(t_max, hgt);
                if (c > 0) t_bmp += abs(hgt - prev);
                prev = hgt;
            }
            out[W] = (uint8_t)t_max; out[W + 1] = (uint8_t)t_hol; out[W + 2] = (uint8_t)t_bmp;
        }
    }
    // ---- phase 5: the tile is contiguous in global memory: full tiles leave as TMA bulk copies issued by one thread ----
    {
        uint8_t* gf = feats + (size_t)base * A * F;
        uint8_t* gl = legal + (size_t)base * A;
        if (nv == EPB) {
            fence_async_smem();            // generic-proxy writes of this thread -> visible to the async proxy
            gf_sync<W, FUSED>();
            if (tid == 0) {
                bulk_s2g(gf, s_featw, (uint32_t)(EPB * A * F));
                bulk_s2g(gl, s_legalw, (uint32_t)(EPB * A));
                if (info_board) bulk_s2g(info_board + (size_t)base * F, s_info, (uint32_t)(EPB * F));
                bulk_commit();
                if (!FUSED) bulk_wait_read();   // shared memory must stay alive until the copies have read it
            }
        } else {
            gf_sync<W, FUSED>();
            const int words = nv * W * F;
            for (int i = tid; i < words; i += T) ((uint32_t*)gf)[i] = s_featw[i];
            if (tid < nv * W) ((uint32_t*)gl)[tid] = s_legalw[tid];
            if (info_board)
                for (int i = tid; i < nv * F; i += T) info_board[(size_t)base * F + i] = s_info[i];
        }
    }
}

template <int W, class COLT>
__global__ void __maxnreg__(W == 10 ? GF_REGS : 96) k_grouped_feats_x(const DevCfg cfg, int64_t n, const uint8_t* __restrict__ hot,
                                                            const uint8_t* __restrict__ board, uint8_t* __restrict__ feats,
                                                            uint8_t* legal, const uint8_t* __restrict__ fill_high,
                                                            uint8_t* __restrict__ info_board) {
    using S = GFeatsSmem<W, COLT>;
    extern __shared__ __align__(16) uint8_t sm[];
    __shared__ int s_nslow[2];
    __shared__ uint32_t s_bot[28];
    __shared__ __align__(16) uint32_t s_sel[(W + P) * 8];
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * S::EPB;
    const int nv = (int)min((int64_t)S::EPB, n - base);
    if (tid < 28 * 8) {
        uint4 v = (&c_prec[0][0])[tid >> 3];
        v.w = prec_w_for_width<W>(v.w);
        ((uint4*)(sm + S::off_prec))[tid] = v;
    }
    if (tid >= 256 && tid < 256 + 28) s_bot[tid - 256] = (&c_bot4[0][0])[tid - 256];
    build_sel_table<(W + 3) / 4>(s_sel, W + P, tid, 32 * W);
    // programmatic dependent launch (see k_step_ws): this grid may be scheduled while the placement step drains
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // (one tile per CTA: a persistent variant -- 4 CTAs per SM looping over tiles, tables staged once -- measured 15 % SLOWER)
    gfeats_tile<W, COLT, false>(cfg, sm, s_nslow, s_bot, s_sel, tid, base, nv, hot + base * 32, board + base * cfg.board_stride,
                                fill_high ? fill_high + base : nullptr, feats, legal, info_board);
}

// ---- fused placement step + enumeration (GroupedActionsObservations.step, wrappers/grouped.py:209-269, followed by the
// observation of the new state) -- one persistent kernel instead of k_step_ws<.., 2> + k_grouped_feats_x.
// CTA = W feature warps + NLW logic warps.  The logic warps (lane = env; warp lw takes every NLW-th tile of the CTA) run ahead:
// each waits for the TMA-staged records of its next tile (NS = 2 NLW stages), applies the placement (decode, legality, hard
// drop, commit, line clear, spawn, autoreset), writes the 5-tuple and sends the records back with bulk stores, while the
// feature warps enumerate the 4W placements of the tile before, straight from the staged records.  The feature kernel is bound
// by integer issue with a third of its issue slots idle (barrier phases); the ~40 M warp instructions of the step hide there,
// and the records are read from HBM once instead of twice.  Named barriers: 1 = feature warps, 2 + s = ready[s] (logic ->
// features), 2 + NS + s = consumed[s] (features -> logic: stage s may be reloaded).
template <int W, class COLT>
struct GFusedSmem {
    using S = GFeatsSmem<W, COLT>;
    static constexpr int NLW = 2, NS = 2 * NLW;
    static constexpr size_t off_tab = (S::bytes + 127) & ~size_t(127);        // rowbytes[112] u32, cells[28] u16 (+8 pad), n[7] i32
    static constexpr size_t off_bar = off_tab + 112 * 4 + 64 + 32;
    static constexpr size_t off_box = off_bar + 64;                           // [32] boxes (unused by the features), [NS][32] fill flags
    static constexpr size_t off_stage = (off_box + 32 * 4 + NS * 32 + 127) & ~size_t(127);
    static __host__ __device__ size_t stage_bytes(const DevCfg& d) { return (((size_t)32 * 32 + 127) & ~size_t(127)) + (((size_t)32 * d.board_stride + 16 + 127) & ~size_t(127)) + (((size_t)32 * d.rng_stride + 127) & ~size_t(127)); }
    static size_t bytes(const DevCfg& d) { return off_stage + NS * stage_bytes(d); }
    static constexpr int threads = 32 * (W + NLW);
};

#ifndef GFU_REGS
#define GFU_REGS 40   // 12 warps x 40 registers: 4 CTAs per SM
#endif
template <int W, class COLT, bool XT>
__global__ void __maxnreg__(GFU_REGS) k_grouped_step_feats(const __grid_constant__ StepParams p, uint8_t* __restrict__ feats, uint8_t* legal,
                                                          uint8_t* __restrict__ info_board) {
    using S = GFeatsSmem<W, COLT>;
    using FS = GFusedSmem<W, COLT>;
    constexpr int E = 32, TF = 32 * W, NLW = FS::NLW, NS = FS::NS;
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ int s_nslow[2];
    __shared__ uint32_t s_bot[28];
    __shared__ __align__(16) uint32_t s_sel[(W + P) * 8];
    const DevCfg& cfg = p.cfg;
    const int tid = threadIdx.x;
    const int BS = cfg.board_stride, RS = cfg.rng_stride;
    uint32_t* s_rowbytes = (uint32_t*)(sm + FS::off_tab);
    unsigned short* s_cells = (unsigned short*)(s_rowbytes + 112);
    int* s_n = (int*)(s_rowbytes + 112 + 16);
    uint64_t* bar = (uint64_t*)(sm + FS::off_bar);
    uint32_t* s_box = (uint32_t*)(sm + FS::off_box);
    uint8_t* s_fill = (uint8_t*)(s_box + 32);          // [NS][32]
    const size_t st_hot = ((size_t)E * 32 + 127) & ~size_t(127), st_brd = ((size_t)E * BS + 16 + 127) & ~size_t(127);
    const size_t st_all = FS::stage_bytes(cfg);
    uint8_t* stage0 = sm + FS::off_stage;

    if (tid < 28 * 8) {
        uint4 v = (&c_prec[0][0])[tid >> 3];
        v.w = prec_w_for_width<W>(v.w);
        ((uint4*)(sm + S::off_prec))[tid] = v;
    }
    if (tid >= 256 && tid < 256 + 28) s_bot[tid - 256] = (&c_bot4[0][0])[tid - 256];
    build_sel_table<(W + 3) / 4>(s_sel, W + P, tid, (int)blockDim.x);
    for (int i = tid; i < 112; i += blockDim.x) s_rowbytes[i] = (&c_rowbytes[0][0][0])[i];
    if (tid < 28) s_cells[tid] = (&c_cells[0][0])[tid];
    if (tid < 7) s_n[tid] = c_n[tid];
    if (tid == 0) {
        for (int s = 0; s < NS; s++) mbar_init(bar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int64_t ntiles = (p.n + E - 1) / E;
    const int64_t G = gridDim.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();

    if (tid >= TF) {
        // ===== logic warps: TMA producer + game logic, ahead of the feature warps =====
        const int lane = tid & 31, lw = (tid - TF) >> 5;
        Tabs tb;
        tb.cells = s_cells; tb.rowbytes = s_rowbytes; tb.n = s_n; tb.ptab = nullptr;
        auto issue_load = [&](int64_t tile, int s) {
            const int64_t base = tile * E;
            const int nv = (int)min((int64_t)E, p.n - base);
            uint8_t* st = stage0 + s * st_all;
            mbar_expect_tx(bar + s, (uint32_t)(nv * (32 + BS + RS)));
            bulk_g2s(st, p.hot + base * 32, (uint32_t)(nv * 32), bar + s);
            bulk_g2s(st + st_hot, p.board + base * BS, (uint32_t)(nv * BS), bar + s);
            bulk_g2s(st + st_hot + st_brd, p.rng + base * RS, (uint32_t)(nv * RS), bar + s);
        };
        if (lane == 0) {   // this warp's first two tiles: k = lw and k = lw + NLW
            if ((int64_t)blockIdx.x + lw * G < ntiles) issue_load((int64_t)blockIdx.x + lw * G, lw);
            if ((int64_t)blockIdx.x + (lw + NLW) * G < ntiles) issue_load((int64_t)blockIdx.x + (lw + NLW) * G, lw + NLW);
        }
        TileStats stt = {0, 0, 0, 0};
        for (int64_t k = lw; blockIdx.x + k * G < ntiles; k += NLW) {
            const int64_t tile = blockIdx.x + k * G, base = tile * E;
            const int s = (int)(k % NS);
            const int nv = (int)min((int64_t)E, p.n - base);
            uint8_t* st = stage0 + s * st_all;
            int action = 0;
            if (lane < nv) action = p.actions[base + lane];
            mbar_wait(bar + s, (uint32_t)((k / NS) & 1));
            uint32_t dirty = 0;
            if (lane < nv)
                dirty = logic_one_env<COLT, false, 2, XT>(p, tb, base + lane, lane, action, (uint32_t*)st, st + st_hot, st + st_hot + st_brd, s_box, stt, base + lane);
            s_fill[s * 32 + lane] = (uint8_t)((dirty >> 2) & 1u);
            const int ndirty = __popc(__ballot_sync(0xffffffffu, (dirty & 1u) != 0));
            fence_async_smem();
            __syncwarp();
            asm volatile("bar.arrive %0, %1;" ::"r"(2 + s), "n"(TF + 32) : "memory");   // ready[s]: the feature warps may read stage s
            // write-back: hot tile, board records (whole tile when most envs committed -- the rule in the grouped mode), rng records
            const bool whole = ndirty >= p.whole_tile_min;
            if (lane == 0) {
                bulk_s2g(p.hot + base * 32, st, (uint32_t)(nv * 32));
                if (whole) bulk_s2g(p.board + base * BS, st + st_hot, (uint32_t)(nv * BS));
            }
            if (lane < nv) {
                if ((dirty & 1) && !whole) bulk_s2g(p.board + (base + lane) * BS, st + st_hot + lane * BS, (uint32_t)BS);
                if (dirty & 2) bulk_s2g(p.rng + (base + lane) * RS, st + st_hot + st_brd + lane * RS, (uint32_t)RS);
            }
            bulk_commit();
            // stage s is free again once the feature warps have copied tile k's columns and the stores above have read it
            if (blockIdx.x + (k + NS) * G < ntiles) {
                asm volatile("bar.sync %0, %1;" ::"r"(2 + NS + s), "n"(TF + 32) : "memory");   // consumed[s]
                bulk_wait_read();
                __syncwarp();
                if (lane == 0) issue_load(blockIdx.x + (k + NS) * G, s);
            }
        }
        bulk_wait_all();
        if (p.stats) flush_stats(p.stats, stt.ep, stt.ret, stt.len, stt.lines);
    } else {
        // ===== feature warps =====
        for (int64_t k = 0; blockIdx.x + k * G < ntiles; k++) {
            const int64_t tile = blockIdx.x + k * G, base = tile * E;
            const int s = (int)(k % NS);
            const int nv = (int)min((int64_t)E, p.n - base);
            const uint8_t* st = stage0 + s * st_all;
            asm volatile("bar.sync %0, %1;" ::"r"(2 + s), "n"(TF + 32) : "memory");       // ready[s]
            if (tid == 0) bulk_wait_read();     // the previous tile's feature stores have left the staging buffers
            gfeats_tile<W, COLT, true>(cfg, sm, s_nslow, s_bot, s_sel, tid, base, nv, st, st + st_hot, s_fill + s * 32, feats, legal, info_board,
                                       blockIdx.x + (k + NS) * G < ntiles ? 2 + NS + s : 0);
        }
        if (tid == 0) bulk_wait_all();
    }
}

}  // namespace tg
