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
__constant__ uint4 c_prec[7][4];

__device__ __forceinline__ uint32_t bytemax_lt128(uint32_t a, uint32_t b) {   // bytewise max, all bytes < 128
    const uint32_t d = (a | 0x80808080u) - b;                    // bit 7 of byte i = (a_i >= b_i), no borrow between bytes
    const uint32_t m = ((d >> 7) & 0x01010101u) * 0xFFu;
    return (a & m) | (b & ~m);
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

template <int W, class COLT>
__global__ void __maxnreg__(W == 10 ? GF_REGS : 96) k_grouped_feats_x(const DevCfg cfg, int64_t n, const uint8_t* __restrict__ hot,
                                                            const uint8_t* __restrict__ board, uint8_t* __restrict__ feats,
                                                            uint8_t* legal, const uint8_t* __restrict__ fill_high,
                                                            uint8_t* __restrict__ info_board) {
    using S = GFeatsSmem<W, COLT>;
    constexpr int EPB = S::EPB, A = S::A, F = S::F, CS = S::CS, HW = S::HW, HS = S::HS;
    constexpr int NH = (W + 3) / 4, VL = W - 4 * (NH - 1), T = 32 * W;
    extern __shared__ __align__(16) uint8_t sm[];
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
    __shared__ int s_nslow;

    const int H = cfg.H, tid = threadIdx.x, e = tid & 31, xb = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * EPB;
    const int nv = (int)min((int64_t)EPB, n - base);
    const bool live = e < nv;
    const COLT field = (COLT(1) << H) - 1;

    // ---- phase 1: thread = (column xb, env e): column -> shared memory, its height / holes with row 0 zeroed (Q1) ----
    if (tid < 28 * 8) s_prec[tid] = (&c_prec[0][0])[tid >> 3];
    if (tid == 0) s_nslow = 0;
    // programmatic dependent launch (see k_step_ws): this grid may be scheduled while the placement step drains
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (live) {
        const COLT col = ((const COLT*)(board + (base + e) * cfg.board_stride))[xb];
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
            s_w0[e] = (*(const uint32_t*)(hot + (base + e) * 32) & 0x7FFFFFFFu) | ((fill_high && fill_high[base + e]) ? 0x80000000u : 0u);
    }
    __syncthreads();
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
            const uint32_t Wlo = hvw[wi], Whi = hvw[wi + 1];
            const uint32_t O4 = __funnelshift_r(Wlo, Whi, sh);
            // holes and max height of the env's board: sums / maxima over the packed bytes (every thread of the env derives them
            // itself -- a per-env scan by three of the ten warps plus a CTA barrier cost more than the ~25 instructions here)
            int holes0 = 0;
            uint32_t mx4 = 0;
            {
                const uint32_t* how = (const uint32_t*)(s_hol + e * W4);
#pragma unroll
                for (int k = 0; k < W4 / 4; k++) holes0 = __dp4a(how[k], 0x01010101u, (uint32_t)holes0);   // bytes beyond W are zero
#pragma unroll
                for (int k = 0; k < HW; k++) mx4 = __vmaxu4(mx4, V[k]);
            }
            const int maxh0 = (int)max(max(mx4 & 255u, (mx4 >> 8) & 255u), max((mx4 >> 16) & 255u, mx4 >> 24));
            // full rows = AND over ALL field columns of (column | piece bits): the columns left / right of the 4-column window
            // are the same for the four rotations (window columns without piece cells contribute themselves, wall columns ones)
            COLT LR = field;
#pragma unroll
            for (int c = 0; c < W; c++) {
                const COLT v = colp[P + c];
                LR &= ((unsigned)(c + P - x) < 4u) ? ~COLT(0) : v;
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int rot = (rot0 + r) & 3;                         // cumulative rot90 presses (wrappers/grouped.py:153-154)
                const uint4 pr = s_prec[(piece * 4 + rot) * 8 + (e & 7)];
                COLT B = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int c = (pr.x >> (4 * k)) & 15;
                    B |= colp[x + (c & 3)] >> (c >> 2);
                }
                const int y = ctz_t<COLT>(B >> 1);                      // while !collision(y+1): y++ from y = 0
                const int jmin = pr.w & 3, jmax = (pr.w >> 2) & 3, mintop = (pr.w >> 4) & 3;
                const int c0 = x + jmin - P, c1 = x + jmax - P;
                uint32_t hw[NH];
                uint32_t summary;
                if (c0 < 0 || c1 >= W) {
                    // collision_with_frame: ones board, row 0 zeroed -> heights H - 1, max H - 1, no holes, no bumpiness
#pragma unroll
                    for (int k = 0; k < NH; k++) hw[k] = (uint32_t)(H - 1) * 0x01010101u;
                    summary = (uint32_t)(H - 1);
                } else {
                    legal_w |= 1u << (8 * r);
#pragma unroll
                    for (int k = 0; k < NH; k++) hw[k] = 0;
                    summary = 0;                                        // game over: zeros board
                    if (!((B >> y) & 1)) {
                        COLT full = LR;
#pragma unroll
                        for (int j = 0; j < 4; j++) full &= cj[j] | ((COLT)((pr.x >> (16 + 4 * j)) & 15u) << y);
                        if (full != 0 || y + mintop == 0) {
                            s_slow[atomicAdd(&s_nslow, 1)] = (unsigned short)(e * A + 4 * xb + r);
                        } else {
                            const uint32_t M4 = pr.y;
                            const uint32_t OM = O4 & M4;
                            const uint32_t T4 = ((uint32_t)(H - y) * 0x01010101u - pr.z) & M4;
                            const uint32_t N4 = bytemax_lt128(OM, T4);
                            // all bytes < 128: signed dot products; sum(new) - sum(old) - 4 cells
                            const int holes = __dp4a((int)OM, (int)0xFFFFFFFFu /* 4 x -1 */, __dp4a((int)N4, 0x01010101, holes0 - 4));
                            const int maxh = max(maxh0, H - y - mintop);
                            const uint32_t Nlo = N4 << sh, Nhi = __funnelshift_l(N4, 0u, sh);
                            const uint32_t Mlo = M4 << sh, Mhi = __funnelshift_l(M4, 0u, sh);
                            const uint32_t Wl = (Wlo & ~Mlo) | Nlo, Wh = (Whi & ~Mhi) | Nhi;
#pragma unroll
                            for (int k = 0; k < NH; k++) hw[k] = (k + 1 == wi) ? Wl : ((k == wi) ? Wh : V[k + 1]);
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
        }
        uint32_t* dst = s_featw + (size_t)e * (W * F) + xb * F;
#pragma unroll
        for (int i = 0; i < F; i++) dst[i] = o[i];
        s_legalw[e * W + xb] = legal_w;
    }
    __syncthreads();
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
        const int ns = s_nslow;
        uint8_t* s_feats = sm + S::off_feats;
        for (int k = tid; k < ns; k += T) {
            const int it = s_slow[k], es = it / A, a = it - es * A;
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
            const int y = ctz_t<COLT>(B >> 1);
            const int jmin = pr.w & 3, c0 = x + jmin - P, c1 = x + (int)((pr.w >> 2) & 3) - P;
            COLT full = field;
            for (int c = 0; c < W; c++) {
                COLT v = colp[P + c];
                const int j = c + P - x;
                if ((unsigned)j < 4u) v |= (COLT)((pr.x >> (16 + 4 * j)) & 15u) << y;
                full &= v;
            }
            const COLT keep = (full != 0 ? ~full : ~COLT(1)) & field;
            uint8_t* out = s_feats + (size_t)it * F;
            int s_max = 0, s_hol = 0, s_bmp = 0, prev = 0;
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
                s_hol += hol; s_max = max(s_max, hgt);
                if (c > 0) s_bmp += abs(hgt - prev);
                prev = hgt;
            }
            out[W] = (uint8_t)s_max; out[W + 1] = (uint8_t)s_hol; out[W + 2] = (uint8_t)s_bmp;
        }
    }
    // ---- phase 5: the tile is contiguous in global memory: full tiles leave as TMA bulk copies issued by one thread ----
    {
        uint8_t* gf = feats + (size_t)base * A * F;
        uint8_t* gl = legal + (size_t)base * A;
        if (nv == EPB) {
            fence_async_smem();            // generic-proxy writes of this thread -> visible to the async proxy
            __syncthreads();
            if (tid == 0) {
                bulk_s2g(gf, s_featw, (uint32_t)(EPB * A * F));
                bulk_s2g(gl, s_legalw, (uint32_t)(EPB * A));
                if (info_board) bulk_s2g(info_board + (size_t)base * F, s_info, (uint32_t)(EPB * F));
                bulk_commit();
                bulk_wait_read();          // shared memory must stay alive until the copies have read it
            }
        } else {
            __syncthreads();
            const int words = nv * W * F;
            for (int i = tid; i < words; i += T) ((uint32_t*)gf)[i] = s_featw[i];
            if (tid < nv * W) ((uint32_t*)gl)[tid] = s_legalw[tid];
            if (info_board)
                for (int i = tid; i < nv * F; i += T) info_board[(size_t)base * F + i] = s_info[i];
        }
    }
}

}  // namespace tg
