// tg_stepn.cuh -- K consecutive steps of a SMALL batch in one launch (tg_step_n; BASELINE config 2 at 4 K .. 128 K envs).
//
// A per-call step of a few thousand envs is bound by everything but the work: launch, prologue, one TMA round trip for 6 KB of
// state per tile, drain (12 us per call at 4,096 envs for 2.5 us of logic + image work).  Envs are independent, so nothing forces
// a trip through HBM between two steps: here every CTA keeps the records of ITS tiles resident in shared memory for all K steps
// (loaded once, written back once), reads the step's actions, and writes the observation dict + 5-tuple of every step.
//   * NL logic warps: game logic, lane = env; warp w owns the resident tiles j = w, w + NL, ... and walks (step, own tile) in order;
//   * the other warps: observation images of the whole (step, tile) sequence, behind the logic, TMA bulk stores;
//   * the roles meet on counters in shared memory (one per logic warp + one for the image warps): a tile's records are stepped
//     again only after the image warps have finished expanding them (with a single resident tile per CTA the roles alternate;
//     with several they overlap).
// HBM traffic per env-step: observation dict + 10 B of outputs + 4 B of action; the state traffic is amortised over K.
// Outputs: obs / 5-tuple arrays of step k start at element offset k * obs_stride (in envs; 0 = every step overwrites the same
// arrays, n = [K][n] rollout storage).
#pragma once
#include "tg_step.cuh"

namespace tg {

struct StepNParams {
    StepParams sp;          // cfg, state pointers, output base pointers, E, image / table offsets (off_hot / off_brd / off_rng = slot arrays)
    int K;
    int TL;                 // resident tile slots per CTA
    int NL;                 // logic warps
    int64_t obs_stride;     // envs between the observation arrays of consecutive steps
    int64_t out_stride;     // envs between the 5-tuple arrays of consecutive steps
    int off_cnt;            // NL + 1 counters
    int off_dirty;          // [TL][E] accumulated dirty flags
};

__device__ __forceinline__ uint32_t ld_volatile_s(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_s(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

template <int WT, int HT, class COLT>
__global__ void __launch_bounds__(256) k_step_resident(const __grid_constant__ StepNParams q) {
    extern __shared__ __align__(128) uint8_t smem[];
    const StepParams& p = q.sp;
    const DevCfg& cfg = p.cfg;
    const int E = p.E, T = blockDim.x, tid = threadIdx.x;
    const int NL = q.NL;
    const int FT = T - 32 * NL, ft = tid - 32 * NL;
    const int W = WT ? WT : cfg.W, H = HT ? HT : cfg.H;
    const int Wp = W + 2 * P, Hp = H + P;
    const int OB = Hp * Wp, OQ = cfg.OQ, BS = cfg.board_stride, RS = cfg.rng_stride;
    const int BAR_FILL = 1;

    uint8_t* i_board = smem + p.off_iboard;
    uint8_t* i_mask = smem + p.off_imask;
    uint8_t* i_holder = smem + p.off_iholder;
    uint8_t* i_queue = smem + p.off_iqueue;
    uint64_t* bar = (uint64_t*)(smem + p.off_bar);
    uint32_t* s_boxes = (uint32_t*)(smem + p.off_box);            // [TL][E] boxes, then [E] boxes of the previous item
    uint32_t* s_boxprev = s_boxes + q.TL * E;
    uint32_t* s_dirty = (uint32_t*)(smem + q.off_dirty);          // [TL][E]
    uint32_t* s_cnt = (uint32_t*)(smem + q.off_cnt);              // [w] own items finished by logic warp w, [NL] items finished by the image warps
    uint32_t* s_rowbytes = (uint32_t*)(smem + p.off_tab);
    unsigned short* s_cells = (unsigned short*)(s_rowbytes + 112);
    int* s_n = (int*)(s_rowbytes + 112 + 16);
    Tabs tb;
    tb.cells = s_cells; tb.rowbytes = s_rowbytes; tb.n = s_n;

    const int64_t ntiles = (p.n + E - 1) / E;
    const int64_t G = gridDim.x;
    int nt = 0;                                                    // tiles of this CTA: blockIdx.x + j * G
    for (int j = 0; j < q.TL; j++) if ((int64_t)blockIdx.x + j * G < ntiles) nt = j + 1;

    for (int i = tid; i < (q.TL + 1) * E; i += T) s_boxes[i] = 0;       // boxes of the resident tiles + boxes of the previous item
    for (int i = tid; i < q.TL * E; i += T) s_dirty[i] = 0;
    if (tid == 0) {
        for (int w = 0; w <= NL; w++) s_cnt[w] = 0;
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    init_cta(E, W, H, s_rowbytes, s_cells, s_n, i_board, i_mask, tid, T);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    if (ft == 0) {   // all resident tiles with one barrier
        uint32_t bytes = 0;
        for (int j = 0; j < nt; j++) {
            const int64_t base = ((int64_t)blockIdx.x + j * G) * E;
            bytes += (uint32_t)((int)min((int64_t)E, p.n - base) * (32 + BS + RS));
        }
        mbar_expect_tx(bar, bytes);
        for (int j = 0; j < nt; j++) {
            const int64_t base = ((int64_t)blockIdx.x + j * G) * E;
            const int nv = (int)min((int64_t)E, p.n - base);
            bulk_g2s(smem + p.off_hot + j * p.st_hot, p.hot + base * 32, (uint32_t)(nv * 32), bar);
            bulk_g2s(smem + p.off_brd + j * p.st_brd, p.board + base * BS, (uint32_t)(nv * BS), bar);
            bulk_g2s(smem + p.off_rng + j * p.st_rng, p.rng + base * RS, (uint32_t)(nv * RS), bar);
        }
    }
    const int items = q.K * nt;

    if (tid < 32 * NL) {
        // ===== logic warps =====
        const int lw = tid >> 5, lane = tid & 31;
        TileStats st = {0, 0, 0, 0};
        mbar_wait(bar, 0);
        uint32_t own = 0;
        for (int k = 0; k < q.K; k++) {
            for (int j = lw; j < nt; j += NL) {
                const int64_t base = ((int64_t)blockIdx.x + j * G) * E;
                const int nv = (int)min((int64_t)E, p.n - base);
                int action = 0;
                if (lane < nv) action = p.actions[(int64_t)k * p.n + base + lane];
                // the image warps must be done with this slot's records of the previous step: item (k - 1, j)
                if (k > 0) { const uint32_t need = (uint32_t)((k - 1) * nt + j + 1); while (ld_volatile_s(s_cnt + NL) < need) {} }
                uint32_t dirty = 0;
                if (lane < nv)
                    dirty = logic_one_env<COLT, false, 0>(p, tb, base + lane, lane, action, (uint32_t*)(smem + p.off_hot + j * p.st_hot),
                                                          smem + p.off_brd + j * p.st_brd, smem + p.off_rng + j * p.st_rng, s_boxes + j * E, st,
                                                          (int64_t)k * q.out_stride + base + lane);
                if (lane < E) s_dirty[j * E + lane] |= dirty;
                __syncwarp();
                __threadfence_block();
                own++;
                if (lane == 0) st_volatile_s(s_cnt + lw, own);
            }
        }
        if (p.stats) flush_stats(p.stats, st.ep, st.ret, st.len, st.lines);
    } else {
        // ===== image / store warps =====
        const bool leader = (ft == 0);
        mbar_wait(bar, 0);                                  // acquire the TMA writes of the resident records
        const bool every_step = q.obs_stride != 0 || q.K == 1;   // obs_stride 0: one set of arrays, only the last step's dict is kept
        int nv_prev = 0, i = 0;
        const int own_q = nt / NL, own_r = nt - own_q * NL;   // logic warp w owns own_q + (w < own_r) of the CTA's tiles
        for (int k = 0; k < q.K; k++) {
            int ow = 0, jq = 0;                              // j % NL and j / NL, kept incrementally (three divisions per item otherwise)
            for (int j = 0; j < nt; j++, i++, jq += (ow + 1 == NL), ow = (ow + 1 == NL ? 0 : ow + 1)) {
                const int64_t base = ((int64_t)blockIdx.x + j * G) * E;
                const int nv = (int)min((int64_t)E, p.n - base);
                const int64_t ob = (int64_t)k * q.obs_stride + base;
                // item (k, j) belongs to logic warp j % NL and is its (k * own tiles + j / NL)-th item
                const uint32_t need = (uint32_t)(k * (own_q + (ow < own_r ? 1 : 0)) + jq + 1);
                if (!every_step && k != q.K - 1) {          // nothing to emit: only release the slot (the box stays with the last emitted item)
                    if (leader) { while (ld_volatile_s(s_cnt + ow) < need) {} st_volatile_s(s_cnt + NL, (uint32_t)(i + 1)); }
                    continue;
                }
                if (leader) { while (ld_volatile_s(s_cnt + ow) < need) {} __threadfence_block(); }
                bulk_wait_read();                           // stores of the previous item have left the image buffers
                named_sync(BAR_FILL, FT);
                __threadfence_block();
                const uint32_t* s_hot = (const uint32_t*)(smem + p.off_hot + j * p.st_hot);
                const uint8_t* s_brd = smem + p.off_brd + j * p.st_brd;
                mask_clear_boxes(s_boxprev, nv_prev, i_mask, OB, Wp, ft, FT);
                fill_images<WT, HT>(cfg, nv, s_hot, s_brd, s_rowbytes, i_board, i_holder, i_queue, ft, FT);
                named_sync(BAR_FILL, FT);
                mask_set_and_overlay(s_boxes + j * E, nv, s_cells, i_board, i_mask, OB, Wp, ft, FT);
                for (int t = ft; t < nv; t += FT) s_boxprev[t] = s_boxes[j * E + t];
                fence_async_smem();
                named_sync(BAR_FILL, FT);
                if (leader) st_volatile_s(s_cnt + NL, (uint32_t)(i + 1));    // the records of slot j may be stepped again
                tile_store<true>(p.o_board + ob * OB, i_board, (uint32_t)(nv * OB), leader, ft, FT);
                tile_store<true>(p.o_mask + ob * OB, i_mask, (uint32_t)(nv * OB), leader, ft, FT);
                if (leader) {
                    bulk_s2g_stream(p.o_holder + ob * cfg.OH, i_holder, (uint32_t)(nv * cfg.OH));
                    bulk_s2g_stream(p.o_queue + ob * OQ, i_queue, (uint32_t)(nv * OQ));
                }
                bulk_commit();
                nv_prev = nv;
            }
        }
        bulk_wait_all();
    }
    (void)items;
    // state write-back, once: hot records always, board / rng records of the tiles that changed (the logic warps order their
    // generic-proxy writes before the bulk stores: fence, then the barrier, then one thread issues)
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        for (int j = 0; j < nt; j++) {
            const int64_t base = ((int64_t)blockIdx.x + j * G) * E;
            const int nv = (int)min((int64_t)E, p.n - base);
            uint32_t any = 0;
            for (int t = 0; t < nv; t++) any |= s_dirty[j * E + t];
            bulk_s2g(p.hot + base * 32, smem + p.off_hot + j * p.st_hot, (uint32_t)(nv * 32));
            if (any & 1) bulk_s2g(p.board + base * BS, smem + p.off_brd + j * p.st_brd, (uint32_t)(nv * BS));
            if (any & 2) bulk_s2g(p.rng + base * RS, smem + p.off_rng + j * p.st_rng, (uint32_t)(nv * RS));
        }
        bulk_commit();
        bulk_wait_all();
    }
}

}  // namespace tg
