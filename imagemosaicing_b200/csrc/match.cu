// match.cu — K2: exact brute-force L2 1-NN of SIFT-128 descriptors as a tcgen05 distance-GEMM with a
// fused arg-min epilogue.  Replaces FlannBasedMatcher().match (M/MosaicWithoutPos.cpp:5108-5110).
//
//   |a_i - b_j|^2 = |a_i|^2 + |b_j|^2 - 2 a_i.b_j ;  the u8 x u8 -> s32 dot products come from
//   tcgen05.mma.kind::i8 (exact), accumulators live in TMEM, and the M x N distance matrix is never
//   written anywhere: each epilogue thread owns one query row and keeps a running (min key, arg-min).
//
// Work item  = 256 query rows of one pair (two M=128 accumulator row blocks) x all train tiles (N=128).
// CTA layout = 18 warps, persistent, one CTA per SM:
//   warp 0      TMEM allocator (512 columns = 2 accumulator buffers x 2 row blocks x 128 columns), then TMA producer (A block once
//               per item; B tiles through a kStages-deep smem ring, each with its 128 column keys)
//   warp 1      MMA issuer    (one elected lane; 2 row blocks x 4 K-chunks of 32 per train tile)
//   warps 2-17  epilogue      (two groups of 8, one per TMEM buffer; warp w reads TMEM lanes 32*(w%4).., 64 columns of BOTH row
//               blocks in two passes of 32)
// Pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), A full/empty.
//
// What bounds it — measured, in this order of discovery:
//  * NOT the TMEM read port: tcgen05.ld.32x32b.x32 sustains 39 / 68 / 87 words per clock per SM from 4 / 8 / 16 warps
//    (scripts/microbench/ldtm.cu, profiles/r2_ldtm_microbench.txt); the kernel needs 32 768 words per tile.
//  * The MMA pipeline of THIS configuration (cta_group::1, M=128 N=128 K=32, both operands from shared memory): with the epilogue
//    reduced to "drain TMEM, hand the buffer back" (UAVM_K2_DBG=2) TMA + MMA alone take 0.354 ms per 49 pairs = 1 000 cycles per
//    256 x 128 x 128 tile = 4 190 MAC / clk / SM = 2.38 Pop/s — one kind::i8 MMA every ~125 cycles, where the data-sheet int8
//    rate would allow 64.  At that rate an SS-mode MMA reads 8 KB of operands per 64 cycles = the whole 128 B/clk of shared memory,
//    next to the TMA writes; cuBLASLt's int8 GEMM (2-CTA MMA, operand sharing) reaches 2.83 Pop/s on the same box, bf16 1.60
//    (scripts/microbench/int8_gemm_probe.py).  ncu's sm__pipe_tensor_cycles_active (40 %) does not show this.  Not pursued further:
//    the epilogue below is the next limiter anyway.
//  * The epilogue on top of it: 16 warps (four per scheduler).  First version: one row block and 64 columns per warp, column
//    keys fetched with LDG (2 057 cycles per tile); then keys delivered to shared memory with the B tile (cp.async.bulk on the
//    same mbarrier): 1 429.  ncu then showed the LSU data pipe 80 % busy — a broadcast key costs one wavefront per key and
//    warp — and the accumulator buffer held until the reduction was done.  Now a warp reads 32 columns of BOTH row blocks (half
//    the key loads), hands the TMEM buffer back as soon as its tcgen05.ld has landed, prefetches the keys one step ahead and
//    reduces with a 3-input max of 64 a.b + ckey (keys stored negated): 1 270 cycles per tile, 0.449 ms per 49 pairs.  With all
//    16 warps in lockstep on one tile the TMEM drain (~380 cycles at 87 words/clk) and the reduction (~600, ALU pipe) of a tile
//    are serialised; two groups of 8 warps, one per TMEM buffer, overlap them: 1 200 cycles per tile, 0.425 ms per 49 pairs
//    (0.70 in round 1) = 1.98 PFLOP/s = 0.82 of what the MMA pipeline above delivers alone; configs[3] 0.154 ms.
//    Measured and rejected: one mbarrier pair per row block (0.199 vs 0.181 ms on configs[3]); 8- and 16-column software
//    pipelines of the TMEM loads inside a warp (0.470 ms: they delay the buffer hand-over); 18 instead of 20 warps alone -5 %.
//
#include <stdlib.h>
#include "internal.h"
#include "ptx.cuh"

using namespace uavm::ptx;

namespace {

constexpr int kStages = 6;            // B ring depth (16 KB per stage)
constexpr int kTileN = 128;           // train rows per tile
constexpr int kBlockM = 256;          // query rows per work item
constexpr int kEpiWarps = 16;
constexpr int kCtlWarps = 2;           // warp 0: TMEM allocator + TMA producer, warp 1: barrier init + MMA issuer
constexpr int kThreads = 32 * (kCtlWarps + kEpiWarps);      // 576: 112 registers per thread (640 threads capped the epilogue at 96)
constexpr uint32_t kABytes = kBlockM * 128;
constexpr uint32_t kBBytes = kTileN * 128;
constexpr uint32_t kKeyBytes = kTileN * 4;
// The producer runs at most kStages tiles ahead of the MMA and the MMA at most 2 tiles (TMEM buffers) ahead of the epilogue:
// a key ring of 16 tiles is never overwritten before the epilogue has read it.
constexpr int kKeyStages = 16;
constexpr size_t kSmemBytes = 1024 /*align slack*/ + kABytes + kStages * kBBytes + kKeyStages * kKeyBytes + 2 * kBlockM * 4 * 8 /*per-slice results*/ + 256 /*barriers*/;

struct __align__(8) Barriers {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t a_full;
    uint64_t a_empty;
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1)
k2_match_tcgen05(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_t,
                 const int32_t* __restrict__ ckey, const int32_t* __restrict__ norm,
                 const MatchItem* __restrict__ items, int n_items,
                 int32_t* __restrict__ out_idx, int32_t* __restrict__ out_d2, int dbg)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kABytes;
    int32_t* smem_key = reinterpret_cast<int32_t*>(smem + kABytes + kStages * kBBytes);                   // [kKeyStages][128]
    int2* smem_res = reinterpret_cast<int2*>(smem + kABytes + kStages * kBBytes + kKeyStages * kKeyBytes); // [2 (item parity)][256 rows][2 halves]
    Barriers* bar = reinterpret_cast<Barriers*>(smem + kABytes + kStages * kBBytes + kKeyStages * kKeyBytes + 2 * kBlockM * 4 * 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmap_q);
        prefetch_tensormap(&tmap_t);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
        mbar_init(&bar->a_full, 1);
        mbar_init(&bar->a_empty, 1);
        for (int b = 0; b < 2; b++) { mbar_init(&bar->tmem_full[b], 1); mbar_init(&bar->tmem_empty[b], kEpiWarps / 2); }
        fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(&bar->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bar->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t fill = 0, it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
                const MatchItem w = items[item];
                mbar_wait(&bar->a_empty, (it & 1) ^ 1);            // previous item's MMAs are done with A
                mbar_arrive_expect_tx(&bar->a_full, kABytes);
                tma_load_2d(smem_a, &tmap_q, &bar->a_full, 0, w.q_row);
                for (int t = 0; t < w.n_tiles; t++, fill++) {
                    const uint32_t s = fill % kStages;
                    mbar_wait(&bar->empty[s], ((fill / kStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bar->full[s], kBBytes + kKeyBytes);
                    tma_load_2d(smem_b + s * kBBytes, &tmap_t, &bar->full[s], 0, w.t_row + t * kTileN);
                    bulk_load_1d(smem_key + (fill % kKeyStages) * kTileN, ckey + w.t_row + t * kTileN, kKeyBytes, &bar->full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_i8(128, kTileN);
            const uint32_t a_addr = smem_u32(smem_a);
            const uint32_t b_addr = smem_u32(smem_b);
            uint32_t cons = 0, acc = 0, it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
                const int n_tiles = items[item].n_tiles;
                mbar_wait(&bar->a_full, it & 1);
                for (int t = 0; t < n_tiles; t++, cons++, acc++) {
                    const uint32_t s = cons % kStages;
                    const uint32_t buf = acc & 1;
                    mbar_wait(&bar->tmem_empty[buf], ((acc >> 1) & 1) ^ 1);   // epilogue drained this buffer
                    mbar_wait(&bar->full[s], (cons / kStages) & 1);           // TMA landed this B tile
                    tc_fence_after();
#pragma unroll
                    for (int rb = 0; rb < 2; rb++) {
                        const uint32_t d_tmem = tmem_base + buf * 256 + rb * 128;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const uint64_t da = make_smem_desc_sw128(a_addr + rb * (128 * 128) + k * 32);
                            const uint64_t db = make_smem_desc_sw128(b_addr + s * kBBytes + k * 32);
                            mma_i8_ss(d_tmem, da, db, idesc, k > 0 ? 1u : 0u);
                        }
                    }
                    tc_commit(&bar->empty[s]);          // smem slot reusable once these MMAs complete
                    tc_commit(&bar->tmem_full[buf]);    // accumulators ready for the epilogue
                }
                tc_commit(&bar->a_empty);               // A block reusable
            }
        }
        __syncwarp();
    } else if (warp >= kCtlWarps) {
        // ===================== epilogue: TMEM -> registers -> running arg-min =====================
        // Two groups of 8 warps, one per TMEM buffer: group g reduces the tiles that land in buffer g (the MMA warp alternates
        // buffers tile by tile, across items), so while one group drains its buffer the other one is computing on the previous tile; with all 16 warps in lockstep on one tile the TMEM drain (~380 cycles) and the
        // reduction (~600) of every tile were serialised.
        // This warp: TMEM lane quarter `quarter` (a warp may only read lanes 32 (warp % 4) ..), columns [64 half, 64 half + 64) of
        // BOTH row blocks in two passes of 32 columns — one key load serves two accumulators (a broadcast key costs one LSU
        // wavefront per key and warp: 80 % LSU utilisation when a warp owned one row block).
        const int ew = warp - kCtlWarps;
        const int quarter = warp & 3;
        const int g = ew >> 3;
        const int half = (ew >> 2) & 1;
        const int slice = 2 * g + half;
        const uint32_t col0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + g * 256 + half * 64;
        const uint32_t key_base = smem_u32(smem_key) + half * 64 * 4;
        const int r0 = quarter * 32 + lane;               // this thread's row inside a row block
        uint32_t use = 0, tile_base = 0, it = 0;          // uses of buffer g so far; tiles of earlier items (key ring position)
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
            const MatchItem w = items[item];
            int bk0 = 0x7fffffff, bj0 = 0, bk1 = 0x7fffffff, bj1 = 0;              // running (distance part, column) of row blocks 0 / 1
            for (int t = (g ^ (int)tile_base) & 1; t < w.n_tiles; t += 2, use++) {      // tile t of this item lands in buffer (tile_base + t) & 1
                const uint32_t kaddr = key_base + ((tile_base + t) % kKeyStages) * kKeyBytes;
                // the keys of this tile landed with its B tile (same mbarrier), which the MMA thread observed before it issued the
                // MMAs whose completion tmem_full signals.  (Waiting on full[] here as well would be wrong: the producer may
                // already have re-armed that stage twice, and a parity wait cannot tell phase k from phase k + 2.)
                mbar_wait(&bar->tmem_full[g], use & 1);
                tc_fence_after();
#pragma unroll
                for (int pass = 0; pass < 2; pass++) {
                    uint32_t v0[32], v1[32];
                    tmem_ld_32x32b_x32(col0 + 32 * pass, v0);                     // row block 0
                    tmem_ld_32x32b_x32(col0 + 32 * pass + 128, v1);               // row block 1, same train columns
                    int4 kc = lds128(kaddr + 128 * pass);                         // first four keys while the loads fly
                    tmem_ld_wait_on(v0, v1);
                    if (pass == 1) {
                        // all of this warp's accumulators are in registers: hand the buffer back BEFORE reducing them — the MMA of
                        // the tile after next needs it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bar->tmem_empty[g]);
                    }
                    pin_below(v0, v1);
                    if (dbg) { bk0 ^= (int)(v0[7] ^ v1[3]); continue; }            // timing experiments only (results are wrong)
                    // q = 64 a.b + ckey = -(32 (|b|^2 - 2 a.b) + (j & 31)): the maximum over the 32 columns is the nearest column
                    // and, among equals, the lowest one
                    int m0 = (int)0x80000000, m1 = (int)0x80000000;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int4 c = kc;
                        if (i < 7) kc = lds128(kaddr + 128 * pass + 16 * (i + 1)); // next four keys one step ahead of their use
                        m0 = __vimax3_s32(m0, 64 * (int)v0[4 * i + 0] + c.x, 64 * (int)v0[4 * i + 1] + c.y);
                        m1 = __vimax3_s32(m1, 64 * (int)v1[4 * i + 0] + c.x, 64 * (int)v1[4 * i + 1] + c.y);
                        m0 = __vimax3_s32(m0, 64 * (int)v0[4 * i + 2] + c.z, 64 * (int)v0[4 * i + 3] + c.w);
                        m1 = __vimax3_s32(m1, 64 * (int)v1[4 * i + 2] + c.z, 64 * (int)v1[4 * i + 3] + c.w);
                    }
                    // -m = 32 (|b|^2 - 2 a.b) + (j & 31) of the winning column: distance part and in-chunk index.  Strict <: an
                    // earlier chunk keeps the row on equal distance.
                    const int P0 = -m0, P1 = -m1;
                    const int k0 = P0 >> 5, k1 = P1 >> 5;
                    const int jb = t * kTileN + half * 64 + 32 * pass;
                    if (k0 < bk0) { bk0 = k0; bj0 = jb + (P0 & 31); }
                    if (k1 < bk1) { bk1 = k1; bj1 = jb + (P1 & 31); }
                }
            }
            tile_base += w.n_tiles;
            // merge the four slices of a row (tile parity x column half): smaller key wins, equal keys keep the lower column (the
            // slices visit disjoint columns, each in ascending order with a strict <)
            int2* res = smem_res + (it & 1) * (kBlockM * 4);
            res[r0 * 4 + slice] = make_int2(bk0, bj0);
            res[(128 + r0) * 4 + slice] = make_int2(bk1, bj1);
            named_bar_sync(1, 32 * kEpiWarps);
            if (slice < 2) {
                const int row = slice * 128 + r0;
                if (row < w.q_valid) {
                    int2 best = res[row * 4];
#pragma unroll
                    for (int q = 1; q < 4; q++) {
                        const int2 o = res[row * 4 + q];
                        if (o.x < best.x || (o.x == best.x && o.y < best.y)) best = o;
                    }
                    out_idx[w.out_off + row] = best.y;
                    out_d2[w.out_off + row] = best.x + norm[w.q_row + row];
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

int uavm_launch_match(uavm_ctx* ctx, uavm_pairbatch* pb)
{
    if (pb->n_items == 0) return UAVM_OK;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    // UAVM_K2_DBG != 0: timing experiment (wrong results) — the epilogue only drains TMEM and hands the buffers back, which leaves
    // the TMA + MMA pipeline alone: 0.354 ms per 49 pairs = 1 000 cycles per 256 x 128 x 128 tile (see the header)
    static const int dbg = getenv("UAVM_K2_DBG") ? atoi(getenv("UAVM_K2_DBG")) : 0;
    if (!ctx->k2_attr_set) {                              // function attributes are per device: tracked per context
        UAVM_CUDA(ctx, cudaFuncSetAttribute(k2_match_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        ctx->k2_attr_set = true;
    }
    int grid = pb->n_items < ctx->sm_count ? pb->n_items : ctx->sm_count;
    k2_match_tcgen05<<<grid, kThreads, kSmemBytes, ctx->stream>>>(pb->fs->tmap_q, pb->fs->tmap_t, pb->fs->d_ckey, pb->fs->d_norm, pb->d_items, pb->n_items,
                                                                   pb->d_train_idx, pb->d_d2, dbg);
    UAVM_CHECK_LAUNCH(ctx);
    pb->fs->pool_read_since_wait = true;                  // a later host upload into the pool has to wait for this launch
    return UAVM_OK;
}
