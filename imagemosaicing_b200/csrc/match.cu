// match.cu — K2: exact brute-force L2 1-NN of SIFT-128 descriptors as a tcgen05 distance-GEMM with a
// fused arg-min epilogue.  Replaces FlannBasedMatcher().match (M/MosaicWithoutPos.cpp:5108-5110).
//
//   |a_i - b_j|^2 = |a_i|^2 + |b_j|^2 - 2 a_i.b_j ;  the u8 x u8 -> s32 dot products come from
//   tcgen05.mma.kind::i8 (exact), accumulators live in TMEM, and the M x N distance matrix is never
//   written anywhere: each epilogue thread owns one query row and keeps a running (min key, arg-min).
//
// Work item  = 256 query rows of one pair (two M=128 accumulator row blocks) x all train tiles (N=128).
// CTA layout = 20 warps, persistent, one CTA per SM:
//   warp 0      TMA producer  (A block once per item; B tiles through a kStages-deep smem ring, each with its 128 column keys)
//   warp 1      MMA issuer    (one elected lane; 2 row blocks x 4 K-chunks of 32 per train tile)
//   warp 2      TMEM allocator (512 columns = 2 accumulator buffers x 2 row blocks x 128 columns)
//   warps 4-19  epilogue      (warp w reads TMEM lanes 32*(w%4).., row block ((w-4)/4)&1, column half (w-4)/8)
// Pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), A full/empty.
//
// What bounds it (scripts/microbench/ldtm.cu, profiles/r2_ldtm_microbench.txt): NOT the TMEM read port — tcgen05.ld.32x32b.x32
// sustains 39 / 68 / 87 words per clock per SM from 4 / 8 / 16 warps, against the 16 words per clock this kernel needs at the
// MMA rate.  The first version (8 epilogue warps, column keys fetched with 32 LDG.128 per tile and thread) ran at 2057 cycles per
// 256 x 128 tile with the tensor pipe 28 % busy: two warps per scheduler cannot hide the latency chain wait -> tcgen05.ld ->
// key loads -> min tree.  Now 16 epilogue warps (four per scheduler) each own 64 columns of a row block, and the column keys
// arrive in shared memory with the B tile (cp.async.bulk on the same mbarrier), so the epilogue reads them with broadcast LDS:
// 1429 cycles per tile, 0.51 ms per 49 pairs (was 0.70), 1.65 PFLOP/s.  Measured and rejected: handing the two row blocks of a
// tile over separately (one mbarrier pair per row block: 0.199 vs 0.181 ms on the 32k x 32k case).
//
// Key trick: column key ckey[j] = 32*|b_j|^2 + (j & 31) (K1), packed = ckey[j] - 64*dot
//          = 32*(|b_j|^2 - 2 dot) + (j & 31): ONE integer min over a 32-column chunk gives the minimum
//   distance and, on ties, the lowest column; chunks/tiles are visited in ascending j with a strict <.
#include "internal.h"
#include "ptx.cuh"

using namespace uavm::ptx;

namespace {

constexpr int kStages = 6;            // B ring depth (16 KB per stage)
constexpr int kTileN = 128;           // train rows per tile
constexpr int kBlockM = 256;          // query rows per work item
constexpr int kEpiWarps = 16;
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr uint32_t kABytes = kBlockM * 128;
constexpr uint32_t kBBytes = kTileN * 128;
constexpr uint32_t kKeyBytes = kTileN * 4;
// The producer runs at most kStages tiles ahead of the MMA and the MMA at most 2 tiles (TMEM buffers) ahead of the epilogue:
// a key ring of 16 tiles is never overwritten before the epilogue has read it.
constexpr int kKeyStages = 16;
constexpr size_t kSmemBytes = 1024 /*align slack*/ + kABytes + kStages * kBBytes + kKeyStages * kKeyBytes + 2 * kBlockM * 2 * 8 /*half results*/ + 256 /*barriers*/;

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
                 int32_t* __restrict__ out_idx, int32_t* __restrict__ out_d2)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kABytes;
    int32_t* smem_key = reinterpret_cast<int32_t*>(smem + kABytes + kStages * kBBytes);                   // [kKeyStages][128]
    int2* smem_res = reinterpret_cast<int2*>(smem + kABytes + kStages * kBBytes + kKeyStages * kKeyBytes); // [2 (item parity)][256 rows][2 halves]
    Barriers* bar = reinterpret_cast<Barriers*>(smem + kABytes + kStages * kBBytes + kKeyStages * kKeyBytes + 2 * kBlockM * 2 * 8);

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
        for (int b = 0; b < 2; b++) { mbar_init(&bar->tmem_full[b], 1); mbar_init(&bar->tmem_empty[b], kEpiWarps); }
        fence_mbar_init();
    }
    if (warp == 2) {
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
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> registers -> running arg-min =====================
        const int ew = warp - 4;
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int rb = (ew >> 2) & 1;                 // accumulator row block
        const int ch = ew >> 3;                       // column half of the tile: columns [64 ch, 64 ch + 64)
        const int row_in_block = rb * 128 + quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        uint32_t acc = 0, it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, it++) {
            const MatchItem w = items[item];
            int best_key = 0x7fffffff;
            int best_j = 0;
            for (int t = 0; t < w.n_tiles; t++, acc++) {
                const uint32_t buf = acc & 1;
                // the keys of this tile landed with its B tile (same mbarrier), which the MMA thread observed before it issued the
                // MMAs whose completion tmem_full signals.  (Waiting on full[] here as well would be wrong: the producer may
                // already have re-armed that stage twice, and a parity wait cannot tell phase k from phase k + 2.)
                mbar_wait(&bar->tmem_full[buf], (acc >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_addr + buf * 256 + rb * 128 + ch * 64;
                uint32_t v0[32], v1[32];
                tmem_ld_32x32b_x32(taddr, v0);
                tmem_ld_32x32b_x32(taddr + 32, v1);
                const int4* __restrict__ ck = reinterpret_cast<const int4*>(smem_key + (acc % kKeyStages) * kTileN + ch * 64);
                tmem_ld_wait();
                int m0 = 0x7fffffff, m1 = 0x7fffffff;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int4 c0 = ck[i], c1 = ck[8 + i];                        // broadcast LDS.128
                    const int p0 = c0.x - 64 * (int)v0[4 * i + 0];
                    const int p1 = c0.y - 64 * (int)v0[4 * i + 1];
                    const int p2 = c0.z - 64 * (int)v0[4 * i + 2];
                    const int p3 = c0.w - 64 * (int)v0[4 * i + 3];
                    m0 = __vimin3_s32(m0, p0, p1);
                    m0 = __vimin3_s32(m0, p2, p3);
                    const int q0 = c1.x - 64 * (int)v1[4 * i + 0];
                    const int q1 = c1.y - 64 * (int)v1[4 * i + 1];
                    const int q2 = c1.z - 64 * (int)v1[4 * i + 2];
                    const int q3 = c1.w - 64 * (int)v1[4 * i + 3];
                    m1 = __vimin3_s32(m1, q0, q1);
                    m1 = __vimin3_s32(m1, q2, q3);
                }
                const int k0 = m0 >> 5, k1 = m1 >> 5;
                if (k0 < best_key) { best_key = k0; best_j = t * kTileN + ch * 64 + (m0 & 31); }
                if (k1 < best_key) { best_key = k1; best_j = t * kTileN + ch * 64 + 32 + (m1 & 31); }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar->tmem_empty[buf]);
            }
            // merge the two column halves of a row: smaller key wins, equal keys keep the lower column (the halves visit
            // disjoint columns, each in ascending order with a strict <)
            int2* res = smem_res + (it & 1) * (kBlockM * 2);
            res[row_in_block * 2 + ch] = make_int2(best_key, best_j);
            named_bar_sync(1, 32 * kEpiWarps);
            if (ch == 0 && row_in_block < w.q_valid) {
                const int2 o = res[row_in_block * 2 + 1];
                if (o.x < best_key || (o.x == best_key && o.y < best_j)) { best_key = o.x; best_j = o.y; }
                out_idx[w.out_off + row_in_block] = best_j;
                out_d2[w.out_off + row_in_block] = best_key + norm[w.q_row + row_in_block];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

int uavm_launch_match(uavm_ctx* ctx, uavm_pairbatch* pb)
{
    if (pb->n_items == 0) return UAVM_OK;
    UAVM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->k2_attr_set) {                              // function attributes are per device: tracked per context
        UAVM_CUDA(ctx, cudaFuncSetAttribute(k2_match_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        ctx->k2_attr_set = true;
    }
    int grid = pb->n_items < ctx->sm_count ? pb->n_items : ctx->sm_count;
    k2_match_tcgen05<<<grid, kThreads, kSmemBytes, ctx->stream>>>(pb->fs->tmap_q, pb->fs->tmap_t, pb->fs->d_ckey,
                                                                   pb->fs->d_norm, pb->d_items, pb->n_items,
                                                                   pb->d_train_idx, pb->d_d2);
    UAVM_CHECK_LAUNCH(ctx);
    pb->fs->pool_read_since_wait = true;                  // a later host upload into the pool has to wait for this launch
    return UAVM_OK;
}
