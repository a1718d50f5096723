// 3x3 convolution (stride 1 or 2) + folded BN + ReLU on tcgen05 with PIXEL-GROUP ROWS
// (north_star (a); network.py:19-25 for the op, :184-188 for the stride-2 layers).
//
// Measured on B200 (profiles/r1_rate_probe.log): an M=128, K=16 tcgen05.mma from shared memory
// costs max(~48, N/2) cycles, so with N = Cout = 16 / 32 the tensor pipe idles 5/6 resp. 2/3 of
// the time; TMA moves <= 1 box row per cycle, so 32-byte rows cap it at ~30 B/cycle/SM.  Both
// problems go away when one shared-memory ROW holds G = 64 / Cin horizontally adjacent pixels
// (always 128 bytes, 128-byte swizzle) and one UMMA row produces the G (stride 1) or G/2
// (stride 2) output pixels of that group at once:
//     N = Gout x Cout = 64 for every layer up to 64 channels,
//     K = the input pixels the group touches along x (G+2 for stride 1, G+1 for stride 2) x Cin,
// with the 3x3 taps placed block-wise in an expanded weight matrix (zero blocks where a tap does
// not connect an input pixel of the slice to an output pixel of the group).  Each K-slice (one
// input pixel of the window) is one UMMA whose A descriptor is just a byte offset into the halo
// patch: row offset -1/0/+1 groups, sub-pixel offset inside the 128-byte row, 32-byte K-steps.
// MMA instructions per 128 output pixels drop from 9 to 4.5 (16->16), 7.5 (16->32 s2), 12 from
// 18 (32->32); stride 2 needs no strided TMA traversal (TF SAME on even sizes: pad 0 before, 1
// after, supplied by TMA's zero fill, as is the pad 1/1 of stride 1).
//
// One tile = 16 output rows x 8 groups (M = 128: TMEM lane = row*8 + group).  One 4-D TMA box
// load brings the (18 or 33) x (10 or 9) x 128 B patch; weights stay resident in shared memory.
// Epilogue (4 warps): one tcgen05.ld of all 64 columns, packed FFMA2 scale/shift, F2FP.RELU to
// 16 bit, swizzled st.shared into a staging tile and ONE TMA store per tile (coalesced, clipped
// at the image border by the tensor map).
#pragma once
#include "tc_common.cuh"

namespace ukbb {

struct ConvGroupParams {
    int tiles_x, tiles_y, n_tiles;
    const float* scale;             // [COUT]
    const float* shift;             // [COUT]
    // split ("x3") mode: hi / lo planes.  Input lo plane = slices [lo_n, ...) of the patch tensor map; the output is written with
    // direct 256-bit stores (no staging tile: the shared memory holds two patches per stage and both weight sets instead)
    int lo_n;
    uint32_t* out;                  // hi plane, [n][ho][wog groups][64 elements]
    long long out_lo;               // offset of the lo plane in 32-bit words
    int ho, wog;
};

// SPLIT: operands are hi + lo pairs of 16-bit values and every K-slice needs three products (hi.hi + lo.hi + hi.lo, tc_conv.cu).
// An N = 64 UMMA from shared memory is bound by the operand port (4 KB of A + 2 KB of B = 48 cycles for 32 cycles of tensor work), so
// the two products that share A_hi are ONE UMMA with N = 128: the weight tile of a K-slice is stored as [w_hi rows | w_lo rows] and
// A_hi . [w_hi ; w_lo]^T lands in accumulator columns [0, 64) and [64, 128) (4 + 4 KB = 64 cycles instead of 2 x 48); A_lo . w_hi^T
// adds into columns [0, 64).  The epilogue sums the two column blocks.  112 instead of 144 port cycles per K-slice.
// TROWS: output rows per tile (<= 16; the UMMA still has M = 128 = 16 rows x 8 groups, rows >= TROWS are idle): the 64 -> 64 split
// instance holds 144 KB of weights and fits two (hi, lo) patch stages only with 14-row tiles.
// F8 (x2 scheme, tc_common.cuh): the lo planes hold FP8 correction operands; per K-slice ONE kind::f8f6f4 UMMA (A_lo8 . B_lo8) replaces the
// two correction products, all of them are issued before the FP16 pass whose first instruction rescales the accumulator.
// PAIR (x2 scheme): clusters of two CTAs execute every UMMA together (tcgen05 cta_group::2, M = 256 = both CTAs' tiles): a CTA keeps
// only ITS 32 of the 64 rows of every weight tile, so an instruction reads 4 KB of A + 1 KB of B per CTA instead of 4 + 2 KB (these N = 64
// UMMAs are bound by the shared-memory operand port: 48 cycles for 32 cycles of tensor work), and the resident weights take half the
// shared memory -- the 64 -> 64 instance gets its 16-row tiles and the staging tiles back.  Protocol as in conv_halo.cuh.
template <int CC, int COUT, int STRIDE, bool SPLIT = false, bool F8 = false, bool PAIR = false>
struct ConvGroupCfg {
    static_assert(!F8 || SPLIT, "the FP8 correction scheme is a split-operand scheme");
    static_assert(!PAIR || F8, "the CTA-pair variant exists for the x2 scheme");
    static constexpr bool MERGE = SPLIT && !F8;                // x3: [w_hi ; w_lo] tiles, N = 128 UMMAs, two accumulator column blocks
    static constexpr int G = 64 / CC;                          // input pixels per 128-byte row
    static constexpr int GOUT = STRIDE == 1 ? G : G / 2;       // output pixels per UMMA row
    static constexpr int N = GOUT * COUT;
    static constexpr int J = STRIDE == 1 ? G + 2 : G + 1;      // K-slices per kernel row
    static constexpr int KS = CC / 16;
    static constexpr int TROWS = (SPLIT && CC == 64 && STRIDE == 1 && !PAIR) ? 14 : 16;
    static constexpr int PU = STRIDE == 1 ? 10 : 9;            // patch groups per row
    static constexpr int PR = STRIDE == 1 ? TROWS + 2 : 2 * TROWS + 1;   // patch rows
    static constexpr int PATCH_TX = PR * PU * 128;
    static constexpr int PATCH_BYTES = (PATCH_TX + 1023) / 1024 * 1024;
    static constexpr int A_STAGE_BYTES = (SPLIT ? 2 : 1) * PATCH_BYTES;   // hi patch | lo patch
    static constexpr int B_ROW = CC * 2;
    static constexpr int B_ROWS = PAIR ? N / 2 : N;            // weight rows of a tile held by one CTA
    static constexpr int B_TILE = (B_ROWS * B_ROW + 1023) / 1024 * 1024;
    static constexpr int NB_TILES = 3 * J;
    static constexpr int B_SET = NB_TILES * B_TILE;
    static constexpr int B_BYTES = (SPLIT ? 2 : 1) * B_SET;    // hi tiles | lo tiles
    // SPLIT: the output leaves through one (hi, lo) pair of staging tiles and two TMA stores where two patch stages leave room for them
    // (32 -> 32): a thread owns a 128-byte row and its 32-byte global stores cost one LSU wavefront per lane (32 lines per instruction);
    // the other split instances keep the direct 256-bit stores.
    static constexpr bool STAGED = SPLIT && TROWS == 16 && 2 * A_STAGE_BYTES + B_BYTES + 2 * 128 * 128 + 1024 + 256 + 2 * N * 4 <= 227 * 1024;
    static constexpr int OUT_BYTES = (SPLIT && !STAGED) ? 0 : 128 * 128;    // staging tile [128 rows][128 B]
    static constexpr int A_MAX = ((SPLIT ? 224 : 200) * 1024 - B_BYTES - 2 * OUT_BYTES) / A_STAGE_BYTES;
    static constexpr int A_STAGES = A_MAX > 4 ? 4 : A_MAX;
    static constexpr int ACC_STAGES = 2;
    static constexpr int ACC_COLS = MERGE ? 2 * N : N;         // x3: columns [0, N) = hi.hi + lo.hi, [N, 2N) = hi.lo
    static constexpr int TMEM_COLS = 2 * ACC_COLS;
    static constexpr int SMEM_BYTES = A_STAGES * A_STAGE_BYTES + B_BYTES + 2 * OUT_BYTES + 1024 + 256 + 2 * N * 4;
    static_assert(N == 64, "pixel-group kernel is built for N = Gout * Cout = 64");
    static_assert(STRIDE == 1 || G >= 2, "stride 2 needs at least two pixels per row");
    static_assert(A_STAGES >= 2, "need at least two patch stages");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

namespace tc {
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
// (a0, a1) * (b0, b1) + (c0, c1) in one FFMA2, then round both to 16 bit with ReLU (and clamp to
// the finite FP16 range in FP16 mode): two instructions per pair of outputs.
template <bool F16>
__device__ __forceinline__ uint32_t bn_relu_pack(uint32_t a0, uint32_t a1, float2 sc, float2 sh) {
    uint64_t a, b, c, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(sc.x), "f"(sc.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(sh.x), "f"(sh.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d));
    uint32_t r;
    if (F16) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// four channels of a 16-channel group (quad q): BN + ReLU, then split_pack4 (tc_common.cuh)
template <bool F16, bool F8>
__device__ __forceinline__ void bn_relu_split4(const uint32_t* v, float4 sc, float4 sh, uint32_t* oh, uint32_t* ol, int q) {
    uint64_t a, b, c, d0, d1;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(v[0]), "r"(v[1]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(sc.x), "f"(sc.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(sh.x), "f"(sh.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d0) : "l"(a), "l"(b), "l"(c));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(v[2]), "r"(v[3]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(sc.z), "f"(sc.w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(sh.z), "f"(sh.w));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d1) : "l"(a), "l"(b), "l"(c));
    float x0, x1, x2, x3;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(d0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(x3) : "l"(d1));
    split_pack4<F16, F8, true>(x0, x1, x2, x3, oh, ol, q);
}
// one 256-bit global store (STG.256): a full 32-byte sector per thread
__device__ __forceinline__ void stg256(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
                 "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
}  // namespace tc

// static round-robin tile walk without divisions in the loop
struct TileWalk {
    int tx, ty, n, dx, dy, dn, tiles_x, tiles_y;
    __device__ __forceinline__ void init(int first, int step, int tiles_x_, int tiles_y_) {
        tiles_x = tiles_x_; tiles_y = tiles_y_;
        tx = first % tiles_x; ty = (first / tiles_x) % tiles_y; n = first / (tiles_x * tiles_y);
        dx = step % tiles_x; dy = (step / tiles_x) % tiles_y; dn = step / (tiles_x * tiles_y);
    }
    __device__ __forceinline__ void next() {
        tx += dx; if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
        ty += dy; if (ty >= tiles_y) { ty -= tiles_y; ++n; }
        n += dn;
    }
};

template <int CC, int COUT, int STRIDE, bool F16, bool SPLIT = false, bool F8 = false, bool PAIR = false>
__global__ void __launch_bounds__(256, 1)
conv_group_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_out, const ConvGroupParams p) {
    using namespace tc;
    using Cfg = ConvGroupCfg<CC, COUT, STRIDE, SPLIT, F8, PAIR>;
    constexpr bool MERGE = Cfg::MERGE;
    constexpr int AST = Cfg::A_STAGES, G = Cfg::G, J = Cfg::J, KS = Cfg::KS, PU = Cfg::PU, N = Cfg::N, TROWS = Cfg::TROWS;
    griddep_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = smem_base + AST * Cfg::A_STAGE_BYTES;
    const uint32_t out_base = b_base + Cfg::B_BYTES;
    const uint32_t bar_base = out_base + 2 * Cfg::OUT_BYTES;
    // barriers: a_full[AST] a_empty[AST] tfull[2] tempty[2] wfull | tmem slot
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (AST + s); };
    auto tfull = [&](int a) { return bar_base + 8u * (2 * AST + a); };
    auto tempty = [&](int a) { return bar_base + 8u * (2 * AST + 2 + a); };
    const uint32_t wfull = bar_base + 8u * (2 * AST + 4);
    const uint32_t tmem_slot = bar_base + 8u * (2 * AST + 5);
    float* s_scale = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));   // [N] expanded (column -> channel)
    float* s_shift = s_scale + N;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); tma_prefetch_desc(&map_out); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < AST; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), PAIR ? 8 : 4); }   // PAIR: the leader's barrier, both epilogues
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) { tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
    }
    if (warp == 3) {
        for (int c = lane; c < N; c += 32) { s_scale[c] = p.scale[c % COUT]; s_shift[c] = p.shift[c % COUT]; }
    }
    tc_fence_before();
    __syncthreads();
    // PAIR: both CTAs' barriers are initialised before any remote arrive / commit, and the collective cta_group::2 allocation has written
    // the tensor-memory address into BOTH CTAs' slots before either reads it
    if (PAIR) cluster_sync();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp != 0) griddep_wait();                       // the producer waits after it has issued the weight loads
    // tile walk: round-robin over CTAs, or (PAIR) pairs of tiles (2 k, 2 k + 1) round-robin over clusters; with an odd tile count the
    // peer of the last pair walks a tile beyond the tensor (its loads are harmless, its epilogue stores nothing)
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    const int n_slices = p.n_tiles / (p.tiles_x * p.tiles_y);
    const int walk_first = PAIR ? 2 * ((int)blockIdx.x >> 1) + (int)crank : (int)blockIdx.x;
    const int walk_step = PAIR ? ((int)gridDim.x >> 1) * 2 : (int)gridDim.x;
    const int n_units = PAIR ? (p.n_tiles + 1) >> 1 : p.n_tiles, unit0 = PAIR ? (int)blockIdx.x >> 1 : (int)blockIdx.x, unit_step = PAIR ? (int)gridDim.x >> 1 : (int)gridDim.x;
    const int my_tiles = unit0 < n_units ? (n_units - unit0 + unit_step - 1) / unit_step : 0;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            if (PAIR) {                                  // this CTA's 32 rows of every tile (box of the weight map: 32 rows); all on the leader's barrier
                if (crank == 0) mbar_arrive_expect_tx(wfull, 2 * Cfg::NB_TILES * N * Cfg::B_ROW);
                for (int t = 0; t < 2 * Cfg::NB_TILES; ++t)
                    tma_load_2d_pair(b_base + t * Cfg::B_TILE, &map_b, mapa_rank(wfull, 0), 0, t * N + (int)crank * (N / 2));
            } else {
            mbar_arrive_expect_tx(wfull, (SPLIT ? 2 : 1) * Cfg::NB_TILES * N * Cfg::B_ROW);
            for (int t = 0; t < (SPLIT ? 2 : 1) * Cfg::NB_TILES; ++t) {
                // global: hi tiles, then lo tiles.  Shared (SPLIT): tile t = [hi 64 rows | lo 64 rows], one N = 128 B operand
                const int tt = t % Cfg::NB_TILES, pl = t / Cfg::NB_TILES;
                tma_load_2d(b_base + (MERGE ? tt * 2 * Cfg::B_TILE + pl * Cfg::B_TILE : t * Cfg::B_TILE), &map_b, wfull, 0, t * N);
            }
            }
            griddep_wait();
            TileWalk w;
            w.init(walk_first, walk_step, p.tiles_x, p.tiles_y);
            int as = 0;
            uint32_t aph = 0;
            for (int i = 0; i < my_tiles; ++i) {
                mbar_wait(a_empty(as), aph ^ 1);
                const uint32_t dst = smem_base + as * Cfg::A_STAGE_BYTES;
                const int cx = STRIDE == 1 ? w.tx * 8 - 1 : w.tx * 8, cy = STRIDE == 1 ? w.ty * TROWS - 1 : w.ty * 2 * TROWS;
                if (PAIR) {                              // the patches of both CTAs complete on the leader's barrier
                    if (crank == 0) mbar_arrive_expect_tx(a_full(as), 4 * Cfg::PATCH_TX);
                    const uint32_t lbar = mapa_rank(a_full(as), 0);
                    tma_load_4d_pair(dst, &map_a, lbar, 0, cx, cy, w.n);
                    tma_load_4d_pair(dst + Cfg::PATCH_BYTES, &map_a, lbar, 0, cx, cy, p.lo_n + w.n);
                    if (++as == AST) { as = 0; aph ^= 1; }
                    w.next();
                    continue;
                }
                mbar_arrive_expect_tx(a_full(as), (SPLIT ? 2 : 1) * Cfg::PATCH_TX);
                tma_load_4d(dst, &map_a, a_full(as), 0, cx, cy, w.n);
                if (SPLIT) tma_load_4d(dst + Cfg::PATCH_BYTES, &map_a, a_full(as), 0, cx, cy, p.lo_n + w.n);
                if (++as == AST) { as = 0; aph ^= 1; }
                w.next();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
        const bool leader = elect_one();
        const uint32_t idesc = F16 ? make_idesc_f16(128, N) : make_idesc_bf16(128, N);
        const uint32_t idesc2 = F16 ? make_idesc_f16(128, 2 * N) : make_idesc_bf16(128, 2 * N);       // x3: A_hi . [w_hi ; w_lo]^T
        const uint32_t idesc8 = make_idesc_e4m3(128, N);                                              // x2: FP8 correction pass
        constexpr uint32_t blayout = Cfg::B_ROW == 128 ? 2u : Cfg::B_ROW == 64 ? 4u : 6u;
        constexpr uint32_t a_hi = (uint32_t)(((STRIDE == 1 ? PU : 2 * PU) * 128) >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t b_hi = (uint32_t)((8 * Cfg::B_ROW) >> 4) | (1u << 14) | (blayout << 29);
        const uint32_t b_lo = ((b_base & 0x3FFFF) >> 4) | (1u << 16);
        int as = 0, acc = 0;
        uint32_t aph = 0, acc_ph = 0;
        if (PAIR) {
            // ---- CTA pair: the leader issues M = 256 UMMAs over both CTAs' tiles; the peer's MMA warp has nothing to do
            const uint32_t idesc_p = make_idesc_f16(256, N), idesc8_p = make_idesc_e4m3(256, N);
            if (crank == 0) {
                mbar_wait(wfull, 0);
                tc_fence_after();
                for (int i = 0; i < my_tiles; ++i) {
                    mbar_wait(tempty(acc), acc_ph ^ 1);
                    mbar_wait(a_full(as), aph);
                    tc_fence_after();
                    const uint32_t d = tmem_base + acc * Cfg::ACC_COLS;
                    const uint32_t a_lo = (((smem_base + as * Cfg::A_STAGE_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
                    if (leader) {
#pragma unroll
                        for (int pass = 0; pass < 2; ++pass)
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int j = 0; j < J; ++j) {
                                const int jj = STRIDE == 1 ? j - 1 : j;
                                const int ro = jj < 0 ? -1 : jj / G;
                                const int sub = jj - ro * G;
                                const int arow = STRIDE == 1 ? ky * PU + 1 + ro : ky * PU + ro;
#pragma unroll
                                for (int ks = 0; ks < KS; ++ks) {
                                    const uint32_t ao = (uint32_t)((arow * 128 + sub * CC * 2 + ks * 32) >> 4);
                                    const uint32_t bo = (uint32_t)(((ky * J + j) * Cfg::B_TILE + ks * 32) >> 4);
                                    if (pass == 0) umma_pair_lohi<1>(d, a_lo + (Cfg::PATCH_BYTES >> 4) + ao, a_hi, b_lo + (Cfg::B_SET >> 4) + bo, b_hi, idesc8_p, (ky | j | ks) != 0 ? 1u : 0u);
                                    else if ((ky | j | ks) == 0) umma_pair_lohi<2>(d, a_lo + ao, a_hi, b_lo + bo, b_hi, idesc_p, 1u);
                                    else umma_pair_lohi<0>(d, a_lo + ao, a_hi, b_lo + bo, b_hi, idesc_p, 1u);
                                }
                            }
                        umma_commit_pair(a_empty(as), (uint16_t)3);
                        umma_commit_pair(tfull(acc), (uint16_t)3);
                    }
                    __syncwarp();
                    if (++as == AST) { as = 0; aph ^= 1; }
                    if (++acc == 2) { acc = 0; acc_ph ^= 1; }
                }
            }
        } else {
        mbar_wait(wfull, 0);
        tc_fence_after();
        for (int i = 0; i < my_tiles; ++i) {
            mbar_wait(tempty(acc), acc_ph ^ 1);
            mbar_wait(a_full(as), aph);
            tc_fence_after();
            const uint32_t d = tmem_base + acc * Cfg::ACC_COLS;
            const uint32_t a_lo = (((smem_base + as * Cfg::A_STAGE_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
            if (leader) {
#pragma unroll
                for (int pass = F8 ? 0 : 1; pass < 2; ++pass)            // x2: pass 0 = FP8 corrections, pass 1 = FP16 main term
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        // input pixel of this K-slice relative to the first pixel of the group: j - 1 (stride 1), j (stride 2)
                        const int jj = STRIDE == 1 ? j - 1 : j;
                        const int ro = jj < 0 ? -1 : jj / G;
                        const int sub = jj - ro * G;
                        const int arow = STRIDE == 1 ? ky * PU + 1 + ro : ky * PU + ro;
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint32_t ao = (uint32_t)((arow * 128 + sub * CC * 2 + ks * 32) >> 4);
                            const uint32_t bo = (uint32_t)(((ky * J + j) * (MERGE ? 2 : 1) * Cfg::B_TILE + ks * 32) >> 4);
                            if (F8) {
                                if (pass == 0) umma_f8_lohi(d, a_lo + (Cfg::PATCH_BYTES >> 4) + ao, a_hi, b_lo + (Cfg::B_SET >> 4) + bo, b_hi, idesc8, (ky | j | ks) != 0 ? 1u : 0u);
                                else if ((ky | j | ks) == 0) umma_f16_lohi_rescale(d, a_lo + ao, a_hi, b_lo + bo, b_hi, idesc);
                                else umma_bf16_lohi(d, a_lo + ao, a_hi, b_lo + bo, b_hi, idesc, 1u);
                            } else if (SPLIT) {
                                umma_bf16_lohi(d, a_lo + ao, a_hi, b_lo + bo, b_hi, idesc2, (ky | j | ks) != 0 ? 1u : 0u);           // hi . [hi ; lo]
                                umma_bf16_lohi(d, a_lo + (Cfg::PATCH_BYTES >> 4) + ao, a_hi, b_lo + bo, b_hi, idesc, 1u);            // lo . hi
                            } else {
                                umma_bf16_lohi(d, a_lo + ao, a_hi, b_lo + bo, b_hi, idesc, (ky | j | ks) != 0 ? 1u : 0u);
                            }
                        }
                    }
                umma_commit(a_empty(as));
                umma_commit(tfull(acc));
            }
            __syncwarp();
            if (++as == AST) { as = 0; aph ^= 1; }
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp - 4;
        const int r = q * 32 + lane;                         // TMEM lane = tile row * 8 + group
        const bool issuer = threadIdx.x == 128;
        TileWalk w;
        w.init(walk_first, walk_step, p.tiles_x, p.tiles_y);
        int acc = 0;
        uint32_t acc_ph = 0;
        // folded BN of the COUT channels in registers (SPLIT): two broadcast LDS.128 per four outputs were a tenth of the epilogue
        float4 scv[SPLIT ? COUT / 4 : 1], shv[SPLIT ? COUT / 4 : 1];
        if (SPLIT) {
#pragma unroll
            for (int c = 0; c < COUT / 4; ++c) { scv[c] = *reinterpret_cast<const float4*>(s_scale + 4 * c); shv[c] = *reinterpret_cast<const float4*>(s_shift + 4 * c); }
        }
        for (int i = 0; i < my_tiles; ++i) {
            mbar_wait(tfull(acc), acc_ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC_COLS;
            if (SPLIT) {
                // acc = columns [0, 64) + columns [64, 128) (x3); hi / lo pieces, 128 contiguous bytes each per thread (this row's group):
                // four 256-bit stores per plane, or (STAGED) swizzled staging tiles and two TMA stores
                const int row = r >> 3, y = w.ty * TROWS + row, gx = w.tx * 8 + (r & 7);
                const bool live = row < TROWS && y < p.ho && gx < p.wog && w.n < n_slices;
                uint32_t* dst = p.out + (((size_t)w.n * p.ho + y) * p.wog + gx) * 32;
                const uint32_t srow = out_base + r * 128;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    uint32_t v[16], v2[16], oh[8], ol[8];
                    tmem_ld16(taddr + 16 * c8, v);
                    if (MERGE) tmem_ld16(taddr + 64 + 16 * c8, v2);
                    tmem_ld_wait();
                    if (c8 == 3) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {                             // accumulator is in registers: release it to the MMA warp
                            if (PAIR) mbar_arrive_cluster(mapa_rank(tempty(acc), 0));
                            else mbar_arrive(tempty(acc));
                        }
                    }
                    if (MERGE) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(v2[c]));
                    }
#pragma unroll
                    for (int c = 0; c < 16; c += 4) bn_relu_split4<F16, F8>(v + c, scv[((16 * c8 + c) % COUT) / 4], shv[((16 * c8 + c) % COUT) / 4], oh, ol, c / 4);
                    if (Cfg::STAGED) {
                        if (c8 == 0) {                               // the stores of the previous tile have read the staging tiles
                            if (issuer) bulk_wait_read<0>();
                            named_bar_sync(1, 128);
                        }
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint32_t d = srow + ((uint32_t)((2 * c8 + k) ^ (r & 7)) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(oh[4 * k]), "r"(oh[4 * k + 1]), "r"(oh[4 * k + 2]),
                                         "r"(oh[4 * k + 3]) : "memory");
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d + Cfg::OUT_BYTES), "r"(ol[4 * k]), "r"(ol[4 * k + 1]),
                                         "r"(ol[4 * k + 2]), "r"(ol[4 * k + 3]) : "memory");
                        }
                    } else if (live) {
                        stg256(dst + 8 * c8, oh);
                        stg256(dst + p.out_lo + 8 * c8, ol);
                    }
                }
                if (Cfg::STAGED) {
                    fence_proxy_async();
                    named_bar_sync(1, 128);
                    if (issuer && w.n < n_slices) {
                        tma_store_4d(&map_out, out_base, 0, w.tx * 8, w.ty * TROWS, w.n);
                        tma_store_4d(&map_out, out_base + Cfg::OUT_BYTES, 0, w.tx * 8, w.ty * TROWS, w.n + p.lo_n);
                        bulk_commit();
                    }
                }
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
                w.next();
                continue;
            }
            uint32_t v[64];
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + 32, v + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));         // accumulator is in registers: release it to the MMA warp
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
            uint32_t o[32];
#pragma unroll
            for (int c = 0; c < 64; c += 4) {
                const float4 sc = *reinterpret_cast<const float4*>(s_scale + c);
                const float4 sh = *reinterpret_cast<const float4*>(s_shift + c);
                o[c / 2] = bn_relu_pack<F16>(v[c], v[c + 1], make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
                o[c / 2 + 1] = bn_relu_pack<F16>(v[c + 2], v[c + 3], make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
            }
            // staging set (i & 1) was read by the store of tile i - 2, which the issuer waited for before the last barrier
            const uint32_t row = out_base + (i & 1) * Cfg::OUT_BYTES + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t dst = row + ((uint32_t)(j ^ (r & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[4 * j]), "r"(o[4 * j + 1]), "r"(o[4 * j + 2]),
                             "r"(o[4 * j + 3])
                             : "memory");
            }
            fence_proxy_async();
            if (issuer) bulk_wait_read<0>();                 // store of tile i - 1 has left its staging set
            named_bar_sync(1, 128);
            if (issuer) {
                tma_store_4d(&map_out, out_base + (i & 1) * Cfg::OUT_BYTES, 0, w.tx * 8, w.ty * TROWS, w.n);
                bulk_commit();
            }
            w.next();
        }
        if (issuer) bulk_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync();                            // the peer may still arrive on this CTA's barriers / read its shared memory
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace ukbb
