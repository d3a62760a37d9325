// tcgen05 / TMA / TMEM sweep GEMM for sm_100a with the fused epilogue.
//
//   out[M x N] = epi( sum_src sum_(pa,pb) A_src,pa[M x K] * B_src,pb[N x K]^T )
//
// A: state planes (fp16, K-major, row = chain), B: weight planes (fp16 of s*W, K-major:
// W[V x ldH] for the down-sweep, Wt[H x ldV] for the up-sweep).  The hi/lo
// operand planes are extra MMA passes into the same fp32 TMEM accumulator.
//
// Persistent CTAs (one per SM), warp-specialised:
//   warps 0 .. EPW-1  epilogue          (tcgen05.ld 32x32b -> bias/activation/Philox draw -> fp16 planes)
//   warp EPW          TMA producer      (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp EPW+1        MMA issuer        (tcgen05.mma kind::f16, M=128, N=BN, K=16 per instruction)
//   warp EPW+2        TMEM allocator
// Two TMEM accumulator stages (2 x BN columns) let the MMAs of tile j+1 run
// under the (sampler-bound) epilogue of tile j.
#pragma once
#include <cuda.h>

#include <unordered_map>

#include "epilogue.cuh"

namespace scdbm {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled costs a few microseconds on the host; the latency-bound training loops (BASELINE config C1:
// ~30 descriptors per minibatch) re-use the same weight planes and minibatch buffers, so descriptors are cached by
// everything they depend on (address, extents, row pitch, box).  A descriptor is a pure function of that key.
struct TcMapKey {
    const void* ptr;
    int64_t rows, cols, ld;
    int32_t box_rows, box_cols;
    bool operator==(const TcMapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && box_cols == o.box_cols;
    }
};
struct TcMapKeyHash {
    size_t operator()(const TcMapKey& k) const {
        uint64_t h = (uint64_t)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
        h ^= (uint64_t)k.rows * 0xC2B2AE3D27D4EB4Full + ((uint64_t)k.cols << 17) + ((uint64_t)k.ld << 41) + ((uint64_t)k.box_rows << 7) + (uint64_t)k.box_cols;
        return (size_t)(h ^ (h >> 29));
    }
};
struct TcContext {
    PFN_tmapEncodeTiled encode = nullptr;
    bool attr_set[24] = {};
    bool nb_cg2 = true;     // NB class with long contractions: two SMs per 256-row tile (context option "two_sm")
    std::unordered_map<TcMapKey, CUtensorMap, TcMapKeyHash> maps;
};

struct TcOperandPlanes {
    MatView a;                       // state (1 or 2 planes)
    const plane_t *b0, *b1;    // weight planes [N x ldb]
    int64_t ldb, N, K;
};

struct TcGradArgs {
    MatView v, h, vm, hm;
    float pos_scale, neg_scale;   // out = pos_scale * v'h + neg_scale * vm'hm  (neg_scale carries the minus sign)
    float* out;                   // fp32 [V x H], row-major
    int64_t V, H;
    float* partial;               // split-K workspace [splits][V][H] (nullptr: splits == 1)
    int32_t splits;
    float* da = nullptr;          // non-null (small batches): trailing CTAs of the same launch also produce the bias gradients
    float* db = nullptr;          //   da[V] = pos_scale colsum(v) + neg_scale colsum(vm),  db[H] likewise from h, hm
};

constexpr int TC_BM = 128, TC_BK = 64;
constexpr int TC_A_TILE = TC_BM * TC_BK * 2;  // 16 KB
constexpr int TC_MAX_STAGES = 8;
// two-pass NB epilogue, per epilogue warp: a queue of 8-byte entries {accumulator, Philox word} with the entry's position
// inside the warp's 32 x CW block packed into bits the draw does not use, a small queue of 16-byte survivor
// states {u, P, q, k | position} for draws that outlive the unrolled first steps, and a 32 x CW fp16 staging
// tile of the lo count plane
#ifndef SCDBM_NB_STEPS
#define SCDBM_NB_STEPS 3
#endif
#ifndef SCDBM_NB_EPW      // epilogue warps of the wide NB sampling kernel; its tile is 8 * SCDBM_NB_EPW columns
#define SCDBM_NB_EPW 28
#endif
constexpr int TC_QCAP = 176, TC_SCAP = 64, TC_NB_STEPS = SCDBM_NB_STEPS;
#ifndef SCDBM_NB_DYN      // 1: the epilogue warps of a row quadrant claim (tile, slice) items dynamically
#define SCDBM_NB_DYN 1
#endif
#ifndef SCDBM_NB_CW       // columns per epilogue warp and tile of the wide NB kernel (16 or 32)
#define SCDBM_NB_CW 32
#endif
#ifndef SCDBM_NB_ASTAGES  // its TMEM accumulator stages (stages x tile width <= 512 columns)
#define SCDBM_NB_ASTAGES 2
#endif
constexpr int TC_NB_EPW = SCDBM_NB_EPW, TC_NB_BN = (SCDBM_NB_CW / 4) * TC_NB_EPW;
static_assert(SCDBM_NB_ASTAGES * TC_NB_BN <= 512, "NB kernel: accumulator stages exceed TMEM");
// Long contractions (K >= 512: the 1 024-unit hidden layer of configs C4 / C5): the 28 epilogue warps' queues leave room for ONE
// 73 KB operand stage, so every k-block's load -> MMA round trip (~2 us) is exposed: 16 of them per tile, 3x the
// epilogue's time.  There the NB class runs 20 epilogue warps on 160-column tiles: two operand stages fit, and three
// accumulator stages (480 TMEM columns).  (16 warps / 128 columns / three operand stages measured 4 % slower.)
#ifndef SCDBM_NBK_EPW
#define SCDBM_NBK_EPW 20
#endif
constexpr int TC_NBK_EPW = SCDBM_NBK_EPW, TC_NBK_BN = 8 * TC_NBK_EPW, TC_NBK_MIN_KB = 8;
// ... and the slice's per-gene constants (nbc: 16 B, threshold: 4 B per column): with the shared-memory carve-out at its
// maximum the L1 is too small to keep the per-gene tables, and every gather from them was an L2 round trip
__host__ __device__ constexpr int tc_epi_smem(int cw) { return TC_QCAP * 8 + TC_SCAP * 16 + 32 * cw * 2 + cw * 20; }
constexpr int TC_SMEM_BUDGET = 200 * 1024;

struct alignas(64) TcMaps {
    CUtensorMap a[2][2];
    CUtensorMap b[2][2];
};

struct TcParams {
    int32_t nsrc;
    int32_t kblocks[2];
    int32_t nA[2];
    uint32_t combos[2];   // bit (pa*2+pb)
    int64_t M, N;
    int32_t m_tiles, n_tiles;
    int32_t stages, stage_bytes;
    int32_t a_tiles;      // A-plane tiles reserved per stage (1 or 2)
    int32_t split_hi[2];  // 1: COUNT source -- the (rarely used) hi plane is contracted in a second k-pass through the
                          //    same single A slot instead of occupying a second A tile in every stage
    int32_t pair;         // 1: a work tile is 256 rows = two M=128 MMA tiles sharing every B (weight) tile load:
                          //    a third less L2 -> shared-memory operand traffic (the up-sweep is L2-bandwidth-bound)
    const uint32_t* a_hiflag[2];   // per source: COUNT operand's per-row-tile "hi plane in use" flags (nullptr: always)
    int32_t ksplit;       // > 1: split-K for latency-bound shapes with few tiles -- work item = (tile, split); each split contracts
                          //      a slice of the k-blocks and stores raw fp32 partial sums (EPI_RAW_F32) into its own slab
    int32_t sched;        // 1: a CTA owns whole m-tiles and walks all n-tiles (balanced over genes, A re-used
                          //    back-to-back); 0: flat round-robin over (m, n) tiles, n fastest (few m-tiles)
    int32_t tposed;       // two-SM kernel, deterministic class: the MMA's roles are swapped (weights as the "A" operand), the
                          //    accumulator holds out^T (lane = output unit) and the epilogue's global accesses are coalesced
    int32_t mfast;        // flat order with the ROW tile fastest: the concurrently running CTAs share one weight tile and
                          //    re-stream the (smaller) state operand instead of all weight planes
    EpiParams epi;
};

inline void tc_context_init(TcContext& t) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
        t.encode = (PFN_tmapEncodeTiled)fn;
}

// ------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Every barrier wait: try_wait with a suspend-time hint (the hardware may park the warp; each mbarrier event of the CTA
// wakes it, so a waiting warp still polls -- measured harmless: the polls do not take issue slots from working warps),
// and a bounded number of wake-ups -- a protocol bug becomes a launch error, never a hang.
#ifndef SCDBM_PARK_NS
#define SCDBM_PARK_NS 20000
#endif
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"((uint32_t)SCDBM_PARK_NS)
            : "memory");
        if (done) break;
        if (++spins > (1u << 22)) __trap();   // never hang the GPU: a protocol bug becomes a launch error
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- cluster / cta_group::2 forms (two-SM kernels)
constexpr uint32_t C2_PEER_MASK = 0xFEFFFFFFu; // clears the CTA-rank bit of a shared::cluster address -> the leader's copy

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar_leader, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar_leader), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_commit_cg2(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
        "r"(cta)
        : "memory");
}

// K-major, 128B-swizzled operand tile: 8-row groups of 1024 B (SBO), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                    // version
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Mean-field extras of the deterministic class: the hoisted addend and the previous value are read once, straight from
// HBM, by a loop with one 16-byte load per array in flight per thread.  Each lane asks L2 for its row's part of the
// warp's block (CW floats) while the warp still waits for the accumulator.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void epi_prefetch_extras(const EpiParams& e, int64_t row, int64_t M, int64_t col, int64_t N, int cw) {
    if (row >= M || col >= N) return;
    if (e.addend) {
        const float* a = e.addend + row * e.ld_add + col;
        for (int b = 0; b < cw; b += 32) prefetch_l2(a + b);
    }
    if (e.prev.f) {
        const float* o = e.prev.f + row * e.prev.ld + col;
        for (int b = 0; b < cw; b += 32) prefetch_l2(o + b);
    }
}

// pass A of the NB epilogue for one element: `word >= T` (the draw is not decided by the zero shortcut) -> append the
// 8-byte entry {x, y} to the warp's queue at shared address qbase + 8 * slot.  One compare feeds the ballot and
// predicates the store; no branch.
__device__ __forceinline__ void nb_enqueue(uint32_t word, uint32_t T, uint32_t x, uint32_t y, uint32_t lanes_below, uint32_t qbase, int& qn) {
    uint32_t cnt;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 m, t;\n\t"
        "setp.ge.u32 p, %2, %3;\n\t"
        "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
        "and.b32 t, m, %6;\n\t"
        "popc.b32 t, t;\n\t"
        "add.u32 t, t, %1;\n\t"
        "shl.b32 t, t, 3;\n\t"
        "add.u32 t, t, %7;\n\t"
        "@p st.shared.v2.u32 [t], {%4, %5};\n\t"
        "popc.b32 %0, m;\n\t}"
        : "=r"(cnt)
        : "r"(qn), "r"(word), "r"(T), "r"(x), "r"(y), "r"(lanes_below), "r"(qbase)
        : "memory");
    qn += (int)cnt;
}

// tile schedule shared by the three warp roles: it = 0, 1, ... -> flat tile id m * n_tiles + n, or -1 at the end
// (bid, nb: this CTA's -- or cluster's -- index and their number)
__device__ __forceinline__ int tc_tile(const TcParams& p, int it, int bid, int nb) {
    if (p.sched) {
        const int m = bid + (it / p.n_tiles) * nb;
        return m < p.m_tiles ? m * p.n_tiles + it % p.n_tiles : -1;
    }
    const int t = bid + it * nb;
    if (t >= p.m_tiles * p.n_tiles * max(p.ksplit, 1)) return -1;     // split-K: t is a work item, see tc_split_range
    if (p.mfast) return (t % p.m_tiles) * p.n_tiles + t / p.m_tiles;   // row tile fastest: the CTAs share a weight tile
    return t;
}

// split-K: work item id -> (tile, split) and the split's range of the tile's flattened k-block sequence
__device__ __forceinline__ void tc_split_range(const TcParams& p, int work, int tkb, int& tile, int& lo, int& hi) {
    if (p.ksplit <= 1) { tile = work; lo = 0; hi = tkb; return; }
    tile = work / p.ksplit;
    const int sp = work % p.ksplit;
    lo = (int)((int64_t)tkb * sp / p.ksplit);
    hi = (int)((int64_t)tkb * (sp + 1) / p.ksplit);
}

// hi count plane needed for work tile m_blk of source s?  (pair: either 128-row half flagged)
__device__ __forceinline__ bool tc_use_hi(const TcParams& p, int s, int m_blk) {
    if (!p.a_hiflag[s] || p.nA[s] != 2) return p.nA[s] == 2;
    if (!p.pair) return (p.a_hiflag[s][m_blk] & 1u) != 0u;
    const int64_t last = (p.M + TC_BM - 1) / TC_BM - 1;
    return (p.a_hiflag[s][2 * m_blk] & 1u) != 0u || (2 * m_blk + 1 <= last && (p.a_hiflag[s][2 * m_blk + 1] & 1u) != 0u);
}

// Long contractions are accumulated in chunks.  tcgen05.mma adds into its fp32 TMEM accumulator with truncation, so a
// same-sign sum over thousands of k-steps drifts by ~0.5 ulp per step (measured: 4e-5 relative at K = 20 000, which a
// sigmoid in its tail turns into 3e-4 -- outside gate G1).  A tile with more than TC_PROMO_KB k-blocks (K > 4 608) is therefore
// contracted in chunks of <= TC_CHUNK_KB k-blocks, each into a fresh accumulator stage; the epilogue warps add the
// chunks in registers (round-to-nearest fp32) and put the total back into TMEM before the usual epilogue runs.
constexpr int TC_PROMO_KB = 72, TC_CHUNK_KB = 32;
__device__ __forceinline__ int tc_tile_kblocks(const TcParams& p, int m_blk) {
    int t = 0;
    for (int s = 0; s < p.nsrc; ++s) t += p.kblocks[s] * ((p.split_hi[s] && tc_use_hi(p, s, m_blk)) ? 2 : 1);
    return t;
}
__host__ __device__ __forceinline__ int tc_chunks(int tkb) { return tkb > TC_PROMO_KB ? (tkb + TC_CHUNK_KB - 1) / TC_CHUNK_KB : 1; }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ kernel
// EK: epilogue class -- 0 generic (all modes, accurate math, explicit uniforms), 1 Philox-fed NB draw
// (fast math), 2 Philox-fed Bernoulli draw (fast math).  Separate instantiations keep each kernel's
// code small enough for the instruction cache.
// EPW: number of epilogue warps (4 per column slice).  The sampling epilogue is latency-bound, so the NB
// kernel runs 24 epilogue warps on 192-column tiles; the others 16 warps on 64/128-column tiles.
// CG2 (NB class with long contractions only): two CTAs of a cluster work on a 256-row tile with one cta_group::2 MMA per
// k-step, each staging its own 128 rows and half of the weight tile (protocol: see gemm_tc2.cuh); launched with a
// cluster dimension of 2.
template <int BN, int EK, int EPW, bool CG2 = false>
__global__ void __launch_bounds__(32 * (4 + EPW), 1) k_gemm_tc(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p) {
    static_assert(!CG2 || EK == 1, "the two-SM form of this kernel serves the NB class; see k_gemm_cg2 for the others");
    constexpr int NS = EPW / 4;          // column slices per tile
    constexpr int CW = BN / NS;          // columns per epilogue warp per tile (16 or 32)
    constexpr int B_TILE = (CG2 ? BN / 2 : BN) * TC_BK * 2;     // CG2: this CTA's half of the weight tile
    constexpr int ASTAGES = (EK == 1 && BN == TC_NB_BN) ? SCDBM_NB_ASTAGES : (EK == 1 && BN == TC_NBK_BN) ? 3 : 2;    // TMEM accumulator stages
    constexpr uint32_t TMEM_COLS = (EK == 1) ? ((ASTAGES * BN <= 256) ? 256 : 512) : ((2 * ASTAGES * BN <= 256) ? 256 : 512);
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((CG2 ? 2 * TC_BM : TC_BM) >> 4) << 24);   // c = F32, a = b = F16, K-major
    static_assert(CW == 16 || CW == 32, "epilogue slice must be 16 or 32 columns");

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES], bar_empty[TC_MAX_STAGES], bar_tfull[4], bar_tempty[4];
    __shared__ uint32_t tmem_holder;
    __shared__ int dyn_item[4];     // EK == 1: next (tile, slice) item of each row quadrant
    if constexpr (EK == 0 || EK == 3) {
        if (p.epi.skip) {      // speculatively launched mean-field iteration after convergence (whole grid)
            pdl_wait();        // the flag is written by the preceding launch
            if (*p.epi.skip) return;
        }
    }

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // Warp roles: epilogue warps are 0 .. EPW-1; the service warps sit ABOVE them (svc = warp - EPW: 0 TMA producer,
    // 1 MMA issuer, 2 TMEM allocator).  The warp scheduler prefers the highest warp index among the eligible warps of
    // a sub-partition, so the few instructions that keep the tensor pipe fed never queue behind the sampling epilogue.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int svc = warp - EPW;
    const uint32_t rank = CG2 ? cluster_ctarank() : 0u;                       // CTA of the pair; 0 = leader
    const int bid = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, nbid = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (svc == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(smem_u32(&bar_full[i]), 1);
            mbar_init(smem_u32(&bar_empty[i]), 1);
        }
        for (int i = 0; i < ASTAGES; ++i) {
            mbar_init(smem_u32(&bar_tfull[i]), 1);
            mbar_init(smem_u32(&bar_tempty[i]), CG2 ? 2 * EPW : EPW);
        }
        for (int i = 0; i < 4; ++i) dyn_item[i] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    } else if (svc == 2) {
        if constexpr (CG2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if constexpr (CG2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;
    pdl_trigger();      // this CTA holds its TMEM columns: the next launch may start setting up
    pdl_wait();         // operands and bias / flag tables come from the preceding launches

    if (svc == 0) {
        if (lane == 0) {
            // ===================== TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0, work; (work = tc_tile(p, it, bid, nbid)) >= 0; ++it) {
                const int rows = p.pair ? 2 * TC_BM : TC_BM;
                const uint32_t a_bytes = (uint32_t)rows * TC_BK * 2;
                int tile, klo, khi;
                tc_split_range(p, work, tc_tile_kblocks(p, (p.ksplit > 1 ? work / p.ksplit : work) / p.n_tiles), tile, klo, khi);
                const int m0 = CG2 ? (tile / p.n_tiles) * 2 * TC_BM + (int)rank * TC_BM : (tile / p.n_tiles) * rows;
                const int n0 = (tile % p.n_tiles) * BN + (CG2 ? (int)rank * (BN / 2) : 0);
                int f = 0;      // position in the tile's flattened k-block sequence
                for (int s = 0; s < p.nsrc; ++s) {
                    // counts >= 2048 are rare: the hi plane of a COUNT operand is contracted only where this row tile has one
                    const bool hi = tc_use_hi(p, s, tile / p.n_tiles);
                    const int npass = p.split_hi[s] ? (hi ? 2 : 1) : 1;
                    const int nA = p.split_hi[s] ? 1 : (hi ? p.nA[s] : 1);
                    const uint32_t tx = (uint32_t)nA * a_bytes + 2 * B_TILE;
                    for (int pass = 0; pass < npass; ++pass)
                        for (int kb = 0; kb < p.kblocks[s]; ++kb, ++f) {
                            if (f < klo || f >= khi) continue;
                            mbar_wait_parked(smem_u32(&bar_empty[stage]), phase ^ 1u);
                            const uint32_t full = smem_u32(&bar_full[stage]);
                            const uint32_t sa = smem_base + (uint32_t)stage * (uint32_t)p.stage_bytes;
                            if constexpr (CG2) {     // both CTAs' loads complete on the leader's barrier
                                const uint32_t full_leader = full & C2_PEER_MASK;
                                if (rank == 0) mbar_expect_tx(full, 2u * tx);
                                for (int pa = 0; pa < nA; ++pa) tma_load_2d_cg2(sa + pa * a_bytes, &maps.a[s][pa + pass], full_leader, kb * TC_BK, m0);
                                for (int pb = 0; pb < 2; ++pb)
                                    tma_load_2d_cg2(sa + (uint32_t)p.a_tiles * a_bytes + pb * B_TILE, &maps.b[s][pb], full_leader, kb * TC_BK, n0);
                            } else {
                            mbar_expect_tx(full, tx);
                            for (int pa = 0; pa < nA; ++pa) tma_load_2d(sa + pa * a_bytes, &maps.a[s][pa + pass], full, kb * TC_BK, m0);
                            for (int pb = 0; pb < 2; ++pb)
                                tma_load_2d(sa + (uint32_t)p.a_tiles * a_bytes + pb * B_TILE, &maps.b[s][pb], full, kb * TC_BK, n0);
                            }
                            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        }
                }
            }
        }
    } else if (svc == 1) {
        if (lane == 0 && rank == 0) {
            // ===================== MMA issuer (CG2: the leader CTA issues for the pair)
            int stage = 0, astage = 0;
            uint32_t phase = 0, aphase = 0;
            for (int it = 0, work; (work = tc_tile(p, it, bid, nbid)) >= 0; ++it) {
                int tile, klo, khi;
                tc_split_range(p, work, tc_tile_kblocks(p, (p.ksplit > 1 ? work / p.ksplit : work) / p.n_tiles), tile, klo, khi);
                const int m_blk = tile / p.n_tiles;
                const int halves = p.pair ? 2 : 1;
                const uint32_t a_bytes = (uint32_t)halves * TC_A_TILE;
                // chunked accumulation (see tc_chunks): every chunk gets a fresh accumulator stage
                const int tkb = khi - klo;
                const int nch = tc_chunks(tkb), per = (tkb + nch - 1) / nch;
                int f = 0;
                int done = 0, in_chunk = 0;
                uint32_t tmem_d = 0, accumulate = 0;
                for (int s = 0; s < p.nsrc; ++s) {
                    const bool hi = tc_use_hi(p, s, m_blk);
                    const uint32_t combos = (p.split_hi[s] || !hi) ? (p.combos[s] & 0x3u) : p.combos[s];
                    const int nkb = p.kblocks[s] * ((p.split_hi[s] && hi) ? 2 : 1);    // second pass: the hi count plane in the same A slot
                    for (int kb = 0; kb < nkb; ++kb, ++f) {
                        if (f < klo || f >= khi) continue;
                        if (in_chunk == 0) {
                            mbar_wait_parked(smem_u32(&bar_tempty[astage]), aphase ^ 1u);
                            tc_fence_after();
                            tmem_d = tmem_base + (uint32_t)(astage * halves) * BN;
                            accumulate = 0;
                        }
                        mbar_wait_parked(smem_u32(&bar_full[stage]), phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + (uint32_t)stage * (uint32_t)p.stage_bytes;
                        const uint32_t sb = sa + (uint32_t)p.a_tiles * a_bytes;
                        for (int h = 0; h < halves; ++h) {        // both row halves re-use the B tile in shared memory
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                if (!((combos >> c) & 1u)) continue;
                                const uint32_t a0 = sa + (c >> 1) * a_bytes + h * TC_A_TILE, b0 = sb + (c & 1) * B_TILE;
#pragma unroll
                                for (int k = 0; k < TC_BK / 16; ++k)
                                    if constexpr (CG2)
                                        tc_mma_f16_cg2(tmem_d + h * BN, umma_desc_k_sw128(a0 + k * 32), umma_desc_k_sw128(b0 + k * 32), IDESC,
                                                       (accumulate | (uint32_t)(c != __ffs(combos) - 1) | (uint32_t)k) ? 1u : 0u);
                                    else
                                    tc_mma_f16(tmem_d + h * BN, umma_desc_k_sw128(a0 + k * 32), umma_desc_k_sw128(b0 + k * 32), IDESC,
                                               (accumulate | (uint32_t)(c != __ffs(combos) - 1) | (uint32_t)k) ? 1u : 0u);
                            }
                        }
                        accumulate = 1;
                        if constexpr (CG2) tc_commit_cg2(smem_u32(&bar_empty[stage]), 3);
                        else tc_commit(smem_u32(&bar_empty[stage]));   // smem slot is free once these MMAs retire
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        ++done;
                        if (++in_chunk == per || done == tkb) {
                            if constexpr (CG2) tc_commit_cg2(smem_u32(&bar_tfull[astage]), 3);
                            else tc_commit(smem_u32(&bar_tfull[astage]));      // chunk (or whole accumulator) ready for the epilogue warps
                            if (++astage == ASTAGES) { astage = 0; aphase ^= 1u; }
                            in_chunk = 0;
                        }
                    }
                }
            }
        }
    } else if (svc < 0) {
        // ===================== epilogue warps
        // TMEM lanes tie a warp to its row quadrant (warp index mod 4); the column slice rotates per tile so that no
        // warp keeps the expensive genes for the whole kernel
        const int q = warp & 3;
        int slice = warp >> 2;
        int astage = 0;
        uint32_t aphase = 0;
        float delta_local = 0.0f;
        // two-pass NB epilogue state (EK == 1): per-warp queues in dynamic shared memory behind the GEMM stages
        uint8_t* wsm = smem_raw + (smem_base - smem_u32(smem_raw)) + (uint32_t)p.stages * (uint32_t)p.stage_bytes +
                       (uint32_t)warp * (uint32_t)tc_epi_smem(CW);
        uint2* Q = reinterpret_cast<uint2*>(wsm);                                   // {accumulator | row bit 4, Philox word | column, row bits 0-3}
        uint4* S = reinterpret_cast<uint4*>(Q + TC_QCAP);                           // survivors {u, P, q, k << 16 | position}
        uint16_t* s_lo = reinterpret_cast<uint16_t*>(S + TC_SCAP);                  // [32][CW] fp16 staging tile, 16 B aligned
        float4* sN = reinterpret_cast<float4*>(s_lo + 32 * CW);                     // the slice's {a, r, 1e-6/r, r-1}
        uint32_t* sT = reinterpret_cast<uint32_t*>(sN + CW);                        // the slice's pass-A thresholds
        int32_t flag_tile = -1;
        bool dirty = true;
        uint32_t cur_m0 = 0;      // (the NB class addresses rows with 32 bits: M < 2^31)
        uint32_t cur_c0 = 0;                                                        // global column of the warp's slice
        const uint32_t lt = (1u << lane) - 1u;
        const float acc_scale = p.epi.acc_scale ? __ldg(p.epi.acc_scale) : 1.0f;   // 1/s of the fp16 weight planes
        int qn = 0, sn = 0;
        const uint32_t qbase = smem_u32(Q);
        // Entry position = index into the warp's [32][CW] staging tile (row * CW + column, 10 bits).  The tile is stored
        // swizzled -- 4-byte word w of row r sits at word w ^ ((r >> 1) & (CW/2 - 1)) -- because the entries of one pass-B
        // batch mostly share a column: their 2-byte result stores would otherwise hit two banks (16-way conflict).
        auto swz = [&](int idx) { return idx ^ ((((idx / CW) >> 1) & (CW / 2 - 1)) << 1); };
        // A finished draw: lo plane into the staging tile; counts >= 2048 (rare) patch the hi plane in global memory
        auto put_big = [&](float k, int idx) {
            const EpiParams& e = p.epi;
            const float hi = floorf(k * (1.0f / COUNT_BASE)) * COUNT_BASE;
            const int64_t r2 = (int64_t)cur_m0 + q * 32 + idx / CW, c2 = (int64_t)cur_c0 + idx % CW;
            if (r2 < p.M && c2 < p.N) {
                e.out.p1[r2 * e.out.ld + c2] = plane_encode(hi);
                if (e.out.hiflag) e.out.hiflag[r2 / TC_BM] = 3u;
            }
            s_lo[swz(idx)] = __half_as_ushort(__float2half_rn(k - hi));
        };
        auto put = [&](float k, int idx) {
            if (k >= COUNT_BASE) put_big(k, idx);
            else s_lo[swz(idx)] = __half_as_ushort(__float2half_rn(k));
        };
        // second stage: `cnt` (<= 32) survivor states from S[base...], one per lane; the chop-down search runs to the end
        auto stage2 = [&](int base, int cnt) {
            if (lane < cnt) {
                const uint4 st = S[base + lane];
                const int idx = (int)(st.w & 0x3FFu);
                float u = __uint_as_float(st.x), P = __uint_as_float(st.y);
                const float qq = __uint_as_float(st.z);
                const float qr1 = qq * sN[idx % CW].w;
                float k = (float)(st.w >> 16);
                float rk = fast_rcp(k + 1.0f);
                while (u >= P && P > PMIN_F) {
                    u -= P;
                    k += 1.0f;
                    const float ratio = fmaf(qr1, rk, qq);
                    rk = fast_rcp(k + 1.0f);
                    P *= ratio;
                }
                put(k, idx);
            }
        };
        // first stage over queue entries [0, n): NB mean, P0, then TC_NB_STEPS speculative steps of the search (no
        // divergence: most draws end here); the rest are compacted into S and finished 32 at a time by stage2
        auto stage1 = [&](int n) {
            const EpiParams& e = p.epi;
            const float mu_lim = e.mu_inv_max - 1e-6f;
            __syncwarp();
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                bool surv = false;
                uint4 st = make_uint4(0u, 0u, 0u, 0u);
                if (i < n) {
                    const uint2 en = Q[i];
                    const int idx = (int)((en.y & 0x1FFu) | ((en.x & 1u) << 9));
                    const uint32_t col = cur_c0 + (uint32_t)(idx % CW);
                    const float4 c = sN[idx % CW];                                   // {a, r, 1e-6/r, r-1}
                    float u = u23(en.y);
                    const float x = fmaf(e.wscale, __uint_as_float(en.x) * acc_scale, c.x);
                    const float E = fast_ex2(x * 1.4426950408889634f);
                    float om = 1.0f - E;
                    // x -> 0-: 1 - e^x cancels; -expm1(x) = -x (1 + x/2 + ... + x^5/720), |err| < 2e-9 on (-1/8, 0)
                    if (x > -0.125f)
                        om = -x * fmaf(x, fmaf(x, fmaf(x, fmaf(x, fmaf(x, 1.0f / 720.0f, 1.0f / 120.0f), 1.0f / 24.0f), 1.0f / 6.0f), 0.5f), 1.0f);
                    if (c.y * E > mu_lim * om) {                                      // mean above mu_inv_max: gamma-Poisson, out of line
                        const RngAddr g{e.k0, e.k1, cur_m0 + (uint32_t)(q * 32 + idx / CW) + e.row_offset, col, e.tick, e.stream & 0xFFu};
                        put(nb_draw_slow(x, c.y, u, e.pshift, e.mu_inv_max, g), idx);
                    } else {
                        const float pp = om * fast_rcp(fmaf(c.z, om, 1.0f)) - e.pshift;
                        const float qq = 1.0f - pp, qr1 = qq * c.w;
                        // the first TC_NB_STEPS steps of the chop-down search, computed speculatively (no predicated
                        // updates): `go` = "the search is still running"; the count is the number of steps taken
                        // (written as its fp16 word: 0, 1, 2, 3), the survivor state is the last (u, P)
                        float P = fast_ex2(c.y * fast_lg2(pp));
                        bool go = (P > PMIN_F) && (u >= P);
                        uint32_t k16 = 0u;
#pragma unroll
                        for (int sidx = 0; sidx < TC_NB_STEPS; ++sidx) {
                            constexpr uint32_t F16[4] = {0x3C00u, 0x4000u, 0x4200u, 0x4400u};   // 1, 2, 3, 4
                            k16 = go ? F16[sidx] : k16;
                            u -= P;
                            P *= fmaf(qr1, 1.0f / (float)(sidx + 1), qq);
                            go = go && (P > PMIN_F) && (u >= P);
                        }
                        s_lo[swz(idx)] = (uint16_t)k16;                              // a survivor's slot is rewritten by stage2
                        surv = go;
                        st = make_uint4(__float_as_uint(u), __float_as_uint(P), __float_as_uint(qq), ((uint32_t)TC_NB_STEPS << 16) | (uint32_t)idx);
                    }
                }
                const uint32_t smask = __ballot_sync(0xffffffffu, surv);
                if (surv) S[sn + __popc(smask & lt)] = st;
                sn += __popc(smask);
                if (sn >= 32) {
                    __syncwarp();
                    sn -= 32;
                    stage2(sn, 32);
                    __syncwarp();
                }
            }
        };
        // mid-tile: whole batches only, the remainder moves to the front of the queue
        auto drain_full = [&]() {
            const int rem = qn & 31, full = qn - rem;
            stage1(full);
            __syncwarp();
            uint2 t = make_uint2(0u, 0u);
            if (lane < rem) t = Q[full + lane];
            __syncwarp();
            if (lane < rem) Q[lane] = t;
            qn = rem;
        };
        auto drain_all = [&]() {
            stage1(qn);
            qn = 0;
            __syncwarp();
            if (sn > 0) { stage2(0, sn); sn = 0; }
            __syncwarp();
        };
        for (int it0 = 0;; ++it0) {
            int it = it0, dyn_slice = 0;
            if constexpr (EK == 1 && SCDBM_NB_DYN) {
                // NB draw: the cost of a 32-row x CW-gene block follows its genes (zero fraction, search length), so a static
                // warp -> slice map leaves warps waiting for the slowest one of their tile (20 % of the warp samples in capture
                // prof_r02c sat in the accumulator wait).  The NS warps of a row quadrant instead claim (tile, slice) items in
                // order from a shared counter.  A warp claiming an item of tile `it` implies an item of tile it-1 was finished
                // (NS items, NS-1 other warps), hence every earlier phase of this TMEM stage's barrier has completed.
                int item = 0;
                if (lane == 0) item = atomicAdd(&dyn_item[q], 1);
                item = __shfl_sync(0xffffffffu, item, 0);
                it = item / NS;
                dyn_slice = item % NS;
                astage = it % ASTAGES;
                aphase = (uint32_t)(it / ASTAGES) & 1u;
            }
            const int work = tc_tile(p, it, bid, nbid);
            if (work < 0) break;
            int tile, klo, khi;
            tc_split_range(p, work, EK == 1 ? 0 : tc_tile_kblocks(p, (p.ksplit > 1 ? work / p.ksplit : work) / p.n_tiles), tile, klo, khi);
            const int64_t row_shift = p.ksplit > 1 ? (int64_t)(work % p.ksplit) * p.M : 0;    // split-K: slab of raw partial sums
            const int halves = (EK != 1 && p.pair) ? 2 : 1;
            const int64_t m0 = CG2 ? (int64_t)(tile / p.n_tiles) * (2 * TC_BM) + (int64_t)rank * TC_BM : (int64_t)(tile / p.n_tiles) * (TC_BM * halves);
            const int64_t n0 = (int64_t)(tile % p.n_tiles) * BN;
            slice = (EK == 1 && SCDBM_NB_DYN) ? dyn_slice : ((warp >> 2) + it) % NS;
            if constexpr (EK == 1) {
                // lo count plane: staged in the warp's shared-memory tile (zeroed here, patched by pass B, written
                // out coalesced at the end of the tile).  hi plane (counts >= 2048, rare): coalesced zero-fill now where
                // the row tile is marked (full 32 B sectors: 4 lanes x 16 B per row, 8 rows per store), scattered patches from pass B.
                cur_m0 = (uint32_t)m0; cur_c0 = (uint32_t)n0 + (uint32_t)(slice * CW);
                const MatView& o = p.epi.out;
                // per-gene constants of this slice (both tables are padded past the last tile): loads issued now, parked in
                // shared memory after the zero-fill
                float4 cn = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                uint32_t ct = 0u;
                if (lane < CW) { cn = __ldg(p.epi.nbc + cur_c0 + lane); ct = __ldg(p.epi.nbT + cur_c0 + lane); }
                // the hi plane of a row tile is all zero unless its flag says otherwise (bit 1: stale entries of an earlier
                // sweep); read once per row tile -- a flag raised by this launch marks entries written into a clean plane
                if (tile / p.n_tiles != flag_tile) {
                    flag_tile = tile / p.n_tiles;
                    dirty = !o.hiflag || (*reinterpret_cast<const volatile uint32_t*>(o.hiflag + m0 / TC_BM) & 2u) != 0u;
                }
#pragma unroll
                for (int i = 0; i < (32 * CW * 2) / (32 * 16); ++i) {
                    const int idx = i * 32 + lane;                   // 16-byte chunk index within the region
                    const int rr = idx / (CW / 8), ch = idx % (CW / 8);
                    const int64_t r2 = m0 + q * 32 + rr, c2 = n0 + slice * CW + ch * 8;
                    reinterpret_cast<uint4*>(s_lo)[idx] = make_uint4(0u, 0u, 0u, 0u);
                    if (dirty && r2 < p.M && c2 < o.ld) *reinterpret_cast<uint4*>(o.p1 + r2 * o.ld + c2) = make_uint4(0u, 0u, 0u, 0u);
                }
                if (lane < CW) { sN[lane] = cn; sT[lane] = ct; }
                __syncwarp();
            }
            float bias_l = 0.0f, r_l = 1.0f;       // this lane's column of the warp's slice: summed biases, inverse dispersion
            if constexpr (EK != 1) {
                const int64_t cl = n0 + slice * CW + lane;
                if (lane < CW && cl < p.N) {
                    if (p.epi.bias0) bias_l = __ldg(p.epi.bias0 + cl);
                    if (p.epi.bias1) bias_l += __ldg(p.epi.bias1 + cl);
                    if (p.epi.r && (p.epi.mode == EPI_NBMEAN || p.epi.mode == EPI_NB_DRAW)) r_l = __ldg(p.epi.r + cl);
                }
            }
            if constexpr (EK == 3) {
                for (int h = 0; h < halves; ++h) epi_prefetch_extras(p.epi, m0 + h * TC_BM + q * 32 + lane, p.M, n0 + slice * CW, p.N, CW);
            }
            if constexpr (EK != 1) {
                // chunked accumulation: add all but the last chunk in registers, then fold the last one in and put
                // the total back into the last chunk's TMEM stage
                const int nch = tc_chunks(khi - klo);
                if (nch > 1) {
                    float sum[2][CW];      // [row half of a pair-mode work tile][column of the warp's slice]
#pragma unroll
                    for (int i = 0; i < CW; ++i) sum[0][i] = sum[1][i] = 0.0f;
                    for (int c = 0; c < nch; ++c) {
                        mbar_wait_parked(smem_u32(&bar_tfull[astage]), aphase);
                        tc_fence_after();
                        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(astage * halves * BN + slice * CW);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (h < halves) {
#pragma unroll
                                for (int g = 0; g < CW / 16; ++g) {
                                    uint32_t r[16];
                                    tmem_ld16(ta + h * BN + g * 16, r);
                                    tmem_wait_ld();
#pragma unroll
                                    for (int i = 0; i < 16; ++i) sum[h][g * 16 + i] += __uint_as_float(r[i]);
                                }
                            }
                        }
                        if (c + 1 < nch) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[astage]));
                            if (++astage == ASTAGES) { astage = 0; aphase ^= 1u; }
                        } else {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if (h < halves) {
#pragma unroll
                                    for (int g = 0; g < CW / 16; ++g) {
                                        uint32_t r[16];
#pragma unroll
                                        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(sum[h][g * 16 + i]);
                                        tmem_st16(ta + h * BN + g * 16, r);
                                    }
                                }
                            }
                            tmem_wait_st();
                        }
                    }
                } else {
                    mbar_wait_parked(smem_u32(&bar_tfull[astage]), aphase);
                    tc_fence_after();
                }
            } else {
                mbar_wait_parked(smem_u32(&bar_tfull[astage]), aphase);
                tc_fence_after();
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(astage * halves * BN + slice * CW);
            const int64_t row = m0 + q * 32 + lane;
            float rowsum = 0.0f;
            if constexpr (EK == 1) {
                // ---- two-pass NB draw.  Pass A (all elements): Philox word vs the per-gene integer threshold T of
                // a lower bound L of the zero probability (u < L  <=>  word < T, exactly the fp32 comparison of the
                // 23-bit uniform; T is padded to whole tiles with 0xFFFFFFFF): below => the draw is 0 (the staging
                // tile is pre-zeroed).  The other ~20 % are appended (ballot + popc) to the warp's queue and drawn
                // densely, one per lane, in pass B.  The entry's position rides in bits the draw ignores: the 9 low
                // bits of the word (the uniform uses the top 23) and the last mantissa bit of the accumulator.
                const EpiParams& e = p.epi;
                // position of (lane, column c) in the staging tile: lane * CW + c; bit 9 travels in the accumulator's last bit
#pragma unroll 1      // (2-way unrolling: more spills at the 64-register cap, measured 6 % slower)
                for (int g = 0; g < CW / 4; ++g) {
                    uint32_t r[4];
                    tmem_ld4(taddr + g * 4, r);
                    const uint32_t col0 = cur_c0 + (uint32_t)(g * 4);
                    const uint4 w = philox_group_rk(e, (uint32_t)row, col0 >> 2);
                    const uint4 T = *reinterpret_cast<const uint4*>(sT + g * 4);
                    tmem_wait_ld();
                    const uint32_t wj[4] = {w.x, w.y, w.z, w.w}, Tj[4] = {T.x, T.y, T.z, T.w};
                    const uint32_t pos0 = (uint32_t)(lane * CW + g * 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        nb_enqueue(wj[j], Tj[j], (r[j] & ~1u) | (pos0 >> 9), (wj[j] & ~0x1FFu) | ((pos0 + (uint32_t)j) & 0x1FFu), lt, qbase, qn);
                    if (qn > TC_QCAP - 128) drain_full();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG2) mbar_arrive_cluster(smem_u32(&bar_tempty[astage]), 0);
                    else mbar_arrive(smem_u32(&bar_tempty[astage]));
                }
                if (++astage == ASTAGES) { astage = 0; aphase ^= 1u; }
                drain_all();
                {   // coalesced write-out of the staged lo plane (undoing the word swizzle of each row)
                    const MatView& o = p.epi.out;
#pragma unroll
                    for (int i = 0; i < (32 * CW * 2) / (32 * 16); ++i) {
                        const int idx = i * 32 + lane;
                        const int rr = idx / (CW / 8), ch = idx % (CW / 8);
                        const int sw = (rr >> 1) & (CW / 2 - 1);                       // word XOR of this row
                        uint4 v = reinterpret_cast<const uint4*>(s_lo)[rr * (CW / 8) + (ch ^ (sw >> 2))];
                        if (sw & 1) v = make_uint4(v.y, v.x, v.w, v.z);
                        if (sw & 2) v = make_uint4(v.z, v.w, v.x, v.y);
                        const int64_t r2 = m0 + q * 32 + rr, c2 = n0 + slice * CW + ch * 8;
                        if (r2 < p.M && c2 < o.ld) *reinterpret_cast<uint4*>(o.p0 + r2 * o.ld + c2) = v;
                    }
                    __syncwarp();
                }
            } else {
            // one 4-column group at a time straight out of TMEM: a single copy of the epilogue body
#pragma unroll 1
            for (int hg = 0; hg < halves * (CW / 4); ++hg) {
                const int h = hg / (CW / 4), g = hg % (CW / 4);     // row half (pair mode), 4-column group
                const int64_t rowh = row + h * TC_BM;
                uint32_t r[4];
                tmem_ld4(taddr + h * BN + g * 4, r);   // warp-collective: outside any row/column predicate
                float bias4[4], r4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    bias4[j] = __shfl_sync(0xffffffffu, bias_l, g * 4 + j);
                    r4[j] = __shfl_sync(0xffffffffu, r_l, g * 4 + j);
                }
                tmem_wait_ld();
                const int64_t col0 = n0 + slice * CW + g * 4;
                const int nvalid = (int)max((int64_t)0, min((int64_t)4, p.N - col0));
                if (rowh < p.M && nvalid > 0) {
                    const float acc[4] = {__uint_as_float(r[0]) * acc_scale, __uint_as_float(r[1]) * acc_scale, __uint_as_float(r[2]) * acc_scale,
                                          __uint_as_float(r[3]) * acc_scale};
                    if constexpr (EK == 2) epi_bern4(p.epi, rowh, col0, acc, nvalid, bias4);
                    else if constexpr (EK == 3) epi_det4(p.epi, rowh + row_shift, col0, acc, nvalid, delta_local, bias4, r4);
                    else rowsum += epi_apply4(p.epi, rowh + row_shift, col0, acc, nvalid, delta_local, bias4, r4);
                }
                if constexpr (EK == 0) {
                    if (g == CW / 4 - 1) {                           // end of this row's slice: flush the AIS partial row sum
                        if (p.epi.mode == EPI_AIS_RATIO && rowh < p.M) atomicAdd(&p.epi.logw[rowh], (double)rowsum);
                        rowsum = 0.0f;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[astage]));   // TMEM stage can be overwritten
            if (++astage == ASTAGES) { astage = 0; aphase ^= 1u; }
            }
        }
        if (p.epi.prev.p0 && p.epi.delta_bits) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) delta_local = fmaxf(delta_local, __shfl_xor_sync(0xffffffffu, delta_local, off));
            if (lane == 0 && delta_local > 0.0f) atomicMax(p.epi.delta_bits, __float_as_uint(delta_local));
        }
    }
    tc_fence_before();
    if constexpr (CG2) cluster_sync_all(); else __syncthreads();
    if (svc == 2) {
        tc_fence_after();
        if constexpr (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------- host
inline bool tc_make_map_box(TcContext& t, CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols) {
    const TcMapKey key{ptr, rows, cols, ld, box_rows, box_cols};
    auto it = t.maps.find(key);
    if (it != t.maps.end()) {
        *map = it->second;
        return true;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    if (t.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (t.maps.size() > 8192) t.maps.clear();
    t.maps.emplace(key, *map);
    return true;
}
inline bool tc_make_map(TcContext& t, CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    return tc_make_map_box(t, map, ptr, rows, cols, ld, box_rows, TC_BK);
}

// the tcgen05 kernel tiles any sweep whose operands exist as fp16 planes
inline bool tc_supported(const GemmArgs& g, int nsrc) {
    if (nsrc < 1 || nsrc > 2 || g.M < 1 || g.N < 1) return false;
    for (int s = 0; s < nsrc; ++s)
        if (g.src[s].K < 1) return false;
    return true;
}

template <int BN, int EK, int EPW, bool CG2 = false>
inline cudaError_t tc_launch_bn(TcContext& t, const TcMaps& maps, TcParams& p, int num_sms, cudaStream_t stream) {
    const int slot = CG2 ? 11 : (EK == 1 && BN == TC_NBK_BN) ? 10 : (BN == 64 ? 0 : BN == 128 ? 4 : 8) + EK;   // wide NB kernel: slot 9
    constexpr int B_TILE = (CG2 ? BN / 2 : BN) * TC_BK * 2;
    const int units = CG2 ? num_sms / 2 : num_sms;      // CTAs, or clusters of two
    const int qbytes = (EK == 1) ? EPW * tc_epi_smem(BN / (EPW / 4)) : 0;
    p.a_tiles = std::max(p.split_hi[0] ? 1 : p.nA[0], p.nsrc > 1 ? (p.split_hi[1] ? 1 : p.nA[1]) : 1);
    p.stage_bytes = p.a_tiles * (p.pair ? 2 : 1) * TC_A_TILE + 2 * B_TILE;
    // (the NB class takes all 227 KB: its queues leave little room for operand stages)
    p.stages = std::max(1, std::min(TC_MAX_STAGES, ((EK == 1 ? 225 * 1024 : TC_SMEM_BUDGET + 20 * 1024) - qbytes) / p.stage_bytes));
    // no more stages than k-blocks of a tile: a small shared-memory request keeps latency-bound launches cheap
    p.stages = std::max(1, std::min(p.stages, p.kblocks[0] * (p.split_hi[0] ? 2 : 1) + (p.nsrc > 1 ? p.kblocks[1] * (p.split_hi[1] ? 2 : 1) : 0)));
    const int smem = p.stages * p.stage_bytes + 1024 + qbytes;
    if (!t.attr_set[slot]) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<BN, EK, EPW, CG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return e;
        t.attr_set[slot] = true;
    }
    // EK == 1 (NB draw): a CTA owns whole row tiles and walks all column tiles -- every CTA sees every gene, which balances
    // the sampling epilogue.  Other classes: flat order with the column tile fastest, so the CTAs that share a row tile
    // (A operand) stream it at the same time and the second reader hits L2 (owning it sequentially re-read A from
    // HBM: in pair mode 148 CTAs x 1 MB of A exceed the 126 MB L2).
    p.sched = (EK == 1 && p.m_tiles >= units && p.ksplit <= 1) ? 1 : 0;
    // Flat order, which operand is re-streamed from L2 / HBM?  Column tile fastest: the ~#SM concurrent tiles span all
    // weight tiles of one or two row tiles, so the weight planes are streamed once per row-tile group (AIS / PCD Gibbs at
    // C4: 16 384 x 20 000 x 1 024 read 3.3 GB of the 82 MB weight planes from HBM, ncu capture prof_r02_aisnb).  Row tile
    // fastest re-streams the state operand instead: chosen when that one is the smaller.
    p.mfast = (!p.sched && p.ksplit <= 1 && p.m_tiles > 1 && p.n_tiles > 1 && p.M * (int64_t)p.a_tiles < p.N * 2) ? 1 : 0;
    const int grid = p.sched ? std::min(p.m_tiles, units) : std::min(p.m_tiles * p.n_tiles * std::max(p.ksplit, 1), units);
    if constexpr (CG2) {      // clusters of two CTAs (plain stream order: the cluster launch does not take the PDL attribute)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * grid);
        cfg.blockDim = dim3(32 * (4 + EPW));
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, k_gemm_tc<BN, EK, EPW, CG2>, maps, p);
    }
    return launch_pdl(k_gemm_tc<BN, EK, EPW>, dim3(grid), dim3(32 * (4 + EPW)), (size_t)smem, stream, maps, p);
}

// epilogue class: the fast sampling kernels need Philox-fed draws into the canonical plane layout
inline int tc_epilogue_class(const GemmArgs& g, const TcOperandPlanes* ops, int nsrc) {
    const EpiParams& e = g.epi;
    int ek = 0;
    if (!e.uniforms && !e.addend && !e.prev.p0) {
        if (e.mode == EPI_NB_DRAW && e.factor == 1.0f && e.nbc && e.nbT && !e.bias1 && e.out.p1 && !e.out.f && e.out.kind == MAT_COUNT &&
            nsrc == 1 && ops[0].a.kind == MAT_BINARY && e.wscale >= 0.0f && e.wscale <= 1.0f) ek = 1;
        if (e.mode == EPI_SIGM_BERN && e.bias0 && !e.out.p1 && !e.out.f) ek = 2;
    }
    // deterministic class: inputs / potentials / NB means / raw accumulators, mean-field extras allowed
    if (ek == 0 && !e.uniforms && (e.mode == EPI_INPUT || e.mode == EPI_SIGM || e.mode == EPI_NBMEAN || e.mode == EPI_RAW_F32)) ek = 3;
    return ek;
}

// split-K plan for latency-bound deterministic sweeps: few output tiles, long contraction, no mean-field extras.
// Returns the number of splits (1: none).
inline int tc_plan_split(const GemmArgs& g, const TcOperandPlanes* ops, int nsrc, int num_sms) {
    const EpiParams& e = g.epi;
    const int ek = tc_epilogue_class(g, ops, nsrc);
    // deterministic class, or the Philox-fed Bernoulli draw (the finishing launch draws: same counters, same samples up to
    // the rounding of the partial sums)
    if ((ek != 3 && ek != 2) || e.mode == EPI_RAW_F32 || e.addend || e.prev.p0 || e.skip) return 1;
    const int64_t tiles = ((g.M + TC_BM - 1) / TC_BM) * ((g.N + 127) / 128);
    if (tiles * 4 > num_sms) return 1;
    int kb = 0;
    for (int s = 0; s < nsrc; ++s) kb += (int)((ops[s].K + TC_BK - 1) / TC_BK);
    if (ek == 2 && kb < 8) return 1;
    const int S = (int)std::min<int64_t>({16, kb / 2, num_sms / tiles});
    return S >= 2 ? S : 1;
}

inline cudaError_t launch_gemm_tc(TcContext& t, const TcOperandPlanes* ops, int nsrc, const GemmArgs& g, int num_sms,
                                  cudaStream_t stream, int ksplit = 1) {
    if (!t.encode) return cudaErrorNotSupported;
    const int ek = tc_epilogue_class(g, ops, nsrc);
    // latency-bound launches (a handful of tiles): 64-column tiles halve each epilogue warp's serial work and double the CTAs
    const bool few_tiles = ek != 1 && ((g.M + TC_BM - 1) / TC_BM) * ((g.N + 127) / 128) * std::max(ksplit, 1) * 4 <= num_sms;
    const bool long_k = ek == 1 && g.N >= TC_NBK_BN && (ops[0].K + TC_BK - 1) / TC_BK >= TC_NBK_MIN_KB;
    // ... and two SMs per 256-row tile when there is at least one wave of such tiles
    static const bool nb_cg2_on = [] { const char* v = getenv("SCDBM_NB_CG2"); return !(v && v[0] == '0'); }();
    const bool nb_cg2 = long_k && nb_cg2_on && t.nb_cg2 && ((g.M + 2 * TC_BM - 1) / (2 * TC_BM)) * ((g.N + TC_NBK_BN - 1) / TC_NBK_BN) >= num_sms / 2;
    const int BN = (g.N <= 64 || few_tiles) ? 64 : long_k ? TC_NBK_BN : (ek == 1 && g.N >= TC_NB_BN) ? TC_NB_BN : 128;
    TcMaps maps;
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.nsrc = nsrc;
    p.M = g.M;
    p.N = g.N;
    // pair mode: 256-row work tiles (two MMA row tiles per weight-tile load) when there is enough work to keep
    // every SM busy with pairs; the NB sampling kernel keeps single tiles (its two 224-column stages fill TMEM)
    p.pair = (ek != 1 && g.M >= (int64_t)2 * TC_BM * num_sms) ? 1 : 0;
    const int rows_per_tile = p.pair ? 2 * TC_BM : TC_BM;
    p.m_tiles = (int)((g.M + rows_per_tile - 1) / rows_per_tile);
    if (nb_cg2) p.m_tiles = (int)((g.M + 2 * TC_BM - 1) / (2 * TC_BM));     // cluster tiles; each CTA's TMA box stays 128 rows
    p.n_tiles = (int)((g.N + BN - 1) / BN);
    p.epi = g.epi;
    p.ksplit = ksplit;
    for (int s = 0; s < nsrc; ++s) {
        const MatView& a = ops[s].a;
        p.kblocks[s] = (int)((ops[s].K + TC_BK - 1) / TC_BK);
        p.a_hiflag[s] = (a.kind == MAT_COUNT) ? a.hiflag : nullptr;
        p.nA[s] = a.p1 ? 2 : 1;
        // plane products: binary x (hi,lo); count (lo,hi) x (hi,lo); real (hi,lo) x (hi,lo) minus lo*lo
        p.combos[s] = a.p1 ? (a.kind == MAT_COUNT ? 0xFu : 0x7u) : 0x3u;
        p.split_hi[s] = (a.p1 && a.kind == MAT_COUNT) ? 1 : 0;
        bool ok = tc_make_map(t, &maps.a[s][0], a.p0, g.M, ops[s].K, a.ld, rows_per_tile);
        ok = ok && tc_make_map(t, &maps.a[s][1], a.p1 ? a.p1 : a.p0, g.M, ops[s].K, a.ld, rows_per_tile);
        ok = ok && tc_make_map(t, &maps.b[s][0], ops[s].b0, ops[s].N, ops[s].K, ops[s].ldb, nb_cg2 ? BN / 2 : BN);
        ok = ok && tc_make_map(t, &maps.b[s][1], ops[s].b1, ops[s].N, ops[s].K, ops[s].ldb, nb_cg2 ? BN / 2 : BN);
        if (!ok) return cudaErrorInvalidValue;
    }
    if (nsrc == 1) { maps.a[1][0] = maps.a[0][0]; maps.a[1][1] = maps.a[0][1]; maps.b[1][0] = maps.b[0][0]; maps.b[1][1] = maps.b[0][1]; }
    if (BN == 64) {
        if (ek == 1) return tc_launch_bn<64, 1, 16>(t, maps, p, num_sms, stream);
        if (ek == 2) return tc_launch_bn<64, 2, 16>(t, maps, p, num_sms, stream);
        if (ek == 3) return tc_launch_bn<64, 3, 16>(t, maps, p, num_sms, stream);
        return tc_launch_bn<64, 0, 16>(t, maps, p, num_sms, stream);
    }
    if (BN == TC_NBK_BN && ek == 1 && nb_cg2) return tc_launch_bn<TC_NBK_BN, 1, TC_NBK_EPW, true>(t, maps, p, num_sms, stream);
    if (BN == TC_NBK_BN && ek == 1) return tc_launch_bn<TC_NBK_BN, 1, TC_NBK_EPW>(t, maps, p, num_sms, stream);
    if (BN == TC_NB_BN && ek == 1) return tc_launch_bn<TC_NB_BN, 1, TC_NB_EPW>(t, maps, p, num_sms, stream);
    if (ek == 1) return tc_launch_bn<128, 1, 16>(t, maps, p, num_sms, stream);
    if (ek == 2) return tc_launch_bn<128, 2, 16>(t, maps, p, num_sms, stream);
    if (ek == 3) return tc_launch_bn<128, 3, 16>(t, maps, p, num_sms, stream);
    return tc_launch_bn<128, 0, 16>(t, maps, p, num_sms, stream);
}

// =====================================================================================
// Gradient GEMM (K5): dW[V x H] = s0 * v' h + s1 * vm' hm, contraction over the batch axis.
// States are stored row = sample, unit-contiguous, i.e. both operands are MN-major for this product:
// TMA boxes of 64 samples x 64 units (128 B inner, 128B swizzle) land as the canonical MN-major UMMA
// layout (LBO = 8 KB between 64-unit blocks, SBO = 1 KB between 8-sample groups); a_major = b_major = 1
// in the instruction descriptor.  Positive and negative phase accumulate into two TMEM accumulators and
// are combined with their scales in the epilogue.  Split-K over CTAs with a deterministic reduction.
struct alignas(64) TcGradMaps {
    CUtensorMap a[2][2];   // [source][plane]: v-like operand (V units)
    CUtensorMap b[2][2];   // [source][plane]: h-like operand (H units)
};
struct TcGradParams {
    int64_t V, H;
    int32_t kblocks[2];
    int32_t nA[2], nB[2];
    uint32_t combos[2];    // bit (pa * 2 + pb): plane product A_pa x B_pb is contracted
    const uint32_t* a_hiflag[2];
    float scale[2];
    float* out;
    float* partial;
    int32_t m_tiles, n_tiles, splits, stages;
    int32_t nbias;         // trailing CTAs that compute bias gradients (32 columns each), 0: none
    MatView bv, bh, bvm, bhm;
    float* da;
    float* db;
};
constexpr int TG_BN = 128, TG_THREADS = 32 * (4 + 16);
constexpr int TG_PLANE = 128 * TC_BK * 2;           // 16 KB: [2 unit blocks][64 samples][64 units] fp16
__host__ __device__ constexpr int tg_stage_bytes(int halves) { return (halves + 2) * TG_PLANE; }   // one A plane (all row halves) + 2 B planes

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(8192 >> 4) << 16;    // leading byte offset: next 64-unit block
    d |= (uint64_t)(1024 >> 4) << 32;    // stride byte offset: next group of 8 samples
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Bias gradients of a small batch by one trailing CTA of the gradient launch (the launch is latency-bound: a separate
// kernel costs more than these few loads): 32 columns x 20 row lanes, fixed summation order.
__device__ __noinline__ void grad_bias_cta(const TcGradParams& p, int blk) {
    __shared__ float red[20][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int64_t V = p.bv.cols, H = p.bh.cols;
    const int64_t c = (int64_t)blk * 32 + cx;
    const bool vis = c < V;
    const MatView& P = vis ? p.bv : p.bh;
    const MatView& N = vis ? p.bvm : p.bhm;
    const int64_t cc = vis ? c : c - V;
    float sp = 0.0f, sn = 0.0f;
    if (c < V + H) {
        for (int64_t r = ry; r < P.rows; r += 20) sp += mat_load(P, r, cc);
        for (int64_t r = ry; r < N.rows; r += 20) sn += mat_load(N, r, cc);
    }
    red[ry][cx] = p.scale[0] * sp + p.scale[1] * sn;
    __syncthreads();
    if (ry == 0 && c < V + H) {
        float tot = 0.0f;
#pragma unroll
        for (int j = 0; j < 20; ++j) tot += red[j][cx];
        (vis ? p.da : p.db)[cc] = tot;
    }
}

// HALVES: 128-row halves of a work tile (1 or 2).  Two halves share every B (hidden-state) tile load: 64 instead of 96
// B/clk/SM of L2 -> shared-memory operand traffic, which bounds this kernel at the C4 shapes.
// A k-block (64 samples) is contracted in one or two PASSES through one shared-memory stage: pass pa stages A plane pa
// (all row halves) and the B planes it multiplies.  Pass 1 runs only where it is needed: a COUNT operand's hi plane where
// the 128-sample block is flagged, a REAL operand's lo plane always.
// The sample axis is contracted in chunks of <= TC_CHUNK_KB k-blocks, chunks never straddle the positive / negative
// source, each chunk goes into a fresh TMEM stage (2 stages x HALVES x 128 columns) and the epilogue warps add
// scale[source] x chunk in registers (fp32 promotion: tcgen05.mma's own accumulation truncates, see tc_chunks).
template <int HALVES>
__global__ void __launch_bounds__(TG_THREADS, 1) k_grad_tc(const __grid_constant__ TcGradMaps maps, const __grid_constant__ TcGradParams p) {
    constexpr uint32_t IDESC = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(TG_BN >> 3) << 17) |
                               ((uint32_t)(TC_BM >> 4) << 24);   // c = F32, a = b = F16, both MN-major
    constexpr uint32_t TG_TMEM = 2 * HALVES * TG_BN;             // 256 or 512 columns
    constexpr int STAGE = tg_stage_bytes(HALVES);
    constexpr int ROWS = HALVES * TC_BM;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES], bar_empty[TC_MAX_STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_holder;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int svc = warp - 16;    // epilogue warps 0-15, service warps above them (see k_gemm_tc)
    const int nwork = p.m_tiles * p.n_tiles * p.splits;
    const int ngemm = (int)gridDim.x - p.nbias;
    if ((int)blockIdx.x >= ngemm) {
        pdl_trigger();
        pdl_wait();
        grad_bias_cta(p, (int)blockIdx.x - ngemm);
        return;
    }

    if (svc == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(smem_u32(&bar_full[i]), 1); mbar_init(smem_u32(&bar_empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bar_tfull[i]), 1); mbar_init(smem_u32(&bar_tempty[i]), 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    } else if (svc == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(TG_TMEM) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;
    pdl_trigger();      // this CTA holds its TMEM columns: the next launch may start setting up
    pdl_wait();         // operands and bias / flag tables come from the preceding launches

    // k-block range of source s for split `sp`
    auto kb_range = [&](int s, int sp, int& lo, int& hi) {
        lo = (int)((int64_t)p.kblocks[s] * sp / p.splits);
        hi = (int)((int64_t)p.kblocks[s] * (sp + 1) / p.splits);
    };
    // B-plane masks of the two passes of k-block kb of source s (0: pass not run)
    auto pass_masks = [&](int s, int kb, uint32_t& m0, uint32_t& m1) {
        m0 = p.combos[s] & 3u;
        m1 = (p.combos[s] >> 2) & 3u;
        if (p.a_hiflag[s] && (p.a_hiflag[s][kb / 2] & 1u) == 0u) m1 = 0u;      // COUNT operand: hi plane unused in this 128-sample block
    };

    if (svc == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < nwork; w += ngemm) {
                const int tile = w / p.splits, sp = w % p.splits;
                const int m0 = (tile / p.n_tiles) * ROWS, n0 = (tile % p.n_tiles) * TG_BN;
                for (int s = 0; s < 2; ++s) {
                    int lo, hi;
                    kb_range(s, sp, lo, hi);
                    for (int kb = lo; kb < hi; ++kb) {
                        uint32_t pm[2];
                        pass_masks(s, kb, pm[0], pm[1]);
                        for (int pa = 0; pa < 2; ++pa) {
                            if (!pm[pa]) continue;
                            const int nb = (pm[pa] & 2u) ? 2 : 1;
                            mbar_wait_parked(smem_u32(&bar_empty[stage]), phase ^ 1u);
                            const uint32_t full = smem_u32(&bar_full[stage]);
                            const uint32_t sa = smem_base + (uint32_t)stage * STAGE, sb = sa + HALVES * TG_PLANE;
                            mbar_expect_tx(full, (uint32_t)((HALVES + nb) * TG_PLANE));
                            for (int blk = 0; blk < 2 * HALVES; ++blk)
                                tma_load_2d(sa + blk * 8192, &maps.a[s][pa], full, m0 + blk * 64, kb * TC_BK);
                            for (int pb = 0; pb < nb; ++pb)
                                for (int blk = 0; blk < 2; ++blk)
                                    tma_load_2d(sb + pb * TG_PLANE + blk * 8192, &maps.b[s][pb], full, n0 + blk * 64, kb * TC_BK);
                            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (svc == 1) {
        if (lane == 0) {
            int stage = 0, astage = 0;
            uint32_t phase = 0, aphase = 0;
            for (int w = blockIdx.x; w < nwork; w += ngemm) {
                const int sp = w % p.splits;
                for (int s = 0; s < 2; ++s) {
                    int lo, hi;
                    kb_range(s, sp, lo, hi);
                    const int n = hi - lo;
                    if (n <= 0) continue;
                    const int nch = (n + TC_CHUNK_KB - 1) / TC_CHUNK_KB, per = (n + nch - 1) / nch;
                    int in_chunk = 0;
                    uint32_t accumulate = 0;
                    for (int kb = lo; kb < hi; ++kb) {
                        if (in_chunk == 0) {
                            mbar_wait_parked(smem_u32(&bar_tempty[astage]), aphase ^ 1u);
                            tc_fence_after();
                            accumulate = 0;
                        }
                        const uint32_t tmem_d = tmem_base + (uint32_t)(astage * HALVES) * TG_BN;
                        uint32_t pm[2];
                        pass_masks(s, kb, pm[0], pm[1]);
                        for (int pa = 0; pa < 2; ++pa) {
                            if (!pm[pa]) continue;
                            mbar_wait_parked(smem_u32(&bar_full[stage]), phase);
                            tc_fence_after();
                            const uint32_t sa = smem_base + (uint32_t)stage * STAGE, sb = sa + HALVES * TG_PLANE;
#pragma unroll
                            for (int h = 0; h < HALVES; ++h) {
                                uint32_t acc_h = accumulate;
#pragma unroll
                                for (int pb = 0; pb < 2; ++pb) {
                                    if (!((pm[pa] >> pb) & 1u)) continue;
#pragma unroll
                                    for (int k = 0; k < TC_BK / 16; ++k) {   // 16 samples = two 8-sample groups = 2 KB
                                        tc_mma_f16(tmem_d + h * TG_BN, umma_desc_mn_sw128(sa + h * TG_PLANE + k * 2048),
                                                   umma_desc_mn_sw128(sb + pb * TG_PLANE + k * 2048), IDESC, acc_h);
                                        acc_h = 1;
                                    }
                                }
                            }
                            accumulate = 1;
                            tc_commit(smem_u32(&bar_empty[stage]));
                            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        }
                        if (++in_chunk == per || kb + 1 == hi) {
                            tc_commit(smem_u32(&bar_tfull[astage]));
                            if (++astage == 2) { astage = 0; aphase ^= 1u; }
                            in_chunk = 0;
                        }
                    }
                }
            }
        }
    } else if (svc < 0) {
        const int q = warp & 3, slice = warp >> 2;
        int astage = 0;
        uint32_t aphase = 0;
        for (int w = blockIdx.x; w < nwork; w += ngemm) {
            const int tile = w / p.splits, sp = w % p.splits;
            const int64_t m0 = (int64_t)(tile / p.n_tiles) * ROWS, nn0 = (int64_t)(tile % p.n_tiles) * TG_BN;
            float sum[HALVES][32];
#pragma unroll
            for (int h = 0; h < HALVES; ++h)
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[h][i] = 0.0f;
            for (int s = 0; s < 2; ++s) {
                int lo, hi;
                kb_range(s, sp, lo, hi);
                const int n = hi - lo;
                if (n <= 0) continue;
                const int nch = (n + TC_CHUNK_KB - 1) / TC_CHUNK_KB;
                const float sc = p.scale[s];
                for (int c = 0; c < nch; ++c) {
                    mbar_wait_parked(smem_u32(&bar_tfull[astage]), aphase);
                    tc_fence_after();
                    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(astage * HALVES) * TG_BN + (uint32_t)(slice * 32);
#pragma unroll
                    for (int h = 0; h < HALVES; ++h)
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            uint32_t r[16];
                            tmem_ld16(ta + h * TG_BN + g * 16, r);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 16; ++i) sum[h][g * 16 + i] = fmaf(sc, __uint_as_float(r[i]), sum[h][g * 16 + i]);
                        }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[astage]));
                    if (++astage == 2) { astage = 0; aphase ^= 1u; }
                }
            }
            float* dst = (p.splits > 1 ? p.partial + (int64_t)sp * p.V * p.H : p.out);
            const int64_t c0 = nn0 + slice * 32;
            const bool vec = c0 + 32 <= p.H && (p.H & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0u;
#pragma unroll
            for (int h = 0; h < HALVES; ++h) {
                const int64_t row = m0 + h * TC_BM + q * 32 + lane;
                if (row < p.V) {
                    float* d = dst + row * p.H + c0;
                    if (vec) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(d + i) = make_float4(sum[h][i], sum[h][i + 1], sum[h][i + 2], sum[h][i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c0 + i < p.H) d[i] = sum[h][i];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (svc == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TG_TMEM) : "memory");
    }
}

// out[i] = sum over splits (fixed order) of partial[s][i]
__global__ void k_grad_reduce(const float* __restrict__ partial, int splits, int64_t n, float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float t = 0.0f;
        for (int s = 0; s < splits; ++s) t += partial[(int64_t)s * n + i];
        out[i] = t;
    }
}

inline bool tc_make_map_mn(TcContext& t, CUtensorMap* map, const void* ptr, int64_t samples, int64_t units, int64_t ld) {
    return tc_make_map_box(t, map, ptr, samples, units, ld, TC_BK, 64);      // box: 64 units (inner) x 64 samples
}

inline bool tc_grad_supported(const TcGradArgs& a) {
    return a.V >= 1 && a.H >= 1 && a.v.rows >= 1 && a.vm.rows >= 1 && a.v.p0 && a.h.p0 && a.vm.p0 && a.hm.p0;
}

// two-SM variant (gemm_tc2.cuh)
inline bool tc_grad_cg2_eligible(const TcGradArgs& a, int num_sms);
inline int tc_grad_cg2_splits(const TcGradArgs& a, int num_sms);
inline cudaError_t launch_grad_cg2(TcContext& t, const TcGradMaps& maps, TcGradParams& p, int num_sms, cudaStream_t stream);
inline bool& tc_grad_cg2_enabled() {
    static bool on = [] { const char* v = getenv("SCDBM_GRAD_CG2"); return !(v && v[0] == '0'); }();
    return on;
}

// number of K splits that fills the GPU when there are few output tiles
// 256-row work tiles when they still fill the GPU
inline int tc_grad_halves(const TcGradArgs& a, int num_sms) {
    const int64_t tiles2 = ((a.V + 2 * TC_BM - 1) / (2 * TC_BM)) * ((a.H + TG_BN - 1) / TG_BN);
    return tiles2 >= num_sms ? 2 : 1;
}
inline int tc_grad_splits(const TcGradArgs& a, int num_sms) {
    if (tc_grad_cg2_enabled() && tc_grad_cg2_eligible(a, num_sms)) return tc_grad_cg2_splits(a, num_sms);
    const int rows = tc_grad_halves(a, num_sms) * TC_BM;
    const int tiles = (int)((a.V + rows - 1) / rows) * (int)((a.H + TG_BN - 1) / TG_BN);
    const int kb = (int)((a.v.rows + TC_BK - 1) / TC_BK);
    if (tiles >= num_sms) return 1;
    return std::max(1, std::min({(num_sms + tiles - 1) / tiles, kb / 4 + 1, 16}));
}

inline cudaError_t launch_grad_tc(TcContext& t, const TcGradArgs& a, int num_sms, cudaStream_t stream) {
    if (!t.encode) return cudaErrorNotSupported;
    TcGradMaps maps;
    TcGradParams p;
    memset(&p, 0, sizeof(p));
    const int halves = tc_grad_halves(a, num_sms);
    p.V = a.V; p.H = a.H;
    p.out = a.out; p.partial = a.partial; p.splits = std::max(1, a.splits);
    p.m_tiles = (int)((a.V + halves * TC_BM - 1) / (halves * TC_BM));
    p.n_tiles = (int)((a.H + TG_BN - 1) / TG_BN);
    p.scale[0] = a.pos_scale; p.scale[1] = a.neg_scale;
    p.nbias = (a.da && a.db) ? (int)((a.v.cols + a.h.cols + 31) / 32) : 0;
    p.bv = a.v; p.bh = a.h; p.bvm = a.vm; p.bhm = a.hm;
    p.da = a.da; p.db = a.db;
    const MatView* av[2] = {&a.v, &a.vm};
    const MatView* bv[2] = {&a.h, &a.hm};
    for (int s = 0; s < 2; ++s) {
        const MatView& A = *av[s];
        const MatView& B = *bv[s];
        p.kblocks[s] = (int)((A.rows + TC_BK - 1) / TC_BK);
        p.nA[s] = A.p1 ? 2 : 1;
        p.nB[s] = B.p1 ? 2 : 1;
        // plane products; real x real drops lo*lo
        uint32_t c = 0;
        for (int pa = 0; pa < p.nA[s]; ++pa)
            for (int pb = 0; pb < p.nB[s]; ++pb)
                if (!(pa == 1 && pb == 1 && A.kind == MAT_REAL && B.kind == MAT_REAL)) c |= 1u << (pa * 2 + pb);
        p.combos[s] = c;
        p.a_hiflag[s] = (A.kind == MAT_COUNT) ? A.hiflag : nullptr;
        bool ok = tc_make_map_mn(t, &maps.a[s][0], A.p0, A.rows, a.V, A.ld);
        ok = ok && tc_make_map_mn(t, &maps.a[s][1], A.p1 ? A.p1 : A.p0, A.rows, a.V, A.ld);
        ok = ok && tc_make_map_mn(t, &maps.b[s][0], B.p0, B.rows, a.H, B.ld);
        ok = ok && tc_make_map_mn(t, &maps.b[s][1], B.p1 ? B.p1 : B.p0, B.rows, a.H, B.ld);
        if (!ok) return cudaErrorInvalidValue;
    }
    if (p.nbias == 0 && tc_grad_cg2_enabled() && tc_grad_cg2_eligible(a, num_sms)) {
        cudaError_t e = launch_grad_cg2(t, maps, p, num_sms, stream);
        if (e != cudaSuccess || p.splits == 1) return e;
        const int64_t n = a.V * a.H;
        return launch_pdl(k_grad_reduce, dim3((unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8)), dim3(256), 0, stream, a.partial, p.splits, n, a.out);
    }
    p.stages = halves == 2 ? 3 : 4;
    const int smem = p.stages * tg_stage_bytes(halves) + 1024;
    const int slot = halves == 2 ? 15 : 14;
    if (!t.attr_set[slot]) {
        cudaError_t e = halves == 2 ? cudaFuncSetAttribute(k_grad_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                                    : cudaFuncSetAttribute(k_grad_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        t.attr_set[slot] = true;
    }
    const int nwork = p.m_tiles * p.n_tiles * p.splits;
    cudaError_t le;
    if (halves == 2) le = launch_pdl(k_grad_tc<2>, dim3(std::min(nwork, num_sms) + p.nbias), dim3(TG_THREADS), (size_t)smem, stream, maps, p);
    else le = launch_pdl(k_grad_tc<1>, dim3(std::min(nwork, num_sms) + p.nbias), dim3(TG_THREADS), (size_t)smem, stream, maps, p);
    if (le != cudaSuccess) return le;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || p.splits == 1) return e;
    const int64_t n = a.V * a.H;
    le = launch_pdl(k_grad_reduce, dim3((unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8)), dim3(256), 0, stream, a.partial, p.splits, n, a.out);
    if (le != cudaSuccess) return le;
    return cudaGetLastError();
}

}  // namespace scdbm
