// Tensor-core candidate pass for the full-ranking evaluation (tcgen05 / TMEM / TMA, sm_100a).
//
// The exact ranking (topk.cu) evaluates every user x item score as a k-sequential fp32 fmaf chain on
// the FFMA pipe: 2*U*I*d flops, ~21 TFLOP/s.  fp32 has no tensor-core MMA; TF32 (10 mantissa bits)
// changes ~11 % of the top-20 lists when used directly.  This pass therefore only NOMINATES: it
// computes all scores with tcgen05.mma.kind::tf32 (fp32 accumulators in TMEM), masks the seen
// items, and keeps the K' > K best approximate scores per user.  gr_topk_rescore then re-scores
// those K' candidates with the exact chain, ranks them canonically and PROVES the result: every
// item outside the candidate list has approximate score <= a_min (the K'-th kept), and
// |s_tf32 - s_exact| <= eps * |u| * max|i|, so if a_min + bound < (K-th exact score) no outside item
// can enter or tie into the top-K.  Users for which the proof fails are flagged and re-ranked by the
// exact kernel; results are therefore always bit-identical to the exact path.
//
// One CTA = 128 eval users (UMMA M = 128) against one item range (blockIdx.y; 1-4 ranges per user tile fill
// the SMs evenly) in tiles of 64 items (UMMA N = 64):
//   warps 0-1            : load the user tile once (rows gathered through eval_users) into the K-major
//                          SWIZZLE_128B canonical layout (32 tf32 per 128-byte row, 8-row atoms, 16-byte
//                          chunks XOR-swizzled); then
//   warp 0, one lane     : TMA producer — per item tile d/32 cp.async.bulk.tensor.2d boxes (32 floats x 64
//                          rows, SWIZZLE_128B; rows past the catalogue are zero-filled by the hardware) into
//                          a two-stage ring, completion by mbarrier complete_tx
//   warp 1, one lane     : MMA issuer — d/8 tcgen05.mma (K = 8 per instruction) per tile into one of four
//                          64-column TMEM accumulators, tcgen05.commit to the stage / accumulator mbarriers
//   warps 2-5  epilogue  : tcgen05.ld of a whole 64-column tile into registers (the accumulator is released at
//                          once); thread = TMEM lane = one user.  Selection is
//                          max-first: masked columns (seen items, catalogue tail) are set to -inf on the rare
//                          tiles that have any, eight 8-column FMNMX trees give the group maxima, and only a
//                          group whose maximum beats the user's admission threshold is staged (8 values) and
//                          appended to the user's pending queue; the queues are drained in lock-step into the
//                          users' candidate sets (unsorted K' entries + tracked minimum) in shared memory.
// Two CTAs are resident per SM when the candidate sets fit (d = 64: K' <= 32), so eight epilogue warps
// share the four schedulers and one CTA's MMA overlaps the other's selection.
#include <cuda.h>
#include <math_constants.h>

#include <cstdlib>
#include <cstring>

#include "gr_common.cuh"

namespace gr {

constexpr int TC_M = 128;       // users per CTA
constexpr int TC_N = 64;        // items per tile
constexpr int TC_KB = 32;       // tf32 elements per 128-byte swizzle row
constexpr int TC_STAGES = 2;    // item-tile stages in shared memory
constexpr int TC_ACC = 4;       // TMEM accumulators of TC_N columns each
constexpr int TC_THREADS = 192; // warps 0-1: user tile, then TMA producer (warp 0) / MMA issuer (warp 1), one lane each; 2-5: epilogue
constexpr int TC_GROUP = 8;     // columns per selection group
constexpr int TC_KPRIME_MAX = 64;
constexpr uint32_t TC_A_TILE = TC_M * 128;   // one k-block (32 tf32) of the user tile: 16 KB
constexpr uint32_t TC_B_TILE = TC_N * 128;   // one k-block of an item tile: 8 KB

struct TcArgs {
    const float *user_emb;
    long long ldu;
    const float *item_emb;  // row i = item id i
    long long ldi;
    int d;
    const int64_t *eval_users;
    int n_eval;
    int n_items;
    const int64_t *seen_indptr;
    const int32_t *seen_items;
    int kprime;          // candidates kept per user
    int split_items;     // items per blockIdx.y (multiple of TC_N); gridDim.y item ranges fill the SMs evenly
    float *cand_scores;  // [n_eval][gridDim.y][kprime] approximate scores (unsorted)
    int *cand_ids;       // [n_eval][gridDim.y][kprime]
    int *cand_cnt;       // [n_eval][gridDim.y]
    int *error;          // device flag: 1 = the kernel could not run (re-scoring flags every row)
    int debug;           // GR_TC_DEBUG bits (experiments): 1 = epilogue skips selection, 2 = producers skip loads,
                         // 4 = take the could-not-run path, 8 = queues are dropped instead of drained,
                         // 16 = filter only (no appends), bits 16.. = producer back-off in ns
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one arrival + `bytes` of pending TMA traffic: the phase completes when the copies have landed
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// waiting role that is NOT on the critical path (TMA producer, MMA issuer).  Measured (ncu source counters): a poll
// loop with nanosleep between polls — test_wait + nanosleep(100), and before that try_wait with a suspend-time hint
// (NANOSLEEP.SYNCS, woken by every mbarrier event of the CTA) — iterated every ~25 ns and executed 39 % of the
// kernel's instructions on the two schedulers these single-lane warps share with epilogue warps, while the
// epilogue's plain try_wait loop (the hardware suspends the thread inside try_wait) cost 1 %.  So: the same loop.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity, unsigned sleep_ns = 0) {
    if (sleep_ns == 0) {
        mbar_wait(bar, parity);
        return;
    }
    uint32_t done = 0;
    while (true) {      // GR_TC_DEBUG bits 16..: poll with a nanosleep of that many ns (experiments)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(sleep_ns);
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one TMA box: coordinates (c0 = element along d, c1 = item row) of the tensor map -> shared memory;
// the bytes are reported to `bar` (complete_tx)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// 32 columns of this warp's 32 TMEM lanes into r[]; the registers are valid only after tmem_ld_wait(r)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]),
          "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]),
          "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
        : "r"(taddr)
        : "memory");
}
// tcgen05.wait::ld; the registers of the load are in/out operands so that no use of them is scheduled above it
__device__ __forceinline__ void tmem_ld_wait(float (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
                   "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]),
                   "+f"(r[16]), "+f"(r[17]), "+f"(r[18]), "+f"(r[19]), "+f"(r[20]), "+f"(r[21]), "+f"(r[22]), "+f"(r[23]),
                   "+f"(r[24]), "+f"(r[25]), "+f"(r[26]), "+f"(r[27]), "+f"(r[28]), "+f"(r[29]), "+f"(r[30]), "+f"(r[31])
                 :
                 : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address
// and offsets in 16-byte units, stride between 8-row atoms = 1024 B, version 1, layout type 2.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t desc = 0;
    desc |= (uint64_t)((smem_addr & 0x3ffff) >> 4);  // start address, bits [0,14)
    desc |= (uint64_t)0 << 16;                       // leading byte offset (unused for swizzled K-major)
    desc |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset, bits [32,46)
    desc |= (uint64_t)1 << 46;                       // version
    desc |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return desc;
}
// byte offset of 16-byte chunk `c16` (0..7) of row `row` inside a [rows][128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_offset(int row, int c16) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((c16 ^ (row & 7)) << 4));
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, N at bits [17,23) (N>>3), M at [24,29) (M>>4)
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((TC_N >> 3) << 17) | ((TC_M >> 4) << 24);

// Per-user candidate set = the K' best approximate scores seen so far as an UNSORTED array, position-major in
// shared memory (entry p of user m at [p * 128 + m], conflict-free for a warp), plus — in registers — the worst
// kept score (the admission threshold) and its position.  A new entry overwrites the worst one and the minimum is
// recomputed by a scan: K' independent shared-memory loads and a compare/select tree, ~100 cycles of dependent
// latency and the SAME instruction stream for every lane.  (The first versions kept a binary heap: fewer
// instructions per insertion, but a sift-down is five dependent load -> compare -> store rounds with data-dependent
// trip counts, ~3x the latency, and the lock-step drain of a warp is bound by exactly that latency.)
// Ties are broken arbitrarily: the nomination only has to guarantee "every rejected or evicted item has
// approximate score <= the final threshold".
struct SelState {
    int cnt;
    float thr;
    int minpos;
};

__device__ __forceinline__ SelState tc_set_insert(float *hs, int *hi, int m, int K, int cnt, int minpos, float s, int id) {
    const int pos = cnt < K ? cnt : minpos;       // filling: append; full: replace the worst entry
    hs[pos * TC_M + m] = s;
    hi[pos * TC_M + m] = id;
    SelState st;
    st.cnt = cnt < K ? cnt + 1 : cnt;
    st.thr = -CUDART_INF_F;
    st.minpos = 0;
    if (st.cnt == K) {                            // (K is a multiple of 8)
        float mn = CUDART_INF_F;
        int mp = 0;
        for (int p0 = 0; p0 < K; p0 += 8) {
            float e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) e[j] = hs[(p0 + j) * TC_M + m];
            // (value, position) minimum tree over the eight entries
            const bool l01 = e[1] < e[0], l23 = e[3] < e[2], l45 = e[5] < e[4], l67 = e[7] < e[6];
            const float v01 = l01 ? e[1] : e[0], v23 = l23 ? e[3] : e[2], v45 = l45 ? e[5] : e[4], v67 = l67 ? e[7] : e[6];
            const int p01 = l01 ? 1 : 0, p23 = l23 ? 3 : 2, p45 = l45 ? 5 : 4, p67 = l67 ? 7 : 6;
            const bool la = v23 < v01, lb = v67 < v45;
            const float va = la ? v23 : v01, vb = lb ? v67 : v45;
            const int pa = la ? p23 : p01, pb = lb ? p67 : p45;
            const bool lc = vb < va;
            const float vc = lc ? vb : va;
            const int pc = lc ? pb : pa;
            if (vc < mn) {
                mn = vc;
                mp = p0 + pc;
            }
        }
        st.thr = mn;
        st.minpos = mp;
    }
    return st;
}

__device__ __forceinline__ float max8(const float *v) {
    return fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
}

// Columns that beat a user's (possibly stale) threshold are only APPENDED to the user's pending queue
// (position-major like the candidate set: entry q of user m at [q * 128 + m]); the insertions run when some
// lane's queue is nearly full: all lanes then pop their queues in lockstep, so a push
// iteration serves every lane that has work instead of the one lane that happened to hit in this chunk
// (one push per chunk with 1 of 32 lanes active was the dominant cost of the first version).  A stale
// threshold only admits extra queue entries; each is re-checked against the current root when popped.
constexpr int TC_PEND = 16;     // queue entries per user

__device__ __forceinline__ void pend_append(uint32_t slot, float s, int id) {
    asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.b32 [%0+%3], %2;" ::"r"(slot), "f"(s), "r"(id), "n"(TC_PEND * TC_M * 4)
                 : "memory");
}

__device__ __noinline__ SelState tc_drain(float *hs, int *hi, const float *ps, const int *pi, int m, int K, int cnt,
                                          float thr, int minpos, int np) {
    for (int q = 0; __any_sync(0xffffffffu, q < np); ++q) {
        if (q < np) {
            const float sj = ps[q * TC_M + m];
            if (sj > thr) {
                const SelState st = tc_set_insert(hs, hi, m, K, cnt, minpos, sj, pi[q * TC_M + m]);
                cnt = st.cnt;
                thr = st.thr;
                minpos = st.minpos;
            }
        }
    }
    SelState out;
    out.cnt = cnt;
    out.thr = thr;
    out.minpos = minpos;
    return out;
}

__global__ void __launch_bounds__(TC_THREADS, 2) topk_tc_candidates_kernel(const TcArgs a,
                                                                           const __grid_constant__ CUtensorMap tmap) {
    // SWIZZLE_128B operand tiles need 1024-byte alignment; the kernel has no static shared memory, so
    // the dynamic window starts at the CTA's (1 KB-granular) allocation.  Checked, not assumed.
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn;
    if ((smem_u32(smem_dyn) & 1023u) != 0 || (a.debug & 4)) {   // never seen (debug bit 4 forces it for the test); the re-scoring pass then hands every row to the exact kernel
        if (threadIdx.x == 0) *a.error = 1;
        return;
    }
    const int d = a.d;
    const int nkb = d / TC_KB;                       // k-blocks of 32 tf32
    unsigned char *sA = smem_raw;                    // [nkb][128 rows][128 B]
    unsigned char *sB = sA + (size_t)nkb * TC_A_TILE;             // [stages][nkb][64 rows][128 B]
    float *ls = reinterpret_cast<float *>(sB + (size_t)TC_STAGES * nkb * TC_B_TILE);     // kept scores [kprime][128]
    int *li = reinterpret_cast<int *>(ls + (size_t)a.kprime * TC_M);                      // kept ids    [kprime][128]
    float *pend_s = reinterpret_cast<float *>(li + (size_t)a.kprime * TC_M);              // pending scores [TC_PEND][128]
    int *pend_i = reinterpret_cast<int *>(pend_s + TC_PEND * TC_M);                       // pending ids    [TC_PEND][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(pend_i + TC_PEND * TC_M);
    uint64_t *b_full = bars, *b_empty = bars + TC_STAGES, *t_full = bars + 2 * TC_STAGES,
             *t_empty = bars + 2 * TC_STAGES + TC_ACC, *a_full = bars + 2 * TC_STAGES + 2 * TC_ACC;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_full + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned sleep_ns = (unsigned)(a.debug >> 16);
    const int row0 = blockIdx.x * TC_M;
    const int item_lo = (int)min((long long)a.n_items, (long long)blockIdx.y * a.split_items);
    const int item_hi = (int)min((long long)a.n_items, (long long)item_lo + a.split_items);
    const int n_it = item_hi - item_lo;
    const int n_tiles = (n_it + TC_N - 1) / TC_N;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&b_full[s], 1);     // one arrive.expect_tx by the TMA producer
            mbar_init(&b_empty[s], 1);    // tcgen05.commit
        }
        for (int s = 0; s < TC_ACC; ++s) {
            mbar_init(&t_full[s], 1);     // tcgen05.commit
            mbar_init(&t_empty[s], 4);    // one arrival per epilogue warp
        }
        mbar_init(a_full, 64);            // the two loader warps of the user tile
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TC_ACC * TC_N);  // four 64-column fp32 accumulators
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 2) {
        // ===== user tile, once (rows gathered through eval_users) =====
        const int f4_per_row = d / 4;
        for (int idx = tid; idx < TC_M * f4_per_row; idx += 64) {
            const int r = idx / f4_per_row, f = idx % f4_per_row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < a.n_eval) v = __ldg(reinterpret_cast<const float4 *>(a.user_emb + a.eval_users[row0 + r] * a.ldu) + f);
            *reinterpret_cast<float4 *>(sA + (size_t)(f >> 3) * TC_A_TILE + sw128_offset(r, f & 7)) = v;
        }
        fence_proxy_async();
        mbar_arrive(a_full);
        if (warp == 0 && lane == 0) {
            // ===== TMA producer =====
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % TC_STAGES;
                if (t >= TC_STAGES) mbar_wait_relaxed(&b_empty[s], ((t / TC_STAGES) - 1) & 1, sleep_ns);
                unsigned char *dst = sB + (size_t)s * nkb * TC_B_TILE;
                if ((a.debug & 2) && t >= TC_STAGES) {
                    mbar_arrive(&b_full[s]);
                    continue;
                }
                mbar_arrive_expect_tx(&b_full[s], (uint32_t)nkb * TC_B_TILE);
                for (int kb = 0; kb < nkb; ++kb)
                    tma_load_2d(dst + (size_t)kb * TC_B_TILE, &tmap, &b_full[s], kb * TC_KB, item_lo + t * TC_N);
            }
        } else if (warp == 1 && lane == 0) {
            // ===== MMA issuer =====
            mbar_wait_relaxed(a_full, 0, sleep_ns);
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % TC_STAGES, acc = t % TC_ACC;
                mbar_wait_relaxed(&b_full[s], (t / TC_STAGES) & 1, sleep_ns);
                if (t >= TC_ACC) mbar_wait_relaxed(&t_empty[acc], ((t / TC_ACC) - 1) & 1, sleep_ns);
                tc_fence_after();
                const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB + (size_t)s * nkb * TC_B_TILE);
                const uint32_t tmem_d = tmem_base + (uint32_t)acc * TC_N;
                for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
                    for (int k = 0; k < TC_KB / 8; ++k) {   // UMMA K = 8 tf32 = 32 bytes
                        const uint64_t da = make_sw128_desc(a_base + kb * TC_A_TILE + k * 32);
                        const uint64_t db = make_sw128_desc(b_base + kb * TC_B_TILE + k * 32);
                        umma_tf32(tmem_d, da, db, kIdescTf32, (kb | k) ? 1u : 0u);
                    }
                }
                umma_commit(&b_empty[s]);   // shared-memory stage may be refilled once these MMAs retire
                umma_commit(&t_full[acc]);  // accumulator ready for the epilogue
            }
        }
    } else {
        // ===== epilogue (warps 2-5): thread = TMEM lane = one user row; warp w reads TMEM lanes 32 (w % 4).. =====
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const uint32_t tmem_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const bool user_ok = row0 + m < a.n_eval;
        const int K = a.kprime;
        int cnt = 0, minpos = 0;     // kept candidates, position of the worst one
        // next free entry of this user's pending queue, as a shared-space byte address (ids 8 KB further)
        const uint32_t wp0 = smem_u32(pend_s + m), wp_limit = wp0 + (TC_PEND - TC_GROUP) * TC_M * 4;
        uint32_t wp = wp0;
        // admission threshold: the worst kept score once the set is full; a row beyond n_eval never admits anything
        float thr = user_ok ? -CUDART_INF_F : CUDART_INF_F;
        constexpr int kNoSeen = 0x7fffffff;
        int sc = 0, se = 0;
        if (user_ok && a.seen_indptr) {
            long long lo = a.seen_indptr[row0 + m], hi = a.seen_indptr[row0 + m + 1];
            se = (int)hi;
            while (lo < hi) {   // first seen id >= item_lo
                const long long mid = (lo + hi) >> 1;
                if (a.seen_items[mid] < item_lo) lo = mid + 1; else hi = mid;
            }
            sc = (int)lo;
        }
        int next_seen = sc < se ? a.seen_items[sc] : kNoSeen;     // absolute item id

        // selection over one 64-column tile held in registers (scores of items cbase .. cbase + 63).  The whole tile
        // is one straight-line block up to the group branches, so the eight FMNMX trees overlap (two epilogue warps
        // per scheduler cannot hide dependent-issue latency by themselves: "wait" was the top stall reason).
        auto select = [&](float (&v)[TC_N], const int cbase) {
            if (a.debug & 1) return;
            // columns that must not be nominated: seen items of this user, padding beyond the range
            unsigned kill0 = 0, kill1 = 0;
            while (next_seen < cbase + TC_N) {        // next_seen is prefetched: no load on the common path
                const int o = next_seen - cbase;
                if (o >= 32) kill1 |= 1u << (o - 32);
                else if (o >= 0) kill0 |= 1u << o;
                ++sc;
                next_seen = sc < se ? a.seen_items[sc] : kNoSeen;
            }
            const int room = item_hi - cbase;
            if (room < TC_N) {
                kill0 |= room <= 0 ? 0xffffffffu : (room >= 32 ? 0u : ~((1u << room) - 1u));
                kill1 |= room <= 32 ? 0xffffffffu : ~((1u << (room - 32)) - 1u);
            }
            if (kill0 | kill1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if ((kill0 >> j) & 1u) v[j] = -CUDART_INF_F;
                    if ((kill1 >> j) & 1u) v[32 + j] = -CUDART_INF_F;
                }
            }
            // max-first filter: one FMNMX tree per 8 columns, one compare per group
            unsigned hit = 0;
#pragma unroll
            for (int gq = 0; gq < TC_N / TC_GROUP; ++gq) hit |= (unsigned)(max8(v + gq * TC_GROUP) > thr) << gq;
            if (!__any_sync(0xffffffffu, hit != 0)) return;       // the warp stays converged here
            if (a.debug & 16) return;
            // groups some lane passed: predicated appends to the pending queues (warp-uniform branch per group)
#pragma unroll
            for (int gq = 0; gq < TC_N / TC_GROUP; ++gq) {
                if (__any_sync(0xffffffffu, (hit >> gq) & 1u)) {
#pragma unroll
                    for (int j = 0; j < TC_GROUP; ++j) {
                        if (v[gq * TC_GROUP + j] > thr) {
                            pend_append(wp, v[gq * TC_GROUP + j], cbase + gq * TC_GROUP + j);
                            wp += TC_M * 4;
                        }
                    }
                    if (__any_sync(0xffffffffu, wp > wp_limit)) {   // the next group may not fit
                        if (!(a.debug & 8)) {
                            const SelState st = tc_drain(ls, li, pend_s, pend_i, m, K, cnt, thr, minpos, (int)(wp - wp0) / (TC_M * 4));
                            cnt = st.cnt;
                            thr = st.thr;
                            minpos = st.minpos;
                        }
                        wp = wp0;
                    }
                }
            }
        };

        float v[TC_N];
        for (int t = 0; t < n_tiles; ++t) {
            const int acc = t % TC_ACC;
            mbar_wait(&t_full[acc], (t / TC_ACC) & 1);
            tc_fence_after();
            tmem_ld32_issue(tmem_row + (uint32_t)(acc * TC_N), *reinterpret_cast<float(*)[32]>(v));
            tmem_ld32_issue(tmem_row + (uint32_t)(acc * TC_N + 32), *reinterpret_cast<float(*)[32]>(v + 32));
            tmem_ld_wait(*reinterpret_cast<float(*)[32]>(v));
            tmem_ld_wait(*reinterpret_cast<float(*)[32]>(v + 32));
            // the accumulator is free as soon as the tile is in registers: the MMA of tile t + 4 can start
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc]);
            select(v, item_lo + t * TC_N);
        }
        {
            const SelState st = tc_drain(ls, li, pend_s, pend_i, m, K, cnt, thr, minpos, (int)(wp - wp0) / (TC_M * 4));
            cnt = st.cnt;
        }
        if (user_ok) {
            const size_t seg = (size_t)(row0 + m) * gridDim.y + blockIdx.y;
            a.cand_cnt[seg] = cnt;
            for (int p = 0; p < K; ++p) {
                a.cand_scores[seg * K + p] = p < cnt ? ls[p * TC_M + m] : -CUDART_INF_F;
                a.cand_ids[seg * K + p] = p < cnt ? li[p * TC_M + m] : -1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, TC_ACC * TC_N);
    }
}

// ---------------------------------------------------------------------------------------------------
// exact re-scoring of the candidates + proof of completeness
// ---------------------------------------------------------------------------------------------------
struct RescoreArgs {
    const float *user_emb;
    long long ldu;
    const float *item_emb;
    long long ldi;
    int d;
    const int64_t *eval_users;
    int n_eval;
    int n_items;
    const float *cand_scores;
    const int *cand_ids;
    const int *cand_cnt;
    int kprime, k, n_seg;
    float eps;
    const float *max_item_norm;  // device scalar
    const int *error;            // set by the nomination kernel when it could not run
    int64_t *out_ids;
    float *out_scores;
    int *flags;       // [n_eval] 1 = not proven, re-rank exactly
    int *n_flagged;   // device counter
};

// One warp per user; NQ candidates per lane (S segments of K' candidates, S * K' <= 32 * NQ).
template <int NQ>
__global__ void __launch_bounds__(256) topk_rescore_kernel(const RescoreArgs a) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= a.n_eval) return;
    if (*a.error) {                                     // no candidates: every row goes to the exact kernel
        if (lane == 0) {
            a.flags[row] = 1;
            atomicAdd(a.n_flagged, 1);
        }
        return;
    }
    const float *u = a.user_emb + a.eval_users[row] * a.ldu;
    const int K = a.k, KP = a.kprime, S = a.n_seg;
    const int total = S * KP;
    int id[NQ];
    float sc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int c = lane + 32 * q;
        id[q] = -1;
        if (c < total && (c % KP) < a.cand_cnt[(size_t)row * S + c / KP]) id[q] = a.cand_ids[(size_t)row * total + c];
        sc[q] = -CUDART_INF_F;
        if (id[q] >= 0) {
            const float4 *it4 = reinterpret_cast<const float4 *>(a.item_emb + (long long)id[q] * a.ldi);
            const float4 *u4 = reinterpret_cast<const float4 *>(u);
            float s = 0.f;
            for (int k4 = 0; k4 < a.d / 4; ++k4) {          // the exact chain, k-sequential (d % 32 == 0)
                const float4 iv = __ldg(it4 + k4), uv = __ldg(u4 + k4);
                s = __fmaf_rn(uv.x, iv.x, s);
                s = __fmaf_rn(uv.y, iv.y, s);
                s = __fmaf_rn(uv.z, iv.z, s);
                s = __fmaf_rn(uv.w, iv.w, s);
            }
            sc[q] = s;
        }
    }
    float un = 0.f;
    for (int k = lane; k < a.d; k += 32) un += __ldg(u + k) * __ldg(u + k);
#pragma unroll
    for (int o = 16; o; o >>= 1) un += __shfl_xor_sync(0xffffffffu, un, o);
    // canonical rank of each candidate among the candidates
    float kth = -CUDART_INF_F;
    int have_k = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        int rank = 0;
#pragma unroll
        for (int q2 = 0; q2 < NQ; ++q2) {
            for (int o = 0; o < 32; ++o) {
                const float os = __shfl_sync(0xffffffffu, sc[q2], o);
                const int oid = __shfl_sync(0xffffffffu, id[q2], o);
                if (oid >= 0 && id[q] >= 0 && (os > sc[q] || (os == sc[q] && oid < id[q]))) ++rank;
            }
        }
        if (id[q] >= 0 && rank < K) {
            a.out_ids[(size_t)row * K + rank] = id[q];
            a.out_scores[(size_t)row * K + rank] = sc[q];
        }
        const bool is_kth = id[q] >= 0 && rank == K - 1;
        const unsigned bal = __ballot_sync(0xffffffffu, is_kth);
        if (bal) {
            kth = __shfl_sync(0xffffffffu, sc[q], __ffs(bal) - 1);
            have_k = 1;
        }
    }
    // proof: an item outside segment s's candidates has approximate score <= that segment's worst kept
    // candidate (only full segments rejected anything), so exact <= a_out + bound for every outside item
    float a_out = -CUDART_INF_F;
    for (int sgm = 0; sgm < S; ++sgm) {
        if (a.cand_cnt[(size_t)row * S + sgm] < KP) continue;
        float a_min = CUDART_INF_F;
        for (int c = lane; c < KP; c += 32) a_min = fminf(a_min, a.cand_scores[((size_t)row * S + sgm) * KP + c]);
#pragma unroll
        for (int o = 16; o; o >>= 1) a_min = fminf(a_min, __shfl_xor_sync(0xffffffffu, a_min, o));
        a_out = fmaxf(a_out, a_min);
    }
    const float bound = a.eps * sqrtf(un) * __ldg(a.max_item_norm) * 1.0001f;
    bool proven;
    if (a_out == -CUDART_INF_F) proven = have_k;        // every unmasked item was a candidate
    else proven = have_k && (a_out + bound < kth);
    if (!isfinite(kth)) proven = false;                 // -inf inside the list: leave it to the exact kernel
    if (lane == 0) {
        a.flags[row] = proven ? 0 : 1;
        if (!proven) atomicAdd(a.n_flagged, 1);
    }
}

__global__ void max_row_norm_kernel(const float *x, long long ld, int n, int d, float *out) {
    const int lane = threadIdx.x & 31;
    float best = 0.f;
    for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
        float s = 0.f;
        for (int k = lane; k < d; k += 32) { const float v = x[(long long)r * ld + k]; s += v * v; }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        best = fmaxf(best, sqrtf(s));
    }
    if (lane == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(best));   // non-negative floats order as ints
}

static size_t tc_smem_bytes(int d, int kprime) {
    return (size_t)(d / TC_KB) * ((size_t)TC_A_TILE + (size_t)TC_STAGES * TC_B_TILE) + (size_t)kprime * TC_M * 8 +
           (size_t)TC_PEND * TC_M * 8 + 256;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libgr_b200.so does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// item table [n_items][d] fp32 (row stride ldi) as a 2-D tensor map with boxes of 32 floats x TC_N rows,
// SWIZZLE_128B: the box lands in shared memory in exactly the K-major layout make_sw128_desc describes
static bool make_item_tensor_map(CUtensorMap *tm, const float *item_emb, int64_t ldi, int64_t n_items, int d) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)n_items};
    const cuuint64_t gstride[1] = {(cuuint64_t)ldi * 4};
    const cuuint32_t box[2] = {(cuuint32_t)TC_KB, (cuuint32_t)TC_N};
    const cuuint32_t estride[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(item_emb), gdim, gstride, box, estride,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
constexpr size_t kSmemTwoCtas = (228 * 1024) / 2 - 1024;   // per CTA when two share an SM (1 KB reserved each)
constexpr size_t kSmemOneCta = 227 * 1024;

}  // namespace gr

using namespace gr;

// Tensor-core nomination + exact re-scoring.  Outputs: topk_ids/topk_scores for the proven rows,
// flags[row] = 1 and *n_flagged for rows that must be re-ranked with gr_score_topk.  d must be a
// multiple of 32 (d = 32 or 64 fit shared memory), k <= kprime <= 64.
// workspace: gr_topk_tc_workspace_bytes(n_eval, kprime).
// item ranges per user tile: the count (with S * K' <= 128 candidates per user for the re-scoring pass)
// that minimises the number of CTA waves per unit of work
static int tc_splits(int64_t n_eval, int32_t d, int32_t kprime) {
    const long long tiles = (n_eval + TC_M - 1) / TC_M;
    if (tiles <= 0) return 1;
    const int per_sm = (tc_smem_bytes(d, kprime) <= kSmemTwoCtas) ? 2 : 1;
    const long long slots = (long long)sm_count() * per_sm;
    const int s_max = 128 / kprime > 0 ? 128 / kprime : 1;
    int best = 1;
    double best_cost = 1e30;
    for (int sp = 1; sp <= s_max; ++sp) {
        const double cost = (double)((tiles * sp + slots - 1) / slots) / sp;
        if (cost < best_cost * 0.95) { best_cost = cost; best = sp; }   // more ranges only for a >5 % gain
    }
    return best;
}

// sized for the largest number of item ranges (128 / K'), so it does not depend on the catalogue
extern "C" size_t gr_topk_tc_workspace_bytes(int64_t n_eval, int32_t kprime) {
    if (n_eval < 0 || kprime <= 0 || kprime > TC_KPRIME_MAX) return 0;
    const int s_max = 128 / kprime;
    return (size_t)n_eval * s_max * kprime * 8 + (size_t)n_eval * s_max * 4 + 256;
}

extern "C" int gr_topk_tc_supported(int32_t d, int32_t kprime) {
    if (!encode_tiled_fn()) return 0;      // the driver has no cuTensorMapEncodeTiled: exact kernel only
    return (d > 0 && d % TC_KB == 0 && kprime > 0 && kprime % 8 == 0 && kprime <= TC_KPRIME_MAX &&
            tc_smem_bytes(d, kprime) <= kSmemOneCta) ? 1 : 0;
}

extern "C" int gr_score_topk_tc(const float *user_emb, int64_t ldu, const float *item_emb, int64_t ldi, int32_t d,
                                const int64_t *eval_users, int64_t n_eval, int64_t n_items,
                                const int64_t *seen_indptr, const int32_t *seen_items, int32_t k, int32_t kprime,
                                int64_t *topk_ids, float *topk_scores, int32_t *flags, int32_t *n_flagged,
                                void *workspace, size_t workspace_bytes, void *stream) {
    if (!user_emb || !item_emb || !eval_users || !topk_ids || !topk_scores || !flags || !n_flagged || !workspace)
        return GR_ERR_INVALID;
    if ((seen_indptr == nullptr) != (seen_items == nullptr)) return GR_ERR_INVALID;
    if (n_eval < 0 || n_items <= 0 || k <= 0 || k > kprime) return GR_ERR_INVALID;
    if (!gr_topk_tc_supported(d, kprime) || (ldu & 3) || (ldi & 3) || ldu < d || ldi < d) return GR_ERR_UNSUPPORTED;
    if (!aligned16(user_emb) || !aligned16(item_emb)) return GR_ERR_INVALID;
    if (n_eval > 0x7fffffffLL || n_items > 0x7fffffffLL) return GR_ERR_OVERFLOW;
    if (workspace_bytes < gr_topk_tc_workspace_bytes(n_eval, kprime)) return GR_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GR_CUDA_CHECK(cudaMemsetAsync(n_flagged, 0, 4, s));
    if (n_eval == 0) return GR_OK;
    const long long tiles_total = (n_items + TC_N - 1) / TC_N;
    int n_seg = tc_splits(n_eval, d, kprime);
    { const char *e = getenv("GR_TC_SPLITS"); if (e && atoi(e) > 0) n_seg = atoi(e) < 128 / kprime ? atoi(e) : 128 / kprime; }
    if (n_seg > tiles_total / 16) n_seg = (int)(tiles_total / 16 > 0 ? tiles_total / 16 : 1);   // >= 1024 items per range
    const long long split_items = ((tiles_total + n_seg - 1) / n_seg) * TC_N;
    n_seg = (int)((n_items + split_items - 1) / split_items);
    float *cand_scores = static_cast<float *>(workspace);
    int *cand_ids = reinterpret_cast<int *>(cand_scores + (size_t)n_eval * n_seg * kprime);
    int *cand_cnt = cand_ids + (size_t)n_eval * n_seg * kprime;
    float *max_norm = reinterpret_cast<float *>(cand_cnt + (size_t)n_eval * n_seg);
    int *tc_error = reinterpret_cast<int *>(max_norm + 1);
    GR_CUDA_CHECK(cudaMemsetAsync(max_norm, 0, 8, s));      // max norm + error flag
    max_row_norm_kernel<<<sm_count() * 4, 256, 0, s>>>(item_emb, ldi, (int)n_items, d, max_norm);
    GR_LAUNCH_CHECK();

    TcArgs a;
    a.user_emb = user_emb; a.ldu = ldu; a.item_emb = item_emb; a.ldi = ldi; a.d = d;
    a.eval_users = eval_users; a.n_eval = (int)n_eval; a.n_items = (int)n_items;
    a.seen_indptr = seen_indptr; a.seen_items = seen_items; a.kprime = kprime;
    a.split_items = (int)split_items;
    a.cand_scores = cand_scores; a.cand_ids = cand_ids; a.cand_cnt = cand_cnt; a.error = tc_error;
    { const char *e = getenv("GR_TC_DEBUG"); a.debug = e ? atoi(e) : 0; }
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (!make_item_tensor_map(&tmap, item_emb, ldi, n_items, d)) return GR_ERR_UNSUPPORTED;
    const size_t smem = tc_smem_bytes(d, kprime);
    GR_CUDA_CHECK(cudaFuncSetAttribute(topk_tc_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_tc_candidates_kernel<<<dim3((unsigned)((n_eval + TC_M - 1) / TC_M), (unsigned)n_seg), TC_THREADS, smem, s>>>(a, tmap);
    GR_LAUNCH_CHECK();

    RescoreArgs r;
    r.user_emb = user_emb; r.ldu = ldu; r.item_emb = item_emb; r.ldi = ldi; r.d = d;
    r.eval_users = eval_users; r.n_eval = (int)n_eval; r.n_items = (int)n_items;
    r.cand_scores = cand_scores; r.cand_ids = cand_ids; r.cand_cnt = cand_cnt; r.kprime = kprime; r.k = k;
    r.n_seg = n_seg;
    // TF32 drops 13 mantissa bits of both operands (relative error < 2^-10 each), so every product is within
    // 2^-9 (1 + 2^-11) of the fp32 one and, by Cauchy-Schwarz, the score within 2^-9 |u| |i|; fp32
    // accumulation adds ~d * 2^-24.  eps = 1.5 * 2^-9 leaves a 50 % margin.
    r.eps = 1.5f / 512.0f;
    r.max_item_norm = max_norm; r.error = tc_error; r.out_ids = topk_ids; r.out_scores = topk_scores; r.flags = flags; r.n_flagged = n_flagged;
    if (n_seg * kprime <= 64) topk_rescore_kernel<2><<<(unsigned)((n_eval + 7) / 8), 256, 0, s>>>(r);
    else topk_rescore_kernel<4><<<(unsigned)((n_eval + 7) / 8), 256, 0, s>>>(r);
    GR_LAUNCH_CHECK();
    return GR_OK;
}
