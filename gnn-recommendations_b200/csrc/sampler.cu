// Host-side BPR batch sampler: Trainer._sample_batch (src/training/trainer.py:146-197) driven by
// the SAME mt19937 stream torch's global CPU generator produces, so sampled indices are
// bit-identical to the reference for a given torch.manual_seed.
//
// The caller passes the generator state exactly as torch.get_rng_state() lays it out
// (at::mt19937: state[624] as uint64 words, `left`, `next`) and writes the advanced state back
// with torch.set_rng_state(); torch.randint(0, n, ...) is `mt() % n` (one 32-bit output) for
// n < 2^28 and `((uint64)mt() << 32 | mt()) % n` (two outputs) for n >= 2^28 — the rule of the
// installed torch (2.11), verified against torch.randint in the tests.
#include <stdint.h>

#include "gr_common.cuh"

namespace {

struct TorchMt {
    uint64_t *state;  // 624 words, low 32 bits used
    int32_t *left;
    uint64_t *next;

    void next_state() {
        constexpr int N = 624, M = 397;
        auto mix = [](uint64_t u, uint64_t v) -> uint32_t {
            const uint32_t y = ((uint32_t)u & 0x80000000u) | ((uint32_t)v & 0x7fffffffu);
            return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
        };
        for (int i = 0; i < N - M; ++i) state[i] = (uint32_t)state[i + M] ^ mix(state[i], state[i + 1]);
        for (int i = N - M; i < N - 1; ++i) state[i] = (uint32_t)state[i + M - N] ^ mix(state[i], state[i + 1]);
        state[N - 1] = (uint32_t)state[M - 1] ^ mix(state[N - 1], state[0]);
        *left = N;
        *next = 0;
    }
    uint32_t operator()() {
        if (--(*left) == 0) next_state();
        uint32_t y = (uint32_t)state[(*next)++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
};

inline int64_t draw_below(TorchMt &mt, int64_t range, int64_t &draws) {
    if (range >= (1LL << 28)) {
        const uint64_t hi = mt(), lo = mt();
        draws += 2;
        return (int64_t)(((hi << 32) | lo) % (uint64_t)range);
    }
    ++draws;
    return (int64_t)(mt() % (uint32_t)range);
}

inline bool contains_sorted(const int32_t *a, int64_t n, int64_t v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t m = (lo + hi) >> 1;
        if (a[m] < v) lo = m + 1; else hi = m;
    }
    return lo < n && a[lo] == v;
}

}  // namespace

// All pointers are HOST pointers.  pos_indptr/pos_items: per-user sorted positive item ids
// (the set trainer.py:169-172 builds).  users/pos/neg: int64[batch] outputs, negative_samples = 1
// (any other value is shape-invalid in the reference's loss).
// Returns the number of 32-bit draws consumed, or a negative error code.
extern "C" int64_t gr_sample_bpr_batch(uint64_t *mt_state_host, int32_t *mt_left_host, uint64_t *mt_next_host,
                                       const int64_t *train_user_host, const int64_t *train_item_host,
                                       int64_t n_train, int64_t n_items, int64_t batch,
                                       const int64_t *pos_indptr_host, const int32_t *pos_items_host,
                                       int64_t *users_host, int64_t *pos_host, int64_t *neg_host) {
    if (!mt_state_host || !mt_left_host || !mt_next_host || !train_user_host || !train_item_host ||
        !pos_indptr_host || !users_host || !pos_host || !neg_host)
        return GR_ERR_INVALID;
    if (n_train <= 0 || n_items <= 0 || batch < 0) return GR_ERR_INVALID;
    TorchMt mt{mt_state_host, mt_left_host, mt_next_host};
    if (batch > n_train) batch = n_train;  // trainer.py:161
    int64_t draws = 0;
    for (int64_t b = 0; b < batch; ++b) {  // indices = torch.randint(0, len(train), (B,))   :162
        const int64_t idx = draw_below(mt, n_train, draws);
        users_host[b] = train_user_host[idx];
        pos_host[b] = train_item_host[idx];
    }
    for (int64_t b = 0; b < batch; ++b) {  // one draw + up to 10 redraws, the last unchecked :181-187
        const int64_t u = users_host[b];
        const int32_t *ps = pos_items_host + pos_indptr_host[u];
        const int64_t np = pos_indptr_host[u + 1] - pos_indptr_host[u];
        int64_t cand = draw_below(mt, n_items, draws);
        for (int t = 0; t < 10; ++t) {
            if (!contains_sorted(ps, np, cand)) break;
            cand = draw_below(mt, n_items, draws);
        }
        neg_host[b] = cand;
    }
    return draws;
}
