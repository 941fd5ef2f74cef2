// host.cpp -- host-side driver pieces of the hot path (include/gsgp_host.h): RNG, constants,
// initial-population trees and the per-generation plan.  Pure C++, no CUDA.
#include "../../include/gsgp_host.h"
#include "program.hpp"
#include "trees.hpp"

#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <new>
#include <charconv>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cmath>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct gsgph_rng {
    // Go's additive lagged Fibonacci source, s[n] = s[n-607] + s[n-273], kept as the last 607 outputs in time order
    // and advanced 607 outputs at a time (dependence-free loops the compiler vectorises) instead of one modular
    // step per draw: same stream, cheaper per draw.
    uint64_t h[607];
    // Raw outputs are read through a queue: [buf, end) holds what the source has produced, cur the next value the
    // logical stream consumes.  While `recording` (gsgph_planner) consumed values stay in the queue, so a caller can
    // note a position (tell) and go back to it (seek): the planner draws a generation before it knows that
    // generation's elite index and, when the guess was wrong, consumes the same raw values again from the first
    // individual whose draws change, in the reference's order.  The logical stream every caller sees is always the
    // source's stream, whatever was re-read.
    uint64_t *buf = nullptr, *cur = nullptr, *end = nullptr;
    size_t cap = 0;
    bool recording = false;
    ~gsgph_rng() { free(buf); }
    void seed_from_go_layout(const uint64_t *vec)               // rngSource's vec with tap = 0, feed = 607 - 273
    {
        for (int i = 0; i < 607; ++i) h[i] = vec[(333 - i + 607) % 607];
        cur = end = buf;
    }
    __attribute__((noinline)) void more()                       // cur == end: the next 607 outputs
    {
        size_t size = recording ? (size_t)(end - buf) : 0;
        if (size + 607 > cap) {
            cap = std::max<size_t>(2 * cap, size + 4 * 607);
            buf = static_cast<uint64_t *>(realloc(buf, cap * sizeof(uint64_t)));
        }
        for (int i = 0; i < 273; ++i) h[i] += h[i + 334];
        for (int i = 273; i < 546; ++i) h[i] += h[i - 273];
        for (int i = 546; i < 607; ++i) h[i] += h[i - 273];
        memcpy(buf + size, h, sizeof(h));
        cur = buf + size;
        end = cur + 607;
    }
    uint64_t next()
    {
        if (__builtin_expect(cur == end, 0)) more();
        return *cur++;
    }
    void tape_begin()                                           // start recording; consumed values are dropped
    {
        const size_t left = (size_t)(end - cur);
        if (left) memmove(buf, cur, left * sizeof(uint64_t));
        cur = buf; end = buf + left;
        recording = true;
    }
    size_t tell() const { return (size_t)(cur - buf); }
    void seek(size_t pos) { cur = buf + pos; }
    int64_t int63() { return (int64_t)(next() & 0x7fffffffffffffffull); }
    int32_t int31() { return (int32_t)(int63() >> 32); }
    double float64()
    {
        for (;;) {
            const double f = (double)int63() / 9223372036854775808.0;
            if (f != 1.0) return f;
        }
    }
    // Int31n (Go math/rand v1): rejection above the largest multiple of n, then v % n.  The driver draws with
    // a handful of distinct n (2, 4, nvar, nconst, population): the threshold and a 64-bit reciprocal for the
    // modulo are cached per n (exact: q = (v * m) >> 63 is floor(v / n) for v < 2^31, m = ceil(2^63 / n)).
    struct IntnCache { int32_t n = 0, mx = 0; uint64_t m = 0; } cache[4];
    int32_t intn(int32_t n)
    {
        if ((n & (n - 1)) == 0) return int31() & (n - 1);
        IntnCache &c = cache[(uint32_t)n & 3u];
        if (c.n != n) {
            c.n = n;
            c.mx = (int32_t)((1u << 31) - 1 - (1u << 31) % (uint32_t)n);
            c.m = (((uint64_t)1 << 63) + (uint64_t)n - 1) / (uint64_t)n;
        }
        return intn_tail(int31(), n, c);
    }
    int32_t intn_tail(int32_t v, int32_t n, const IntnCache &c)
    {
        while (v > c.mx) v = int31();
        const uint32_t q = (uint32_t)(((unsigned __int128)(uint64_t)v * c.m) >> 63);
        return v - (int32_t)(q * (uint32_t)n);
    }
};

namespace {

constexpr int32_t kFunc = 4;

struct TreeBuilder {
    gsgph_rng *r;
    const gsgph_params *p;
    std::vector<int32_t> t;

    int32_t choose_terminal()                                   // main_cpu.go:466-474
    {
        if (p->num_constants == 0) return kFunc + r->intn(p->nvar);
        if (r->float64() < 0.7) return kFunc + r->intn(p->nvar);
        return kFunc + p->nvar + r->intn(p->num_constants);
    }
    // kind 0: create_grow_tree_arrays (:609-639, Intn(2)==0 -> terminal)
    // kind 1: create_grow_tree       (:568-607, Intn(2)==0 -> function) then tree_to_array (:675-693)
    // kind 2: create_full_tree       (:642-672)                          then tree_to_array
    void node(int32_t depth, int32_t max_depth, int kind)
    {
        bool func;
        if (kind == 2) func = (depth == 0 && depth < max_depth) || depth != max_depth;
        else if (depth == 0 && !p->zero_depth) func = true;
        else if (depth == max_depth) func = false;
        else { const int32_t c = r->intn(2); func = (kind == 0) ? (c != 0) : (c == 0); }
        if (!func) { t.push_back(choose_terminal()); return; }
        const int32_t op = r->intn(kFunc);                      // choose_function :458-460
        const size_t pos = t.size();
        t.push_back(op); t.push_back(0); t.push_back(0);
        for (int c = 1; c <= 2; ++c) {
            t[pos + c] = (int32_t)t.size();
            node(depth + 1, max_depth, kind);
        }
    }
    const std::vector<int32_t> &build(int32_t max_depth, int kind) { t.clear(); node(0, max_depth, kind); return t; }
};

// create_grow_tree_arrays (:609-639) straight into `t` (room for 4 * 2^max_depth ints), iterative: pre-order with an
// explicit stack of (slot to patch, depth), so the draws come in the reference's order (decision, then the symbol,
// left subtree before right).  Random trees make "function or terminal?" a coin flip per node; with no constants
// both outcomes take exactly one more draw, so the node is written without branching on the coin.
inline int32_t grow_tree_array(gsgph_rng *r, const gsgph_params *p, int32_t *t)
{
    const int32_t max_depth = p->max_depth_creation, nvar = std::max(p->nvar, 1);   // (no variables: x0, as Intn(0) would panic)
    const bool plain = p->num_constants == 0;
    const int32_t mx = (int32_t)((1u << 31) - 1 - (1u << 31) % (uint32_t)nvar);
    const uint64_t m = (((uint64_t)1 << 63) + (uint64_t)nvar - 1) / (uint64_t)nvar;
    int32_t patch[64];
    int32_t dep[64];
    int32_t n = 0, sp = 1;
    patch[0] = -1; dep[0] = 0;
    int32_t scratch = 0;
    while (sp > 0) {
        --sp;
        const int32_t slot = patch[sp], d = dep[sp];
        *(slot >= 0 ? t + slot : &scratch) = n;                              // the parent's child index (:617,:620 / :632,:635)
        bool func;
        if (d == 0 && !p->zero_depth) func = true;                          // :610
        else if (d == max_depth) func = false;                              // :623
        else func = r->intn(2) != 0;                                        // :626-638
        if (plain) {                                                        // choose_function :458-460 / choose_terminal :466-468
            int32_t v = r->int31();
            if (__builtin_expect(!func && v > mx, 0)) { do v = r->int31(); while (v > mx); }   // Int31n's rejection (rare)
            const int32_t q = (int32_t)(((unsigned __int128)(uint64_t)v * m) >> 63);         // v / nvar (see IntnCache)
            t[n] = func ? (v & (kFunc - 1)) : kFunc + (v - q * nvar);
        } else if (func) t[n] = r->intn(kFunc);
        else if (r->float64() < 0.7) t[n] = kFunc + r->intn(nvar);          // :469-473
        else t[n] = kFunc + nvar + r->intn(p->num_constants);
        patch[sp] = n + 2; dep[sp] = d + 1;                                 // right child: popped after the whole left subtree
        patch[sp + 1] = n + 1; dep[sp + 1] = d + 1;
        sp += func ? 2 : 0;
        n += func ? 3 : 1;
    }
    return n;
}

inline bool better(const gsgph_params *p, double a, double b) { return p->minimization_problem ? a < b : a > b; }

int32_t tournament(gsgph_rng *r, const gsgph_params *p, const double *fit)      // :830-840
{
    int32_t best = r->intn(p->population_size);
    for (int i = 1; i < p->tournament_size; ++i) {
        const int32_t next = r->intn(p->population_size);
        if (better(p, fit[next], fit[best])) best = next;
    }
    return best;
}


// Where the interpreter programs of a plan's trees go when the host stages them itself (gsgp_generation_compiled)
struct ProgSink {
    gsgp::Ins *ins = nullptr;
    int32_t cap = 0, used = 0;
    int32_t *off = nullptr, *desc = nullptr;                    // per tree slot 2*k + which
    int32_t n_symbols = 0;
    std::vector<uint8_t> need;
    std::vector<gsgp::TreeFrame> frames;
    static int32_t n_ins_of(int32_t tree_len) { return tree_len <= 0 ? 0 : tree_len == 1 ? 1 : (tree_len - 1) / 4; }
    bool compile(int32_t slot, const int32_t *tree, int32_t len, int32_t at)      // false: not a tree
    {
        int n = 0, nd = 0;
        const char *err = "";
        if (!gsgp::compile_prefix_tree(tree, len, n_symbols, ins + at, n, nd, need, frames, err) || n != n_ins_of(len)) return false;
        off[slot] = at; desc[slot] = gsgp::prog_desc(n, nd);
        return true;
    }
};

// One generation's decisions (:1447-1470), individual by individual in the reference's draw order.
struct PlanWriter {
    gsgph_rng *r;
    const gsgph_params *p;
    const double *fit;
    int32_t index_best;
    gsgp_plan_entry *plan;
    int32_t *pool;
    int32_t cap, used = 0;
    bool overflow = false;                                      // the tree pool is too small
    TreeBuilder tb;
    int64_t tree_max;
    ProgSink *sink = nullptr;                                   // optional: compile each tree as it is drawn
    PlanWriter(gsgph_rng *r_, const gsgph_params *p_, const double *fit_, int32_t best, gsgp_plan_entry *plan_,
               int32_t *pool_, int32_t cap_, int32_t used_)
        : r(r_), p(p_), fit(fit_), index_best(best), plan(plan_), pool(pool_), cap(cap_), used(used_), tb{r_, p_, {}},
          tree_max(p_->max_depth_creation >= 0 && p_->max_depth_creation < 20 ? ((int64_t)4 << p_->max_depth_creation) : INT32_MAX) {}
    void new_tree(int32_t &off, int32_t &len)
    {
        if (used + tree_max <= cap) {                           // room for any tree: built in place
            off = used; len = grow_tree_array(r, p, pool + used); used += len;
            return;
        }
        const std::vector<int32_t> &t = tb.build(p->max_depth_creation, 0);
        if (used + (int64_t)t.size() > cap) { overflow = true; off = -1; len = 0; return; }
        memcpy(pool + used, t.data(), t.size() * sizeof(int32_t));
        off = used; len = (int32_t)t.size(); used += len;
    }
    void compile_trees_of(int32_t k)
    {
        const gsgp_plan_entry &e = plan[k];
        const int32_t off[2] = {e.tree0_off, e.tree1_off}, len[2] = {e.tree0_len, e.tree1_len};
        for (int w = 0; w < 2; ++w) {
            sink->off[2 * k + w] = 0; sink->desc[2 * k + w] = 0;
            if (len[w] <= 0 || overflow) continue;
            if (sink->used + ProgSink::n_ins_of(len[w]) > sink->cap) { overflow = true; continue; }
            sink->compile(2 * k + w, pool + off[w], len[w], sink->used);
            sink->used += ProgSink::n_ins_of(len[w]);
        }
    }
    void one(int32_t k)
    {
        draw_one(k);
        if (sink) compile_trees_of(k);
    }
    void draw_one(int32_t k)
    {
        gsgp_plan_entry &e = plan[k];
        e.op = GSGP_OP_REPR; e.elite = (k == index_best); e.p1 = -1; e.p2 = -1;
        e.tree0_off = -1; e.tree0_len = 0; e.tree1_off = -1; e.tree1_len = 0; e.ms = 0.0;
        const double rand_num = r->float64();                                   // :1451
        if (rand_num < p->p_crossover) {                                        // :1453
            e.op = GSGP_OP_XO;
            if (k != index_best) {
                new_tree(e.tree0_off, e.tree0_len);                             // :879
                e.p1 = tournament(r, p, fit);                                   // :881
                e.p2 = tournament(r, p, fit);                                   // :882
            } else e.p1 = k;
        } else if (rand_num < p->p_crossover + p->p_mutation) {                 // :1459
            e.op = GSGP_OP_MUT;
            e.p1 = (k != index_best) ? tournament(r, p, fit) : k;               // :847-850
            if (k != index_best) {
                e.ms = r->float64();                                            // :919
                new_tree(e.tree0_off, e.tree0_len);                             // :921
                new_tree(e.tree1_off, e.tree1_len);                             // :922
            }
        } else {
            e.op = GSGP_OP_REPR;                                                // :1467-1469
            e.p1 = (k != index_best) ? tournament(r, p, fit) : k;
        }
    }
    void range(int32_t k_from)
    {
        for (int32_t k = k_from; k < p->population_size && !overflow; ++k) one(k);
    }
};

}  // namespace

extern "C" {

gsgph_rng *gsgph_rng_new(int64_t seed)                         // rngSource.Seed: Go's own stream for this seed
{
    static const uint64_t cooked[607] = {
#include "go_rng_cooked.inc.h"
    };
    auto seedrand = [](int32_t x) {                             // x[n+1] = 48271 * x[n] mod (2^31 - 1)
        const int32_t hi = x / 44488, lo = x % 44488;
        x = 48271 * lo - 3399 * hi;
        return x < 0 ? x + 2147483647 : x;
    };
    gsgph_rng *r = new gsgph_rng();
    uint64_t vec[607];
    seed %= 2147483647;
    if (seed < 0) seed += 2147483647;
    if (seed == 0) seed = 89482311;
    int32_t x = (int32_t)seed;
    for (int i = -20; i < 607; ++i) {
        x = seedrand(x);
        if (i < 0) continue;
        uint64_t u = (uint64_t)x << 40;
        x = seedrand(x); u ^= (uint64_t)x << 20;
        x = seedrand(x); u ^= (uint64_t)x;
        vec[i] = u ^ cooked[i];
    }
    r->seed_from_go_layout(vec);
    return r;
}
void gsgph_rng_free(gsgph_rng *r) { delete r; }
double gsgph_rng_float64(gsgph_rng *r) { return r->float64(); }
int32_t gsgph_rng_intn(gsgph_rng *r, int32_t n) { return n > 0 ? r->intn(n) : -1; }

int gsgph_create_T_F(gsgph_rng *r, const gsgph_params *p, double *sym_val)
{
    const int32_t n = kFunc + p->nvar + p->num_constants;
    for (int32_t i = 0; i < kFunc + p->nvar; ++i) sym_val[i] = 0.0;
    for (int32_t i = kFunc + p->nvar; i < n; ++i)               // :450-454
        sym_val[i] = p->min_random_constant + r->float64() * (p->max_random_constant - p->min_random_constant);
    return 0;
}

int gsgph_initialize_population(gsgph_rng *r, const gsgph_params *p, int32_t n_seeds, int32_t *pool,
                                int32_t cap, int32_t *off, int32_t *len, int32_t *pool_len)
{
    TreeBuilder tb{r, p, {}};
    const int32_t P = p->population_size, d = p->max_depth_creation;
    int32_t made = 0, used = 0;
    bool overflow = false;
    auto add = [&](int32_t max_depth, int kind) {
        const std::vector<int32_t> &t = tb.build(max_depth, kind);
        if (used + (int64_t)t.size() > cap) { overflow = true; return; }
        memcpy(pool + used, t.data(), t.size() * sizeof(int32_t));
        off[made] = used; len[made] = (int32_t)t.size();
        used += (int32_t)t.size(); ++made;
    };
    const int32_t todo = P - n_seeds;
    if (p->init_type == 0) { while (made < todo && !overflow) add(d, 1); }            // create_grow_pop :477-483
    else if (p->init_type == 1) { while (made < todo && !overflow) add(d, 2); }       // create_full_pop :486-492
    else {                                                                             // create_ramped_pop :495-538
        int32_t sub_pop, rem, min_depth;
        if (!p->zero_depth) { sub_pop = todo / d; rem = todo % d; min_depth = 1; }
        else { sub_pop = todo / (d + 1); rem = todo % (d + 1); min_depth = 0; }
        for (int32_t j = d; j >= min_depth && !overflow; --j) {
            const int32_t n = (j < d) ? sub_pop : sub_pop + rem;
            const int32_t n_full = (int32_t)std::ceil(n * 0.5), n_grow = (int32_t)std::floor(n * 0.5);
            for (int32_t k = 0; k < n_full && !overflow; ++k) add(j, 2);
            for (int32_t k = 0; k < n_grow && !overflow; ++k) add(j, 1);
        }
    }
    if (pool_len) *pool_len = used;
    return overflow ? 1 : (made == todo ? 0 : 2);
}

int gsgph_build_plan(gsgph_rng *r, const gsgph_params *p, const double *fit, int32_t index_best,
                     gsgp_plan_entry *plan, int32_t *pool, int32_t cap, int32_t *pool_len)
{
    PlanWriter w(r, p, fit, index_best, plan, pool, cap, 0);
    w.range(0);
    if (pool_len) *pool_len = w.used;
    return w.overflow ? 1 : 0;
}

int gsgph_build_effects_plan(gsgph_rng *r, const gsgph_params *p, int32_t test_crossover,
                             gsgp_plan_entry *plan, int32_t *pool, int32_t cap, int32_t *pool_len)
{
    TreeBuilder tb{r, p, {}};
    int32_t used = 0;
    bool overflow = false;
    auto new_tree = [&](int32_t &off, int32_t &len) {
        const std::vector<int32_t> &t = tb.build(p->max_depth_creation, 0);
        if (used + (int64_t)t.size() > cap) { overflow = true; off = -1; len = 0; return; }
        memcpy(pool + used, t.data(), t.size() * sizeof(int32_t));
        off = used; len = (int32_t)t.size(); used += len;
    };
    const int32_t index_best = 0;                                               // operators_effects.go:197 (never assigned)
    for (int32_t k = 0; k < p->population_size && !overflow; ++k) {
        gsgp_plan_entry &e = plan[k];
        e.op = test_crossover ? GSGP_OP_XO : GSGP_OP_MUT; e.elite = (k == index_best); e.p1 = k; e.p2 = -1;
        e.tree0_off = -1; e.tree0_len = 0; e.tree1_off = -1; e.tree1_len = 0; e.ms = 0.0;
        if (e.elite) continue;
        if (test_crossover) {
            new_tree(e.tree0_off, e.tree0_len);                                 // :803
            e.p1 = r->intn(p->population_size);                                 // :805 random_selection
            e.p2 = r->intn(p->population_size);                                 // :806
        } else {
            e.ms = r->float64();                                                // :838 (after reproduction(k): self copy)
            new_tree(e.tree0_off, e.tree0_len);                                 // :840
            new_tree(e.tree1_off, e.tree1_len);                                 // :841
        }
    }
    if (pool_len) *pool_len = used;
    return overflow ? 1 : 0;
}

int32_t gsgph_best_individual(const gsgph_params *p, const double *fit)
{
    int32_t best = 0;
    for (int32_t i = 1; i < p->population_size; ++i) if (better(p, fit[i], fit[best])) best = i;
    return best;
}

// ---- pipelined planner ------------------------------------------------------------------------------------------
// gsgph_build_plan is serial with the device: generation g+1's plan needs generation g's fitness.  But only the
// tournament WINNERS need fitness values; the operator draw, the random trees, the mutation step and the tournament
// CANDIDATES need nothing but the RNG -- and the elite index (the elite slot skips its draws, :847,:877,:918).  So the
// planner draws generation g+1 on a worker thread while the device computes generation g, under a guess of the elite
// index, and finishes the plan when the fitness arrives: winners among the recorded candidates and, when the guess
// was wrong, a repair.  The repair re-reads the recorded raw RNG values (gsgph_rng's queue) from the first slot whose
// draws change.  Past that slot the individuals parse a SHIFTED stream -- until an individual happens to start at a
// raw position where some individual of the draft started: from there on the draft's parse is the right one again
// (same values, same order) and its entries are copied, shifted by the index difference, up to the next slot that
// differs (the other elite index).  A wrong guess therefore costs a few dozen re-drawn individuals, not the suffix.
// Plans, tree pools and the RNG stream are exactly those of gsgph_build_plan.
struct gsgph_planner {
    gsgph_rng *r;
    gsgph_params p;
    struct Rec { size_t pos; int32_t pool_used; int32_t ncand; };           // per draft individual: where its draws / trees start
    std::vector<Rec> rec;                                                   // P + 1: the last entry is the draft's end
    std::vector<int32_t> cand;                                              // [P][2][tournament_size]
    struct Buf {
        std::vector<gsgp_plan_entry> plan; std::vector<int32_t> pool;
        std::vector<gsgp::Ins> prog; std::vector<int32_t> prog_off, prog_desc; int32_t n_ins = 0;    // with programs enabled
    } buf[3];
    size_t pool_max = 0;
    // programs (gsgph_planner_enable_programs): the draft's trees are compiled on the worker + helper threads
    bool programs = false;
    int32_t n_symbols = 0;
    std::vector<int32_t> prog_start;                                        // P + 1: instructions before draft individual k
    std::vector<std::thread> helpers;
    std::vector<ProgSink> sinks;                                            // per compile thread (scratch) + [n] for the caller's thread
    uint64_t job_epoch = 0;
    int job_pending = 0;
    bool job_quit = false;
    std::mutex jmu;
    std::condition_variable jcv, jdone;
    int out = 0, draft = 1;                                                 // buffers: last result (the caller reads it), draft
    int32_t guess = -1, used = 0;
    bool pending = false, overflow = false;
    int64_t hits = 0, misses = 0, redrawn = 0;
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    enum { IDLE, REQUESTED, QUIT } state = IDLE;

    void draw_candidates(int32_t *c)
    {
        for (int i = 0; i < p.tournament_size; ++i) c[i] = r->intn(p.population_size);     // :831,:834
    }
    int32_t winner(const int32_t *c, const double *fit) const                              // :830-840
    {
        int32_t best = c[0];
        for (int i = 1; i < p.tournament_size; ++i)
            if (better(&p, fit[c[i]], fit[best])) best = c[i];
        return best;
    }
    void size_pool(int b, size_t n)
    {
        buf[b].pool.resize(n);
        if (programs) buf[b].prog.resize(n / 4 + 2 * (size_t)p.population_size + 16);   // one instruction per function node
    }
    bool grow_pool(int b)                                                   // false: already at the worst-case size
    {
        if (buf[b].pool.size() >= pool_max) return false;
        size_pool(b, std::min(pool_max, buf[b].pool.size() * 2));
        return true;
    }
    ProgSink *sink_for(int b, ProgSink &s)                                  // the caller thread's sink onto buffer b
    {
        if (!programs) return nullptr;
        s.ins = buf[b].prog.data(); s.cap = (int32_t)buf[b].prog.size(); s.used = 0;
        s.off = buf[b].prog_off.data(); s.desc = buf[b].prog_desc.data(); s.n_symbols = n_symbols;
        return &s;
    }
    // the draft's trees -> programs, in place at offsets known from the tree lengths; individuals split over the threads
    void compile_part(int part, int parts)
    {
        const int32_t P = p.population_size;
        Buf &B = buf[draft];
        ProgSink &s = sinks[(size_t)part];
        s.ins = B.prog.data(); s.off = B.prog_off.data(); s.desc = B.prog_desc.data(); s.n_symbols = n_symbols;
        const int32_t k0 = (int32_t)((int64_t)P * part / parts), k1 = (int32_t)((int64_t)P * (part + 1) / parts);
        for (int32_t k = k0; k < k1; ++k) {
            const gsgp_plan_entry &e = B.plan[(size_t)k];
            int32_t at = prog_start[(size_t)k];
            if (e.tree0_len > 0) { s.compile(2 * k, B.pool.data() + e.tree0_off, e.tree0_len, at); at += ProgSink::n_ins_of(e.tree0_len); }
            if (e.tree1_len > 0) s.compile(2 * k + 1, B.pool.data() + e.tree1_off, e.tree1_len, at);
        }
    }
    void compile_draft()
    {
        const int32_t P = p.population_size;
        Buf &B = buf[draft];
        int32_t acc = 0;
        for (int32_t k = 0; k < P; ++k) {
            prog_start[(size_t)k] = acc;
            B.prog_off[2 * (size_t)k] = B.prog_off[2 * (size_t)k + 1] = 0;
            B.prog_desc[2 * (size_t)k] = B.prog_desc[2 * (size_t)k + 1] = 0;
            acc += ProgSink::n_ins_of(B.plan[(size_t)k].tree0_len) + ProgSink::n_ins_of(B.plan[(size_t)k].tree1_len);
        }
        prog_start[(size_t)P] = acc;
        B.n_ins = acc;
        const int parts = (int)helpers.size() + 1;
        if (parts > 1 && P >= 256) {
            { std::lock_guard<std::mutex> lk(jmu); job_pending = parts - 1; ++job_epoch; }
            jcv.notify_all();
            compile_part(0, parts);
            std::unique_lock<std::mutex> lk(jmu);
            jdone.wait(lk, [&] { return job_pending == 0; });
        } else compile_part(0, 1);
    }
    void helper_loop(int part)
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(jmu);
                jcv.wait(lk, [&] { return job_epoch != seen || job_quit; });
                if (job_quit) return;
                seen = job_epoch;
            }
            compile_part(part, (int)helpers.size() + 1);
            { std::lock_guard<std::mutex> lk(jmu); if (--job_pending == 0) jdone.notify_one(); }
        }
    }
    // everything of one generation that does not need fitness values, into the draft buffers
    void draw()
    {
        r->tape_begin();
        for (;;) {
            draw_once();
            if (!overflow || !grow_pool(draft)) break;
            r->seek(0);                                                     // same raw values, larger pool
        }
        if (programs && !overflow) compile_draft();
    }
    void draw_once()
    {
        const int32_t P = p.population_size, ts = p.tournament_size;
        PlanWriter w(r, &p, nullptr, guess, buf[draft].plan.data(), buf[draft].pool.data(), (int32_t)buf[draft].pool.size(), 0);
        for (int32_t k = 0; k < P && !w.overflow; ++k) {
            gsgp_plan_entry &e = w.plan[k];
            int32_t *c = cand.data() + (size_t)k * 2 * ts;
            rec[(size_t)k] = {r->tell(), w.used, 0};
            e.op = GSGP_OP_REPR; e.elite = (k == guess); e.p1 = k; e.p2 = -1;
            e.tree0_off = -1; e.tree0_len = 0; e.tree1_off = -1; e.tree1_len = 0; e.ms = 0.0;
            const double rand_num = r->float64();                                   // :1451
            if (rand_num < p.p_crossover) {                                         // :1453
                e.op = GSGP_OP_XO;
                if (e.elite) continue;
                w.new_tree(e.tree0_off, e.tree0_len);                               // :879
                draw_candidates(c); draw_candidates(c + ts);                        // :881-882
                rec[(size_t)k].ncand = 2;
            } else if (rand_num < p.p_crossover + p.p_mutation) {                   // :1459
                e.op = GSGP_OP_MUT;
                if (e.elite) continue;
                draw_candidates(c);                                                 // :847-850
                rec[(size_t)k].ncand = 1;
                e.ms = r->float64();                                                // :919
                w.new_tree(e.tree0_off, e.tree0_len);                               // :921
                w.new_tree(e.tree1_off, e.tree1_len);                               // :922
            } else {
                if (e.elite) continue;                                              // :1467-1469
                draw_candidates(c);
                rec[(size_t)k].ncand = 1;
            }
        }
        rec[(size_t)P] = {r->tell(), w.used, 0};
        used = w.used; overflow = w.overflow;
    }
    void resolve(gsgp_plan_entry &e, int32_t k_draft, const double *fit) const
    {
        const int32_t *c = cand.data() + (size_t)k_draft * 2 * p.tournament_size;
        const int32_t n = rec[(size_t)k_draft].ncand;
        if (n >= 1) e.p1 = winner(c, fit);
        if (n == 2) e.p2 = winner(c + p.tournament_size, fit);
    }
    // the draft was drawn for elite index `guess`, the elite is `best`: the plan for `best` into buffer b.
    // Returns true when b's tree pool is too small.
    bool repair(int b, const double *fit, int32_t best)
    {
        const int32_t P = p.population_size;
        const gsgp_plan_entry *dplan = buf[draft].plan.data();
        const int32_t *dpool = buf[draft].pool.data();
        const int32_t m = std::max(0, std::min(best, guess));               // slots before m: draws unchanged
        PlanWriter w(r, &p, fit, best, buf[b].plan.data(), buf[b].pool.data(), (int32_t)buf[b].pool.size(), 0);
        ProgSink &sk = sinks.back();
        w.sink = sink_for(b, sk);
        auto copy_run = [&](int32_t k, int32_t kd, int32_t n) {             // draft slots [kd, kd+n) become slots [k, k+n)
            const int32_t u0 = rec[(size_t)kd].pool_used, u1 = rec[(size_t)(kd + n)].pool_used, shift = w.used - u0;
            if ((int64_t)w.used + (u1 - u0) > w.cap) { w.overflow = true; return; }
            memcpy(w.pool + w.used, dpool + u0, sizeof(int32_t) * (size_t)(u1 - u0));
            memcpy(w.plan + k, dplan + kd, sizeof(gsgp_plan_entry) * (size_t)n);
            for (int32_t i = 0; i < n; ++i) {
                gsgp_plan_entry &e = w.plan[k + i];
                if (e.tree0_len) e.tree0_off += shift;
                if (e.tree1_len) e.tree1_off += shift;
                resolve(e, kd + i, fit);
            }
            w.used += u1 - u0;
            if (w.sink) {                                                   // the run's programs, shifted the same way
                const int32_t q0 = prog_start[(size_t)kd], q1 = prog_start[(size_t)(kd + n)], qshift = sk.used - q0;
                if (sk.used + (q1 - q0) > sk.cap) { w.overflow = true; return; }
                memcpy(sk.ins + sk.used, buf[draft].prog.data() + q0, sizeof(gsgp::Ins) * (size_t)(q1 - q0));
                const int32_t *doff = buf[draft].prog_off.data(), *ddesc = buf[draft].prog_desc.data();
                for (int32_t t = 0; t < 2 * n; ++t) {
                    const int32_t d = ddesc[2 * kd + t];
                    sk.desc[2 * k + t] = d;
                    sk.off[2 * k + t] = d ? doff[2 * kd + t] + qshift : 0;
                }
                sk.used += q1 - q0;
            }
        };
        copy_run(0, 0, m);
        int32_t k = m, kd = m;
        r->seek(rec[(size_t)m].pos);
        while (k < P && !w.overflow) {
            const size_t pos = r->tell();
            while (kd < P && rec[(size_t)kd].pos < pos) ++kd;
            // a run of draft individuals that parse identically here: it starts at this raw position, holds neither
            // the guessed elite (its draft skipped the draws) nor -- at its new index -- the actual one
            int32_t n = 0;
            if (kd < P && rec[(size_t)kd].pos == pos) {
                n = std::min(P - k, P - kd);
                if (guess >= kd) n = std::min(n, guess - kd);
                if (best >= k) n = std::min(n, best - k);
            }
            if (n > 0) {
                copy_run(k, kd, n);
                k += n; kd += n;
                r->seek(rec[(size_t)kd].pos);
            } else {
                w.one(k++);
                ++redrawn;
            }
        }
        used = w.used;
        if (w.sink) buf[b].n_ins = sk.used;
        return w.overflow;
    }
    void loop()
    {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return state != IDLE; });
            if (state == QUIT) return;
            lk.unlock();
            draw();
            lk.lock();
            state = IDLE;
            cv.notify_all();
        }
    }
};

gsgph_planner *gsgph_planner_new(gsgph_rng *r, const gsgph_params *p)
{
    if (!r || !p || p->population_size <= 0 || p->tournament_size <= 0 || p->nvar <= 0 || p->num_constants < 0 ||
        p->max_depth_creation < 0 || p->max_depth_creation > 20) return nullptr;
    gsgph_planner *pl = new gsgph_planner();
    pl->r = r; pl->p = *p;
    const size_t P = (size_t)p->population_size;
    const size_t tree_cap = ((size_t)4 << p->max_depth_creation);           // a depth-d tree has at most 4*2^d - 3 ints
    pl->rec.resize(P + 1);
    pl->cand.resize(P * 2 * (size_t)p->tournament_size);
    pl->pool_max = std::min<size_t>(2 * P * tree_cap, (size_t)INT32_MAX);
    for (auto &b : pl->buf) { b.plan.resize(P); b.pool.resize(std::min<size_t>(pl->pool_max, 64 * P + 2 * tree_cap)); }
    pl->sinks.resize(1);
    pl->worker = std::thread([pl] { pl->loop(); });
    return pl;
}

int gsgph_planner_enable_programs(gsgph_planner *pl, int32_t n_threads)
{
    if (!pl) return 1;
    if (pl->programs) return 0;
    if (pl->pending) return 2;                                              // before the first prefetch
    const size_t P = (size_t)pl->p.population_size;
    pl->programs = true;
    pl->n_symbols = kFunc + pl->p.nvar + pl->p.num_constants;
    pl->prog_start.resize(P + 1);
    for (int b = 0; b < 3; ++b) {
        pl->buf[b].prog_off.assign(2 * P, 0); pl->buf[b].prog_desc.assign(2 * P, 0);
        pl->size_pool(b, pl->buf[b].pool.size());
    }
    if (n_threads <= 0) n_threads = std::max(1, std::min(4, (int)std::thread::hardware_concurrency() / 4));
    n_threads = std::min(n_threads, 16);
    pl->sinks.resize((size_t)n_threads + 1);
    for (int i = 1; i < n_threads; ++i) pl->helpers.emplace_back([pl, i] { pl->helper_loop(i); });
    return 0;
}

int gsgph_planner_programs(const gsgph_planner *pl, const uint32_t **programs, int32_t *n_instructions,
                           const int32_t **prog_off, const int32_t **prog_desc)
{
    if (!pl || !pl->programs) return 1;
    const gsgph_planner::Buf &B = pl->buf[pl->out];
    if (programs) *programs = reinterpret_cast<const uint32_t *>(B.prog.data());
    if (n_instructions) *n_instructions = B.n_ins;
    if (prog_off) *prog_off = B.prog_off.data();
    if (prog_desc) *prog_desc = B.prog_desc.data();
    return 0;
}

static void planner_wait(gsgph_planner *pl)
{
    std::unique_lock<std::mutex> lk(pl->mu);
    pl->cv.wait(lk, [&] { return pl->state == gsgph_planner::IDLE; });
}

void gsgph_planner_free(gsgph_planner *pl)
{
    if (!pl) return;
    planner_wait(pl);
    {
        std::lock_guard<std::mutex> lk(pl->mu);
        pl->state = gsgph_planner::QUIT;
    }
    pl->cv.notify_all();
    pl->worker.join();
    {
        std::lock_guard<std::mutex> lk(pl->jmu);
        pl->job_quit = true;
    }
    pl->jcv.notify_all();
    for (auto &t : pl->helpers) t.join();
    // a prefetched generation that was never asked for: its draws go back to the stream
    if (pl->pending) pl->r->seek(pl->rec[0].pos);
    pl->r->recording = false;
    delete pl;
}

int gsgph_planner_prefetch(gsgph_planner *pl, int32_t index_best_guess)
{
    if (!pl) return 1;
    planner_wait(pl);
    if (pl->pending) return 2;                                              // one generation ahead, not more
    pl->guess = index_best_guess;
    pl->draft = (pl->out + 1) % 3;
    pl->pending = true;
    {
        std::lock_guard<std::mutex> lk(pl->mu);
        pl->state = gsgph_planner::REQUESTED;
    }
    pl->cv.notify_all();
    return 0;
}

int gsgph_planner_plan(gsgph_planner *pl, const double *fit, int32_t index_best, const gsgp_plan_entry **plan_out,
                       const int32_t **pool_out, int32_t *pool_len)
{
    if (!pl || !fit) return 1;
    planner_wait(pl);
    const int32_t P = pl->p.population_size;
    bool overflow = false;
    int res;
    if (!pl->pending) {                                                     // nothing prefetched: the plain serial build
        res = (pl->out + 1) % 3;
        pl->r->tape_begin();
        for (;;) {
            pl->r->seek(0);
            PlanWriter w(pl->r, &pl->p, fit, index_best, pl->buf[res].plan.data(), pl->buf[res].pool.data(), (int32_t)pl->buf[res].pool.size(), 0);
            w.sink = pl->sink_for(res, pl->sinks.back());
            w.range(0);
            pl->used = w.used; overflow = w.overflow;
            if (w.sink) pl->buf[res].n_ins = w.sink->used;
            if (!overflow || !pl->grow_pool(res)) break;
        }
    } else if (pl->overflow) {                                              // only at the worst-case pool size
        pl->pending = false;
        res = pl->draft; overflow = true;
    } else if (index_best == pl->guess) {                                   // right guess: winners among the drawn candidates
        pl->pending = false;
        res = pl->draft;
        gsgp_plan_entry *plan = pl->buf[res].plan.data();
        for (int32_t k = 0; k < P; ++k) pl->resolve(plan[k], k, fit);
        ++pl->hits;
    } else {                                                                // wrong guess: repaired into the third buffer
        pl->pending = false;
        res = 3 - pl->out - pl->draft;
        const int32_t draft_used = pl->used;
        for (;;) {
            overflow = pl->repair(res, fit, index_best);
            if (!overflow || !pl->grow_pool(res)) break;
            pl->used = draft_used;
        }
        ++pl->misses;
    }
    pl->r->recording = false;
    pl->out = res;
    if (plan_out) *plan_out = pl->buf[res].plan.data();
    if (pool_out) *pool_out = pl->buf[res].pool.data();
    if (pool_len) *pool_len = pl->used;
    return overflow ? 3 : 0;
}

int gsgph_compile_plan(const gsgp_plan_entry *plan, int32_t population_size, const int32_t *tree_pool, int32_t pool_len,
                       int32_t n_symbols, uint32_t *programs, int32_t cap, int32_t *n_instructions, int32_t *prog_off,
                       int32_t *prog_desc)
{
    if (!plan || !programs || !prog_off || !prog_desc || population_size < 0) return 1;
    ProgSink s;
    s.ins = reinterpret_cast<gsgp::Ins *>(programs); s.cap = cap; s.off = prog_off; s.desc = prog_desc; s.n_symbols = n_symbols;
    for (int32_t k = 0; k < population_size; ++k) {
        const gsgp_plan_entry &e = plan[k];
        const bool computed = !e.elite && (e.op == GSGP_OP_XO || e.op == GSGP_OP_MUT);
        const int32_t off[2] = {e.tree0_off, e.tree1_off};
        const int32_t len[2] = {computed ? e.tree0_len : 0, computed && e.op == GSGP_OP_MUT ? e.tree1_len : 0};
        for (int w = 0; w < 2; ++w) {
            prog_off[2 * k + w] = 0; prog_desc[2 * k + w] = 0;
            if (len[w] <= 0) continue;
            if (!tree_pool || off[w] < 0 || (int64_t)off[w] + len[w] > pool_len) return 2;
            if ((int64_t)s.used + len[w] / 4 + 1 > cap) return 3;
            int n = 0, nd = 0;
            const char *err = "";
            if (!gsgp::compile_prefix_tree(tree_pool + off[w], len[w], n_symbols, s.ins + s.used, n, nd, s.need, s.frames, err)) return 2;
            prog_off[2 * k + w] = s.used; prog_desc[2 * k + w] = gsgp::prog_desc(n, nd);
            s.used += n;
        }
    }
    if (n_instructions) *n_instructions = s.used;
    return 0;
}

// gsgp_b200.cu: gsgp_generation_compiled without the per-instruction check, for programs this library compiled itself
extern "C" int gsgp_internal_generation_trusted(gsgp_ctx *c, const gsgp_plan_entry *plan, const int32_t *tree_pool, int32_t pool_len,
                                                const uint32_t *programs, int32_t n_instructions, const int32_t *prog_off,
                                                const int32_t *prog_desc, double *fit_out, double *fit_test_out, int32_t *index_best_out);

int gsgph_replay_generations(gsgph_planner *pl, gsgp_ctx *ctx, int32_t n_generations, double *fit, double *fit_test,
                             int32_t *index_best, double *best_fit_history, double *best_fit_test_history)
{
    if (!pl || !ctx || !fit || !fit_test || !index_best || n_generations < 0) return 1;
    static const bool trace = getenv("GSGP_TRACE") != nullptr;
    double us_plan = 0, us_prefetch = 0, us_call = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    for (int32_t g = 0; g < n_generations; ++g) {
        const gsgp_plan_entry *plan; const int32_t *pool; int32_t pool_len;
        const auto t0 = now();
        if (int rc = gsgph_planner_plan(pl, fit, *index_best, &plan, &pool, &pool_len)) return rc;
        const auto t1 = now();
        gsgph_planner_prefetch(pl, *index_best);                                // the next generation's draws start now
        const auto t2 = now();
        int rc;
        if (pl->programs) {
            const gsgph_planner::Buf &B = pl->buf[pl->out];
            // the planner's own programs: no per-instruction re-validation (gsgp_b200.cu)
            rc = gsgp_internal_generation_trusted(ctx, plan, pool, pool_len, reinterpret_cast<const uint32_t *>(B.prog.data()), B.n_ins,
                                                  B.prog_off.data(), B.prog_desc.data(), fit, fit_test, index_best);
        } else rc = gsgp_generation(ctx, plan, pool, pool_len, fit, fit_test, index_best);
        if (rc) return 4;                                                       // gsgp_last_error(ctx) says why
        if (best_fit_history) best_fit_history[g] = fit[*index_best];           // :1495-1496
        if (best_fit_test_history) best_fit_test_history[g] = fit_test[*index_best];
        us_plan += us(t0, t1); us_prefetch += us(t1, t2); us_call += us(t2, now());
    }
    if (trace && n_generations > 0)
        fprintf(stderr, "gsgph_replay_generations: per generation plan %.1f us, prefetch %.1f us, generation call %.1f us\n",
                us_plan / n_generations, us_prefetch / n_generations, us_call / n_generations);
    return 0;
}

// ---- one plan for every rank of a node ("host broadcasts the same plan to each GPU") ------------------------
// The plan of a generation is identical on every rank, so one process (the writer) builds it and the others (readers)
// take it from a POSIX shared-memory segment instead of each running a planner and its compile threads.  Two slots:
// sequence s lives in slot s & 1; the writer reuses a slot once every reader has released sequence s - 2.
}  // extern "C"  (the struct has a member template)
struct gsgph_channel {
    static constexpr int kMaxReaders = 64;
    struct Header {
        uint64_t magic;
        int32_t P, cap_pool, cap_ins, n_readers;
        uint64_t slot_bytes, total_bytes;
        std::atomic<uint64_t> ready;                     // 1 once the writer has initialised the segment
        std::atomic<uint64_t> published[2];              // the sequence number each slot holds (0: none yet)
        std::atomic<uint64_t> released[kMaxReaders];     // the highest sequence each reader has released
    };
    static constexpr uint64_t kMagic = 0x67736770706c616eull;   // "gsgpplan"
    Header *h = nullptr;
    unsigned char *base = nullptr;
    std::string name;
    bool writer = false;
    int32_t reader = -1;

    static size_t align64(size_t x) { return (x + 63) & ~(size_t)63; }
    static size_t header_bytes() { return align64(sizeof(Header)); }
    // slot: [pool_len, n_ins][plan P][prog_off 2P][prog_desc 2P][pool cap_pool][programs 2 x cap_ins]
    size_t off_plan() const { return 64; }
    size_t off_poff() const { return off_plan() + align64(sizeof(gsgp_plan_entry) * (size_t)h->P); }
    size_t off_pdesc() const { return off_poff() + align64(8 * (size_t)h->P); }
    size_t off_pool() const { return off_pdesc() + align64(8 * (size_t)h->P); }
    size_t off_prog() const { return off_pool() + align64(4 * (size_t)h->cap_pool); }
    static size_t slot_size(int32_t P, int32_t cap_pool, int32_t cap_ins)
    {
        return 64 + align64(sizeof(gsgp_plan_entry) * (size_t)P) + 2 * align64(8 * (size_t)P) + align64(4 * (size_t)cap_pool) +
               align64(8 * (size_t)cap_ins);
    }
    unsigned char *slot(uint64_t seq) const { return base + header_bytes() + (size_t)(seq & 1) * h->slot_bytes; }

    template <typename F> static bool wait_until(F done, int timeout_ms)      // spin briefly, then sleep in 20 us steps
    {
        for (int i = 0; i < 2000; ++i) { if (done()) return true; _mm_pause(); }
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            if (done()) return true;
            if (timeout_ms >= 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(timeout_ms)) return false;
            std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
    }
};
extern "C" {

gsgph_channel *gsgph_channel_create(const char *name, int32_t population_size, int32_t cap_pool_ints, int32_t cap_instructions,
                                    int32_t n_readers)
{
    if (!name || population_size < 1 || cap_pool_ints < 1 || cap_instructions < 1 || n_readers < 0 ||
        n_readers > gsgph_channel::kMaxReaders) return nullptr;
    const size_t slot = gsgph_channel::slot_size(population_size, cap_pool_ints, cap_instructions);
    const size_t total = gsgph_channel::header_bytes() + 2 * slot;
    shm_unlink(name);                                                          // a stale segment of a dead run
    const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return nullptr;
    if (ftruncate(fd, (off_t)total) != 0) { close(fd); shm_unlink(name); return nullptr; }
    void *m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) { shm_unlink(name); return nullptr; }
    gsgph_channel *c = new gsgph_channel();
    c->base = static_cast<unsigned char *>(m); c->h = new (m) gsgph_channel::Header();
    c->name = name; c->writer = true;
    c->h->magic = gsgph_channel::kMagic; c->h->P = population_size; c->h->cap_pool = cap_pool_ints; c->h->cap_ins = cap_instructions;
    c->h->n_readers = n_readers; c->h->slot_bytes = slot; c->h->total_bytes = total;
    c->h->published[0].store(0); c->h->published[1].store(0);
    for (auto &r : c->h->released) r.store(0);
    c->h->ready.store(1, std::memory_order_release);
    return c;
}

gsgph_channel *gsgph_channel_attach(const char *name, int32_t reader_index, int32_t timeout_ms)
{
    if (!name || reader_index < 0 || reader_index >= gsgph_channel::kMaxReaders) return nullptr;
    int fd = -1;
    struct stat st{};
    // the writer may not have created (or sized) the segment yet
    if (!gsgph_channel::wait_until([&] {
            if (fd < 0) fd = shm_open(name, O_RDWR, 0600);
            return fd >= 0 && fstat(fd, &st) == 0 && (size_t)st.st_size >= gsgph_channel::header_bytes();
        }, timeout_ms)) { if (fd >= 0) close(fd); return nullptr; }
    void *hm = mmap(nullptr, gsgph_channel::header_bytes(), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (hm == MAP_FAILED) { close(fd); return nullptr; }
    auto *hh = static_cast<gsgph_channel::Header *>(hm);
    const bool ok = gsgph_channel::wait_until([&] { return hh->ready.load(std::memory_order_acquire) == 1; }, timeout_ms) &&
                    hh->magic == gsgph_channel::kMagic && reader_index < hh->n_readers;
    const size_t total = ok ? (size_t)hh->total_bytes : 0;
    munmap(hm, gsgph_channel::header_bytes());
    if (!ok) { close(fd); return nullptr; }
    void *m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return nullptr;
    gsgph_channel *c = new gsgph_channel();
    c->base = static_cast<unsigned char *>(m); c->h = static_cast<gsgph_channel::Header *>(m);
    c->name = name; c->reader = reader_index;
    return c;
}

void gsgph_channel_close(gsgph_channel *c)
{
    if (!c) return;
    const size_t total = (size_t)c->h->total_bytes;
    munmap(c->base, total);
    if (c->writer) shm_unlink(c->name.c_str());
    delete c;
}

int gsgph_channel_publish(gsgph_channel *c, uint64_t seq, const gsgp_plan_entry *plan, const int32_t *pool, int32_t pool_len,
                          const uint32_t *programs, int32_t n_instructions, const int32_t *prog_off, const int32_t *prog_desc,
                          int32_t timeout_ms)
{
    if (!c || !c->writer || seq < 1 || !plan || pool_len < 0 || n_instructions < 0 || (pool_len > 0 && !pool) ||
        (n_instructions > 0 && (!programs || !prog_off || !prog_desc))) return 1;
    if (pool_len > c->h->cap_pool || n_instructions > c->h->cap_ins) return 3;
    const int32_t n_readers = c->h->n_readers;
    if (seq > 2 && !gsgph_channel::wait_until([&] {                             // the slot's previous occupant is seq - 2
            for (int32_t r = 0; r < n_readers; ++r) if (c->h->released[r].load(std::memory_order_acquire) + 2 < seq) return false;
            return true;
        }, timeout_ms)) return 5;
    unsigned char *s = c->slot(seq);
    const size_t P = (size_t)c->h->P;
    reinterpret_cast<int32_t *>(s)[0] = pool_len; reinterpret_cast<int32_t *>(s)[1] = n_instructions;
    memcpy(s + c->off_plan(), plan, sizeof(gsgp_plan_entry) * P);
    if (n_instructions > 0) {
        memcpy(s + c->off_poff(), prog_off, 8 * P);
        memcpy(s + c->off_pdesc(), prog_desc, 8 * P);
        memcpy(s + c->off_prog(), programs, 8 * (size_t)n_instructions);
    }
    if (pool_len > 0) memcpy(s + c->off_pool(), pool, 4 * (size_t)pool_len);
    c->h->published[seq & 1].store(seq, std::memory_order_release);
    return 0;
}

int gsgph_channel_acquire(gsgph_channel *c, uint64_t seq, const gsgp_plan_entry **plan, const int32_t **pool, int32_t *pool_len,
                          const uint32_t **programs, int32_t *n_instructions, const int32_t **prog_off, const int32_t **prog_desc,
                          int32_t timeout_ms)
{
    if (!c || c->writer || seq < 1) return 1;
    if (!gsgph_channel::wait_until([&] { return c->h->published[seq & 1].load(std::memory_order_acquire) >= seq; }, timeout_ms)) return 5;
    if (c->h->published[seq & 1].load(std::memory_order_acquire) != seq) return 6;   // overwritten: this reader skipped a release
    const unsigned char *s = c->slot(seq);
    if (plan) *plan = reinterpret_cast<const gsgp_plan_entry *>(s + c->off_plan());
    if (pool) *pool = reinterpret_cast<const int32_t *>(s + c->off_pool());
    if (pool_len) *pool_len = reinterpret_cast<const int32_t *>(s)[0];
    if (programs) *programs = reinterpret_cast<const uint32_t *>(s + c->off_prog());
    if (n_instructions) *n_instructions = reinterpret_cast<const int32_t *>(s)[1];
    if (prog_off) *prog_off = reinterpret_cast<const int32_t *>(s + c->off_poff());
    if (prog_desc) *prog_desc = reinterpret_cast<const int32_t *>(s + c->off_pdesc());
    return 0;
}

int gsgph_channel_release(gsgph_channel *c, uint64_t seq)
{
    if (!c || c->writer || c->reader < 0) return 1;
    c->h->released[c->reader].store(seq, std::memory_order_release);
    return 0;
}

int gsgph_replay_generations_shared(gsgph_planner *pl, gsgph_channel *ch, gsgp_ctx *ctx, uint64_t first_seq, int32_t n_generations,
                                    double *fit, double *fit_test, int32_t *index_best, double *best_fit_history,
                                    double *best_fit_test_history, int32_t timeout_ms)
{
    if (!ch || !ctx || !fit || !fit_test || !index_best || n_generations < 0 || first_seq < 1) return 1;
    if (ch->writer ? (!pl || !pl->programs) : pl != nullptr) return 1;         // the writer plans (programs on), readers do not
    for (int32_t g = 0; g < n_generations; ++g) {
        const uint64_t seq = first_seq + (uint64_t)g;
        const gsgp_plan_entry *plan; const int32_t *pool, *prog_off, *prog_desc; const uint32_t *programs;
        int32_t pool_len = 0, n_ins = 0;
        if (ch->writer) {
            if (int rc = gsgph_planner_plan(pl, fit, *index_best, &plan, &pool, &pool_len)) return rc;
            const gsgph_planner::Buf &B = pl->buf[pl->out];
            programs = reinterpret_cast<const uint32_t *>(B.prog.data()); n_ins = B.n_ins;
            prog_off = B.prog_off.data(); prog_desc = B.prog_desc.data();
            if (int rc = gsgph_channel_publish(ch, seq, plan, pool, pool_len, programs, n_ins, prog_off, prog_desc, timeout_ms)) return rc;
            gsgph_planner_prefetch(pl, *index_best);                            // the next generation's draws start now
        } else if (int rc = gsgph_channel_acquire(ch, seq, &plan, &pool, &pool_len, &programs, &n_ins, &prog_off, &prog_desc, timeout_ms)) return rc;
        const int rc = gsgp_generation_compiled(ctx, plan, pool, pool_len, programs, n_ins, prog_off, prog_desc, fit, fit_test, index_best);
        if (!ch->writer) gsgph_channel_release(ch, seq);                        // the call has copied what it needs
        if (rc) return 4;                                                       // gsgp_last_error(ctx) says why
        if (best_fit_history) best_fit_history[g] = fit[*index_best];           // :1495-1496
        if (best_fit_test_history) best_fit_test_history[g] = fit_test[*index_best];
    }
    return 0;
}

void gsgph_planner_stats(const gsgph_planner *pl, int64_t *hits, int64_t *misses, int64_t *redrawn)
{
    if (hits) *hits = pl ? pl->hits : 0;
    if (misses) *misses = pl ? pl->misses : 0;
    if (redrawn) *redrawn = pl ? pl->redrawn : 0;
}

// strconv.FormatFloat(v, 'g', -1, 64) as fmt's %v prints it: shortest round-trip digits, exponent
// form when exp < -4 || exp >= 6, at least two exponent digits.
int gsgph_compile_tree(const int32_t *tree, int32_t len, int32_t n_symbols, uint32_t *code_ab, int32_t cap,
                       int32_t *n_ins, int32_t *spill_levels)
{
    if (!tree || len <= 0) return 1;
    const char *err = "";
    std::vector<gsgp::Ins> prog((size_t)len / 4 + 2);
    std::vector<uint8_t> need;
    std::vector<gsgp::TreeFrame> frames;
    int n = 0, nd = 0;
    if (!gsgp::compile_prefix_tree(tree, len, n_symbols, prog.data(), n, nd, need, frames, err)) return 1;
    // the device-side builder (plan_draw_kernel) goes through the postfix form: both must agree instruction for instruction
    {
        std::vector<uint16_t> post;
        std::string e2;
        int32_t sp = 0;
        if (!gsgp::prefix_to_postfix(tree, len, n_symbols, post, sp, e2)) return 1;
        std::vector<gsgp::Ins> prog2(post.size() / 2 + 1);
        std::vector<uint32_t> scratch(post.size());
        std::vector<gsgp::CompileFrame> fr(post.size() / 2 + 2);
        int n2 = 0, nd2 = 0;
        if (!gsgp::compile_postfix(post.data(), (int)post.size(), prog2.data(), n2, nd2, scratch.data(), fr.data(), (int)fr.size())) return 2;
        if (n2 != n || nd2 != nd) return 4;
        for (int i = 0; i < n; ++i) if (prog[i].code != prog2[i].code || prog[i].ab != prog2[i].ab) return 4;
    }
    if (n > cap) return 3;
    for (int i = 0; i < n; ++i) { code_ab[2 * i] = prog[i].code; code_ab[2 * i + 1] = prog[i].ab; }
    if (n_ins) *n_ins = n;
    if (spill_levels) *spill_levels = nd;
    return 0;
}

// ---- ingest ------------------------------------------------------------------------------------------
namespace {

inline bool is_space(char ch) { return ch == ' ' || ch == '\n' || ch == '\t' || ch == '\r' || ch == '\v' || ch == '\f'; }

// one token [first, last) -> value; false on a token strconv.ParseFloat would reject
inline bool parse_token(const char *first, const char *last, double &v)
{
    if (*first == '+' && last - first > 1) ++first;                    // strconv.ParseFloat accepts a leading '+'
    const auto r = std::from_chars(first, last, v);
    if (r.ec == std::errc::result_out_of_range) {                      // ParseFloat: +-Inf with ErrRange; atof keeps the value
        v = std::strtod(std::string(first, last).c_str(), nullptr);
        return true;
    }
    return r.ec == std::errc() && r.ptr == last;
}

// tokens of text[b, e) parsed into out (the header reader's helper); returns tokens seen, -1 on a bad token
int64_t scan_tokens(const char *text, int64_t b, int64_t e, double *out)
{
    int64_t n = 0, i = b;
    while (i < e) {
        while (i < e && is_space(text[i])) ++i;
        if (i >= e) break;
        int64_t j = i;
        while (j < e && !is_space(text[j])) ++j;
        if (out && !parse_token(text + i, text + j, out[n])) return -1;
        ++n;
        i = j;
    }
    return n;
}

// tokens that START in [a, b) of text[0, size): a non-space byte preceded by a space (or by the start of the text).
// Branch-free over independent byte pairs, so the compiler vectorises it (several GB/s per thread): cheap enough to
// run as a separate counting pass, which lets every thread parse straight into its final place.
inline unsigned space_flag(unsigned char ch) { return (unsigned)(ch == ' ') | (unsigned)((unsigned)(ch - 9u) <= 4u); }
int64_t count_starts(const char *text, int64_t a, int64_t b)
{
    int64_t n = 0;
    if (a == 0 && b > 0) { n += !space_flag((unsigned char)text[0]); a = 1; }
    const unsigned char *u = reinterpret_cast<const unsigned char *>(text);
    int64_t i = a;
#if defined(__SSE2__)
    {   // 16 byte pairs per step: space(u[i-1]) & !space(u[i]) -> mask -> popcount
        const __m128i nine = _mm_set1_epi8(9), four = _mm_set1_epi8(4), blank = _mm_set1_epi8(' ');
        auto spaces = [&](__m128i x) {
            const __m128i d = _mm_sub_epi8(x, nine);                               // (x - 9) <= 4 unsigned: \t \n \v \f \r
            return _mm_or_si128(_mm_cmpeq_epi8(x, blank), _mm_cmpeq_epi8(_mm_min_epu8(d, four), d));
        };
        for (; i + 16 <= b; i += 16) {
            const __m128i prev = _mm_loadu_si128(reinterpret_cast<const __m128i *>(u + i - 1));
            const __m128i cur = _mm_loadu_si128(reinterpret_cast<const __m128i *>(u + i));
            n += __builtin_popcount((unsigned)_mm_movemask_epi8(_mm_andnot_si128(spaces(cur), spaces(prev))));
        }
    }
#endif
    for (; i < b; ++i) n += space_flag(u[i - 1]) & (space_flag(u[i]) ^ 1u);
    return n;
}

// the first max_tokens tokens that start in [a, b) parsed into out; returns how many, or -1 on a bad token.
// from_chars finds the end of a number itself; the byte after it must be whitespace or the end of the text.
int64_t parse_range(const char *text, int64_t size, int64_t a, int64_t b, double *out, int64_t max_tokens)
{
    const char *p = text + a, *pb = text + b, *pe = text + size;
    if (a > 0 && !is_space(text[a - 1])) while (p < pe && !is_space(*p)) ++p;    // straddles a: the previous range's token
    int64_t n = 0;
    while (n < max_tokens) {
        while (p < pe && is_space(*p)) ++p;
        if (p >= pb || p >= pe) break;
        const char *first = p;
        if (*first == '+' && pe - first > 1 && !is_space(first[1])) ++first;      // strconv.ParseFloat accepts a leading '+'
        double v = 0.0;
        const auto r = std::from_chars(first, pe, v);
        const char *q = r.ptr;
        if ((r.ec != std::errc() && r.ec != std::errc::result_out_of_range) || q == first || (q != pe && !is_space(*q))) return -1;
        if (r.ec == std::errc::result_out_of_range) v = std::strtod(std::string(first, q).c_str(), nullptr);   // +-Inf / 0, like ParseFloat's value
        out[n++] = v;
        p = q;
    }
    return n;
}

// a file's bytes: mapped (page-cache pages, faulted in by the parsing threads in parallel) or, where mmap is not
// possible, read into memory
struct FileView {
    const char *data = nullptr; size_t size = 0;
    void *map = nullptr; std::string owned;
    bool open(const char *path)
    {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                map = m; data = static_cast<const char *>(m); size = (size_t)st.st_size;
                ::close(fd);
                return true;
            }
        }
        char chunk[1 << 16];                                            // pipes, empty files, mmap refused
        for (;;) {
            const ssize_t got = ::read(fd, chunk, sizeof chunk);
            if (got < 0) { ::close(fd); return false; }
            if (got == 0) break;
            owned.append(chunk, (size_t)got);
        }
        ::close(fd);
        data = owned.data(); size = owned.size();
        return true;
    }
    ~FileView() { if (map) munmap(map, size); }
};

}  // namespace

int gsgph_parse_floats(const char *text, int64_t size, double *out, int64_t cap, int64_t *n_out, int32_t n_threads)
{
    if (n_threads <= 0) n_threads = (int32_t)std::min<unsigned>(8, std::max<unsigned>(1, std::thread::hardware_concurrency()));
    if (size < (1 << 16)) n_threads = 1;
    std::vector<int64_t> cut((size_t)n_threads + 1), count((size_t)n_threads, 0), first((size_t)n_threads + 1, 0);
    for (int t = 0; t <= n_threads; ++t) cut[(size_t)t] = size * t / n_threads;
    auto run = [&](auto &&fn) {
        std::vector<std::thread> th;
        for (int t = 1; t < n_threads; ++t) th.emplace_back(fn, t);
        fn(0);
        for (auto &x : th) x.join();
    };
    // pass 1 (vectorised): tokens starting in each byte range -> where each range's values go
    run([&](int t) { count[(size_t)t] = count_starts(text, cut[(size_t)t], cut[(size_t)t + 1]); });
    for (int t = 0; t < n_threads; ++t) first[(size_t)t + 1] = first[(size_t)t] + count[(size_t)t];
    if (n_out) *n_out = first[(size_t)n_threads];
    // pass 2: every range parses straight into its place; tokens beyond cap are ignored (callers read a prefix)
    std::vector<int> bad((size_t)n_threads, 0);
    run([&](int t) {
        const int64_t f0 = first[(size_t)t], want = std::min(count[(size_t)t], cap - f0);
        if (want > 0 && parse_range(text, size, cut[(size_t)t], cut[(size_t)t + 1], out + f0, want) != want) bad[(size_t)t] = 1;
    });
    for (int b : bad) if (b) return 2;
    return 0;
}

int gsgph_read_semantic(const char *path, int32_t n, double *out)       // read_sem main_cpu.go:710-733
{
    FileView f;
    if (!path || !f.open(path)) return 1;                               // panic(err)
    int64_t got = 0;
    const int rc = gsgph_parse_floats(f.data, (int64_t)f.size, out, n, &got, 0);
    if (rc) return 2;                                                   // "Cannot parse semantic value"
    return got >= n ? 0 : 3;                                            // "Not enough values when reading semantic file"
}

int gsgph_read_dataset(const char *train_path, const char *test_path, int32_t *nvar, int32_t *nrow, int32_t *nrow_test,
                       double **rowmajor)                                // read_input_data main_cpu.go:381-421
{
    FileView a, b;
    if (!train_path || !test_path || !a.open(train_path) || !b.open(test_path)) return 1;
    auto header = [](const FileView &s, int32_t &v, int32_t &r, int64_t &body) {
        double h[2];
        // the two header tokens: parse a short prefix sequentially
        int64_t i = 0, seen = 0;
        while (i < (int64_t)s.size && seen < 2) {
            while (i < (int64_t)s.size && is_space(s.data[(size_t)i])) ++i;
            int64_t j = i;
            while (j < (int64_t)s.size && !is_space(s.data[(size_t)j])) ++j;
            if (j > i) { if (scan_tokens(s.data, i, j, h + seen) != 1) return false; ++seen; }
            i = j;
        }
        if (seen < 2) return false;
        v = (int32_t)h[0]; r = (int32_t)h[1]; body = i;
        return v >= 0 && r >= 0 && h[0] == (double)v && h[1] == (double)r;      // atoi: integers
    };
    int32_t va = 0, ra = 0, vb = 0, rb = 0; int64_t ba = 0, bb = 0;
    if (!header(a, va, ra, ba) || !header(b, vb, rb, bb)) return 2;
    if (va != vb) return 4;                                             // "Train and Test datasets must have the same number of variables"
    const size_t stride = (size_t)va + 1, na = (size_t)ra * stride, nb = (size_t)rb * stride;
    double *m = (double *)malloc(sizeof(double) * std::max<size_t>(1, na + nb));
    if (!m) return 5;
    int64_t ga = 0, gb = 0;
    int rc = gsgph_parse_floats(a.data + ba, (int64_t)a.size - ba, m, (int64_t)na, &ga, 0);
    if (!rc) rc = gsgph_parse_floats(b.data + bb, (int64_t)b.size - bb, m + na, (int64_t)nb, &gb, 0);
    if (rc || ga < (int64_t)na || gb < (int64_t)nb) { free(m); return rc ? 2 : 3; }
    *nvar = va; *nrow = ra; *nrow_test = rb; *rowmajor = m;
    return 0;
}

void gsgph_free(void *p) { free(p); }

namespace {
// strconv.FormatFloat(v, 'g', -1, 64) as fmt's %v prints it: shortest round-trip digits (std::to_chars), exponent
// form when exp < -4 || exp >= 6, at least two exponent digits.  Returns the number of chars written (<= 24).
inline int format_go_v(double v, char *out)
{
    if (std::isnan(v)) { memcpy(out, "NaN", 3); return 3; }
    if (std::isinf(v)) { memcpy(out, v > 0 ? "+Inf" : "-Inf", 4); return 4; }
    if (v == 0) { if (std::signbit(v)) { memcpy(out, "-0", 2); return 2; } out[0] = '0'; return 1; }
    // shortest digits in scientific form, written where the result goes: [-]d[.ddd]e[+-]XX[X]; then the decimal point
    // is moved in place for the exponents Go prints without one
    char *o = out;
    const auto r = std::to_chars(o, o + 32, v, std::chars_format::scientific);
    char *endp = r.ptr;
    if (*o == '-') ++o;                                     // o: first digit
    char *e = endp - 4;                                     // 'e' sits 4 or 5 chars from the end (2- or 3-digit exponent)
    if (*e != 'e') --e;
    int x = e[2] - '0';
    for (char *q = e + 3; q < endp; ++q) x = x * 10 + (*q - '0');
    if (e[1] == '-') x = -x;
    if (x < -4 || x >= 6) return (int)(endp - out);         // exponent form: Go's %v layout is to_chars' own
    const int nd = (int)(e - o) - (e - o > 1 ? 1 : 0);      // digits (the '.' does not count)
    if (x >= 0) {                                           // d.ddd -> dddd[.dd] (zero-padded when digits run out)
        if (nd <= x + 1) {                                  // all digits in front of the point: drop it, pad
            if (nd > 1) memmove(o + 1, o + 2, (size_t)nd - 1);
            for (int i = nd; i <= x; ++i) o[i] = '0';
            return (int)(o + x + 1 - out);
        }
        if (x > 0) { memmove(o + 1, o + 2, (size_t)x); o[x + 1] = '.'; }
        return (int)(o + nd + 1 - out);
    }
    // 0.000ddd: digits move right by -x + 1 (the point inside them disappears)
    char d[24];
    d[0] = o[0];
    if (nd > 1) memcpy(d + 1, o + 2, (size_t)nd - 1);
    o[0] = '0'; o[1] = '.';
    for (int i = 0; i < -x - 1; ++i) o[2 + i] = '0';
    memcpy(o + 1 - x, d, (size_t)nd);
    return (int)(o + 1 - x + nd - out);
}
}  // namespace

int gsgph_format_float(double v, char *buf, int32_t cap)
{
    char tmp[32];
    const int n = format_go_v(v, tmp);
    if (cap > 0) { const int m = std::min(n, cap - 1); memcpy(buf, tmp, (size_t)m); buf[m] = 0; }
    return n;
}

// Semantic.String() (main_cpu.go:123-126): the values in %v form joined by commas; what the best-of-generation
// semantic lines are made of (:1499-1500).  Multi-threaded; returns non-zero if cap is too small (25 bytes per value
// always suffice).
int gsgph_format_semantic(const double *v, int64_t n, char *out, int64_t cap, int64_t *len_out, int32_t n_threads)
{
    if (n_threads <= 0) n_threads = (int32_t)std::min<unsigned>(8, std::max<unsigned>(1, std::thread::hardware_concurrency()));
    if (n < (1 << 14)) n_threads = 1;
    std::vector<std::string> part((size_t)n_threads);
    auto work = [&](int t) {
        const int64_t a = n * t / n_threads, b = n * (t + 1) / n_threads;
        std::string &s = part[(size_t)t];
        s.resize((size_t)(b - a) * 25);
        char *o = &s[0];
        for (int64_t i = a; i < b; ++i) { if (i) *o++ = ','; o += format_go_v(v[i], o); }
        s.resize((size_t)(o - s.data()));
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    int64_t total = 0;
    for (auto &s : part) total += (int64_t)s.size();
    if (len_out) *len_out = total;
    if (total > cap) return 1;
    int64_t at = 0;
    for (auto &s : part) { memcpy(out + at, s.data(), s.size()); at += (int64_t)s.size(); }
    return 0;
}

// ---- output path: the reference's line writers, off the generation loop's critical path ---------
// main_cpu.go writes, every generation, the best individual's fitness (:1495-1496) and its whole train / test semantic
// (:1499-1500) as text through a writer that is a gzip stream when the path ends in ".gz" (:1227-1241; the defaults are
// semantictrain.txt.gz / semantictest.txt.gz, :175-176).  At 1M+ cases that is tens of MB of text per generation:
// formatting and deflate would dominate a fast loop.  gsgph_writer takes a COPY of the values and returns; a background
// thread formats (multi-threaded shortest-digits %v) and compresses.  Compression is parallel: the text is cut into
// blocks, each deflated as its own gzip member by a pool thread, members written in order -- a concatenation of gzip
// members is a valid gzip file (Go's compress/gzip and Python's gzip read it as one stream).
struct gsgph_writer {
    FILE *f = nullptr;
    bool gz = false;
    int32_t n_threads = 4;
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv_put, cv_get;
    struct Job { std::vector<double> v; bool semantic; };
    std::deque<Job> q;
    size_t queued_bytes = 0;
    bool closing = false, failed = false;
    std::string pending;                               // text not yet written / compressed (small lines are batched)

    static constexpr size_t kBlock = (size_t)4 << 20;  // text bytes per gzip member
    static constexpr size_t kMaxQueued = (size_t)512 << 20;

    // One gzip member holding text[0, n) as ONE dynamic-Huffman deflate block of literals (RFC 1951 3.2.7, RFC 1952).
    // The output lines are decimal digits with no repeats worth matching: LZ77 at zlib's fastest level finds almost
    // nothing (ratio 0.49 at 53 MB/s per core) while entropy coding alone gives 0.47 -- so the member is built directly:
    // byte histogram -> Huffman code lengths -> table-driven bit packing.  Returns false when the code would need
    // more than 15 bits (the caller falls back to zlib).
    static bool huffman_member(const char *src, size_t n, std::string &out)
    {
        if (n == 0 || n >= ((size_t)1 << 32)) return false;
        const unsigned char *u = reinterpret_cast<const unsigned char *>(src);
        // histogram (four tables: consecutive equal bytes do not wait on each other's increments)
        uint32_t h4[4][256] = {};
        size_t i = 0;
        for (; i + 4 <= n; i += 4) { ++h4[0][u[i]]; ++h4[1][u[i + 1]]; ++h4[2][u[i + 2]]; ++h4[3][u[i + 3]]; }
        for (; i < n; ++i) ++h4[0][u[i]];
        uint64_t freq[257];
        for (int c = 0; c < 256; ++c) freq[c] = (uint64_t)h4[0][c] + h4[1][c] + h4[2][c] + h4[3][c];
        freq[256] = 1;                                                          // end of block
        // Huffman code lengths: repeatedly merge the two lightest nodes (<= 257 leaves: a linear scan is fine)
        int parent[2 * 257], nn = 0, live[257], nlive = 0;
        uint64_t w[2 * 257];
        int leaf_of[257];
        for (int c = 0; c < 257; ++c) if (freq[c]) { w[nn] = freq[c]; parent[nn] = -1; leaf_of[c] = nn; live[nlive++] = nn; ++nn; } else leaf_of[c] = -1;
        uint8_t len[257] = {};
        if (nlive == 1) return false;                                           // (cannot happen: EOB + >= 1 literal)
        while (nlive > 1) {
            int a = 0, b = 1;
            if (w[live[b]] < w[live[a]]) std::swap(a, b);
            for (int k = 2; k < nlive; ++k) {
                if (w[live[k]] < w[live[a]]) { b = a; a = k; }
                else if (w[live[k]] < w[live[b]]) b = k;
            }
            w[nn] = w[live[a]] + w[live[b]]; parent[nn] = -1;
            parent[live[a]] = nn; parent[live[b]] = nn;
            const int hi = std::max(a, b), lo = std::min(a, b);
            live[lo] = nn; live[hi] = live[--nlive];
            ++nn;
        }
        int max_len = 0;
        for (int c = 0; c < 257; ++c) {
            if (leaf_of[c] < 0) continue;
            int d = 0;
            for (int x = leaf_of[c]; parent[x] >= 0; x = parent[x]) ++d;
            len[c] = (uint8_t)d;
            max_len = std::max(max_len, d);
        }
        if (max_len > 15) return false;
        // canonical codes (RFC 1951 3.2.2), stored bit-reversed: deflate packs Huffman codes starting from their MSB
        auto canonical = [](const uint8_t *l, int count, uint16_t *code) {
            int bl_count[16] = {}, next[16] = {};
            for (int c = 0; c < count; ++c) ++bl_count[l[c]];
            bl_count[0] = 0;
            for (int b = 1, cd = 0; b < 16; ++b) { cd = (cd + bl_count[b - 1]) << 1; next[b] = cd; }
            for (int c = 0; c < count; ++c) {
                if (!l[c]) { code[c] = 0; continue; }
                int v = next[l[c]]++, r = 0;
                for (int b = 0; b < l[c]; ++b) { r = (r << 1) | (v & 1); v >>= 1; }
                code[c] = (uint16_t)r;
            }
        };
        uint16_t code[257];
        canonical(len, 257, code);
        // the 257 + 2 code lengths, run-length coded with symbols 16 / 17 / 18 (3.2.7); the two distance codes have
        // length 1 (what zlib sends when no distance is used)
        uint8_t all[259];
        memcpy(all, len, 257); all[257] = 1; all[258] = 1;
        struct Sym { uint8_t s, extra, nbits; };
        std::vector<Sym> rle;
        for (int k = 0; k < 259;) {
            int run = 1;
            while (k + run < 259 && all[k + run] == all[k]) ++run;
            const uint8_t v = all[k];
            int left = run;
            if (v == 0) {
                while (left >= 11) { const int r = std::min(left, 138); rle.push_back({18, (uint8_t)(r - 11), 7}); left -= r; }
                if (left >= 3) { rle.push_back({17, (uint8_t)(left - 3), 3}); left = 0; }
                while (left-- > 0) rle.push_back({0, 0, 0});
            } else {
                rle.push_back({v, 0, 0}); --left;
                while (left >= 3) { const int r = std::min(left, 6); rle.push_back({16, (uint8_t)(r - 3), 2}); left -= r; }
                while (left-- > 0) rle.push_back({v, 0, 0});
            }
            k += run;
        }
        // a fixed complete code for the 19 code-length symbols: the 13 that can occur often get 4 bits, 6 get 5 bits
        uint8_t cl_len[19];
        for (int c = 0; c < 19; ++c) cl_len[c] = 5;
        for (int c : {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 16, 17, 18}) cl_len[c] = 4;
        uint16_t cl_code[19];
        canonical(cl_len, 19, cl_code);
        // worst case: every literal 15 bits
        out.resize(10 + 8 + 400 + (n * 15 + 7) / 8 + 16);
        unsigned char *o = reinterpret_cast<unsigned char *>(&out[0]);
        static const unsigned char head[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff};
        memcpy(o, head, 10); o += 10;
        uint64_t acc = 0; int nb = 0;
        auto put = [&](uint32_t v, int bits) {                                  // LSB-first bit stream
            acc |= (uint64_t)v << nb; nb += bits;
            if (nb >= 32) { memcpy(o, &acc, 4); o += 4; acc >>= 32; nb -= 32; }
        };
        put(1, 1); put(2, 2);                                                   // BFINAL = 1, BTYPE = 10 (dynamic)
        put(0, 5); put(1, 5); put(19 - 4, 4);                                   // HLIT = 257, HDIST = 2, HCLEN = 19
        static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        for (int k = 0; k < 19; ++k) put(cl_len[order[k]], 3);
        for (const Sym &y : rle) { put(cl_code[y.s], cl_len[y.s]); if (y.nbits) put(y.extra, y.nbits); }
        // the literals: code | length << 16 per byte, two symbols per flush check (2 x 15 bits on top of < 32)
        uint32_t tab[256];
        for (int c = 0; c < 256; ++c) tab[c] = (uint32_t)code[c] | ((uint32_t)len[c] << 16);
        i = 0;
        for (; i + 2 <= n; i += 2) {
            const uint32_t t0 = tab[u[i]], t1 = tab[u[i + 1]];
            acc |= (uint64_t)(t0 & 0xFFFFu) << nb; nb += (int)(t0 >> 16);
            acc |= (uint64_t)(t1 & 0xFFFFu) << nb; nb += (int)(t1 >> 16);
            if (nb >= 32) { memcpy(o, &acc, 4); o += 4; acc >>= 32; nb -= 32; }
        }
        for (; i < n; ++i) put(code[u[i]], len[u[i]]);
        put(code[256], len[256]);                                               // end of block
        while (nb > 0) { *o++ = (unsigned char)(acc & 0xFF); acc >>= 8; nb -= 8; }
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), u, (uInt)n), isize = (uint32_t)n;
        memcpy(o, &crc, 4); memcpy(o + 4, &isize, 4); o += 8;                   // little-endian host (x86-64)
        out.resize((size_t)(o - reinterpret_cast<unsigned char *>(&out[0])));
        return true;
    }
    static bool deflate_member(const char *src, size_t n, std::string &out)
    {
        static const bool use_zlib = getenv("GSGP_GZIP_ZLIB") != nullptr;       // A/B: zlib's own encoder for everything
        if (!use_zlib && huffman_member(src, n, out)) return true;
        z_stream zs{};
        if (deflateInit2(&zs, Z_BEST_SPEED, Z_DEFLATED, 15 + 16, 8, Z_HUFFMAN_ONLY) != Z_OK) return false;
        out.resize(deflateBound(&zs, (uLong)n) + 64);
        zs.next_in = reinterpret_cast<Bytef *>(const_cast<char *>(src)); zs.avail_in = (uInt)n;
        zs.next_out = reinterpret_cast<Bytef *>(&out[0]); zs.avail_out = (uInt)out.size();
        const int rc = deflate(&zs, Z_FINISH);
        out.resize(zs.total_out);
        deflateEnd(&zs);
        return rc == Z_STREAM_END;
    }
    // writes text[0, n): plain, or as ceil(n / kBlock) gzip members compressed in parallel
    void emit(const char *text, size_t n)
    {
        if (!n || failed) return;
        if (!gz) { if (fwrite(text, 1, n, f) != n) failed = true; return; }
        const size_t nb = (n + kBlock - 1) / kBlock;
        std::vector<std::string> out(nb);
        std::vector<char> ok(nb, 1);
        auto work = [&](size_t t, size_t stride) {
            for (size_t b = t; b < nb; b += stride)
                ok[b] = deflate_member(text + b * kBlock, std::min(kBlock, n - b * kBlock), out[b]);
        };
        const size_t nt = std::min<size_t>((size_t)std::max(1, n_threads), nb);
        std::vector<std::thread> th;
        for (size_t t = 1; t < nt; ++t) th.emplace_back(work, t, nt);
        work(0, nt);
        for (auto &x : th) x.join();
        for (size_t b = 0; b < nb; ++b)
            if (!ok[b] || fwrite(out[b].data(), 1, out[b].size(), f) != out[b].size()) { failed = true; return; }
    }
    void flush_pending(bool force)
    {
        if (pending.size() >= kBlock || (force && !pending.empty())) { emit(pending.data(), pending.size()); pending.clear(); }
    }
    void run()
    {
        std::string text;
        for (;;) {
            Job job;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_get.wait(lk, [&] { return closing || !q.empty(); });
                if (q.empty()) break;
                job = std::move(q.front()); q.pop_front();
                queued_bytes -= job.v.size() * sizeof(double);
                cv_put.notify_all();
            }
            const int64_t n = (int64_t)job.v.size();
            if (job.semantic) {
                text.resize((size_t)n * 25 + 2);
                int64_t len = 0;
                gsgph_format_semantic(job.v.data(), n, &text[0], (int64_t)text.size() - 1, &len, n_threads);
                text[(size_t)len] = '\n';
                if ((size_t)len + 1 >= kBlock) {                         // a big line: its own members, in order
                    flush_pending(true);
                    emit(text.data(), (size_t)len + 1);
                } else {
                    pending.append(text.data(), (size_t)len + 1);
                    flush_pending(false);
                }
            } else {
                char buf[40];
                for (double x : job.v) { const int k = format_go_v(x, buf); buf[k] = '\n'; pending.append(buf, (size_t)k + 1); }
                flush_pending(false);
            }
        }
        flush_pending(true);
    }
};

int gsgph_gzip_member(const char *src, int64_t n, char *out, int64_t cap, int64_t *len_out)
{
    if (!src || n < 0 || !out) return 1;
    std::string m;
    if (n == 0 || !gsgph_writer::deflate_member(src, (size_t)n, m)) {
        if (n != 0) return 2;
        static const unsigned char empty[20] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        m.assign(reinterpret_cast<const char *>(empty), sizeof empty);
    }
    if (len_out) *len_out = (int64_t)m.size();
    if ((int64_t)m.size() > cap) return 3;
    memcpy(out, m.data(), m.size());
    return 0;
}

gsgph_writer *gsgph_writer_open(const char *path, int32_t n_threads)
{
    if (!path) return nullptr;
    FILE *f = fopen(path, "wb");                                         // os.Create, main_cpu.go:1229
    if (!f) return nullptr;
    gsgph_writer *w = new gsgph_writer();
    w->f = f;
    const size_t len = strlen(path);
    w->gz = len >= 3 && strcmp(path + len - 3, ".gz") == 0;              // :1237
    w->n_threads = n_threads > 0 ? n_threads : (int32_t)std::min<unsigned>(8, std::max<unsigned>(1, std::thread::hardware_concurrency()));
    w->worker = std::thread([w] { w->run(); });
    return w;
}

static int writer_put(gsgph_writer *w, const double *v, int64_t n, bool semantic)
{
    if (!w || (!v && n > 0) || n < 0) return 1;
    gsgph_writer::Job job;
    job.v.assign(v, v + n); job.semantic = semantic;
    std::unique_lock<std::mutex> lk(w->mu);
    if (w->closing) return 1;
    w->cv_put.wait(lk, [&] { return w->queued_bytes < gsgph_writer::kMaxQueued; });   // back-pressure, not unbounded memory
    w->queued_bytes += job.v.size() * sizeof(double);
    w->q.push_back(std::move(job));
    w->cv_get.notify_one();
    return w->failed ? 2 : 0;
}

int gsgph_writer_semantic(gsgph_writer *w, const double *v, int64_t n) { return writer_put(w, v, n, true); }
int gsgph_writer_float(gsgph_writer *w, double v) { return writer_put(w, &v, 1, false); }

int gsgph_writer_close(gsgph_writer *w)
{
    if (!w) return 1;
    { std::lock_guard<std::mutex> lk(w->mu); w->closing = true; w->cv_get.notify_all(); }
    if (w->worker.joinable()) w->worker.join();
    bool bad = w->failed;
    if (fclose(w->f) != 0) bad = true;
    delete w;
    return bad ? 2 : 0;
}

// ---- -proto_dump (main_cpu.go:1416-1436,1441-1488,1510-1517; evolution.proto) --------------------------------------
// The protobuf wire format written directly (proto3: packed repeated scalars, zero-valued scalars and empty
// bytes / repeated fields left out, fields in number order -- what proto.Marshal emits for these messages).
struct gsgph_proto {
    std::string evo, pop, ind, rt;       // serialized Evolution so far / open Population / scratch
    bool open = false;
    int32_t n_pop = 0;

    static void varint(std::string &o, uint64_t v)
    {
        while (v >= 0x80) { o.push_back((char)(v | 0x80)); v >>= 7; }
        o.push_back((char)v);
    }
    static void tag(std::string &o, uint32_t field, uint32_t wire) { varint(o, (uint64_t)field << 3 | wire); }
    static void f64(std::string &o, uint32_t field, double v)                 // double field = N; (left out when +0.0)
    {
        uint64_t b; memcpy(&b, &v, 8);
        if (b == 0) return;
        tag(o, field, 1); o.append(reinterpret_cast<const char *>(&b), 8);    // little-endian host (x86-64)
    }
    static void packed_f64(std::string &o, uint32_t field, const double *v, int64_t n)
    {
        if (n <= 0) return;
        tag(o, field, 2); varint(o, (uint64_t)n * 8);
        o.append(reinterpret_cast<const char *>(v), (size_t)n * 8);
    }
    static void packed_i32(std::string &o, uint32_t field, const int32_t *v, int64_t n, std::string &tmp)
    {
        if (n <= 0) return;
        tmp.clear();
        for (int64_t i = 0; i < n; ++i) varint(tmp, (uint64_t)(int64_t)v[i]); // int32: sign-extended to 64 bits
        tag(o, field, 2); varint(o, tmp.size()); o.append(tmp);
    }
    static void bytes(std::string &o, uint32_t field, const char *v, size_t n)
    {
        tag(o, field, 2); varint(o, n); o.append(v, n);
    }
};

gsgph_proto *gsgph_proto_new(void) { return new gsgph_proto(); }
void gsgph_proto_free(gsgph_proto *d) { delete d; }

int gsgph_proto_begin_population(gsgph_proto *d, int32_t generation)
{
    if (!d || d->open) return 1;
    d->open = true;
    d->pop.clear();
    if (generation != 0) { gsgph_proto::tag(d->pop, 1, 0); gsgph_proto::varint(d->pop, (uint64_t)(int64_t)generation); }
    return 0;
}

int gsgph_proto_add_individual(gsgph_proto *d, int32_t op, double fitness_train, double fitness_test,
                               const double *semantic_train, int64_t nrow, const double *semantic_test, int64_t nrow_test,
                               const uint8_t *contrib, int64_t contrib_len,
                               int32_t n_rt, const int32_t *const *rt_data, const int32_t *rt_len,
                               const double *const *rt_semantic_train, const double *const *rt_semantic_test)
{
    if (!d || !d->open || op < 0 || op > 3 || nrow < 0 || nrow_test < 0 || contrib_len < 0 || n_rt < 0) return 1;
    if ((nrow > 0 && !semantic_train) || (nrow_test > 0 && !semantic_test) || (contrib_len > 0 && !contrib)) return 1;
    if (n_rt > 0 && (!rt_data || !rt_len)) return 1;
    std::string &o = d->ind;
    o.clear();
    if (op != 0) { gsgph_proto::tag(o, 1, 0); gsgph_proto::varint(o, (uint64_t)op); }
    gsgph_proto::f64(o, 2, fitness_train);
    gsgph_proto::f64(o, 3, fitness_test);
    gsgph_proto::packed_f64(o, 4, semantic_train, nrow);
    gsgph_proto::packed_f64(o, 5, semantic_test, nrow_test);
    if (contrib_len > 0) gsgph_proto::bytes(o, 6, reinterpret_cast<const char *>(contrib), (size_t)contrib_len);
    std::string tmp;
    for (int32_t t = 0; t < n_rt; ++t) {                                       // RandomTree{data, semantic_train, semantic_test}
        if (rt_len[t] < 0 || (rt_len[t] > 0 && !rt_data[t])) return 1;
        d->rt.clear();
        gsgph_proto::packed_i32(d->rt, 1, rt_data[t], rt_len[t], tmp);
        if (rt_semantic_train && rt_semantic_train[t]) gsgph_proto::packed_f64(d->rt, 2, rt_semantic_train[t], nrow);
        if (rt_semantic_test && rt_semantic_test[t]) gsgph_proto::packed_f64(d->rt, 3, rt_semantic_test[t], nrow_test);
        gsgph_proto::bytes(o, 7, d->rt.data(), d->rt.size());
    }
    if (o.size() + d->pop.size() + d->evo.size() > (size_t)0x7fffffff - 64) return 4;   // protobuf's 2 GB limit (proto.Marshal fails too)
    gsgph_proto::bytes(d->pop, 2, o.data(), o.size());
    return 0;
}

int gsgph_proto_end_population(gsgph_proto *d)
{
    if (!d || !d->open) return 1;
    gsgph_proto::bytes(d->evo, 1, d->pop.data(), d->pop.size());
    d->pop.clear();
    d->open = false;
    ++d->n_pop;
    return 0;
}

int64_t gsgph_proto_size(const gsgph_proto *d) { return d ? (int64_t)d->evo.size() : -1; }

int gsgph_proto_bytes(const gsgph_proto *d, uint8_t *out, int64_t cap)
{
    if (!d || d->open || (!out && cap > 0) || cap < (int64_t)d->evo.size()) return 1;
    if (!d->evo.empty()) memcpy(out, d->evo.data(), d->evo.size());
    return 0;
}

int gsgph_proto_write(const gsgph_proto *d, const char *path)
{
    if (!d || d->open || !path) return 1;
    FILE *f = fopen(path, "wb");
    if (!f) return 2;
    const bool ok = d->evo.empty() || fwrite(d->evo.data(), 1, d->evo.size(), f) == d->evo.size();
    return (fclose(f) == 0 && ok) ? 0 : 2;
}

// ---- contributions (main_cpu.go:128-149,857,863-873,885,909,1196,1272-1283,1390-1397,1411,1503) ---------------------
// contrib[k] is a vector of num_models big integers: crossover adds the parents' vectors, everything else copies one.
// The values at most double per generation, so they are kept as fixed-width little-endian limb arrays [P][M][W] that
// widen by one limb when an addition carries out of the top (every ~64 generations) instead of as heap big.Ints.
struct gsgph_contrib {
    int32_t P = 0, M = 0, W = 1;
    std::vector<uint64_t> cur, nxt;

    void widen()                                                              // W -> W + 1 on `cur`
    {
        const size_t n = (size_t)P * M;
        std::vector<uint64_t> w(n * (size_t)(W + 1), 0);
        for (size_t i = 0; i < n; ++i) memcpy(&w[i * (size_t)(W + 1)], &cur[i * (size_t)W], (size_t)W * 8);
        cur.swap(w);
        ++W;
        nxt.assign(cur.size(), 0);
    }
    // rows [k0, k1) of the next generation; returns false if an addition carried out of the top limb
    bool rows(const gsgp_plan_entry *plan, int32_t k0, int32_t k1)
    {
        const size_t row = (size_t)M * W;
        bool ok = true;
        for (int32_t k = k0; k < k1; ++k) {
            const gsgp_plan_entry &e = plan[k];
            uint64_t *d = &nxt[(size_t)k * row];
            // the elite keeps its own vector whatever the operator (:847,:857 with i == old_i, :909); device-mode plans
            // (gsgp_get_plan) do not define p1 for it
            const uint64_t *a = &cur[(size_t)(e.elite ? k : e.p1) * row];
            if (e.op == GSGP_OP_XO && !e.elite) {                             // merge_contribs :869-873,885
                const uint64_t *b = &cur[(size_t)e.p2 * row];
                if (W == 1) {
                    uint64_t over = 0;
                    for (size_t i = 0; i < row; ++i) { d[i] = a[i] + b[i]; over |= (d[i] < a[i]); }
                    ok = ok && !over;
                } else {
                    for (int32_t m = 0; m < M; ++m) {
                        uint64_t carry = 0;
                        for (int32_t w = 0; w < W; ++w) {
                            const size_t i = (size_t)m * W + w;
                            const uint64_t t = a[i] + b[i], u = t + carry;
                            carry = (uint64_t)(t < a[i]) | (uint64_t)(u < t);
                            d[i] = u;
                        }
                        ok = ok && !carry;
                    }
                }
            } else {
                memcpy(d, a, row * 8);                                        // copy_contrib :857,863-867,909
            }
        }
        return ok;
    }
};

gsgph_contrib *gsgph_contrib_new(int32_t population_size, int32_t n_seeds)
{
    if (population_size < 1 || n_seeds < 0 || n_seeds > population_size) return nullptr;
    gsgph_contrib *c = new gsgph_contrib();
    c->P = population_size; c->M = n_seeds + 1;                               // init_tables(len(sem_seed) + 1) :1380
    c->cur.assign((size_t)c->P * c->M, 0);
    c->nxt.assign(c->cur.size(), 0);
    for (int32_t i = 0; i < c->P; ++i) c->cur[(size_t)i * c->M + (i < n_seeds ? i + 1 : 0)] = 1;   // :1390-1397
    return c;
}
void gsgph_contrib_free(gsgph_contrib *c) { delete c; }
int32_t gsgph_contrib_num_models(const gsgph_contrib *c) { return c ? c->M : -1; }
int32_t gsgph_contrib_limbs(const gsgph_contrib *c) { return c ? c->W : -1; }

int gsgph_contrib_apply(gsgph_contrib *c, const gsgp_plan_entry *plan, int32_t n_threads)
{
    if (!c || !plan) return 1;
    for (int32_t k = 0; k < c->P; ++k) {
        const gsgp_plan_entry &e = plan[k];
        if (e.elite) continue;
        if (e.p1 < 0 || e.p1 >= c->P) return 2;
        if (e.op == GSGP_OP_XO && (e.p2 < 0 || e.p2 >= c->P)) return 2;
    }
    if (n_threads <= 0) n_threads = (int32_t)std::min<unsigned>(8, std::max<unsigned>(1, std::thread::hardware_concurrency()));
    if ((int64_t)c->P * c->M * c->W < (1 << 16)) n_threads = 1;               // small tables: not worth the threads
    n_threads = std::min(n_threads, c->P);
    for (;;) {
        bool ok = true;
        if (n_threads == 1) ok = c->rows(plan, 0, c->P);
        else {
            std::vector<std::thread> th;
            std::vector<char> oks((size_t)n_threads, 1);
            for (int32_t t = 0; t < n_threads; ++t)
                th.emplace_back([&, t] { oks[(size_t)t] = c->rows(plan, (int32_t)((int64_t)c->P * t / n_threads), (int32_t)((int64_t)c->P * (t + 1) / n_threads)); });
            for (auto &x : th) x.join();
            for (char o : oks) ok = ok && o;
        }
        if (ok) break;
        c->widen();                                                           // a value outgrew W limbs: redo one limb wider
    }
    c->cur.swap(c->nxt);                                                      // update_tables :1196
    return 0;
}

// Contribution.String() (:141-146): the decimal values joined by commas
int gsgph_contrib_format(const gsgph_contrib *c, int32_t ind, char *out, int64_t cap, int64_t *len_out)
{
    if (!c || ind < 0 || ind >= c->P || (!out && cap > 0)) return 1;
    std::string s;
    std::vector<uint64_t> v((size_t)c->W);
    std::vector<uint64_t> chunks;
    for (int32_t m = 0; m < c->M; ++m) {
        memcpy(v.data(), &c->cur[((size_t)ind * c->M + m) * (size_t)c->W], (size_t)c->W * 8);
        int32_t top = c->W;
        chunks.clear();
        while (top > 0 && v[(size_t)top - 1] == 0) --top;
        while (top > 0) {                                                     // divide by 10^19, most significant limb first
            unsigned __int128 rem = 0;
            for (int32_t w = top - 1; w >= 0; --w) {
                const unsigned __int128 x = (rem << 64) | v[(size_t)w];
                v[(size_t)w] = (uint64_t)(x / 10000000000000000000ull);
                rem = x % 10000000000000000000ull;
            }
            chunks.push_back((uint64_t)rem);
            while (top > 0 && v[(size_t)top - 1] == 0) --top;
        }
        if (m) s.push_back(',');
        if (chunks.empty()) s.push_back('0');
        else {
            char b[24];
            s.append(b, (size_t)snprintf(b, sizeof b, "%llu", (unsigned long long)chunks.back()));
            for (size_t i = chunks.size() - 1; i-- > 0;) s.append(b, (size_t)snprintf(b, sizeof b, "%019llu", (unsigned long long)chunks[i]));
        }
    }
    if (len_out) *len_out = (int64_t)s.size();
    if ((int64_t)s.size() > cap) return 3;
    memcpy(out, s.data(), s.size());
    return 0;
}

// time.Duration.String() (the executiontime lines, main_cpu.go:1413,1505: fmt.Fprintln(file, time.Since(start))):
// below one second a single unit (ns, µs, ms) with the fraction's trailing zeros dropped, otherwise [h]m]s with up to nine
// fractional digits of the seconds; "0s" for zero.
int gsgph_format_duration(int64_t ns, char *out, int32_t cap)
{
    char buf[40];
    int w = (int)sizeof buf;
    uint64_t u = ns < 0 ? (uint64_t)0 - (uint64_t)ns : (uint64_t)ns;
    auto frac = [&](int prec) {                        // fmtFrac: u / 10^prec, fraction without trailing zeros
        bool print = false;
        for (int i = 0; i < prec; ++i) {
            const int digit = (int)(u % 10);
            print = print || digit != 0;
            if (print) buf[--w] = (char)('0' + digit);
            u /= 10;
        }
        if (print) buf[--w] = '.';
    };
    auto integer = [&](uint64_t v) {                   // fmtInt
        if (v == 0) { buf[--w] = '0'; return; }
        while (v > 0) { buf[--w] = (char)('0' + v % 10); v /= 10; }
    };
    if (u < 1000000000ull) {
        int prec;
        buf[--w] = 's';
        if (u == 0) { buf[--w] = '0'; prec = -1; }
        else if (u < 1000ull) { prec = 0; buf[--w] = 'n'; }
        else if (u < 1000000ull) { prec = 3; buf[--w] = (char)0xB5; buf[--w] = (char)0xC2; }   // U+00B5 micro sign
        else { prec = 6; buf[--w] = 'm'; }
        if (prec >= 0) { frac(prec); integer(u); }
    } else {
        buf[--w] = 's';
        frac(9);                                       // u: whole seconds
        integer(u % 60);
        u /= 60;
        if (u > 0) {
            buf[--w] = 'm';
            integer(u % 60);
            u /= 60;
            if (u > 0) { buf[--w] = 'h'; integer(u); }
        }
    }
    if (ns < 0) buf[--w] = '-';
    const int n = (int)sizeof buf - w;
    if (!out || cap < n + 1) return -1;
    memcpy(out, buf + w, (size_t)n);
    out[n] = 0;
    return n;
}

}  // extern "C"
