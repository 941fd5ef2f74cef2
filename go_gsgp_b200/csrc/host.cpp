// host.cpp -- host-side driver pieces of the hot path (include/gsgp_host.h): RNG, constants,
// initial-population trees and the per-generation plan.  Pure C++, no CUDA.
#include "../../include/gsgp_host.h"
#include "program.hpp"
#include "trees.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct gsgph_rng {
    uint64_t vec[607];
    int tap, feed;
    uint64_t next()
    {
        if (--tap < 0) tap += 607;
        if (--feed < 0) feed += 607;
        const uint64_t x = vec[feed] + vec[tap];
        vec[feed] = x;
        return x;
    }
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
        int32_t v = int31();
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

}  // namespace

extern "C" {

gsgph_rng *gsgph_rng_new(int64_t seed)
{
    gsgph_rng *r = new gsgph_rng();
    uint64_t x = (uint64_t)seed;
    for (int i = 0; i < 607; ++i) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        r->vec[i] = z ^ (z >> 31);
    }
    r->tap = 0; r->feed = 607 - 273;
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
    TreeBuilder tb{r, p, {}};
    int32_t used = 0;
    bool overflow = false;
    auto new_tree = [&](int32_t &off, int32_t &len) {
        const std::vector<int32_t> &t = tb.build(p->max_depth_creation, 0);
        if (used + (int64_t)t.size() > cap) { overflow = true; off = -1; len = 0; return; }
        memcpy(pool + used, t.data(), t.size() * sizeof(int32_t));
        off = used; len = (int32_t)t.size(); used += len;
    };
    for (int32_t k = 0; k < p->population_size; ++k) {
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
        if (overflow) break;
    }
    if (pool_len) *pool_len = used;
    return overflow ? 1 : 0;
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

int gsgph_format_float(double v, char *buf, int32_t cap)
{
    if (std::isnan(v)) return snprintf(buf, cap, "NaN");
    if (std::isinf(v)) return snprintf(buf, cap, v > 0 ? "+Inf" : "-Inf");
    if (v == 0) return snprintf(buf, cap, std::signbit(v) ? "-0" : "0");
    char e[40];
    for (int prec = 0; prec < 17; ++prec) {
        snprintf(e, sizeof e, "%.*e", prec, v);
        if (strtod(e, nullptr) == v) break;
    }
    std::string mant, out;
    const char *q = e;
    if (*q == '-') { out += '-'; ++q; }
    for (; *q && *q != 'e'; ++q) if (*q != '.') mant += *q;
    const int x = atoi(q + 1);
    while (mant.size() > 1 && mant.back() == '0') mant.pop_back();
    if (x < -4 || x >= 6) {
        out += mant[0];
        if (mant.size() > 1) { out += '.'; out.append(mant, 1, std::string::npos); }
        char ex[16]; snprintf(ex, sizeof ex, "e%c%02d", x < 0 ? '-' : '+', std::abs(x));
        out += ex;
    } else if (x < 0) {
        out += "0."; out.append((size_t)(-x - 1), '0'); out += mant;
    } else {
        for (int i = 0; i <= x; ++i) out += i < (int)mant.size() ? mant[i] : '0';
        if ((int)mant.size() > x + 1) { out += '.'; out.append(mant, (size_t)x + 1, std::string::npos); }
    }
    return snprintf(buf, cap, "%s", out.c_str());
}

}  // extern "C"
