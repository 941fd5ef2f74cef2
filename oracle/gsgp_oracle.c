/*
 * gsgp_oracle.c -- CPU ORACLE (test infrastructure, see gsgp_oracle.h).
 *
 * Plain-C restatement of akiross/go-gsgp main_cpu.go's generation-loop path.  PARITY UNPINNED by
 * the reference's own tests (it ships none for this path); pinned by hand-derived KATs and, for the
 * random stream, by known outputs of real Go programs (gsgp_oracle.h).
 * Build with -ffp-contract=off: Go on amd64 never fuses a*b+c, and neither may this file.
 */
#define _GNU_SOURCE
#include "gsgp_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== RNG ============ */

/* Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Go math/rand (v1) seed table (math/rand/rng.go: rngCooked).  Go's stdlib is not in /root/reference and Go is not
 * installed; the table is recomputed from its definition by tools/gen_go_rng_cooked.py and pinned by known outputs of
 * Go programs (tests/test_oracle_kat.py::test_go_math_rand_known_answers). */
static const uint64_t go_rng_cooked[607] = {
#include "go_rng_cooked.inc.h"
};

static int32_t go_seedrand(int32_t x)           /* x[n+1] = 48271 * x[n] mod (2^31 - 1) */
{
    int32_t hi = x / 44488, lo = x % 44488;
    x = 48271 * lo - 3399 * hi;
    return x < 0 ? x + 2147483647 : x;
}

/* ORC_RNG_ALFG: Go's rngSource.Seed (what rand.Seed(cfg.rng_seed), main_cpu.go:1367, runs): vec[607], tap = 0,
 * feed = 607 - 273; the stream is bit-identical to Go's. */
void orc_rng_seed(orc_rng *r, int kind, int64_t seed)
{
    memset(r, 0, sizeof *r);
    r->kind = kind;
    if (kind == ORC_RNG_ALFG) {
        seed %= 2147483647;
        if (seed < 0) seed += 2147483647;
        if (seed == 0) seed = 89482311;
        int32_t x = (int32_t)seed;
        for (int i = -20; i < 607; ++i) {
            x = go_seedrand(x);
            if (i >= 0) {
                uint64_t u = (uint64_t)x << 40;
                x = go_seedrand(x);
                u ^= (uint64_t)x << 20;
                x = go_seedrand(x);
                u ^= (uint64_t)x;
                r->vec[i] = u ^ go_rng_cooked[i];
            }
        }
        r->tap = 0;
        r->feed = 607 - 273;
    } else {
        r->key[0] = (uint32_t)((uint64_t)seed);
        r->key[1] = (uint32_t)((uint64_t)seed >> 32);
    }
}

void orc_rng_seek(orc_rng *r, uint32_t s0, uint32_t s1, uint32_t s2)
{
    r->stream[0] = s0; r->stream[1] = s1; r->stream[2] = s2;
    r->block = 0;
    r->have = 0;
}

uint64_t orc_rng_uint64(orc_rng *r)
{
    r->n_draws++;
    if (r->kind == ORC_RNG_ALFG) {            /* Go rngSource.Uint64 */
        if (--r->tap < 0) r->tap += 607;
        if (--r->feed < 0) r->feed += 607;
        uint64_t x = r->vec[r->feed] + r->vec[r->tap];
        r->vec[r->feed] = x;
        return x;
    }
    if (r->have == 0) {
        uint32_t ctr[4] = { r->block++, r->stream[0], r->stream[1], r->stream[2] };
        orc_philox4x32_10(ctr, r->key, r->out);
        r->have = 2;
    }
    uint64_t v = (r->have == 2) ? ((uint64_t)r->out[0] | ((uint64_t)r->out[1] << 32))
                                : ((uint64_t)r->out[2] | ((uint64_t)r->out[3] << 32));
    r->have--;
    return v;
}

int64_t orc_rng_int63(orc_rng *r) { return (int64_t)(orc_rng_uint64(r) & 0x7fffffffffffffffull); }
int32_t orc_rng_int31(orc_rng *r) { return (int32_t)(orc_rng_int63(r) >> 32); }

double orc_rng_float64(orc_rng *r)           /* Go Rand.Float64: redraw on 1.0 */
{
    for (;;) {
        double f = (double)orc_rng_int63(r) / 9223372036854775808.0;
        if (f != 1.0) return f;
    }
}

int32_t orc_rng_intn(orc_rng *r, int32_t n)  /* Go Rand.Intn -> Int31n for n <= 2^31-1 */
{
    if (n <= 0) { fprintf(stderr, "orc_rng_intn: invalid argument\n"); abort(); }
    if ((n & (n - 1)) == 0) return orc_rng_int31(r) & (n - 1);
    int32_t max = (int32_t)((1u << 31) - 1 - (1u << 31) % (uint32_t)n);
    int32_t v = orc_rng_int31(r);
    while (v > max) v = orc_rng_int31(r);
    return v % n;
}

/* ======================================================================== state ========== */

#define N_FUNC 4   /* main_cpu.go:433-440: + - * / (sqrt commented out) */

typedef void (*pool_fn)(void *arg, int w);

struct orc_state {
    orc_config cfg;
    int32_t nvar, nrow, nrow_test;
    double *set;                 /* row-major (nrow+nrow_test) x (nvar+1): Instance{vars,y} :64-67 */
    int32_t n_var, n_const, n_symbols;
    double *sym_value;           /* Symbol.value per id (:96-103); 0 for functions/variables */
    double *fit, *fit_test, *fit_new, *fit_test_new;                /* :205-208 */
    double **sem_train, **sem_train_new, **sem_test, **sem_test_new; /* :210-213 */
    int32_t index_best;          /* :228 */
    int32_t generation;
    int32_t n_seeds;
    orc_rng rng;
    int32_t **init_tree; int32_t *init_len; uint8_t *seeded;
    orc_plan_entry *plan;
    int32_t *pool; int32_t pool_len, pool_cap;
    int keep_tree_sem;
    double **rt_sem[2];          /* [which][ind] -> nrow+nrow_test values, last generation */
    /* goroutine stand-in: persistent worker pool */
    int n_threads; pthread_t *threads;
    pthread_mutex_t mu; pthread_cond_t cv_work, cv_done;
    pool_fn job; void *job_arg; int job_epoch; int job_pending; int job_n; int shutdown;
    /* scratch */
    double *rt1_train, *rt1_test, *rt2_train, *rt2_test;
};

typedef struct { orc_state *s; int id; } worker_arg;

static void *pool_main(void *p)
{
    worker_arg *wa = (worker_arg *)p;
    orc_state *s = wa->s; int id = wa->id; free(wa);
    int seen = 0;
    pthread_mutex_lock(&s->mu);
    for (;;) {
        while (!s->shutdown && s->job_epoch == seen) pthread_cond_wait(&s->cv_work, &s->mu);
        if (s->shutdown) break;
        seen = s->job_epoch;
        pool_fn fn = s->job; void *arg = s->job_arg; int n = s->job_n;
        pthread_mutex_unlock(&s->mu);
        for (int w = id; w < n; w += s->n_threads + 1) fn(arg, w);
        pthread_mutex_lock(&s->mu);
        if (--s->job_pending == 0) pthread_cond_signal(&s->cv_done);
    }
    pthread_mutex_unlock(&s->mu);
    return NULL;
}

/* run fn(arg, w) for w in [0, n): the reference's "go func(...)" fan-out + wg.Wait()
 * (main_cpu.go:804-820, 955-972).  The caller takes part as worker n_threads. */
static void pool_run(orc_state *s, pool_fn fn, void *arg, int n)
{
    if (s->n_threads == 0 || n <= 1) { for (int w = 0; w < n; ++w) fn(arg, w); return; }
    pthread_mutex_lock(&s->mu);
    s->job = fn; s->job_arg = arg; s->job_n = n; s->job_pending = s->n_threads; s->job_epoch++;
    pthread_cond_broadcast(&s->cv_work);
    pthread_mutex_unlock(&s->mu);
    for (int w = s->n_threads; w < n; w += s->n_threads + 1) fn(arg, w);
    pthread_mutex_lock(&s->mu);
    while (s->job_pending) pthread_cond_wait(&s->cv_done, &s->mu);
    pthread_mutex_unlock(&s->mu);
}

void orc_config_defaults(orc_config *c)      /* main_cpu.go:156-183 */
{
    memset(c, 0, sizeof *c);
    c->population_size = 200; c->init_type = 2; c->p_crossover = 0.6; c->p_mutation = 0.3;
    c->max_depth_creation = 6; c->tournament_size = 4; c->zero_depth = 0;
    c->num_random_constants = 0; c->min_random_constant = -100; c->max_random_constant = 100;
    c->minimization_problem = 1; c->error_measure = ORC_MSE; c->n_workers = 1;
    c->rng_kind = ORC_RNG_ALFG; c->rng_seed = 1;
}

static void *xcalloc(size_t n, size_t sz)
{
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}

/* init_tables main_cpu.go:1257-1284 */
orc_state *orc_new(const orc_config *cfg, int32_t nvar, int32_t nrow, int32_t nrow_test,
                   const double *rowmajor)
{
    orc_state *s = (orc_state *)xcalloc(1, sizeof *s);
    s->cfg = *cfg; s->nvar = nvar; s->nrow = nrow; s->nrow_test = nrow_test;
    size_t n = (size_t)(nrow + nrow_test) * (size_t)(nvar + 1);
    s->set = (double *)xcalloc(n, sizeof(double));
    memcpy(s->set, rowmajor, n * sizeof(double));
    int P = cfg->population_size;
    s->fit = xcalloc(P, 8); s->fit_test = xcalloc(P, 8);
    s->fit_new = xcalloc(P, 8); s->fit_test_new = xcalloc(P, 8);
    s->sem_train = xcalloc(P, sizeof(double *)); s->sem_train_new = xcalloc(P, sizeof(double *));
    s->sem_test = xcalloc(P, sizeof(double *));  s->sem_test_new = xcalloc(P, sizeof(double *));
    for (int i = 0; i < P; ++i) {
        s->sem_train[i] = xcalloc(nrow, 8);     s->sem_train_new[i] = xcalloc(nrow, 8);
        s->sem_test[i] = xcalloc(nrow_test, 8); s->sem_test_new[i] = xcalloc(nrow_test, 8);
    }
    s->init_tree = xcalloc(P, sizeof(int32_t *)); s->init_len = xcalloc(P, 4);
    s->seeded = xcalloc(P, 1);
    s->plan = xcalloc(P, sizeof(orc_plan_entry));
    s->rt1_train = xcalloc(nrow, 8); s->rt1_test = xcalloc(nrow_test, 8);
    s->rt2_train = xcalloc(nrow, 8); s->rt2_test = xcalloc(nrow_test, 8);
    orc_rng_seed(&s->rng, cfg->rng_kind, cfg->rng_seed);     /* rand.Seed :1367 */
    pthread_mutex_init(&s->mu, NULL);
    pthread_cond_init(&s->cv_work, NULL); pthread_cond_init(&s->cv_done, NULL);
    if (cfg->n_workers > 1) {
        s->n_threads = cfg->n_workers - 1;
        s->threads = xcalloc(s->n_threads, sizeof(pthread_t));
        for (int t = 0; t < s->n_threads; ++t) {
            worker_arg *wa = xcalloc(1, sizeof *wa); wa->s = s; wa->id = t;
            pthread_create(&s->threads[t], NULL, pool_main, wa);
        }
    }
    return s;
}

void orc_free(orc_state *s)
{
    if (!s) return;
    if (s->n_threads) {
        pthread_mutex_lock(&s->mu); s->shutdown = 1;
        pthread_cond_broadcast(&s->cv_work); pthread_mutex_unlock(&s->mu);
        for (int t = 0; t < s->n_threads; ++t) pthread_join(s->threads[t], NULL);
        free(s->threads);
    }
    int P = s->cfg.population_size;
    for (int i = 0; i < P; ++i) {
        free(s->sem_train[i]); free(s->sem_train_new[i]);
        free(s->sem_test[i]);  free(s->sem_test_new[i]);
        free(s->init_tree[i]);
        for (int w = 0; w < 2; ++w) if (s->rt_sem[w]) free(s->rt_sem[w][i]);
    }
    free(s->rt_sem[0]); free(s->rt_sem[1]);
    free(s->sem_train); free(s->sem_train_new); free(s->sem_test); free(s->sem_test_new);
    free(s->fit); free(s->fit_test); free(s->fit_new); free(s->fit_test_new);
    free(s->init_tree); free(s->init_len); free(s->seeded); free(s->plan); free(s->pool);
    free(s->rt1_train); free(s->rt1_test); free(s->rt2_train); free(s->rt2_test);
    free(s->sym_value); free(s->set);
    free(s);
}

orc_rng *orc_get_rng(orc_state *s) { return &s->rng; }

/* create_T_F main_cpu.go:426-455: ids 0..3 functions, 4..4+V-1 variables, then constants, each
 * constant drawing one Float64 (:451). */
void orc_create_T_F(orc_state *s)
{
    s->n_var = s->nvar;
    s->n_const = s->cfg.num_random_constants;                 /* :1299 */
    s->n_symbols = N_FUNC + s->n_var + s->n_const;
    free(s->sym_value);
    s->sym_value = xcalloc(s->n_symbols, 8);
    for (int i = N_FUNC + s->n_var; i < s->n_symbols; ++i)
        s->sym_value[i] = s->cfg.min_random_constant +
            orc_rng_float64(&s->rng) * (s->cfg.max_random_constant - s->cfg.min_random_constant);
}

static int32_t choose_function(orc_state *s) { return orc_rng_intn(&s->rng, N_FUNC); } /* :458-460 */

static int32_t choose_terminal(orc_state *s)                                           /* :466-474 */
{
    if (s->n_const == 0) return N_FUNC + orc_rng_intn(&s->rng, s->n_var);
    if (orc_rng_float64(&s->rng) < 0.7) return N_FUNC + orc_rng_intn(&s->rng, s->n_var);
    return N_FUNC + s->n_var + orc_rng_intn(&s->rng, s->n_const);
}

/* growable int32 buffer used as the Go slice being appended to */
typedef struct { int32_t *v; int32_t len, cap; int fixed; } ibuf;
static void ib_push(ibuf *b, int32_t x)
{
    if (b->len == b->cap) {
        if (b->fixed) { fprintf(stderr, "oracle: tree buffer overflow\n"); abort(); }
        b->cap = b->cap ? b->cap * 2 : 64;
        b->v = realloc(b->v, (size_t)b->cap * 4);
    }
    b->v[b->len++] = x;
}

/* Shared builder for the pointer-prefix array format (main_cpu.go:609-639 and, for the Node
 * builders :568-607/:642-672 followed by tree_to_array :675-693, which re-encodes the Node tree
 * depth-first into the very same layout; the RNG draw order is identical, so the arrays are built
 * directly).  mode 0 = create_grow_tree_arrays (Intn(2)==0 -> terminal, :626),
 *           mode 1 = create_grow_tree Node builder (Intn(2)==0 -> function, :588),
 *           mode 2 = create_full_tree (:642-672).
 * Function node = [op, idx_child1, idx_child2, child1..., child2...], absolute indices. */
static void build_tree(orc_state *s, ibuf *b, int32_t depth, int32_t max_depth, int mode)
{
    int is_func;
    if (mode == 2) {
        if (depth == 0 && depth < max_depth) is_func = 1;        /* :643 */
        else if (depth == max_depth) is_func = 0;                /* :655 */
        else is_func = 1;                                        /* :662 */
    } else {
        if (depth == 0 && !s->cfg.zero_depth) is_func = 1;       /* :569 / :610 */
        else if (depth == max_depth) is_func = 0;                /* :581 / :623 */
        else {
            int32_t c = orc_rng_intn(&s->rng, 2);                /* :588 / :626 */
            is_func = (mode == 0) ? (c != 0) : (c == 0);
        }
    }
    if (!is_func) { ib_push(b, choose_terminal(s)); return; }
    int32_t op = choose_function(s);
    int32_t pos = b->len;
    ib_push(b, op); ib_push(b, 0); ib_push(b, 0);                /* arity 2: id + 2 child slots */
    for (int c = 1; c <= 2; ++c) {
        b->v[pos + c] = b->len;                                  /* tree[c] = len(tree)+base :617 */
        build_tree(s, b, depth + 1, max_depth, mode);
    }
}

int32_t orc_create_grow_tree_arrays(orc_state *s, int32_t depth, int32_t max_depth,
                                    int32_t base_index, int32_t *out, int32_t cap)
{
    (void)base_index;    /* the reference's base_index only makes child indices absolute; the
                            builder above writes into one buffer, so they already are */
    ibuf b = { out, 0, cap, 1 };
    build_tree(s, &b, depth, max_depth, 0);
    return b.len;
}

static void set_init_tree(orc_state *s, int32_t ind, const int32_t *t, int32_t len)
{
    free(s->init_tree[ind]);
    s->init_tree[ind] = xcalloc(len, 4);
    memcpy(s->init_tree[ind], t, (size_t)len * 4);
    s->init_len[ind] = len;
}

void orc_set_initial_tree(orc_state *s, int32_t ind, const int32_t *tree, int32_t len)
{
    set_init_tree(s, ind, tree, len);
    s->seeded[ind] = 0;
}

int orc_seed_semantic(orc_state *s, int32_t ind, const double *sem)   /* :1383-1388 */
{
    if (ind < 0 || ind >= s->cfg.population_size) return -1;
    memcpy(s->sem_train[ind], sem, (size_t)s->nrow * 8);
    memcpy(s->sem_test[ind], sem + s->nrow, (size_t)s->nrow_test * 8);
    s->seeded[ind] = 1;
    return 0;
}

static void add_individual(orc_state *s, int32_t *num_ind, int32_t max_depth, int mode)
{
    ibuf b = { NULL, 0, 0, 0 };
    build_tree(s, &b, 0, max_depth, mode);
    set_init_tree(s, *num_ind, b.v, b.len);
    free(b.v);
    (*num_ind)++;
}

/* initialize_population :556-565, create_grow_pop :477-483, create_full_pop :486-492,
 * create_ramped_pop :495-538.  Seeded individuals occupy 0..n_seeds-1 (:549,:1383-1388). */
void orc_initialize_population(orc_state *s, int32_t n_seeds)
{
    int32_t P = s->cfg.population_size, d = s->cfg.max_depth_creation;
    int32_t num_ind = n_seeds;
    s->n_seeds = n_seeds;
    if (s->cfg.init_type == 0) { while (num_ind < P) add_individual(s, &num_ind, d, 1); return; }
    if (s->cfg.init_type == 1) { while (num_ind < P) add_individual(s, &num_ind, d, 2); return; }
    int32_t sub_pop, r, min_depth;
    if (!s->cfg.zero_depth) { sub_pop = (P - num_ind) / d; r = (P - num_ind) % d; min_depth = 1; }
    else { sub_pop = (P - num_ind) / (d + 1); r = (P - num_ind) % (d + 1); min_depth = 0; }
    for (int32_t j = d; j >= min_depth; --j) {
        int32_t n = (j < d) ? sub_pop : sub_pop + r;
        int32_t n_full = (int32_t)ceil((double)n * 0.5), n_grow = (int32_t)floor((double)n * 0.5);
        for (int32_t k = 0; k < n_full; ++k) add_individual(s, &num_ind, j, 2);
        for (int32_t k = 0; k < n_grow; ++k) add_individual(s, &num_ind, j, 1);
    }
}

/* ======================================================================== evaluation ===== */

static inline double protected_division(double num, double den)   /* :736-741 */
{
    if (den == 0) return 1;
    return num / den;
}

/* eval_arrays :755-775 with terminal_value :745-753.  The reference dispatches on the symbol
 * NAME; ids 0..3 are "+","-","*","/" by construction (:434-437).  Left operand first. */
double orc_eval_arrays(const orc_state *s, const int32_t *tree, int32_t start, int32_t i)
{
    int32_t id = tree[start];
    switch (id) {
    case 0: { double a = orc_eval_arrays(s, tree, tree[start + 1], i);
              double b = orc_eval_arrays(s, tree, tree[start + 2], i); return a + b; }
    case 1: { double a = orc_eval_arrays(s, tree, tree[start + 1], i);
              double b = orc_eval_arrays(s, tree, tree[start + 2], i); return a - b; }
    case 2: { double a = orc_eval_arrays(s, tree, tree[start + 1], i);
              double b = orc_eval_arrays(s, tree, tree[start + 2], i); return a * b; }
    case 3: { double a = orc_eval_arrays(s, tree, tree[start + 1], i);
              double b = orc_eval_arrays(s, tree, tree[start + 2], i);
              return protected_division(a, b); }
    default:
        if (id >= N_FUNC && id < N_FUNC + s->n_var)
            return s->set[(size_t)i * (size_t)(s->nvar + 1) + (size_t)(id - N_FUNC)];
        return s->sym_value[id];
    }
}

typedef struct { const orc_state *s; const int32_t *tree; int32_t size, offs, block; double *val; } sea_job;

static void sea_worker(void *p, int w)
{
    sea_job *j = (sea_job *)p;
    int32_t start = j->block * w, end = j->block * (w + 1);
    if (end > j->size) end = j->size;                            /* :810-812 */
    for (int32_t i = j->offs + start; i < j->offs + end; ++i)
        j->val[i - j->offs] = orc_eval_arrays(j->s, j->tree, 0, i);
}

/* semantic_evaluate_array :797-827 (the caller provides val instead of make()) */
void orc_semantic_evaluate_array(const orc_state *s, const int32_t *tree,
                                 int32_t sem_size, int32_t sem_offs, double *val)
{
    int W = s->cfg.n_workers;
    if (W > 1) {
        sea_job j = { s, tree, sem_size, sem_offs, (sem_size + W - 1) / W, val };
        pool_run((orc_state *)s, sea_worker, &j, W);
    } else {
        for (int32_t i = sem_offs; i < sem_size + sem_offs; ++i)
            val[i - sem_offs] = orc_eval_arrays(s, tree, 0, i);
    }
}

/* Go's math.Exp returns 1+x for |x| < 2^-28 (stdlib exp.go, NearZero); everything else is libm's
 * correctly-specialised exp.  The rule matters only for exp64's "r == 1 && v != 0" test below. */
static int g_exp_kind = 0;
void orc_set_exp_kind(int kind) { g_exp_kind = kind; }

/* Go's portable math.Exp (stdlib exp.go: the FreeBSD e_exp.c port -- what Go runs where it has no assembly routine;
 * amd64 / arm64 / s390x have their own), restated from its published algorithm: k = round(x / ln 2), r = x - k ln 2 in
 * two pieces, a degree-5 rational correction, ldexp.  It differs from glibc's exp in the last bit for a few percent of
 * the arguments: the second exp (ORC_EXP_GO_PORTABLE) exists to show what that last bit does to a run -- nothing,
 * tests/test_oracle_kat.py::test_run_does_not_depend_on_the_last_bit_of_exp. */
static double go_exp_portable(double x)
{
    const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10, Log2e = 1.44269504088896338700e+00;
    const double Overflow = 7.09782712893383973096e+02, Underflow = -7.45133219101941108420e+02;
    const double P1 = 1.66666666666666657415e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    if (isnan(x) || (isinf(x) && x > 0)) return x;
    if (isinf(x)) return 0;
    if (x > Overflow) return INFINITY;
    if (x < Underflow) return 0;
    if (x > -0x1p-28 && x < 0x1p-28) return 1 + x;
    int k = 0;
    if (x < 0) k = (int)(Log2e * x - 0.5);
    else if (x > 0) k = (int)(Log2e * x + 0.5);
    double hi = x - (double)k * Ln2Hi, lo = (double)k * Ln2Lo;
    double r = hi - lo, t = r * r;
    double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y = 1 - ((lo - (r * c) / (2 - c)) - hi);
    return ldexp(y, k);
}

static inline double go_exp(double x)
{
    if (g_exp_kind == ORC_EXP_GO_PORTABLE) return go_exp_portable(x);
    if (x > -0x1p-28 && x < 0x1p-28) return 1 + x;
    return exp(x);
}

double orc_exp64(double v)                                       /* :50-61 */
{
    double r = go_exp(v);
    if ((r == 0 && v > 0) || (isinf(r) && r > 0)) return DBL_MAX;
    if (r == 1 && v != 0) return 0;
    return r;
}

double orc_sigmoid(double r) { return 1 / (1 + orc_exp64(-r)); }  /* :894 */

static inline double dist(int measure, double a, double b)       /* :290-292 */
{
    switch (measure) {
    case ORC_MAE: return fabs(a - b);
    case ORC_MRE: return fabs(a - b) / a;
    default:      return (a - b) * (a - b);
    }
}

static inline double post_error(int measure, double d)           /* :235, :1326-1328 */
{
    return measure == ORC_RMSE ? sqrt(d) : d;
}

typedef struct { const orc_state *s; const double *sem; int32_t size, offs, block; double *par; } fit_job;

static void fit_worker(void *p, int w)
{
    fit_job *j = (fit_job *)p;
    const orc_state *s = j->s;
    int32_t start = j->block * w, end = j->block * (w + 1);
    if (end > j->size) end = j->size;
    size_t stride = (size_t)s->nvar + 1;
    double acc = 0;
    for (int32_t i = j->offs + start; i < j->offs + end; ++i)
        acc += dist(s->cfg.error_measure, s->set[(size_t)i * stride + s->nvar], j->sem[i - j->offs]);
    j->par[w] = acc;
}

/* fitness_of_semantic_train_nls :950-985 and _test_nls :1106-1141 share this body: sequential
 * sum when W == 1, W blocked partials added in worker order otherwise (:973-977). */
static double fitness_nls(const orc_state *s, const double *sem, int32_t size, int32_t offs)
{
    int W = s->cfg.n_workers;
    double d = 0;
    if (W > 1) {
        double *par = (double *)alloca((size_t)W * 8);
        memset(par, 0, (size_t)W * 8);
        fit_job j = { s, sem, size, offs, (size + W - 1) / W, par };
        pool_run((orc_state *)s, fit_worker, &j, W);
        d = par[0];
        for (int i = 1; i < W; ++i) d += par[i];
    } else {
        size_t stride = (size_t)s->nvar + 1;
        for (int32_t i = offs; i < offs + size; ++i)
            d += dist(s->cfg.error_measure, s->set[(size_t)i * stride + s->nvar], sem[i - offs]);
    }
    return post_error(s->cfg.error_measure, d / (double)size);
}

double orc_fitness_train(const orc_state *s, const double *sem) { return fitness_nls(s, sem, s->nrow, 0); }
double orc_fitness_test(const orc_state *s, const double *sem)  { return fitness_nls(s, sem, s->nrow_test, s->nrow); }

/* ---- linear scaling ------------------------------------------------------------------------ */
typedef struct { const orc_state *s; const double *sem; int32_t size, offs, block;
                 double *s_out, *s_tar, *s_oxo, *s_oxt; } ls_sum_job;

static void ls_sum_worker(void *p, int w)                        /* :1009-1024 */
{
    ls_sum_job *j = (ls_sum_job *)p;
    const orc_state *s = j->s;
    int32_t start = j->block * w, end = j->block * (w + 1);
    if (end > j->size) end = j->size;
    size_t stride = (size_t)s->nvar + 1;
    double so = 0, st = 0, soo = 0, sot = 0;
    for (int32_t i = j->offs + start; i < j->offs + end; ++i) {
        double t = s->set[(size_t)i * stride + s->nvar], y = j->sem[i - j->offs];
        so += y; st += t; soo += y * y; sot += y * t;
    }
    j->s_out[w] = so; j->s_tar[w] = st; j->s_oxo[w] = soo; j->s_oxt[w] = sot;
}

typedef struct { const orc_state *s; const double *sem; int32_t size, offs, block; double a, b; double *par; } ls_err_job;

static void ls_err_worker(void *p, int w)                        /* :1059-1071, :1153-1165 */
{
    ls_err_job *j = (ls_err_job *)p;
    const orc_state *s = j->s;
    int32_t start = j->block * w, end = j->block * (w + 1);
    if (end > j->size) end = j->size;
    size_t stride = (size_t)s->nvar + 1;
    double acc = 0;
    for (int32_t i = j->offs + start; i < j->offs + end; ++i)
        acc += dist(s->cfg.error_measure, s->set[(size_t)i * stride + s->nvar], j->a + j->b * j->sem[i - j->offs]);
    j->par[w] = acc;
}

static double ls_error(const orc_state *s, const double *sem, int32_t size, int32_t offs, double a, double b)
{
    int W = s->cfg.n_workers;
    double d = 0;
    if (W > 1) {
        double *par = (double *)alloca((size_t)W * 8);
        memset(par, 0, (size_t)W * 8);
        ls_err_job j = { s, sem, size, offs, (size + W - 1) / W, a, b, par };
        pool_run((orc_state *)s, ls_err_worker, &j, W);
        d = par[0];
        for (int i = 1; i < W; ++i) d += par[i];
    } else {
        size_t stride = (size_t)s->nvar + 1;
        for (int32_t i = offs; i < offs + size; ++i)
            d += dist(s->cfg.error_measure, s->set[(size_t)i * stride + s->nvar], a + b * sem[i - offs]);
    }
    return post_error(s->cfg.error_measure, d / (double)size);
}

/* fitness_of_semantic_train_ls :991-1104.  W > 1: four blocked sums, slope from the sum form (:1039-1050);
 * W == 1: means first, then the centred products (:1076-1095).  b = 0 when den == 0. */
double orc_fitness_train_ls(const orc_state *s, const double *sem, double *a_out, double *b_out)
{
    const int32_t size = s->nrow, offs = 0;
    int W = s->cfg.n_workers;
    size_t stride = (size_t)s->nvar + 1;
    double a, b;
    if (W > 1) {
        double *buf = (double *)alloca((size_t)W * 32);
        memset(buf, 0, (size_t)W * 32);
        ls_sum_job j = { s, sem, size, offs, (size + W - 1) / W, buf, buf + W, buf + 2 * W, buf + 3 * W };
        pool_run((orc_state *)s, ls_sum_worker, &j, W);
        double tot_out = j.s_out[0], tot_tar = j.s_tar[0], tot_oxo = j.s_oxo[0], tot_oxt = j.s_oxt[0];
        for (int i = 1; i < W; ++i) { tot_out += j.s_out[i]; tot_tar += j.s_tar[i]; tot_oxo += j.s_oxo[i]; tot_oxt += j.s_oxt[i]; }
        double avg_out = tot_out / (double)size, avg_tar = tot_tar / (double)size;
        double num = tot_oxt - tot_tar * avg_out - tot_out * avg_tar + (double)size * avg_out * avg_tar;
        double den = tot_oxo - 2.0 * tot_out * avg_out + (double)size * avg_out * avg_out;
        b = (den != 0) ? num / den : 0;
        a = avg_tar - b * avg_out;
    } else {
        double avg_out = 0, avg_tar = 0;
        for (int32_t i = offs; i < size + offs; ++i) { avg_out += sem[i - offs]; avg_tar += s->set[(size_t)i * stride + s->nvar]; }
        avg_out /= (double)size; avg_tar /= (double)size;
        double num = 0, den = 0;
        for (int32_t i = offs; i < offs + size; ++i) {
            double odiff = sem[i - offs] - avg_out;
            num += (s->set[(size_t)i * stride + s->nvar] - avg_tar) * odiff;
            den += odiff * odiff;
        }
        b = (den != 0) ? num / den : 0;
        a = avg_tar - b * avg_out;
    }
    if (a_out) *a_out = a;
    if (b_out) *b_out = b;
    return ls_error(s, sem, size, offs, a, b);
}

/* fitness_of_semantic_test_ls :1142-1177: the train a, b applied to the test cases */
double orc_fitness_test_ls(const orc_state *s, const double *sem, double a, double b)
{
    return ls_error(s, sem, s->nrow_test, s->nrow, a, b);
}

/* fitness_of_semantic_train / _test as main() binds them (:1311-1317) */
static void fitness_pair(const orc_state *s, const double *sem_train, const double *sem_test, double *fit, double *fit_test)
{
    if (s->cfg.use_linear_scaling) {
        double a, b;
        *fit = orc_fitness_train_ls(s, sem_train, &a, &b);
        *fit_test = orc_fitness_test_ls(s, sem_test, a, b);
    } else {
        *fit = orc_fitness_train(s, sem_train);
        *fit_test = orc_fitness_test(s, sem_test);
    }
}

/* evaluate :780-795 */
void orc_evaluate(orc_state *s)
{
    for (int i = 0; i < s->cfg.population_size; ++i) {
        if (!s->seeded[i] && s->init_tree[i]) {
            orc_semantic_evaluate_array(s, s->init_tree[i], s->nrow, 0, s->sem_train[i]);
            orc_semantic_evaluate_array(s, s->init_tree[i], s->nrow_test, s->nrow, s->sem_test[i]);
        }
        fitness_pair(s, s->sem_train[i], s->sem_test[i], &s->fit[i], &s->fit_test[i]);
    }
    s->index_best = orc_best_individual(s);                      /* :1403 */
}

/* ======================================================================== selection ====== */

int orc_better(const orc_state *s, double f1, double f2)         /* :1206-1212 */
{
    return s->cfg.minimization_problem ? (f1 < f2) : (f1 > f2);
}

int32_t orc_tournament_selection(orc_state *s)                   /* :830-840 */
{
    int32_t best_index = orc_rng_intn(&s->rng, s->cfg.population_size);
    for (int i = 1; i < s->cfg.tournament_size; ++i) {
        int32_t next = orc_rng_intn(&s->rng, s->cfg.population_size);
        if (orc_better(s, s->fit[next], s->fit[best_index])) best_index = next;
    }
    return best_index;
}

int32_t orc_best_individual(const orc_state *s)                  /* :1180-1188 */
{
    int32_t best_index = 0;
    for (int i = 1; i < s->cfg.population_size; ++i)
        if (orc_better(s, s->fit[i], s->fit[best_index])) best_index = i;
    return best_index;
}

/* ======================================================================== operators ====== */

void orc_gs_crossover_vec(const double *p1, const double *p2, const double *r, double *out, int64_t n)
{
    for (int64_t j = 0; j < n; ++j) {                            /* :893-896 */
        double sigmoid = 1 / (1 + orc_exp64(-r[j]));
        out[j] = p1[j] * sigmoid + p2[j] * (1 - sigmoid);
    }
}

void orc_gs_mutation_vec(const double *t, const double *r1, const double *r2, double ms, double *out, int64_t n)
{
    for (int64_t j = 0; j < n; ++j) {                            /* :932-936 */
        double sigmoid1 = 1 / (1 + orc_exp64(-r1[j]));
        double sigmoid2 = 1 / (1 + orc_exp64(-r2[j]));
        out[j] = t[j] + ms * (sigmoid1 - sigmoid2);
    }
}

static void keep_rt(orc_state *s, int which, int32_t ind, const double *tr, const double *te)
{
    if (!s->keep_tree_sem) return;
    if (!s->rt_sem[which]) s->rt_sem[which] = xcalloc(s->cfg.population_size, sizeof(double *));
    if (!s->rt_sem[which][ind]) s->rt_sem[which][ind] = xcalloc((size_t)s->nrow + s->nrow_test, 8);
    memcpy(s->rt_sem[which][ind], tr, (size_t)s->nrow * 8);
    memcpy(s->rt_sem[which][ind] + s->nrow, te, (size_t)s->nrow_test * 8);
}

static void copy_individual(orc_state *s, int32_t dst, int32_t src)  /* :853-860, :906-912 */
{
    memcpy(s->sem_train_new[dst], s->sem_train[src], (size_t)s->nrow * 8);
    memcpy(s->sem_test_new[dst], s->sem_test[src], (size_t)s->nrow_test * 8);
    s->fit_new[dst] = s->fit[src];
    s->fit_test_new[dst] = s->fit_test[src];
}

/* the arithmetic of geometric_semantic_crossover :887-903 once tree and parents are known */
static void xo_apply(orc_state *s, int32_t i, int32_t p1, int32_t p2, const int32_t *rt1)
{
    orc_semantic_evaluate_array(s, rt1, s->nrow, 0, s->rt1_train);
    orc_semantic_evaluate_array(s, rt1, s->nrow_test, s->nrow, s->rt1_test);
    orc_gs_crossover_vec(s->sem_train[p1], s->sem_train[p2], s->rt1_train, s->sem_train_new[i], s->nrow);
    orc_gs_crossover_vec(s->sem_test[p1], s->sem_test[p2], s->rt1_test, s->sem_test_new[i], s->nrow_test);
    fitness_pair(s, s->sem_train_new[i], s->sem_test_new[i], &s->fit_new[i], &s->fit_test_new[i]);   /* :897,:903 */
    keep_rt(s, 0, i, s->rt1_train, s->rt1_test);
}

/* the arithmetic of geometric_semantic_mutation :924-944 (after reproduction copied the parent) */
static void mut_apply(orc_state *s, int32_t i, double ms, const int32_t *rt1, const int32_t *rt2)
{
    orc_semantic_evaluate_array(s, rt1, s->nrow, 0, s->rt1_train);
    orc_semantic_evaluate_array(s, rt1, s->nrow_test, s->nrow, s->rt1_test);
    orc_semantic_evaluate_array(s, rt2, s->nrow, 0, s->rt2_train);
    orc_semantic_evaluate_array(s, rt2, s->nrow_test, s->nrow, s->rt2_test);
    orc_gs_mutation_vec(s->sem_train_new[i], s->rt1_train, s->rt2_train, ms, s->sem_train_new[i], s->nrow);
    orc_gs_mutation_vec(s->sem_test_new[i], s->rt1_test, s->rt2_test, ms, s->sem_test_new[i], s->nrow_test);
    fitness_pair(s, s->sem_train_new[i], s->sem_test_new[i], &s->fit_new[i], &s->fit_test_new[i]);   /* :937,:944 */
    keep_rt(s, 0, i, s->rt1_train, s->rt1_test);
    keep_rt(s, 1, i, s->rt2_train, s->rt2_test);
}

static int32_t pool_new_tree(orc_state *s, int32_t *len_out)
{
    ibuf b = { s->pool, s->pool_len, s->pool_cap, 0 };
    int32_t off = b.len;
    /* trees are built with indices relative to their own start, as the reference does with
     * base_index = 0 (:879,:921-922): build into a scratch and append */
    ibuf t = { NULL, 0, 0, 0 };
    build_tree(s, &t, 0, s->cfg.max_depth_creation, 0);
    for (int32_t j = 0; j < t.len; ++j) ib_push(&b, t.v[j]);
    *len_out = t.len;
    free(t.v);
    s->pool = b.v; s->pool_len = b.len; s->pool_cap = b.cap;
    return off;
}

static void update_tables(orc_state *s)                           /* :1191-1197 */
{
    double *t;
    t = s->fit; s->fit = s->fit_new; s->fit_new = t;
    t = s->fit_test; s->fit_test = s->fit_test_new; s->fit_test_new = t;
    double **q;
    q = s->sem_train; s->sem_train = s->sem_train_new; s->sem_train_new = q;
    q = s->sem_test; s->sem_test = s->sem_test_new; s->sem_test_new = q;
}

/* main GP cycle body :1447-1470 + :1490-1492, with the per-individual decisions recorded. */
void orc_generation(orc_state *s)
{
    int32_t P = s->cfg.population_size;
    s->pool_len = 0;
    for (int32_t k = 0; k < P; ++k) {
        orc_plan_entry *e = &s->plan[k];
        memset(e, 0, sizeof *e);
        e->p1 = e->p2 = -1; e->tree0_off = e->tree1_off = -1;
        if (s->rng.kind == ORC_RNG_PHILOX) orc_rng_seek(&s->rng, (uint32_t)(s->generation + 1), (uint32_t)k, 0);
        double rand_num = orc_rng_float64(&s->rng);               /* :1451 */
        e->elite = (k == s->index_best);
        if (rand_num < s->cfg.p_crossover) {                      /* :1453 */
            e->op = ORC_OP_XO;
            if (k != s->index_best) {                             /* :877 */
                e->tree0_off = pool_new_tree(s, &e->tree0_len);   /* :879 */
                e->p1 = orc_tournament_selection(s);              /* :881 */
                e->p2 = orc_tournament_selection(s);              /* :882 */
                xo_apply(s, k, e->p1, e->p2, s->pool + e->tree0_off);
            } else {
                e->p1 = k;
                copy_individual(s, k, k);                         /* :906-912 */
            }
        } else if (rand_num < s->cfg.p_crossover + s->cfg.p_mutation) {   /* :1459 */
            e->op = ORC_OP_MUT;
            e->p1 = (k != s->index_best) ? orc_tournament_selection(s) : k;  /* :847-850 */
            copy_individual(s, k, e->p1);
            if (k != s->index_best) {                             /* :918 */
                e->ms = orc_rng_float64(&s->rng);                 /* :919 */
                e->tree0_off = pool_new_tree(s, &e->tree0_len);   /* :921 */
                e->tree1_off = pool_new_tree(s, &e->tree1_len);   /* :922 */
                mut_apply(s, k, e->ms, s->pool + e->tree0_off, s->pool + e->tree1_off);
            }
        } else {                                                  /* :1467-1469 */
            e->op = ORC_OP_REPR;
            e->p1 = (k != s->index_best) ? orc_tournament_selection(s) : k;
            copy_individual(s, k, e->p1);
        }
    }
    update_tables(s);                                             /* :1490 */
    s->index_best = orc_best_individual(s);                       /* :1492 */
    s->generation++;
}

/* operators_effects.go's single step (its main, :1246-1256): every individual gets the same operator once, parents
 * drawn UNIFORMLY (random_selection :783-786), no tournament and no operator draw.  test_crossover != 0:
 * geometric_semantic_crossover(k) for every k (:800-806: tree, p1 = Intn(P), p2 = Intn(P), in that order); else
 * reproduction(k) -- a plain self copy there (:790-797: no draw) -- then geometric_semantic_mutation(k) (:836-842:
 * ms = Float64(), rt1, rt2).  index_best is never assigned in that program (its zero value, individual 0, plays the
 * elite: :801,:837).  Records the plan and tree pool like orc_generation; fit / fit_test then hold fit_new / fit_test_new
 * of the step (the program prints fit[k] fit_new[k], :1258-1261). */
void orc_effects_step(orc_state *s, int test_crossover)
{
    int32_t P = s->cfg.population_size;
    s->pool_len = 0;
    s->index_best = 0;
    for (int32_t k = 0; k < P; ++k) {
        orc_plan_entry *e = &s->plan[k];
        memset(e, 0, sizeof *e);
        e->p1 = k; e->p2 = -1; e->tree0_off = e->tree1_off = -1;
        e->elite = (k == s->index_best);
        if (test_crossover) {
            e->op = ORC_OP_XO;
            if (k != s->index_best) {
                e->tree0_off = pool_new_tree(s, &e->tree0_len);           /* :803 */
                e->p1 = orc_rng_intn(&s->rng, P);                         /* :805 random_selection */
                e->p2 = orc_rng_intn(&s->rng, P);                         /* :806 */
                xo_apply(s, k, e->p1, e->p2, s->pool + e->tree0_off);
            } else {
                copy_individual(s, k, k);
            }
        } else {
            e->op = ORC_OP_MUT;
            copy_individual(s, k, k);                                     /* reproduction(i): self copy, :790-797 */
            if (k != s->index_best) {
                e->ms = orc_rng_float64(&s->rng);                         /* :838 */
                e->tree0_off = pool_new_tree(s, &e->tree0_len);           /* :840 */
                e->tree1_off = pool_new_tree(s, &e->tree1_len);           /* :841 */
                mut_apply(s, k, e->ms, s->pool + e->tree0_off, s->pool + e->tree1_off);
            }
        }
    }
    update_tables(s);
    s->generation++;
}

void orc_generation_replay(orc_state *s, const orc_plan_entry *plan, const int32_t *tree_pool)
{
    int32_t P = s->cfg.population_size;
    for (int32_t k = 0; k < P; ++k) {
        const orc_plan_entry *e = &plan[k];
        if (e->op == ORC_OP_XO && !e->elite) {
            xo_apply(s, k, e->p1, e->p2, tree_pool + e->tree0_off);
        } else {
            copy_individual(s, k, e->p1);
            if (e->op == ORC_OP_MUT && !e->elite)
                mut_apply(s, k, e->ms, tree_pool + e->tree0_off, tree_pool + e->tree1_off);
        }
    }
    if (plan != s->plan) memcpy(s->plan, plan, (size_t)P * sizeof *plan);
    update_tables(s);
    s->index_best = orc_best_individual(s);
    s->generation++;
}

/* ======================================================================== accessors ====== */

void orc_keep_tree_semantics(orc_state *s, int flag) { s->keep_tree_sem = flag; }
/* test hooks: drive the driver loop from a poked state (selection edge cases) */
void orc_test_set_index_best(orc_state *s, int32_t b) { s->index_best = b; }
void orc_test_set_generation(orc_state *s, int32_t g) { s->generation = g; }
int32_t orc_generation_index(const orc_state *s) { return s->generation; }
int32_t orc_index_best(const orc_state *s) { return s->index_best; }
const double *orc_fit(const orc_state *s) { return s->fit; }
const double *orc_fit_test(const orc_state *s) { return s->fit_test; }
const double *orc_sem_train(const orc_state *s, int32_t ind) { return s->sem_train[ind]; }
const double *orc_sem_test(const orc_state *s, int32_t ind) { return s->sem_test[ind]; }
const orc_plan_entry *orc_last_plan(const orc_state *s) { return s->plan; }
const int32_t *orc_last_tree_pool(const orc_state *s, int32_t *len) { if (len) *len = s->pool_len; return s->pool; }
const int32_t *orc_initial_tree(const orc_state *s, int32_t ind, int32_t *len)
{ if (len) *len = s->init_len[ind]; return s->init_tree[ind]; }
const double *orc_symbol_values(const orc_state *s, int32_t *n) { if (n) *n = s->n_symbols; return s->sym_value; }
const double *orc_last_tree_semantic(const orc_state *s, int32_t ind, int32_t which)
{ return (s->rt_sem[which] ? s->rt_sem[which][ind] : NULL); }

/* ======================================================================== formats ======== */

/* Go fmt %v for float64 = strconv.FormatFloat(v, 'g', -1, 64): the shortest round-trip digits,
 * exponent form when exp < -4 || exp >= 6 (strconv/ftoa.go: "if precision was the shortest
 * possible, use precision 6 for this decision"), at least two exponent digits.  Go stdlib, not in
 * /root/reference; note SURVEY.md Appendix A.15 quotes 21, which is encoding/json's threshold. */
int orc_format_float(double v, char *buf, size_t cap)
{
    if (isnan(v)) return snprintf(buf, cap, "NaN");
    if (isinf(v)) return snprintf(buf, cap, v > 0 ? "+Inf" : "-Inf");
    if (v == 0) return snprintf(buf, cap, signbit(v) ? "-0" : "0");
    char e[40];
    int prec;
    for (prec = 1; prec <= 17; ++prec) {
        snprintf(e, sizeof e, "%.*e", prec - 1, v);
        if (strtod(e, NULL) == v) break;
    }
    /* e = [-]d[.ddd]e[+-]XX */
    char digits[24]; int nd = 0; int neg = 0; const char *p = e;
    if (*p == '-') { neg = 1; ++p; }
    for (; *p && *p != 'e'; ++p) if (*p != '.') digits[nd++] = *p;
    int x = atoi(p + 1);
    while (nd > 1 && digits[nd - 1] == '0') --nd;             /* shortest: strip trailing zeros */
    char out[64]; int o = 0;
    if (neg) out[o++] = '-';
    if (x < -4 || x >= 6) {
        out[o++] = digits[0];
        if (nd > 1) { out[o++] = '.'; for (int i = 1; i < nd; ++i) out[o++] = digits[i]; }
        o += snprintf(out + o, sizeof out - o, "e%c%02d", x < 0 ? '-' : '+', x < 0 ? -x : x);
    } else if (x < 0) {
        out[o++] = '0'; out[o++] = '.';
        for (int i = 0; i < -x - 1; ++i) out[o++] = '0';
        for (int i = 0; i < nd; ++i) out[o++] = digits[i];
    } else {
        for (int i = 0; i <= x; ++i) out[o++] = i < nd ? digits[i] : '0';
        if (nd > x + 1) { out[o++] = '.'; for (int i = x + 1; i < nd; ++i) out[o++] = digits[i]; }
    }
    out[o] = 0;
    return snprintf(buf, cap, "%s", out);
}

/* read_input_data :381-421: whitespace tokens; nvar, nrow, then nrow x (nvar+1) floats */
int orc_read_dataset(const char *train, const char *test, int32_t *nvar, int32_t *nrow,
                     int32_t *nrow_test, double **rowmajor)
{
    FILE *a = fopen(train, "r"), *b = fopen(test, "r");
    if (!a || !b) { if (a) fclose(a); if (b) fclose(b); return -1; }
    int va, vb, ra, rb, rc = 0;
    if (fscanf(a, "%d %d", &va, &ra) != 2 || fscanf(b, "%d %d", &vb, &rb) != 2 || va != vb) rc = -2;
    double *m = NULL;
    if (!rc) {
        size_t stride = (size_t)va + 1, n = (size_t)(ra + rb) * stride;
        m = (double *)xcalloc(n, 8);
        for (size_t i = 0; i < (size_t)ra * stride && !rc; ++i) if (fscanf(a, "%lf", &m[i]) != 1) rc = -3;
        for (size_t i = (size_t)ra * stride; i < n && !rc; ++i) if (fscanf(b, "%lf", &m[i]) != 1) rc = -3;
    }
    fclose(a); fclose(b);
    if (rc) { free(m); return rc; }
    *nvar = va; *nrow = ra; *nrow_test = rb; *rowmajor = m;
    return 0;
}

/* read_sem :710-733: one value per line, nrow+nrow_test lines */
int orc_read_sem(const char *path, int32_t n, double *out)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[256]; int i = 0;
    while (i < n && fgets(line, sizeof line, f)) {
        char *end; double v = strtod(line, &end);
        if (end == line) { fclose(f); return -2; }
        out[i++] = v;
    }
    fclose(f);
    return i == n ? 0 : -3;
}
