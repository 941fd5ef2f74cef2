/*
 * gsgp_oracle.h -- CPU ORACLE for the go-gsgp generation-loop hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It is a plain-C restatement of the
 * reference's algorithm (akiross/go-gsgp, main_cpu.go).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it, and only as the checker
 * (or as the timed CPU baseline) -- never as a fallback for the CUDA path.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, fixtures or runnable tests for this
 * path (main_test.go does not compile at the reference snapshot) and there is no Go toolchain in
 * this image, so the reference itself cannot be run.  The oracle is pinned by hand-derived
 * known-answer tests (tests/test_oracle_kat.py), by the 4-row toy
 * dataset of main_test.go:107-112 and -- for the random stream, i.e. every draw the reference makes
 * after rand.Seed (main_cpu.go:1367) -- by known outputs of real Go programs (below).
 *
 * Every function cites the reference file:line it restates.
 */
#ifndef GSGP_ORACLE_H
#define GSGP_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ RNG ------------------ */
/* The reference draws everything from Go's global math/rand stream (main_cpu.go:1367).  The uint64
 * SOURCE is injectable; the derived draws (Int63, Int31, Float64, Intn) restate Go's math/rand v1
 * derivations on top of whichever source is plugged in.
 *   ORC_RNG_ALFG    Go's rngSource itself: additive lagged Fibonacci (607,273) seeded as rand.Seed
 *                   does.  Its 607-word seed table (rngCooked, Go stdlib -- not in /root/reference,
 *                   Go not installed) is recomputed from its definition (the generator's state 7.8e12
 *                   steps after srand(1)) by tools/gen_go_rng_cooked.py; the stream is pinned by known
 *                   outputs of Go programs (rand.Seed(1): Int63 5577006791947779410, ..., Intn(100)
 *                   81 87 47 59 81 18 25 40 56 0, Float64 0.6046602879796196, ...; seed 42: 5 87),
 *                   tests/test_oracle_kat.py::test_go_math_rand_known_answers.
 *   ORC_RNG_PHILOX  counter-based Philox4x32-10 keyed by seed, one stream per
 *                   (generation, individual): the stream the CUDA selection kernel uses. */
enum { ORC_RNG_ALFG = 0, ORC_RNG_PHILOX = 1 };

typedef struct {
    int      kind;
    /* ALFG state */
    uint64_t vec[607];
    int      tap, feed;
    /* Philox state */
    uint32_t key[2];
    uint32_t stream[3];   /* counter words 1..3 */
    uint32_t block;       /* counter word 0     */
    uint32_t out[4];
    int      have;        /* unread 64-bit halves left in out[] (0..2) */
    uint64_t n_draws;     /* number of uint64 values consumed (diagnostics) */
} orc_rng;

void     orc_rng_seed(orc_rng *r, int kind, int64_t seed);
void     orc_rng_seek(orc_rng *r, uint32_t s0, uint32_t s1, uint32_t s2); /* Philox: select stream */
uint64_t orc_rng_uint64(orc_rng *r);
int64_t  orc_rng_int63(orc_rng *r);
int32_t  orc_rng_int31(orc_rng *r);
double   orc_rng_float64(orc_rng *r);
int32_t  orc_rng_intn(orc_rng *r, int32_t n);
void     orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* ------------------------------------------------------------------ config --------------- */
enum { ORC_MSE = 0, ORC_RMSE = 1, ORC_MAE = 2, ORC_MRE = 3 };
enum { ORC_OP_INIT = 0, ORC_OP_XO = 1, ORC_OP_MUT = 2, ORC_OP_REPR = 3 }; /* evolution.proto:30-35 */

typedef struct {                       /* main_cpu.go:70-94, defaults :156-183 */
    int32_t population_size;
    int32_t init_type;                 /* 0 grow, 1 full, 2 ramped */
    double  p_crossover, p_mutation;
    int32_t max_depth_creation;
    int32_t tournament_size;
    int32_t zero_depth;
    int32_t num_random_constants;
    double  min_random_constant, max_random_constant;
    int32_t minimization_problem;
    int32_t error_measure;             /* ORC_MSE.. */
    int32_t n_workers;                 /* -workers: summation order + fan-out (main_cpu.go:180) */
    int32_t rng_kind;                  /* ORC_RNG_ALFG / ORC_RNG_PHILOX */
    int64_t rng_seed;
    int32_t use_linear_scaling;        /* -linsc (main_cpu.go:181,1311-1317) */
    int32_t reserved;
} orc_config;

void orc_config_defaults(orc_config *c);

/* One record per individual per generation: what the sequential driver decided
 * (main_cpu.go:1447-1470) -- this is the "generation plan" the C-ABI consumes in replay mode. */
typedef struct {
    int32_t op;        /* operator drawn: ORC_OP_XO / MUT / REPR                                  */
    int32_t elite;     /* 1 if k == index_best (operator degenerates to a copy of itself)          */
    int32_t p1, p2;    /* XO: parents.  MUT/REPR: p1 = reproduced source, p2 = -1                  */
    int32_t tree0_off, tree0_len;  /* into the generation's tree pool, -1/0 if unused              */
    int32_t tree1_off, tree1_len;
    double  ms;        /* mutation step drawn at main_cpu.go:919                                   */
} orc_plan_entry;

typedef struct orc_state orc_state;

/* ------------------------------------------------------------------ lifecycle ------------ */
orc_state *orc_new(const orc_config *cfg, int32_t nvar, int32_t nrow, int32_t nrow_test,
                   const double *rowmajor /* (nrow+nrow_test) x (nvar+1), target last, train first */);
void       orc_free(orc_state *s);
orc_rng   *orc_get_rng(orc_state *s);

/* main_cpu.go:426-455 (draws the constants), :1378-1388 (seeding), :1400-1403 (init + evaluate) */
void orc_create_T_F(orc_state *s);
int  orc_seed_semantic(orc_state *s, int32_t ind, const double *sem /* nrow+nrow_test */);
void orc_initialize_population(orc_state *s, int32_t n_seeds);
void orc_evaluate(orc_state *s);
/* Same as orc_evaluate but with caller-supplied trees for individuals [first, first+n): used by
 * parity tests so both sides evaluate the very same initial trees. */
void orc_set_initial_tree(orc_state *s, int32_t ind, const int32_t *tree, int32_t len);

/* main_cpu.go:1440-1492: one generation, sequential driver, draws from the state's RNG.
 * In ORC_RNG_PHILOX mode the stream is re-seeked to (generation, k) before each individual. */
void orc_generation(orc_state *s);
/* Apply an externally supplied plan (no RNG draws): the replay path. */
/* operators_effects.go:783-806,836-842,1246-1256: the one-step experiment (uniform parents); records plan + tree pool */
void orc_effects_step(orc_state *s, int test_crossover);
void orc_generation_replay(orc_state *s, const orc_plan_entry *plan, const int32_t *tree_pool);

/* ------------------------------------------------------------------ primitives ----------- */
/* which exp sits under exp64 (process-wide; tests only): glibc's exp + Go's |x| < 2^-28 rule (default), or Go's portable
 * math.Exp restated.  Go's own exp is not glibc's; both are accurate to under an ulp (see gsgp_oracle.c). */
enum { ORC_EXP_LIBM = 0, ORC_EXP_GO_PORTABLE = 1 };
void    orc_set_exp_kind(int kind);
double  orc_exp64(double v);                                              /* main_cpu.go:50-61   */
double  orc_sigmoid(double r);                                            /* :894                */
int32_t orc_create_grow_tree_arrays(orc_state *s, int32_t depth, int32_t max_depth,
                                    int32_t base_index, int32_t *out, int32_t cap); /* :609-639  */
double  orc_eval_arrays(const orc_state *s, const int32_t *tree, int32_t start, int32_t i); /* :755-775 */
void    orc_semantic_evaluate_array(const orc_state *s, const int32_t *tree,
                                    int32_t sem_size, int32_t sem_offs, double *val); /* :797-827 */
double  orc_fitness_train(const orc_state *s, const double *sem);          /* :950-985            */
double  orc_fitness_test(const orc_state *s, const double *sem);           /* :1106-1141          */
/* linear scaling (-linsc): train returns the fitness and the scaling a, b; test applies the train a, b */
double  orc_fitness_train_ls(const orc_state *s, const double *sem, double *a, double *b);   /* :991-1104  */
double  orc_fitness_test_ls(const orc_state *s, const double *sem, double a, double b);      /* :1142-1177 */
int32_t orc_tournament_selection(orc_state *s);                            /* :830-840            */
int32_t orc_best_individual(const orc_state *s);                           /* :1180-1188          */
int     orc_better(const orc_state *s, double f1, double f2);              /* :1206-1212          */
/* stand-alone GS operators on caller vectors (n elements), used by primitive parity tests */
void    orc_gs_crossover_vec(const double *p1, const double *p2, const double *r, double *out, int64_t n); /* :893-896 */
void    orc_gs_mutation_vec(const double *t, const double *r1, const double *r2, double ms, double *out, int64_t n); /* :932-936 */

/* ------------------------------------------------------------------ accessors ------------ */
int32_t        orc_generation_index(const orc_state *s);
int32_t        orc_index_best(const orc_state *s);
const double  *orc_fit(const orc_state *s);
const double  *orc_fit_test(const orc_state *s);
const double  *orc_sem_train(const orc_state *s, int32_t ind);
const double  *orc_sem_test(const orc_state *s, int32_t ind);
const orc_plan_entry *orc_last_plan(const orc_state *s);
const int32_t *orc_last_tree_pool(const orc_state *s, int32_t *len);
const int32_t *orc_initial_tree(const orc_state *s, int32_t ind, int32_t *len);
const double  *orc_symbol_values(const orc_state *s, int32_t *n_symbols);
/* semantics of the random trees of individual `ind` in the last generation (proto dump fields) */
const double  *orc_last_tree_semantic(const orc_state *s, int32_t ind, int32_t which);
void           orc_keep_tree_semantics(orc_state *s, int flag);   /* off by default (O(P*N) memory) */
void           orc_test_set_index_best(orc_state *s, int32_t b);  /* test hooks */
void           orc_test_set_generation(orc_state *s, int32_t g);

/* ------------------------------------------------------------------ formats -------------- */
/* Go's fmt "%v" for float64 (shortest round-trip %g: exponent form when exp < -4 || exp >= 6). */
int  orc_format_float(double v, char *buf, size_t cap);                    /* main_cpu.go:123-126,1405 */
/* dataset reader main_cpu.go:381-421; returns malloc'd row-major matrix */
int  orc_read_dataset(const char *train, const char *test, int32_t *nvar, int32_t *nrow,
                      int32_t *nrow_test, double **rowmajor);
int  orc_read_sem(const char *path, int32_t n, double *out);                /* :710-733           */

#ifdef __cplusplus
}
#endif
#endif
