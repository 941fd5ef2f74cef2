/*
 * gsgp_host.h -- host-side half of the hot path: the parts of main_cpu.go's driver that stay on the
 * CPU (they define the RNG draw order) restated in C++ for hosts that cannot link Go: the seeded
 * RNG, create_T_F's constants, the initial-population tree builders and the per-generation
 * "plan" = the decisions of the sequential loop at main_cpu.go:1447-1470 without its arithmetic.
 * A Go host (main_b200.go) keeps its own main_cpu.go code for all of this and only
 * needs gsgp_b200.h; the C++ CLI driver and the Python harness use these entry points instead.
 * Exported from the same libgsgp_b200.so.  No CUDA involved.
 */
#ifndef GSGP_HOST_H
#define GSGP_HOST_H

#include <stdint.h>
#include "gsgp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsgph_rng gsgph_rng;

/* math/rand (v1) restated: Go's additive lagged Fibonacci source (607,273), seeded as rand.Seed(seed) does, and Go's
 * derivations of Float64/Intn.  The stream is Go's own, bit for bit: the 607-word seed table of
 * the Go stdlib is recomputed from its definition by tools/gen_go_rng_cooked.py (csrc/go_rng_cooked.inc.h) and pinned
 * by known outputs of Go programs (rand.Seed(1): Intn(100) = 81 87 47 59 81 ..., Float64 = 0.6046602879796196 ...;
 * tests/test_host_driver.py::test_host_rng_is_gos_seeded_stream).  A C++ host over this library therefore makes the
 * draws main_cpu.go makes for the same -seed. */
gsgph_rng *gsgph_rng_new(int64_t seed);                    /* rand.Seed   main_cpu.go:1367 */
void       gsgph_rng_free(gsgph_rng *r);
double     gsgph_rng_float64(gsgph_rng *r);
int32_t    gsgph_rng_intn(gsgph_rng *r, int32_t n);

typedef struct {
    int32_t population_size, nvar, num_constants;
    int32_t max_depth_creation, tournament_size;
    int32_t zero_depth, minimization_problem, init_type;   /* init_type: 0 grow, 1 full, 2 ramped */
    double  p_crossover, p_mutation;
    double  min_random_constant, max_random_constant;
} gsgph_params;

/* create_T_F (main_cpu.go:426-455): Symbol.value per id; draws one Float64 per constant */
int gsgph_create_T_F(gsgph_rng *r, const gsgph_params *p, double *sym_val /* 4+nvar+num_constants */);

/* initialize_population (:556-565) + tree_to_array (:675-693) for individuals n_seeds..P-1:
 * tree t of individual n_seeds+t is pool[off[t] .. off[t]+len[t]).  Returns non-zero if cap is too small. */
int gsgph_initialize_population(gsgph_rng *r, const gsgph_params *p, int32_t n_seeds, int32_t *pool,
                                int32_t cap, int32_t *off, int32_t *len, int32_t *pool_len);

/* one generation's decisions (:1447-1470): operator draw, tournament_selection (:830-840) on the OLD
 * fit[], create_grow_tree_arrays (:609-639), mutation step (:919); elite slots only draw the operator */
int gsgph_build_plan(gsgph_rng *r, const gsgph_params *p, const double *fit, int32_t index_best,
                     gsgp_plan_entry *plan, int32_t *pool, int32_t cap, int32_t *pool_len);

int32_t gsgph_best_individual(const gsgph_params *p, const double *fit);     /* :1180-1188 */

/* ---- pipelined planner: gsgph_build_plan's result, with the fitness-independent draws made while the device works
 * Generation g+1's plan needs generation g's fitness only for the tournament WINNERS (:835) and for the elite index
 * (the elite slot skips its draws, :847,:877,:918).  gsgph_planner_prefetch starts a worker thread drawing the next
 * generation -- operators, random trees, mutation steps, tournament CANDIDATES -- under a guess of its elite index
 * (the index just used: the elite keeps its slot unless a strictly better child appears); gsgph_planner_plan, called
 * when the fitness is there, picks the winners and, if the guess was wrong, redraws the individuals from the first
 * affected slot on from the recorded raw RNG values.  Plans, tree pools and the RNG stream are exactly those of
 * gsgph_build_plan called once per generation (tests/test_host_driver.py).  While a planner exists the rng belongs to
 * it between prefetch and plan.  Typical loop:
 *     gsgph_planner_plan(pl, fit, best, &plan, &pool, &len);      // generation g+1 (first call: drawn on the spot)
 *     gsgph_planner_prefetch(pl, best);                           // g+2's draws start now ...
 *     gsgp_generation(ctx, plan, pool, len, fit, fit_test, &best);// ... and run under this call
 * A Go host gets the same effect with a goroutine and a rand.Source wrapper that records Int63. */
typedef struct gsgph_planner gsgph_planner;
gsgph_planner *gsgph_planner_new(gsgph_rng *r, const gsgph_params *p);        /* NULL on bad parameters */
void gsgph_planner_free(gsgph_planner *pl);                 /* draws of an unused prefetch go back to the stream */
int gsgph_planner_prefetch(gsgph_planner *pl, int32_t index_best_guess);      /* 2: a prefetch is already pending */
/* plan/pool stay valid until the next gsgph_planner_plan call returns; 3: tree pool exhausted */
int gsgph_planner_plan(gsgph_planner *pl, const double *fit, int32_t index_best, const gsgp_plan_entry **plan,
                       const int32_t **pool, int32_t *pool_len);
void gsgph_planner_stats(const gsgph_planner *pl, int64_t *hits, int64_t *misses, int64_t *redrawn_individuals);
/* Also compile each plan's trees into the interpreter's programs, ahead of time like the draws (worker + n_threads - 1
 * helper threads; <= 0: default), for gsgp_generation_compiled: the call then starts the kernels at once.  Call before
 * the first prefetch.  gsgph_planner_programs returns the programs of the plan gsgph_planner_plan returned last
 * (same lifetime). */
int gsgph_planner_enable_programs(gsgph_planner *pl, int32_t n_threads);
int gsgph_planner_programs(const gsgph_planner *pl, const uint32_t **programs, int32_t *n_instructions,
                           const int32_t **prog_off, const int32_t **prog_desc);
/* The generation loop of main_cpu.go:1440-1506 in host-replay mode, n times: plan (pipelined) -> gsgp_generation[_compiled]
 * -> fitness back.  fit / fit_test / index_best: the current generation's on entry (gsgp_fitness_all or the previous
 * call), the last generation's on return; *_history (n values each, may be NULL) = fit[best], fit_test[best] per
 * generation, the fitnesstrain / fitnesstest lines (:1495-1496).  4: a gsgp_* call failed (gsgp_last_error).
 * On return the following generation's draws are already under way (used by the next call or by gsgph_planner_plan;
 * gsgph_planner_free returns them to the stream). */
int gsgph_replay_generations(gsgph_planner *pl, gsgp_ctx *ctx, int32_t n_generations, double *fit, double *fit_test,
                             int32_t *index_best, double *best_fit_history, double *best_fit_test_history);
/* ---- one plan for every rank of a node (multi-GPU host replay) ------------------------------------------------------
 * A generation's plan is identical on every rank (same RNG stream, same fitness), so one process -- the writer -- runs
 * the planner and the others take plan + tree pool + programs from a POSIX shared-memory segment instead of each running
 * a planner and its compile threads.  Sequence numbers start at 1 and grow by one per generation; sequence s lives in
 * slot s & 1, which the writer reuses once every reader has released s - 2.  Pointers returned by _acquire point into
 * the segment and stay valid until the reader releases that sequence.  timeout_ms < 0 waits for ever.
 * Returns: 1 bad arguments, 3 plan larger than the segment's capacities, 5 timeout, 6 the slot was overwritten. */
typedef struct gsgph_channel gsgph_channel;
gsgph_channel *gsgph_channel_create(const char *name /* "/..." as for shm_open */, int32_t population_size,
                                    int32_t cap_pool_ints, int32_t cap_instructions, int32_t n_readers);   /* writer */
gsgph_channel *gsgph_channel_attach(const char *name, int32_t reader_index, int32_t timeout_ms);           /* reader */
void gsgph_channel_close(gsgph_channel *c);                                    /* the writer also unlinks the segment */
int gsgph_channel_publish(gsgph_channel *c, uint64_t seq, const gsgp_plan_entry *plan, const int32_t *pool, int32_t pool_len,
                          const uint32_t *programs, int32_t n_instructions, const int32_t *prog_off, const int32_t *prog_desc,
                          int32_t timeout_ms);
int gsgph_channel_acquire(gsgph_channel *c, uint64_t seq, const gsgp_plan_entry **plan, const int32_t **pool, int32_t *pool_len,
                          const uint32_t **programs, int32_t *n_instructions, const int32_t **prog_off, const int32_t **prog_desc,
                          int32_t timeout_ms);
int gsgph_channel_release(gsgph_channel *c, uint64_t seq);
/* gsgph_replay_generations across the ranks of a node: the writer (pl = its planner, programs enabled) plans, publishes
 * and runs its own generation; a reader (pl = NULL) acquires, runs gsgp_generation_compiled and releases.  first_seq:
 * 1 for the first call, then first_seq + n_generations of the previous call -- the same on every rank. */
int gsgph_replay_generations_shared(gsgph_planner *pl, gsgph_channel *ch, gsgp_ctx *ctx, uint64_t first_seq, int32_t n_generations,
                                    double *fit, double *fit_test, int32_t *index_best, double *best_fit_history,
                                    double *best_fit_test_history, int32_t timeout_ms);
/* the same staging for hosts that build their plans themselves (a Go host's goroutine): every tree of `plan` through
 * gsgph_compile_tree's compiler, programs concatenated in slot order (slot = 2*k + which).  programs: 2 uint32 per
 * instruction, cap instructions (pool_len / 4 + 2 * population_size always suffices); 2: malformed tree, 3: cap */
int gsgph_compile_plan(const gsgp_plan_entry *plan, int32_t population_size, const int32_t *tree_pool, int32_t pool_len,
                       int32_t n_symbols, uint32_t *programs, int32_t cap, int32_t *n_instructions, int32_t *prog_off,
                       int32_t *prog_desc);

/* the one-step experiment of operators_effects.go:1246-1256 as a plan for gsgp_generation: test_crossover != 0 ->
 * every individual is a crossover of two UNIFORMLY drawn parents (random_selection :783-786, :800-806); else every
 * individual is copied onto itself and mutated (:790-797, :836-842).  index_best is never assigned in that program,
 * so individual 0 plays the elite (:197,:801,:837). */
int gsgph_build_effects_plan(gsgph_rng *r, const gsgph_params *p, int32_t test_crossover,
                             gsgp_plan_entry *plan, int32_t *pool, int32_t cap, int32_t *pool_len);

/* The interpreter's program for one tree (csrc/program.hpp), as the shim builds it before upload:
 * pointer-prefix array (main_cpu.go:609-639 format) -> validated postfix -> accumulator program.
 * code_ab receives 2 uint32 per instruction {level<<8 | flags | aop, a | b<<16}; returns non-zero on a
 * malformed tree or when cap (instructions) is too small.  No CUDA involved: lets hosts and CPU tests inspect
 * exactly what the device will run. */
int gsgph_compile_tree(const int32_t *tree, int32_t len, int32_t n_symbols, uint32_t *code_ab, int32_t cap,
                       int32_t *n_ins, int32_t *spill_levels);

/* ---- ingest: the reference's text formats, parsed fast ---------------------------------------------
 * Whole-file read, tokens split over worker threads, std::from_chars (correctly rounded, like strconv.ParseFloat).
 * Same files as read_input_data (main_cpu.go:381-421: `nvar`, `nrow`, then nrow x (nvar+1) whitespace-separated
 * floats, train file then test file) and read_sem (:710-733: one value per line, nrow+nrow_test of them). */
int gsgph_read_dataset(const char *train_path, const char *test_path, int32_t *nvar, int32_t *nrow, int32_t *nrow_test,
                       double **rowmajor /* (nrow+nrow_test) x (nvar+1), malloc'd: release with gsgph_free */);
int gsgph_read_semantic(const char *path, int32_t n, double *out);   /* non-zero: cannot open / bad value / too few values */
void gsgph_free(void *p);
/* the tokenizer/parser behind both: n_threads <= 0 picks a default */
int gsgph_parse_floats(const char *text, int64_t size, double *out, int64_t cap, int64_t *n_out, int32_t n_threads);

/* Go fmt %v of a float64 (fitness / semantic output lines, main_cpu.go:123-126,1405-1409) */
int gsgph_format_float(double v, char *buf, int32_t cap);
/* time.Duration.String() of a nanosecond count: the executiontime lines (main_cpu.go:1413,1505, parsed by the harness's
 * reporter.py:87-121), e.g. "1.234567ms", "2m3.5s", "1h0m0s".  Returns the length, -1 if cap (40 always suffices) is short. */
int gsgph_format_duration(int64_t ns, char *out, int32_t cap);
/* Semantic.String() (main_cpu.go:123-126): %v values joined by commas -- the best-of-generation semantic line
 * (:1499-1500).  Multi-threaded (n_threads <= 0: default); 25 bytes per value always suffice; non-zero if cap is short. */
int gsgph_format_semantic(const double *v, int64_t n, char *out, int64_t cap, int64_t *len_out, int32_t n_threads);

/* ---- asynchronous line writers for the per-generation outputs (main_cpu.go:1227-1241,1494-1500) ----------------
 * A path ending in ".gz" is written as gzip (the reference's default for the semantic files, :175-176), anything
 * else as plain text.  The calls copy their input and return; a background thread formats and compresses (parallel
 * gzip members), so the generation loop never waits for text.  gsgph_writer_close drains and closes; 0 = all written. */
typedef struct gsgph_writer gsgph_writer;
gsgph_writer *gsgph_writer_open(const char *path, int32_t n_threads);          /* NULL if the file cannot be created */
int gsgph_writer_semantic(gsgph_writer *w, const double *v, int64_t n);        /* fmt.Fprintln(file, Semantic) :1499-1500 */
int gsgph_writer_float(gsgph_writer *w, double v);                            /* fmt.Fprintln(file, fit[best]) :1495-1496 */
int gsgph_writer_close(gsgph_writer *w);
/* the writers' compressor on its own: src[0, n) as one gzip member (what a ".gz" writer emits per 4 MB of text).  The
 * lines are decimal digits -- nothing for LZ77 to match -- so the member is one dynamic-Huffman block of literals built
 * directly (histogram, code lengths, table-driven bit packing), with zlib's Z_HUFFMAN_ONLY as the fallback for input
 * whose code would exceed 15 bits.  cap >= n + n/8 + 512 always suffices; 3: cap too small (len_out = needed). */
int gsgph_gzip_member(const char *src, int64_t n, char *out, int64_t cap, int64_t *len_out);

/* ---- -proto_dump (main_cpu.go:1416-1436,1441-1488,1510-1517; messages of evolution.proto) ------------------------
 * pb.Evolution built population by population and written in the protobuf wire format proto.Marshal produces (proto3:
 * packed repeated scalars, zero scalars / empty fields left out), so the reference's analysis tooling reads the file
 * unchanged.  One population is open at a time: begin, add every individual in index order, end.
 *   op                 pb.Individual_Operator: 0 INIT, 1 XO, 2 MUT, 3 REPR (:1425,1454,1460,1468)
 *   contrib            contrib[k].AsBytes() of the host's bookkeeping (may be NULL / 0)
 *   rt_*               the operator's random trees {rt, sem_rt_train, sem_rt_test} (:1456-1466): n_rt = 0, 1 or 2;
 *                      rt_semantic_train[t] has nrow values, rt_semantic_test[t] nrow_test (either array may be NULL)
 * All pointers are read during the call only.  Non-zero: bad arguments / wrong call order; gsgph_proto_add_individual 4:
 * the dump would pass protobuf's 2 GB message limit (where the reference's proto.Marshal fails); gsgph_proto_write 2: I/O. */
typedef struct gsgph_proto gsgph_proto;
gsgph_proto *gsgph_proto_new(void);                                            /* new(pb.Evolution) :1419 */
void gsgph_proto_free(gsgph_proto *d);
int gsgph_proto_begin_population(gsgph_proto *d, int32_t generation);          /* new(pb.Population) :1420-1421,1443-1444 */
int gsgph_proto_add_individual(gsgph_proto *d, int32_t op, double fitness_train, double fitness_test,
                               const double *semantic_train, int64_t nrow, const double *semantic_test, int64_t nrow_test,
                               const uint8_t *contrib, int64_t contrib_len,
                               int32_t n_rt, const int32_t *const *rt_data, const int32_t *rt_len,
                               const double *const *rt_semantic_train, const double *const *rt_semantic_test);
int gsgph_proto_end_population(gsgph_proto *d);                                /* append(pb_evo.Generations, pb_pop) :1436,1487 */
int64_t gsgph_proto_size(const gsgph_proto *d);                                /* bytes of the closed populations so far */
int gsgph_proto_bytes(const gsgph_proto *d, uint8_t *out, int64_t cap);        /* proto.Marshal(pb_evo) :1511 */
int gsgph_proto_write(const gsgph_proto *d, const char *path);                 /* ioutil.WriteFile :1515 */

/* ---- contributions (main_cpu.go:128-149: Contribution = []*big.Int, one slot per model) --------------------------
 * Which seeded model (slot i + 1 for seed file i, slot 0 for GP itself) an individual descends from, counted with
 * multiplicity: crossover adds the parents' vectors (merge_contribs :869-873,885), reproduction / mutation / the elite
 * copy one (copy_contrib :857,909).  The values double per generation at most; they are kept as fixed-width limb
 * arrays that widen on demand, so a generation of config 3 (pop 2,000 x 501 models) is one streaming pass.
 *   gsgph_contrib_new     init_tables + the initial slots (:1272-1283,1390-1397)
 *   gsgph_contrib_apply   one generation from its plan (op / elite / p1 / p2), then the swap of update_tables (:1196);
 *                         an elite entry keeps its own vector (its p1 / p2 are not read: gsgp_get_plan leaves them
 *                         undefined in device mode); 2: a parent index out of range
 *   gsgph_contrib_format  Contribution.String() of one individual (:141-146, the line written at :1411,1503 and the
 *                         bytes of pb.Individual.contrib :1430,1479); 3: cap too small (len_out = needed) */
typedef struct gsgph_contrib gsgph_contrib;
gsgph_contrib *gsgph_contrib_new(int32_t population_size, int32_t n_seeds);
void gsgph_contrib_free(gsgph_contrib *c);
int32_t gsgph_contrib_num_models(const gsgph_contrib *c);
int32_t gsgph_contrib_limbs(const gsgph_contrib *c);                           /* 64-bit limbs per value right now */
int gsgph_contrib_apply(gsgph_contrib *c, const gsgp_plan_entry *plan, int32_t n_threads);
int gsgph_contrib_format(const gsgph_contrib *c, int32_t ind, char *out, int64_t cap, int64_t *len_out);

#ifdef __cplusplus
}
#endif
#endif
