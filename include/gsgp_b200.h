/*
 * gsgp_b200.h -- C-ABI of the B200-native go-gsgp generation-loop hot path.
 *
 * The reference (akiross/go-gsgp) has no plugin/FFI interface; its only precedent for swapping the
 * hot path is main_gpu.go, which re-implements init_tables / evaluate / reproduction /
 * geometric_semantic_crossover / geometric_semantic_mutation / update_tables on device buffers
 * behind the same main() (main_gpu.go:688-853,897-914, cgo preamble :24-27).  This header is the
 * seam a third Go file (main_b200.go) binds through cgo: plain pointers and
 * sizes, caller-owned host buffers valid only for the duration of a call, device memory owned by
 * the opaque context.  Every entry point names the reference code it replaces (file:line into the
 * reference tree).
 *
 * Conventions
 *   - return 0 on success, non-zero on failure; gsgp_last_error() gives the message.  The reference
 *     panics on every failure (e.g. main_cpu.go:383-386,725,730); the Go wrapper panics on non-zero.
 *   - no CPU fallback: every call needs a CUDA device (sm_100a build).
 *   - one caller thread per context (main_gpu.go:929 pins the goroutine for the same reason).
 *   - symbol ids: 0..3 = + - * / (main_cpu.go:434-437), 4..4+nvar-1 variables, then constants.
 *   - trees: the reference's pointer-prefix int32 arrays (create_grow_tree_arrays,
 *     main_cpu.go:609-639; wire format RandomTree.data, evolution.proto:19).
 *   - semantic vectors: train cases first, then test cases (main_cpu.go:1386-1387).
 *   - multi-GPU: fitness cases are sharded; rank r of W owns train rows [r*nrow/W,(r+1)*nrow/W) and
 *     the same split of the test rows (gsgp_shard_range).  Counts in gsgp_config are GLOBAL.
 */
#ifndef GSGP_B200_H
#define GSGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSGP_ABI_VERSION 1

enum { GSGP_MSE = 0, GSGP_RMSE = 1, GSGP_MAE = 2, GSGP_MRE = 3 };          /* main_cpu.go:1319-1331 */
enum { GSGP_OP_INIT = 0, GSGP_OP_XO = 1, GSGP_OP_MUT = 2, GSGP_OP_REPR = 3 }; /* evolution.proto:30-35 */
enum { GSGP_RNG_HOST_REPLAY = 0,   /* host (Go math/rand) makes every draw and passes a plan        */
       GSGP_RNG_DEVICE_PHILOX = 1  /* selection + random trees drawn on the device (Philox4x32-10)  */ };

typedef struct gsgp_ctx gsgp_ctx;

typedef struct {
    int32_t population_size;        /* config.population_size          main_cpu.go:157 */
    int32_t nvar;                   /* nvar                            :399            */
    int32_t nrow, nrow_test;        /* GLOBAL train / test case counts :404-405        */
    int32_t max_depth_creation;     /* :162 */
    int32_t tournament_size;        /* :163 */
    int32_t zero_depth;             /* :164 */
    int32_t minimization_problem;   /* :169 */
    int32_t error_measure;          /* GSGP_MSE..                      :179,:1319-1331 */
    int32_t num_constants;          /* NUM_CONSTANT_SYMBOLS            :1299           */
    double  p_crossover, p_mutation;/* :160-161 (device-RNG mode only)                 */
    int64_t rng_seed;               /* :172     (device-RNG mode only)                 */
    int32_t rng_mode;               /* GSGP_RNG_*                                       */
    int32_t device;                 /* CUDA device ordinal                              */
    int32_t rank, world_size;       /* case sharding; world_size 1 = single GPU         */
    uint8_t nccl_id[128];           /* ncclUniqueId from gsgp_nccl_unique_id (world_size > 1) */
    int64_t workspace_bytes;        /* budget for the random-tree semantic workspace; 0 = auto */
    int32_t flags;                  /* GSGP_FLAG_* */
    int32_t reserved;
} gsgp_config;

enum { GSGP_FLAG_KEEP_TREE_SEMANTICS = 1, /* keep R of the whole last generation (proto dump)                  */
       GSGP_FLAG_LINEAR_SCALING = 2,      /* -linsc: fitness on a + b*semantic (main_cpu.go:181,991-1104,1142-1177) */
       GSGP_FLAG_KEEP_RUN_HISTORY = 4,    /* keep, per generation, the decisions and the best individual's semantic on the
                                             device, so that gsgp_run_generations(n) loses none of the reference's
                                             per-generation outputs (semantic lines :1499-1500, contributions :885,:1503) */
       GSGP_FLAG_SINGLE_BUFFER_STORE = 8  /* one semantic store instead of main_cpu.go's cur/new pair (:205-228,1191-1197):
                                             rows are [cases | one free block of cases]; a generation is computed block by
                                             block, each child block written where the previous block has been consumed,
                                             so the rows shift by one block per generation (alternately right and left).
                                             Chosen automatically when two buffers do not fit the device (config 5: 10M
                                             cases x pop 10,000 on 8 GPUs = 100 GB of semantics per GPU) */ };

/* One entry per individual per generation: the decisions of the sequential driver loop
 * (main_cpu.go:1447-1470).  In host-replay mode the host fills it; in device mode the selection
 * kernel does and gsgp_get_plan returns it (the host needs p1/p2 for merge_contribs, :885). */
typedef struct {
    int32_t op;        /* operator drawn: GSGP_OP_XO / MUT / REPR                              */
    int32_t elite;     /* 1 if k == index_best: the operator degenerates to a self copy (:877,:847,:918) */
    int32_t p1, p2;    /* XO: tournament winners (:881-882).  MUT/REPR: p1 = reproduced source (:849) */
    int32_t tree0_off, tree0_len;  /* XO rt1 / MUT rt1: offset,len in tree_pool; -1,0 if unused */
    int32_t tree1_off, tree1_len;  /* MUT rt2                                                   */
    double  ms;        /* mutation step = rand.Float64() (:919)                                 */
} gsgp_plan_entry;

/* ---- library ------------------------------------------------------------------------------ */
int32_t     gsgp_abi_version(void);
void        gsgp_config_defaults(gsgp_config *cfg);            /* main_cpu.go:156-183 defaults   */
const char *gsgp_last_error(const gsgp_ctx *ctx);              /* ctx may be NULL (create errors) */
int gsgp_nccl_unique_id(uint8_t out[128]);                     /* rank 0 calls, host broadcasts   */
void gsgp_shard_range(int32_t n, int32_t rank, int32_t world, int32_t *lo, int32_t *hi);

/* ---- lifecycle: init_tables (main_cpu.go:1257-1284) ---------------------------------------- */
int  gsgp_create(const gsgp_config *cfg, gsgp_ctx **out);
void gsgp_destroy(gsgp_ctx *ctx);
int  gsgp_local_counts(const gsgp_ctx *ctx, int32_t *nrow_local, int32_t *nrow_test_local);
/* how the semantic store (init_tables :1257-1284) was laid out: buffers = 2 (cur/new) or 1 (single-buffer store),
 * bytes per buffer, bytes of the free case block the single buffer carries (0 with two buffers) */
int  gsgp_store_info(const gsgp_ctx *ctx, int32_t *buffers, int64_t *buffer_bytes, int64_t *scratch_bytes);
/* how the per-generation partial sums travel between ranks: 0 = single GPU, 1 = ncclAllGather, 2 = fused into the
 * reduction kernel as direct stores into the peers' buffers (CUDA IPC over NVLink; GSGP_EXCHANGE=nccl forces 1) */
int  gsgp_exchange_mode(const gsgp_ctx *ctx);

/* ---- inputs -------------------------------------------------------------------------------- */
/* the in-memory `set` built by read_input_data (main_cpu.go:381-421): row-major
 * (nrow+nrow_test) x (nvar+1), target last, train rows first, GLOBAL rows (every rank passes the
 * same matrix; each keeps its shard).  Same layout main_gpu.go:1036-1046 uploads. */
int gsgp_set_dataset(gsgp_ctx *ctx, const double *rowmajor);
/* shard-only variant: local train rows and local test rows, each row-major x (nvar+1) */
int gsgp_set_dataset_shard(gsgp_ctx *ctx, const double *train_rows, const double *test_rows);
/* Symbol.value per symbol id (create_T_F main_cpu.go:450-454; main_gpu.go:1049-1059) */
int gsgp_set_constants(gsgp_ctx *ctx, const double *sym_val, int32_t n_symbols);
/* semantic feeding (main_cpu.go:1383-1388): sem = nrow+nrow_test GLOBAL values, train first */
int gsgp_seed_semantic(gsgp_ctx *ctx, int32_t ind, const double *sem);
/* the same from the reference's semantic files (read_sem main_cpu.go:710-733, one value per line): file i seeds
 * individual first_ind + i.  Parsing (multi-threaded, include/gsgp_host.h) of file i+1 overlaps the upload of file i
 * through two pinned buffers.  Fails like the reference panics: unreadable file, bad value, too few values. */
int gsgp_seed_semantic_files(gsgp_ctx *ctx, int32_t first_ind, int32_t n_files, const char *const *paths);

/* ---- evaluate() (main_cpu.go:780-795) ------------------------------------------------------ */
/* tree -> semantics for individuals [first_ind, first_ind+n): the tree_to_array +
 * semantic_evaluate_array half of evaluate() (:784-786) */
int gsgp_eval_initial(gsgp_ctx *ctx, int32_t first_ind, int32_t n, const int32_t *tree_pool,
                      int32_t pool_len, const int32_t *tree_off, const int32_t *tree_len);
/* fitness of every individual + best_individual(): the tail of evaluate() (:788-790, :1403).
 * Any output pointer may be NULL. */
int gsgp_fitness_all(gsgp_ctx *ctx, double *fit, double *fit_test, int32_t *index_best);

/* ---- one generation (main_cpu.go:1447-1492) ------------------------------------------------- */
/* host-replay mode: apply `plan` (population_size entries) = reproduction (:844-861),
 * geometric_semantic_crossover (:876-914), geometric_semantic_mutation (:917-947) for every k,
 * then update_tables (:1191-1197) and best_individual (:1180-1188).
 * fit_out / fit_test_out (population_size each) and index_best_out may be NULL. */
int gsgp_generation(gsgp_ctx *ctx, const gsgp_plan_entry *plan, const int32_t *tree_pool,
                    int32_t pool_len, double *fit_out, double *fit_test_out, int32_t *index_best_out);
/* gsgp_generation with the staging done by the caller: the random trees of `plan` arrive already compiled into the
 * interpreter's programs (gsgph_compile_trees in gsgp_host.h, or the pipelined planner gsgph_planner_*, which compiles
 * while the device computes the previous generation), so the call starts the kernels at once instead of compiling
 * ~1.2 trees per individual first (0.11 ms of an 0.89 ms call at population 1000).
 *   programs      2 uint32 per instruction (gsgph_compile_tree's format), n_instructions of them
 *   prog_off      per tree slot 2*k + which: first instruction of the slot's program
 *   prog_desc     per tree slot: instruction count | spill levels << 24; slots without a tree are ignored
 * The programs are checked for what memory safety needs (ranges, operand ids, spill levels), not re-derived from the
 * trees: a program that does not belong to its tree yields that program's semantics.  tree_pool is still required
 * (gsgp_get_tree, proto dump). */
int gsgp_generation_compiled(gsgp_ctx *ctx, const gsgp_plan_entry *plan, const int32_t *tree_pool, int32_t pool_len,
                             const uint32_t *programs, int32_t n_instructions, const int32_t *prog_off,
                             const int32_t *prog_desc, double *fit_out, double *fit_test_out, int32_t *index_best_out);
/* device-RNG mode: run n whole generations on the device (selection, random trees, operators,
 * fitness, best) without host involvement; asynchronous, use gsgp_sync or any getter to wait. */
int gsgp_run_generations(gsgp_ctx *ctx, int32_t n);
int gsgp_sync(gsgp_ctx *ctx);

/* ---- outputs -------------------------------------------------------------------------------- */
int gsgp_get_fitness(gsgp_ctx *ctx, double *fit, double *fit_test, int32_t *index_best);
/* semantic of one individual, LOCAL shard: nrow_local train values then nrow_test_local test
 * values (best-of-generation writer main_cpu.go:1499-1500, proto dump :1477-1478) */
int gsgp_get_semantic(gsgp_ctx *ctx, int32_t ind, double *out);
/* semantic of random tree `which` (0/1) of individual ind in the last generation
 * (sem_rt1_* / sem_rt2_*, main_cpu.go:222-226,1456-1466); needs GSGP_FLAG_KEEP_TREE_SEMANTICS */
int gsgp_get_tree_semantic(gsgp_ctx *ctx, int32_t ind, int32_t which, double *out);
/* last generation's decisions (device mode: what the selection kernel drew) */
int gsgp_get_plan(gsgp_ctx *ctx, gsgp_plan_entry *plan_out);
/* random tree `which` of individual ind in the last generation, reference array format */
int gsgp_get_tree(gsgp_ctx *ctx, int32_t ind, int32_t which, int32_t *out, int32_t cap, int32_t *len);
/* per-generation best-of-generation record kept on the device (fitness files, main_cpu.go:1495-1496):
 * entries for generations [first, first+n) where 0 = initial population */
int gsgp_get_history(gsgp_ctx *ctx, int32_t first, int32_t n, double *best_fit, double *best_fit_test,
                     int32_t *best_index);
int32_t gsgp_generation_index(const gsgp_ctx *ctx);
/* with GSGP_FLAG_KEEP_RUN_HISTORY: the best individual's semantic (local shard, layout of gsgp_get_semantic) and the
 * decisions of any past generation (0 = initial population: semantic only) */
int gsgp_get_best_semantic_of(gsgp_ctx *ctx, int32_t generation, double *out);
int gsgp_get_plan_of(gsgp_ctx *ctx, int32_t generation, gsgp_plan_entry *plan_out);

/* ---- primitives (parity tests; also usable by operators_effects-style hosts) ------------------ */
/* semantic_evaluate_array (main_cpu.go:797-827) for n trees; out = n x (nrow_local+nrow_test_local) */
int gsgp_eval_trees(gsgp_ctx *ctx, int32_t n, const int32_t *tree_pool, int32_t pool_len,
                    const int32_t *tree_off, const int32_t *tree_len, double *out);
/* overwrite fit[] used by selection (tests) */
int gsgp_set_fitness(gsgp_ctx *ctx, const double *fit, const double *fit_test);
/* device-mode selection + tree generation only, for generation index g (no semantic update) */
int gsgp_draw_plan(gsgp_ctx *ctx, int32_t generation, gsgp_plan_entry *plan_out);

/* ---- measurement ----------------------------------------------------------------------------- */
void *gsgp_stream(gsgp_ctx *ctx);                        /* cudaStream_t all kernels are issued on */
/* CUDA-event timing of each kernel family on that stream: enable, run, then read. */
enum { GSGP_K_INTERP = 0, GSGP_K_GSOP = 1, GSGP_K_FINALIZE = 2, GSGP_K_PLAN = 3, GSGP_K_COUNT = 4 };
/* on: 0 off; 1 the generation kernels (GSGP_K_INTERP, GSGP_K_GSOP) only -- two timing events per launch cost
 * ~2 us of stream time each, so the small kernels are left alone; 2 every kernel */
int gsgp_profile_enable(gsgp_ctx *ctx, int32_t on);
int gsgp_profile_read(gsgp_ctx *ctx, int32_t kernel, int64_t *launches, double *total_ms,
                      double *algorithmic_bytes);
int64_t gsgp_launch_count(const gsgp_ctx *ctx);          /* kernels launched since create */

#ifdef __cplusplus
}
#endif
#endif
