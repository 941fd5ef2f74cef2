/* TEST STUB -- not product code, never shipped, never a fallback.
 * An LD_PRELOAD shim that stands in for the device half of libgsgp_b200.so (the gsgp_* entry points the example host
 * calls) so that examples/gsgp_b200_main.cpp's HOST half -- Go's random stream, the pipelined planner, contributions,
 * the asynchronous writers -- can be run end to end on a box without a GPU (tests/test_example_host.py).  Fitness and
 * semantics are a fixed integer hash of (generation, individual); nothing here computes GSGP. */
#include <stdint.h>
#include <stdlib.h>
#include "../include/gsgp_b200.h"

typedef struct { int32_t P, n, gen; uint64_t hash; } stub_ctx;

static void stub_mix(stub_ctx *c, const void *p, size_t n)           /* FNV-1a over what a generation call received */
{
    const unsigned char *b = p;
    for (size_t i = 0; i < n; ++i) { c->hash ^= b[i]; c->hash *= 1099511628211ull; }
}
uint64_t stub_plan_hash(const gsgp_ctx *ctx) { return ((const stub_ctx *)ctx)->hash; }

static double stub_fit(int32_t gen, int32_t k, int32_t which)      /* mirrored in the test */
{
    return (double)(((uint32_t)k * 2654435761u + (uint32_t)gen * 40503u + (uint32_t)which * 97u) % 1000u) / 8.0 + 1.0;
}
static void stub_fill(stub_ctx *c, double *fit, double *fit_test, int32_t *best)
{
    int32_t b = 0;
    for (int32_t k = 0; k < c->P; ++k) {
        const double f = stub_fit(c->gen, k, 0);
        if (fit) fit[k] = f;
        if (fit_test) fit_test[k] = stub_fit(c->gen, k, 1);
        if (f < stub_fit(c->gen, b, 0)) b = k;                      /* best_individual: first strictly better */
    }
    if (best) *best = b;
}

void gsgp_config_defaults(gsgp_config *cfg) { cfg->flags = 0; cfg->device = 0; cfg->rank = 0; cfg->world_size = 1; }
const char *gsgp_last_error(const gsgp_ctx *ctx) { (void)ctx; return "stub"; }
int gsgp_create(const gsgp_config *cfg, gsgp_ctx **out)
{
    stub_ctx *c = calloc(1, sizeof *c);
    c->P = cfg->population_size; c->n = cfg->nrow + cfg->nrow_test; c->hash = 14695981039346656037ull;
    *out = (gsgp_ctx *)c;
    return 0;
}
void gsgp_destroy(gsgp_ctx *ctx) { free(ctx); }
int gsgp_set_dataset(gsgp_ctx *ctx, const double *rowmajor) { (void)ctx; return rowmajor ? 0 : 1; }
int gsgp_set_constants(gsgp_ctx *ctx, const double *sym_val, int32_t n) { (void)ctx; (void)n; return sym_val ? 0 : 1; }
int gsgp_seed_semantic_files(gsgp_ctx *ctx, int32_t first, int32_t n, const char *const *paths) { (void)ctx; (void)first; (void)n; return paths ? 0 : 1; }
int gsgp_eval_initial(gsgp_ctx *ctx, int32_t first_ind, int32_t n, const int32_t *tree_pool, int32_t pool_len,
                      const int32_t *tree_off, const int32_t *tree_len)
{
    (void)ctx; (void)first_ind;
    for (int32_t i = 0; i < n; ++i)                                  /* the trees must lie inside the pool */
        if (tree_off[i] < 0 || tree_len[i] < 1 || tree_off[i] + tree_len[i] > pool_len || tree_pool[tree_off[i]] < 0) return 1;
    return 0;
}
int gsgp_fitness_all(gsgp_ctx *ctx, double *fit, double *fit_test, int32_t *index_best)
{
    stub_fill((stub_ctx *)ctx, fit, fit_test, index_best);
    return 0;
}
int gsgp_generation_compiled(gsgp_ctx *ctx, const gsgp_plan_entry *plan, const int32_t *tree_pool, int32_t pool_len,
                             const uint32_t *programs, int32_t n_instructions, const int32_t *prog_off,
                             const int32_t *prog_desc, double *fit_out, double *fit_test_out, int32_t *index_best_out)
{
    stub_ctx *c = (stub_ctx *)ctx;
    if (!plan || !tree_pool || pool_len < 0 || n_instructions < 0) return 1;
    stub_mix(c, plan, sizeof(gsgp_plan_entry) * (size_t)c->P);
    stub_mix(c, tree_pool, 4 * (size_t)pool_len);
    stub_mix(c, programs, 8 * (size_t)n_instructions);
    stub_mix(c, prog_off, 8 * (size_t)c->P);
    stub_mix(c, prog_desc, 8 * (size_t)c->P);
    c->gen++;
    stub_fill(c, fit_out, fit_test_out, index_best_out);
    return 0;
}
/* the library's own replay loop (gsgph_replay_generations) hands its planner's programs to this internal twin */
int gsgp_internal_generation_trusted(gsgp_ctx *ctx, const gsgp_plan_entry *plan, const int32_t *tree_pool, int32_t pool_len,
                                     const uint32_t *programs, int32_t n_instructions, const int32_t *prog_off,
                                     const int32_t *prog_desc, double *fit_out, double *fit_test_out, int32_t *index_best_out)
{
    return gsgp_generation_compiled(ctx, plan, tree_pool, pool_len, programs, n_instructions, prog_off, prog_desc, fit_out, fit_test_out, index_best_out);
}
int gsgp_get_semantic(gsgp_ctx *ctx, int32_t ind, double *out)
{
    stub_ctx *c = (stub_ctx *)ctx;
    for (int32_t i = 0; i < c->n; ++i) out[i] = (double)c->gen + (double)ind * 0.5 + (double)i * 0.25;
    return 0;
}
