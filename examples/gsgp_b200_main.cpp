// gsgp_b200_main.cpp -- a compiled host for the C-ABI: the L2 loop of main_cpu.go (:1364-1507) restated in C++ over
// include/gsgp_b200.h + include/gsgp_host.h, i.e. what main_b200.go does from Go.  It is an ABI
// example and link test, not the reference's CLI: a handful of positional options, the reference's file formats in
// (read_input_data :381-421, read_sem :710-733) and out (best-of-generation fitness / semantic lines :1404-1413,
// :1494-1505).  The RNG is the library's restatement of math/rand (gsgp_host.h): Go's own stream for the seed, so the
// draws are the ones main_cpu.go makes with the same -seed.
//
//   gsgp_b200_main <train> <test> <out_dir> [generations=10] [population=200] [seed=1] [linsc=0] [seed files...]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/gsgp_b200.h"
#include "../include/gsgp_host.h"

static void die(gsgp_ctx *c, const char *what)
{
    fprintf(stderr, "panic: %s: %s\n", what, gsgp_last_error(c));   // the reference panics on every failure
    exit(2);
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s train test out_dir [gens] [pop] [seed] [linsc] [semantic files...]\n", argv[0]); return 1; }
    const std::string out_dir = argv[3];
    const int gens = argc > 4 ? atoi(argv[4]) : 10, pop = argc > 5 ? atoi(argv[5]) : 200;
    const long seed = argc > 6 ? atol(argv[6]) : 1;
    const int linsc = argc > 7 ? atoi(argv[7]) : 0;
    const int n_seeds = argc > 8 ? argc - 8 : 0;

    int32_t nvar = 0, nrow = 0, nrow_test = 0;
    double *set = nullptr;
    if (gsgph_read_dataset(argv[1], argv[2], &nvar, &nrow, &nrow_test, &set)) { fprintf(stderr, "panic: cannot read the dataset\n"); return 2; }
    const int64_t N = (int64_t)nrow + nrow_test;

    gsgph_params p{};
    p.population_size = pop; p.nvar = nvar; p.num_constants = 0; p.max_depth_creation = 6; p.tournament_size = 4;
    p.zero_depth = 0; p.minimization_problem = 1; p.init_type = 2; p.p_crossover = 0.6; p.p_mutation = 0.3;
    p.min_random_constant = -100; p.max_random_constant = 100;                  // defaults main_cpu.go:157-169
    gsgph_rng *rng = gsgph_rng_new(seed);                                       // rand.Seed :1367

    gsgp_config cfg;
    gsgp_config_defaults(&cfg);
    cfg.population_size = pop; cfg.nvar = nvar; cfg.nrow = nrow; cfg.nrow_test = nrow_test;
    cfg.error_measure = GSGP_RMSE;                                              // as the harness passes, runner.py:335
    cfg.rng_mode = GSGP_RNG_HOST_REPLAY;
    if (linsc) cfg.flags |= GSGP_FLAG_LINEAR_SCALING;
    gsgp_ctx *ctx = nullptr;
    if (gsgp_create(&cfg, &ctx)) die(nullptr, "gsgp_create");                   // init_tables :1380
    if (gsgp_set_dataset(ctx, set)) die(ctx, "gsgp_set_dataset");
    std::vector<double> sym((size_t)4 + nvar);
    gsgph_create_T_F(rng, &p, sym.data());                                      // :1371
    if (gsgp_set_constants(ctx, sym.data(), (int32_t)sym.size())) die(ctx, "gsgp_set_constants");
    const auto start = std::chrono::steady_clock::now();                        // start = time.Now() :1374-1375
    if (n_seeds && gsgp_seed_semantic_files(ctx, 0, n_seeds, argv + 8)) die(ctx, "gsgp_seed_semantic_files");   // :1383-1388

    // initialize_population (:1400) + evaluate (:1402)
    const int n_trees = pop - n_seeds;
    std::vector<int32_t> pool((size_t)n_trees * 260 + 16), off((size_t)n_trees + 1), len((size_t)n_trees + 1);
    int32_t used = 0;
    if (gsgph_initialize_population(rng, &p, n_seeds, pool.data(), (int32_t)pool.size(), off.data(), len.data(), &used)) { fprintf(stderr, "panic: tree pool\n"); return 2; }
    if (gsgp_eval_initial(ctx, n_seeds, n_trees, pool.data(), used, off.data(), len.data())) die(ctx, "gsgp_eval_initial");
    std::vector<double> fit((size_t)pop), fit_test((size_t)pop), sem((size_t)N);
    int32_t index_best = 0;
    if (gsgp_fitness_all(ctx, fit.data(), fit_test.data(), &index_best)) die(ctx, "gsgp_fitness_all");

    // the reference's writers (:1227-1241), asynchronous: the rows are copied, text is produced off the loop
    gsgph_writer *f_train = gsgph_writer_open((out_dir + "/fitnesstrain.txt").c_str(), 0), *f_test = gsgph_writer_open((out_dir + "/fitnesstest.txt").c_str(), 0);
    gsgph_writer *s_train = gsgph_writer_open((out_dir + "/semantic_train.txt").c_str(), 0), *s_test = gsgph_writer_open((out_dir + "/semantic_test.txt").c_str(), 0);
    if (!f_train || !f_test || !s_train || !s_test) { fprintf(stderr, "panic: cannot create the output files\n"); return 2; }
    // contributions (:1390-1397,1411,1503): which seeded model each individual descends from, updated from every plan
    gsgph_contrib *contrib = gsgph_contrib_new(pop, n_seeds);
    FILE *f_contrib = fopen((out_dir + "/contributions.txt").c_str(), "w");
    if (!contrib || !f_contrib) { fprintf(stderr, "panic: cannot create the contributions file\n"); return 2; }
    FILE *f_time = fopen((out_dir + "/executiontime.txt").c_str(), "w");        // :1413,1505: time.Since(start) per line
    if (!f_time) { fprintf(stderr, "panic: cannot create the execution time file\n"); return 2; }
    std::vector<char> line(64);
    auto write_best = [&] {                                                     // :1404-1413, :1494-1505
        int64_t need = 0;
        if (gsgph_contrib_format(contrib, index_best, line.data(), (int64_t)line.size(), &need) == 3) {
            line.resize((size_t)need);
            gsgph_contrib_format(contrib, index_best, line.data(), (int64_t)line.size(), &need);
        }
        fwrite(line.data(), 1, (size_t)need, f_contrib); fputc('\n', f_contrib);
        gsgph_writer_float(f_train, fit[(size_t)index_best]);
        gsgph_writer_float(f_test, fit_test[(size_t)index_best]);
        if (gsgp_get_semantic(ctx, index_best, sem.data())) die(ctx, "gsgp_get_semantic");
        gsgph_writer_semantic(s_train, sem.data(), nrow);
        gsgph_writer_semantic(s_test, sem.data() + nrow, nrow_test);
        char dur[40];
        gsgph_format_duration(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - start).count(), dur, sizeof dur);
        fprintf(f_time, "%s\n", dur);
    };
    write_best();

    // main GP cycle :1440-1506.  The plan builder is pipelined: while the device computes generation g the planner's
    // worker threads draw generation g+1's operators / trees / tournament candidates and compile the trees; when the
    // fitness is back it picks the winners (and repairs a wrong elite guess) -- same plans as gsgph_build_plan.
    gsgph_planner *planner = gsgph_planner_new(rng, &p);
    if (!planner || gsgph_planner_enable_programs(planner, 0)) { fprintf(stderr, "panic: planner\n"); return 2; }
    for (int g = 0; g < gens; ++g) {
        const gsgp_plan_entry *plan; const int32_t *trees, *prog_off, *prog_desc; const uint32_t *programs;
        int32_t tlen = 0, n_ins = 0;
        if (gsgph_planner_plan(planner, fit.data(), index_best, &plan, &trees, &tlen)) { fprintf(stderr, "panic: tree pool\n"); return 2; }
        gsgph_planner_prefetch(planner, index_best);
        gsgph_planner_programs(planner, &programs, &n_ins, &prog_off, &prog_desc);
        if (gsgp_generation_compiled(ctx, plan, trees, tlen, programs, n_ins, prog_off, prog_desc,
                                     fit.data(), fit_test.data(), &index_best)) die(ctx, "gsgp_generation_compiled");
        if (gsgph_contrib_apply(contrib, plan, 0)) { fprintf(stderr, "panic: contributions\n"); return 2; }   // :857,885,909,1196
        write_best();
    }
    gsgph_planner_free(planner);
    if (gsgph_writer_close(f_train) | gsgph_writer_close(f_test) | gsgph_writer_close(s_train) | gsgph_writer_close(s_test)) {
        fprintf(stderr, "panic: cannot write the output files\n"); return 2;
    }
    fclose(f_time);
    if (fclose(f_contrib) != 0) { fprintf(stderr, "panic: cannot write the contributions file\n"); return 2; }
    gsgph_contrib_free(contrib);
    gsgp_destroy(ctx);
    gsgph_rng_free(rng);
    gsgph_free(set);
    return 0;
}
