/*
 * oracle_cli.c -- command-line driver of the CPU ORACLE (test infrastructure; see gsgp_oracle.h).
 *
 * Restates main() of main_cpu.go:1286-1528 for the pieces the hot path touches: dataset files in,
 * best-of-generation fitness / semantic files out (Go %v formatting), positional semantic-seed
 * files.  Used (a) for CLI-level parity tests against the CUDA host driver and (b) as the timed
 * "C restatement of main_cpu.go" CPU baseline (bench.py cpu_baseline / --impl reference).
 * It is NOT go-gsgp and is never labelled as such.
 *
 * usage: oracle_cli -train_file F -test_file F [-population_size N] [-max_number_generations N]
 *          [-p_crossover X] [-p_mutation X] [-max_depth_creation N] [-tournament_size N]
 *          [-init_type N] [-num_random_constants N] [-error_measure MSE|RMSE|MAE|MRE]
 *          [-seed N] [-workers N] [-rng alfg|philox] [-out_file_train_fitness F] ... [seed files]
 */
#define _GNU_SOURCE
#include "gsgp_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>

static double now_s(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static FILE *create_or_die(const char *path)       /* create_or_panic main_cpu.go:1227-1241 (no gzip here) */
{
    if (!path || !*path) return NULL;
    FILE *f = fopen(path, "w");
    if (!f) { perror(path); exit(2); }
    return f;
}

static void write_float_line(FILE *f, double v)
{
    if (!f) return;
    char b[64]; orc_format_float(v, b, sizeof b); fprintf(f, "%s\n", b);
}

static void write_sem_line(FILE *f, const double *v, int n)   /* Semantic.String :123-126 */
{
    if (!f) return;
    char b[64];
    for (int i = 0; i < n; ++i) { orc_format_float(v[i], b, sizeof b); fprintf(f, i ? ",%s" : "%s", b); }
    fputc('\n', f);
}

int main(int argc, char **argv)
{
    orc_config cfg; orc_config_defaults(&cfg);
    const char *train = NULL, *test = NULL;
    const char *of_train = "fitnesstrain.txt", *of_test = "fitnesstest.txt";
    const char *of_sem_train = "", *of_sem_test = "", *of_timing = "execution_time.txt";
    int gens = 300, json = 0;
    const char *seeds[4096]; int n_seeds = 0;
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        if (a[0] != '-') { if (n_seeds < 4096) seeds[n_seeds++] = a; continue; }
        while (*a == '-') ++a;
        const char *v = (i + 1 < argc) ? argv[i + 1] : "";
        if (!strcmp(a, "json")) { json = 1; continue; }
        if (!strcmp(a, "zero_depth")) { cfg.zero_depth = 1; continue; }
        ++i;
        if (!strcmp(a, "train_file")) train = v;
        else if (!strcmp(a, "test_file")) test = v;
        else if (!strcmp(a, "population_size")) cfg.population_size = atoi(v);
        else if (!strcmp(a, "max_number_generations")) gens = atoi(v);
        else if (!strcmp(a, "init_type")) cfg.init_type = atoi(v);
        else if (!strcmp(a, "p_crossover")) cfg.p_crossover = atof(v);
        else if (!strcmp(a, "p_mutation")) cfg.p_mutation = atof(v);
        else if (!strcmp(a, "max_depth_creation")) cfg.max_depth_creation = atoi(v);
        else if (!strcmp(a, "tournament_size")) cfg.tournament_size = atoi(v);
        else if (!strcmp(a, "num_random_constants")) cfg.num_random_constants = atoi(v);
        else if (!strcmp(a, "min_random_constant")) cfg.min_random_constant = atof(v);
        else if (!strcmp(a, "max_random_constant")) cfg.max_random_constant = atof(v);
        else if (!strcmp(a, "minimization_problem")) cfg.minimization_problem = atoi(v);
        else if (!strcmp(a, "seed")) cfg.rng_seed = atoll(v);
        else if (!strcmp(a, "workers")) cfg.n_workers = atoi(v);
        else if (!strcmp(a, "rng")) cfg.rng_kind = !strcasecmp(v, "philox") ? ORC_RNG_PHILOX : ORC_RNG_ALFG;
        else if (!strcmp(a, "error_measure")) {
            cfg.error_measure = !strcasecmp(v, "RMSE") ? ORC_RMSE : !strcasecmp(v, "MAE") ? ORC_MAE
                              : !strcasecmp(v, "MRE") ? ORC_MRE : ORC_MSE;
        }
        else if (!strcmp(a, "out_file_train_fitness")) of_train = v;
        else if (!strcmp(a, "out_file_test_fitness")) of_test = v;
        else if (!strcmp(a, "out_file_train_semantic")) of_sem_train = v;
        else if (!strcmp(a, "out_file_test_semantic")) of_sem_test = v;
        else if (!strcmp(a, "out_file_exec_timing")) of_timing = v;
        else { fprintf(stderr, "unknown flag -%s\n", a); return 2; }
    }
    if (!train || !test) { fprintf(stderr, "Please specify -train_file and -test_file\n"); return 2; }

    int32_t nvar, nrow, nrow_test; double *m;
    if (orc_read_dataset(train, test, &nvar, &nrow, &nrow_test, &m)) { fprintf(stderr, "cannot read datasets\n"); return 2; }
    FILE *f_tr = create_or_die(of_train), *f_te = create_or_die(of_test);
    FILE *f_str = create_or_die(of_sem_train), *f_ste = create_or_die(of_sem_test);
    FILE *f_tm = create_or_die(of_timing);

    orc_state *s = orc_new(&cfg, nvar, nrow, nrow_test, m);
    orc_create_T_F(s);                                               /* :1371 */
    double t0 = now_s();                                             /* :1375 */
    double *sem = (double *)malloc(sizeof(double) * (size_t)(nrow + nrow_test));
    for (int i = 0; i < n_seeds; ++i) {                              /* :1383-1388 */
        if (orc_read_sem(seeds[i], nrow + nrow_test, sem)) { fprintf(stderr, "Not enough values when reading semantic file\n"); return 2; }
        orc_seed_semantic(s, i, sem);
    }
    orc_initialize_population(s, n_seeds);                           /* :1400 */
    orc_evaluate(s);                                                 /* :1402-1403 */
    double t_init = now_s() - t0;
    int b = orc_index_best(s);
    write_float_line(f_tr, orc_fit(s)[b]); write_float_line(f_te, orc_fit_test(s)[b]);
    write_sem_line(f_str, orc_sem_train(s, b), nrow); write_sem_line(f_ste, orc_sem_test(s, b), nrow_test);
    if (f_tm) fprintf(f_tm, "%.9gs\n", now_s() - t0);
    double t1 = now_s();
    for (int g = 0; g < gens; ++g) {                                 /* :1440-1506 */
        orc_generation(s);
        b = orc_index_best(s);
        write_float_line(f_tr, orc_fit(s)[b]); write_float_line(f_te, orc_fit_test(s)[b]);
        write_sem_line(f_str, orc_sem_train(s, b), nrow); write_sem_line(f_ste, orc_sem_test(s, b), nrow_test);
        if (f_tm) fprintf(f_tm, "%.9gs\n", now_s() - t0);
    }
    double t_loop = now_s() - t1;
    if (json) {
        double updates = (double)cfg.population_size * (double)(nrow + nrow_test) * (double)gens;
        printf("{\"impl\": \"C restatement of main_cpu.go (oracle)\", \"workers\": %d, \"pop\": %d, \"cases\": %d, "
               "\"gens\": %d, \"init_s\": %.6f, \"loop_s\": %.6f, \"updates_per_s\": %.6e}\n",
               cfg.n_workers, cfg.population_size, nrow + nrow_test, gens, t_init, t_loop,
               t_loop > 0 ? updates / t_loop : 0.0);
    }
    if (f_tr) fclose(f_tr);
    if (f_te) fclose(f_te);
    if (f_str) fclose(f_str);
    if (f_ste) fclose(f_ste);
    if (f_tm) fclose(f_tm);
    free(sem); free(m); orc_free(s);
    return 0;
}
