// main_b200_glue.go -- the bodies go-gsgp's B200 build replaces, bound to libgsgp_b200.so through cgo.
//
// This is NOT a stand-alone Go file: go-gsgp keeps its whole program in one `package main` file per backend
// (main_cpu.go, main_gpu.go -- the legacy CUDA backend repeats ~850 lines of main_cpu.go), and this repository may not
// carry the reference's sources.  integration/make_main_b200.py assembles the real file on a checkout of go-gsgp:
//
//     python integration/make_main_b200.py /path/to/go-gsgp        # writes /path/to/go-gsgp/main_b200.go
//     cd /path/to/go-gsgp && go build -tags b200 -o gsgp-b200 main_b200.go
//
// It copies main_cpu.go, swaps the build tag for `b200`, adds the cgo preamble, removes the six functions that
// main_gpu.go also re-implements (main_gpu.go:688-853,897-914) -- init_tables, evaluate, reproduction,
// geometric_semantic_crossover, geometric_semantic_mutation, update_tables -- and appends everything below the marker
// line.  Flags, configuration.ini, the dataset / semantic readers, create_T_F, the tree builders, every rand.* call,
// tournament_selection, the contributions, best_individual and all writers stay main_cpu.go's own code.
//
// Written against include/gsgp_b200.h (ABI version 1).  There is no Go toolchain in the build image of this
// repository: the file has never been compiled.  The same call sequence runs in the tests through the ctypes mirror
// (go_gsgp_b200/engine.py) and through examples/gsgp_b200_main.cpp.
//
// What changes for the program: the float arithmetic between the random draws is deferred to the end of the
// generation (one gsgp_generation call in update_tables), which is legal because every child depends only on the OLD
// generation (sem_*_cases, fit and index_best are frozen during the k loop, main_cpu.go:1447-1484).  Every rand.*
// call stays where main_cpu.go makes it, so a run consumes the same random stream.
// With -proto_dump the per-individual record main() builds inside the k loop (main_cpu.go:1472-1483) reads fit_new[k]
// and sem_*_cases_new[k] before the generation has been computed; make_main_b200.py moves that block behind
// update_tables() (it then reads fit[k] / sem_*_cases[k] of the new generation).

// ---- cgo preamble (make_main_b200.py puts it between `package main` and the import block) ----
// #cgo CFLAGS: -I${SRCDIR}/gsgp-b200/include
// #cgo LDFLAGS: -L${SRCDIR}/gsgp-b200/go_gsgp_b200 -lgsgp_b200 -Wl,-rpath,${SRCDIR}/gsgp-b200/go_gsgp_b200
// #include <stdlib.h>
// #include "gsgp_b200.h"
// import "C"

// ==== everything below this line is appended to the assembled file =================================================
// GSGP-B200-GLUE-BEGIN

var (
	b200_ctx   *C.gsgp_ctx
	b200_plan  []C.gsgp_plan_entry // one entry per individual, rebuilt every generation
	b200_trees []C.int32_t         // the generation's random trees, reference array format
	b200_keep  bool                // -proto_dump: every individual's semantic comes back to the host
)

// the reference panics on every failure (e.g. main_cpu.go:383-386,725,730)
func b200_must(rc C.int) {
	if rc != 0 {
		panic(C.GoString(C.gsgp_last_error(b200_ctx)))
	}
}

func b200_append_tree(t []cInt) (C.int32_t, C.int32_t) {
	off := len(b200_trees)
	for _, v := range t {
		b200_trees = append(b200_trees, C.int32_t(v))
	}
	return C.int32_t(off), C.int32_t(len(t))
}

// Allocate the host-side tables and create the device context.  The semantic tables of main_cpu.go:1257-1284
// (4 x population x cases float64) live on the device; the host keeps the slice headers (rows are filled on demand:
// the best individual's row every generation, every row with -proto_dump) and the fitness / contribution tables.
func init_tables(num_models int) {
	if num_models < 1 {
		panic("Cannot use less than one model")
	}
	P := *config.population_size
	fit = make([]cFloat64, P)
	fit_test = make([]cFloat64, P)
	fit_new = make([]cFloat64, P)
	fit_test_new = make([]cFloat64, P)
	sem_train_cases = make([]Semantic, P)
	sem_train_cases_new = make([]Semantic, P)
	sem_test_cases = make([]Semantic, P)
	sem_test_cases_new = make([]Semantic, P)
	contrib = make([]Contribution, P)
	contrib_new = make([]Contribution, P)
	for i := 0; i < P; i++ {
		contrib[i] = NewContribution(num_models)
		contrib_new[i] = NewContribution(num_models)
	}
	b200_plan = make([]C.gsgp_plan_entry, P)
	b200_keep = *config.proto_dump != ""

	var cfg C.gsgp_config
	C.gsgp_config_defaults(&cfg)
	cfg.population_size = C.int32_t(P)
	cfg.nvar = C.int32_t(nvar)
	cfg.nrow = C.int32_t(nrow)
	cfg.nrow_test = C.int32_t(nrow_test)
	cfg.max_depth_creation = C.int32_t(*config.max_depth_creation)
	cfg.tournament_size = C.int32_t(*config.tournament_size)
	if *config.zero_depth {
		cfg.zero_depth = 1
	}
	if *config.minimization_problem {
		cfg.minimization_problem = 1
	} else {
		cfg.minimization_problem = 0
	}
	switch strings.ToUpper(*config.error_measure) { // main_cpu.go:1319-1331
	case "MSE":
		cfg.error_measure = C.GSGP_MSE
	case "RMSE":
		cfg.error_measure = C.GSGP_RMSE
	case "MAE":
		cfg.error_measure = C.GSGP_MAE
	case "MRE":
		cfg.error_measure = C.GSGP_MRE
	}
	cfg.num_constants = C.int32_t(NUM_CONSTANT_SYMBOLS)
	cfg.rng_mode = C.GSGP_RNG_HOST_REPLAY // Go's math/rand makes every draw
	if *config.use_linear_scaling {
		cfg.flags |= C.GSGP_FLAG_LINEAR_SCALING
	}
	if b200_keep {
		cfg.flags |= C.GSGP_FLAG_KEEP_TREE_SEMANTICS
	}
	if rc := C.gsgp_create(&cfg, &b200_ctx); rc != 0 {
		panic(C.GoString(C.gsgp_last_error(nil)))
	}

	// the dataset, row-major with the target last (as main_gpu.go:1036-1046 flattens `set`), train rows first
	flat := make([]C.double, (nrow+nrow_test)*(nvar+1))
	for i := 0; i < nrow+nrow_test; i++ {
		for j := 0; j < nvar; j++ {
			flat[i*(nvar+1)+j] = C.double(set[i].vars[j])
		}
		flat[i*(nvar+1)+nvar] = C.double(set[i].y_value)
	}
	b200_must(C.gsgp_set_dataset(b200_ctx, &flat[0]))
	// Symbol.value per id (create_T_F has run: main_cpu.go:1371 precedes init_tables :1380)
	vals := make([]C.double, len(symbols))
	for i, s := range symbols {
		vals[i] = C.double(s.value)
	}
	b200_must(C.gsgp_set_constants(b200_ctx, &vals[0], C.int32_t(len(vals))))
}

// evaluate(): main_cpu.go:780-795.  Seeded individuals (p.individuals[i] == nil) already have their semantic in
// sem_train_cases[i] / sem_test_cases[i] (main() sliced it out of read_sem's vector, :1383-1388): upload it; the rest
// are evaluated on the device from their tree arrays.  Fitness and the best individual come back in one call.
func evaluate(p *Population) {
	P := *config.population_size
	var pool []C.int32_t
	var off, length []C.int32_t
	first := -1
	for i := 0; i < P; i++ {
		if p.individuals[i] == nil {
			sem := make([]C.double, nrow+nrow_test)
			for j := 0; j < nrow; j++ {
				sem[j] = C.double(sem_train_cases[i][j])
			}
			for j := 0; j < nrow_test; j++ {
				sem[nrow+j] = C.double(sem_test_cases[i][j])
			}
			b200_must(C.gsgp_seed_semantic(b200_ctx, C.int32_t(i), &sem[0]))
			sem_train_cases[i], sem_test_cases[i] = nil, nil
			continue
		}
		if first < 0 {
			first = i // seeds occupy individuals 0..S-1 (:1383), the trees are contiguous after them
		}
		arr := tree_to_array(p.individuals[i])
		off = append(off, C.int32_t(len(pool)))
		length = append(length, C.int32_t(len(arr)))
		for _, v := range arr {
			pool = append(pool, C.int32_t(v))
		}
	}
	if first >= 0 {
		b200_must(C.gsgp_eval_initial(b200_ctx, C.int32_t(first), C.int32_t(len(off)), &pool[0], C.int32_t(len(pool)), &off[0], &length[0]))
	}
	var best C.int32_t
	b200_must(C.gsgp_fitness_all(b200_ctx, (*C.double)(&fit[0]), (*C.double)(&fit_test[0]), &best))
	b200_fetch_rows(cInt(best))
}

// the rows the host reads after a generation: the best individual's (main_cpu.go:1408-1409,1499-1500), every
// individual's with -proto_dump (:1472-1483)
func b200_fetch_rows(best cInt) {
	buf := make([]C.double, nrow+nrow_test)
	fetch := func(k int) {
		b200_must(C.gsgp_get_semantic(b200_ctx, C.int32_t(k), &buf[0]))
		tr, te := make(Semantic, nrow), make(Semantic, nrow_test)
		for j := range tr {
			tr[j] = cFloat64(buf[j])
		}
		for j := range te {
			te[j] = cFloat64(buf[nrow+j])
		}
		sem_train_cases[k], sem_test_cases[k] = tr, te
	}
	if b200_keep {
		for k := 0; k < *config.population_size; k++ {
			fetch(k)
		}
	} else {
		fetch(int(best))
	}
}

// reproduction(): main_cpu.go:844-861 -- the draws and the contribution copy stay, the row and fitness copies
// become a plan entry (the device copies the fitness, it does not recompute it: :859-860)
func reproduction(i cInt) {
	old_i := i
	e := &b200_plan[old_i]
	e.op, e.elite = C.GSGP_OP_REPR, 1
	if i != index_best {
		i = tournament_selection()
		e.elite = 0
	}
	copy_contrib(contrib_new[old_i], contrib[i])
	e.p1, e.p2 = C.int32_t(i), -1
	e.tree0_off, e.tree0_len, e.tree1_off, e.tree1_len, e.ms = -1, 0, -1, 0, 0
}

// geometric_semantic_crossover(): main_cpu.go:876-914
func geometric_semantic_crossover(i cInt) {
	e := &b200_plan[i]
	e.op, e.elite, e.p1, e.p2 = C.GSGP_OP_XO, 1, C.int32_t(i), -1
	e.tree0_off, e.tree0_len, e.tree1_off, e.tree1_len, e.ms = -1, 0, -1, 0, 0
	if i != index_best {
		rt1 = create_grow_tree_arrays(0, cInt(*config.max_depth_creation), 0)
		p1 := tournament_selection()
		p2 := tournament_selection()
		merge_contribs(contrib_new[i], contrib[p1], contrib[p2])
		e.elite, e.p1, e.p2 = 0, C.int32_t(p1), C.int32_t(p2)
		e.tree0_off, e.tree0_len = b200_append_tree(rt1)
	} else {
		copy_contrib(contrib_new[i], contrib[i])
	}
}

// geometric_semantic_mutation(): main_cpu.go:917-947 (called after reproduction(i), which has filled the entry)
func geometric_semantic_mutation(i cInt) {
	if i != index_best {
		mut_step := cFloat64(rand.Float64())
		rt1 = create_grow_tree_arrays(0, cInt(*config.max_depth_creation), 0)
		rt2 = create_grow_tree_arrays(0, cInt(*config.max_depth_creation), 0)
		e := &b200_plan[i]
		e.op, e.ms = C.GSGP_OP_MUT, C.double(mut_step)
		e.tree0_off, e.tree0_len = b200_append_tree(rt1)
		e.tree1_off, e.tree1_len = b200_append_tree(rt2)
	} else {
		b200_plan[i].op = C.GSGP_OP_MUT // the elite is reproduced, not mutated: a self copy either way
	}
}

// update_tables(): main_cpu.go:1191-1197 -- the whole generation runs here, in one call; fit / fit_test receive the new
// generation's values directly (no fit_new pair to swap), the contributions swap as before
func update_tables() {
	var best C.int32_t
	var pool *C.int32_t
	if len(b200_trees) > 0 {
		pool = &b200_trees[0]
	}
	b200_must(C.gsgp_generation(b200_ctx, &b200_plan[0], pool, C.int32_t(len(b200_trees)),
		(*C.double)(&fit[0]), (*C.double)(&fit_test[0]), &best))
	b200_trees = b200_trees[:0]
	contrib, contrib_new = contrib_new, contrib
	for k := range sem_train_cases { // last generation's rows are stale
		sem_train_cases[k], sem_test_cases[k] = nil, nil
	}
	b200_fetch_rows(cInt(best)) // main() prints sem_*_cases[index_best]; best_individual() (:1180-1188) picks the same index
}

// the random trees' semantics of individual k of the generation just computed (-proto_dump, main_cpu.go:1459-1467):
// which = 0 -> rt1, 1 -> rt2
func b200_tree_semantic(k int, which int) (Semantic, Semantic) {
	buf := make([]C.double, nrow+nrow_test)
	b200_must(C.gsgp_get_tree_semantic(b200_ctx, C.int32_t(k), C.int32_t(which), &buf[0]))
	tr, te := make(Semantic, nrow), make(Semantic, nrow_test)
	for j := range tr {
		tr[j] = cFloat64(buf[j])
	}
	for j := range te {
		te[j] = cFloat64(buf[nrow+j])
	}
	return tr, te
}
// GSGP-B200-GLUE-END
