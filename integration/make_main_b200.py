#!/usr/bin/env python
"""Assembles go-gsgp's main_b200.go on a checkout of the reference (this repository carries none of its sources).

    python integration/make_main_b200.py /path/to/go-gsgp [--out main_b200.go] [--lib-dir gsgp-b200]

main_b200.go = main_cpu.go with
  * the build tag `gc,!gccgo` replaced by `b200` (and main_cpu.go's own tag should gain `,!b200`: printed as a hint),
  * the cgo preamble inserted after `package main` (include / library paths relative to --lib-dir inside the checkout),
  * the six functions main_gpu.go also re-implements removed (init_tables, evaluate, reproduction,
    geometric_semantic_crossover, geometric_semantic_mutation, update_tables),
  * the -proto_dump record of the k loop (main_cpu.go:1472-1483) moved behind update_tables(), where the generation's
    fitness and semantics exist,
  * integration/main_b200_glue.go's bodies appended.
Build: go build -tags b200 -o gsgp-b200 main_b200.go      (needs libgsgp_b200.so built: python go_gsgp_b200/build.py)
"""
import argparse
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPLACED = ("init_tables", "evaluate", "reproduction", "geometric_semantic_crossover", "geometric_semantic_mutation", "update_tables")


def remove_func(src: str, name: str) -> str:
    """drops `func name(...) {...}` together with the comment lines right above it"""
    m = re.search(r"^func %s\(" % re.escape(name), src, re.M)
    if not m:
        raise SystemExit(f"main_cpu.go has no func {name}")
    i = src.index("{", m.end())
    depth, j = 0, i
    while True:                                   # gofmt'd source: braces in strings / runes do not occur in these bodies
        ch = src[j]
        depth += ch == "{"
        depth -= ch == "}"
        j += 1
        if depth == 0:
            break
    start = m.start()
    while True:                                   # the doc comment above
        prev = src.rfind("\n", 0, start - 1)
        line = src[prev + 1:start - 1]
        if line.startswith("//"):
            start = prev + 1
        else:
            break
    return src[:start] + src[j:].lstrip("\n")


def move_proto_record(src: str) -> str:
    """the per-individual pb.Individual of the k loop reads fit_new[k] / sem_*_cases_new[k] / sem_rt*: rebuild it from the
    plan after update_tables().  Leaves the source alone (with a warning) when the anchors are not found."""
    a = src.find("\t\t\t// If required, save this individual\n")
    b = src.find("\t\tupdate_tables()\n")
    if a < 0 or b < 0 or b < a:
        sys.stderr.write("warning: -proto_dump block not found; the dump would read stale per-individual values\n")
        return src
    end = src.index("\t\t\t}\n", src.index("pb_pop.Individuals = append(pb_pop.Individuals, ind)", a)) + len("\t\t\t}\n")
    src = src[:a] + "\t\t\tops_used = append(ops_used, operator_used)\n" + src[end:]
    src = src.replace("\t\tfor k := 0; k < *config.population_size; k++ {\n\t\t\tvar operator_used pb.Individual_Operator",
                      "\t\tops_used := make([]pb.Individual_Operator, 0, *config.population_size)\n"
                      "\t\tfor k := 0; k < *config.population_size; k++ {\n\t\t\tvar operator_used pb.Individual_Operator", 1)
    b = src.find("\t\tupdate_tables()\n")
    after = ("\t\tupdate_tables()\n"
             "\t\tif *config.proto_dump != \"\" {\n"
             "\t\t\tfor k := 0; k < *config.population_size; k++ {\n"
             "\t\t\t\tvar random_trees []*pb.RandomTree\n"
             "\t\t\t\tif b200_plan[k].elite == 0 && ops_used[k] != pb.Individual_REPR {\n"
             "\t\t\t\t\ttr, te := b200_tree_semantic(k, 0)\n"
             "\t\t\t\t\trandom_trees = append(random_trees, &pb.RandomTree{b200_tree_of(k, 0), tr, te})\n"
             "\t\t\t\t\tif ops_used[k] == pb.Individual_MUT {\n"
             "\t\t\t\t\t\ttr, te = b200_tree_semantic(k, 1)\n"
             "\t\t\t\t\t\trandom_trees = append(random_trees, &pb.RandomTree{b200_tree_of(k, 1), tr, te})\n"
             "\t\t\t\t\t}\n"
             "\t\t\t\t}\n"
             "\t\t\t\tpb_pop.Individuals = append(pb_pop.Individuals, &pb.Individual{ops_used[k], fit[k], fit_test[k],\n"
             "\t\t\t\t\tsem_train_cases[k], sem_test_cases[k], contrib_new[k].AsBytes(), random_trees})  // the OLD generation's contribution of slot k, as main_cpu.go:1479 records it (contrib was swapped above)\n"
             "\t\t\t}\n"
             "\t\t}\n")
    src = src[:b] + after + src[b + len("\t\tupdate_tables()\n"):]
    # the now-unused random_trees of the k loop
    src = src.replace("\t\t\tvar random_trees []*pb.RandomTree        // Random trees used\n", "", 1)
    src = re.sub(r"\t\t\t\trandom_trees = \[\]\*pb\.RandomTree\{\n(?:\t\t\t\t\t\{rt\d, sem_rt\d_train, sem_rt\d_test\},\n)+\t\t\t\t\}\n", "", src)
    return src


TREE_OF = '''
// the random tree of individual k as the host drew it (kept by the library: GSGP_FLAG_KEEP_TREE_SEMANTICS runs)
func b200_tree_of(k int, which int) []cInt {
	buf := make([]C.int32_t, 4<<uint(*config.max_depth_creation))
	var n C.int32_t
	b200_must(C.gsgp_get_tree(b200_ctx, C.int32_t(k), C.int32_t(which), &buf[0], C.int32_t(len(buf)), &n))
	out := make([]cInt, n)
	for i := range out {
		out[i] = cInt(buf[i])
	}
	return out
}
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("checkout")
    ap.add_argument("--out", default="main_b200.go")
    ap.add_argument("--lib-dir", default="gsgp-b200", help="where this repository sits inside the checkout")
    a = ap.parse_args()
    src = open(os.path.join(a.checkout, "main_cpu.go")).read()
    if "// +build gc,!gccgo" not in src:
        raise SystemExit("main_cpu.go: build tag line not found")
    src = src.replace("// +build gc,!gccgo", "// +build b200", 1)
    preamble = ("package main\n\n"
                f"// #cgo CFLAGS: -I${{SRCDIR}}/{a.lib_dir}/include\n"
                f"// #cgo LDFLAGS: -L${{SRCDIR}}/{a.lib_dir}/go_gsgp_b200 -lgsgp_b200 -Wl,-rpath,${{SRCDIR}}/{a.lib_dir}/go_gsgp_b200\n"
                "// #include <stdlib.h>\n// #include \"gsgp_b200.h\"\nimport \"C\"\n")
    src = src.replace("package main\n", preamble, 1)
    for f in REPLACED:
        src = remove_func(src, f)
    src = move_proto_record(src)
    glue = open(os.path.join(HERE, "main_b200_glue.go")).read()
    body = glue[glue.index("// GSGP-B200-GLUE-BEGIN"):glue.index("// GSGP-B200-GLUE-END")]
    out = os.path.join(a.checkout, a.out) if not os.path.isabs(a.out) else a.out
    open(out, "w").write(src.rstrip("\n") + "\n\n" + body + TREE_OF)
    print(out)
    print("hint: give main_cpu.go the tag `// +build gc,!gccgo,!b200` so that `go build -tags b200` sees one main()")


if __name__ == "__main__":
    main()
