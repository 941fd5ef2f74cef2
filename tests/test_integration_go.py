"""integration/make_main_b200.py: assembles main_b200.go on a checkout of go-gsgp.  There is no Go toolchain here, so
the test checks the assembly itself -- and only where the reference tree exists (not on the GPU box)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_glue_file_declares_every_replaced_function():
    glue = open(os.path.join(ROOT, "integration", "main_b200_glue.go")).read()
    for f in ("init_tables", "evaluate", "reproduction", "geometric_semantic_crossover", "geometric_semantic_mutation", "update_tables"):
        assert re.search(r"^func %s\(" % f, glue, re.M), f
    # every C symbol the glue calls is declared by the header
    header = open(os.path.join(ROOT, "include", "gsgp_b200.h")).read()
    for sym in set(re.findall(r"C\.(gsgp_[a-z_0-9]+)\(", glue)):
        assert re.search(r"\b%s\(" % sym, header), sym
    for const in set(re.findall(r"C\.(GSGP_[A-Z_0-9]+)", glue)):
        assert const in header, const


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "main_cpu.go")), reason="needs the reference checkout")
def test_assembled_file_has_one_definition_of_everything(tmp_path):
    out = tmp_path / "main_b200.go"
    subprocess.run([sys.executable, os.path.join(ROOT, "integration", "make_main_b200.py"), REF, "--out", str(out)], check=True,
                   capture_output=True)
    src = out.read_text()
    assert "// +build b200" in src and 'import "C"' in src
    names = re.findall(r"^func (?:\([^)]*\) )?([A-Za-z_0-9]+)\(", src, re.M)
    dup = {n for n in names if names.count(n) > 1 and n not in ("String", "Close", "Write")}
    assert not dup, dup
    for f in ("init_tables", "evaluate", "reproduction", "geometric_semantic_crossover", "geometric_semantic_mutation",
              "update_tables", "tournament_selection", "create_grow_tree_arrays", "best_individual", "main"):
        assert names.count(f) == 1, f
    assert src.count("{") == src.count("}")
    # the host arithmetic of the replaced functions is gone: no semantic loop is left on the host
    assert "sem_train_cases_new[i][j]" not in src and "semantic_evaluate_array(rt1" not in src
    # the -proto_dump record was moved behind update_tables()
    assert src.index("ops_used[k], fit[k], fit_test[k]") > src.index("\t\tupdate_tables()\n")
