// trees.hpp -- host-side validation + re-encoding of the reference's tree arrays.
//
// The reference passes random trees around as pointer-prefix int32 arrays
// (create_grow_tree_arrays main_cpu.go:609-639, tree_to_array :675-693):
//   function node = [op, idx_child1, idx_child2, <child1...>, <child2...>] (absolute indices),
//   terminal      = [id].
// eval_arrays (:755-775) walks that form recursively per fitness case.  The device interpreter
// instead runs a postfix program (left subtree, right subtree, op) with an explicit operand
// stack, so the shim converts once per tree, checking every index on the way: a malformed array
// is reported as an error instead of faulting on the device.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gsgp {

constexpr int kNumFunc = 4;   // + - * /  (main_cpu.go:433-440)

// Appends the postfix program of `tree[0..len)` to `out`; returns false and sets err on a
// malformed tree.  max_sp = operand-stack depth the program needs.
inline bool prefix_to_postfix(const int32_t *tree, int32_t len, int32_t n_symbols,
                              std::vector<uint16_t> &out, int32_t &max_sp, std::string &err)
{
    struct Frame { int32_t pos; int32_t next; };
    std::vector<Frame> stack;
    max_sp = 0;
    int32_t sp = 0, visited = 0;
    if (len <= 0) { err = "empty tree"; return false; }
    auto visit = [&](int32_t pos) -> bool {
        if (pos < 0 || pos >= len) { err = "tree index out of range"; return false; }
        if (++visited > len) { err = "tree has a cycle"; return false; }
        const int32_t id = tree[pos];
        if (id < 0 || id >= n_symbols || id > 0xFFFF) { err = "unknown symbol id in tree"; return false; }
        if (id >= kNumFunc) {
            out.push_back((uint16_t)id);
            if (++sp > max_sp) max_sp = sp;
            return true;
        }
        if (pos + 2 >= len) { err = "truncated function node"; return false; }
        stack.push_back({pos, 1});
        return true;
    };
    if (!visit(0)) return false;
    while (!stack.empty()) {
        Frame &f = stack.back();
        if (f.next > 2) {
            out.push_back((uint16_t)tree[f.pos]);
            --sp;
            stack.pop_back();
            continue;
        }
        const int32_t child = tree[f.pos + f.next];
        if (child <= f.pos) { err = "child index does not point forward"; return false; }
        ++f.next;
        if (!visit(child)) return false;
    }
    if (sp != 1) { err = "malformed tree"; return false; }
    return true;
}

}  // namespace gsgp
